#!/bin/bash
# ncu --set full captures of the kernels added in round 2 (run under gpurun from the repo root)
out=gpurun_out
NCU="ncu --clock-control none --set full --import-source on"
timeout 600 $NCU -k regex:"k_verify_each_comb|k_each_key" -c 3 -f -o $out/prof_each_comb_r2 python tools/each_bench.py > $out/prof_each_comb_r2.log 2>&1
timeout 600 $NCU -k regex:"k_batch_torsion|k_transcript_blocks" -c 3 -f -o $out/prof_batches_r2 python bench.py --workload verify --no-extras --steps 1 --warmup 1 > $out/prof_batches_r2.log 2>&1
ls -la $out/*.ncu-rep | tail -4
