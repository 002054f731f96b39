#!/bin/bash
# One profiling pass on the B200 (run under gpurun from the repo root): launch lists of one MSM and one verify_batch
# bench step, ncu --set full captures of the dominant kernels, the TMA A/B, the reference arm.  Usage:
#   tools/profile_round.sh <tag>          -> gpurun_out/*_<tag>.*
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
# launch lists (cold-cache, serialised: compare shares)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $out/launches_msm_$tag.csv \
    python bench.py --workload msm --no-extras --steps 3 --warmup 1 > $out/launches_msm_$tag.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $out/launches_verify_$tag.csv \
    python bench.py --workload verify --no-extras --steps 2 --warmup 1 > $out/launches_verify_$tag.log 2>&1
# full captures, one launch each (skip the warm-up launches)
timeout 600 $NCU --set full --import-source on -k regex:k_bucket_accumulate -s 2 -c 1 -f -o $out/prof_bucket_$tag \
    python bench.py --workload msm --no-extras --steps 2 --warmup 1 > $out/prof_bucket_$tag.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_prep_R -s 1 -c 1 -f -o $out/prof_prep_R_$tag \
    python bench.py --workload verify --no-extras --steps 1 --warmup 1 > $out/prof_prep_R_$tag.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_transcript -s 1 -c 1 -f -o $out/prof_transcript_$tag \
    python bench.py --workload verify --no-extras --steps 1 --warmup 1 > $out/prof_transcript_$tag.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"k_chunk_reduce|k_combine|k_finish_windows|k_plain_sum|k_heavy_fixup" -s 10 -c 10 -f -o $out/prof_tail_$tag \
    python bench.py --workload msm --no-extras --steps 2 --warmup 1 > $out/prof_tail_$tag.log 2>&1
# TMA A/B and the reference arm
timeout 600 python tools/ab_tma.py > $out/ab_tma_$tag.json 2> $out/ab_tma_$tag.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_$tag.json 2> $out/bench_reference_$tag.err
ls -la $out | tail -30
