tag=r2h
out=gpurun_out
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $out/launches_msm_$tag.csv python bench.py --workload msm --no-extras --steps 3 --warmup 1 > $out/launches_msm_$tag.log 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum -c 900 --csv --log-file $out/launches_verify_$tag.csv python bench.py --workload verify --no-extras --steps 2 --warmup 1 > $out/launches_verify_$tag.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_bucket_accumulate -s 2 -c 1 -f -o $out/prof_bucket_$tag python bench.py --workload msm --no-extras --steps 2 --warmup 1 > $out/prof_bucket_$tag.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:"k_batch_torsion|k_transcript_blocks|k_combine" -s 3 -c 4 -f -o $out/prof_new_$tag python bench.py --workload verify --no-extras --steps 1 --warmup 1 > $out/prof_new_$tag.log 2>&1
ls -la $out | tail -8
