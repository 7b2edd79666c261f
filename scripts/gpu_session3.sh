#!/bin/bash
# GPU session 3: full GPU suite, bench (+ reference arm), launch list at batch 8, source-level
# capture of the fused upsample lateral conv.
O=gpurun_out/${1:-s3}
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -25 | tee $O/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench.err
tail -24 $O/bench.err; cat $O/bench.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; cat $O/bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_b8.csv python scripts/profile_step.py 8 > $O/ncu_launches.log 2>&1
tail -1 $O/ncu_launches.log
timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_kernel -s 19 -c 1 -o $O/prof_convup_b8 -f python scripts/profile_step.py 8 > $O/ncu_convup.log 2>&1
tail -1 $O/ncu_convup.log
