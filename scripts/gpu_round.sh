#!/bin/bash
# One GPU session: tests, bench (both arms), ncu launch list + full captures.  Output -> gpurun_out/
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
python bench.py --steps 10 --warmup 3 --profile-ops > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
tail -22 gpurun_out/bench_ours.err; cat gpurun_out/bench_ours.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cat gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_b8.csv python scripts/profile_step.py 8 > gpurun_out/ncu_launches.log 2>&1
tail -1 gpurun_out/ncu_launches.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -c 2 \
    -o gpurun_out/prof_conv -f python scripts/profile_step.py 8 > gpurun_out/ncu_conv.log 2>&1
tail -1 gpurun_out/ncu_conv.log
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_kernel -s 45 -c 5 \
    -o gpurun_out/prof_xfmr -f python scripts/profile_step.py 8 > gpurun_out/ncu_xfmr.log 2>&1
tail -1 gpurun_out/ncu_xfmr.log
ls -la gpurun_out | head -20
