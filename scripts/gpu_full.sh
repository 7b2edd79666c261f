#!/bin/bash
# full GPU check of a build: kernel + model tests, bench line, launch list at batch 8, batch-1 latency
O=gpurun_out/${1:-full}
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --profile-ops --no-c5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
head -16 $O/bench.err; cut -c1-330 $O/bench.json
timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
timeout 300 python scripts/segment_probe.py 50 2>&1 | tail -1 | tee -a $O/segments.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_b8.csv python scripts/profile_step.py 8 > $O/ncu_launches_b8.log 2>&1
tail -1 $O/ncu_launches_b8.log
