#!/bin/bash
# Evidence session for the final build: full GPU suite, bench, ncu --set full of one whole batch-64
# forward (raw CSV -> scripts/summarize_ncu_raw.py), launch list at batch 8.
O=gpurun_out/${1:-ev}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/pytest_gpu.log
timeout 500 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench.err
tail -24 $O/bench.err | head -14; cut -c1-400 $O/bench.json
timeout 1200 ncu --set full --clock-control none --profile-from-start off --csv --page raw \
    --log-file $O/ncu_b64_raw.csv python scripts/profile_step.py 64 > $O/ncu_b64.log 2>&1
tail -1 $O/ncu_b64.log; ls -la $O/ncu_b64_raw.csv
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize_small.py > $O/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a $O/memcheck.log; tail -4 $O/memcheck.log
timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
timeout 300 python scripts/segment_probe.py 50 2>&1 | tail -1 | tee -a $O/segments.jsonl
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_b8.csv python scripts/profile_step.py 8 > $O/ncu_launches_b8.log 2>&1
tail -1 $O/ncu_launches_b8.log
