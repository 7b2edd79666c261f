#!/bin/bash
O=gpurun_out/${1:-lat}
mkdir -p $O
for pdl in 0 1; do for ns in 0 1; do
  OPP_PDL=$pdl OPP_NSPLIT=$ns timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
done; done
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -4 | tee $O/pytest_gpu.log
timeout 500 python bench.py --steps 10 --warmup 3 --profile-ops --no-c5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
head -8 $O/bench.err; cut -c1-330 $O/bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_b1.csv python scripts/profile_step.py 1 > $O/ncu_launches_b1.log 2>&1
tail -1 $O/ncu_launches_b1.log
