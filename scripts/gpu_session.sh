#!/bin/bash
# One GPU session: full GPU test suite, variant timing probe, bench (both arms), ncu launch list.
# Output -> gpurun_out/s1/
O=gpurun_out/${1:-s1}
mkdir -p $O gpurun_out/variants
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $O/pytest_gpu.log
timeout 400 python scripts/variant_probe.py probe "" colmax=0,lse_cols=0 lazy=1 kv1=1 fine_attn_vec=1 > $O/probe.log 2>&1
echo "probe exit=$?"; cp gpurun_out/variants/probe.json $O/probe.json
timeout 300 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench.err
tail -25 $O/bench.err; cat $O/bench.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_b8.csv python scripts/profile_step.py 8 > $O/ncu_launches.log 2>&1
tail -1 $O/ncu_launches.log
