#!/bin/bash
# GPU session 4: full GPU suite (incl. LoFTR 2D-2D, full attention), bench, launch lists (batch 8 and 1).
O=gpurun_out/${1:-s4}
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -40 | tee $O/pytest_gpu.log
timeout 400 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench.err
tail -24 $O/bench.err; cat $O/bench.json
for b in 8 1; do
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $O/launches_b$b.csv python scripts/profile_step.py $b > $O/ncu_launches_b$b.log 2>&1
tail -1 $O/ncu_launches_b$b.log
done
