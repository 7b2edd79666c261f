#!/bin/bash
O=gpurun_out/${1:-s6}
mkdir -p $O
timeout 600 python -m pytest tests/test_loftr_gpu.py tests/test_pnp_gpu.py -q -m gpu 2>&1 | tail -8 | tee $O/pytest_loftr.log
timeout 500 python bench.py --steps 10 --warmup 3 > $O/bench.json 2> $O/bench.err
tail -3 $O/bench.err; cat $O/bench.json
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke.log
