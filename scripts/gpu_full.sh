#!/bin/bash
# full GPU check of a build: kernel + model tests, smoke(), the default bench line
O=gpurun_out/${1:-full}
mkdir -p $O
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1 | tee $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
cut -c1-330 $O/bench.json
