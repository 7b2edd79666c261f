#!/bin/bash
# quick check of a build: bench line as the driver runs it (defaults), then smoke()
O=gpurun_out/${1:-q}
mkdir -p $O
timeout 900 python bench.py > $O/bench.json 2> $O/bench.err
cut -c1-330 $O/bench.json; tail -2 $O/bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.log
