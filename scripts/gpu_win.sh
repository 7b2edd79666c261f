#!/bin/bash
O=gpurun_out/${1:-win}
mkdir -p $O
timeout 600 python tests/kernel_checks.py conv_win 2>&1 | tail -15 | tee $O/conv_win.log
timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "fine_windows or graph_mode or golden" 2>&1 | tail -15 | tee $O/pytest_win.log
timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
OPP_B200_FINE_WINDOWS=dense timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
timeout 500 python bench.py --steps 10 --warmup 3 --profile-ops --no-c5 --no-cpu-baseline > $O/bench.json 2> $O/bench.err
head -12 $O/bench.err; cut -c1-330 $O/bench.json
