#!/bin/bash
O=gpurun_out/${1:-lat}
mkdir -p $O
timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
OPP_B200_TWO_STREAMS=0 timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
timeout 300 python scripts/segment_probe.py 50 2>&1 | tail -1 | tee -a $O/segments.jsonl
timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "fine_windows or graph_mode or resident" 2>&1 | tail -5 | tee $O/pytest.log
