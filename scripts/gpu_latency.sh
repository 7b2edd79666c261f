#!/bin/bash
# batch-1 latency check of a build: the GPU suite, latency + stage probes
O=gpurun_out/${1:-lat}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -s -k "properties_at_baseline" 2>&1 | grep -i "batch independence\|passed\|failed\|Error" | tee $O/pytest_prop.log
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/pytest.log
timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
timeout 300 python scripts/segment_probe.py 50 2>&1 | tail -1 | tee -a $O/segments.jsonl
