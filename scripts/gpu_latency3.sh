#!/bin/bash
O=gpurun_out/${1:-lat}
mkdir -p $O
timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
timeout 300 python scripts/segment_probe.py 50 2>&1 | tail -1 | tee -a $O/segments.jsonl
OPP_NSPLIT=0 timeout 300 python scripts/segment_probe.py 50 2>&1 | tail -1 | tee -a $O/segments.jsonl
