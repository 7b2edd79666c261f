#!/bin/bash
# check of a build with the batch-1 view: the GPU suite, latency + stage probes, the bench line
O=gpurun_out/${1:-lat}
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 | tee $O/pytest.log
timeout 300 python scripts/latency_probe.py 50 2>&1 | tail -1 | tee -a $O/latency.jsonl
timeout 300 python scripts/segment_probe.py 50 2>&1 | tail -1 | tee -a $O/segments.jsonl
timeout 600 python bench.py --steps 10 --warmup 3 --profile-ops > $O/bench.json 2> $O/bench.err
cut -c1-300 $O/bench.json
