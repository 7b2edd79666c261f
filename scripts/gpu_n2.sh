#!/bin/bash
# 2-GPU check: smoke(), then the bench exactly as the driver launches it at N = 2
O=gpurun_out/${1:-n2}
mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $O/smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err
tail -3 $O/bench_n2.err; cut -c1-400 $O/bench_n2.json
