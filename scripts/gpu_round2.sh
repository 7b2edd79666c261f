#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_model_gpu.py -x -q 2>&1 | tail -4
python scripts/time_forward.py 16 fp16 2>&1 | grep -v Warn | head -8
python scripts/time_forward.py 16 fp16x3 2>&1 | grep -v Warn | head -14
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -3 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2>/dev/null; cat gpurun_out/bench_ref_n2.json
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -3 gpurun_out/bench_n1.err; cat gpurun_out/bench_n1.json
