"""Batch-1 latency view (BASELINE configs[1]): wall clock per forward in CUDA-graph mode, the GPU
time of the captured graph alone, and bit-equality of the graph outputs with the eager path.
    [OPP_PDL=0] [OPP_NSPLIT=0] python scripts/latency_probe.py [n]"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle, workload  # noqa: E402  (test infrastructure: the planted workload)
from onepose_plus_plus_b200 import OnePosePlus_model  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
sd = workload.synthetic_state_dict(0)
m = OnePosePlus_model(oracle.DEFAULT_CONFIG)
m.load_state_dict(sd)
m = m.eval().cuda()
data, _ = workload.planted_workload(sd, 512, 512, 5000, 3000, batch=1)
d = {k: v.cuda() for k, v in data.items()}
img = (d["query_image"] * 255).round().clamp(0, 255).to(torch.uint8)
m.conf_matrix_mode = "lazy"
m.set_bank(d["keypoints3d"], d["descriptors3d_db"], d["descriptors3d_coarse_db"])
q = {"query_image": img, "query_image_scale": d["query_image_scale"]}
ref = dict(q)
m(ref)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(n):
    e = dict(q)
    m(e)
torch.cuda.synchronize()
eager_ms = (time.perf_counter() - t0) / n * 1e3
m.enable_cuda_graphs(True)
for _ in range(3):
    g = dict(q)
    m(g)
torch.cuda.synchronize()
same = all(torch.equal(ref[k], g[k]) for k in ("b_ids", "i_ids", "j_ids", "mconf", "mkpts_query_f", "expec_f",
                                               "mkpts_3d_db", "mkpts_query_c"))
t0 = time.perf_counter()
for _ in range(n):
    g = dict(q)
    m(g)
torch.cuda.synchronize()
graph_ms = (time.perf_counter() - t0) / n * 1e3
ent = next(iter(m._graphs.values()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    ent["graph"].replay()
e1.record()
torch.cuda.synchronize()
gpu_ms = e0.elapsed_time(e1) / n
print(json.dumps({"OPP_PDL": os.environ.get("OPP_PDL", "1"), "OPP_NSPLIT": os.environ.get("OPP_NSPLIT", "1"),
                  "eager_ms": round(eager_ms, 4), "graph_wall_ms": round(graph_ms, 4),
                  "graph_gpu_ms": round(gpu_ms, 4), "matches": int(g["b_ids"].numel()),
                  "graph_equals_eager": bool(same)}))
