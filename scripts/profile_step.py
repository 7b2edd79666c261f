"""One profiled forward for ncu: warm up, then bracket a single forward with cudaProfilerStart/Stop.
    ncu --profile-from-start off ... python scripts/profile_step.py [batch] [precision]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle, workload  # noqa: E402  (workload + checkpoint generators only)
from onepose_plus_plus_b200 import OnePosePlus_model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
sd = workload.synthetic_state_dict(0)
m = OnePosePlus_model(oracle.DEFAULT_CONFIG, precision=precision)
m.load_state_dict(sd)
m = m.eval().cuda()
data, _ = workload.planted_workload(sd, 512, 512, 5000, 3000, batch=B)
d = {k: v.cuda() for k, v in data.items()}
for k in ("keypoints3d", "descriptors3d_db", "descriptors3d_coarse_db"):   # one object shared by the batch
    d[k] = d[k][:1].contiguous()
for _ in range(2):
    m(dict(d))
torch.cuda.synchronize()
torch.cuda.profiler.start()
out = dict(d)
m(out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled forward: B", B, "M", out["b_ids"].numel(), precision)
