"""Small end-to-end forwards for compute-sanitizer (memcheck): eager, window / dense fine head,
CUDA-graph mode with a resident bank, a batch of 2 with distinct objects.
    compute-sanitizer --tool memcheck python scripts/sanitize_small.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle, workload  # noqa: E402  (test infrastructure: planted workloads)
from onepose_plus_plus_b200 import OnePosePlus_model  # noqa: E402

sd = workload.synthetic_state_dict(0)
m = OnePosePlus_model(oracle.DEFAULT_CONFIG)
m.load_state_dict(sd)
m = m.eval().cuda()
cases = [workload.planted_workload(sd, 128, 160, 400, 200, batch=1)[0],
         workload.hetero_workload(sd, 96, 128, 300, 150, batch=2)]
for d in cases:
    d = d[0] if isinstance(d, tuple) else d
    dd = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}
    for mode in ("sparse", "dense"):
        m.fine_windows = mode
        out = dict(dd)
        m(out)
        torch.cuda.synchronize()
        print(f"eager fine_windows={mode}: M={out['b_ids'].numel()}")
    m.fine_windows = "auto"
d = cases[0]
dd = {k: v.cuda() for k, v in d.items()}
m.set_bank(dd["keypoints3d"], dd["descriptors3d_db"], dd["descriptors3d_coarse_db"])
m.conf_matrix_mode = "lazy"
m.enable_cuda_graphs(True)
for _ in range(2):
    out = {"query_image": dd["query_image"], "query_image_scale": dd["query_image_scale"]}
    m(out)
torch.cuda.synchronize()
print(f"graph mode: M={out['b_ids'].numel()}")
print("sanitize_small done")
