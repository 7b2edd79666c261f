"""Forward timing + per-op breakdown: python scripts/time_forward.py [batch] [precision]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle, workload
from onepose_plus_plus_b200 import OnePosePlus_model, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
precision = sys.argv[2] if len(sys.argv) > 2 else "fp16x3"
sd = workload.synthetic_state_dict(0)
m = OnePosePlus_model(oracle.DEFAULT_CONFIG, precision=precision); m.load_state_dict(sd); m = m.eval().cuda()
data, _ = workload.planted_workload(sd, 512, 512, 5000, 3000, batch=B)
d = {k: v.cuda() for k, v in data.items()}
for _ in range(3): m(dict(d))
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
for _ in range(5): out = dict(d); m(out)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"B={B} {precision} OPP_CLUSTER={os.environ.get('OPP_CLUSTER','default')}: {ms:.3f} ms/forward, {B/ms*1e3:.1f} img/s, M={out['b_ids'].numel()}")
_lib.profile_ops(lambda: m(dict(d)), sys.stdout)
