"""Where the batch-1 graph time goes: capture the stages of the forward as separate CUDA graphs
(backbone, coarse transformer with/without the side stream, coarse matching, fine stage) and time
the replay of each with CUDA events.    python scripts/segment_probe.py [n]"""
import json
import os
import sys

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
scale = d["query_image_scale"].float().contiguous()
m.conf_matrix_mode = "lazy"
m.set_bank(d["keypoints3d"], d["descriptors3d_db"], d["descriptors3d_coarse_db"])
m({"query_image": img, "query_image_scale": scale})       # plan + workspace
torch.cuda.synchronize()
dev = img.device
res = {}


def timed(name, fn):
    fn()                       # warm-up outside the capture
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / n * 1e3, 1)   # us
    return out, g


with torch.no_grad():
    (q2, x1_lat, (hc, wc)), g_bb = timed("backbone_trunk", lambda: m._backbone(img, defer_fine=True))
    bank = m._resident_bank_state()
    S, N = hc * wc, bank["N"]

    def xf(two):
        m._side_stream = m._aux_stream(dev) if two else None
        try:
            return m._coarse_transformer(q2, bank, 1, S, N)
        finally:
            m._side_stream = None

    (_, _), _g = timed("transformer_one_stream", lambda: xf(False))
    (c2, c3), g_xf = timed("transformer_two_streams", lambda: xf(True))
    out = {}

    def cm():
        out.clear()
        return m._coarse_matching(c2, c3, bank, scale, 1, N, hc, wc, 8.0, out)

    (count, cap), g_cm = timed("coarse_matching", cm)
    ids = (out["b_ids"], out["i_ids"], out["j_ids"], out["mkpts_query_c"])
    fcap = min(cap, min(N, S))
    # run the real coarse stage once so that the device-side match count is the real one
    g_bb.replay()        # (the transformer ping-pongs through its own input buffer)
    g_xf.replay()
    g_cm.replay()
    torch.cuda.synchronize()
    hf, wf = x1_lat.shape[1:3]
    fw, _g = timed("fine_head_windows", lambda: m._fine_head_windows(x1_lat, out["b_ids"], out["j_ids"], fcap, wc, 4,
                                                                  count=count))
    _, _g = timed("fine_head_dense", lambda: m._fine_head_dense(x1_lat))
    _, g_f = timed("fine", lambda: m._fine(fw, bank, ids, fcap, scale, hc, wc, (512, 512), {}, count=count,
                                           windows_hw=(hf, wf)))
res["matches"] = int(count.item())
res["sum_us"] = round(res["backbone_trunk"] + res["transformer_two_streams"] + res["coarse_matching"]
                      + res["fine_head_windows"] + res["fine"], 1)
print(json.dumps(res))
