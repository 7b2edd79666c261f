import sys, time, torch
sys.path.insert(0, '.')
from oracle import oracle, workload
from tests import golden_io, parity
sd = workload.synthetic_state_dict(0)
for case in golden_io.cases():
    data, z = golden_io.load(case)
    got = parity.run_cuda(data)
    rep = parity.report(got, {k: z[k] for k in z.files})
    print(case, rep)
    conf = got["conf_matrix"].cpu()
    print("  conf rowmax maxdiff", (conf.max(2).values - torch.from_numpy(z["conf_rowmax"])).abs().max().item())
    if rep["indices_equal"]:
        rel = ((got["mconf"].cpu() - torch.from_numpy(z["mconf"])).abs() / torch.from_numpy(z["mconf"])).max().item()
        print("  mconf max rel", rel, " expec_f per col", (got["expec_f"].cpu() - torch.from_numpy(z["expec_f"])).abs().max(0).values.tolist())
for (h, w, n, npl, B) in [(256, 320, 1500, 700, 2), (512, 512, 5000, 3000, 1)]:
    data, meta = workload.planted_workload(sd, h, w, n, npl, batch=B)
    d_or = {k: v.clone() for k, v in data.items()}
    t = time.time(); oracle.forward(sd, d_or); t_or = time.time() - t
    got = parity.run_cuda(data)
    rep = parity.report(got, d_or)
    print((h, w, n, B), "oracle %.2fs" % t_or, rep)
    if not rep["indices_equal"]:
        a = set(zip(got["b_ids"].tolist(), got["i_ids"].tolist(), got["j_ids"].tolist()))
        b = set(zip(d_or["b_ids"].tolist(), d_or["i_ids"].tolist(), d_or["j_ids"].tolist()))
        print("  only cuda", len(a - b), "only oracle", len(b - a))
        print("  oracle mconf of missing:", [round(d_or["conf_matrix"][x].item(), 4) for x in list(b - a)[:8]])
        print("  cuda   mconf of extra  :", [round(got["conf_matrix"][x].item(), 4) for x in list(a - b)[:8]])
    conf_d = (got["conf_matrix"].cpu() - d_or["conf_matrix"]).abs().max().item()
    print("  conf maxdiff", conf_d)
# timing
from onepose_plus_plus_b200 import _lib
for B in (1, 16):
    data, meta = workload.planted_workload(sd, 512, 512, 5000, 3000, batch=B)
    d = {k: v.cuda() for k, v in data.items()}
    m = parity.cuda_model()
    for _ in range(3): m(dict(d))
    torch.cuda.synchronize(); t = time.time()
    for _ in range(5): m(dict(d))
    torch.cuda.synchronize(); print(f"B={B} forward ms", (time.time() - t) / 5 * 1000, "precision", m.precision)
    _lib.profile_ops(lambda: m(dict(d)), sys.stdout)
