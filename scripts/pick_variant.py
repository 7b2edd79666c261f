"""Reads gpurun_out/variants/{f000,f111}.json (scripts/variant_probe.py) and decides feature by
feature from the per-op times: L = coalesced LayerNorm-epilogue I/O (opp_linear_ln), C = coalesced
conf_matrix store (opp_sim_conf), V = conv epilogue on two warp groups + vectorised pe
(opp_conv2d_nhwc), and the runtime option conv1_staged (opp_conv1_7x7).  Prints shell assignments
selecting the matching prebuilt library: `export OPP_B200_LIB=... OPP_CONV1_STAGED=...`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VAR = os.path.join(ROOT, "gpurun_out", "variants")


def load(tag):
    try:
        return json.loads(open(os.path.join(VAR, tag + ".json")).read())
    except (OSError, ValueError):
        return None


def log(*a):
    print("#", *a, file=sys.stderr)


base, full = load("f000"), load("f111")
if base is None:
    log("no f000 probe result")
    sys.exit(1)
for r in (base, full):
    if r:
        log(r["tag"], {k: (v if v == "ok" else v[:90]) for k, v in r["checks"].items()},
            {k: round(v["ms_per_forward"], 2) for k, v in r["timing"].items()})

D0, D1 = "conv1_staged=0", "conv1_staged=1"


def ok(r, *names):
    return r is not None and all(r["checks"].get(n) == "ok" for n in names)


def ops(r, label):
    return r["timing"][label]["ops_ms"] if r and label in r["timing"] else None


L = C = V = S = 0
b0, f0 = ops(base, D0), ops(full, D0)
if b0 and f0 and ok(full, f"golden[{D0}]", f"kv_state[{D0}]", f"conv1[{D0}]"):
    log("linear_ln", b0["opp_linear_ln"], "->", f0["opp_linear_ln"], " sim_conf", b0["opp_sim_conf"], "->",
        f0["opp_sim_conf"], " conv2d", b0["opp_conv2d_nhwc"], "->", f0["opp_conv2d_nhwc"])
    if ok(full, "linear_ln") and f0["opp_linear_ln"] < 0.99 * b0["opp_linear_ln"]:
        L = 1
    if ok(full, "sim") and f0["opp_sim_conf"] < 0.99 * b0["opp_sim_conf"]:
        C = 1
    if ok(full, "conv") and f0["opp_conv2d_nhwc"] <= 1.005 * b0["opp_conv2d_nhwc"]:
        V = 1
b1 = ops(base, D1)
if b0 and b1 and ok(base, f"conv1[{D1}]", f"golden[{D1}]"):
    log("conv1_7x7", b0["opp_conv1_7x7"], "->", b1["opp_conv1_7x7"])
    if b1["opp_conv1_7x7"] < 0.95 * b0["opp_conv1_7x7"]:
        S = 1
print(f"export OPP_B200_LIB={ROOT}/variants/libopp_f{L}{C}{V}.so OPP_CONV1_STAGED={S}"
      f"  # ln_staged={L} conf_staged={C} conv2+pe_vec={V} conv1_staged={S}")
