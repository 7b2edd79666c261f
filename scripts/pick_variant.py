"""Reads gpurun_out/variants/{base,ln2conv2}.json (scripts/variant_probe.py) and decides, kernel by
kernel, which build options pay off: LayerNorm epilogue warp groups (opp_linear_ln time), conv
epilogue warp groups (opp_conv2d_nhwc time) and the KV-state kernel (kv_mma).  Prints shell
assignments selecting the matching prebuilt library: `export OPP_B200_LIB=... OPP_KV_MMA=...`."""
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


base, two = load("base"), load("ln2conv2")
if base is None:
    log("no base probe result")
    sys.exit(1)
for r in (base, two):
    if r:
        log(r["tag"], {k: (v if v == "ok" else v[:80]) for k, v in r["checks"].items()},
            {k: round(v["ms_per_forward"], 2) for k, v in r["timing"].items() if isinstance(v, dict)})


def ok(r, *names):
    return r is not None and all(r["checks"].get(n) == "ok" for n in names)


kv = 0
t0, t1 = base["timing"].get("kv_mma=0"), base["timing"].get("kv_mma=1")
if ok(base, "kv_state[kv_mma=1]", "golden[kv_mma=1]") and t1 and (t0 is None or
                                                                  t1["ms_per_forward"] < t0["ms_per_forward"]):
    kv = 1
ln = conv = 1
key = "kv_mma=0" if (two and "kv_mma=0" in two["timing"] and t0) else "kv_mma=1"
if two and key in two["timing"] and key in base["timing"] and ok(two, "golden[kv_mma=0]"):
    ob, ot = base["timing"][key]["ops_ms"], two["timing"][key]["ops_ms"]
    log("linear_ln ms base/2-group:", ob.get("opp_linear_ln"), ot.get("opp_linear_ln"),
        " conv ms base/2-group:", ob.get("opp_conv2d_nhwc"), ot.get("opp_conv2d_nhwc"))
    if ok(two, "linear_ln") and ot["opp_linear_ln"] < 0.97 * ob["opp_linear_ln"]:
        ln = 2
    if ok(two, "conv") and ot["opp_conv2d_nhwc"] < 0.97 * ob["opp_conv2d_nhwc"]:
        conv = 2
name = {(1, 1): "base", (2, 1): "ln2", (1, 2): "conv2", (2, 2): "ln2conv2"}[(ln, conv)]
print(f"export OPP_B200_LIB={ROOT}/variants/libopp_{name}.so OPP_KV_MMA={kv}  # ln_groups={ln} conv_groups={conv}")
