"""Reads gpurun_out/variants/probe.json (scripts/variant_probe.py run on the default build with one
option set per experimental feature) and decides feature by feature from the per-op times:
  upsample_rows -> opp_upsample2x_add      conv1_px4 -> opp_conv1_7x7      fine_attn_vec -> opp_fine_attention
  colmax        -> opp_sim_conf + opp_sim_conf_colmax + opp_best_finalize + opp_match_select(_colmax)
  lse_cols      -> opp_sim_lse + opp_sim_lse_cols + opp_lse_finalize + opp_lse_col_finalize
Prints shell assignments: `export OPP_UPSAMPLE_ROWS=.. OPP_CONV1_PX4=.. OPP_B200_COLMAX=..`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    r = json.loads(open(os.path.join(ROOT, "gpurun_out", "variants", "probe.json")).read())
except (OSError, ValueError):
    print("# no probe result", file=sys.stderr)
    sys.exit(1)
print("#", {k: (v if v == "ok" else v[:90]) for k, v in r["checks"].items()}, file=sys.stderr)
print("#", {k: round(v["ms_per_forward"], 2) for k, v in r["timing"].items()}, file=sys.stderr)
base = r["timing"].get("default", {}).get("ops_ms")
FEATURES = {"upsample_rows": (["opp_upsample2x_add"], "upsample_rows", "OPP_UPSAMPLE_ROWS"),
            "conv1_px4": (["opp_conv1_7x7"], "conv1_px4", "OPP_CONV1_PX4"),
            "colmax": (["opp_sim_conf", "opp_sim_conf_colmax", "opp_best_finalize", "opp_match_select",
                        "opp_match_select_colmax"], "sim_colmax", "OPP_B200_COLMAX"),
            "fine_attn_vec": (["opp_fine_attention"], "fine_attn_vec", "OPP_FINE_ATTN_VEC"),
            "kv1": (["opp_linear_act_f16", "opp_linear_act_f16_out1", "opp_kv_partial"], "kv_single_plane",
                    "OPP_B200_KV1"),
            "lse_cols": (["opp_sim_lse", "opp_sim_lse_cols", "opp_lse_finalize", "opp_lse_col_finalize"],
                         "sim_lse_cols", "OPP_B200_LSECOLS")}
out = {}
for opt, (ops_, check, env) in FEATURES.items():
    label = f"{opt}=1"
    t = r["timing"].get(label, {}).get("ops_ms")
    ok = all(r["checks"].get(f"{c}[{label}]") == "ok" for c in (check, "golden", "kv_state", "conv1", "upsample"))
    on = 0
    if base and t and ok:
        tb, tn = sum(base.get(o, 0.0) for o in ops_), sum(t.get(o, 0.0) for o in ops_)
        print(f"# {opt}: {tb:.3f} -> {tn:.3f} ms", file=sys.stderr)
        on = int(tn < 0.97 * tb)
    out[env] = on
# compile-time candidate: the residual-staged conv epilogue (probed with default options)
try:
    rs = json.loads(open(os.path.join(ROOT, "gpurun_out", "variants", "residstaged.json")).read())
    t = rs["timing"].get("default", {}).get("ops_ms")
    good = all(v == "ok" for v in rs["checks"].values())
    if base and t and good:
        print(f"# conv2d default -> resid-staged: {base['opp_conv2d_nhwc']:.3f} -> {t['opp_conv2d_nhwc']:.3f} ms",
              file=sys.stderr)
        if t["opp_conv2d_nhwc"] < 0.985 * base["opp_conv2d_nhwc"]:
            out["OPP_B200_LIB"] = rs["lib"]
except (OSError, ValueError, KeyError):
    pass
print("export " + " ".join(f"{k}={v}" for k, v in out.items()))
