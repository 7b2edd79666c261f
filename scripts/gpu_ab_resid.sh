#!/bin/bash
# A/B of a compile-time variant library against the default one: kernel checks, golden parity,
# per-op CUDA-event times at batch 64 (scripts/variant_probe.py), twice each to see the noise.
O=gpurun_out/${1:-ab}
mkdir -p $O gpurun_out/variants
for rep in 1 2; do
  timeout 300 python scripts/variant_probe.py default$rep "" > $O/default$rep.log 2>&1; cp gpurun_out/variants/default$rep.json $O/
  OPP_B200_LIB=$PWD/variants/libopp_residstaged.so timeout 300 python scripts/variant_probe.py residstaged$rep "" > $O/residstaged$rep.log 2>&1; cp gpurun_out/variants/residstaged$rep.json $O/
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/variants/*[12].json")):
    r=json.load(open(f)); t=r["timing"].get("default",{})
    bad=[k for k,v in r["checks"].items() if v!="ok"]
    print(f, "ms/fwd", round(t.get("ms_per_forward",0),2), "conv2d", t.get("ops_ms",{}).get("opp_conv2d_nhwc"), "bad", bad)
PY
