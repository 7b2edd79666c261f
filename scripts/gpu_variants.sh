#!/bin/bash
# One short GPU session over the experimental (built, off-by-default) paths of the default library:
# probe each (its kernel check + golden parity + per-op timing at batch 64), pick feature by
# feature, then run the GPU test suite, smoke, the bench and ncu captures with the picked options.
# Everything lands in gpurun_out/variants/.  Afterwards: flip the OPP_*_DEFAULT macros / the model
# default for the adopted ones, rebuild, `python scripts/sass_diff.py` against this build.
mkdir -p gpurun_out/variants
V=gpurun_out/variants
timeout 200 python scripts/variant_probe.py probe "" upsample_rows=1 conv1_px4=1 colmax=1 lse_cols=1 fine_attn_vec=1 kv1=1 \
  > $V/probe.log 2>&1
echo "probe exit=$?"
# compile-time candidate (scripts/build_variants.sh): same probe with default options
if [ -f variants/libopp_residstaged.so ]; then
  OPP_B200_LIB=$PWD/variants/libopp_residstaged.so timeout 100 python scripts/variant_probe.py residstaged "" \
    > $V/residstaged.log 2>&1
  echo "residstaged exit=$?"
fi
python scripts/pick_variant.py > $V/best.env 2> $V/pick.log
cat $V/pick.log $V/best.env
[ -s $V/best.env ] || exit 1
source $V/best.env
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $V/pytest_best.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $V/smoke_best.log
timeout 200 python bench.py --steps 10 --warmup 3 > $V/bench_best.json 2> $V/bench_best.err
cat $V/bench_best.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $V/launches_b8.csv python scripts/profile_step.py 8 > $V/ncu_launches.log 2>&1
tail -1 $V/ncu_launches.log
