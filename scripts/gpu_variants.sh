#!/bin/bash
# One short GPU session over the library variants built by scripts/build_variants.sh:
# probe the default build and the two-warp-group build (checks + per-op timing), pick per kernel, run the GPU test suite and a short
# bench with it.  Everything lands in gpurun_out/variants/.
mkdir -p gpurun_out/variants
for v in base ln2conv2; do
  OPP_B200_LIB=$PWD/variants/libopp_$v.so timeout 120 python scripts/variant_probe.py $v \
    > gpurun_out/variants/$v.log 2>&1
  echo "$v exit=$? $(tail -c 600 gpurun_out/variants/$v.log | tr '\n' ' ' | tail -c 400)"
done
python scripts/pick_variant.py > gpurun_out/variants/best.env 2> gpurun_out/variants/pick.log
cat gpurun_out/variants/pick.log gpurun_out/variants/best.env
if [ -s gpurun_out/variants/best.env ]; then
  source gpurun_out/variants/best.env
  timeout 240 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/variants/pytest_best.log
  timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/variants/bench_best.json \
    2> gpurun_out/variants/bench_best.err
  cat gpurun_out/variants/bench_best.json
fi
# evidence for profiles/ (last: these are the first to go if the session runs out of time)
if [ -s gpurun_out/variants/best.env ]; then
  timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/variants/launches_b8.csv python scripts/profile_step.py 8 > gpurun_out/variants/ncu_launches.log 2>&1
  tail -1 gpurun_out/variants/ncu_launches.log
  timeout 150 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:"kv_partial|EpiLN" -c 6 -o gpurun_out/variants/prof_kv_ln -f python scripts/profile_step.py 8 \
    > gpurun_out/variants/ncu_kv_ln.log 2>&1
  tail -1 gpurun_out/variants/ncu_kv_ln.log
fi
