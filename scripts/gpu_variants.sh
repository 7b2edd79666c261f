#!/bin/bash
# One short GPU session over the library variants built by scripts/build_variants.sh: probe the
# default build (f000) and the all-features build (f111) (kernel checks + golden parity + per-op
# timing), pick feature by feature, then run the GPU test suite, smoke, the bench and the ncu
# captures with the picked configuration.  Everything lands in gpurun_out/variants/.
mkdir -p gpurun_out/variants
V=gpurun_out/variants
OPP_B200_LIB=$PWD/variants/libopp_f000.so timeout 100 python scripts/variant_probe.py f000 conv1_staged=0 conv1_staged=1 > $V/f000.log 2>&1
echo "f000 exit=$?"
OPP_B200_LIB=$PWD/variants/libopp_f111.so timeout 100 python scripts/variant_probe.py f111 conv1_staged=0 > $V/f111.log 2>&1
echo "f111 exit=$?"
python scripts/pick_variant.py > $V/best.env 2> $V/pick.log
cat $V/pick.log $V/best.env
[ -s $V/best.env ] || exit 1
source $V/best.env
timeout 200 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee $V/pytest_best.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $V/smoke_best.log
timeout 200 python bench.py --steps 10 --warmup 3 > $V/bench_best.json 2> $V/bench_best.err
cat $V/bench_best.json
# evidence for profiles/ (last: these are the first to go if the session runs out of time)
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file $V/launches_b8.csv python scripts/profile_step.py 8 > $V/ncu_launches.log 2>&1
tail -1 $V/ncu_launches.log
timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"EpiStoreF16|EpiQ|EpiLN" -c 10 -o $V/prof_xfmr -f python scripts/profile_step.py 8 > $V/ncu_xfmr.log 2>&1
tail -1 $V/ncu_xfmr.log
timeout 120 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:"EpiConf|EpiLse|conv1_7x7" -c 5 -o $V/prof_sim -f python scripts/profile_step.py 8 > $V/ncu_sim.log 2>&1
tail -1 $V/ncu_sim.log
