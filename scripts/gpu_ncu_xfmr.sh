#!/bin/bash
# ncu --set full of the first coarse-transformer layer (5 GEMMs per sequence) and of the four
# dual-softmax passes, default build.  -k matches the base kernel name only (template arguments
# are not part of it), so the launches are selected by position among the gemm_kernel launches:
# 21 convolutions, then 60 transformer GEMMs, then EpiLse x2, EpiConf x2.
mkdir -p gpurun_out/final
timeout 70 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:gemm_kernel -s 21 -c 10 -o gpurun_out/final/prof_xfmr -f python scripts/profile_step.py 8 \
  > gpurun_out/final/ncu_xfmr.log 2>&1
tail -2 gpurun_out/final/ncu_xfmr.log
timeout 60 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:gemm_kernel -s 81 -c 4 -o gpurun_out/final/prof_sim -f python scripts/profile_step.py 8 \
  > gpurun_out/final/ncu_sim.log 2>&1
tail -2 gpurun_out/final/ncu_sim.log
