#!/bin/bash
# GPU session 2: full GPU suite (incl. PnP), KV1 adoption check, ncu --set full of one whole
# batch-64 forward (raw CSV), source-level capture of the K=256 transformer GEMMs.
O=gpurun_out/${1:-s2}
mkdir -p $O
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee $O/pytest_gpu.log
OPP_B200_KV1=1 timeout 600 python -m pytest tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -8 | tee $O/pytest_kv1.log
timeout 900 ncu --set full --clock-control none --profile-from-start off --csv --page raw \
    --log-file $O/ncu_b64_raw.csv python scripts/profile_step.py 64 > $O/ncu_b64.log 2>&1
tail -2 $O/ncu_b64.log; ls -la $O/ncu_b64_raw.csv
timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -k regex:gemm_kernel -s 28 -c 5 -o $O/prof_xfmr_b64 -f python scripts/profile_step.py 64 > $O/ncu_xfmr.log 2>&1
tail -2 $O/ncu_xfmr.log; ls -la $O/
