#!/bin/bash
# Builds epilogue warp-group variants of libopp_b200.so into variants/ (git-ignored, but shipped to
# the GPU box by gpurun) so that one GPU session can check and time them side by side
# (scripts/gpu_variants.sh).  Only opp_gemm.cu depends on the flags.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
build() {
  name=$1; shift
  OPP_OUT=$PWD/variants/libopp_$name.so OPP_OBJ=$PWD/variants/obj_$name \
    bash onepose_plus_plus_b200/csrc/build.sh "$@" > variants/build_$name.log 2>&1 && echo "built $name"
}
build base &
build ln2 -DOPP_LN_GROUPS=2 &
build conv2 -DOPP_CONV_GROUPS=2 &
build ln2conv2 -DOPP_LN_GROUPS=2 -DOPP_CONV_GROUPS=2 &
wait
ls -la variants/*.so
