#!/bin/bash
# Builds compile-time variants of libopp_b200.so into variants/ (git-ignored, but shipped to the
# GPU box by gpurun) so that one GPU session can check and time them next to the default build
# (scripts/gpu_variants.sh).  Current candidates:
#   residstaged: -DOPP_CONV_RESID_STAGED=1  (BasicBlock residual through the transpose buffer)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
rm -f variants/libopp_*.so
build() {
  name=$1; shift
  OPP_OUT=$PWD/variants/libopp_$name.so OPP_OBJ=$PWD/variants/obj_$name \
    bash onepose_plus_plus_b200/csrc/build.sh "$@" > variants/build_$name.log 2>&1 && echo "built $name"
}
build residstaged -DOPP_CONV_RESID_STAGED=1 &
wait
cp onepose_plus_plus_b200/libopp_b200.so variants/libopp_default.so
ls -la variants/*.so
