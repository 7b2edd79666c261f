#!/bin/bash
# Builds feature variants of libopp_b200.so into variants/ (git-ignored, but shipped to the GPU box
# by gpurun) so that one GPU session can check and time them side by side (scripts/gpu_variants.sh).
# Name f<L><C><V>:  L = OPP_LN_STAGED (coalesced LayerNorm-epilogue I/O), C = OPP_CONF_STAGED
# (coalesced conf_matrix store), V = conv epilogue on two warp groups + vectorised pe loads.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
rm -f variants/libopp_*.so
build() {
  name=$1; shift
  OPP_OUT=$PWD/variants/libopp_$name.so OPP_OBJ=$PWD/variants/obj_$name \
    bash onepose_plus_plus_b200/csrc/build.sh "$@" > variants/build_$name.log 2>&1 && echo "built $name"
}
for L in 0 1; do for C in 0 1; do for V in 0 1; do
  flags="-DOPP_LN_STAGED=$L -DOPP_CONF_STAGED=$C"
  [ $V = 1 ] && flags="$flags -DOPP_CONV_GROUPS=2 -DOPP_PE_VEC=1"
  build f$L$C$V $flags &
done; done; wait; done
ls -la variants/*.so
