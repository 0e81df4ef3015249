#!/bin/bash
# Kernel-variant experiments: builds voxelyze_b200/lib/variants/lib<name>.so from the product sources with extra -D flags.
# usage: tools/build_variant.sh <name> [-DFLAG=...]...     run with VX_PRODUCT_SO=<that .so> (capi.load_product honours it)
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
mkdir -p voxelyze_b200/lib/variants
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -diag-suppress 177 \
  -ccbin /usr/bin/g++ -I include -I voxelyze_b200/csrc "$@" -o voxelyze_b200/lib/variants/lib$NAME.so voxelyze_b200/csrc/vx_capi.cu
echo voxelyze_b200/lib/variants/lib$NAME.so
