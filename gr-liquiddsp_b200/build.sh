#!/bin/bash
# Build liblqb200.so (sm_100a only) in-tree.  Usage: gr-liquiddsp_b200/build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$HERE/lib"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
SRCS="$HERE/csrc/lqb_tables.cpp $HERE/csrc/lqb_api.cu $HERE/csrc/lqb_rx_seek.cu $HERE/csrc/lqb_rx_payload.cu $HERE/csrc/lqb_rx_fec.cu $HERE/csrc/lqb_rx_soft.cu"
for f in lqb_tx.cu lqb_debug.cu lqb_liquid_compat.cpp; do [ -f "$HERE/csrc/$f" ] && SRCS="$SRCS $HERE/csrc/$f"; done
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false \
      -Xcompiler -fPIC,-ffp-contract=off,-Wall -Xptxas -v \
      -shared -o "${LQB_OUT:-$HERE/lib/liblqb200.so}" $SRCS "$@"
