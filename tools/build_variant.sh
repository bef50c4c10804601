#!/bin/bash
# usage: tools/build_variant.sh NAME -DFLAG=.. ...   -> openmm_sdm_plugin_b200/libsdmb200_NAME.so
# (pair-kernel A/B variants: only kernels_rows.cu is recompiled with the extra flags)
set -e
cd "$(dirname "$0")/../openmm_sdm_plugin_b200/csrc"
NAME=$1; shift
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v "$@" -c kernels_rows.cu -o /tmp/kc_$NAME.o 2> /tmp/kc_$NAME.log
grep -A3 "pair_row_kernelILi8ELb1ELb1ELb0" /tmp/kc_$NAME.log | grep -E "registers|spill" | head -8
$NVCC $ARCH -shared -cudart static -o ../libsdmb200_$NAME.so api.o kernels_fused.o kernels_elementwise.o kernels_md.o kernels_restraints.o kernels_pme.o pairlist.o /tmp/kc_$NAME.o
echo built libsdmb200_$NAME.so
