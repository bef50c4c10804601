#!/bin/bash
# usage: tools/build_variant_fused.sh NAME -DFLAG=.. ...   -> openmm_sdm_plugin_b200/libsdmb200_NAME.so
# (displaced-atom kernel A/B variants: only kernels_fused.cu is recompiled with the extra flags)
set -e
cd "$(dirname "$0")/../openmm_sdm_plugin_b200/csrc"
NAME=$1; shift
NVCC=/usr/local/cuda/bin/nvcc
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC "$@" -c kernels_fused.cu -o /tmp/kf_$NAME.o
$NVCC $ARCH -shared -cudart static -o ../libsdmb200_$NAME.so api.o /tmp/kf_$NAME.o kernels_elementwise.o kernels_rows.o kernels_md.o kernels_restraints.o kernels_pme.o kernels_gb.o pairlist.o -ldl
echo built libsdmb200_$NAME.so
