#!/bin/bash
# usage (under gpurun): tools/gpu_prof.sh TAG [LIB]  -- one `ncu --set full` capture of the pair kernel
TAG=${1:-x}
[ -n "$2" ] && export SDMB200_LIB=$PWD/$2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_${KERNEL:-row}_kernel -s 3 -c 1 \
  -o gpurun_out/prof_pair_$TAG -f python bench.py --no-cpu-baseline --no-single-lambda --steps 2 --warmup 1 \
  > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-200
