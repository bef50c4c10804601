#!/bin/bash
# usage (under gpurun): tools/gpu_buildcost.sh -- cost of a list build at R = 1 and 16, plus row-kernel variants for DESIGN.md section 8
python tools/build_cost.py --replicas 16 2>&1 | tee gpurun_out/build_cost.txt
python tools/build_cost.py --replicas 1 2>&1 | tee -a gpurun_out/build_cost.txt
for v in "SDMB200_ROW_GROUP=2" "SDMB200_ROW_CHUNK=8" "SDMB200_ROW_CHUNK=16" "SDMB200_PAIR_RESIDENT=16" "SDMB200_PAIR_RESIDENT=20" "SDMB200_ROW_LPT=0" "X=0"; do
  env $v python tools/single_lambda.py --replicas 16 --steps 100 2>&1 | tail -1 | tee -a gpurun_out/row_variants.txt
done
