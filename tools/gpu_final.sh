#!/bin/bash
# usage (under gpurun): tools/gpu_final.sh TAG -- the evidence set of a round: GPU tests, default bench line, reference arm,
# ncu full capture of the pair kernel, launch list of a bench run, list-build cost, dynamics sweep.  Everything lands in gpurun_out/.
TAG=${1:-r02}
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 | tee gpurun_out/${TAG}_gputests.txt
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -2 gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -2 gpurun_out/${TAG}_bench_reference.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_row_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_pair_row -f \
  python bench.py --no-cpu-baseline --no-single-lambda --no-md-loop --no-cfg3 --no-sweep --no-elementwise --steps 2 --warmup 1 > gpurun_out/${TAG}_prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --no-cpu-baseline --no-single-lambda --no-md-loop --no-cfg3 --no-sweep --no-elementwise --steps 20 --warmup 1 > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 60 > gpurun_out/${TAG}_launches_summary.txt; head -12 gpurun_out/${TAG}_launches_summary.txt
(python tools/build_cost.py --replicas 16; python tools/build_cost.py --replicas 1; SDMB200_ASYNC_BUILD=0 SDMB200_KD_SORTS=1 python tools/build_cost.py --replicas 16; SDMB200_ASYNC_BUILD=0 SDMB200_KD_SORTS=1 python tools/build_cost.py --replicas 1) > gpurun_out/${TAG}_build_cost.txt 2>&1; cat gpurun_out/${TAG}_build_cost.txt
python tools/md_sweep.py --grid 0.06:20,0.10:20,0.14:40,0.20:40,0.24:40,0.28:60 > gpurun_out/${TAG}_md_sweep.jsonl 2>&1; tail -6 gpurun_out/${TAG}_md_sweep.jsonl | cut -c1-200
python tools/md_sweep.py --free-solute --grid 0.20:40 > gpurun_out/${TAG}_md_sweep_free_solute.jsonl 2>&1; tail -1 gpurun_out/${TAG}_md_sweep_free_solute.jsonl | cut -c1-200
for v in "SDMB200_ROW_GROUP=2" "SDMB200_ROW_CHUNK=8" "SDMB200_ROW_CHUNK=16" "SDMB200_PAIR_RESIDENT=16" "SDMB200_PAIR_RESIDENT=20" "SDMB200_ROW_LPT=0" "X=0"; do
  env $v python tools/single_lambda.py --replicas 16 --steps 100 2>&1 | tail -1
done > gpurun_out/${TAG}_row_variants.txt; cat gpurun_out/${TAG}_row_variants.txt
