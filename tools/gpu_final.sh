#!/bin/bash
# usage (under gpurun): tools/gpu_final.sh TAG -- the evidence set of a round: GPU tests, default bench line, reference arm,
# ncu full capture of the pair kernel, launch list of a bench run, microbenchmarks.  Everything lands in gpurun_out/.
TAG=${1:-r02}
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 | tee gpurun_out/${TAG}_gputests.txt
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -2 gpurun_out/${TAG}_bench_n1.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -2 gpurun_out/${TAG}_bench_reference.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_row_kernel -s 3 -c 1 -o gpurun_out/${TAG}_prof_pair_row -f \
  python bench.py --no-cpu-baseline --no-single-lambda --no-md-loop --no-cfg3 --no-sweep --no-elementwise --steps 2 --warmup 1 > gpurun_out/${TAG}_prof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
  python bench.py --no-cpu-baseline --no-single-lambda --no-md-loop --no-cfg3 --no-sweep --no-elementwise --steps 20 --warmup 1 > gpurun_out/${TAG}_launches.log 2>&1
python tools/launch_summary.py gpurun_out/${TAG}_launches.csv 60 > gpurun_out/${TAG}_launches_summary.txt; head -12 gpurun_out/${TAG}_launches_summary.txt
tools/microbench/red_rates > gpurun_out/${TAG}_microbench_red_rates.txt 2>&1
tools/microbench/fp32_rates > gpurun_out/${TAG}_microbench_fp32_rates.txt 2>&1
python tools/md_sweep.py --grid 0.06:20,0.10:20,0.14:40,0.20:40,0.24:40 > gpurun_out/${TAG}_md_sweep.jsonl 2>&1; tail -5 gpurun_out/${TAG}_md_sweep.jsonl | cut -c1-220
