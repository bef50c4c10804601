#!/bin/bash
# usage (under gpurun): tools/gpu_sweep.sh OUT.jsonl  -- cfg4 / cfg5: synthetic explicit-solvent boxes,
# roofline fraction versus size for 1-replica and 8-replica batches, plus cfg4 (50k atoms, 16 replicas)
OUT=${1:-gpurun_out/sweep.jsonl}
: > $OUT
for N in 5000 10000 20000 50000 100000 200000 500000; do
  for R in 1 8; do
    python bench.py --workload synthetic:$N --replicas $R --steps 10 --warmup 3 --no-cpu-baseline >> $OUT 2>> gpurun_out/sweep.err
  done
done
python bench.py --workload synthetic:50000 --replicas 16 --steps 10 --warmup 3 --no-cpu-baseline >> $OUT 2>> gpurun_out/sweep.err
python bench.py --workload cfg2 --replicas 1 --steps 20 --warmup 3 --no-cpu-baseline >> $OUT 2>> gpurun_out/sweep.err
python bench.py --workload cfg1 --replicas 16 --steps 20 --warmup 3 --no-cpu-baseline >> $OUT 2>> gpurun_out/sweep.err
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); r=d["roofline"]
    print("%-70s R=%2d  evals/s %9.1f  ms/step %8.3f  pair_ms %7.3f  frac %.4f  e2e %9.1f" % (d["config"]["workload"][:70], d["config"]["replicas_per_gpu"], d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
PY
tail -3 gpurun_out/sweep.err
