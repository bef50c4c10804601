#!/bin/bash
# usage (under gpurun): tools/gpu_ab2.sh lib1.so lib2.so ...  -- resident leg only, once per library variant (twice each, interleaved)
for rep in 1 2; do
for L in "$@"; do
  SDMB200_LIB=$PWD/$L python bench.py --no-cpu-baseline --no-single-lambda --no-md-loop --no-cfg3 --no-sweep --no-elementwise --steps 60 --warmup 5 > gpurun_out/ab.json 2> gpurun_out/ab.err || tail -3 gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json"))
r=d["roofline"]
print("%-45s evals/s %.0f  ms/step %.4f  pair_ms %.4f  frac %.4f" % ("$L", d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"]))
PY
done
done
