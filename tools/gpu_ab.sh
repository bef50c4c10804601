#!/bin/bash
# usage (under gpurun): tools/gpu_ab.sh lib1.so lib2.so ...  -- default bench once per library variant
for L in "$@"; do
  SDMB200_LIB=$PWD/$L python bench.py --no-cpu-baseline --steps 40 --warmup 5 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json"))
r=d["roofline"]
print("%-45s evals/s %.0f  ms/step %.4f  pair_ms %.4f  frac %.4f" % ("$L", d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"]))
PY
done
