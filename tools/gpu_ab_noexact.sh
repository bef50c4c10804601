#!/bin/bash
# usage (under gpurun): tools/gpu_ab_noexact.sh lib.so ...  -- pair-kernel time only, K=20 steps, optional SDMB200_* env
for L in "$@"; do
  SDMB200_LIB=$PWD/$L python bench.py --no-cpu-baseline --no-single-lambda --steps 20 --warmup 3 --e2e-depth 1 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json"))
r=d["roofline"]
print("%-45s ms/step %.4f  pair_ms %.4f" % ("$L", d["ms_per_step"], r["kernel_ms"]))
PY
done
