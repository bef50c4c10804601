#!/bin/bash
# usage (under gpurun): tools/gpu_env_ab.sh VAR v1 v2 ...  -- default bench once per value of an environment knob
VAR=$1; shift
for V in "$@"; do
  env $VAR=$V python bench.py --no-cpu-baseline --no-single-lambda --steps 30 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab.json")); r=d["roofline"]
print("$VAR=$V  evals/s %.0f  ms/step %.4f  pair_ms %.4f  frac %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"]))
PY
done
