#!/bin/bash
# usage (under gpurun): tools/gpu_env_sweep.sh VAR v1 v2 ...  -- default bench once per value of an environment knob
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v python bench.py --no-cpu-baseline --no-single-lambda --steps 40 --warmup 5 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    r=d["roofline"]
    print("%-28s evals/s %.0f  ms/step %.4f  pair_ms %.4f  frac %.4f e2e %.0f" % ("$VAR=$v", d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
except Exception as ex:
    print("$VAR=$v", "failed", ex); print(open("gpurun_out/ab.err").read()[-1500:])
PY
done
