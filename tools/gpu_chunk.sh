#!/bin/bash
# usage (under gpurun): tools/gpu_chunk.sh "8 12 16 24 32 auto" [replicas]  -- unit-length sweep, single-lambda leg and main leg
R=${2:-16}
for C in $1; do
  if [ "$C" = auto ]; then unset SDMB200_CHUNK; else export SDMB200_CHUNK=$C; fi
  python bench.py --no-cpu-baseline --replicas $R --steps 30 --warmup 5 > gpurun_out/chunk.json 2> gpurun_out/chunk.err
  python - <<PY
import json
d=json.load(open("gpurun_out/chunk.json"))
r=d["roofline"]; s=d.get("single_lambda") or {}
print("chunk %-5s R=%-3s ms/step %.4f pair_ms %.4f | single: ms/step %.4f evals/s %.0f" % ("$C", "$R", d["ms_per_step"], r["kernel_ms"], s.get("ms_per_step",0), s.get("value",0)))
PY
done
