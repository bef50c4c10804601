#!/bin/bash
# usage (under gpurun): tools/gpu_e2e.sh G1 G2 ...  -- e2e leg with different numbers of replica groups
for G in "$@"; do
  python bench.py --no-cpu-baseline --no-single-lambda --e2e-groups $G > gpurun_out/e2e.json 2> gpurun_out/e2e.err
  python - <<PY
import json
d=json.load(open("gpurun_out/e2e.json"))
print("groups %d  resident %.0f evals/s (%.4f ms)   e2e %.0f evals/s (%.4f ms)" % ($G, d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
  tail -2 gpurun_out/e2e.err
done
