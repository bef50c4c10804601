#!/bin/bash
# usage (under gpurun): tools/gpu_quick.sh TAG  -- GPU parity tests + default bench, output under gpurun_out/
TAG=${1:-x}
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -6
python bench.py --no-cpu-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$TAG.json"))
r=d["roofline"]
print("evals/s %.0f  ms/step %.4f  pair_ms %.4f  frac %.4f  share %.3f  e2e %.0f" % (d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], r["kernel_share_of_step"], d["e2e"]["value"]))
PY
tail -3 gpurun_out/bench_$TAG.err
