#!/bin/bash
# usage (under gpurun): tools/gpu_rows.sh  -- parity tests with the row kernel, then the default bench for kernel variants
python -m pytest tests/test_gpu_parity.py tests/test_gpu_cluster.py tests/test_gpu_edge.py -m gpu -q --tb=short -x 2>&1 | tail -15
for V in "cluster 1" "rows 1" "rows 2"; do
  set -- $V
  SDMB200_PAIR_KERNEL=$1 SDMB200_ROW_GROUP=$2 python bench.py --no-cpu-baseline --steps 40 --warmup 5 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/ab.json"))
    r=d["roofline"]
    print("%-12s evals/s %.0f  ms/step %.4f  pair_ms %.4f  frac %.4f e2e %.0f" % ("$V", d["value"], d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]))
except Exception as ex:
    print("$V", "failed", ex); print(open("gpurun_out/ab.err").read()[-2000:])
PY
done
