#!/bin/bash
# usage (under gpurun): tools/gpu_quick2.sh -- GPU tests, build cost, short bench
python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -15
python tools/build_cost.py --replicas 16 2>&1 | tee gpurun_out/build_cost.txt
python tools/build_cost.py --replicas 1 2>&1 | tee -a gpurun_out/build_cost.txt
python bench.py --no-cpu-baseline --no-sweep --no-elementwise --no-cfg3 > gpurun_out/q.json 2> gpurun_out/q.err; tail -3 gpurun_out/q.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/q.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'pair',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value']),'f32',round(d['e2e']['f32_io']['value']),'single',d['single_lambda']['ms_per_step'],d['single_lambda']['ms_per_step_between_list_builds'],'md',d['md_loop']['ms_per_step'])
P
