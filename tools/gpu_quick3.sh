#!/bin/bash
python tools/md_sweep.py --steps 200 --grid 0.06:20,0.10:20,0.10:40,0.14:40,0.20:40 2>&1 | tee gpurun_out/md_sweep_frozen.jsonl | cut -c1-260
python bench.py --no-cpu-baseline --no-sweep --no-elementwise > gpurun_out/q.json 2> gpurun_out/q.err; tail -3 gpurun_out/q.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/q.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms',round(d['ms_per_step'],4),'pair',round(d['roofline']['kernel_ms'],4),'frac',round(d['roofline']['frac'],4),'e2e',round(d['e2e']['value']),'flush',round(d['e2e']['with_flush_kernel']['value']),'f32',round(d['e2e']['f32_io']['value']),'single',d['single_lambda']['ms_per_step'],d['single_lambda']['ms_per_step_between_list_builds'],'md',d['md_loop']['ms_per_step'],d['md_loop']['list_builds'],'cfg3',d['cfg3']['ms_per_step'])
P
