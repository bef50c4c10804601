#!/bin/bash
# usage (under gpurun): tools/gpu_side_sweep.sh "ENV1=.. ENV2=.." ...  -- resident leg + replay time per environment setting
for e in "$@"; do
  r=$(env $e python tools/build_cost.py --replicas 16 2>&1 | grep replay | tail -1 | awk '{print $8}')
  env $e python bench.py --no-cpu-baseline --no-single-lambda --no-md-loop --no-cfg3 --no-sweep --no-elementwise --steps 60 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-60s replay $r  value %d  ms/step %.4f  e2e %d' % ('$e', d['value'], d['ms_per_step'], d['e2e']['value']))"
done
