#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python bench.py --mode weak --scenes 8 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench20_weak8.json 2> gpurun_out/bench20.err; tail -3 gpurun_out/bench20.err
python -c "
import json
d=json.loads(open('gpurun_out/bench20_weak8.json').read().strip().splitlines()[-1])
for key in ('stage_rooflines','stage_rooflines_large_batch','stage_rooflines_large_batch_morton_order'):
    for k,v in d[key].items(): print(key[-14:], k[:60], round(v['ms']*1000,1),'us', round(v['frac'],3))
"
