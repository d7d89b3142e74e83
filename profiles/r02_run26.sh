#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1000 python -m pytest tests -m gpu -q -k "stages or containment or stress or point_level or ensemble or scene" ) > gpurun_out/pytest_gpu26.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu26.log; tail -5 gpurun_out/pytest_gpu26.log
timeout 900 python bench.py --mode weak --scenes 8 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench26_weak8.json 2> /dev/null
python -c "
import json
d=json.loads(open('gpurun_out/bench26_weak8.json').read().strip().splitlines()[-1])
for key in ('stage_rooflines','stage_rooflines_large_batch','stage_rooflines_large_batch_morton_order'):
    for k,v in d[key].items():
        if 'pooling' in k: print(key[-14:], k[:40], round(v['ms']*1000,1),'us', round(v['frac'],3))
"
