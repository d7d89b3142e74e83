#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
for S in 8 16 24 48; do
timeout 900 python bench.py --scenes $S --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench18_s$S.json 2> /dev/null
python -c "
import json
d=json.loads(open('gpurun_out/bench18_s$S.json').read().strip().splitlines()[-1]); print('scenes per pass', $S, round(d['value'],3), round(d['ms_per_step'],1), 'e2e', round(d['e2e']['value'],3))
"
done
