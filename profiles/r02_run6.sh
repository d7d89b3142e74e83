#!/bin/bash
# round 2, GPU call 6: full parity suite, single-scene latency vs stream groups, e2e breakdown, ncu launch list
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu6.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu6.log
tail -5 gpurun_out/pytest_gpu6.log
for G in 1 2 4; do
GAPRO_GP_STREAMS=$G timeout 300 python tests/latency_probe.py >> gpurun_out/latency_probe.log 2>&1
done
timeout 300 python tests/latency_probe.py default >> gpurun_out/latency_probe.log 2>&1
cat gpurun_out/latency_probe.log
timeout 600 python bench.py --mode weak --scenes 8 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench6_weak8.json 2> gpurun_out/bench6_weak8.err
python -c "
import json
d=json.loads(open('gpurun_out/bench6_weak8.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e'], d['latency'])
for k,v in d['stage_rooflines'].items(): print('1x', k[:60], round(v['ms']*1000,1),'us', round(v['frac'],3))
for k,v in d['stage_rooflines_large_batch'].items(): print('8x', k[:60], round(v['ms']*1000,1),'us', round(v['frac'],3))
"
