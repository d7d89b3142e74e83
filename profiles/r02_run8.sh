#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu8.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu8.log
tail -25 gpurun_out/pytest_gpu8.log
rm -f gpurun_out/latency_probe.log
timeout 300 python tests/latency_probe.py >> gpurun_out/latency_probe.log 2>&1
GAPRO_GP_SMALL=0 timeout 300 python tests/latency_probe.py >> gpurun_out/latency_probe.log 2>&1
cat gpurun_out/latency_probe.log
B="timeout 600 python bench.py --mode weak --scenes 8 --steps 3 --warmup 3 --no-cpu-baseline --no-latency"
$B > gpurun_out/bench8_weak8.json 2> gpurun_out/bench8_weak8.err
GAPRO_GP_SMALL=0 $B > gpurun_out/bench8_weak8_nosmall.json 2> gpurun_out/bench8_weak8_nosmall.err
for f in gpurun_out/bench8_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['gpu_launches'], d['roofline']['phases_ms'])
" 2>&1 | tail -1; done
