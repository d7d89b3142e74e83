#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu9.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu9.log
tail -12 gpurun_out/pytest_gpu9.log
rm -f gpurun_out/latency_probe.log; timeout 300 python tests/latency_probe.py >> gpurun_out/latency_probe.log 2>&1; cat gpurun_out/latency_probe.log
timeout 600 python bench.py --workload c1_deep --mode weak --scenes 8 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench9_c1deep.json 2> gpurun_out/bench9_c1deep.err
python -c "
import json
d=json.loads(open('gpurun_out/bench9_c1deep.json').read().strip().splitlines()[-1]); print(round(d['value'],3), round(d['ms_per_step'],1), d['roofline']['phases_ms'])
"
