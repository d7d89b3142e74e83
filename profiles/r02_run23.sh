#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1000 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -8 gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --no-latency --workload c1_deep --mode weak --scenes 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1deep_8scenes.json 2> /dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_c1deep_8scenes.json').read().strip().splitlines()[-1]); print(round(d['value'],3), round(d['ms_per_step'],1), d['roofline']['phases_ms'])
"
