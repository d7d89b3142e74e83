#!/bin/bash
# round 2, final verification: the full GPU suite, smoke(), the default bench command as the driver runs it
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 ) > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -16 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
timeout 1200 python bench.py --gpus 1 --steps 3 --warmup 3 > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err
timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu_final.json').read().strip().splitlines()[-1]); print(round(d['value'],3), round(d['ms_per_step'],1), d['e2e'], d['latency']['ms'], d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])
r=json.loads(open('gpurun_out/r02_bench_reference.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['ms_per_step'], r['config']==d['config'], r['scene_fraction_per_step'])
"
