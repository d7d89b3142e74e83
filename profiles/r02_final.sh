#!/bin/bash
# round 2, last GPU call: full GPU suite, smoke(), the default bench line and the round-1-shaped pass with the final code
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1000 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log; tail -5 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --mode weak --scenes 8 --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/r02_bench_weak8.json 2> /dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu.json').read().strip().splitlines()[-1]); print(round(d['value'],3), round(d['ms_per_step'],1), d['e2e'], d['latency']['ms'], d['cpu_baseline']['value'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches'], d['clocks'])
for key in ('stage_rooflines','stage_rooflines_large_batch','stage_rooflines_large_batch_morton_order'):
    for k,v in d[key].items(): print(key[-14:], k[:50], round(v['ms']*1000,1),'us', round(v['frac'],3), v.get('points'))
w=json.loads(open('gpurun_out/r02_bench_weak8.json').read().strip().splitlines()[-1]); print('weak8', round(w['value'],3), round(w['ms_per_step'],1), w['roofline']['frac'], w['roofline']['gemm_family'], w['roofline']['phases_ms'])
"
