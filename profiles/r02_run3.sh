#!/bin/bash
# round 2, GPU call 3: tcgen05 digit-plane path inside the GP fit - parity suite, accuracy probe, A/B bench lines
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tests/oz_gp_probe.py > gpurun_out/oz_gp_probe.log 2>&1
echo "probe rc=$?" >> gpurun_out/oz_gp_probe.log
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu3.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu3.log
B="timeout 600 python bench.py --mode weak --scenes 8 --steps 3 --warmup 3 --no-cpu-baseline --no-latency"
$B > gpurun_out/bench3_weak8_oz1024.json 2> gpurun_out/bench3_weak8_oz1024.err
GAPRO_GP_OZAKI=0 $B > gpurun_out/bench3_weak8_dmma.json 2> gpurun_out/bench3_weak8_dmma.err
GAPRO_GP_OZAKI_MIN_M=512 $B > gpurun_out/bench3_weak8_oz512.json 2> gpurun_out/bench3_weak8_oz512.err
GAPRO_GP_OZAKI_MIN_M=768 $B > gpurun_out/bench3_weak8_oz768.json 2> gpurun_out/bench3_weak8_oz768.err
timeout 600 python bench.py --workload c4 --total-scenes 2 --scenes 2 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench3_c4.json 2> gpurun_out/bench3_c4.err
cat gpurun_out/oz_gp_probe.log; tail -15 gpurun_out/pytest_gpu3.log
for f in gpurun_out/bench3_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['roofline']['phases_ms'])
" 2>&1 | tail -1; done
