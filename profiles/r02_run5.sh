#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python tests/oz_gp_probe.py > gpurun_out/oz_gp_probe2.log 2>&1
echo "probe rc=$?" >> gpurun_out/oz_gp_probe2.log
for S in 7 8; do
GAPRO_GP_OZAKI=1 GAPRO_GP_OZAKI_S=$S GAPRO_GP_OZAKI_MIN_M=1024 timeout 600 python bench.py --workload c4 --total-scenes 2 --scenes 2 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench5_c4_S$S.json 2> gpurun_out/bench5_c4_S$S.err
done
cat gpurun_out/oz_gp_probe2.log
for f in gpurun_out/bench5_*.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value'],3), round(d['ms_per_step'],1), d['roofline']['phases_ms'])
" 2>&1 | tail -1; done
