#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "stages or containment or stress" ) > gpurun_out/pytest_gpu14.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu14.log
tail -5 gpurun_out/pytest_gpu14.log
timeout 600 python bench.py --mode weak --scenes 8 --steps 1 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench14_weak8.json 2> /dev/null
python -c "
import json
d=json.loads(open('gpurun_out/bench14_weak8.json').read().strip().splitlines()[-1])
for k,v in d['stage_rooflines'].items(): print('1x', k[:70], round(v['ms']*1000,1),'us', round(v['frac'],3))
for k,v in d['stage_rooflines_large_batch'].items(): print('8x', k[:70], round(v['ms']*1000,1),'us', round(v['frac'],3))
"
export GAPRO_GP_STREAMS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_kgrad_wide|k_build_wide" -c 4 -o /tmp/deep python tests/ncu_target.py c1_deep 2 1 > /tmp/n6.log 2>&1; tail -2 /tmp/n6.log
ncu -i /tmp/deep.ncu-rep --page raw --csv > gpurun_out/r02_deep_raw.csv 2>/dev/null
gzip -f gpurun_out/r02_deep_raw.csv; ls -la gpurun_out/r02_deep_raw.csv.gz
