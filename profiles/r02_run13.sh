#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q -k "tcgen05 or golden or small_region or deep" ) > gpurun_out/pytest_gpu13.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu13.log
tail -6 gpurun_out/pytest_gpu13.log
GAPRO_GP_OZAKI=1 timeout 600 python bench.py --no-cpu-baseline --no-latency --workload c4 --total-scenes 2 --scenes 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_2scenes_tcgen05.json 2> /dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_c4_2scenes_tcgen05.json').read().strip().splitlines()[-1]); print(round(d['value'],3), round(d['ms_per_step'],1), d['roofline']['phases_ms'])
"
export GAPRO_GP_STREAMS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_kgrad_wide -c 2 -o /tmp/deep1 python tests/ncu_target.py c1_deep 2 1 > /tmp/n6.log 2>&1; tail -3 /tmp/n6.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_build_wide -c 2 -o /tmp/deep2 python tests/ncu_target.py c1_deep 2 1 > /tmp/n7.log 2>&1; tail -3 /tmp/n7.log
ncu -i /tmp/deep1.ncu-rep --page raw --csv > gpurun_out/r02_deep_kgrad_raw.csv 2>/dev/null
ncu -i /tmp/deep2.ncu-rep --page raw --csv > gpurun_out/r02_deep_build_raw.csv 2>/dev/null
GAPRO_GP_OZAKI=1 timeout 300 ncu --set full --clock-control none -k regex:"k_oz_vecmax|k_oz_zero" -c 4 -o /tmp/ozs python tests/ncu_target.py c4 1 1 > /tmp/n8.log 2>&1
ncu -i /tmp/ozs.ncu-rep --page raw --csv > gpurun_out/r02_tcgen05_scale_raw.csv 2>/dev/null
gzip -f gpurun_out/r02_deep_kgrad_raw.csv gpurun_out/r02_deep_build_raw.csv gpurun_out/r02_tcgen05_scale_raw.csv
ls -la gpurun_out/*.gz | tail -4
