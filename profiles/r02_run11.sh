#!/bin/bash
# round 2, GPU call 11 (one B200): final default bench line (48-scene job), deep-feature captures, digit-plane probe with S=8
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
timeout 600 python bench.py --no-cpu-baseline --no-latency --workload c1_deep --mode weak --scenes 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1deep_8scenes.json 2> /dev/null
export GAPRO_GP_STREAMS=1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_kgrad_wide|k_build_wide" -c 4 -o /tmp/deep python tests/ncu_target.py c1_deep 2 1 > /tmp/n6.log 2>&1
ncu -i /tmp/deep.ncu-rep --page raw --csv > gpurun_out/r02_deep_raw.csv 2>/dev/null
gzip -f gpurun_out/r02_deep_raw.csv
unset GAPRO_GP_STREAMS
timeout 300 python tests/oz_probe.py > gpurun_out/r02_oz_probe.log 2>&1
python -c "
import json
for f in ('r02_bench_1gpu','r02_bench_c1deep_8scenes'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, round(d['value'],3), round(d['ms_per_step'],1), d['e2e']['value'], d.get('latency',{}) and d['latency'].get('ms'), d['cpu_baseline'] and d['cpu_baseline']['value'], d['roofline']['phases_ms'])
"
grep "4096\|cuBLAS" gpurun_out/r02_oz_probe.log
