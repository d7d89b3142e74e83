#!/bin/bash
# Round-2 evidence capture (one B200).  Outputs land in gpurun_out/ and are copied into profiles/ by hand; the tables
# are written by profiles/summarize.py.  Nothing printed under ncu is used as a bench value.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
set -x
( time timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu.log
tail -4 gpurun_out/r02_pytest_gpu.log
# bench lines: the default job, the round-1-shaped pass, deep features, the two stress configs (DMMA and tcgen05)
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
B="timeout 900 python bench.py --no-cpu-baseline --no-latency"
$B --mode weak --scenes 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_weak8.json 2> /dev/null
$B --workload c1 --mode weak --scenes 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1_8scenes.json 2> /dev/null
$B --workload c1_deep --mode weak --scenes 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_c1deep_8scenes.json 2> /dev/null
$B --workload c4 --total-scenes 2 --scenes 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_2scenes.json 2> /dev/null
GAPRO_GP_OZAKI=1 $B --workload c4 --total-scenes 2 --scenes 2 --steps 3 --warmup 3 > gpurun_out/r02_bench_c4_2scenes_tcgen05.json 2> /dev/null
$B --workload c5 --total-scenes 1 --scenes 1 --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/r02_bench_c5_1scene.json 2> /dev/null
GAPRO_GP_OZAKI=1 $B --workload c5 --total-scenes 1 --scenes 1 --steps 1 --warmup 1 --e2e-steps 1 > gpurun_out/r02_bench_c5_1scene_tcgen05.json 2> /dev/null
# launch list of the bench command (first 40000 launches: two passes of the job)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-latency > /tmp/b.log 2>&1
gzip -f gpurun_out/r02_launches.csv
# full captures of the top kernels (one stream group so that launches are in phase order)
export GAPRO_GP_STREAMS=1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:k_gemm -c 8 -o /tmp/gemm python tests/ncu_target.py c3 8 1 > /tmp/n1.log 2>&1
ncu -i /tmp/gemm.ncu-rep --page raw --csv > gpurun_out/r02_gemm_raw.csv 2>/dev/null
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_rl_|k_kgrad|k_build|k_colstats|k_grad_m|k_adam_small|k_small_fit" -c 16 -o /tmp/misc python tests/ncu_target.py c3 8 1 > /tmp/n2.log 2>&1
ncu -i /tmp/misc.ncu-rep --page raw --csv > gpurun_out/r02_gpmisc_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_occupancy|k_pool_feats|k_broadcast|k_densify|k_extent|k_compact|k_resolve|k_floor" -c 12 -o /tmp/scene python tests/ncu_target.py c3 8 1 > /tmp/n3.log 2>&1
ncu -i /tmp/scene.ncu-rep --page raw --csv > gpurun_out/r02_scene_raw.csv 2>/dev/null
GAPRO_OCCUPANCY=points timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_occ_|k_grid_build" -c 6 -o /tmp/occp python tests/ncu_target.py c3 8 1 > /tmp/n4.log 2>&1
ncu -i /tmp/occp.ncu-rep --page raw --csv > gpurun_out/r02_occ_points_raw.csv 2>/dev/null
GAPRO_GP_OZAKI=1 timeout 500 ncu --set full --clock-control none --import-source on -k regex:"k_oz_" -c 30 -o /tmp/oz python tests/ncu_target.py c4 1 1 > /tmp/n5.log 2>&1
ncu -i /tmp/oz.ncu-rep --page raw --csv > gpurun_out/r02_tcgen05_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_kgrad_wide|k_build" -c 4 -o /tmp/deep python tests/ncu_target.py c1_deep 2 1 > /tmp/n6.log 2>&1
ncu -i /tmp/deep.ncu-rep --page raw --csv > gpurun_out/r02_deep_raw.csv 2>/dev/null
gzip -f gpurun_out/r02_gemm_raw.csv gpurun_out/r02_gpmisc_raw.csv gpurun_out/r02_scene_raw.csv gpurun_out/r02_occ_points_raw.csv gpurun_out/r02_tcgen05_raw.csv gpurun_out/r02_deep_raw.csv
timeout 200 python tests/oz_probe.py > gpurun_out/r02_oz_probe.log 2>&1
ls -la gpurun_out | tail -30
