# Round-1 evidence capture (one B200).  Outputs land in gpurun_out/ and are copied into profiles/ by hand;
# tables are then written by profiles/summarize.py.  Pass "gp" to skip the element-wise / scene captures.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r01_bench_1gpu.json 2> gpurun_out/bench_err.log
tail -c 600 gpurun_out/bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r01_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /tmp/b.log 2>&1
gzip -f gpurun_out/r01_launches.csv
export GAPRO_GP_STREAMS=1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_gemm -c 8 -o /tmp/gemm python tests/ncu_target.py c3 8 1 > /tmp/n1.log 2>&1
ncu -i /tmp/gemm.ncu-rep --page raw --csv > gpurun_out/r01_gemm_raw.csv 2>/dev/null
gzip -f gpurun_out/r01_gemm_raw.csv
if [ "$1" != "gp" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_rl_|k_kgrad|k_build|k_colstats|k_grad_m|k_adam_small" -c 14 -o /tmp/misc python tests/ncu_target.py c3 8 1 > /tmp/n2.log 2>&1
ncu -i /tmp/misc.ncu-rep --page raw --csv > gpurun_out/r01_gpmisc_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_occupancy|k_pool_feats|k_broadcast|k_densify|k_extent|k_compact|k_resolve|k_floor" -c 12 -o /tmp/scene python tests/ncu_target.py c3 8 1 > /tmp/n3.log 2>&1
ncu -i /tmp/scene.ncu-rep --page raw --csv > gpurun_out/r01_scene_raw.csv 2>/dev/null
gzip -f gpurun_out/r01_gpmisc_raw.csv gpurun_out/r01_scene_raw.csv
fi
ls -la gpurun_out
