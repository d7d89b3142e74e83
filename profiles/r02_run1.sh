#!/bin/bash
# round 2, GPU call 1: parity suite + first bench lines (strong default, round-1-shaped weak, c4)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_strong32.json 2> gpurun_out/bench_strong32.err
python bench.py --mode weak --scenes 8 --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_weak8.json 2> gpurun_out/bench_weak8.err
python bench.py --workload c4 --total-scenes 2 --scenes 2 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err
tail -5 gpurun_out/pytest_gpu.log
