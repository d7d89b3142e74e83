#!/bin/bash
# round 2, GPU call 2: parity suite with the point-order occupancy, stage rooflines, first tcgen05 digit-plane probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
timeout 600 python bench.py --mode weak --scenes 8 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench2_weak8.json 2> gpurun_out/bench2_weak8.err
timeout 180 python tests/oz_probe.py > gpurun_out/oz_probe.log 2>&1
echo "oz_probe rc=$?" >> gpurun_out/oz_probe.log
nvidia-smi > gpurun_out/smi_after.txt 2>&1
tail -3 gpurun_out/pytest_gpu2.log; tail -30 gpurun_out/oz_probe.log
