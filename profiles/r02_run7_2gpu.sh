#!/bin/bash
# round 2, GPU call 7 (2 GPUs): the sharded job through the CLI and the bench at N=2 vs N=1
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
R=$(pwd)
T=$(mktemp -d)
python - <<PY
import os, sys, torch
sys.path.insert(0, "$R")
from gapro_b200 import synthetic
root = os.path.join("$T", "dataset", "scannetv2")
for d in ("train", "superpoints", "scans_transform"):
    os.makedirs(os.path.join(root, d))
for i in range(14):
    sc = synthetic.make_scene(300 + i, "small" if i % 3 else "tiny")
    scan = "scene%04d_00" % i
    torch.save((sc.xyz_raw, sc.rgb, sc.sem, sc.inst), os.path.join(root, "train", scan + "_inst_nostuff.pth"))
    torch.save(sc.spp, os.path.join(root, "superpoints", scan + ".pth"))
    os.makedirs(os.path.join(root, "scans_transform", scan))
    open(os.path.join(root, "scans_transform", scan, scan + ".txt"), "w").write(
        "axisAlignment = " + " ".join(repr(float(x)) for x in sc.axis_align.reshape(-1)) + "\n")
# one broken scene: no labelled instance -> must be reported, not fatal
sc = synthetic.make_scene(399, "tiny")
scan = "scene0099_00"
torch.save((sc.xyz_raw, sc.rgb, sc.sem, sc.inst * 0 - 100.0), os.path.join(root, "train", scan + "_inst_nostuff.pth"))
torch.save(sc.spp, os.path.join(root, "superpoints", scan + ".pth"))
os.makedirs(os.path.join(root, "scans_transform", scan))
open(os.path.join(root, "scans_transform", scan, scan + ".txt"), "w").write(
    "axisAlignment = " + " ".join(repr(float(x)) for x in sc.axis_align.reshape(-1)) + "\n")
PY
( cd $T && PYTHONPATH=$R timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 -m gapro_b200.gen_ps --seed 7 --eval_pslabel --batch_scenes 4 ) > gpurun_out/cli_2gpu.log 2>&1
echo "cli rc=$? outputs: $(ls $T/dataset/scannetv2/gaussian_process_kl_pseudo_labels | wc -l)" >> gpurun_out/cli_2gpu.log
# resume: second run must find nothing to do
( cd $T && PYTHONPATH=$R timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 -m gapro_b200.gen_ps --seed 7 --batch_scenes 4 ) >> gpurun_out/cli_2gpu.log 2>&1
echo "resume rc=$?" >> gpurun_out/cli_2gpu.log
tail -12 gpurun_out/cli_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench7_strong32_n2.json 2> gpurun_out/bench7_strong32_n2.err
timeout 900 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/bench7_strong32_n1.json 2> gpurun_out/bench7_strong32_n1.err
for f in gpurun_out/bench7_*.json; do python -c "
import json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['n_gpus'], round(d['value'],3), round(d['ms_per_step'],1), d['balance'], d['e2e']['value'])
" 2>&1 | tail -1; done
tail -3 gpurun_out/bench7_strong32_n2.err
