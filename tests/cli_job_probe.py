"""Development probe (GPU): the whole job through the CLI on a reference-shaped dataset tree of synthetic configs[2]
scenes - disk -> loader threads -> pinned H2D -> hot path -> D2H -> torch.save - wall-clock scenes/s."""
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gapro_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tmp = tempfile.mkdtemp()
root = os.path.join(tmp, "dataset", "scannetv2")
for d in ("train", "superpoints", "scans_transform"):
    os.makedirs(os.path.join(root, d))
for i in range(n):
    sc = synthetic.make_scene(1000 + i, synthetic.c3_config(i))
    scan = "scene%04d_00" % i
    torch.save((sc.xyz_raw, sc.rgb, sc.sem, sc.inst), os.path.join(root, "train", scan + "_inst_nostuff.pth"))
    torch.save(sc.spp, os.path.join(root, "superpoints", scan + ".pth"))
    os.makedirs(os.path.join(root, "scans_transform", scan))
    with open(os.path.join(root, "scans_transform", scan, scan + ".txt"), "w") as f:
        f.write("axisAlignment = " + " ".join(repr(float(x)) for x in sc.axis_align.reshape(-1)) + "\n")
env = dict(os.environ, PYTHONPATH=ROOT)
for extra in ([], ["--device_boxes"]):
    out = os.path.join(tmp, "out" + ("_dev" if extra else ""))
    t0 = time.perf_counter()
    r = subprocess.run([sys.executable, "-m", "gapro_b200.gen_ps", "--save_folder", out, "--seed", "1"] + extra, cwd=tmp,
                       env=env, capture_output=True, text=True)
    dt = time.perf_counter() - t0
    done = len([f for f in os.listdir(out) if f.endswith(".pth")]) if os.path.isdir(out) else 0
    print(f"CLI {' '.join(extra) or '(host boxes)'}: {done}/{n} scenes in {dt:.1f} s wall incl. interpreter + CUDA start-up "
          f"= {done / dt:.2f} scenes/s; rc={r.returncode}; {r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr[-300:]}",
          flush=True)
