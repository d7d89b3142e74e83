"""Development aid: per-phase spans of one multi-stream GP stage (GAPRO_GP_TIMELINE=<file>)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gapro_b200 import _lib, synthetic                               # noqa: E402
from gapro_b200.engine import get_engine                             # noqa: E402
from gapro_b200.gen_ps import synthetic_inputs, to_scene_inputs      # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.txt"
os.environ["GAPRO_GP_TIMELINE"] = out
dev = torch.device("cuda:0")
eng = get_engine(dev)
lib = _lib.load()
scenes = [to_scene_inputs(synthetic_inputs(synthetic.make_scene(1000 + i, synthetic.c3_config(i))), dev, noise_seed=i)
          for i in range(8)]
for _ in range(2):
    eng.run(scenes, thresh_spp_occu=0.999, training_iter=50)
torch.cuda.synchronize()
lib.gapro_gp_set_profiling(1)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
eng.run(scenes, thresh_spp_occu=0.999, training_iter=50)
b.record()
torch.cuda.synchronize()
n = 32
ms = (ctypes.c_double * n)()
fa = (ctypes.c_double * n)()
fe = (ctypes.c_double * n)()
lib.gapro_gp_get_profile(ms, fa, fe, n)
lib.gapro_gp_set_profiling(0)
print("step ms", a.elapsed_time(b))
T = np.loadtxt(out)
names = lib.gapro_gp_phase_names().decode().split(",") if hasattr(lib.gapro_gp_phase_names, "restype") else None
slot, grp, t0, t1 = T[:, 0].astype(int), T[:, 1].astype(int), T[:, 2], T[:, 3]
print("span of GP stage ms", t1.max() - t0.min())
gemm = np.isin(slot, [2, 3, 5, 6, 8, 9, 11, 12])
# coverage: fraction of the stage during which >= 1 tile-product span is open, and mean number open
ev = sorted([(x, 1) for x in t0[gemm]] + [(x, -1) for x in t1[gemm]])
open_n, last, cover, area = 0, ev[0][0], 0.0, 0.0
for x, d in ev:
    if open_n > 0:
        cover += x - last
    area += open_n * (x - last)
    last = x
    open_n += d
print("time with >=1 tile product open: %.1f ms; mean open: %.2f" % (cover, area / max(cover, 1e-9)))
for s in sorted(set(slot)):
    sel = slot == s
    print("slot %2d: spans %5d  sum %.1f ms  mean %.3f ms" % (s, sel.sum(), (t1 - t0)[sel].sum(), (t1 - t0)[sel].mean()))
for g in sorted(set(grp)):
    sel = (grp == g)
    ch = sel & (slot == 1)
    print("group %d: chol mean %.3f ms, build mean %.3f, A mean %.3f, step period %.3f" % (
        g, (t1 - t0)[ch].mean(), (t1 - t0)[sel & (slot == 0)].mean(), (t1 - t0)[sel & (slot == 2)].mean(),
        np.diff(t0[sel & (slot == 0)]).mean()))
