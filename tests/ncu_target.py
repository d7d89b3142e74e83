"""Short single-pass workload for ncu captures (development aid):
    ncu --set full -k regex:k_gemm -s 40 -c 3 -o gpurun_out/prof python tests/ncu_target.py [config] [n_scenes]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gapro_b200 import synthetic                                     # noqa: E402
from gapro_b200.engine import get_engine                             # noqa: E402
from gapro_b200.gen_ps import synthetic_inputs, to_scene_inputs      # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
eng = get_engine(dev)
def _cfg(i):
    return synthetic.c3_config(i) if cfg == "c3" else cfg


deep = cfg == "c1_deep"
scenes = [to_scene_inputs(synthetic_inputs(synthetic.make_scene(1000 + i, _cfg(i)), use_deepfeat=deep), dev, noise_seed=i)
          for i in range(n)]
eng.run(scenes, thresh_spp_occu=0.999, training_iter=iters)
torch.cuda.synchronize()
print("done", eng.last_stats["n_regions"], eng.last_stats["launches"])
