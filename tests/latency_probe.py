"""Development probe (GPU): one configs[1] scene through the reference's per-scene call - wall-clock latency and
launch count (GAPRO_GP_STREAMS selects the number of stream groups)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gapro_b200 import synthetic  # noqa: E402
from gapro_b200.engine import get_engine  # noqa: E402
from gapro_b200.gen_ps import synthetic_inputs, to_scene_inputs  # noqa: E402
from gapro_b200.gen_ps_utils import gen_pseudo_label_gaussian_process  # noqa: E402

dev = torch.device("cuda:0")
sc = to_scene_inputs(synthetic_inputs(synthetic.make_scene(1000, "c1")), dev, noise_seed=1)
ts = []
for r in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    gen_pseudo_label_gaussian_process(sc.coords_float, sc.mask_feats, sc.spp, sc.instance_cls, sc.instance_box,
                                      sc.instance_box_volume, sc.wall_box, sc.wall_box_volume, instance_classes=18,
                                      dataset_name="scannetv2", ground_h=0.1, training_iter=50, thresh_spp_occu=0.999,
                                      noise_seed=1)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
st = get_engine(dev).last_stats
print(f"streams={os.environ.get('GAPRO_GP_STREAMS', 'default')}: c1 scene {np.median(ts[1:]):.1f} ms "
      f"(runs {[round(t, 1) for t in ts]}), {st['n_regions']} regions, {st['launches']} launches")
