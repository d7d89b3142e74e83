"""Development aid: the HBM-bound stage kernels timed alone (bench.py's stage_rooflines) on the bench batch and
on the same batch 8 times over."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                                          # noqa: E402
from gapro_b200 import _lib                                           # noqa: E402
from gapro_b200.engine import get_engine                              # noqa: E402
from gapro_b200.gen_ps import to_scene_inputs                         # noqa: E402

dev = torch.device("cuda:0")
eng = get_engine(dev)
lib = _lib.load()
stream = torch.cuda.current_stream(dev).cuda_stream
scenes = [to_scene_inputs(inp, dev, noise_seed=i) for i, inp in enumerate(bench.make_input(i, "c3") for i in range(8))]
flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flush = lambda: flush_buf.zero_()
for mult in (1, 8):
    eng.run(scenes * mult, stages_only=True, keep=True, thresh_spp_occu=0.999, training_iter=50)
    r = bench.stage_rooflines(eng, lib, stream, flush)
    print(mult, json.dumps({k: {"us": round(v["ms"] * 1e3, 1), "GB/s": round(v["achieved"]), "frac": round(v["frac"], 3),
                                "frac_gather": round(v.get("frac_of_random_gather_peak", 0), 3)}
                            for k, v in r.items()}))
