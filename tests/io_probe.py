"""Host I/O pipeline measurement: disk -> host arrays (load_scene: torch.load of the scene /
superpoint files, axis alignment, boxes) with the CLI's worker-thread prefetch, and host arrays -> disk
(save_pseudo_labels), on a reference-shaped dataset/scannetv2 tree of synthetic c1 scenes.
    python tests/io_probe.py [n_scenes] [workers]"""
import json
import os
import os.path as osp
import sys
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, ROOT)
from gapro_b200 import gen_ps, synthetic  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    workers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    with tempfile.TemporaryDirectory() as tmp:
        root = osp.join(tmp, "dataset", "scannetv2")
        for d in ("train", "superpoints", "scans_transform", "out"):
            os.makedirs(osp.join(root, d))
        nbytes = 0
        for i in range(n):
            sc = synthetic.make_scene(1000 + i, "c1")
            scan = "scene%04d_00" % i
            f1 = osp.join(root, "train", scan + "_inst_nostuff.pth")
            torch.save((sc.xyz_raw, sc.rgb, sc.sem, sc.inst), f1)
            f2 = osp.join(root, "superpoints", scan + ".pth")
            torch.save(sc.spp, f2)
            os.makedirs(osp.join(root, "scans_transform", scan))
            with open(osp.join(root, "scans_transform", scan, scan + ".txt"), "w") as f:
                f.write("axisAlignment = " + " ".join(repr(float(x)) for x in sc.axis_align.reshape(-1)) + "\n")
            nbytes += os.path.getsize(f1) + os.path.getsize(f2)
        items = [(osp.join(root, "train", "scene%04d_00_inst_nostuff.pth" % i), "scene%04d_00" % i) for i in range(n)]
        res = {}
        for w in (1, workers):
            pool = ThreadPoolExecutor(max_workers=w)
            t0 = time.perf_counter()
            loaded = list(pool.map(lambda it: gen_ps.load_scene(it[0], it[1], data_root=root), items))
            dt = time.perf_counter() - t0
            res[f"load_scene_{w}_threads"] = {"scenes_per_s": n / dt, "MB_per_s": nbytes / dt / 1e6}
        N, S = len(loaded[0][0]["xyz"]), len(np.unique(loaded[0][0]["spp"]))
        out = (torch.zeros(N, dtype=torch.int32), torch.zeros(N, dtype=torch.int32), torch.ones(N), torch.zeros(S), torch.ones(S))
        t0 = time.perf_counter()
        for i in range(n):
            gen_ps.save_pseudo_labels(osp.join(root, "out", "scene%04d_00.pth" % i), out)
        dt = time.perf_counter() - t0
        sz = os.path.getsize(osp.join(root, "out", "scene0000_00.pth"))
        res["save_pseudo_labels"] = {"scenes_per_s": n / dt, "MB_per_s": n * sz / dt / 1e6}
        res["scene_file_MB"] = nbytes / n / 1e6
        res["label_file_MB"] = sz / 1e6
        res["host_cores"] = len(os.sched_getaffinity(0))
        print(json.dumps(res))


if __name__ == "__main__":
    main()
