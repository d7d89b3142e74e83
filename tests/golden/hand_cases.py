"""Hand-built geometries that drive every branch of the pair loop of gen_pseudo_label_gaussian_process
(gen_ps_utils.py:385-446): nesting both ways, the IoU >= 0.6 skip with the smallest-volume fallback, three
mutually overlapping boxes (strict-'<' merge in event order), a box without superpoints of its own, walls.
Shared by tests/golden/make_ref_golden.py (which runs the REFERENCE on them) and the parity tests."""
import numpy as np


def grid_points(lo, hi, n=4):
    ax = [np.linspace(lo[d] + 0.02, hi[d] - 0.02, n) for d in range(3)]
    return np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)


def _case(points, boxes, cls=None, walls=None, group=1, seed=0):
    """points -> (xyz f64, feats f32[6], spp i64 with gaps, cls i64, boxes f32, volumes f32, walls, wall volumes).
    One extra point far below pins the floor slab (gen_ps_utils.py:317-326) away from the geometry.
    `group` consecutive points share a superpoint; colours are a fixed pseudo-random pattern."""
    pts = np.concatenate([np.asarray(points, np.float64), [[points[:, 0].mean(), points[:, 1].mean(), -5.0]]])
    n = len(pts)
    spp = (np.arange(n) // group) * 7 + 3
    rng = np.random.default_rng([77, seed])
    rgb = rng.uniform(-1, 1, (n, 3))
    feats = np.concatenate([pts, rgb], 1).astype(np.float32)
    boxes = np.asarray(boxes, dtype=np.float32)
    vol = np.prod(np.clip(boxes[:, 3:] - boxes[:, :3], 0, None), axis=1).astype(np.float32)
    cls = np.arange(len(boxes), dtype=np.int64) if cls is None else np.asarray(cls, dtype=np.int64)
    if walls is None:
        wb, wv = [], []
    else:
        wb = np.asarray(walls, np.float32)
        wv = np.prod(wb[:, 3:] - wb[:, :3], axis=1).astype(np.float32)
    return dict(xyz=pts, mask_feats=feats, spp=spp.astype(np.int64), instance_cls=cls, instance_box=boxes,
                instance_box_volume=vol, wall_box=wb, wall_volume=wv)


def hand_cases():
    out = {}
    big, small = [0, 0, 1, 4, 4, 3], [1, 1, 1.5, 2, 2, 2.5]
    nested_pts = np.concatenate([grid_points(big[:3], big[3:], 6), grid_points(small[:3], small[3:], 3)])
    out["nested_b1_in_b2"] = _case(nested_pts, [small, big], cls=[1, 2], seed=1)
    out["nested_b2_in_b1"] = _case(nested_pts, [big, small], cls=[1, 2], seed=2)
    a, b = [0, 0, 1, 2, 2, 3], [0.15, 0, 1, 2.2, 2, 3]
    out["high_iou_skip"] = _case(np.concatenate([grid_points(a[:3], a[3:], 5), grid_points(b[:3], b[3:], 5)]), [a, b], seed=3)
    a, b, c = [0, 0, 1, 2, 2, 2], [1.5, 0, 1, 3.5, 2, 2], [1.2, 1.5, 1, 2.5, 3.5, 2]
    three = np.concatenate([grid_points(a[:3], a[3:], 6), grid_points(b[:3], b[3:], 6), grid_points(c[:3], c[3:], 6)])
    out["three_way_merge"] = _case(three, [a, b, c], cls=[3, 3, 7], seed=4)
    out["three_way_grouped"] = _case(three, [a, b, c], cls=[0, 5, 17], group=3, seed=5)
    a, b, c = [0, 0, 1, 2, 1, 2], [1.0, 0, 1, 3.0, 1, 2], [2.0, 0, 1, 4, 1, 2]
    out["no_exclusive_spp"] = _case(np.concatenate([grid_points([0, 0, 1], [1.9, 1, 2], 5), grid_points([2.05, 0, 1], [4, 1, 2], 5)]),
                                    [a, b, c], seed=6)
    a, b = [0, 0, 1, 2, 2, 2], [1.5, 0, 1, 3.5, 2, 2]
    wall = [[-0.2, -0.2, 0.9, 0.3, 2.2, 2.1]]
    out["with_wall"] = _case(np.concatenate([grid_points(a[:3], a[3:], 6), grid_points(b[:3], b[3:], 6)]), [a, b],
                             cls=[2, 9], walls=wall, seed=7)
    return out


CLI_SCANS = ["scene0000_00", "scene0001_00", "scene0002_01"]
CLI_SEED = 7


def cli_noise_seed(scan, seed=CLI_SEED):
    """noise seed the product CLI derives for a scan from --seed (gapro_b200/gen_ps.py)."""
    import zlib
    return (zlib.crc32(scan.encode()) ^ seed) & 0x7FFFFFFF


def write_cli_dataset(root):
    """dataset/scannetv2 as gen_ps.py:27-58 expects it under `root` (a pathlib.Path): three tiny synthetic scans,
    a planes json for the first one.  Returns {scan: Scene}."""
    import json

    import torch

    from gapro_b200 import synthetic
    ds = root / "dataset" / "scannetv2"
    for sub in ("train", "superpoints", "scans_transform", "scannet_planes"):
        (ds / sub).mkdir(parents=True, exist_ok=True)
    scenes = {}
    for i, scan in enumerate(CLI_SCANS):
        sc = synthetic.make_scene(40 + i, "tiny")
        scenes[scan] = sc
        torch.save((sc.xyz_raw, sc.rgb, sc.sem, sc.inst), str(ds / "train" / f"{scan}_inst_nostuff.pth"))
        torch.save(sc.spp, str(ds / "superpoints" / f"{scan}.pth"))
        (ds / "scans_transform" / scan).mkdir(exist_ok=True)
        (ds / "scans_transform" / scan / f"{scan}.txt").write_text(
            "axisAlignment = " + " ".join(repr(float(x)) for x in sc.axis_align.ravel()) + "\nnumColorFrames = 1\n")
    # planes file of the first scan: the four walls of its bounding rectangle in the RAW (y-up) frame of the file
    sc = scenes[CLI_SCANS[0]]
    lo, hi = sc.xyz_raw.min(0), sc.xyz_raw.max(0)
    base = [[lo[0], lo[1]], [hi[0], lo[1]], [hi[0], hi[1]], [lo[0], hi[1]]]
    verts = [[float(x), float(lo[2]), float(-y)] for x, y in base] + [[float(x), float(hi[2]), float(-y)] for x, y in base]
    quads = [[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [0, 1, 2, 3]]
    (ds / "scannet_planes" / f"{CLI_SCANS[0]}.json").write_text(json.dumps({"verts": verts, "quads": quads}))
    return scenes
