"""Deterministic ScanNet-shaped synthetic scenes.

A scene is what `/root/reference/gapro/gen_ps.py:45-53` loads from disk: raw
(un-aligned) xyz float64, rgb float64 in [-1, 1], semantic / instance labels
float64, raw superpoint ids int64 (with gaps), an axis-alignment 4x4, and
optionally 32-d deep features float32.  `prepare_inputs` then performs the host
steps of gen_ps.py:55-89 (feature concat BEFORE alignment, alignment, boxes).

Everything is numpy + a seeded `default_rng`, so the same seed yields the same
scene here and on the GPU box.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class SceneConfig:
    n_points: int = 150_000
    n_objects: int = 30
    s_target: int = 4_000
    feat_dim: int = 6              # 6 = [raw xyz, rgb]; 32 = synthetic "deep" features
    n_walls: int = 4
    overlap: float = 0.3           # fraction of objects pushed into a neighbour
    n_nested: int = 0              # objects placed strictly inside an earlier box
    room: tuple = ((6.0, 10.0), (5.0, 8.0), (2.6, 3.2))
    size_range: tuple = (0.3, 2.0)
    inst_id_gaps: bool = True
    shuffle: bool = True
    n_giants: int = 0              # elevated multi-layer structures overlapping pairwise (the 8k-row regions of C4)
    giant_layers: int = 6
    max_height: float | None = None   # cap on the regular objects' height (keeps them below the giants)


# named configurations of BASELINE.json:configs
CONFIGS = {
    "c1": SceneConfig(),
    "c1_deep": SceneConfig(feat_dim=32),
    # heavy-overlap stress (BASELINE configs[3]): 80 boxes = 58 overlapping + 20 nested + 2 elevated shelf
    # structures whose pair is ONE region of ~8k superpoints (M >= 4096 training rows + the overlap zone)
    "c4": SceneConfig(n_objects=78, s_target=17_500, n_nested=20, overlap=0.6, n_points=400_000, n_giants=2,
                      max_height=1.0, size_range=(0.3, 1.4)),
    "c5": SceneConfig(n_points=1_000_000, n_objects=120, s_target=30_000, overlap=0.5,
                      room=((14.0, 20.0), (10.0, 16.0), (2.8, 3.4))),
    "tiny": SceneConfig(n_points=6_000, n_objects=6, s_target=300, overlap=0.5, n_nested=1),
    "small": SceneConfig(n_points=20_000, n_objects=10, s_target=700, overlap=0.4, n_nested=1),
}


def c3_config(i: int) -> SceneConfig:
    """Scene i of the 1201-scene ScanNetv2-train-shaped batch: N ~ U(50k, 250k)."""
    r = np.random.default_rng([3, i])
    n = int(r.integers(50_000, 250_001))
    k = int(r.integers(12, 41))
    return SceneConfig(n_points=n, n_objects=k, s_target=int(4000 * n / 150_000))


@dataclass
class Scene:
    xyz_raw: np.ndarray
    rgb: np.ndarray
    sem: np.ndarray
    inst: np.ndarray
    spp: np.ndarray
    axis_align: np.ndarray
    wall_box: np.ndarray
    wall_volume: np.ndarray
    deep_feats: np.ndarray | None = None
    meta: dict = field(default_factory=dict)


def _place_objects(rng, cfg, room):
    rx, ry, rz = room
    lo_s, hi_s = cfg.size_range
    boxes = []
    n_plain = cfg.n_objects - cfg.n_nested
    for k in range(n_plain):
        size = rng.uniform(lo_s, hi_s, 3)
        size[2] = min(size[2], rz - 0.3)
        if cfg.max_height is not None:
            size[2] = min(size[2], cfg.max_height)
        if boxes and rng.random() < cfg.overlap:
            # push into a neighbour: overlap 10-50 % of the smaller extent along one axis
            nb = boxes[int(rng.integers(len(boxes)))]
            c = 0.5 * (nb[:3] + nb[3:])
            ax = int(rng.integers(2))
            sgn = 1.0 if rng.random() < 0.5 else -1.0
            ov = rng.uniform(0.1, 0.5) * min(size[ax], nb[3 + ax] - nb[ax])
            cx = c.copy()
            cx[ax] = c[ax] + sgn * (0.5 * (nb[3 + ax] - nb[ax]) + 0.5 * size[ax] - ov)
            cx[1 - ax] = c[1 - ax] + rng.uniform(-0.3, 0.3) * size[1 - ax]
        else:
            cx = np.array([rng.uniform(0.2 + size[0] / 2, rx - 0.2 - size[0] / 2),
                           rng.uniform(0.2 + size[1] / 2, ry - 0.2 - size[1] / 2), 0.0])
        lo = np.array([cx[0] - size[0] / 2, cx[1] - size[1] / 2, 0.0])
        hi = np.array([cx[0] + size[0] / 2, cx[1] + size[1] / 2, size[2]])
        lo[:2] = np.clip(lo[:2], 0.05, None)
        hi[0] = min(hi[0], rx - 0.05)
        hi[1] = min(hi[1], ry - 0.05)
        if np.any(hi - lo < 0.15):
            hi = lo + np.maximum(hi - lo, 0.15)
        boxes.append(np.concatenate([lo, hi]))
    for k in range(cfg.n_nested):
        # a smaller object strictly inside (by > 0.1 m) an earlier, big enough box
        big = [b for b in boxes if np.all(b[3:] - b[:3] > 0.65)]
        if not big:
            big = boxes
        nb = big[int(rng.integers(len(big)))]
        ext = nb[3:] - nb[:3]
        size = np.maximum(ext * rng.uniform(0.25, 0.55, 3), 0.12)
        room_ = np.maximum(ext - size - 0.3, 0.0)
        lo = nb[:3] + 0.15 + rng.uniform(0, 1, 3) * room_
        boxes.append(np.concatenate([lo, lo + size]))
    return np.stack(boxes)


def _sample_faces(rng, faces, n, cell):
    """faces: list of (origin(3), u(3), v(3)) rectangles.  Returns points, face id,
    in-face cell id."""
    areas = np.array([np.linalg.norm(np.cross(u, v)) for _, u, v in faces])
    which = rng.choice(len(faces), size=n, p=areas / areas.sum())
    a = rng.random(n)
    b = rng.random(n)
    org = np.stack([f[0] for f in faces])[which]
    U = np.stack([f[1] for f in faces])[which]
    V = np.stack([f[2] for f in faces])[which]
    pts = org + a[:, None] * U + b[:, None] * V
    lu = np.linalg.norm(U, axis=1)
    lv = np.linalg.norm(V, axis=1)
    nu = np.maximum(np.ceil(lu / cell), 1)
    ci = np.minimum((a * nu).astype(np.int64), nu.astype(np.int64) - 1)
    cj = (b * np.maximum(np.ceil(lv / cell), 1)).astype(np.int64)
    cell_id = cj * nu.astype(np.int64) + ci
    return pts, which, cell_id


def make_scene(seed: int, cfg: SceneConfig | str = "c1") -> Scene:
    if isinstance(cfg, str):
        cfg = CONFIGS[cfg]
    rng = np.random.default_rng([1234, int(seed)])
    room = np.array([rng.uniform(*cfg.room[0]), rng.uniform(*cfg.room[1]), rng.uniform(*cfg.room[2])])
    rx, ry, rz = room
    obj = _place_objects(rng, cfg, room)
    K = len(obj)

    # surfaces: id 0 floor, 1..4 walls, then 5 faces per object (top + 4 sides)
    surfaces = [(np.array([0.0, 0, 0]), np.array([rx, 0, 0]), np.array([0, ry, 0]))]
    surfaces += [
        (np.array([0.0, 0, 0]), np.array([rx, 0, 0]), np.array([0, 0, rz])),
        (np.array([0.0, ry, 0]), np.array([rx, 0, 0]), np.array([0, 0, rz])),
        (np.array([0.0, 0, 0]), np.array([0, ry, 0]), np.array([0, 0, rz])),
        (np.array([rx, 0, 0]), np.array([0, ry, 0]), np.array([0, 0, rz])),
    ]
    owner = [-1] * 5                       # -1 floor/wall
    for k, b in enumerate(obj):
        lo, hi = b[:3], b[3:]
        e = hi - lo
        surfaces += [
            (np.array([lo[0], lo[1], hi[2]]), np.array([e[0], 0, 0]), np.array([0, e[1], 0])),   # top
            (np.array([lo[0], lo[1], lo[2]]), np.array([e[0], 0, 0]), np.array([0, 0, e[2]])),
            (np.array([lo[0], hi[1], lo[2]]), np.array([e[0], 0, 0]), np.array([0, 0, e[2]])),
            (np.array([lo[0], lo[1], lo[2]]), np.array([0, e[1], 0]), np.array([0, 0, e[2]])),
            (np.array([hi[0], lo[1], lo[2]]), np.array([0, e[1], 0]), np.array([0, 0, e[2]])),
        ]
        owner += [k] * 5
    # giants: shelf-like structures above the regular objects, horizontal layers + 4 sides, consecutive ones
    # overlapping by ~0.2 of the room length (IoU < 0.6, not nested): their pair is one very large GP region
    for gi in range(cfg.n_giants):
        w = 0.55 if cfg.n_giants > 1 else 0.8
        x0 = 0.05 * rx + (0.9 - w) * rx * (gi / max(cfg.n_giants - 1, 1))
        lo = np.array([x0, 0.35, 1.25])
        hi = np.array([x0 + w * rx, ry - 0.35, min(rz - 0.15, 2.45)])
        e = hi - lo
        for z in np.linspace(lo[2], hi[2], cfg.giant_layers):
            surfaces.append((np.array([lo[0], lo[1], z]), np.array([e[0], 0, 0]), np.array([0, e[1], 0])))
        surfaces += [
            (np.array([lo[0], lo[1], lo[2]]), np.array([e[0], 0, 0]), np.array([0, 0, e[2]])),
            (np.array([lo[0], hi[1], lo[2]]), np.array([e[0], 0, 0]), np.array([0, 0, e[2]])),
            (np.array([lo[0], lo[1], lo[2]]), np.array([0, e[1], 0]), np.array([0, 0, e[2]])),
            (np.array([hi[0], lo[1], lo[2]]), np.array([0, e[1], 0]), np.array([0, 0, e[2]])),
        ]
        owner += [K + gi] * (cfg.giant_layers + 4)
    n_reg = K
    K = K + cfg.n_giants
    owner = np.array(owner)
    total_area = sum(np.linalg.norm(np.cross(u, v)) for _, u, v in surfaces)
    cell = float(np.sqrt(total_area / cfg.s_target))
    pts, face, cell_id = _sample_faces(rng, surfaces, cfg.n_points, cell)
    pts = pts + rng.normal(0.0, 0.005, pts.shape)
    inst = owner[face].astype(np.int64)

    # every object must own at least a few points (boxes come from points)
    for k in range(n_reg):
        if not np.any(inst == k):
            j = int(rng.integers(len(pts)))
            f = 5 + 5 * k
            pts[j] = surfaces[f][0] + 0.5 * surfaces[f][1] + 0.5 * surfaces[f][2]
            face[j], cell_id[j], inst[j] = f, 0, k

    # superpoints: one id per (face, cell), spread over a gappy raw id range
    key = face.astype(np.int64) * 1_000_003 + cell_id
    uniq, dense = np.unique(key, return_inverse=True)
    raw_ids = np.sort(rng.choice(3 * len(uniq) + 7, size=len(uniq), replace=False)).astype(np.int64)
    raw_ids = raw_ids[rng.permutation(len(uniq))]
    spp = raw_ids[dense]

    # colours / labels
    base = rng.uniform(-0.8, 0.8, (K + 2, 3))
    col_idx = np.where(inst >= 0, inst + 2, np.where(face == 0, 0, 1))
    rgb = np.clip(base[col_idx] + rng.normal(0, 0.05, pts.shape), -1.0, 1.0)
    obj_cls = rng.integers(0, 18, K)
    sem = np.where(inst >= 0, obj_cls[np.clip(inst, 0, None)] + 2, np.where(face == 0, 1, 0)).astype(np.float64)
    inst_lab = inst.astype(np.float64)
    if cfg.inst_id_gaps and K >= 4:
        remap = np.arange(K) + (np.arange(K) >= K // 2)          # one unused id in the middle
        inst_lab = np.where(inst >= 0, remap[np.clip(inst, 0, None)], -100).astype(np.float64)
    else:
        inst_lab[inst < 0] = -100.0

    # axis alignment: aligned = R * raw + t  (random z-rotation + translation)
    th = rng.uniform(0, 2 * np.pi)
    A = np.eye(4)
    A[:2, :2] = [[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]]
    A[:3, 3] = rng.uniform(-3, 3, 3)
    Ainv = np.linalg.inv(A)
    xyz_raw = (np.concatenate([pts, np.ones((len(pts), 1))], 1) @ Ainv.T)[:, :3]

    if cfg.shuffle:
        p = rng.permutation(len(pts))
        xyz_raw, rgb, sem, inst_lab, spp, inst, pts = xyz_raw[p], rgb[p], sem[p], inst_lab[p], spp[p], inst[p], pts[p]

    # wall boxes (float32, what get_wall_boxes would hand over), in ALIGNED frame of the
    # re-aligned cloud: recompute extents from aligned points to stay consistent.
    aligned = (np.concatenate([xyz_raw, np.ones((len(pts), 1))], 1) @ A.T)[:, :3]
    wall_box = np.zeros((0, 6), np.float32)
    wall_vol = np.zeros((0,), np.float32)
    if cfg.n_walls > 0:
        t = 0.1
        lo, hi = aligned.min(0), aligned.max(0)
        wb = [
            [lo[0], lo[1] - t, lo[2], hi[0], lo[1] + t, hi[2]],
            [lo[0], hi[1] - t, lo[2], hi[0], hi[1] + t, hi[2]],
            [lo[0] - t, lo[1], lo[2], lo[0] + t, hi[1], hi[2]],
            [hi[0] - t, lo[1], lo[2], hi[0] + t, hi[1], hi[2]],
        ][: cfg.n_walls]
        wall_box = np.array(wb, dtype=np.float32)
        wall_vol = np.prod(wall_box[:, 3:] - wall_box[:, :3], axis=1).astype(np.float32)

    deep = None
    if cfg.feat_dim == 32:
        r2 = np.random.default_rng(77)       # fixed "network"
        onehot = np.zeros((len(pts), 8))
        onehot[np.arange(len(pts)), np.clip(inst, -1, None) % 8] = (inst >= 0)
        x = np.concatenate([xyz_raw, rgb, onehot], 1)
        W1 = r2.normal(0, 0.5, (x.shape[1], 48))
        W2 = r2.normal(0, 0.3, (48, 32))
        deep = (np.tanh(x @ W1) @ W2 + rng.normal(0, 0.02, (len(pts), 32))).astype(np.float32)
    elif cfg.feat_dim != 6:
        raise ValueError("feat_dim must be 6 or 32")

    return Scene(xyz_raw=xyz_raw, rgb=rgb, sem=sem, inst=inst_lab, spp=spp, axis_align=A,
                 wall_box=wall_box, wall_volume=wall_vol, deep_feats=deep,
                 meta=dict(seed=seed, room=room, n_objects=K, cell=cell))
