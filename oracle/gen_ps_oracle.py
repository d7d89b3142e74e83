"""Oracle for the per-scene pseudo-label pipeline (TEST INFRASTRUCTURE — see
oracle/__init__.py).

Restates, in numpy on the CPU:
  * `gen_pseudo_label_gaussian_process`  /root/reference/gapro/gen_ps_utils.py:293-482
  * `getInstanceInfo`                    /root/reference/gapro/gen_ps_utils.py:195-239
  * `batch_giou_cross` (IoU part)        /root/reference/gapro/gen_ps_utils.py:33-61
  * `is_box1_in_box2`                    /root/reference/gapro/gen_ps_utils.py:75-76
  * `is_within_bb_torch`                 /root/reference/gapro/gen_ps_utils.py:79-80
and the torch_scatter==2.0.9 semantics the reference relies on (CPU build:
`scatter(reduce="mean")` = index-order sum in the source dtype divided by the
count clamped at 1; `scatter_min` = strict `<` update in index order, i.e. the
first minimal element wins ties).

dtype rules follow what /root/reference/gapro/gen_ps.py:79-89 hands over:
coords float64, features float32, boxes float32 promoted to float64 by the
concatenation with the float64 floor box (gen_ps_utils.py:321-345).
"""
from __future__ import annotations

import numpy as np

from . import gp_oracle

MARGIN = 0.005            # gen_ps_utils.py:350
IOU_GATE = 0.0001         # :393
NEST_OFFSET = 0.1         # :411,418
IOU_SKIP = 0.6            # :425
MAX_NUM = 1000000         # :308


# ----------------------------------------------------------------------------- helpers
def get_instance_info(xyz, instance_label, semantic_label, dataset_name="scannetv2"):
    """gen_ps_utils.py:195-239.  Boxes are listed in increasing GT id, empty ids
    skipped (so box index != GT id when ids have gaps); class = semantic label of
    the instance's first point, minus 2 for scannetv2 unless -100."""
    xyz = np.asarray(xyz)
    instance_label = np.asarray(instance_label)
    semantic_label = np.asarray(semantic_label)
    n_inst = int(instance_label.max()) + 1
    cls, box, vol = [], [], []
    corners = np.full((xyz.shape[0], 6), -100.0, dtype=np.float32)
    for i in range(n_inst):
        idx = np.flatnonzero(instance_label == i)
        if idx.size == 0:
            continue
        pts = xyz[idx]
        lo, hi = pts.min(0), pts.max(0)
        corners[idx, :3] = lo - pts
        corners[idx, 3:] = hi - pts
        box.append(np.concatenate([lo, hi]))
        cls.append(semantic_label[idx[0]])
        vol.append(np.prod(np.clip(hi - lo, 0.0, None)))
    if not cls:
        return None
    cls = np.array(cls)
    if dataset_name == "scannetv2":
        cls[cls != -100] -= 2
    return n_inst, cls, np.stack(box, 0), np.array(vol), corners


def box_iou_cross(boxes):
    """IoU half of gen_ps_utils.py:33-50 for boxes x boxes (float64), '+1e-6' in
    the denominator; diagonal zeroed as at :386."""
    b1 = boxes[:, None, :]
    b2 = boxes[None, :, :]
    e = np.clip(np.minimum(b1[..., 3:], b2[..., 3:]) - np.maximum(b1[..., :3], b2[..., :3]), 0.0, None)
    inter = e[..., 0] * e[..., 1] * e[..., 2]
    d1 = np.clip(b1[..., 3:] - b1[..., :3], 0.0, None)
    d2 = np.clip(b2[..., 3:] - b2[..., :3], 0.0, None)
    v1 = d1[..., 0] * d1[..., 1] * d1[..., 2]
    v2 = d2[..., 0] * d2[..., 1] * d2[..., 2]
    union = v1 + v2 - inter
    iou = inter / (union + 1e-6)
    np.fill_diagonal(iou, 0.0)
    return iou


def box1_in_box2(b1, b2, offset):
    """gen_ps_utils.py:75-76."""
    return bool(np.all(b1[:3] + offset >= b2[:3]) and np.all(b1[3:] - offset <= b2[3:]))


def scatter_sum_index_order(src, index, n_out):
    """Sum rows of `src` into `n_out` bins, accumulating in src.dtype strictly in
    increasing row order (torch_scatter CPU semantics)."""
    src = np.asarray(src)
    order = np.argsort(index, kind="stable")
    sidx = index[order]
    counts = np.bincount(index, minlength=n_out)
    starts = np.concatenate([[0], np.cumsum(counts)[:-1]])
    out = np.zeros((n_out,) + src.shape[1:], dtype=src.dtype)
    live = np.flatnonzero(counts > 0)
    k = 0
    while live.size:
        out[live] = out[live] + src[order[starts[live] + k]]
        k += 1
        live = live[counts[live] > k]
    del sidx
    return out, counts


def build_boxes(coords, instance_box, instance_box_volume, instance_cls, wall_box, wall_box_volume,
                instance_classes, ground_h):
    """gen_ps_utils.py:317-347: floor slab from the scene extent, concatenation
    (float32 boxes widened to float64), class 18 for wall/floor."""
    lo = coords.min(0)
    hi = coords.max(0)
    floor = np.array([[lo[0], lo[1], lo[2], hi[0], hi[1], lo[2] + ground_h]], dtype=np.float64)
    floor_vol = np.prod(np.clip(floor[:, 3:] - floor[:, :3], 0.001, None), axis=1)
    inst_box = np.asarray(instance_box, dtype=np.float32).astype(np.float64).reshape(-1, 6)
    inst_vol = np.asarray(instance_box_volume, dtype=np.float32).astype(np.float64).reshape(-1)
    cls = np.asarray(instance_cls, dtype=np.int64).reshape(-1)
    n_wall = len(wall_box)
    if n_wall > 0:
        wb = np.asarray(wall_box, dtype=np.float32).astype(np.float64).reshape(-1, 6)
        wv = np.asarray(wall_box_volume, dtype=np.float32).astype(np.float64).reshape(-1)
        boxes = np.concatenate([inst_box, wb, floor])
        vols = np.concatenate([inst_vol, wv, floor_vol])
    else:
        boxes = np.concatenate([inst_box, floor])
        vols = np.concatenate([inst_vol, floor_vol])
    boxes_cls = np.concatenate([cls, np.full(n_wall + 1, instance_classes, dtype=np.int64)])
    return boxes, boxes_cls, vols


def containment(coords, boxes):
    """gen_ps_utils.py:349-351 (margins evaluated in float64)."""
    lo = boxes[None, :, :3] - MARGIN
    hi = boxes[None, :, 3:] + MARGIN
    p = coords[:, None, :]
    return np.all(p >= lo, axis=-1) & np.all(p <= hi, axis=-1)


def pooled_occupancy(occ, spp_dense, n_spp, thresh):
    """gen_ps_utils.py:359-363: float32 mean of 0/1 then `>= thresh` with the
    python scalar cast to float32 (torch scalar-compare rule)."""
    cnt_in = np.zeros((n_spp, occ.shape[1]), dtype=np.int64)
    np.add.at(cnt_in, spp_dense, occ.astype(np.int64))
    cnt = np.bincount(spp_dense, minlength=n_spp)
    mean = cnt_in.astype(np.float32) / np.maximum(cnt, 1).astype(np.float32)[:, None]
    occ_spp = mean >= np.float32(thresh)
    return occ_spp, cnt_in, cnt


# ----------------------------------------------------------------------------- pipeline
def gen_pseudo_label_oracle(coords_float, mask_feats, spp, instance_cls, instance_box, instance_box_volume,
                            wall_box, wall_box_volume, instance_classes=18, dataset_name="scannetv2",
                            ground_h=0.1, training_iter=50, thresh_spp_occu=0.8, noise_seed=0,
                            jitter_zz=1e-4, jitter_xx=1e-4, fit_fn=None, return_debug=False):
    """Returns (sem[N] i32, inst[N] i32, prob[N] f32, mu[S] f32, var[S] f32) like
    gen_ps_utils.py:482.  `noise_seed` seeds ONE numpy Generator per scene; each GP
    region draws its M standard normals from it in loop order (the reference draws
    from the unseeded global torch RNG inside gpytorch)."""
    coords = np.asarray(coords_float, dtype=np.float64)
    feats = np.asarray(mask_feats).astype(np.float32)
    spp_raw = np.asarray(spp).astype(np.int64)
    n_fg = len(instance_box)
    _, spp_dense = np.unique(spp_raw, return_inverse=True)            # :312
    spp_dense = spp_dense.reshape(-1)
    S = int(spp_dense.max()) + 1
    boxes, boxes_cls, boxes_vol = build_boxes(coords, instance_box, instance_box_volume, instance_cls,
                                              wall_box, wall_box_volume, instance_classes, ground_h)
    B = len(boxes)
    occ = containment(coords, boxes)                                  # :349
    fsum, cnt = scatter_sum_index_order(feats, spp_dense, S)          # :357
    feats_spp = fsum / np.maximum(cnt, 1).astype(np.float32)[:, None]
    occ_spp, cnt_in, _ = pooled_occupancy(occ, spp_dense, S, thresh_spp_occu)   # :359-362
    n_bbs = occ_spp.sum(1)                                            # :363

    inst = np.full(S, -100, dtype=np.int32)                           # :365-369
    determined = np.zeros(S, dtype=np.int64)
    prob = np.zeros(S, dtype=np.float32)
    mu = np.full(S, -100.0, dtype=np.float32)
    var = np.full(S, -100.0, dtype=np.float32)
    one = n_bbs == 1                                                  # :373-377
    inst[one] = np.argmax(occ_spp[one], axis=1).astype(np.int32)
    prob[one] = 1
    determined[one] = MAX_NUM
    none = n_bbs == 0                                                 # :381-383
    inst[none] = -1
    prob[none] = 1
    determined[none] = MAX_NUM

    iou = box_iou_cross(boxes)                                        # :385-386
    visited = np.zeros(B, dtype=bool)
    rng = np.random.default_rng(noise_seed)
    fit = fit_fn or (lambda X, n1, Xt, nz: gp_oracle.fit_region_autograd(
        X, n1, Xt, nz, iters=training_iter, jitter_zz=jitter_zz, jitter_xx=jitter_xx, policy="fp64"))
    events, regions = [], []
    for b1 in range(B):                                               # :390
        overlap = np.flatnonzero((iou[b1] > IOU_GATE) & ~visited)     # :393-394
        if overlap.size == 0:
            visited[b1] = True
            continue
        for b2 in overlap:
            b2 = int(b2)
            inter = np.flatnonzero(occ_spp[:, b1] & occ_spp[:, b2])   # :403-405
            if inter.size == 0:
                continue
            if box1_in_box2(boxes[b1], boxes[b2], NEST_OFFSET):       # :411-416
                inst[inter] = b1
                determined[inter] = MAX_NUM
                prob[inter] = 1
                visited[b1] = True
                events.append(("nest", b1, b2, b1))
                break
            if box1_in_box2(boxes[b2], boxes[b1], NEST_OFFSET):       # :418-423
                inst[inter] = b2
                determined[inter] = MAX_NUM
                prob[inter] = 1
                visited[b2] = True
                events.append(("nest", b1, b2, b2))
                continue
            if iou[b1, b2] >= IOU_SKIP:                               # :425
                continue
            b1_inds = np.flatnonzero((inst == b1) & (n_bbs == 1))     # :428-429
            b2_inds = np.flatnonzero((inst == b2) & (n_bbs == 1))
            if b1_inds.size == 0 or b2_inds.size == 0:
                continue
            train_x = np.concatenate([feats_spp[b1_inds], feats_spp[b2_inds]])   # gaussian_process_utils.py:386-395
            noise = rng.standard_normal(len(train_x)).astype(np.float32)
            res = fit(train_x, len(b1_inds), feats_spp[inter], noise)
            ow = prob[inter] < res["conf"]                            # :438 (float32, strict)
            sel = inter[ow]
            inst[sel] = np.where(res["label"][ow], b2, b1)            # :440-441
            prob[sel] = res["conf"][ow]
            mu[sel] = res["mu"][ow]
            var[sel] = res["var"][ow]
            determined[sel] = len(inter)                              # :446
            events.append(("gp", b1, b2, len(train_x)))
            regions.append(dict(b1=b1, b2=b2, b1_inds=b1_inds, b2_inds=b2_inds, inter=inter, res=res,
                                noise=noise, overwrite=ow))
        visited[b1] = True

    left = (n_bbs > 1) & (determined == 0)                            # :450-464
    for s in np.flatnonzero(left):
        best, best_v = -1, np.inf
        for b in np.flatnonzero(occ_spp[s]):
            if boxes_vol[b] < best_v:
                best, best_v = int(b), boxes_vol[b]
        inst[s] = best
    prob[left] = 1.0

    sem_spp = np.full(S, -100, dtype=np.int32)                        # :467-476
    inst_spp = np.full(S, -100, dtype=np.int32)
    pos = inst >= 0
    sem_spp[pos] = boxes_cls[inst[pos]].astype(np.int32)
    sem_spp[inst == -1] = instance_classes
    inst_spp[pos] = inst[pos]
    bg = inst_spp >= n_fg
    inst_spp[bg] = -100
    sem_spp[inst_spp >= n_fg] = instance_classes
    out = (sem_spp[spp_dense].astype(np.int32), inst_spp[spp_dense].astype(np.int32),
           prob[spp_dense].astype(np.float32), mu, var)               # :478-482
    if return_debug:
        return out, dict(spp_dense=spp_dense, boxes=boxes, boxes_vol=boxes_vol, boxes_cls=boxes_cls,
                         occ_spp=occ_spp, cnt_in=cnt_in, cnt=cnt, n_bbs=n_bbs, feats_spp=feats_spp,
                         iou=iou, events=events, regions=regions, inst_spp_raw=inst, prob_spp=prob)
    return out
