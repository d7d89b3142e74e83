"""Oracle for the heuristic labelers (TEST INFRASTRUCTURE — see oracle/__init__.py).

Restates, in numpy on the CPU:
  * `gen_pseudo_label_box2mask`  /root/reference/gapro/gen_ps_utils.py:242-290
  * `gen_pseudo_label`           /root/reference/gapro/gen_ps_utils.py:485-569
  * `spp_align_label`            /root/reference/gapro/gen_ps_utils.py:99-129
dtype rules: coordinates float64; `instance_box` is float32 there (gen_ps.py:80), so the +-0.005
margins and the box centres are float32 values, compared / subtracted against float64 points.
"""
from __future__ import annotations

import numpy as np


def point_containment(coords, instance_box):
    """gen_ps_utils.py:248-250 / :504-506."""
    box = np.asarray(instance_box, dtype=np.float32)
    lo = (box[:, :3] - np.float32(0.005)).astype(np.float64)
    hi = (box[:, 3:] + np.float32(0.005)).astype(np.float64)
    p = np.asarray(coords, dtype=np.float64)[:, None, :]
    return np.all(p >= lo[None], axis=-1) & np.all(p <= hi[None], axis=-1)


def spp_align(spp_dense, label, n_classes, occ_spp=None):
    """gen_ps_utils.py:99-121: per-superpoint majority vote over the per-point labels (label 0 =
    background); with `occ_spp` (n_boxes, S) the votes of box labels are zeroed where the box does not
    hold the superpoint; first maximum wins."""
    S = int(spp_dense.max()) + 1
    count = np.zeros((n_classes, S), dtype=np.int64)
    np.add.at(count, (label, spp_dense), 1)
    count = count.astype(np.float32)
    if occ_spp is not None:
        count[1:, :] = count[1:, :] * occ_spp.astype(np.float32)
    label_spp = np.argmax(count, axis=0)
    return label_spp[spp_dense], label_spp


def heuristic_labels(coords_float, spp, instance_cls, instance_box, instance_box_volume, instance_classes=18,
                     dataset_name="scannetv2", heuristic_rule="volume", box2mask=False):
    """`box2mask=True`: gen_pseudo_label_box2mask; else gen_pseudo_label with the given rule."""
    coords = np.asarray(coords_float, dtype=np.float64)
    box = np.asarray(instance_box, dtype=np.float32)
    vol = np.asarray(instance_box_volume, dtype=np.float32)
    cls = np.asarray(instance_cls, dtype=np.int64)
    N, K = len(coords), len(box)
    occ = point_containment(coords, box)
    n_bbs = occ.sum(1)
    inst = np.full(N, -100, dtype=np.int64)
    one = n_bbs == 1
    inst[one] = np.argmax(occ[one], axis=1)
    inst[n_bbs == 0] = -1
    multi = np.flatnonzero(n_bbs > 1)
    if box2mask or heuristic_rule == "volume":
        for i in multi:                        # scatter_min: strict '<' in box order, first minimum wins
            bs = np.flatnonzero(occ[i])
            inst[i] = bs[np.argmin(vol[bs])]
    elif heuristic_rule == "dist":
        centre = ((box[:, :3] + box[:, 3:]) / np.float32(2.0)).astype(np.float64)
        # gen_ps_utils.py:526 indexes coords_float with `point_inds`, which are row numbers of the COMPACTED
        # sub-matrix bb_occupancy[num_BBs_per_point > 1] (:516), not point ids: the k-th multi-box point is
        # measured from the coordinates of point k.  Restated as it executes.
        for rank, i in enumerate(multi):
            bs = np.flatnonzero(occ[i])
            d = coords[rank][None, :] - centre[bs]
            d2 = (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
            inst[i] = bs[np.argmin(d2)]
    elif heuristic_rule == "none":
        inst[multi] = -2
    else:
        raise Exception
    if dataset_name == "scannetv2":
        _, dense = np.unique(np.asarray(spp), return_inverse=True)
        dense = dense.reshape(-1)
        occ_spp = None
        if not box2mask:
            S = int(dense.max()) + 1
            cnt_in = np.zeros((K, S), dtype=np.int64)
            np.add.at(cnt_in.T, dense, occ.astype(np.int64))
            cnt = np.bincount(dense, minlength=S)
            occ_spp = (cnt_in.astype(np.float32) / cnt.astype(np.float32)[None, :]) >= np.float32(0.7)
        label = np.where(inst >= 0, inst + 1, 0)
        label, _ = spp_align(dense, label, K + 1, occ_spp)
        inst = np.where(label > 0, label - 1, -1)
    sem_out = np.full(N, -100, dtype=np.int32)
    inst_out = np.full(N, -100, dtype=np.int32)
    pos = inst >= 0
    sem_out[pos] = cls[inst[pos]].astype(np.int32)
    sem_out[inst == -1] = instance_classes
    inst_out[pos] = inst[pos]
    return sem_out, inst_out
