"""Host-side planning between the occupancy stage and the GP stage: run the pair state
machine per scene (C ABI, host code) and lay out the index lists the device fills.

Pure numpy + the host entry point gapro_enumerate_events; no device work, so it is
covered by the CPU test-suite.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def enumerate_events(boxes_h, excl_h, inter_h, box_off, stride):
    """Events of every scene of a batch, in the loop order of
    /root/reference/gapro/gen_ps_utils.py:390-448.
    boxes_h (Bt,6) f64, excl_h (Bt,) i32, inter_h (n_scenes, stride, stride) i32."""
    lib = _lib.load()
    ns = len(box_off) - 1
    ev_scene, ev_kind, ev_b1, ev_b2 = [], [], [], []
    ev_off = np.zeros(ns + 1, dtype=np.int32)
    for i in range(ns):
        b0, B = int(box_off[i]), int(box_off[i + 1] - box_off[i])
        cap = B * B + 1
        k_ = np.empty(cap, dtype=np.int32)
        b1_ = np.empty(cap, dtype=np.int32)
        b2_ = np.empty(cap, dtype=np.int32)
        bx = np.ascontiguousarray(boxes_h[b0:b0 + B], dtype=np.float64)
        ex = np.ascontiguousarray(excl_h[b0:b0 + B], dtype=np.int32)
        ic = np.ascontiguousarray(inter_h[i], dtype=np.int32)
        n_ev = _lib.check(lib.gapro_enumerate_events(bx.ctypes.data, B, ex.ctypes.data, ic.ctypes.data, stride,
                                                     k_.ctypes.data, b1_.ctypes.data, b2_.ctypes.data, cap),
                          "gapro_enumerate_events")
        ev_scene.append(np.full(n_ev, i, dtype=np.int32))
        ev_kind.append(k_[:n_ev].copy())
        ev_b1.append(b1_[:n_ev].copy())
        ev_b2.append(b2_[:n_ev].copy())
        ev_off[i + 1] = ev_off[i] + n_ev
    cat = lambda xs: np.concatenate(xs) if xs else np.zeros(0, np.int32)
    return dict(ev_off=ev_off, ev_scene=cat(ev_scene), ev_kind=cat(ev_kind), ev_b1=cat(ev_b1), ev_b2=cat(ev_b2))


def plan_lists(ev, box_off, excl_h, inter_h):
    """Offsets of every index list in one flat buffer laid out as
        [GP intersections (GP order) | nest intersections | GP training rows (b1 excl ++ b2 excl)]
    so that the GP stage sees plain CSR arrays (test_off / train_off) over its regions."""
    ev_scene, ev_kind, ev_b1, ev_b2 = ev["ev_scene"], ev["ev_kind"], ev["ev_b1"], ev["ev_b2"]
    n_ev = len(ev_kind)
    is_gp = ev_kind == _lib.EV_GP
    gp_ev = np.flatnonzero(is_gp)
    nest_ev = np.flatnonzero(~is_gp)
    R = len(gp_ev)
    if n_ev:
        inter_len = inter_h[ev_scene, np.minimum(ev_b1, ev_b2), np.maximum(ev_b1, ev_b2)].astype(np.int64)
    else:
        inter_len = np.zeros(0, np.int64)
    ev_list_off = np.zeros(n_ev, dtype=np.int64)
    order = np.concatenate([gp_ev, nest_ev]).astype(np.int64)
    offs = np.concatenate([[0], np.cumsum(inter_len[order])]).astype(np.int64)
    ev_list_off[order] = offs[:-1]
    n_inter_total = int(offs[-1])
    test_off = np.zeros(R + 1, dtype=np.int32)
    test_off[1:] = np.cumsum(inter_len[gp_ev])
    gb0 = np.asarray(box_off)[ev_scene[gp_ev]] if R else np.zeros(0, np.int32)
    m1 = excl_h[gb0 + ev_b1[gp_ev]].astype(np.int64) if R else np.zeros(0, np.int64)
    m2 = excl_h[gb0 + ev_b2[gp_ev]].astype(np.int64) if R else np.zeros(0, np.int64)
    train_off = np.zeros(R + 1, dtype=np.int32)
    train_off[1:] = np.cumsum(m1 + m2)
    ev_gp_off = np.full(n_ev, -1, dtype=np.int32)
    ev_gp_off[gp_ev] = test_off[:-1]
    return dict(
        n_events=n_ev, n_regions=R, gp_ev=gp_ev, inter_len=inter_len, ev_list_off=ev_list_off,
        n_inter_total=n_inter_total, n_test_total=int(test_off[-1]), n_train_total=int(train_off[-1]),
        test_off=test_off, train_off=train_off, m1=m1, m2=m2, ev_gp_off=ev_gp_off,
        region_scene=ev_scene[gp_ev] if R else np.zeros(0, np.int32),
        list_scene=np.concatenate([ev_scene, ev_scene[gp_ev], ev_scene[gp_ev]]).astype(np.int32),
        list_b1=np.concatenate([ev_b1, ev_b1[gp_ev], ev_b2[gp_ev]]).astype(np.int32),
        list_b2=np.concatenate([ev_b2, np.full(2 * R, -1)]).astype(np.int32),
        list_off=np.concatenate([ev_list_off, n_inter_total + train_off[:-1],
                                 n_inter_total + train_off[:-1] + m1]).astype(np.int32),
    )
