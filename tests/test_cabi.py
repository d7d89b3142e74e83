"""CPU tests of the C-ABI shared library: it loads, exports every symbol the header declares,
and its HOST entry points (IoU, pair state machine) agree with the oracle.  No device calls."""
import os
import re

import numpy as np
import pytest

from gapro_b200 import _lib, plan, synthetic
from gapro_b200.gen_ps import synthetic_inputs
from oracle import gen_ps_oracle as O
from tests.conftest import oracle_args

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "gapro_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gapro_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(lib):
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/gapro_b200.h but not exported"
    assert set(syms) == set(_lib.EXPORTS), "ctypes signature table and header disagree"
    assert lib.gapro_version() >= 100


def test_box_iou_is_bit_exact_with_oracle(lib):
    rng = np.random.default_rng(0)
    lo = rng.uniform(0, 5, (40, 3))
    boxes = np.concatenate([lo, lo + rng.uniform(0.05, 3, (40, 3))], 1)
    boxes[5] = boxes[4]                       # identical boxes
    boxes[7, 3:] = boxes[7, :3]               # degenerate (zero volume)
    iou = np.zeros((40, 40))
    assert lib.gapro_box_iou(boxes.ctypes.data, 40, iou.ctypes.data) == 0
    assert (iou == O.box_iou_cross(boxes)).all()
    assert (np.diag(iou) == 0).all()


def counts_from_occupancy(occ_spp, n_bbs, stride):
    B = occ_spp.shape[1]
    excl = (occ_spp & (n_bbs == 1)[:, None]).sum(0).astype(np.int32)
    inter = np.zeros((stride, stride), dtype=np.int32)
    o = occ_spp.astype(np.int32)
    full = o.T @ o
    for b1 in range(B):
        for b2 in range(b1 + 1, B):
            inter[b1, b2] = full[b1, b2]
    return excl, inter


@pytest.mark.parametrize("name,seed", [("tiny", 0), ("tiny", 3), ("small", 1), ("small", 4)])
def test_event_enumeration_equals_oracle_loop(lib, name, seed):
    inp = synthetic_inputs(synthetic.make_scene(seed, name))
    fake = lambda X, n1, Xt, nz: dict(conf=np.full(len(Xt), .7, np.float32), label=np.ones(len(Xt), bool),
                                      mu=np.zeros(len(Xt), np.float32), var=np.ones(len(Xt), np.float32))
    _, dbg = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999, fit_fn=fake, return_debug=True)
    B = len(dbg["boxes"])
    stride = 32 * ((B + 31) // 32)
    excl, inter = counts_from_occupancy(dbg["occ_spp"], dbg["n_bbs"], stride)
    ev = plan.enumerate_events(dbg["boxes"], excl, inter[None], np.array([0, B], np.int32), stride)
    kinds = {0: "nest", 1: "nest", 2: "gp"}
    got = [(kinds[int(k)], int(a), int(b)) for k, a, b in zip(ev["ev_kind"], ev["ev_b1"], ev["ev_b2"])]
    assert got == [(e[0], e[1], e[2]) for e in dbg["events"]]
    # winners of nest events
    for (k, a, b), e in zip(zip(ev["ev_kind"], ev["ev_b1"], ev["ev_b2"]), dbg["events"]):
        if e[0] == "nest":
            assert (a if k == _lib.EV_NEST_B1 else b) == e[3]
    # planned lists reproduce the oracle's index sets when compaction is emulated in numpy
    pl = plan.plan_lists(ev, np.array([0, B], np.int32), excl, inter[None])
    assert pl["n_regions"] == len(dbg["regions"])
    buf = np.full(pl["n_inter_total"] + pl["n_train_total"], -1, dtype=np.int64)
    for sc, b1, b2, off in zip(pl["list_scene"], pl["list_b1"], pl["list_b2"], pl["list_off"]):
        if b2 < 0:
            idx = np.flatnonzero(dbg["occ_spp"][:, b1] & (dbg["n_bbs"] == 1))
        else:
            idx = np.flatnonzero(dbg["occ_spp"][:, b1] & dbg["occ_spp"][:, b2])
        buf[off:off + len(idx)] = idx
    assert (buf >= 0).all()
    for r, reg in enumerate(dbg["regions"]):
        t0, t1 = pl["train_off"][r], pl["train_off"][r + 1]
        q0, q1 = pl["test_off"][r], pl["test_off"][r + 1]
        train = buf[pl["n_inter_total"] + t0: pl["n_inter_total"] + t1]
        assert (train == np.concatenate([reg["b1_inds"], reg["b2_inds"]])).all() and pl["m1"][r] == len(reg["b1_inds"])
        assert (buf[q0:q1] == reg["inter"]).all()


def test_event_enumeration_equals_oracle_loop_on_random_clutter(lib):
    """25 small scenes with heavy overlap and nesting: the host state machine (gapro_enumerate_events) against the
    oracle's pair loop, event for event (the oracle's loop is itself pinned to the reference's, test_reference_run.py)."""
    fake = lambda X, n1, Xt, nz: dict(conf=np.full(len(Xt), .7, np.float32), label=np.ones(len(Xt), bool),
                                      mu=np.zeros(len(Xt), np.float32), var=np.ones(len(Xt), np.float32))
    cfg = synthetic.SceneConfig(n_points=1500, n_objects=12, s_target=120, overlap=0.7, n_nested=3)
    kinds = {0: "nest", 1: "nest", 2: "gp"}
    n_nest = n_gp = 0
    for seed in range(200, 225):
        inp = synthetic_inputs(synthetic.make_scene(seed, cfg))
        _, dbg = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999, fit_fn=fake, return_debug=True)
        B = len(dbg["boxes"])
        stride = 32 * ((B + 31) // 32)
        excl, inter = counts_from_occupancy(dbg["occ_spp"], dbg["n_bbs"], stride)
        ev = plan.enumerate_events(dbg["boxes"], excl, inter[None], np.array([0, B], np.int32), stride)
        got = [(kinds[int(k)], int(a), int(b)) for k, a, b in zip(ev["ev_kind"], ev["ev_b1"], ev["ev_b2"])]
        assert got == [(e[0], e[1], e[2]) for e in dbg["events"]], seed
        for (k, a, b), e in zip(zip(ev["ev_kind"], ev["ev_b1"], ev["ev_b2"]), dbg["events"]):
            if e[0] == "nest":
                assert (a if k == _lib.EV_NEST_B1 else b) == e[3], seed
        n_nest += sum(e[0] == "nest" for e in dbg["events"])
        n_gp += sum(e[0] == "gp" for e in dbg["events"])
    assert n_nest >= 10 and n_gp >= 50


def test_event_enumeration_hand_cases(lib):
    # nested pair: b0 inside b1 -> NEST_B1 and break; b1 inside b0 -> NEST_B2
    small, big = [1, 1, 1.5, 2, 2, 2.5], [0, 0, 1, 4, 4, 3]
    stride = 32
    inter = np.zeros((1, stride, stride), np.int32)
    inter[0, 0, 1] = 5
    excl = np.array([3, 7], np.int32)
    ev = plan.enumerate_events(np.array([small, big], np.float64), excl, inter, np.array([0, 2], np.int32), stride)
    assert ev["ev_kind"].tolist() == [_lib.EV_NEST_B1] and (ev["ev_b1"][0], ev["ev_b2"][0]) == (0, 1)
    ev = plan.enumerate_events(np.array([big, small], np.float64), excl, inter, np.array([0, 2], np.int32), stride)
    assert ev["ev_kind"].tolist() == [_lib.EV_NEST_B2]
    # overlapping, not nested, IoU < 0.6 -> GP; empty exclusive set or empty intersection -> nothing
    a, b = [0, 0, 0, 2, 2, 2], [1, 1, 1, 3, 3, 3]
    ev = plan.enumerate_events(np.array([a, b], np.float64), excl, inter, np.array([0, 2], np.int32), stride)
    assert ev["ev_kind"].tolist() == [_lib.EV_GP]
    ev = plan.enumerate_events(np.array([a, b], np.float64), np.array([0, 7], np.int32), inter, np.array([0, 2], np.int32), stride)
    assert len(ev["ev_kind"]) == 0
    ev = plan.enumerate_events(np.array([a, b], np.float64), excl, np.zeros_like(inter), np.array([0, 2], np.int32), stride)
    assert len(ev["ev_kind"]) == 0
    # IoU >= 0.6 -> skipped
    c = [0.15, 0, 0, 2.15, 2, 2]
    ev = plan.enumerate_events(np.array([a, c], np.float64), excl, inter, np.array([0, 2], np.int32), stride)
    assert len(ev["ev_kind"]) == 0


def test_errors_are_reported_not_fatal(lib):
    boxes = np.array([[0, 0, 0, 2, 2, 2], [1, 1, 1, 3, 3, 3]], np.float64)
    inter = np.zeros((32, 32), np.int32)
    inter[0, 1] = 1
    excl = np.array([1, 1], np.int32)
    k = np.zeros(1, np.int32)
    rc = lib.gapro_enumerate_events(boxes.ctypes.data, 2, excl.ctypes.data, inter.ctypes.data, 32, k.ctypes.data,
                                    k.ctypes.data, k.ctypes.data, 0)
    assert rc == -4 and b"more than 0 events" in lib.gapro_last_error()
    rc = lib.gapro_enumerate_events(boxes.ctypes.data, 2, excl.ctypes.data, inter.ctypes.data, 1, k.ctypes.data,
                                    k.ctypes.data, k.ctypes.data, 4)
    assert rc == -1
    with pytest.raises(_lib.GaproError):
        _lib.check(rc, "gapro_enumerate_events")
    # workspace queries are pure host arithmetic
    off = np.array([0, 100, 164], np.int32)
    toff = np.array([0, 10, 30], np.int32)
    full = lib.gapro_gp_workspace_bytes(2, off.ctypes.data, toff.ctypes.data, 6)
    need = lib.gapro_gp_min_workspace_bytes(2, off.ctypes.data, toff.ctypes.data, 6)
    assert 0 < need <= full
    assert lib.gapro_densify_workspace_bytes(1000, 2) > 1000 * 28


def test_gp_tile_tables_fit_the_advertised_workspace(lib):
    """Round-1 advisor finding: the tile tables of a chunk split into stream groups overran
    gapro_gp_workspace_bytes by ~50 KB for many small regions.  Host-side replay of the layout arithmetic for the
    cases the advisor computed and a sweep of random region mixes, for every group count."""
    rng = np.random.default_rng(0)
    cases = [[100] * 16, [150] * 40, list(rng.integers(2, 400, 60)), [2] * 500, [64] * 33, [65, 1607, 3, 700] * 5]
    cases += [list(rng.integers(1, 2000, int(rng.integers(1, 120)))) for _ in range(40)]
    for ms in cases:
        n = len(ms)
        tr = np.zeros(n + 1, np.int32)
        tr[1:] = np.cumsum(ms)
        te = np.zeros(n + 1, np.int32)
        te[1:] = np.cumsum(rng.integers(1, 300, n))
        for groups in range(1, 9):
            slack = lib.gapro_gp_debug_aux_slack(n, tr.ctypes.data, te.ctypes.data, groups)
            assert slack >= 0, (ms[:8], groups, slack)
