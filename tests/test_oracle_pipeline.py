"""CPU tests of the pipeline oracle: known-answer cases built by hand for containment, the
occupancy threshold, the pair state machine, the fallback, plus the committed golden scenes."""
import os

import numpy as np
import pytest
import torch

from gapro_b200 import synthetic
from gapro_b200.gen_ps import synthetic_inputs
from oracle import gen_ps_oracle as O
from tests.conftest import oracle_args, rel_err
from tests.golden.make_golden import input_digest

GOLD_DIR = os.path.join(os.path.dirname(__file__), "golden")


def fake_fit(conf=0.75, label=True):
    def f(X, n1, Xt, nz):
        n = len(Xt)
        return dict(conf=np.full(n, conf, np.float32), label=np.full(n, label, bool),
                    mu=np.full(n, 0.5, np.float32), var=np.full(n, 2.0, np.float32))
    return f


def run_boxes(points, spp, boxes, cls=None, thresh=0.999, fit=None, walls=None):
    """Oracle on hand-built boxes.  One extra point far below the scene pins the floor slab
    (gen_ps_utils.py:317-326) away from the test geometry; it is stripped from the outputs."""
    points = np.concatenate([np.asarray(points, np.float64), [[points[:, 0].mean(), points[:, 1].mean(), -5.0]]])
    spp = np.concatenate([np.asarray(spp), [np.max(spp) + 1]])
    out, dbg = _run_boxes(points, spp, boxes, cls, thresh, fit, walls)
    sem, inst, prob, mu, var = out
    keep = dbg["spp_dense"][:-1]
    last = dbg["spp_dense"][-1]
    sel = np.arange(len(mu)) != last
    for k in ("occ_spp", "n_bbs"):
        dbg[k] = dbg[k][sel]
    dbg["occ_spp"] = dbg["occ_spp"][:, :-1]
    dbg["n_bbs"] = dbg["occ_spp"].sum(1)
    del keep
    return (sem[:-1], inst[:-1], prob[:-1], mu[sel], var[sel]), dbg


def _run_boxes(points, spp, boxes, cls, thresh, fit, walls):
    boxes = np.asarray(boxes, dtype=np.float32)
    vol = np.prod(np.clip(boxes[:, 3:] - boxes[:, :3], 0, None), axis=1).astype(np.float32)
    cls = np.arange(len(boxes)) if cls is None else np.asarray(cls)
    wb = np.asarray(walls, np.float32) if walls is not None else []
    wv = np.prod(wb[:, 3:] - wb[:, :3], axis=1) if walls is not None else []
    feats = np.concatenate([points, np.zeros_like(points)], 1).astype(np.float32)
    return O.gen_pseudo_label_oracle(np.asarray(points, np.float64), feats, np.asarray(spp), cls, boxes, vol, wb, wv,
                                     thresh_spp_occu=thresh, fit_fn=fit or fake_fit(), return_debug=True)


def test_containment_margin_is_inclusive_in_float64():
    box = np.array([[0.0, 0.0, 0.0, 1.0, 1.0, 1.0]])
    lo, hi = 0.0 - 0.005, 1.0 + 0.005
    pts = np.array([[lo, 0.5, 0.5], [np.nextafter(lo, -1), 0.5, 0.5], [hi, 0.5, 0.5], [np.nextafter(hi, 2), 0.5, 0.5],
                    [0.5, 0.5, hi], [0.5, hi + 1e-9, 0.5]])
    occ = O.containment(pts, box)
    assert occ[:, 0].tolist() == [True, False, True, False, True, False]


def test_occupancy_threshold_float32_mean():
    occ = np.zeros((2000, 1), dtype=bool)
    occ[:999] = True            # superpoint 0: 999/1000 inside
    occ[1000:1998] = True       # superpoint 1: 998/1000 inside
    spp = np.repeat([0, 1], 1000)
    occ_spp, cnt_in, cnt = O.pooled_occupancy(occ, spp, 2, 0.999)
    assert cnt_in[:, 0].tolist() == [999, 998] and cnt.tolist() == [1000, 1000]
    assert occ_spp[:, 0].tolist() == [True, False]
    # the comparison happens in float32, as torch does for a float32 tensor vs a python scalar
    t = torch.tensor([999.0], dtype=torch.float32) / torch.tensor([1000.0], dtype=torch.float32)
    assert bool((t >= 0.999).item()) is True


def test_scatter_sum_is_index_ordered_float32():
    rng = np.random.default_rng(0)
    src = (rng.normal(size=(5000, 4)) * 1e3).astype(np.float32)
    idx = rng.integers(0, 37, 5000)
    out, cnt = O.scatter_sum_index_order(src, idx, 37)
    ref = np.zeros((37, 4), dtype=np.float32)
    for i in range(5000):      # the definition: sequential float32 adds in row order
        ref[idx[i]] = ref[idx[i]] + src[i]
    assert (out.view(np.uint32) == ref.view(np.uint32)).all()
    assert (cnt == np.bincount(idx, minlength=37)).all()
    t = torch.zeros(37, 4).index_add_(0, torch.from_numpy(idx), torch.from_numpy(src))
    assert np.allclose(t.numpy(), out, rtol=1e-5)


def grid_points(lo, hi, n=4):
    ax = [np.linspace(lo[d] + 0.02, hi[d] - 0.02, n) for d in range(3)]
    return np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)


def test_single_box_and_background():
    pts = np.concatenate([grid_points([0, 0, 0.5], [1, 1, 1.5]), grid_points([5, 5, 2], [6, 6, 3])])
    spp = np.arange(len(pts)) * 7 + 3
    (sem, inst, prob, mu, var), dbg = run_boxes(pts, spp, [[0, 0, 0.5, 1, 1, 1.5]], cls=[4])
    inside = np.arange(len(pts)) < 64
    assert (inst[inside] == 0).all() and (sem[inside] == 4).all()
    assert (inst[~inside] == -100).all() and (sem[~inside] == 18).all()
    assert (prob == 1).all() and (mu == -100).all() and (var == -100).all()
    assert mu.shape == (len(pts),)      # one superpoint per point here; mu/var are per SUPERPOINT


def test_nested_box_takes_intersection_and_breaks():
    # box 0 strictly inside box 1 (by > 0.1): b1=0 in b2=1 -> intersection -> box 0, loop breaks
    big, small = [0, 0, 1, 4, 4, 3], [1, 1, 1.5, 2, 2, 2.5]
    pts = np.concatenate([grid_points(big[:3], big[3:], 6), grid_points(small[:3], small[3:], 3)])
    (sem, inst, prob, mu, var), dbg = run_boxes(pts, np.arange(len(pts)), [small, big], cls=[1, 2])
    assert ("nest", 0, 1, 0) in dbg["events"] and not any(e[0] == "gp" for e in dbg["events"])
    in_small = np.all((pts >= np.array(small[:3]) - 0.005) & (pts <= np.array(small[3:]) + 0.005), axis=1)
    assert (inst[in_small] == 0).all() and (inst[~in_small] == 1).all() and (prob == 1).all()


def test_nested_other_way_marks_b2_visited():
    # box 1 inside box 0: event (nest, 0, 1, winner 1); box 1 is then never a b2 again
    big, small = [0, 0, 1, 4, 4, 3], [1, 1, 1.5, 2, 2, 2.5]
    pts = np.concatenate([grid_points(big[:3], big[3:], 6), grid_points(small[:3], small[3:], 3)])
    (_, inst, _, _, _), dbg = run_boxes(pts, np.arange(len(pts)), [big, small], cls=[1, 2])
    assert dbg["events"][0] == ("nest", 0, 1, 1)
    in_small = np.all((pts >= np.array(small[:3]) - 0.005) & (pts <= np.array(small[3:]) + 0.005), axis=1)
    assert (inst[in_small] == 1).all()


def test_high_iou_pair_is_skipped_and_falls_back_to_smallest_volume():
    a, b = [0, 0, 1, 2, 2, 3], [0.15, 0, 1, 2.2, 2, 3]          # IoU ~ 0.86 >= 0.6, not nested (offset 0.1)
    pts = np.concatenate([grid_points(a[:3], a[3:], 5), grid_points(b[:3], b[3:], 5)])
    (_, inst, prob, mu, _), dbg = run_boxes(pts, np.arange(len(pts)), [a, b])
    assert dbg["iou"][0, 1] >= 0.6 and not dbg["events"]
    both = dbg["n_bbs"] == 2
    assert both.any() and (inst[both] == 0).all()                 # a has the smaller volume (4*2 vs 4.1*2)
    assert (prob == 1).all() and (mu == -100).all()


def test_gp_event_merge_is_strict_and_ordered():
    # three mutually overlapping boxes, none nested, each with superpoints of its own
    a, b, c = [0, 0, 1, 2, 2, 2], [1.5, 0, 1, 3.5, 2, 2], [1.2, 1.5, 1, 2.5, 3.5, 2]
    pts = np.concatenate([grid_points(a[:3], a[3:], 6), grid_points(b[:3], b[3:], 6), grid_points(c[:3], c[3:], 6)])
    calls = []

    def fit(X, n1, Xt, nz):
        calls.append((len(X), n1, len(Xt)))
        conf = np.float32(0.9 if len(calls) == 1 else 0.9)       # equal confidence later: strict '<' keeps the first
        return dict(conf=np.full(len(Xt), conf, np.float32), label=np.ones(len(Xt), bool),
                    mu=np.full(len(Xt), float(len(calls)), np.float32), var=np.ones(len(Xt), np.float32))
    (_, inst, prob, mu, _), dbg = run_boxes(pts, np.arange(len(pts)), [a, b, c], fit=fit)
    gp = [e for e in dbg["events"] if e[0] == "gp"]
    assert [(e[1], e[2]) for e in gp] == [(0, 1), (0, 2), (1, 2)]
    triple = dbg["n_bbs"] == 3
    assert triple.any()
    assert (mu[triple] == 1.0).all() and (inst[triple] == 1).all()     # first pair (0,1) won and was not overwritten
    only_bc = (dbg["occ_spp"][:, 1] & dbg["occ_spp"][:, 2] & ~dbg["occ_spp"][:, 0])
    assert (mu[only_bc] == 3.0).all() and (inst[only_bc] == 2).all()


def test_empty_exclusive_set_skips_gp():
    # box 1 has no superpoint of its own (every point of it also lies in box 0 or box 2)
    a, b, c = [0, 0, 1, 2, 1, 2], [1.0, 0, 1, 3.0, 1, 2], [2.0, 0, 1, 4, 1, 2]
    pts = np.concatenate([grid_points([0, 0, 1], [1.9, 1, 2], 5), grid_points([2.05, 0, 1], [4, 1, 2], 5)])
    (_, inst, prob, _, _), dbg = run_boxes(pts, np.arange(len(pts)), [a, b, c])
    assert not any(e[0] == "gp" and 1 in (e[1], e[2]) for e in dbg["events"])
    assert (prob == 1).all()


def test_wall_and_floor_points_become_background():
    inp = synthetic_inputs(synthetic.make_scene(0, "tiny"))
    (sem, inst, prob, mu, var), dbg = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999,
                                                                fit_fn=fake_fit(), return_debug=True)
    K = len(inp["instance_box"])
    assert inst.max() < K and set(np.unique(inst)) <= set(range(K)) | {-100}
    assert (sem[inst == -100] == 18).all()
    assert len(mu) == int(dbg["spp_dense"].max()) + 1 and sem.dtype == np.int32 and prob.dtype == np.float32


@pytest.mark.parametrize("name,seed,nseed", [("tiny", 3, 5), ("small", 4, 6)])
def test_oracle_reproduces_golden_scene(name, seed, nseed):
    gold = np.load(os.path.join(GOLD_DIR, f"scene_{name}.npz"))
    inp = synthetic_inputs(synthetic.make_scene(seed, name))
    args = oracle_args(inp)
    assert input_digest(args) == str(gold["digest"]), "synthetic generator drifted: regenerate tests/golden"
    res = O.gen_pseudo_label_oracle(*args, thresh_spp_occu=0.999, noise_seed=nseed)
    assert (res[0] == gold["sem"]).all() and (res[1] == gold["inst"]).all()
    assert np.allclose(res[2], gold["prob"], rtol=1e-5)
    g = gold["mu"] != -100
    assert ((res[3] != -100) == g).all()
    assert rel_err(res[3][g], gold["mu"][g]) < 1e-5 and rel_err(res[4][g], gold["var"][g]) < 1e-5


# ------------------------------------------------------------------------------- heuristic labelers (8f)
def test_heuristic_labelers_known_answers():
    from oracle import heuristic_oracle as H
    # two overlapping boxes; superpoint 0 = points only in box 0, superpoint 1 = points in both,
    # superpoint 2 = background
    box = np.array([[0, 0, 0, 2, 2, 2], [1, 0, 0, 4, 2, 2]], np.float32)      # volumes 8 and 12
    vol = np.prod(box[:, 3:] - box[:, :3], axis=1)
    pts = np.array([[0.5, 1, 1], [0.6, 1, 1], [1.5, 1, 1], [1.2, 1, 1], [1.9, 1, 1], [9, 9, 9], [9.5, 9, 9]])
    spp = np.array([10, 10, 20, 20, 20, 30, 30])
    cls = np.array([3, 7])
    sem, inst = H.heuristic_labels(pts, spp, cls, box, vol, box2mask=True)
    assert inst.tolist() == [0, 0, 0, 0, 0, -100, -100] and sem.tolist() == [3, 3, 3, 3, 3, 18, 18]
    sem, inst = H.heuristic_labels(pts, spp, cls, box, vol, heuristic_rule="dist")
    # centres x = 1 and 2.5: 1.5 and 1.2 are nearer to box 0, 1.9 to box 1 -> majority box 0
    assert inst.tolist() == [0, 0, 0, 0, 0, -100, -100]
    sem, inst = H.heuristic_labels(pts, spp, cls, box, vol, heuristic_rule="none")
    assert inst.tolist() == [0, 0, -100, -100, -100, -100, -100] and sem[2] == 18
    sem, inst = H.heuristic_labels(pts, spp, cls, box, vol, heuristic_rule="none", dataset_name="s3dis")
    assert inst.tolist() == [0, 0, -100, -100, -100, -100, -100] and sem[2] == -100 and sem[5] == 18
    # margins are float32: a point exactly on float32(lo - 0.005) is inside
    edge = float(np.float32(1.0) - np.float32(0.005))
    occ = H.point_containment(np.array([[edge, 1, 1], [np.nextafter(edge, 0), 1, 1]]), box[1:])
    assert occ[:, 0].tolist() == [True, False]
    # the 70 % rule: a superpoint with 2 of 3 points in the box cannot take the box label
    pts2 = np.array([[0.5, 1, 1], [0.6, 1, 1], [-3, 1, 1]])
    sem, inst = H.heuristic_labels(pts2, np.zeros(3, int), cls[:1], box[:1], vol[:1], heuristic_rule="volume")
    assert inst.tolist() == [-100, -100, -100]
    sem, inst = H.heuristic_labels(pts2, np.zeros(3, int), cls[:1], box[:1], vol[:1], box2mask=True)
    assert inst.tolist() == [0, 0, 0]
