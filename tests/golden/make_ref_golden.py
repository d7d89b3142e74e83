"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE in this container.

    python tests/golden/make_ref_golden.py          # needs /root/reference (absent on the GPU box)

`/root/reference/gapro/gen_ps_utils.py` and `eval_ps_labels.py` are imported as they lie (nothing is
copied) and executed on CPU on the same seeded synthetic scenes the tests use.  Two of their
dependencies do not exist here, so the script installs stand-ins for them before the import:

  * `torch_scatter` (a compiled extension): `scatter`, `scatter_add`, `scatter_min` written with plain
    sequential loops over the index, i.e. the accumulation order of torch_scatter's CPU kernels
    (one pass over the source in index order; `scatter_min` keeps the FIRST minimum and reports
    `src.size(dim)` for empty groups).  This is the only restated arithmetic in the run.
  * `gaussian_process_utils.fit_gp_spp` (needs gpytorch): replaced by the oracle's GP fit
    (`oracle.gp_oracle.fit_region_autograd`, fp64 policy), fed with standard-normal draws from one
    numpy Generator per scene in call order - the same convention as `gen_pseudo_label_oracle`.

Everything else - densification, floor slab, box concatenation, containment, pooling, occupancy, IoU,
the pair state machine, merge, fallback, broadcast, the heuristic labelers, getInstanceInfo,
batch_giou_cross, the IoU evaluation - is the reference's code.  The fixtures therefore PIN the
scene-level oracle (and through it the CUDA path) to the reference; the inside of the GP fit stays
pinned only by the derivation in oracle/gp_oracle.py ("parity unpinned" for that part).
"""
import hashlib
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference/gapro"

from gapro_b200 import synthetic                       # noqa: E402
from gapro_b200.gen_ps import synthetic_inputs         # noqa: E402
from oracle import gp_oracle as G                      # noqa: E402

MANY_BOXES = synthetic.SceneConfig(n_points=14_000, n_objects=34, s_target=600, overlap=0.5, n_nested=3)


# ------------------------------------------------------------------------------------------------
# stand-in for the absent torch_scatter extension (sequential, index order)
# ------------------------------------------------------------------------------------------------
def _prep(src, index, dim):
    dim = dim if dim >= 0 else src.dim() + dim
    if index.dim() != src.dim():                        # torch_scatter broadcasts a 1-d index along `dim`
        shape = [1] * src.dim()
        shape[dim] = -1
        index = index.reshape(shape).expand_as(src)
    s = src.movedim(dim, 0).contiguous()
    i = index.movedim(dim, 0).contiguous()
    return dim, s, i


def _size(index, dim_size):
    return int(dim_size) if dim_size is not None else (int(index.max()) + 1 if index.numel() else 0)


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    assert out is None
    dim, s, i = _prep(src, index, dim)
    n = _size(index, dim_size)
    flat_s, flat_i = s.reshape(s.shape[0], -1), i.reshape(i.shape[0], -1)
    if reduce in ("sum", "add", "mean"):
        acc = torch.zeros((n, flat_s.shape[1]), dtype=src.dtype)
        cnt = torch.zeros((n, flat_s.shape[1]), dtype=torch.long)
        a, c = acc.numpy(), cnt.numpy()
        sv, iv = flat_s.numpy(), flat_i.numpy()
        cols = np.arange(sv.shape[1])
        for r in range(sv.shape[0]):                    # one pass in index order, like the CPU kernel
            a[iv[r], cols] += sv[r]
            c[iv[r], cols] += 1
        if reduce == "mean":
            c = np.maximum(c, 1)
            if src.dtype.is_floating_point:
                acc = torch.from_numpy(a / c.astype(a.dtype))
            else:
                acc = torch.from_numpy(a // c)          # rounding_mode="floor" for integers
        return acc.reshape((n,) + tuple(s.shape[1:])).movedim(0, dim)
    raise NotImplementedError(reduce)


def scatter_add(src, index, dim=-1, out=None, dim_size=None):
    return scatter(src, index, dim, out, dim_size, "sum")


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    assert out is None
    dim, s, i = _prep(src, index, dim)
    assert s.dim() == 1
    n = _size(index, dim_size)
    sv, iv = s.numpy(), i.numpy()
    big = np.finfo(sv.dtype).max if sv.dtype.kind == "f" else np.iinfo(sv.dtype).max
    val = np.full(n, big, dtype=sv.dtype)
    arg = np.full(n, sv.shape[0], dtype=np.int64)
    for r in range(sv.shape[0]):
        if sv[r] < val[iv[r]]:
            val[iv[r]] = sv[r]
            arg[iv[r]] = r
    val[arg == sv.shape[0]] = 0
    return torch.from_numpy(val), torch.from_numpy(arg)


def install_stand_ins(fit_state):
    ts = types.ModuleType("torch_scatter")
    ts.scatter, ts.scatter_add, ts.scatter_min = scatter, scatter_add, scatter_min
    sys.modules["torch_scatter"] = ts

    def fit_gp_spp(coords_float_spp, feats_spp, b1_inds, b2_inds, intersect_inds, training_iter=50):
        b1, b2 = b1_inds.reshape(-1).numpy(), b2_inds.reshape(-1).numpy()
        f = feats_spp.numpy()
        noise = fit_state["rng"].standard_normal(len(b1) + len(b2)).astype(np.float32)
        r = G.fit_region_autograd(np.concatenate([f[b1], f[b2]]), len(b1), f[intersect_inds.reshape(-1).numpy()],
                                  noise, iters=training_iter)
        fit_state["calls"] += 1
        t = torch.from_numpy
        return t(r["prob"]), t(r["conf"]), t(r["label"]), t(r["mu"]), t(r["var"])

    gp = types.ModuleType("gaussian_process_utils")
    gp.fit_gp_spp = fit_gp_spp
    sys.modules["gaussian_process_utils"] = gp


def digest(arrays):
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(np.asarray(a)).tobytes())
    return h.hexdigest()


def as_ref_tensors(inp):
    """The tensors gen_ps.py:81-92 builds (on CPU instead of .cuda())."""
    t = torch.from_numpy
    return dict(xyz=t(inp["xyz"]), mask_feats=t(np.ascontiguousarray(inp["mask_feats"])).float(), spp=t(inp["spp"]),
                instance_cls=t(inp["instance_cls"]).long(), instance_box=t(inp["instance_box"]).float(),
                instance_box_volume=t(inp["instance_box_volume"]).float(),
                wall_box=t(inp["wall_box"]).float() if len(inp["wall_box"]) else inp["wall_box"],
                wall_volume=t(inp["wall_volume"]).float() if len(inp["wall_box"]) else inp["wall_volume"])


def main():
    if not os.path.isdir(REF):
        raise SystemExit("make_ref_golden.py needs the reference checkout at " + REF)
    state = {"rng": None, "calls": 0}
    install_stand_ins(state)
    sys.path.insert(0, REF)
    import gen_ps_utils as R                              # the reference's module, executed as it lies
    import eval_ps_labels as E

    out = {}
    for name, seed, nseed in [("tiny", 3, 5), ("small", 4, 6)]:
        scene = synthetic.make_scene(seed, name)
        inp = synthetic_inputs(scene)
        T = as_ref_tensors(inp)
        key = digest([inp["xyz"], inp["mask_feats"].astype(np.float32), inp["spp"], inp["instance_cls"].astype(np.int64),
                      inp["instance_box"].astype(np.float32), inp["instance_box_volume"].astype(np.float32),
                      inp["wall_box"], inp["wall_volume"]])
        out[f"{name}_digest"] = np.array(key)
        # --- the GP path, with the CLI's arguments (gen_ps.py:97-111)
        state["rng"], state["calls"] = np.random.default_rng(nseed), 0
        sem, inst, prob, mu, var = R.gen_pseudo_label_gaussian_process(
            T["xyz"], T["mask_feats"], T["spp"], T["instance_cls"], T["instance_box"], T["instance_box_volume"],
            T["wall_box"], T["wall_volume"], instance_classes=18, dataset_name="scannetv2", ground_h=0.1,
            training_iter=50, thresh_spp_occu=0.999)
        out[f"{name}_gp_sem"], out[f"{name}_gp_inst"] = sem.int().numpy(), inst.int().numpy()
        out[f"{name}_gp_prob"], out[f"{name}_gp_mu"], out[f"{name}_gp_var"] = prob.numpy(), mu.numpy(), var.numpy()
        out[f"{name}_gp_regions"] = np.array(state["calls"])
        print(f"{name}: GP path, {state['calls']} regions, {int((inst.numpy() >= 0).sum())} labelled points")
        # --- heuristic labelers
        for ds in ("scannetv2", "s3dis"):
            s2, i2 = R.gen_pseudo_label_box2mask(T["xyz"], T["spp"], T["instance_cls"], T["instance_box"],
                                                 T["instance_box_volume"], instance_classes=18, dataset_name=ds)
            out[f"{name}_b2m_{ds}_sem"], out[f"{name}_b2m_{ds}_inst"] = s2.int().numpy(), i2.int().numpy()
            for rule in ("volume", "dist", "none"):
                s3, i3 = R.gen_pseudo_label(T["xyz"], T["spp"], T["instance_cls"], T["instance_box"],
                                            T["instance_box_volume"], instance_classes=18, dataset_name=ds,
                                            heuristic_rule=rule)
                out[f"{name}_heur_{ds}_{rule}_sem"], out[f"{name}_heur_{ds}_{rule}_inst"] = s3.int().numpy(), i3.int().numpy()
        # --- host helpers
        info = R.getInstanceInfo(inp["xyz"], instance_label=scene.inst.copy(), semantic_label=scene.sem.copy())
        out[f"{name}_info_num"] = np.array(info[0])
        out[f"{name}_info_cls"], out[f"{name}_info_box"], out[f"{name}_info_vol"] = info[1], info[2], info[3]
        out[f"{name}_info_corners_sum"] = np.array(float(np.where(info[4] == -100.0, 0.0, info[4]).astype(np.float64).sum()))
        iou, giou = R.batch_giou_cross(T["instance_box"], T["instance_box"])
        out[f"{name}_iou"], out[f"{name}_giou"] = iou.numpy(), giou.numpy()
        nest = np.array([[bool(R.is_box1_in_box2(a, b, offset=0.1)) for b in T["instance_box"]] for a in T["instance_box"]])
        out[f"{name}_nest"] = nest
        # --- evaluation (gen_ps.py:116-123)
        sem_gt = torch.from_numpy(scene.sem.copy()).int()
        inst_gt = torch.from_numpy(scene.inst.copy()).int()
        sem_gt[sem_gt != -100] -= 2
        sem_gt[(sem_gt == -1) | (sem_gt == -2)] = 18
        cuda = torch.Tensor.cuda                         # eval_ps_labels.py:102 hard-codes .cuda(); no GPU here
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            ious = E.get_miou_scene(sem_gt.long(), inst_gt.long(), sem.long(), inst.long())
        finally:
            torch.Tensor.cuda = cuda
        out[f"{name}_miou"] = ious.numpy() if torch.is_tensor(ious) else np.asarray(ious)
        print(f"{name}: heuristics + helpers done, mean IoU of the GP labels {float(np.mean(out[f'{name}_miou'])):.4f}")
    # --- hand-built geometries: every branch of the pair loop, real (oracle) GP fit, several thresholds
    from tests.golden.hand_cases import hand_cases
    for i, (cname, inp) in enumerate(hand_cases().items()):
        T = as_ref_tensors(inp)
        for thr in (0.999, 0.5):
            state["rng"], state["calls"] = np.random.default_rng(100 + i), 0
            sem, inst, prob, mu, var = R.gen_pseudo_label_gaussian_process(
                T["xyz"], T["mask_feats"], T["spp"], T["instance_cls"], T["instance_box"], T["instance_box_volume"],
                T["wall_box"], T["wall_volume"], instance_classes=18, dataset_name="scannetv2", ground_h=0.1,
                training_iter=50, thresh_spp_occu=thr)
            k = f"hand_{cname}_{thr}"
            out[k + "_sem"], out[k + "_inst"], out[k + "_prob"] = sem.int().numpy(), inst.int().numpy(), prob.numpy()
            out[k + "_mu"], out[k + "_var"], out[k + "_regions"] = mu.numpy(), var.numpy(), np.array(state["calls"])
            print(f"hand case {cname} thresh {thr}: {state['calls']} regions, labels {np.unique(inst.numpy()).tolist()}")
    # --- deep features (D = 32) and the function's default threshold
    scene = synthetic.make_scene(8, synthetic.SceneConfig(n_points=6_000, n_objects=6, s_target=300, overlap=0.5,
                                                          n_nested=1, feat_dim=32))
    inp = synthetic_inputs(scene, use_deepfeat=True)
    T = as_ref_tensors(inp)
    state["rng"], state["calls"] = np.random.default_rng(11), 0
    sem, inst, prob, mu, var = R.gen_pseudo_label_gaussian_process(
        T["xyz"], T["mask_feats"], T["spp"], T["instance_cls"], T["instance_box"], T["instance_box_volume"],
        T["wall_box"], T["wall_volume"])
    out["deep_sem"], out["deep_inst"], out["deep_prob"] = sem.int().numpy(), inst.int().numpy(), prob.numpy()
    out["deep_mu"], out["deep_var"], out["deep_regions"] = mu.numpy(), var.numpy(), np.array(state["calls"])
    print(f"deep-feature scene: D={inp['mask_feats'].shape[1]}, {state['calls']} regions")
    # --- more than 32 boxes (two 32-bit occupancy words on the device side), 41 GP regions, 7 nesting events;
    # minimum posterior margin 1.5e-3, minimum gap of a contested merge 1.2e-2
    inp = synthetic_inputs(synthetic.make_scene(23, MANY_BOXES))
    T = as_ref_tensors(inp)
    state["rng"], state["calls"] = np.random.default_rng(13), 0
    sem, inst, prob, mu, var = R.gen_pseudo_label_gaussian_process(
        T["xyz"], T["mask_feats"], T["spp"], T["instance_cls"], T["instance_box"], T["instance_box_volume"],
        T["wall_box"], T["wall_volume"], thresh_spp_occu=0.999)
    out["many_sem"], out["many_inst"], out["many_prob"] = sem.int().numpy(), inst.int().numpy(), prob.numpy()
    out["many_mu"], out["many_var"], out["many_regions"] = mu.numpy(), var.numpy(), np.array(state["calls"])
    print(f"many-box scene: {len(inp['instance_box']) + len(inp['wall_box']) + 1} boxes, {state['calls']} regions")
    # --- semantic confusion matrix (eval_ps_labels.py:152-172)
    cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        gt = torch.from_numpy(np.random.default_rng(5).integers(0, 19, 4000)).long()
        gt[::17] = -100
        ps = gt.clone()
        flip = torch.from_numpy(np.random.default_rng(6).random(4000) < 0.2)
        ps[flip] = torch.from_numpy(np.random.default_rng(7).integers(0, 19, int(flip.sum()))).long()
        ps[::29] = -100
        conf = E.get_scene_sem_conf(gt.clone(), ps.clone())
    finally:
        torch.Tensor.cuda = cuda
    out["conf_gt"], out["conf_ps"], out["conf_matrix"] = gt.numpy(), ps.numpy(), np.asarray(conf)
    # --- wall boxes from a ScanNet-planes file (scannet_planes.py:177-230); np.int was removed from numpy, restore the alias
    import json
    import tempfile
    import scannet_planes as P
    if not hasattr(np, "int"):
        np.int = int
    rng = np.random.default_rng(21)
    # a slightly rotated room in the raw (y-up) convention, 4 walls + floor + ceiling + a slanted roof quad,
    # a non-planar quad and a triangle
    c, s_ = np.cos(0.3), np.sin(0.3)
    base = np.array([[0, 0], [5.2, 0], [5.2, 3.9], [0, 3.9]]) @ np.array([[c, -s_], [s_, c]]).T + [1.0, -2.0]
    verts = [[float(x), 0.0, float(-y)] for x, y in base] + [[float(x), 2.7, float(-y)] for x, y in base]
    verts += [[2.0, 2.7, 1.0], [3.0, 3.4, 1.0], [3.0, 3.4, 2.5], [2.0, 2.7, 2.5]]          # slanted
    verts += [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [1.0, 1.0, 0.7], [0.0, 1.0, -0.9]]          # not coplanar
    verts = (np.array(verts) + rng.normal(0, 1e-3, (len(verts), 3))).tolist()
    quads = [[0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7], [0, 1, 2, 3], [4, 5, 6, 7],
             [8, 9, 10, 11], [12, 13, 14, 15], [0, 1, 2]]
    plane_text = json.dumps({"verts": verts, "quads": quads})
    A = np.array([[0.94, 0.34, 0, -1.5], [-0.34, 0.94, 0, 2.25], [0, 0, 1, -0.05], [0, 0, 0, 1]])
    align_text = "axisAlignment = " + " ".join(repr(float(v)) for v in A.reshape(-1)) + "\nnumDepthFrames = 1\n"
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.makedirs(os.path.join(tmp, "dataset/scannetv2/scannet_planes"))
        os.makedirs(os.path.join(tmp, "dataset/scannetv2/scans_transform/scene0000_00"))
        open(os.path.join(tmp, "dataset/scannetv2/scannet_planes/scene0000_00.json"), "w").write(plane_text)
        open(os.path.join(tmp, "dataset/scannetv2/scans_transform/scene0000_00/scene0000_00.txt"), "w").write(align_text)
        os.chdir(tmp)
        try:
            wcls, wbox, wvol = P.get_wall_boxes("scene0000_00")
            missing = P.get_wall_boxes("scene0000_01")
        finally:
            os.chdir(cwd)
    assert missing == ([], [], [])
    out["planes_json"], out["planes_align"] = np.array(plane_text), np.array(align_text)
    out["planes_cls"], out["planes_box"], out["planes_vol"] = np.asarray(wcls), np.asarray(wbox), np.asarray(wvol)
    print(f"wall boxes: {len(wbox)} of {len(quads)} quads kept")
    # --- the CLI itself: gen_ps.py run as a script on a dataset/scannetv2 tree.  Environment stand-ins only: no GPU
    # (.cuda() -> identity), torch.load of numpy pickles (weights_only default changed in torch 2.6), and the GP
    # draws re-seeded per scan (hooked on get_wall_boxes, the call gen_ps.py makes once per scan before the labeler)
    import pathlib
    import runpy
    from tests.golden.hand_cases import CLI_SCANS, cli_noise_seed, write_cli_dataset
    real_walls, real_load, cuda = P.get_wall_boxes, torch.load, torch.Tensor.cuda

    def walls_and_reseed(scan_name):
        state["rng"] = np.random.default_rng(cli_noise_seed(scan_name))
        return real_walls(scan_name)

    with tempfile.TemporaryDirectory() as tmp:
        write_cli_dataset(pathlib.Path(tmp))
        os.makedirs(os.path.join(tmp, "out"))
        torch.save(("sentinel",), os.path.join(tmp, "out", CLI_SCANS[1] + ".pth"))      # must be skipped (gen_ps.py:39-41)
        P.get_wall_boxes = walls_and_reseed
        torch.load = lambda f, *a, **k: real_load(f, *a, **{**k, "weights_only": False})
        torch.Tensor.cuda = lambda self, *a, **k: self
        argv = sys.argv
        sys.argv = ["gen_ps.py", "--save_folder", "out", "--eval_pslabel"]
        os.chdir(tmp)
        try:
            runpy.run_path(os.path.join(REF, "gen_ps.py"), run_name="__main__")
        finally:
            os.chdir(cwd)
            sys.argv = argv
            P.get_wall_boxes, torch.load, torch.Tensor.cuda = real_walls, real_load, cuda
        assert real_load(os.path.join(tmp, "out", CLI_SCANS[1] + ".pth"), weights_only=False) == ("sentinel",)
        for scan in (CLI_SCANS[0], CLI_SCANS[2]):
            tup = real_load(os.path.join(tmp, "out", scan + ".pth"), weights_only=False)
            assert len(tup) == 5 and all(isinstance(a, np.ndarray) for a in tup)
            for a, key in zip(tup, ("sem", "inst", "prob", "mu", "var")):
                out[f"cli_{scan}_{key}"] = a
            print(f"CLI {scan}: dtypes {[str(a.dtype) for a in tup]}, shapes {[a.shape for a in tup]}")
    np.savez_compressed(os.path.join(HERE, "ref_outputs.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_outputs.npz"))


if __name__ == "__main__":
    main()
