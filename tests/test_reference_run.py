"""Fixtures produced by EXECUTING the reference's gen_ps_utils.py / eval_ps_labels.py in the build
container (tests/golden/make_ref_golden.py; torch_scatter and the gpytorch fit replaced by stand-ins
documented there).  They pin the scene-level oracle, the host mirror and - on the GPU - the CUDA path
to the reference's own code.  Nothing here reads /root/reference."""
import os

import numpy as np
import pytest
import torch

from gapro_b200 import synthetic
from gapro_b200.gen_ps import synthetic_inputs, to_scene_inputs

from .conftest import oracle_args

GOLD = os.path.join(os.path.dirname(__file__), "golden")
SCENES = [("tiny", 3, 5), ("small", 4, 6)]
HEUR = [(ds, rule) for ds in ("scannetv2", "s3dis") for rule in ("volume", "dist", "none")]


@pytest.fixture(scope="module")
def ref():
    return np.load(os.path.join(GOLD, "ref_outputs.npz"))


def _inputs(name, seed):
    scene = synthetic.make_scene(seed, name)
    return scene, synthetic_inputs(scene)


@pytest.mark.parametrize("name,seed,nseed", SCENES)
def test_oracle_fixture_equals_reference_run(ref, name, seed, nseed):
    """scene_<name>.npz (oracle output, reproduced by test_oracle_reproduces_golden_scene) is bit-identical
    to what the reference's gen_pseudo_label_gaussian_process returned on the same inputs and draws."""
    orc = np.load(os.path.join(GOLD, f"scene_{name}.npz"))
    assert str(orc["digest"]) == str(ref[f"{name}_digest"])
    assert int(orc["n_regions"]) == int(ref[f"{name}_gp_regions"])
    for k in ("sem", "inst", "prob", "mu", "var"):
        a, b = ref[f"{name}_gp_{k}"], orc[k]
        assert a.dtype == b.dtype and np.array_equal(a, b), k


@pytest.mark.parametrize("name,seed,nseed", SCENES)
def test_heuristic_oracle_equals_reference_run(ref, name, seed, nseed):
    from oracle import heuristic_oracle as H
    _, inp = _inputs(name, seed)
    a = oracle_args(inp)
    args = (a[0], a[2], a[3], a[4], a[5])
    for ds in ("scannetv2", "s3dis"):
        sem, inst = H.heuristic_labels(*args, dataset_name=ds, box2mask=True)
        assert np.array_equal(sem, ref[f"{name}_b2m_{ds}_sem"]) and np.array_equal(inst, ref[f"{name}_b2m_{ds}_inst"])
    for ds, rule in HEUR:
        sem, inst = H.heuristic_labels(*args, dataset_name=ds, heuristic_rule=rule)
        assert np.array_equal(sem, ref[f"{name}_heur_{ds}_{rule}_sem"]), (ds, rule)
        assert np.array_equal(inst, ref[f"{name}_heur_{ds}_{rule}_inst"]), (ds, rule)


@pytest.mark.parametrize("name,seed,nseed", SCENES)
def test_host_mirror_equals_reference_run(ref, name, seed, nseed):
    from gapro_b200.eval_ps_labels import get_miou_scene
    from gapro_b200.gen_ps_utils import batch_giou_cross, getInstanceInfo, is_box1_in_box2
    scene, inp = _inputs(name, seed)
    info = getInstanceInfo(inp["xyz"], instance_label=scene.inst.copy(), semantic_label=scene.sem.copy())
    assert info[0] == int(ref[f"{name}_info_num"])
    assert np.array_equal(info[1], ref[f"{name}_info_cls"])
    assert np.array_equal(info[2], ref[f"{name}_info_box"]) and np.array_equal(info[3], ref[f"{name}_info_vol"])
    corners = float(np.where(info[4] == -100.0, 0.0, info[4]).astype(np.float64).sum())
    assert corners == float(ref[f"{name}_info_corners_sum"])
    box = torch.from_numpy(inp["instance_box"]).float()
    iou, giou = batch_giou_cross(box, box)
    assert np.array_equal(iou.numpy(), ref[f"{name}_iou"]) and np.array_equal(giou.numpy(), ref[f"{name}_giou"])
    nest = np.array([[bool(is_box1_in_box2(a, b, offset=0.1)) for b in box] for a in box])
    assert np.array_equal(nest, ref[f"{name}_nest"])
    sem_gt = torch.from_numpy(scene.sem.copy()).int()
    inst_gt = torch.from_numpy(scene.inst.copy()).int()
    sem_gt[sem_gt != -100] -= 2
    sem_gt[(sem_gt == -1) | (sem_gt == -2)] = 18
    ious = get_miou_scene(sem_gt.long(), inst_gt.long(), torch.from_numpy(ref[f"{name}_gp_sem"]).long(),
                          torch.from_numpy(ref[f"{name}_gp_inst"]).long())
    assert np.allclose(ious.numpy(), ref[f"{name}_miou"], rtol=0, atol=1e-6)


def _hand_items():
    from tests.golden.hand_cases import hand_cases
    return [(i, cname, thr) for i, cname in enumerate(hand_cases()) for thr in (0.999, 0.5)]


def _deep_inputs():
    scene = synthetic.make_scene(8, synthetic.SceneConfig(n_points=6_000, n_objects=6, s_target=300, overlap=0.5,
                                                          n_nested=1, feat_dim=32))
    return synthetic_inputs(scene, use_deepfeat=True)


@pytest.mark.parametrize("i,cname,thr", _hand_items())
def test_oracle_equals_reference_run_on_hand_cases(ref, i, cname, thr):
    """Every branch of the pair loop (nesting both ways, IoU skip + volume fallback, strict merge in event order,
    boxes without superpoints of their own, walls), bit for bit against the reference's loop."""
    from oracle import gen_ps_oracle as O
    from tests.golden.hand_cases import hand_cases
    inp = hand_cases()[cname]
    res, dbg = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=thr, noise_seed=100 + i, return_debug=True)
    k = f"hand_{cname}_{thr}"
    assert len(dbg["regions"]) == int(ref[k + "_regions"])
    for got, key in zip(res, ("sem", "inst", "prob", "mu", "var")):
        want = ref[f"{k}_{key}"]
        assert got.dtype == want.dtype and np.array_equal(got, want), key


def test_oracle_equals_reference_run_on_deep_features(ref):
    from oracle import gen_ps_oracle as O
    res, dbg = O.gen_pseudo_label_oracle(*oracle_args(_deep_inputs()), noise_seed=11, return_debug=True)
    assert len(dbg["regions"]) == int(ref["deep_regions"]) > 0
    for got, key in zip(res, ("sem", "inst", "prob", "mu", "var")):
        assert np.array_equal(got, ref[f"deep_{key}"]), key


def _many_box_inputs():
    from tests.golden.make_ref_golden import MANY_BOXES
    return synthetic_inputs(synthetic.make_scene(23, MANY_BOXES))


def test_oracle_equals_reference_run_with_more_than_32_boxes(ref):
    from oracle import gen_ps_oracle as O
    inp = _many_box_inputs()
    assert len(inp["instance_box"]) + len(inp["wall_box"]) + 1 > 32
    res, dbg = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999, noise_seed=13, return_debug=True)
    assert len(dbg["regions"]) == int(ref["many_regions"]) > 30
    for got, key in zip(res, ("sem", "inst", "prob", "mu", "var")):
        assert np.array_equal(got, ref[f"many_{key}"]), key


def test_confusion_matrix_equals_reference_run(ref):
    from gapro_b200.eval_ps_labels import get_scene_sem_conf
    conf = get_scene_sem_conf(torch.from_numpy(ref["conf_gt"]), torch.from_numpy(ref["conf_ps"]))
    assert np.array_equal(conf.numpy(), ref["conf_matrix"])


def test_wall_boxes_equal_reference_run(ref, tmp_path):
    """scannet_planes.get_wall_boxes on the files the reference was run on (planes json + axis alignment)."""
    from gapro_b200 import scannet_planes
    (tmp_path / "planes").mkdir()
    (tmp_path / "tf" / "scene0000_00").mkdir(parents=True)
    (tmp_path / "planes" / "scene0000_00.json").write_text(str(ref["planes_json"]))
    (tmp_path / "tf" / "scene0000_00" / "scene0000_00.txt").write_text(str(ref["planes_align"]))
    cls, box, vol = scannet_planes.get_wall_boxes("scene0000_00", planes_root=str(tmp_path / "planes"),
                                                  transform_root=str(tmp_path / "tf"))
    assert len(ref["planes_box"]) >= 4
    assert np.array_equal(np.asarray(cls), ref["planes_cls"])
    assert np.array_equal(np.asarray(box), ref["planes_box"]) and np.array_equal(np.asarray(vol), ref["planes_vol"])


def test_host_io_and_oracle_equal_reference_cli_run(ref, tmp_path):
    """The reference's gen_ps.py was executed as a script on this dataset tree (make_ref_golden.py); our loader
    (disk layout, alignment, boxes, wall boxes from the planes file) feeding the oracle must reproduce the files it
    saved, bit for bit."""
    from gapro_b200.gen_ps import DATA_ROOT, load_scene
    from oracle import gen_ps_oracle as O
    from tests.golden.hand_cases import CLI_SCANS, cli_noise_seed, write_cli_dataset
    write_cli_dataset(tmp_path)
    root = tmp_path / DATA_ROOT
    for scan in (CLI_SCANS[0], CLI_SCANS[2]):
        inp, _, _ = load_scene(str(root / "train" / f"{scan}_inst_nostuff.pth"), scan, data_root=str(root))
        assert (len(inp["wall_box"]) == 4) == (scan == CLI_SCANS[0])
        res = O.gen_pseudo_label_oracle(*oracle_args(inp), thresh_spp_occu=0.999, noise_seed=cli_noise_seed(scan))
        for got, key in zip(res, ("sem", "inst", "prob", "mu", "var")):
            want = ref[f"cli_{scan}_{key}"]
            assert got.dtype == want.dtype and np.array_equal(got, want), (scan, key)


# ------------------------------------------------------------------------------------------------
# CUDA path against the reference run
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name,seed,nseed", SCENES)
def test_cuda_gp_path_equals_reference_run(dev, lib, ref, name, seed, nseed):
    """Labels bit-exact, prob / mu / var within rel 1e-4 of the reference run (the two differ only in the
    float64 summation order inside the GP fit; every region of these scenes has a posterior margin far
    above the 1e-6 epsilon, so no point is excluded)."""
    from gapro_b200.gen_ps_utils import gen_pseudo_label_gaussian_process
    _, inp = _inputs(name, seed)
    sc = to_scene_inputs(inp, dev)
    sem, inst, prob, mu, var = gen_pseudo_label_gaussian_process(
        sc.coords_float, sc.mask_feats, sc.spp, sc.instance_cls, sc.instance_box, sc.instance_box_volume, sc.wall_box,
        sc.wall_box_volume, instance_classes=18, dataset_name="scannetv2", ground_h=0.1, training_iter=50,
        thresh_spp_occu=0.999, noise_seed=nseed)
    assert sem.dtype == torch.int32 and inst.dtype == torch.int32
    assert np.array_equal(sem.cpu().numpy(), ref[f"{name}_gp_sem"])
    assert np.array_equal(inst.cpu().numpy(), ref[f"{name}_gp_inst"])
    for got, key in ((prob, "prob"), (mu, "mu"), (var, "var")):
        want = ref[f"{name}_gp_{key}"]
        assert got.shape == want.shape
        assert np.allclose(got.cpu().numpy(), want, rtol=1e-4, atol=1e-7), key


@pytest.mark.gpu
@pytest.mark.parametrize("name,seed,nseed", SCENES)
def test_cuda_heuristic_labelers_equal_reference_run(dev, lib, ref, name, seed, nseed):
    from gapro_b200.gen_ps_utils import gen_pseudo_label, gen_pseudo_label_box2mask
    _, inp = _inputs(name, seed)
    sc = to_scene_inputs(inp, dev)
    args = (sc.coords_float, sc.spp, sc.instance_cls, sc.instance_box, sc.instance_box_volume)
    for ds in ("scannetv2", "s3dis"):
        sem, inst = gen_pseudo_label_box2mask(*args, instance_classes=18, dataset_name=ds)
        assert np.array_equal(sem.cpu().numpy(), ref[f"{name}_b2m_{ds}_sem"]), ds
        assert np.array_equal(inst.cpu().numpy(), ref[f"{name}_b2m_{ds}_inst"]), ds
    for ds, rule in HEUR:
        sem, inst = gen_pseudo_label(*args, instance_classes=18, dataset_name=ds, heuristic_rule=rule)
        assert np.array_equal(sem.cpu().numpy(), ref[f"{name}_heur_{ds}_{rule}_sem"]), (ds, rule)
        assert np.array_equal(inst.cpu().numpy(), ref[f"{name}_heur_{ds}_{rule}_inst"]), (ds, rule)


def _run_cuda(dev, inp, thr, nseed):
    from gapro_b200.gen_ps_utils import gen_pseudo_label_gaussian_process
    sc = to_scene_inputs(inp, dev)
    kw = {} if thr is None else dict(thresh_spp_occu=thr)
    return gen_pseudo_label_gaussian_process(
        sc.coords_float, sc.mask_feats, sc.spp, sc.instance_cls, sc.instance_box, sc.instance_box_volume, sc.wall_box,
        sc.wall_box_volume, noise_seed=nseed, **kw)


def _check_cuda(res, ref, prefix):
    sem, inst, prob, mu, var = res
    assert np.array_equal(sem.cpu().numpy(), ref[prefix + "_sem"])
    assert np.array_equal(inst.cpu().numpy(), ref[prefix + "_inst"])
    for got, key in ((prob, "prob"), (mu, "mu"), (var, "var")):
        assert np.allclose(got.cpu().numpy(), ref[f"{prefix}_{key}"], rtol=1e-4, atol=1e-7), key


@pytest.mark.gpu
@pytest.mark.parametrize("i,cname,thr", _hand_items())
def test_cuda_equals_reference_run_on_hand_cases(dev, lib, ref, i, cname, thr):
    from tests.golden.hand_cases import hand_cases
    _check_cuda(_run_cuda(dev, hand_cases()[cname], thr, 100 + i), ref, f"hand_{cname}_{thr}")


@pytest.mark.gpu
def test_cuda_equals_reference_run_on_deep_features(dev, lib, ref):
    _check_cuda(_run_cuda(dev, _deep_inputs(), None, 11), ref, "deep")


@pytest.mark.gpu
def test_cuda_cli_equals_reference_cli_run(dev, lib, ref, tmp_path, monkeypatch):
    """python -m gapro_b200.gen_ps against the files the reference's gen_ps.py script saved for the same dataset
    tree: same names, same 5-tuple of numpy arrays (dtypes, shapes; mu / var per SUPERPOINT), the resume rule."""
    from gapro_b200 import gen_ps
    from tests.golden.hand_cases import CLI_SCANS, CLI_SEED, write_cli_dataset
    write_cli_dataset(tmp_path)
    monkeypatch.chdir(tmp_path)
    for k, v in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")):
        monkeypatch.setenv(k, v)
    (tmp_path / "out").mkdir()
    torch.save(("sentinel",), str(tmp_path / "out" / f"{CLI_SCANS[1]}.pth"))
    gen_ps.main(["--save_folder", "out", "--seed", str(CLI_SEED), "--eval_pslabel"])
    assert torch.load(str(tmp_path / "out" / f"{CLI_SCANS[1]}.pth"), weights_only=False) == ("sentinel",)
    for scan in (CLI_SCANS[0], CLI_SCANS[2]):
        tup = torch.load(str(tmp_path / "out" / f"{scan}.pth"), weights_only=False)
        assert len(tup) == 5 and all(isinstance(a, np.ndarray) for a in tup)
        for a, key in zip(tup, ("sem", "inst", "prob", "mu", "var")):
            want = ref[f"cli_{scan}_{key}"]
            assert a.dtype == want.dtype and a.shape == want.shape, (scan, key)
            if key in ("sem", "inst"):
                assert np.array_equal(a, want), (scan, key)
            else:
                assert np.allclose(a, want, rtol=1e-4, atol=1e-7), (scan, key)


@pytest.mark.gpu
def test_cuda_equals_reference_run_with_more_than_32_boxes(dev, lib, ref):
    """39 boxes (two occupancy words per superpoint), 41 GP regions, nesting events, contested merges: final labels
    against the reference run (minimum posterior margin of the scene 1.5e-3, minimum merge gap 1.2e-2)."""
    _check_cuda(_run_cuda(dev, _many_box_inputs(), 0.999, 13), ref, "many")
