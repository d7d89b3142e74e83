"""CLI and per-scene host steps of the pseudo-label generator — the counterpart of
/root/reference/gapro/gen_ps.py.

    python -m gapro_b200.gen_ps [--save_folder D] [--use_deepfeat] [--deepfeat_folder D] [--eval_pslabel]

keeps the reference flags (gen_ps.py:15-21), the relative dataset layout
(dataset/scannetv2/{train,superpoints,scans_transform}), the output naming
<save_folder>/<scan_name>.pth with scan_name = basename[:12] (gen_ps.py:37-39), the
skip-if-exists resume rule (:39-41) and the saved tuple of five numpy arrays (:126-132).
Additions: scenes are processed in batches on the GPU, and when launched under torchrun
the sorted scene list is sharded across ranks (one process per GPU, no per-step collective;
one final gather of per-scene metadata).
"""
from __future__ import annotations

import argparse
import os
import os.path as osp
import time
import zlib
from glob import glob

import numpy as np
import torch

from .engine import SceneInputs
from .gen_ps_utils import gen_pseudo_labels_batch, getInstanceInfo
from .sharding import gather_records, shard_scenes

DATA_ROOT = "dataset/scannetv2"


def read_axis_align_matrix(meta_file: str) -> np.ndarray:
    """The 'axisAlignment = ...' line of a ScanNet scene meta file (gen_ps.py:58-64)."""
    with open(meta_file) as f:
        for line in f:
            if "axisAlignment" in line:
                vals = [float(x) for x in line.split("=", 1)[1].split()]
                return np.array(vals, dtype=np.float64).reshape(4, 4)
    raise ValueError(f"no axisAlignment line in {meta_file}")


def prepare_inputs(xyz, rgb, semantic_label, instance_label, spp, axis_align_matrix, wall_box=None,
                   wall_volume=None, deep_feats=None, dataset_name="scannetv2"):
    """Host steps of gen_ps.py:48-77 on numpy arrays.  NOTE the reference builds the GP features
    from the UN-aligned xyz (concat at :55 happens before the alignment at :66-69); kept."""
    xyz = np.asarray(xyz)
    mask_feats = np.asarray(deep_feats) if deep_feats is not None else np.concatenate([xyz, np.asarray(rgb)], axis=-1)
    pts = np.ones((xyz.shape[0], 4))
    pts[:, 0:3] = xyz[:, 0:3]
    xyz_al = np.dot(pts, np.asarray(axis_align_matrix).transpose())[:, :3]
    info = getInstanceInfo(xyz_al, instance_label=instance_label, semantic_label=semantic_label,
                           dataset_name=dataset_name)
    if info is None:
        raise ValueError("scene has no labelled instance (getInstanceInfo returned None, gen_ps_utils.py:229-230)")
    _, instance_cls, instance_box, instance_box_volume, _ = info
    return dict(xyz=xyz_al, mask_feats=mask_feats, spp=np.asarray(spp), instance_cls=instance_cls,
                instance_box=instance_box, instance_box_volume=instance_box_volume,
                wall_box=wall_box if wall_box is not None else [], wall_volume=wall_volume if wall_volume is not None else [])


def to_scene_inputs(inp: dict, device, noise_seed=None, pin=False) -> SceneInputs:
    """Host -> device boundary of gen_ps.py:79-89 (dtypes as there: xyz stays float64, boxes and
    features become float32, classes int64)."""
    def dev(a, dtype):
        t = torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
        if pin:
            t = t.pin_memory()
        return t.to(device, non_blocking=pin)

    wall_box, wall_vol = inp["wall_box"], inp["wall_volume"]
    has_wall = len(wall_box) > 0
    return SceneInputs(
        coords_float=dev(inp["xyz"], torch.float64),
        mask_feats=dev(inp["mask_feats"], torch.float32),
        spp=dev(inp["spp"], torch.int64),
        instance_cls=dev(inp["instance_cls"], torch.int64),
        instance_box=dev(inp["instance_box"], torch.float32),
        instance_box_volume=dev(inp["instance_box_volume"], torch.float32),
        wall_box=dev(wall_box, torch.float32) if has_wall else [],
        wall_box_volume=dev(wall_vol, torch.float32) if has_wall else [],
        noise_seed=noise_seed,
    )


def synthetic_inputs(scene, use_deepfeat=False) -> dict:
    """gapro_b200.synthetic.Scene -> the dict prepare_inputs returns."""
    return prepare_inputs(scene.xyz_raw, scene.rgb, scene.sem, scene.inst, scene.spp, scene.axis_align,
                          wall_box=scene.wall_box, wall_volume=scene.wall_volume,
                          deep_feats=scene.deep_feats if use_deepfeat else None)


def load_scene(filename, scan_name, use_deepfeat=False, deepfeat_folder=None, data_root=DATA_ROOT):
    """Disk -> host arrays for one scene (gen_ps.py:45-77): scene tuple, superpoints, optional deep
    features, axis alignment, instance boxes, optional wall boxes.  Returns (inputs, sem, inst)."""
    from .scannet_planes import get_wall_boxes
    xyz, rgb, sem, inst = torch.load(filename, weights_only=False)
    spp = torch.load(osp.join(data_root, "superpoints", scan_name + ".pth"), weights_only=False)
    deep = torch.load(osp.join(deepfeat_folder, scan_name + ".pth"), weights_only=False) if use_deepfeat else None
    A = read_axis_align_matrix(osp.join(data_root, "scans_transform", scan_name, scan_name + ".txt"))
    _, wall_box, wall_vol = get_wall_boxes(scan_name, planes_root=osp.join(data_root, "scannet_planes"),
                                           transform_root=osp.join(data_root, "scans_transform"))
    return prepare_inputs(xyz, rgb, sem, inst, spp, A, wall_box, wall_vol, deep), sem, inst


def save_pseudo_labels(path, result, per_point_uncertainty=False, spp_dense=None):
    """torch.save of the 5-tuple of numpy arrays (gen_ps.py:126-132).  The reference saves mu/var
    per SUPERPOINT although its consumers index them per point;
    per_point_uncertainty=True saves mu[spp], var[spp] instead."""
    sem, inst, prob, mu, var = result
    if per_point_uncertainty:
        mu, var = mu[spp_dense], var[spp_dense]
    arrs = (sem.int().cpu().numpy(), inst.int().cpu().numpy(), prob.cpu().numpy(), mu.cpu().numpy(), var.cpu().numpy())
    tmp = path + ".tmp%d" % os.getpid()
    torch.save(arrs, tmp)
    os.replace(tmp, path)


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def main(argv=None):
    parser = argparse.ArgumentParser("GaPro_GenPS")
    parser.add_argument("--save_folder", type=str, default=osp.join(DATA_ROOT, "gaussian_process_kl_pseudo_labels"))
    parser.add_argument("--use_deepfeat", action="store_true")
    parser.add_argument("--deepfeat_folder", type=str, default=osp.join(DATA_ROOT, "pretrain_maskfeats2"))
    parser.add_argument("--eval_pslabel", action="store_true")
    # additions
    parser.add_argument("--batch_scenes", type=int, default=8, help="scenes per GPU pass")
    parser.add_argument("--seed", type=int, default=None, help="seed of the GP initialisation noise")
    parser.add_argument("--load_workers", type=int, default=8, help="host threads reading / preparing the next batch")
    parser.add_argument("--per_point_uncertainty", action="store_true",
                        help="save mu/var broadcast to points (what the ISBNet/SPFormer loaders index)")
    args = parser.parse_args(argv)

    rank, world, local = _dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    os.makedirs(args.save_folder, exist_ok=True)

    filenames = sorted(glob(osp.join(DATA_ROOT, "train", "*_inst_nostuff.pth")))
    todo = []
    for fn in filenames:
        scan = osp.basename(fn)[:12]
        if not osp.exists(osp.join(args.save_folder, scan + ".pth")):   # resume rule, gen_ps.py:39-41
            todo.append((fn, scan))
    todo = shard_scenes(todo, rank, world)

    from .eval_ps_labels import get_miou_scene
    ious, meta = [], []
    t0 = time.time()
    chunks = [todo[i:i + args.batch_scenes] for i in range(0, len(todo), args.batch_scenes)]
    # host side of the NEXT batch (disk reads, alignment, boxes) runs on worker threads while the GPU
    # works on the current one
    from concurrent.futures import ThreadPoolExecutor
    pool = ThreadPoolExecutor(max_workers=max(1, min(args.load_workers, args.batch_scenes)))
    submit = lambda chunk: [pool.submit(load_scene, fn, scan, args.use_deepfeat, args.deepfeat_folder)
                            for fn, scan in chunk]
    pending = submit(chunks[0]) if chunks else []
    for ci, chunk in enumerate(chunks):
        loaded = [f.result() for f in pending]
        pending = submit(chunks[ci + 1]) if ci + 1 < len(chunks) else []
        scenes, gts = [], []
        for (fn, scan), (inp, sem, inst) in zip(chunk, loaded):
            seed = None if args.seed is None else (zlib.crc32(scan.encode()) ^ args.seed) & 0x7fffffff
            scenes.append(to_scene_inputs(inp, device, noise_seed=seed, pin=True))
            gts.append((sem, inst))
        results = gen_pseudo_labels_batch(scenes, instance_classes=18, ground_h=0.1, training_iter=50,
                                          thresh_spp_occu=0.999, device=device)    # gen_ps.py:106-110
        for (fn, scan), res, sc, (sem, inst) in zip(chunk, results, scenes, gts):
            if args.eval_pslabel:      # gen_ps.py:116-124
                s = torch.from_numpy(np.asarray(sem)).to(device).int()
                g = torch.from_numpy(np.asarray(inst)).to(device).int()
                s[s != -100] -= 2
                s[(s == -1) | (s == -2)] = 18
                iou = get_miou_scene(s.long(), g.long(), res[0].long(), res[1].long())
                print("miou", iou)
                ious.append(iou)
            dense = torch.unique(sc.spp, return_inverse=True)[1] if args.per_point_uncertainty else None
            save_pseudo_labels(osp.join(args.save_folder, scan + ".pth"), res, args.per_point_uncertainty, dense)
            meta.append((scan, int(res[0].numel()), int(res[3].numel())))
    if args.eval_pslabel:
        local_iou = torch.cat(ious) if ious else torch.zeros(0, device=device)
        if world > 1:
            import torch.distributed as dist
            gathered = [None] * world
            dist.all_gather_object(gathered, local_iou.cpu())
            local_iou = torch.cat(gathered)
        if rank == 0 and local_iou.numel():
            print("Mean instance iou of pseudo labels", torch.mean(local_iou.float()).item())
    if world > 1:
        import torch.distributed as dist
        records = gather_records(meta, world)       # the one collective: label metadata
        if rank == 0:
            print(f"{len(records)} scenes / {sum(m[1] for m in records)} points labelled on {world} GPUs "
                  f"in {time.time() - t0:.1f}s")
        dist.destroy_process_group()
    if rank == 0:
        print("Finish")


if __name__ == "__main__":
    main()
