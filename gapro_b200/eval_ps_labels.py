"""Pseudo-label quality metrics — counterpart of
/root/reference/gapro/eval_ps_labels.py:100-172 (`--eval_pslabel` of the CLI).

The reference builds N x K one-hot matrices and multiplies them.  Device tensors go through the CUDA
library (`gapro_eval_miou_scene`, `gapro_eval_sem_conf`: one integer-atomic pass over the points, exact
counts).  Host tensors (the CPU checks against the reference run in tests/) take the same table from a
torch bincount.
"""
from __future__ import annotations

import torch


def _i32(t):
    return t.to(torch.int32).contiguous()


def _miou_scene_cuda(semantic_label, instance_label, ps_semantic_label, ps_instance_label, n_inst, n_ps):
    from . import _lib
    lib = _lib.load()
    dev = instance_label.device
    stream = torch.cuda.current_stream(dev).cuda_stream
    gs, gi, ps, pi = _i32(semantic_label), _i32(instance_label), _i32(ps_semantic_label), _i32(ps_instance_label)
    ws = torch.empty(lib.gapro_eval_workspace_bytes(n_inst, n_ps), dtype=torch.uint8, device=dev)
    out = torch.empty(n_inst, dtype=torch.float32, device=dev)
    valid = torch.empty(n_inst, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.gapro_eval_miou_scene(gs.data_ptr(), gi.data_ptr(), ps.data_ptr(), pi.data_ptr(), gi.numel(),
                                             n_inst, n_ps, out.data_ptr(), valid.data_ptr(), ws.data_ptr(), ws.numel(),
                                             stream), "gapro_eval_miou_scene")
    return out[valid.bool()]


def _first_point_class(inst, sem, n):
    """class[i] = semantic label of the first point of instance i, -1 if the id is unused
    (eval_ps_labels.py:101-108)."""
    cls = torch.full((n,), -1, dtype=torch.float32, device=inst.device)
    valid = inst >= 0
    idx = torch.nonzero(valid).view(-1)
    if idx.numel() == 0:
        return cls
    first = torch.full((n,), inst.numel(), dtype=torch.long, device=inst.device)
    first.scatter_reduce_(0, inst[idx].long(), idx, reduce="amin", include_self=True)
    used = first < inst.numel()
    cls[used] = sem[first[used]].float()
    return cls


def get_miou_scene(semantic_label, instance_label, ps_semantic_label, ps_instance_label):
    """Best IoU of every GT instance with a pseudo instance of the same class
    (eval_ps_labels.py:100-147).  Returns one value per GT instance id in use."""
    n_inst = int(instance_label.max()) + 1
    n_ps = int(ps_instance_label.max()) + 1
    if n_inst <= 0:
        return torch.zeros(0, device=instance_label.device)
    gt_cls = _first_point_class(instance_label, semantic_label, n_inst)
    ps_cls = _first_point_class(ps_instance_label, ps_semantic_label, max(n_ps, 0))
    if n_ps <= 0:
        return torch.zeros(int((gt_cls >= 0).sum()), device=instance_label.device)
    if instance_label.is_cuda:
        return _miou_scene_cuda(semantic_label, instance_label, ps_semantic_label, ps_instance_label, n_inst, n_ps)
    g = torch.where(instance_label < 0, 0, instance_label + 1).long()
    p = torch.where(ps_instance_label < 0, 0, ps_instance_label + 1).long()
    table = torch.bincount(g * (n_ps + 1) + p, minlength=(n_inst + 1) * (n_ps + 1)).view(n_inst + 1, n_ps + 1)
    inter = table[1:, 1:].float()
    area_g = table[1:, :].sum(1).float()
    area_p = table[:, 1:].sum(0).float()
    union = area_g[:, None] + area_p[None, :] - inter
    ious = inter / (union + 1e-4)                     # cal_iou, eval_ps_labels.py:40
    ious = ious * (gt_cls[:, None] == ps_cls[None, :]).float()
    max_ious = ious.max(dim=1)[0]
    return max_ious[gt_cls >= 0]


def get_scene_sem_conf(semantic_label, ps_semantic_label, num_classes=19):
    """Semantic confusion matrix (eval_ps_labels.py:152-172); unlabelled pseudo points count
    as a wrong neighbouring class."""
    if semantic_label.is_cuda:
        from . import _lib
        lib = _lib.load()
        dev = semantic_label.device
        conf = torch.empty(num_classes * num_classes, dtype=torch.int64, device=dev)
        gs, ps = _i32(semantic_label), _i32(ps_semantic_label)
        with torch.cuda.device(dev):
            _lib.check(lib.gapro_eval_sem_conf(gs.data_ptr(), ps.data_ptr(), gs.numel(), num_classes, conf.data_ptr(),
                                               torch.cuda.current_stream(dev).cuda_stream), "gapro_eval_sem_conf")
        return conf.view(num_classes, num_classes)
    keep = semantic_label != -100
    gt = semantic_label[keep].clone()
    ps = ps_semantic_label[keep].clone()
    miss = ps == -100
    ps[miss] = torch.where(gt[miss] < 18, gt[miss] + 1, gt[miss] - 1)
    x = ps.long() + num_classes * gt.long()
    return torch.bincount(x, minlength=num_classes ** 2).reshape(num_classes, num_classes)
