"""Pseudo-label quality metrics on the device — counterpart of
/root/reference/gapro/eval_ps_labels.py:100-172 (`--eval_pslabel` of the CLI).

The reference builds N x K one-hot matrices and multiplies them; here the K x K'
contingency table comes from one integer bincount over the points (exact counts).
"""
from __future__ import annotations

import torch


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
    keep = semantic_label != -100
    gt = semantic_label[keep].clone()
    ps = ps_semantic_label[keep].clone()
    miss = ps == -100
    ps[miss] = torch.where(gt[miss] < 18, gt[miss] + 1, gt[miss] - 1)
    x = ps.long() + num_classes * gt.long()
    return torch.bincount(x, minlength=num_classes ** 2).reshape(num_classes, num_classes)
