"""Wall boxes from ScanNet-Planes quads — host-side counterpart of
/root/reference/gapro/scannet_planes.py:177-230 (tiny numpy geometry, tens of quads per
scene; not accelerated).  Optional input of the pipeline: a scene without a planes JSON
contributes no wall boxes, as in the reference (:180-181)."""
from __future__ import annotations

import json
import os

import numpy as np

PLANES_ROOT = "dataset/scannetv2/scannet_planes"
TRANSFORM_ROOT = "dataset/scannetv2/scans_transform"


def _coplanar(q, tol=100.0):
    """scalar triple product of the three edge vectors from vertex 0 within +-tol (:8-22)."""
    s1, s2, s3 = q[1] - q[0], q[2] - q[0], q[3] - q[0]
    return abs(float(np.dot(s1, np.cross(s2, s3)))) <= tol


def _plane_normal(q):
    """Unit normal of the least-squares plane through 4 vertices (:25-56): fit z = ax + by + c;
    if the normal equations are singular (vertical plane) fit ax + by + 1 = 0 instead."""
    A = np.column_stack([q[:, 0], q[:, 1], np.ones(4)])
    AtA = A.T @ A
    if np.linalg.det(AtA) > 1e-10:
        fit = (np.linalg.inv(AtA) @ A.T) @ q[:, 2]               # (A^T A)^-1 A^T first, then z: the reference's product order
        with np.errstate(divide="ignore", invalid="ignore"):     # plane through the origin: NaN, as in the reference
            n = np.array([fit[0] / fit[2], fit[1] / fit[2], -1.0 / fit[2]])
    else:
        A2 = A[:, :2]
        fit = (np.linalg.inv(A2.T @ A2) @ A2.T) @ -np.ones(4)
        n = np.array([fit[0], fit[1], 0.0])
    with np.errstate(invalid="ignore"):
        return n / np.linalg.norm(n)


def quad_to_box(q, normal):
    """Axis-aligned box spanned by a vertical quad: centre +- width/2 along the in-plane horizontal
    direction, +- height/2 vertically (:101-159)."""
    centre = q.mean(0)
    h = np.array([normal[0], normal[1], 0.0])
    h = h / np.linalg.norm(h)
    h = h / max(np.linalg.norm(h), 1e-6)
    edge = q[0] - q[1]
    cos_t = edge[2] / max(np.linalg.norm(edge), 1e-8)
    l = [np.linalg.norm(q[i] - q[(i + 1) % 4]) for i in range(4)]
    a, b = (l[0] + l[2]) / 2, (l[1] + l[3]) / 2
    height, width = (a, b) if abs(cos_t) > 0.5 else (b, a)
    x = (centre[0] + width * h[1] / 2, centre[0] - width * h[1] / 2)
    y = (centre[1] - width * h[0] / 2, centre[1] + width * h[0] / 2)
    z = (centre[2] + height / 2, centre[2] - height / 2)
    return np.array([min(x), min(y), min(z), max(x), max(y), max(z)])


def wall_boxes_from_planes(plane_dict, axis_align_matrix):
    verts = np.array(plane_dict["verts"], dtype=np.float64)
    verts = np.column_stack([verts[:, 0], -verts[:, 2], verts[:, 1]])     # y <- -z, z <- y (:192-195)
    # (the reference also computes the room centre before the alignment, :216, and never uses it)
    pts = np.ones((verts.shape[0], 4))
    pts[:, :3] = verts
    verts = (pts @ np.asarray(axis_align_matrix).T)[:, :3]
    boxes = []
    for quad in plane_dict["quads"]:
        if len(quad) != 4:
            continue
        q = verts[list(quad)]
        if not _coplanar(q):
            continue
        n = _plane_normal(q)
        if not abs(n[2]) < 0.2:       # vertical planes only (:218-220); a NaN normal is dropped too
            continue
        boxes.append(quad_to_box(q, n))
    if not boxes:
        return [], [], []
    boxes = np.array(boxes)
    cls = np.full(len(boxes), 18, dtype=np.int64)
    vol = np.prod(np.clip(boxes[:, 3:] - boxes[:, :3], 0.0, None), axis=-1)
    return cls, boxes, vol


def get_wall_boxes(scan_name, planes_root=PLANES_ROOT, transform_root=TRANSFORM_ROOT):
    path = os.path.join(planes_root, scan_name + ".json")
    if not os.path.exists(path):
        return [], [], []
    with open(path) as f:
        plane_dict = json.load(f)
    from .gen_ps import read_axis_align_matrix
    A = read_axis_align_matrix(os.path.join(transform_root, scan_name, scan_name + ".txt"))
    return wall_boxes_from_planes(plane_dict, A)
