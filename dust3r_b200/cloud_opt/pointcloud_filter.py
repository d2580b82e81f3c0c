"""Cross-view consistency filter applied after alignment (`clean_pointcloud`, dust3r/cloud_opt/base_opt.py:369-405;
SURVEY §8f rank 3).  A 3-D point of image i that projects into image j IN FRONT of the surface image j sees there
(closer than (1 - tol) x j's depth) contradicts j; if j is the more confident of the two at that pixel, i's
confidence is cut to `bad_conf`.  O(n^2 P) projections: CUDA kernel (csrc/scene_ops.cu) for scenes on the GPU, plain torch
for CPU tensors."""
from __future__ import annotations

import torch

from ..utils.geometry import geotrf


def _see_through(points_i, conf_i, cam_j, K_j, depth_j, conf_j, tol):
    """Boolean map over image i: pixels whose point lands inside image j, in front of j's surface, where j is more
    confident than i."""
    in_j = geotrf(cam_j, points_i)                                   # image i's points in camera j's frame
    z = in_j[..., 2]
    col, row = geotrf(K_j, in_j, norm=1, ncol=2).round().long().unbind(-1)
    H, W = conf_j.shape
    visible = (z > 0) & (col >= 0) & (col < W) & (row >= 0) & (row < H)
    hit = (row[visible], col[visible])
    contradicts = (z[visible] < (1 - tol) * depth_j[hit]) & (conf_i[visible] < conf_j[hit])
    out = visible.clone()
    out[visible] = contradicts
    return out


@torch.no_grad()
def clean_pointcloud(im_confs, K, cams, depthmaps, all_pts3d, tol=0.001, bad_conf=0, dbg=()):
    """im_confs / depthmaps / all_pts3d: per-image (H,W) / (H,W) / (H,W,3) (flattened inputs are reshaped); K (n,3,3);
    cams (n,4,4) world-to-camera.  Returns new confidence maps.  Images are visited in order and later tests see the
    confidences already lowered by earlier ones (the reference's sequential semantics)."""
    n = len(im_confs)
    assert n == len(cams) == len(K) == len(depthmaps) == len(all_pts3d)
    assert 0 <= tol < 1
    if n > 0 and all(torch.is_tensor(c) and c.is_cuda for c in im_confs):
        from .scene_ops import clean_pointcloud as clean_cuda      # one thread per pixel, n launches (csrc/scene_ops.cu)
        return clean_cuda(im_confs, K, cams, depthmaps, all_pts3d, tol=tol, bad_conf=bad_conf)
    conf = [c.clone() for c in im_confs]
    points = [p.view(*c.shape, 3) for p, c in zip(all_pts3d, im_confs)]
    depth = [d.view(*c.shape) for d, c in zip(depthmaps, im_confs)]
    for i in range(n):
        for j in range(n):
            if i == j:
                continue
            bad = _see_through(points[i], conf[i], cams[j], K[j], depth[j], conf[j], tol)
            conf[i][bad] = conf[i][bad].clip_(max=bad_conf)
    return conf
