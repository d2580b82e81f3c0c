"""Focal length from a pointmap and its pixel grid (dust3r/post_process.py:12-60, used by the 'mst' initialisation).

A pinhole camera at the origin maps (x, y, z) to f * (x/z, y/z) + pp.  Given the predicted points of every pixel the
focal is the scalar that best explains the pixel coordinates, estimated either by a robust vote or by iteratively
re-weighted least squares on the reprojection error."""
from __future__ import annotations

import math

import torch

from .utils.geometry import xy_grid


def _vote_focal(centered_px, pts):
    """Every pixel votes twice: u z / x and v z / y.  The median ignores the NaN votes (0/0) of degenerate pixels."""
    z = pts[..., 2]
    votes = torch.cat((centered_px[..., 0] * z / pts[..., 0], centered_px[..., 1] * z / pts[..., 1]), dim=-1)
    return torch.nanmedian(votes, dim=-1).values


def _irls_focal(centered_px, pts, steps=10):
    """Weiszfeld iterations for  min_f  sum_p || px_p - f * (x/z, y/z)_p ||  (L1 of the residual norms): start from
    the least-squares solution, then re-solve with weights 1 / residual."""
    rays = (pts[..., :2] / pts[..., 2:3]).nan_to_num(posinf=0, neginf=0)
    num = (rays * centered_px).sum(dim=-1)      # <ray, px>
    den = rays.square().sum(dim=-1)             # <ray, ray>
    focal = num.mean(dim=1) / den.mean(dim=1)
    for _ in range(steps):
        resid = (centered_px - focal.view(-1, 1, 1) * rays).norm(dim=-1)
        weight = resid.clip(min=1e-8).reciprocal()
        focal = (weight * num).mean(dim=1) / (weight * den).mean(dim=1)
    return focal


def estimate_focal_knowing_depth(pts3d, pp, focal_mode='median', min_focal=0., max_focal=math.inf):
    """pts3d (B,H,W,3) in the camera frame, pp (B,2) principal points -> (B,) focal lengths in pixels, clipped to
    [min_focal, max_focal] x the focal of a 60-degree field of view."""
    B, H, W, C = pts3d.shape
    assert C == 3
    centered_px = xy_grid(W, H, device=pts3d.device).view(1, -1, 2) - pp.view(-1, 1, 2)
    pts = pts3d.flatten(1, 2)
    if focal_mode == 'median':
        with torch.no_grad():
            focal = _vote_focal(centered_px, pts)
    elif focal_mode == 'weiszfeld':
        if pts3d.is_cuda:
            from .cloud_opt.scene_ops import weiszfeld_focal       # one CTA per pointmap (csrc/scene_ops.cu)
            focal = weiszfeld_focal(pts3d, pp.reshape(-1, 2).expand(B, 2))
        else:
            focal = _irls_focal(centered_px, pts)
    else:
        raise ValueError(f'bad {focal_mode=}')
    fov60 = max(H, W) / (2 * math.tan(math.radians(60) / 2))
    return focal.clip(min=min_focal * fov60, max=max_focal * fov60)
