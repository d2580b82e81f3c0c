"""Pose parameterisation shared by the optimizer classes.

A pose row is [qx, qy, qz, qw, sx, sy, sz (, log-scale)]: an XYZW quaternion (normalised when read) and a translation
stored in signed-log space, t = sign(s) * expm1(|s|) -- the encoding the fused alignment kernel differentiates through
(dust3r/cloud_opt/base_opt.py:150-195 defines it).  Pairwise poses carry an extra log-scale; with `norm_pw_scale` the
pairwise scales are divided by their geometric mean and multiplied by `base_scale`, which removes the global-scale
gauge freedom of the objective."""
from __future__ import annotations

import math

import torch

from .commons import rotmat_to_unitquat, signed_expm1, signed_log1p, unitquat_to_rotmat


def rows_to_matrices(rows):
    """(N, >=7) pose rows -> (N, 4, 4) rigid transforms (rotation from the quaternion, translation decoded)."""
    mats = torch.zeros((rows.shape[0], 4, 4), dtype=rows.dtype, device=rows.device)
    mats[:, :3, :3] = unitquat_to_rotmat(rows[:, :4])
    mats[:, :3, 3] = signed_expm1(rows[:, 4:7])
    mats[:, 3, 3] = 1
    return mats


def write_row(row, rotation=None, translation=None, scale=None):
    """Encode rotation (3,3) / translation (3,) / scale into one pose row (a tensor view, modified in place).
    With a scale, the translation is stored for the UNSCALED transform (t / scale) and log(scale) goes last."""
    if rotation is not None:
        row[0:4] = rotmat_to_unitquat(rotation).to(row.device)
    if translation is not None:
        t = torch.as_tensor(translation / (scale or 1), dtype=torch.float32)
        row[4:7] = signed_log1p(t).to(row.device)
    if scale is not None:
        row[-1] = math.log(float(scale))


def split_rigid(mat):
    """(4,4) -> rotation (3,3), translation (3,)."""
    return mat[:3, :3], mat[:3, 3]


def scale_gauge(log_scales, base_scale, normalise):
    """Factor applied to every pairwise scale: base_scale / geometric-mean(scales) when normalising, else 1."""
    if not normalise:
        return 1
    return (math.log(base_scale) - log_scales.mean()).exp()
