"""Host-side helpers of the global aligner (API mirror of dust3r/cloud_opt/commons.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn


def edge_str(i, j):
    return f'{i}_{j}'


def i_j_ij(ij):
    return edge_str(*ij), ij


def edge_conf(conf_i, conf_j, edge):
    return float(conf_i[edge].mean() * conf_j[edge].mean())


def compute_edge_scores(edges, conf_i, conf_j):
    return {(i, j): edge_conf(conf_i, conf_j, e) for e, (i, j) in edges}


def NoGradParamDict(x):
    assert isinstance(x, dict)
    return nn.ParameterDict(x).requires_grad_(False)


def get_imshapes(edges, pred_i, pred_j):
    n_imgs = max(max(e) for e in edges) + 1
    imshapes = [None] * n_imgs
    for e, (i, j) in enumerate(edges):
        shape_i = tuple(pred_i[e].shape[0:2])
        shape_j = tuple(pred_j[e].shape[0:2])
        if imshapes[i]:
            assert imshapes[i] == shape_i, f'incorrect shape for image {i}'
        if imshapes[j]:
            assert imshapes[j] == shape_j, f'incorrect shape for image {j}'
        imshapes[i] = shape_i
        imshapes[j] = shape_j
    return imshapes


_CONF_TRF = {
    'log': lambda x: x.log(),
    'sqrt': lambda x: x.sqrt(),
    'm1': lambda x: x - 1,
    'id': lambda x: x,
    'none': lambda x: x,
}


def get_conf_trf(mode):
    if mode not in _CONF_TRF:
        raise ValueError(f'bad mode for {mode=}')
    return _CONF_TRF[mode]


ALL_DISTS = dict(l1='l1', l2='l2')  # the distances themselves live in the CUDA kernel


def signed_log1p(x):
    sign = torch.sign(x)
    return sign * torch.log1p(torch.abs(x))


def signed_expm1(x):
    sign = torch.sign(x)
    return sign * torch.expm1(torch.abs(x))


def cosine_schedule(t, lr_start, lr_end):
    assert 0 <= t <= 1
    return lr_end + (lr_start - lr_end) * (1 + np.cos(t * np.pi)) / 2


def linear_schedule(t, lr_start, lr_end):
    assert 0 <= t <= 1
    return lr_start + (lr_end - lr_start) * t


# --- rotation helpers the reference takes from `roma` (requirements.txt:3; not a dependency here) ---

def unitquat_to_rotmat(q):
    """XYZW quaternion (normalised here) -> (…,3,3)."""
    q = q / q.norm(dim=-1, keepdim=True)
    x, y, z, w = q.unbind(-1)
    R = torch.stack((1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y),
                     2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x),
                     2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)), dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


def rotmat_to_unitquat(R):
    """(…,3,3) rotation -> XYZW unit quaternion (w >= 0 branch-free Shepperd selection)."""
    R = torch.as_tensor(R, dtype=torch.float32)
    batch = R.shape[:-2]
    m = R.reshape(-1, 3, 3)
    m00, m11, m22 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    # four candidate (un-normalised) quaternions, pick the best conditioned one per matrix
    c0 = torch.stack((1 + m00 - m11 - m22, m[:, 1, 0] + m[:, 0, 1], m[:, 2, 0] + m[:, 0, 2], m[:, 2, 1] - m[:, 1, 2]), -1)
    c1 = torch.stack((m[:, 1, 0] + m[:, 0, 1], 1 - m00 + m11 - m22, m[:, 2, 1] + m[:, 1, 2], m[:, 0, 2] - m[:, 2, 0]), -1)
    c2 = torch.stack((m[:, 2, 0] + m[:, 0, 2], m[:, 2, 1] + m[:, 1, 2], 1 - m00 - m11 + m22, m[:, 1, 0] - m[:, 0, 1]), -1)
    c3 = torch.stack((m[:, 2, 1] - m[:, 1, 2], m[:, 0, 2] - m[:, 2, 0], m[:, 1, 0] - m[:, 0, 1], 1 + m00 + m11 + m22), -1)
    cands = torch.stack((c0, c1, c2, c3), dim=1)                      # (B,4,4)
    diag = torch.stack((c0[:, 0], c1[:, 1], c2[:, 2], c3[:, 3]), -1)  # (B,4)
    pick = diag.argmax(dim=-1)
    q = cands[torch.arange(m.shape[0]), pick]
    q = q / q.norm(dim=-1, keepdim=True)
    return q.reshape(batch + (4,))


def rigid_points_registration(x, y, weights=None, compute_scaling=False):
    """Weighted Kabsch / Umeyama: argmin sum_k w_k |s R x_k + t - y_k|^2  -> (R, t[, s])."""
    if weights is None:
        weights = torch.ones(x.shape[:-1], dtype=x.dtype, device=x.device)
    w = weights[..., None]
    wsum = w.sum(dim=-2, keepdim=True)
    xm = (w * x).sum(dim=-2, keepdim=True) / wsum
    ym = (w * y).sum(dim=-2, keepdim=True) / wsum
    xc, yc = x - xm, y - ym
    M = (w * yc).transpose(-1, -2) @ xc
    U, S, Vh = torch.linalg.svd(M)
    d = torch.sign(torch.linalg.det(U @ Vh))
    D = torch.ones_like(S)
    D[..., -1] = d
    R = U @ torch.diag_embed(D) @ Vh
    if compute_scaling:
        s = (S * D).sum(dim=-1) / (w * xc * xc).sum(dim=(-1, -2))
        t = ym.squeeze(-2) - s[..., None] * (R @ xm.transpose(-1, -2)).squeeze(-1)
        return R, t, s
    t = ym.squeeze(-2) - (R @ xm.transpose(-1, -2)).squeeze(-1)
    return R, t
