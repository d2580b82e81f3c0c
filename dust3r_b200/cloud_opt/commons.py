"""Small host-side pieces of the global aligner: edge keys and scores, the confidence transforms, the scalar encodings
of the pose parameters, the learning-rate schedules (public names as in dust3r/cloud_opt/commons.py), and the three
rotation / registration routines the reference imports from `roma`."""
from __future__ import annotations

import math

import torch
import torch.nn as nn


# ---- pair graph -------------------------------------------------------------------------------------------------
def edge_str(i, j):
    """Key of the directed pair (i, j) in the pred_* / conf_* dictionaries."""
    return '%d_%d' % (i, j) if isinstance(i, int) and isinstance(j, int) else f'{i}_{j}'


def i_j_ij(ij):
    return edge_str(ij[0], ij[1]), ij


def edge_conf(conf_i, conf_j, edge):
    """Score of a pair = product of the mean confidences of its two pointmaps."""
    return float(conf_i[edge].mean() * conf_j[edge].mean())     # fp32 product: the ordering of near-ties depends on it


def compute_edge_scores(edges, conf_i, conf_j):
    """edges yields (key, (i, j)) couples (see i_j_ij) -> {(i, j): score}."""
    return {ij: edge_conf(conf_i, conf_j, key) for key, ij in edges}


def NoGradParamDict(x):
    """Tensors registered on the module (they follow .to(device) and land in the state dict) but never trained."""
    assert isinstance(x, dict)
    holder = nn.ParameterDict(x)
    holder.requires_grad_(False)
    return holder


def get_imshapes(edges, pred_i, pred_j):
    """(H, W) of every image, read off the pointmaps of the pairs it takes part in; all of them must agree."""
    shapes = {}
    for e, (i, j) in enumerate(edges):
        for img, pts in ((i, pred_i[e]), (j, pred_j[e])):
            hw = (int(pts.shape[0]), int(pts.shape[1]))
            seen = shapes.setdefault(img, hw)          # not inside the assert: `python -O` strips those
            assert seen == hw, f'incorrect shape for image {img}'
    return [shapes.get(img) for img in range(max(shapes) + 1)]


# ---- confidence transform (weight of a residual = trf(confidence)) ---------------------------------------------------
_CONF_TRF = {
    'log': torch.log,
    'sqrt': torch.sqrt,
    'm1': lambda c: c - 1,
    'id': lambda c: c,
    'none': lambda c: c,
}


def get_conf_trf(mode):
    try:
        return _CONF_TRF[mode]
    except KeyError:
        raise ValueError(f'bad mode for {mode=}') from None


ALL_DISTS = dict(l1='l1', l2='l2')  # the distances themselves live in the CUDA kernel


# ---- scalar encodings ---------------------------------------------------------------------------------------------
def signed_log1p(x):
    """sign(x) * log(1 + |x|): translations are optimised in this compressed space."""
    return torch.copysign(torch.log1p(x.abs()), x) * (x != 0)


def signed_expm1(x):
    """Inverse of signed_log1p."""
    return torch.copysign(torch.expm1(x.abs()), x) * (x != 0)


# ---- learning-rate schedules over t = iteration / niter in [0, 1] -----------------------------------------------------
def cosine_schedule(t, lr_start, lr_end):
    assert 0 <= t <= 1
    swing = (1 + math.cos(t * math.pi)) / 2          # 1 at t = 0, 0 at t = 1
    return lr_end + (lr_start - lr_end) * swing


def linear_schedule(t, lr_start, lr_end):
    assert 0 <= t <= 1
    return lr_start + t * (lr_end - lr_start)


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
    if x.is_cuda and x.ndim in (2, 3) and x.shape[-1] == 3:
        from .scene_ops import rigid_registration                 # moments in one kernel launch (csrc/scene_ops.cu)
        return rigid_registration(x, y, weights, compute_scaling=compute_scaling)
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
