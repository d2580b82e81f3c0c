"""ORACLE SUPPORT (test infrastructure): local restatement of the three `roma` entry points the
reference's cloud_opt touches.  `roma` (requirements.txt:3 of the reference, unpinned, not vendored)
is absent from this image and cannot be installed offline, so this stub lets the UNMODIFIED reference
cloud_opt import and run when tests/golden/make_golden.py generates fixtures.  Everything produced
through it is labelled "reference cloud_opt + local roma restatement".

Restated from roma's public documentation (quaternions are XYZW, scalar last):
  RigidUnitQuat(q, t).normalize().to_homogeneous()  -> base_opt.py:154
  rotmat_to_unitquat(R)                             -> base_opt.py:169
  rigid_points_registration(x, y, weights, compute_scaling) -> init_im_poses.py:221,315
This is not roma's code; roma itself cannot be run here.  The three entry points are pinned against independent
implementations instead (tests/test_roma_stub.py): scipy.spatial.transform.Rotation for both quaternion maps (XYZW),
scipy's weighted Kabsch and a float64 closed-form weighted Umeyama for the registration, plus an optimality check.
"""
import torch


def _quat_to_rotmat(q):
    x, y, z, w = q.unbind(-1)
    xx, yy, zz = x * x, y * y, z * z
    xy, xz, yz = x * y, x * z, y * z
    wx, wy, wz = w * x, w * y, w * z
    R = torch.stack((1 - 2 * (yy + zz), 2 * (xy - wz), 2 * (xz + wy),
                     2 * (xy + wz), 1 - 2 * (xx + zz), 2 * (yz - wx),
                     2 * (xz - wy), 2 * (yz + wx), 1 - 2 * (xx + yy)), dim=-1)
    return R.reshape(q.shape[:-1] + (3, 3))


class RigidUnitQuat:
    def __init__(self, linear, translation):
        self.linear = linear
        self.translation = translation

    def normalize(self):
        return RigidUnitQuat(self.linear / torch.norm(self.linear, dim=-1, keepdim=True), self.translation)

    def to_homogeneous(self):
        R = _quat_to_rotmat(self.linear)
        batch = R.shape[:-2]
        H = torch.zeros(batch + (4, 4), dtype=R.dtype, device=R.device)
        H[..., :3, :3] = R
        H[..., :3, 3] = self.translation
        H[..., 3, 3] = 1
        return H


def rotmat_to_unitquat(R):
    """Batched rotation matrix -> unit quaternion XYZW (largest-component branch for stability)."""
    R = torch.as_tensor(R)
    batch = R.shape[:-2]
    m = R.reshape(-1, 3, 3)
    m00, m11, m22 = m[:, 0, 0], m[:, 1, 1], m[:, 2, 2]
    cand = torch.stack((1 + m00 - m11 - m22, 1 - m00 + m11 - m22, 1 - m00 - m11 + m22, 1 + m00 + m11 + m22), dim=-1)
    choice = cand.argmax(dim=-1)
    q = torch.empty((m.shape[0], 4), dtype=m.dtype, device=m.device)
    for b in range(m.shape[0]):
        c = int(choice[b])
        M = m[b]
        if c == 0:
            v = torch.stack((cand[b, 0], M[1, 0] + M[0, 1], M[2, 0] + M[0, 2], M[2, 1] - M[1, 2]))
        elif c == 1:
            v = torch.stack((M[1, 0] + M[0, 1], cand[b, 1], M[2, 1] + M[1, 2], M[0, 2] - M[2, 0]))
        elif c == 2:
            v = torch.stack((M[2, 0] + M[0, 2], M[2, 1] + M[1, 2], cand[b, 2], M[1, 0] - M[0, 1]))
        else:
            v = torch.stack((M[2, 1] - M[1, 2], M[0, 2] - M[2, 0], M[1, 0] - M[0, 1], cand[b, 3]))
        q[b] = v / v.norm()
    return q.reshape(batch + (4,))


def rigid_points_registration(x, y, weights=None, compute_scaling=False):
    """Weighted Kabsch/Umeyama: returns (R, t[, s]) minimising sum w |s R x + t - y|^2."""
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
