"""Pins oracle/roma_stub (the local restatement of the three `roma` entry points the reference's cloud_opt calls,
base_opt.py:154,169 and init_im_poses.py:221,315) against independent implementations: scipy's Rotation for the
XYZW quaternion <-> matrix maps, scipy's weighted Kabsch (Rotation.align_vectors) and a float64 closed-form weighted
Umeyama (Umeyama 1991, eqs. 34-42) for the similarity registration.  `roma` itself cannot be installed offline."""
import os
import sys

import numpy as np
import pytest
import torch
from scipy.spatial.transform import Rotation

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'roma_stub'))
import roma  # noqa: E402  (the stub)


def test_unitquat_to_matrix_matches_scipy_xyzw():
    rng = np.random.default_rng(0)
    q = rng.standard_normal((256, 4))
    t = rng.standard_normal((256, 3))
    H = roma.RigidUnitQuat(torch.from_numpy(q), torch.from_numpy(t)).normalize().to_homogeneous().numpy()
    R_ref = Rotation.from_quat(q).as_matrix()          # scipy: scalar-last (x, y, z, w), normalises its input
    assert np.abs(H[:, :3, :3] - R_ref).max() < 1e-14
    assert np.abs(H[:, :3, 3] - t).max() == 0 and np.all(H[:, 3] == np.array([0, 0, 0, 1.0]))
    # fp32, as the reference uses it
    H32 = roma.RigidUnitQuat(torch.from_numpy(q).float(), torch.from_numpy(t).float()).normalize().to_homogeneous().numpy()
    assert np.abs(H32[:, :3, :3] - R_ref).max() < 2e-6


def test_rotmat_to_unitquat_matches_scipy_up_to_sign():
    rng = np.random.default_rng(1)
    R = Rotation.random(512, random_state=2).as_matrix()
    # include rotations by ~pi about each axis (trace close to -1: every branch of the conversion is exercised)
    for ax in np.eye(3):
        R = np.concatenate([R, Rotation.from_rotvec(ax[None] * (np.pi - 1e-4 * rng.random((8, 1)))).as_matrix()])
    q = roma.rotmat_to_unitquat(torch.from_numpy(R)).numpy()
    q_ref = Rotation.from_matrix(R).as_quat()
    assert np.abs(np.linalg.norm(q, axis=-1) - 1).max() < 1e-12
    dot = np.abs((q * q_ref).sum(-1))
    assert dot.min() > 1 - 1e-12
    # and it round-trips through the stub's own quaternion -> matrix map
    back = roma.RigidUnitQuat(torch.from_numpy(q), torch.zeros(len(q), 3, dtype=torch.float64)).to_homogeneous().numpy()
    assert np.abs(back[:, :3, :3] - R).max() < 1e-10


def _umeyama_f64(x, y, w):
    """argmin_{s,R,t} sum_k w_k |s R x_k + t - y_k|^2  (Umeyama 1991 with weights), float64 numpy."""
    x, y, w = np.float64(x), np.float64(y), np.float64(w)
    W = w.sum()
    mx, my = (w[:, None] * x).sum(0) / W, (w[:, None] * y).sum(0) / W
    xc, yc = x - mx, y - my
    var_x = (w * (xc ** 2).sum(-1)).sum() / W
    cov = (w[:, None, None] * yc[:, :, None] * xc[:, None, :]).sum(0) / W
    U, D, Vt = np.linalg.svd(cov)
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    s = np.trace(np.diag(D) @ S) / var_x
    t = my - s * R @ mx
    return R, t, s


def _objective(R, t, s, x, y, w):
    return float((w * ((s * x @ R.T + t - y) ** 2).sum(-1)).sum())


@pytest.mark.parametrize('seed', [0, 1, 2, 3])
def test_weighted_registration_with_scale_matches_closed_form_and_scipy(seed):
    rng = np.random.default_rng(seed)
    n = 4000
    x = rng.standard_normal((n, 3)) * np.array([1.0, 2.0, 0.5]) + np.array([0.3, -1.0, 3.0])
    R0 = Rotation.random(random_state=seed).as_matrix()
    s0, t0 = float(np.exp(rng.normal())), rng.standard_normal(3)
    y = s0 * x @ R0.T + t0 + 0.05 * rng.standard_normal((n, 3))
    w = rng.random(n) * 5 + 0.01
    R, t, s = roma.rigid_points_registration(torch.from_numpy(x), torch.from_numpy(y), weights=torch.from_numpy(w),
                                             compute_scaling=True)
    R, t, s = R.numpy(), t.numpy(), float(s)
    Rr, tr, sr = _umeyama_f64(x, y, w)
    assert np.abs(R - Rr).max() < 1e-10 and np.abs(t - tr).max() < 1e-9 and abs(s - sr) < 1e-10
    # the rotation alone against scipy's weighted Kabsch on the centred clouds
    W = w.sum()
    xc, yc = x - (w[:, None] * x).sum(0) / W, y - (w[:, None] * y).sum(0) / W
    Rk, _ = Rotation.align_vectors(yc, xc, weights=w)
    assert np.abs(R - Rk.as_matrix()).max() < 1e-8
    # optimality: no small perturbation of (s, R, t) lowers the weighted objective
    f0 = _objective(R, t, s, x, y, w)
    for k in range(20):
        dR = Rotation.from_rotvec(1e-3 * rng.standard_normal(3)).as_matrix()
        assert _objective(dR @ R, t + 1e-3 * rng.standard_normal(3), s * (1 + 1e-3 * rng.normal()), x, y, w) >= f0 - 1e-9 * f0
    # fp32 inputs (what init_im_poses passes) stay close to the float64 solution
    R32, t32, s32 = roma.rigid_points_registration(torch.from_numpy(x).float(), torch.from_numpy(y).float(),
                                                   weights=torch.from_numpy(w).float(), compute_scaling=True)
    assert np.abs(R32.numpy() - Rr).max() < 1e-4 and abs(float(s32) - sr) < 1e-4 * sr


def test_registration_without_scale_and_reflection_case():
    rng = np.random.default_rng(7)
    x = rng.standard_normal((500, 3))
    R0 = Rotation.random(random_state=3).as_matrix()
    y = x @ R0.T + np.array([1.0, 2.0, 3.0])
    R, t = roma.rigid_points_registration(torch.from_numpy(x), torch.from_numpy(y))
    assert np.abs(R.numpy() - R0).max() < 1e-10 and np.abs(t.numpy() - [1, 2, 3]).max() < 1e-10
    # mirrored target: the best ROTATION is returned (det = +1), never a reflection
    y_m = y * np.array([1.0, 1.0, -1.0])
    Rm, tm, sm = roma.rigid_points_registration(torch.from_numpy(x), torch.from_numpy(y_m), compute_scaling=True)
    assert abs(float(torch.linalg.det(Rm)) - 1) < 1e-10
    Rr, tr, sr = _umeyama_f64(x, y_m, np.ones(len(x)))
    assert np.abs(Rm.numpy() - Rr).max() < 1e-9 and abs(float(sm) - sr) < 1e-9
