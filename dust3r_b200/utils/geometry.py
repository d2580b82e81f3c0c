"""Small geometry helpers used by the host side of the two hot paths (API mirror of the subset of
dust3r/utils/geometry.py the paths touch: xy_grid :15-37, geotrf :40-101, inv :104-111,
depthmap_to_pts3d :114-162, depthmap_to_absolute_camera_coordinates :165-225)."""
from __future__ import annotations

import numpy as np
import torch


def xy_grid(W, H, device=None, origin=(0, 0), unsqueeze=None, cat_dim=-1, homogeneous=False, **arange_kw):
    """(H,W,2) grid with out[j,i] = (i + origin[0], j + origin[1]); numpy when device is None."""
    if device is None:
        xs = np.arange(origin[0], origin[0] + W, **arange_kw)
        ys = np.arange(origin[1], origin[1] + H, **arange_kw)
        grid = tuple(np.meshgrid(xs, ys, indexing='xy'))
        if homogeneous:
            grid = grid + (np.ones((H, W)),)
        if unsqueeze is not None:
            grid = tuple(np.expand_dims(g, unsqueeze) for g in grid[:2])
        return np.stack(grid, cat_dim) if cat_dim is not None else grid
    xs = torch.arange(origin[0], origin[0] + W, device=device, **arange_kw)
    ys = torch.arange(origin[1], origin[1] + H, device=device, **arange_kw)
    grid = tuple(torch.meshgrid(xs, ys, indexing='xy'))
    if homogeneous:
        grid = grid + (torch.ones((H, W), device=device),)
    if unsqueeze is not None:
        grid = (grid[0].unsqueeze(unsqueeze), grid[1].unsqueeze(unsqueeze))
    return torch.stack(grid, cat_dim) if cat_dim is not None else grid


def geotrf(Trf, pts, ncol=None, norm=False):
    """Apply a (batched) linear / affine / projective transform to points with last dim 2 or 3."""
    assert Trf.ndim >= 2
    if isinstance(Trf, np.ndarray):
        pts = np.asarray(pts)
    else:
        pts = torch.as_tensor(pts, dtype=Trf.dtype)
    out_shape = pts.shape[:-1]
    ncol = ncol or pts.shape[-1]
    d = pts.shape[-1]
    if Trf.ndim >= 3:
        nb = Trf.ndim - 2
        assert Trf.shape[:nb] == pts.shape[:nb], 'batch size does not match'
        Trf = Trf.reshape(-1, Trf.shape[-2], Trf.shape[-1])
        pts = pts.reshape(Trf.shape[0], -1, d) if pts.ndim > 2 else pts[:, None, :]
    else:
        pts = pts.reshape(-1, d)
    Tt = Trf.swapaxes(-1, -2)
    if d + 1 == Trf.shape[-1]:
        res = pts @ Tt[..., :-1, :] + Tt[..., -1:, :]
    elif d == Trf.shape[-1]:
        res = pts @ Tt
    else:
        raise ValueError(f'bad shapes {Trf.shape} x {pts.shape}')
    if norm:
        res = res / res[..., -1:]
        if norm != 1:
            res = res * norm
    return res[..., :ncol].reshape(*out_shape, ncol)


def inv(mat):
    if isinstance(mat, torch.Tensor):
        return torch.linalg.inv(mat)
    if isinstance(mat, np.ndarray):
        return np.linalg.inv(mat)
    raise ValueError(f'bad matrix type = {type(mat)}')


def depthmap_to_pts3d(depth, pseudo_focal, pp=None, **_):
    """depth (B,H,W), pseudo_focal (B,H,W) | (B,1,H,W) | (B,2,H,W) -> (B,H,W,3) camera-frame points."""
    B, H, W = depth.shape
    if pseudo_focal.ndim == 3:
        fx = fy = pseudo_focal
    elif pseudo_focal.ndim == 4:
        fx = pseudo_focal[:, 0]
        fy = pseudo_focal[:, 1] if pseudo_focal.shape[1] == 2 else fx
    else:
        raise NotImplementedError("Error, unknown input focal shape format.")
    assert fx.shape == depth.shape and fy.shape == depth.shape
    gx, gy = xy_grid(W, H, cat_dim=0, device=depth.device)[:, None]
    if pp is None:
        gx = gx - (W - 1) / 2
        gy = gy - (H - 1) / 2
    else:
        gx = gx.expand(B, -1, -1) - pp[:, 0, None, None]
        gy = gy.expand(B, -1, -1) - pp[:, 1, None, None]
    return torch.stack((depth * gx / fx, depth * gy / fy, depth), dim=-1)


def depthmap_to_camera_coordinates(depthmap, camera_intrinsics, pseudo_focal=None):
    camera_intrinsics = np.float32(camera_intrinsics)
    H, W = depthmap.shape
    assert camera_intrinsics[0, 1] == 0.0 and camera_intrinsics[1, 0] == 0.0
    if pseudo_focal is None:
        fu, fv = camera_intrinsics[0, 0], camera_intrinsics[1, 1]
    else:
        assert pseudo_focal.shape == (H, W)
        fu = fv = pseudo_focal
    cu, cv = camera_intrinsics[0, 2], camera_intrinsics[1, 2]
    u, v = np.meshgrid(np.arange(W), np.arange(H))
    z = depthmap
    X_cam = np.stack(((u - cu) * z / fu, (v - cv) * z / fv, z), axis=-1).astype(np.float32)
    return X_cam, depthmap > 0.0


def depthmap_to_absolute_camera_coordinates(depthmap, camera_intrinsics, camera_pose, **kw):
    X_cam, valid = depthmap_to_camera_coordinates(depthmap, camera_intrinsics)
    X_world = X_cam
    if camera_pose is not None:
        R, t = camera_pose[:3, :3], camera_pose[:3, 3]
        X_world = np.einsum("ik, vuk -> vui", R, X_cam) + t[None, None, :]
    return X_world, valid


def find_reciprocal_matches(P1, P2):
    """dust3r/utils/geometry.py:345-361: (reciprocal_in_P2 bool[len(P2)], nn2_in_P1 int[len(P2)], number of matches).
    P2[k] is a reciprocal match when the nearest point of P1 to it, nn2_in_P1[k], has P2[k] as ITS nearest point in P2.
    CUDA tensors (N,3) / (M,3) are matched by a brute-force kernel on the GPU (csrc/scene_ops.cu) and returned as tensors;
    arrays / CPU tensors use scipy's cKDTree exactly like the reference and return numpy."""
    if torch.is_tensor(P1) and P1.is_cuda:
        from ..cloud_opt.scene_ops import nearest_neighbours
        P2 = torch.as_tensor(P2, device=P1.device)
        nn1_in_P2 = nearest_neighbours(P1, P2)
        nn2_in_P1 = nearest_neighbours(P2, P1)
        reciprocal_in_P2 = nn1_in_P2[nn2_in_P1] == torch.arange(len(nn2_in_P1), device=P1.device)
        return reciprocal_in_P2, nn2_in_P1, int(reciprocal_in_P2.sum())
    from scipy.spatial import cKDTree

    def nearest(src, dst):
        return cKDTree(dst).query(src, workers=8)[1]
    P1, P2 = np.asarray(P1), np.asarray(P2)
    to_p2, to_p1 = nearest(P1, P2), nearest(P2, P1)
    mutual_2 = to_p2[to_p1] == np.arange(len(to_p1))
    assert (to_p1[to_p2] == np.arange(len(to_p2))).sum() == mutual_2.sum()     # the relation is symmetric
    return mutual_2, to_p1, mutual_2.sum()
