"""PointCloudOptimizer — stacked-parameter global aligner (API mirror of
dust3r/cloud_opt/optimizer.py).  Parameters keep the reference's names, shapes and parametrisation
(log-depth padded to max_area, poses as XYZW quaternion + signed-log1p translation, focal as
focal_break*log f, principal point as offset/10) so state dicts interchange; the optimisation itself
runs in the fused CUDA step."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .base_opt import BasePCOptimizer
from ..utils.geometry import xy_grid, geotrf
from ..utils.device import to_cpu, to_numpy


def _ravel_hw(tensor, fill=0):
    tensor = tensor.reshape((tensor.shape[0] * tensor.shape[1],) + tuple(tensor.shape[2:]))
    if len(tensor) < fill:
        tensor = torch.cat((tensor, tensor.new_zeros((fill - len(tensor),) + tuple(tensor.shape[1:]))))
    return tensor


def ParameterStack(params, keys=None, is_param=None, fill=0):
    if keys is not None:
        params = [params[k] for k in keys]
    if fill > 0:
        params = [_ravel_hw(p, fill) for p in params]
    requires_grad = params[0].requires_grad
    assert all(p.requires_grad == requires_grad for p in params)
    params = torch.stack(list(params)).float().detach()
    if is_param or requires_grad:
        params = nn.Parameter(params)
        params.requires_grad_(requires_grad)
    return params


class PointCloudOptimizer(BasePCOptimizer):
    def __init__(self, *args, optimize_pp=False, focal_break=20, **kwargs):
        super().__init__(*args, **kwargs)
        self.has_im_poses = True
        self.focal_break = focal_break

        # same draws, same order as optimizer.py:29-33 so a shared torch seed gives a shared start
        depth0 = [torch.randn(H, W) / 10 - 3 for H, W in self.imshapes]
        poses0 = [self.rand_pose(self.POSE_DIM) for _ in range(self.n_imgs)]
        focals0 = [torch.FloatTensor([self.focal_break * np.log(max(H, W))]) for H, W in self.imshapes]
        pp0 = [torch.zeros((2,)) for _ in range(self.n_imgs)]

        self.imshape = self.imshapes[0]
        im_areas = [h * w for h, w in self.imshapes]
        self.max_area = max(im_areas)

        self.im_depthmaps = nn.Parameter(torch.stack([_ravel_hw(d, self.max_area) for d in depth0]).float())
        self.im_poses = nn.Parameter(torch.stack(poses0).float())
        self.im_focals = nn.Parameter(torch.stack(focals0).float())
        self.im_pp = nn.Parameter(torch.stack(pp0).float())
        self.im_pp.requires_grad_(optimize_pp)
        self.register_buffer('_pp', torch.tensor([(w / 2, h / 2) for h, w in self.imshapes]))
        self.register_buffer('_ei', torch.tensor([i for i, j in self.edges]))
        self.register_buffer('_ej', torch.tensor([j for i, j in self.edges]))
        self.total_area_i = sum(im_areas[i] for i, j in self.edges)
        self.total_area_j = sum(im_areas[j] for i, j in self.edges)

    # the reference also registers _grid/_weight_*/_stacked_pred_* buffers (optimizer.py:45-57); here
    # they exist only inside the engine's packed float4 observation buffer.

    def _engine_variant(self):
        return 'stacked'

    def _engine_pix_stride(self):
        return self.max_area

    def _engine_push(self, eng):
        eng.set_params(self.im_depthmaps.data.view(-1), self.im_poses.data, self.im_focals.data, self.im_pp.data,
                       self.pw_poses.data, self.pw_adaptors.data,
                       train_poses=self.im_poses.requires_grad, train_focals=self.im_focals.requires_grad,
                       train_pp=self.im_pp.requires_grad, train_pw=self.pw_poses.requires_grad,
                       train_adaptors=self.pw_adaptors.requires_grad, norm_pw_scale=self.norm_pw_scale)

        def pull():
            s = eng.get_small()
            self.im_poses.data.copy_(s['im_poses'])
            self.im_focals.data.copy_(s['im_focals'])
            self.im_pp.data.copy_(s['im_pp'])
            self.pw_poses.data.copy_(s['pw_poses'])
            self.pw_adaptors.data.copy_(s['pw_adaptors'])
        return pull

    # ---------------------------------------------------------------- presets
    def _check_all_imgs_are_selected(self, msk):
        assert np.all(self._get_msk_indices(msk) == np.arange(self.n_imgs)), 'incomplete mask!'

    def preset_pose(self, known_poses, pose_msk=None):
        self._check_all_imgs_are_selected(pose_msk)
        if isinstance(known_poses, torch.Tensor) and known_poses.ndim == 2:
            known_poses = [known_poses]
        for idx, pose in zip(self._get_msk_indices(pose_msk), known_poses):
            if self.verbose:
                print(f' (setting pose #{idx} = {pose[:3,3]})')
            self._no_grad(self._set_pose(self.im_poses, idx, torch.as_tensor(pose)))
        self.im_poses.requires_grad_(False)
        self.norm_pw_scale = False

    def preset_focal(self, known_focals, msk=None):
        self._check_all_imgs_are_selected(msk)
        for idx, focal in zip(self._get_msk_indices(msk), known_focals):
            if self.verbose:
                print(f' (setting focal #{idx} = {focal})')
            self._no_grad(self._set_focal(idx, focal))
        self.im_focals.requires_grad_(False)

    def preset_principal_point(self, known_pp, msk=None):
        self._check_all_imgs_are_selected(msk)
        for idx, pp in zip(self._get_msk_indices(msk), known_pp):
            if self.verbose:
                print(f' (setting principal point #{idx} = {pp})')
            self._no_grad(self._set_principal_point(idx, pp))
        self.im_pp.requires_grad_(False)

    def _get_msk_indices(self, msk):
        if msk is None:
            return range(self.n_imgs)
        if isinstance(msk, int):
            return [msk]
        if isinstance(msk, (tuple, list)):
            return self._get_msk_indices(np.array(msk))
        if msk.dtype in (bool, torch.bool, np.bool_):
            assert len(msk) == self.n_imgs
            return np.where(msk)[0]
        if np.issubdtype(msk.dtype, np.integer):
            return msk
        raise ValueError(f'bad {msk=}')

    def _no_grad(self, tensor):
        assert tensor.requires_grad, 'it must be True at this point, otherwise no modification occurs'

    # ---------------------------------------------------------------- intrinsics / poses / depth
    def _set_focal(self, idx, focal, force=False):
        param = self.im_focals[idx]
        if param.requires_grad or force:
            param.data[:] = self.focal_break * np.log(float(focal))
        return param

    def get_focals(self):
        return (self.im_focals / self.focal_break).exp()

    def get_known_focal_mask(self):
        return torch.tensor([not self.im_focals.requires_grad] * self.n_imgs)

    def _set_principal_point(self, idx, pp, force=False):
        param = self.im_pp[idx]
        H, W = self.imshapes[idx]
        if param.requires_grad or force:
            param.data[:] = to_cpu(to_numpy(pp) - (W / 2, H / 2)) / 10
        return param

    def get_principal_points(self):
        return self._pp + 10 * self.im_pp

    def get_intrinsics(self):
        K = torch.zeros((self.n_imgs, 3, 3), device=self.device)
        focals = self.get_focals().flatten()
        K[:, 0, 0] = K[:, 1, 1] = focals
        K[:, :2, 2] = self.get_principal_points()
        K[:, 2, 2] = 1
        return K

    def get_im_poses(self):
        return self._get_poses(self.im_poses)

    def _set_depthmap(self, idx, depth, force=False):
        depth = _ravel_hw(depth, self.max_area)
        param = self.im_depthmaps[idx]
        if param.requires_grad or force:
            param.data[:] = depth.log().nan_to_num(neginf=0)
        return param

    def get_depthmaps(self, raw=False):
        res = self.im_depthmaps.exp()
        if not raw:
            res = [dm[:h * w].view(h, w) for dm, (h, w) in zip(res, self.imshapes)]
        return res

    def depth_to_pts3d(self):
        """(n, max_area, 3) world-frame pointmaps.  On a B200 this is one launch of the engine's
        unprojection kernel; before `.to(cuda)` it is evaluated with torch ops (host glue, not timed)."""
        if self.device.type == 'cuda':
            eng = self._get_engine()
            pull = self._engine_push(eng)
            del pull
            return eng.pts3d().view(self.n_imgs, self.max_area, 3)
        focals = self.get_focals().unsqueeze(1)
        pp = self.get_principal_points().unsqueeze(1)
        depth = self.get_depthmaps(raw=True).unsqueeze(-1)
        grid = torch.stack([_ravel_hw(xy_grid(W, H, device=self.device).float(), self.max_area)
                            for H, W in self.imshapes])
        rel = torch.cat((depth * (grid - pp) / focals, depth), dim=-1)
        return geotrf(self.get_im_poses(), rel)
