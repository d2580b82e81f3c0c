"""ModularPointCloudOptimizer — per-image parameters so individual poses / intrinsics can be frozen
(API mirror of dust3r/cloud_opt/modular_optimizer.py).  Same fused CUDA step as PointCloudOptimizer;
the per-image `requires_grad` flags become the kernel's trainable mask, and the objective uses the
base-class normalisation (per-edge pixel means / n_edges, base_opt.py:246-273)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .base_opt import BasePCOptimizer
from ..utils.geometry import geotrf, depthmap_to_pts3d
from ..utils.device import to_cpu, to_numpy


class ModularPointCloudOptimizer(BasePCOptimizer):
    def __init__(self, *args, optimize_pp=False, fx_and_fy=False, focal_brake=20, **kwargs):
        super().__init__(*args, **kwargs)
        self.has_im_poses = True
        self.focal_brake = focal_brake
        self.im_depthmaps = nn.ParameterList(torch.randn(H, W) / 10 - 3 for H, W in self.imshapes)
        self.im_poses = nn.ParameterList(self.rand_pose(self.POSE_DIM) for _ in range(self.n_imgs))
        default_focals = [self.focal_brake * np.log(max(H, W)) for H, W in self.imshapes]
        self.im_focals = nn.ParameterList(torch.FloatTensor([f, f] if fx_and_fy else [f]) for f in default_focals)
        self.im_pp = nn.ParameterList(torch.zeros((2,)) for _ in range(self.n_imgs))
        self.im_pp.requires_grad_(optimize_pp)

    def _engine_push(self, eng):
        dev = self.device
        logd = torch.cat([d.data.reshape(-1) for d in self.im_depthmaps]).contiguous()
        eng.set_params(logd, torch.stack([p.data for p in self.im_poses]),
                       torch.stack([f.data for f in self.im_focals]), torch.stack([p.data for p in self.im_pp]),
                       self.pw_poses.data, self.pw_adaptors.data,
                       train_poses=[p.requires_grad for p in self.im_poses],
                       train_focals=[p.requires_grad for p in self.im_focals],
                       train_pp=[p.requires_grad for p in self.im_pp],
                       train_pw=self.pw_poses.requires_grad, train_adaptors=self.pw_adaptors.requires_grad,
                       norm_pw_scale=self.norm_pw_scale)

        def pull():
            s = eng.get_small()
            off = 0
            for i, (H, W) in enumerate(self.imshapes):
                self.im_depthmaps[i].data.copy_(logd[off:off + H * W].view(H, W))
                off += H * W
                self.im_poses[i].data.copy_(s['im_poses'][i])
                self.im_focals[i].data.copy_(s['im_focals'][i])
                self.im_pp[i].data.copy_(s['im_pp'][i])
            self.pw_poses.data.copy_(s['pw_poses'])
            self.pw_adaptors.data.copy_(s['pw_adaptors'])
        return pull

    # ---------------------------------------------------------------- presets
    def preset_pose(self, known_poses, pose_msk=None):
        if isinstance(known_poses, torch.Tensor) and known_poses.ndim == 2:
            known_poses = [known_poses]
        for idx, pose in zip(self._get_msk_indices(pose_msk), known_poses):
            if self.verbose:
                print(f' (setting pose #{idx} = {pose[:3,3]})')
            self._no_grad(self._set_pose(self.im_poses, idx, torch.as_tensor(pose), force=True))
        n_known_poses = sum((p.requires_grad is False) for p in self.im_poses)
        self.norm_pw_scale = (n_known_poses <= 1)

    def preset_intrinsics(self, known_intrinsics, msk=None):
        if isinstance(known_intrinsics, torch.Tensor) and known_intrinsics.ndim == 2:
            known_intrinsics = [known_intrinsics]
        for K in known_intrinsics:
            assert K.shape == (3, 3)
        self.preset_focal([K.diagonal()[:2].mean() for K in known_intrinsics], msk)
        self.preset_principal_point([K[:2, 2] for K in known_intrinsics], msk)

    def preset_focal(self, known_focals, msk=None):
        for idx, focal in zip(self._get_msk_indices(msk), known_focals):
            if self.verbose:
                print(f' (setting focal #{idx} = {focal})')
            self._no_grad(self._set_focal(idx, focal, force=True))

    def preset_principal_point(self, known_pp, msk=None):
        for idx, pp in zip(self._get_msk_indices(msk), known_pp):
            if self.verbose:
                print(f' (setting principal point #{idx} = {pp})')
            self._no_grad(self._set_principal_point(idx, pp, force=True))

    def _no_grad(self, tensor):
        return tensor.requires_grad_(False)

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

    # ---------------------------------------------------------------- accessors
    def _set_focal(self, idx, focal, force=False):
        param = self.im_focals[idx]
        if param.requires_grad or force:
            param.data[:] = self.focal_brake * np.log(float(focal))
        return param

    def get_focals(self):
        return (torch.stack(list(self.im_focals), dim=0) / self.focal_brake).exp()

    def get_known_focal_mask(self):
        return torch.tensor([not p.requires_grad for p in self.im_focals])

    def _set_principal_point(self, idx, pp, force=False):
        param = self.im_pp[idx]
        H, W = self.imshapes[idx]
        if param.requires_grad or force:
            param.data[:] = to_cpu(to_numpy(pp) - (W / 2, H / 2)) / 10
        return param

    def get_principal_points(self):
        return torch.stack([pp.new_tensor((W / 2, H / 2)) + 10 * pp for pp, (H, W) in zip(self.im_pp, self.imshapes)])

    def get_intrinsics(self):
        K = torch.zeros((self.n_imgs, 3, 3), device=self.device)
        focals = self.get_focals().view(self.n_imgs, -1)
        K[:, 0, 0] = focals[:, 0]
        K[:, 1, 1] = focals[:, -1]
        K[:, :2, 2] = self.get_principal_points()
        K[:, 2, 2] = 1
        return K

    def get_im_poses(self):
        return self._get_poses(torch.stack(list(self.im_poses)))

    def _set_depthmap(self, idx, depth, force=False):
        param = self.im_depthmaps[idx]
        if param.requires_grad or force:
            param.data[:] = depth.log().nan_to_num(neginf=0)
        return param

    def get_depthmaps(self):
        return [d.exp() for d in self.im_depthmaps]

    def depth_to_pts3d(self):
        if self.device.type == 'cuda':
            eng = self._get_engine()
            pull = self._engine_push(eng)
            del pull
            flat = eng.pts3d()
            out, off = [], 0
            for H, W in self.imshapes:
                out.append(flat[off:off + H * W].view(H, W, 3))
                off += H * W
            return out
        focals = self.get_focals()
        pp = self.get_principal_points()
        im_poses = self.get_im_poses()
        depth = self.get_depthmaps()
        def focal_ex(i): return focals[i][..., None, None].expand(1, *focals[i].shape, *self.imshapes[i])
        rel = [depthmap_to_pts3d(depth[i][None], focal_ex(i), pp=pp[i:i + 1])[0] for i in range(im_poses.shape[0])]
        return [geotrf(pose, ptmap) for pose, ptmap in zip(im_poses, rel)]

    def get_pts3d(self):
        return self.depth_to_pts3d()
