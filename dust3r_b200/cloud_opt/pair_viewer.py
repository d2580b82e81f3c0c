"""PairViewer — closed-form scene for one symmetrised pair, no optimisation (API mirror of
dust3r/cloud_opt/pair_viewer.py:18-127; used by BASELINE config 1).  Host code: focal by Weiszfeld
IRLS, relative pose by OpenCV PnP-RANSAC exactly as the reference does it (cv2, CPU)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .base_opt import BasePCOptimizer
from .commons import edge_str
from ..utils.geometry import inv, geotrf, depthmap_to_absolute_camera_coordinates
from ..post_process import estimate_focal_knowing_depth


class PairViewer(BasePCOptimizer):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.is_symmetrized and self.n_edges == 2
        self.has_im_poses = True
        import cv2

        focals, pps, rel_poses, confs = [], [], [], []
        for i in range(self.n_imgs):
            key, rkey = edge_str(i, 1 - i), edge_str(1 - i, i)
            conf = float(self.conf_i[key].mean() * self.conf_j[key].mean())
            if self.verbose:
                print(f'  - {conf=:.3} for edge {i}-{1-i}')
            confs.append(conf)
            H, W = self.imshapes[i]
            pp = torch.tensor((W / 2, H / 2))
            focal = float(estimate_focal_knowing_depth(self.pred_i[key][None], pp, focal_mode='weiszfeld'))
            focals.append(focal)
            pps.append(pp)
            # pose of image i in the frame of image 1-i from 2D-3D matches
            pixels = np.mgrid[:W, :H].T.astype(np.float32)
            pts3d = self.pred_j[rkey].numpy()
            assert pts3d.shape[:2] == (H, W)
            msk = self.get_masks()[i].numpy()
            K = np.float32([(focal, 0, pp[0]), (0, focal, pp[1]), (0, 0, 1)])
            try:
                ok, rvec, tvec, _ = cv2.solvePnPRansac(pts3d[msk], pixels[msk], K, None, iterationsCount=100,
                                                       reprojectionError=5, flags=cv2.SOLVEPNP_SQPNP)
                assert ok
                Rm = cv2.Rodrigues(rvec)[0]
                pose = inv(np.r_[np.c_[Rm, tvec], [(0, 0, 0, 1)]])
            except Exception:
                pose = np.eye(4)
            rel_poses.append(torch.from_numpy(pose.astype(np.float32)))

        if confs[0] > confs[1]:   # scene expressed in camera 0
            im_poses = [torch.eye(4), rel_poses[1]]
            depth = [self.pred_i['0_1'][..., 2], geotrf(inv(rel_poses[1]), self.pred_j['0_1'])[..., 2]]
        else:                     # scene expressed in camera 1
            im_poses = [rel_poses[0], torch.eye(4)]
            depth = [geotrf(inv(rel_poses[0]), self.pred_j['1_0'])[..., 2], self.pred_i['1_0'][..., 2]]

        self.im_poses = nn.Parameter(torch.stack(im_poses, dim=0), requires_grad=False)
        self.focals = nn.Parameter(torch.tensor(focals), requires_grad=False)
        self.pp = nn.Parameter(torch.stack(pps, dim=0), requires_grad=False)
        self.depth = nn.ParameterList(depth)
        for p in self.parameters():
            p.requires_grad = False

    def _set_depthmap(self, idx, depth, force=False):
        if self.verbose:
            print('_set_depthmap is ignored in PairViewer')

    def get_depthmaps(self, raw=False):
        return [d.to(self.device) for d in self.depth]

    def _set_focal(self, idx, focal, force=False):
        self.focals[idx] = focal

    def get_focals(self):
        return self.focals

    def get_known_focal_mask(self):
        return torch.tensor([not p.requires_grad for p in self.focals])

    def get_principal_points(self):
        return self.pp

    def get_intrinsics(self):
        focals, pps = self.get_focals(), self.get_principal_points()
        K = torch.zeros((len(focals), 3, 3), device=self.device)
        for i in range(len(focals)):
            K[i, 0, 0] = K[i, 1, 1] = focals[i]
            K[i, :2, 2] = pps[i]
            K[i, 2, 2] = 1
        return K

    def get_im_poses(self):
        return self.im_poses

    def depth_to_pts3d(self):
        pts3d = []
        for d, intrinsics, im_pose in zip(self.depth, self.get_intrinsics(), self.get_im_poses()):
            pts, _ = depthmap_to_absolute_camera_coordinates(d.cpu().numpy(), intrinsics.cpu().numpy(),
                                                             im_pose.cpu().numpy())
            pts3d.append(torch.from_numpy(pts).to(device=self.device))
        return pts3d

    def forward(self):
        return float('nan')
