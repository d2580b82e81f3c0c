"""PairViewer: the scene of ONE image pair in closed form, no optimisation (GlobalAlignerMode.PairViewer,
dust3r/cloud_opt/pair_viewer.py:18-127; BASELINE config 1).

For each of the two images: focal from its own pointmap (Weiszfeld IRLS), pose from the OTHER pair's prediction of its
pixels (3-D points in the other camera's frame <-> its own pixel grid, OpenCV PnP-RANSAC like the reference).  The
more confident of the two directed pairs fixes the world frame; depths are the z of the points in each camera.
Host code (cv2 on the CPU) -- outside the two hot paths."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from ..post_process import estimate_focal_knowing_depth
from ..utils.geometry import geotrf, inv
from .base_opt import BasePCOptimizer
from .commons import edge_str


def _pnp_cam2world(points, pixels, focal, pp, iterations=100):
    """Camera-to-world pose of a camera that sees world `points` (N,3) at `pixels` (N,2); identity if PnP fails."""
    import cv2
    K = np.float32([(focal, 0, pp[0]), (0, focal, pp[1]), (0, 0, 1)])
    try:
        ok, rvec, tvec, _ = cv2.solvePnPRansac(points, pixels, K, None, iterationsCount=iterations, reprojectionError=5,
                                               flags=cv2.SOLVEPNP_SQPNP)
        if not ok:
            raise RuntimeError('PnP failed')
        world2cam = np.eye(4)
        world2cam[:3, :3] = cv2.Rodrigues(rvec)[0]
        world2cam[:3, 3] = tvec.ravel()
        return torch.from_numpy(np.linalg.inv(world2cam).astype(np.float32))
    except Exception:
        return torch.eye(4)


class PairViewer(BasePCOptimizer):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        assert self.is_symmetrized and self.n_edges == 2, 'PairViewer needs exactly the pairs (0,1) and (1,0)'
        self.has_im_poses = True
        masks = self.get_masks()
        views = [self._solve_view(i, masks[i]) for i in range(2)]
        # the directed pair with the higher confidence product defines the world frame (its first image = identity)
        anchor = 0 if views[0]['conf'] > views[1]['conf'] else 1
        other = 1 - anchor
        fwd = edge_str(anchor, other)
        poses, depths = [None, None], [None, None]
        poses[anchor] = torch.eye(4)
        poses[other] = views[other]['pose']
        depths[anchor] = self.pred_i[fwd][..., 2]
        depths[other] = geotrf(inv(poses[other]), self.pred_j[fwd])[..., 2]
        self.im_poses = nn.Parameter(torch.stack(poses), requires_grad=False)
        self.focals = nn.Parameter(torch.tensor([v['focal'] for v in views]), requires_grad=False)
        self.pp = nn.Parameter(torch.stack([v['pp'] for v in views]), requires_grad=False)
        self.depth = nn.ParameterList(depths)
        for prm in self.parameters():
            prm.requires_grad = False

    def _solve_view(self, i, mask):
        """Confidence of the directed pair starting at image i, focal / principal point of image i and its pose in
        the frame of the other image."""
        own, mirrored = edge_str(i, 1 - i), edge_str(1 - i, i)
        conf = float(self.conf_i[own].mean() * self.conf_j[own].mean())
        if self.verbose:
            print(f'  - {conf=:.3} for edge {i}-{1-i}')
        H, W = self.imshapes[i]
        pp = torch.tensor((W / 2, H / 2))
        focal = float(estimate_focal_knowing_depth(self.pred_i[own][None], pp, focal_mode='weiszfeld'))
        # image i's pixels as predicted by the mirrored pair: 3-D points in the other camera's frame
        points = self.pred_j[mirrored].numpy()
        assert points.shape[:2] == (H, W)
        grid = np.stack(np.meshgrid(np.arange(W, dtype=np.float32), np.arange(H, dtype=np.float32)), axis=-1)
        keep = mask.numpy()
        pose = _pnp_cam2world(points[keep], grid[keep], focal, pp)
        return dict(conf=conf, focal=focal, pp=pp, pose=pose)

    # ---- the optimizer interface, all closed form ----
    def forward(self):
        return float('nan')

    def _set_depthmap(self, idx, depth, force=False):
        if self.verbose:
            print('_set_depthmap is ignored in PairViewer')

    def _set_focal(self, idx, focal, force=False):
        self.focals[idx] = focal

    def get_depthmaps(self, raw=False):
        return [d.to(self.device) for d in self.depth]

    def get_focals(self):
        return self.focals

    def get_known_focal_mask(self):
        return torch.tensor([not f.requires_grad for f in self.focals])

    def get_principal_points(self):
        return self.pp

    def get_im_poses(self):
        return self.im_poses

    def get_intrinsics(self):
        f, pp = self.get_focals(), self.get_principal_points()
        K = torch.zeros((len(f), 3, 3), device=self.device)
        K[:, 0, 0] = K[:, 1, 1] = f.to(self.device)
        K[:, :2, 2] = pp.to(self.device)
        K[:, 2, 2] = 1
        return K

    def depth_to_pts3d(self):
        """Back-project every depth map through its pinhole intrinsics and move it to the world frame."""
        out = []
        for depth, K, cam2world in zip(self.get_depthmaps(), self.get_intrinsics(), self.get_im_poses()):
            H, W = depth.shape
            v, u = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=depth.device),
                                  torch.arange(W, dtype=torch.float32, device=depth.device), indexing='ij')
            cam = torch.stack(((u - K[0, 2]) * depth / K[0, 0], (v - K[1, 2]) * depth / K[1, 1], depth), dim=-1)
            cam2world = cam2world.to(depth.device)
            out.append(cam @ cam2world[:3, :3].T + cam2world[:3, 3])
        return out
