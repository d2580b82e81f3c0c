"""init='mst' host algorithm: CPU tests of the spanning-tree / Procrustes / PnP initialiser."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, has_reference
from dust3r_b200.utils.synth import synth_consistent_scene


def _edges(n):
    e = [(i, j) for i in range(n) for j in range(i)]
    return e + [(j, i) for i, j in e]


def test_mst_init_recovers_consistent_scene():
    """On an exactly consistent scene the initialiser alone must nearly zero the objective's residuals: the
    world pointmaps of every image agree with each aligned pairwise prediction."""
    from dust3r_b200.cloud_opt import global_aligner
    from dust3r_b200.cloud_opt import init_im_poses as init_fun
    from dust3r_b200.utils.geometry import geotrf
    n, H, W = 4, 24, 32
    out, cams, f = synth_consistent_scene(n, _edges(n), H, W, seed=1, noise=0.0)
    import cv2
    cv2.setRNGSeed(0)
    torch.manual_seed(0)
    net = global_aligner(copy.deepcopy(out), 'cpu', verbose=False)
    init_fun.minimum_spanning_tree  # noqa
    pts3d, msp_edges, im_focals, im_poses = init_fun.minimum_spanning_tree(
        net.imshapes, net.edges, net.pred_i, net.pred_j, net.conf_i, net.conf_j, net.im_conf, net.min_conf_thr, 'cpu',
        has_im_poses=True, verbose=False)
    assert len(msp_edges) == n - 1
    assert all(abs(fo - f) / f < 0.05 for fo in im_focals)
    # relative geometry: pairwise camera-centre distances match the ground truth up to one global scale
    c_est = im_poses[:, :3, 3]
    c_gt = cams[:, :3, 3]
    d_est = torch.cdist(c_est, c_est)
    d_gt = torch.cdist(c_gt, c_gt)
    s = (d_est.sum() / d_gt.sum())
    assert torch.allclose(d_est, s * d_gt, atol=0.05 * float(d_gt.max()) * float(s))


@pytest.mark.skipif(not has_reference(), reason='reference not mounted')
def test_mst_init_matches_reference():
    """Same inputs, same cv2 RNG seed -> same initial parameters as the reference initialiser
    (reference cloud_opt + local roma restatement)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'roma_stub'))
    sys.path.insert(0, '/root/reference')
    import cv2
    from dust3r.cloud_opt import global_aligner as ref_aligner
    import dust3r.cloud_opt.init_im_poses as ref_init
    from dust3r_b200.cloud_opt import global_aligner
    import dust3r_b200.cloud_opt.init_im_poses as init_fun
    n, H, W = 4, 24, 32
    out, cams, f = synth_consistent_scene(n, _edges(n), H, W, seed=2, noise=0.005)

    torch.manual_seed(3)
    cv2.setRNGSeed(0)
    ref = ref_aligner(copy.deepcopy(out), 'cpu', verbose=False)
    ref.verbose = False
    ref_init.init_minimum_spanning_tree(ref, niter_PnP=10)

    torch.manual_seed(3)
    cv2.setRNGSeed(0)
    net = global_aligner(copy.deepcopy(out), 'cpu', verbose=False)
    pts3d, _, im_focals, im_poses = init_fun.minimum_spanning_tree(
        net.imshapes, net.edges, net.pred_i, net.pred_j, net.conf_i, net.conf_j, net.im_conf, net.min_conf_thr, 'cpu',
        has_im_poses=True, niter_PnP=10, verbose=False)
    # init_from_pts3d ends with a loss evaluation, which needs the GPU kernel; set parameters only
    net.verbose = False
    try:
        init_fun.init_from_pts3d(net, pts3d, im_focals, im_poses)
    except Exception as e:  # pragma: no cover
        raise
    for k in ('pw_poses', 'im_poses', 'im_focals'):
        a, b = getattr(ref, k).data, getattr(net, k).data
        if k.endswith('poses'):   # quaternion sign is arbitrary
            qa, qb = a[:, :4], b[:, :4]
            sign = torch.sign((qa * qb).sum(-1, keepdim=True))
            b = torch.cat((qb * sign, b[:, 4:]), dim=-1)
        assert torch.allclose(a, b, atol=2e-3, rtol=2e-3), (k, float((a - b).abs().max()))
    assert torch.allclose(ref.im_depthmaps.data, net.im_depthmaps.data, atol=2e-3, rtol=2e-3)


@pytest.mark.skipif(not has_reference(), reason='reference not mounted')
def test_pair_viewer_matches_live_reference():
    """GlobalAlignerMode.PairViewer (closed form, cv2 PnP) against the unmodified reference class on a consistent
    two-view scene, OpenCV's RANSAC seeded identically."""
    import cv2
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'roma_stub'))
    sys.path.insert(0, '/root/reference')
    from dust3r_b200.cloud_opt import global_aligner as ours, GlobalAlignerMode as OurMode
    from dust3r.cloud_opt import global_aligner as theirs, GlobalAlignerMode as RefMode
    out, cams, f = synth_consistent_scene(2, [(0, 1), (1, 0)], 48, 64, seed=3, noise=0.002)
    cv2.setRNGSeed(0)
    a = ours(copy.deepcopy(out), 'cpu', mode=OurMode.PairViewer, verbose=False)
    cv2.setRNGSeed(0)
    b = theirs(copy.deepcopy(out), 'cpu', mode=RefMode.PairViewer, verbose=False)
    assert torch.allclose(a.get_focals(), b.get_focals(), rtol=1e-6)
    assert torch.allclose(a.get_im_poses(), b.get_im_poses(), atol=1e-5)
    assert torch.equal(a.get_principal_points(), b.get_principal_points())
    assert torch.allclose(a.get_intrinsics(), b.get_intrinsics(), rtol=1e-6)
    for x, y in zip(a.get_depthmaps(), b.get_depthmaps()):
        assert torch.allclose(x, y, atol=1e-5)
    for x, y in zip(a.get_pts3d(), b.get_pts3d()):
        assert torch.allclose(x, y, atol=1e-5)
    for x, y in zip(a.get_masks(), b.get_masks()):
        assert torch.equal(x, y)
    assert np.isnan(a())   # no objective: forward() is NaN like the reference's


@pytest.mark.skipif(not has_reference(), reason='reference not mounted')
def test_is_symmetrized_quirks_match_live_reference():
    """Exhaustive over instance lists of length <= 5 on a two-letter alphabet, including the IndexError the reference
    raises for an odd batch of mirrored couples."""
    import itertools
    sys.path.insert(0, '/root/reference')
    from dust3r.utils.misc import is_symmetrized as ref
    from dust3r_b200.utils.misc import is_symmetrized as mine

    def outcome(fn, a, b):
        try:
            return bool(fn(dict(instance=a), dict(instance=b)))
        except IndexError:
            return 'IndexError'
    for n in range(1, 6):
        for a in itertools.product('ab', repeat=n):
            for b in itertools.product('ab', repeat=n):
                assert outcome(mine, list(a), list(b)) == outcome(ref, list(a), list(b)), (a, b)


@pytest.mark.skipif(not has_reference(), reason='reference not mounted')
@pytest.mark.parametrize('fx_and_fy', [False, True])
def test_modular_optimizer_presets_match_live_reference(fx_and_fy):
    """Host-side API of ModularPointCloudOptimizer (presets with int / list / boolean-tensor / array masks, parameter
    encodings, intrinsics, world pointmaps) against the unmodified reference class, same seeds, on the CPU."""
    import warnings
    warnings.filterwarnings('ignore')
    from dust3r_b200.utils.synth import synth_pair_predictions
    sys.path.insert(0, os.path.join(ROOT, 'oracle', 'roma_stub'))
    sys.path.insert(0, '/root/reference')
    from dust3r_b200.cloud_opt import global_aligner as ours, GlobalAlignerMode as OurMode
    from dust3r.cloud_opt import global_aligner as theirs, GlobalAlignerMode as RefMode
    n, H, W = 4, 24, 32
    edges = [(i, j) for i in range(n) for j in range(n) if i != j]
    out = synth_pair_predictions(n, edges, H, W, seed=2)
    torch.manual_seed(0)
    a = ours(copy.deepcopy(out), 'cpu', mode=OurMode.ModularPointCloudOptimizer, verbose=False, fx_and_fy=fx_and_fy, optimize_pp=True)
    torch.manual_seed(0)
    b = theirs(copy.deepcopy(out), 'cpu', mode=RefMode.ModularPointCloudOptimizer, verbose=False, fx_and_fy=fx_and_fy, optimize_pp=True)
    Ks = [torch.tensor([[30. + i, 0, 15 + i], [0, 32. + i, 11 - i], [0, 0, 1.]]) for i in range(2)]
    poses = [torch.eye(4), torch.tensor([[0., -1, 0, 1], [1, 0, 0, 2], [0, 0, 1, 3], [0, 0, 0, 1]])]
    for net in (a, b):
        net.preset_intrinsics(Ks, msk=[1, 3])
        net.preset_pose(poses, pose_msk=torch.tensor([True, False, True, False]))
        net.preset_focal([55.0], msk=0)
        net.preset_principal_point([torch.tensor([14., 13.])], msk=np.array([2]))
    assert a.norm_pw_scale == b.norm_pw_scale
    for name in ('im_poses', 'im_pp', 'im_focals'):
        assert [p.requires_grad for p in getattr(a, name)] == [p.requires_grad for p in getattr(b, name)], name
    for get in ('get_focals', 'get_principal_points', 'get_intrinsics', 'get_im_poses'):
        assert torch.equal(getattr(a, get)(), getattr(b, get)()), get
    for x, y in zip(a.get_pts3d(), b.get_pts3d()):
        assert torch.equal(x, y)
    for x, y in zip(a.get_depthmaps(), b.get_depthmaps()):
        assert torch.equal(x, y)
    assert a.get_known_focal_mask().tolist() == [True, True, False, True]
