"""Deterministic synthetic weights / images / pointmaps.

There are no checkpoints or datasets offline, so tests, goldens and bench.py all draw from here.  Values
depend only on (key, shape, seed) through torch's CPU Philox-free default generator, which is
reproducible across machines for a fixed torch version; goldens additionally store the outputs.
"""
from __future__ import annotations

import hashlib
import math
import torch

from ..config import ModelConfig, state_dict_spec


def _gen(seed: int, key: str) -> torch.Generator:
    h = int.from_bytes(hashlib.sha256(f'{seed}:{key}'.encode()).digest()[:7], 'little')
    g = torch.Generator(device='cpu')
    g.manual_seed(h)
    return g


def synth_state_dict(cfg: ModelConfig, seed: int = 0, dtype=torch.float32):
    """Xavier-like weights (as croco.py:111-127 initialises them) but with non-zero biases and
    non-unit LayerNorm gains so every epilogue path is exercised.  layer_rn.k aliases layer{k+1}_rn."""
    sd = {}
    spec = state_dict_spec(cfg)
    for key, shape in spec.items():
        if '.scratch.layer_rn.' in key:
            continue
        g = _gen(seed, key)
        if key == 'mask_token':
            t = torch.randn(shape, generator=g) * 0.02
        elif '.norm' in key or key.startswith(('enc_norm', 'dec_norm')):
            if key.endswith('weight'):
                t = 1.0 + 0.1 * torch.randn(shape, generator=g)
            else:
                t = 0.05 * torch.randn(shape, generator=g)
        elif key.endswith('bias'):
            t = 0.02 * torch.randn(shape, generator=g)
        else:
            fan_out = shape[0] * (math.prod(shape[2:]) if len(shape) > 2 else 1)
            fan_in = math.prod(shape[1:])
            if '.act_postprocess.0.1.' in key or '.act_postprocess.1.1.' in key:
                # ConvTranspose2d weight is (Cin, Cout, k, k): every output pixel sees Cin taps
                fan_in, fan_out = shape[0], shape[1]
            a = math.sqrt(6.0 / (fan_in + fan_out))
            if key.endswith('.dpt.head.4.weight'):
                # un-normalised DPT trunk reaches std~8 with xavier weights; trained heads emit O(1)
                # log-depths, so damp the last 1x1 conv to keep exp()/expm1() in a sane range
                a *= 0.06
            t = (torch.rand(shape, generator=g) * 2 - 1) * a
        sd[key] = t.to(dtype)
    for key in spec:
        if '.scratch.layer_rn.' in key:
            k = int(key.split('.scratch.layer_rn.')[1].split('.')[0])
            sd[key] = sd[key.replace(f'.scratch.layer_rn.{k}.', f'.scratch.layer{k + 1}_rn.')]
    return {k: sd[k] for k in spec}


def synth_images(n: int, H: int, W: int, seed: int = 0):
    """n images in [-1,1] (the ImgNorm range, dust3r/utils/image.py:23) in load_images' dict format
    (utils/image.py:122-123)."""
    import numpy as np
    out = []
    for i in range(n):
        g = _gen(seed, f'img{i}')
        # smooth-ish content: low-res noise upsampled + fine noise, clipped to [-1,1]
        low = torch.rand((1, 3, max(H // 16, 1), max(W // 16, 1)), generator=g) * 2 - 1
        img = torch.nn.functional.interpolate(low, size=(H, W), mode='bilinear', align_corners=False)
        img = (img + 0.25 * (torch.rand((1, 3, H, W), generator=g) * 2 - 1)).clamp(-1, 1)
        out.append(dict(img=img, true_shape=np.int32([[H, W]]), idx=i, instance=str(i)))
    return out


def synth_pair_predictions(n_imgs: int, edges, H: int, W: int, seed: int = 0):
    """Directly synthesise what inference() would return for `edges` (list of (i,j)), as SURVEY §8d
    prescribes for the alignment benchmark: pts3d ~ N(0,1)+[0,0,3], conf = 1 + 5*U(0,1)."""
    E = len(edges)
    g = _gen(seed, f'pairs{n_imgs}:{E}:{H}x{W}')
    off = torch.tensor([0.0, 0.0, 3.0])
    pts1 = torch.randn((E, H, W, 3), generator=g) + off
    pts2 = torch.randn((E, H, W, 3), generator=g) + off
    conf1 = 1 + 5 * torch.rand((E, H, W), generator=g)
    conf2 = 1 + 5 * torch.rand((E, H, W), generator=g)
    import numpy as np
    ts = torch.from_numpy(np.int32([[H, W]] * E))
    view1 = dict(idx=[int(i) for i, j in edges], instance=[str(i) for i, j in edges], true_shape=ts)
    view2 = dict(idx=[int(j) for i, j in edges], instance=[str(j) for i, j in edges], true_shape=ts)
    pred1 = dict(pts3d=pts1, conf=conf1)
    pred2 = dict(pts3d_in_other_view=pts2, conf=conf2)
    return dict(view1=view1, view2=view2, pred1=pred1, pred2=pred2, loss=None)


def synth_consistent_scene(n_imgs: int, edges, H: int, W: int, seed: int = 0, noise: float = 0.01):
    """A geometrically consistent toy scene: smooth random depth maps seen by cameras on a small arc, exact
    pairwise pointmaps (image i's points in camera i's frame / image j's points in camera i's frame) plus
    a little noise, confidences in [1.5, 6].  Gives the initialisers (MST / PnP / Procrustes) something
    meaningful to recover, unlike synth_pair_predictions' white noise."""
    import numpy as np
    g = _gen(seed, f'scene{n_imgs}:{H}x{W}')
    f = 1.2 * max(H, W)
    vs, us = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing='ij')
    cams, clouds = [], []
    for i in range(n_imgs):
        low = torch.rand((1, 1, 4, 4), generator=g)
        depth = 2.0 + torch.nn.functional.interpolate(low, size=(H, W), mode='bicubic', align_corners=True)[0, 0]
        pts_cam = torch.stack(((us - W / 2) * depth / f, (vs - H / 2) * depth / f, depth), dim=-1)
        ang = 0.25 * (i - (n_imgs - 1) / 2)
        R = torch.tensor([[math.cos(ang), 0, math.sin(ang)], [0, 1, 0], [-math.sin(ang), 0, math.cos(ang)]], dtype=torch.float32)
        t = torch.tensor([1.5 * math.sin(ang), 0.05 * i, 0.3 * (1 - math.cos(ang))], dtype=torch.float32)
        c2w = torch.eye(4)
        c2w[:3, :3], c2w[:3, 3] = R, t
        cams.append(c2w)
        clouds.append(pts_cam)

    def to_frame(pts_cam, c2w_src, c2w_dst):
        world = pts_cam @ c2w_src[:3, :3].T + c2w_src[:3, 3]
        w2c = torch.linalg.inv(c2w_dst)
        return world @ w2c[:3, :3].T + w2c[:3, 3]
    p1, p2, c1, c2 = [], [], [], []
    for (i, j) in edges:
        p1.append(clouds[i] + noise * torch.randn((H, W, 3), generator=g))
        p2.append(to_frame(clouds[j], cams[j], cams[i]) + noise * torch.randn((H, W, 3), generator=g))
        c1.append(1.5 + 4.5 * torch.rand((H, W), generator=g))
        c2.append(1.5 + 4.5 * torch.rand((H, W), generator=g))
    ts = torch.from_numpy(np.int32([[H, W]] * len(edges)))
    out = dict(view1=dict(idx=[int(i) for i, j in edges], instance=[str(i) for i, j in edges], true_shape=ts),
               view2=dict(idx=[int(j) for i, j in edges], instance=[str(j) for i, j in edges], true_shape=ts),
               pred1=dict(pts3d=torch.stack(p1), conf=torch.stack(c1)),
               pred2=dict(pts3d_in_other_view=torch.stack(p2), conf=torch.stack(c2)), loss=None)
    return out, torch.stack(cams), f


def many_ar_inputs(H: int, W: int, seed: int = 9):
    """4 pairs stored in landscape (H x W, W >= H) for the landscape_only=True (ManyAR) path; orientation of
    (view1, view2) per item: LL, LP, PL, PP -- a portrait item is the transposed storage of a (W x H) image."""
    g = torch.Generator().manual_seed(seed)
    img1 = torch.rand((4, 3, H, W), generator=g) * 2 - 1
    img2 = torch.rand((4, 3, H, W), generator=g) * 2 - 1
    L, P = [H, W], [W, H]
    ts1 = torch.tensor([L, L, P, P], dtype=torch.int32)
    ts2 = torch.tensor([L, P, L, P], dtype=torch.int32)
    return (dict(img=img1, true_shape=ts1, instance=['0', '1', '2', '3']),
            dict(img=img2, true_shape=ts2, instance=['4', '5', '6', '7']))


def synth_photo(H: int, W: int, seed: int = 0):
    """A decoded 'photograph': uint8 (H, W, 3) numpy array with natural-image statistics (smooth colour fields, a few hard
    edges, sensor noise) -- input of the load_images preprocessing tests (edges and noise exercise the negative lobes and the
    clipping of the resampling filters, which flat or purely random images do not)."""
    g = _gen(seed, f'photo{H}x{W}')
    y = torch.linspace(0, 1, H)[:, None, None]
    x = torch.linspace(0, 1, W)[None, :, None]
    f = torch.rand((6, 3), generator=g) * 9 + 1
    p = torch.rand((6, 3), generator=g) * 6.28
    field = sum(torch.sin(f[k] * (x if k % 2 else y) * 6.28 + p[k]) for k in range(6)) / 6
    img = 127 + 110 * field
    # hard-edged rectangles (saturated values next to dark ones: ringing gets clipped)
    for _ in range(4):
        r = torch.rand((4,), generator=g)
        y0, x0 = int(r[0] * H * 0.8), int(r[1] * W * 0.8)
        y1, x1 = y0 + 1 + int(r[2] * H * 0.3), x0 + 1 + int(r[3] * W * 0.3)
        img[y0:y1, x0:x1] = torch.randint(0, 2, (3,), generator=g).to(torch.float32) * 255
    img = img + torch.randn((H, W, 3), generator=g) * 12
    return img.clamp(0, 255).to(torch.uint8).numpy()
