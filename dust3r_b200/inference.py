"""inference(): batched pairwise forward (API mirror of dust3r/inference.py:14-78).

Same contract as the reference: takes the `make_pairs` list, returns {'view1','view2','pred1','pred2','loss'}
with every tensor ON CPU, concatenated over pairs in input order (lists when image sizes are mixed).
Differences are internal: each batch is one C-ABI call (d3r_forward_pairs), device->host copies of
the predictions go through pinned staging buffers and overlap the next batch's compute, and
`keep_on_device=True` (extension) skips the host round trip for callers that feed global_aligner next
(SURVEY §8f rank 2)."""
from __future__ import annotations

import torch
import tqdm

from .utils.device import to_cpu, collate_with_cat


def _interleave_imgs(img1, img2):
    res = {}
    for key, value1 in img1.items():
        value2 = img2[key]
        if isinstance(value1, torch.Tensor):
            value = torch.stack((value1, value2), dim=1).flatten(0, 1)
        else:
            value = [x for pair in zip(value1, value2) for x in pair]
        res[key] = value
    return res


def make_batch_symmetric(batch):
    view1, view2 = batch
    return _interleave_imgs(view1, view2), _interleave_imgs(view2, view1)


_IGNORE = {'depthmap', 'dataset', 'label', 'instance', 'idx', 'true_shape', 'rng'}


def loss_of_one_batch(batch, model, criterion, device, symmetrize_batch=False, use_amp=False, ret=None):
    """inference.py:32-52.  `criterion` must be None (training losses are outside the hot paths)."""
    view1, view2 = batch
    for view in batch:
        for name in view.keys():
            if name in _IGNORE:
                continue
            view[name] = view[name].to(device, non_blocking=True)
    if symmetrize_batch:
        view1, view2 = make_batch_symmetric(batch)
    if criterion is not None:
        raise NotImplementedError('training criteria are not part of the inference hot path')
    pred1, pred2 = model(view1, view2)
    result = dict(view1=view1, view2=view2, pred1=pred1, pred2=pred2, loss=None)
    return result[ret] if ret else result


def check_if_same_size(pairs):
    shapes1 = [img1['img'].shape[-2:] for img1, img2 in pairs]
    shapes2 = [img2['img'].shape[-2:] for img1, img2 in pairs]
    return all(shapes1[0] == s for s in shapes1) and all(shapes2[0] == s for s in shapes2)


def _collate_views_pinned(views, pin):
    """collate_with_cat for a list of view dicts; image tensors are concatenated straight into pinned
    host memory so the H2D copy can be asynchronous."""
    out = {}
    for k in views[0]:
        vals = [v[k] for v in views]
        if k == 'img' and pin:
            n = sum(int(t.shape[0]) for t in vals)
            buf = torch.empty((n,) + tuple(vals[0].shape[1:]), dtype=vals[0].dtype, pin_memory=True)
            torch.cat(vals, out=buf)
            out[k] = buf
        else:
            out[k] = collate_with_cat(vals)
    return out


@torch.no_grad()
def inference(pairs, model, device, batch_size=8, verbose=True, keep_on_device=False):
    """inference.py:55-72.  Returns {'view1','view2','pred1','pred2','loss'}; tensors on CPU (pinned) unless
    keep_on_device.  Per batch: pinned H2D of the images -> one fused forward -> predictions copied D2H on a
    side stream into the final (whole pair list) pinned output, overlapping the next batch's compute.  The
    returned views are the caller's own CPU tensors (the reference round-trips them through the GPU)."""
    if verbose:
        print(f'>> Inference with model on {len(pairs)} image pairs')
    multiple_shapes = not check_if_same_size(pairs)
    dev = torch.device(device)
    fused = dev.type == 'cuda' and not multiple_shapes and len(pairs) > 0
    if not fused:
        # mixed image sizes (batch size forced to 1, lists instead of stacked tensors) or non-CUDA stand-in
        # models used by host-side tests: plain reference control flow
        result = []
        bs = 1 if multiple_shapes else batch_size
        for i in tqdm.trange(0, len(pairs), bs, disable=not verbose):
            res = loss_of_one_batch(collate_with_cat(pairs[i:i + bs]), model, None, device)
            result.append(res if keep_on_device else to_cpu(res))
        return collate_with_cat(result, lists=multiple_shapes)

    n = len(pairs)
    view1_all = collate_with_cat([a for a, b in pairs])
    view2_all = collate_with_cat([b for a, b in pairs])
    outs = None
    main = torch.cuda.current_stream(dev)
    side = torch.cuda.Stream(device=dev)
    for i in tqdm.trange(0, n, batch_size, disable=not verbose):
        chunk = pairs[i:i + batch_size]
        v1 = _collate_views_pinned([a for a, b in chunk], pin=True)
        v2 = _collate_views_pinned([b for a, b in chunk], pin=True)
        d1 = dict(v1, img=v1['img'].to(dev, non_blocking=True))
        d2 = dict(v2, img=v2['img'].to(dev, non_blocking=True))
        pred1, pred2 = model(d1, d2)
        flat = {('pred1', k): v for k, v in pred1.items()}
        flat.update({('pred2', k): v for k, v in pred2.items()})
        if outs is None:
            if keep_on_device:
                outs = {key: torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for key, t in flat.items()}
            else:
                outs = {key: torch.empty((n,) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True) for key, t in flat.items()}
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            for key, t in flat.items():
                outs[key][i:i + len(chunk)].copy_(t, non_blocking=True)
                t.record_stream(side)
    side.synchronize()
    res = dict(view1=view1_all, view2=view2_all, pred1={}, pred2={}, loss=None)
    for (which, k), t in outs.items():
        res[which][k] = t
    return res
