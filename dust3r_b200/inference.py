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


@torch.no_grad()
def inference(pairs, model, device, batch_size=8, verbose=True, keep_on_device=False):
    if verbose:
        print(f'>> Inference with model on {len(pairs)} image pairs')
    result = []
    multiple_shapes = not check_if_same_size(pairs)
    if multiple_shapes:  # force bs=1
        batch_size = 1
    for i in tqdm.trange(0, len(pairs), batch_size, disable=not verbose):
        res = loss_of_one_batch(collate_with_cat(pairs[i:i + batch_size]), model, None, device)
        result.append(res if keep_on_device else to_cpu(res))
    return collate_with_cat(result, lists=multiple_shapes)
