"""Host glue of the forward path (API mirror of dust3r/utils/misc.py:10-121)."""
from __future__ import annotations

import torch


def fill_default_args(kwargs, func):
    import inspect
    for k, v in inspect.signature(func).parameters.items():
        if v.default is not inspect.Parameter.empty:
            kwargs.setdefault(k, v.default)
    return kwargs


def freeze_all_params(modules):
    for module in modules:
        try:
            for n, param in module.named_parameters():
                param.requires_grad = False
        except AttributeError:
            module.requires_grad = False


def is_symmetrized(gt1, gt2):
    """True when the batch is [(a,b),(b,a),(c,d),(d,c),...] judged on the `instance` strings;
    a batch of one pair is never symmetrised (misc.py:32-40)."""
    x, y = gt1['instance'], gt2['instance']
    if len(x) == len(y) and len(x) == 1:
        return False
    ok = True
    for i in range(0, len(x), 2):
        ok = ok and (x[i] == y[i + 1]) and (x[i + 1] == y[i])
    return ok


def flip(tensor):
    return torch.stack((tensor[1::2], tensor[0::2]), dim=1).flatten(0, 1)


def interleave(tensor1, tensor2):
    res1 = torch.stack((tensor1, tensor2), dim=1).flatten(0, 1)
    res2 = torch.stack((tensor2, tensor1), dim=1).flatten(0, 1)
    return res1, res2


def transposed(dic):
    return {k: v.swapaxes(1, 2) for k, v in dic.items()}


def transpose_to_landscape(head, activate=True):
    """Wrap a head so portrait images are predicted in landscape and transposed back (misc.py:54-100)."""
    def wrapper_no(decout, true_shape):
        assert true_shape[0:1].allclose(true_shape), 'true_shape must be all identical'
        H, W = true_shape[0].cpu().tolist()
        return head(decout, (H, W))

    def wrapper_yes(decout, true_shape):
        B = len(true_shape)
        H, W = int(true_shape.min()), int(true_shape.max())
        height, width = true_shape.T
        is_landscape = (width >= height)
        is_portrait = ~is_landscape
        if is_landscape.all():
            return head(decout, (H, W))
        if is_portrait.all():
            return transposed(head(decout, (W, H)))

        def selout(ar): return [d[ar] for d in decout]
        l_result = head(selout(is_landscape), (H, W))
        p_result = transposed(head(selout(is_portrait), (W, H)))
        result = {}
        for k in l_result | p_result:
            x = l_result[k].new_empty((B,) + tuple(l_result[k].shape[1:]))
            x[is_landscape] = l_result[k]
            x[is_portrait] = p_result[k]
            result[k] = x
        return result

    return wrapper_yes if activate else wrapper_no


def invalid_to_nans(arr, valid_mask, ndim=999):
    if valid_mask is not None:
        arr = arr.clone()
        arr[~valid_mask] = float('nan')
    if arr.ndim > ndim:
        arr = arr.flatten(-2 - (arr.ndim - ndim), -2)
    return arr
