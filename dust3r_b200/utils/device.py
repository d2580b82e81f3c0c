"""Container helpers of the inference path: moving nested batches between devices and merging per-batch results.
Public names follow dust3r/utils/device.py (to_device / to_cpu / to_numpy / to_cuda / collate_with_cat)."""
from __future__ import annotations

import numpy as np
import torch


def _map_leaves(obj, leaf_fn):
    """Apply leaf_fn to every non-container leaf of nested dict / list / tuple structures, keeping their types."""
    if isinstance(obj, dict):
        return {key: _map_leaves(val, leaf_fn) for key, val in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_leaves(val, leaf_fn) for val in obj)
    return leaf_fn(obj)


def to_device(batch, device, callback=None, non_blocking=False):
    """Move every tensor / ndarray leaf of `batch` to `device`; device == 'numpy' converts tensors to arrays instead.
    Strings, numbers and None pass through."""
    if callback is not None:
        batch = callback(batch)

    def move(leaf):
        if device == 'numpy':
            return leaf.detach().cpu().numpy() if torch.is_tensor(leaf) else leaf
        if isinstance(leaf, np.ndarray):
            leaf = torch.from_numpy(leaf)
        return leaf.to(device, non_blocking=non_blocking) if torch.is_tensor(leaf) else leaf

    return _map_leaves(batch, move)


todevice = to_device


def to_numpy(x):
    return to_device(x, 'numpy')


def to_cpu(x):
    return to_device(x, 'cpu')


def to_cuda(x):
    return to_device(x, 'cuda')


def listify(elems):
    """Flatten one level: [[a, b], [c]] -> [a, b, c]; tensors are split along their first dimension."""
    flat = []
    for group in elems:
        flat.extend(group)
    return flat


def collate_with_cat(parts, lists=False):
    """Merge the per-batch results of a loop into one structure of the same shape as a single result.

    parts is a list (or tuple) of results, or a dict of such lists.  Per field: tensors / arrays are concatenated
    along dim 0 (lists=True: split into one entry per sample instead, for mixed image sizes), dicts and tuples are
    merged member by member, python lists are chained, scalars and strings are kept as the list they came in, and a
    field that is None stays None."""
    if isinstance(parts, dict):
        return {key: collate_with_cat(val, lists=lists) for key, val in parts.items()}
    if not isinstance(parts, (list, tuple)):
        return None
    if not parts:
        return parts
    head, kind = parts[0], type(parts)
    if head is None:
        return None
    if isinstance(head, (bool, int, float, str)):
        return parts
    if isinstance(head, dict):
        return {key: collate_with_cat([part[key] for part in parts], lists=lists) for key in head}
    if isinstance(head, tuple):
        return kind(collate_with_cat(column, lists=lists) for column in zip(*parts))
    if torch.is_tensor(head) or isinstance(head, np.ndarray):
        tensors = [torch.from_numpy(part) if isinstance(part, np.ndarray) else part for part in parts]
        return listify(tensors) if lists else torch.cat(tensors)
    chained = kind()
    for part in parts:
        chained = chained + part
    return chained
