"""Multi-GPU data parallelism of the pairwise forward (one process per GPU, torch.distributed / NCCL).

The reference runs inference on a single device (no collective on the inference path).  Pairs are
independent, so the pair list is split into contiguous per-rank slices (global order preserved -> pair
indexing stays bit-exact), every rank holds a replica of the weights, and ONE all-gather of the per-pair
outputs {pts3d, conf} x 2 (6.29 MB / pair at 512x384) rebuilds the full `inference()` result on every rank
before global alignment (BASELINE north_star; NVLink 5 / NVSwitch: any-to-any full bandwidth, so a plain
ring/NVLS all-gather is bandwidth-optimal).  Works with backend 'nccl' (GPU) and 'gloo' (CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .inference import inference
from .utils.device import collate_with_cat


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous balanced split: the first (n % world) ranks get one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def _all_gather_rows(t: torch.Tensor, counts, group=None):
    """Concatenate per-rank tensors with different leading sizes, in rank order."""
    world = dist.get_world_size(group)
    mx = max(counts)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = torch.empty((world * mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * mx: r * mx + counts[r]] for r in range(world)], dim=0)


@torch.no_grad()
def inference_sharded(pairs, model, device, batch_size=8, verbose=False, group=None, gather_device=None):
    """inference() over this rank's slice of `pairs` + all-gather -> the full result dict on every rank.

    Same return structure as inference(); tensors live on `gather_device` (default: CPU like the reference;
    pass the CUDA device to keep them resident for global_aligner)."""
    if not (dist.is_available() and dist.is_initialized()):
        return inference(pairs, model, device, batch_size=batch_size, verbose=verbose)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = [shard_bounds(len(pairs), world, r)[1] - shard_bounds(len(pairs), world, r)[0] for r in range(world)]
    lo, hi = shard_bounds(len(pairs), world, rank)
    local = inference(pairs[lo:hi], model, device, batch_size=batch_size, verbose=verbose, keep_on_device=True) if hi > lo else None
    backend = dist.get_backend(group)
    comm_dev = torch.device(device) if backend == 'nccl' else torch.device('cpu')
    tmpl = None
    if local is None:   # rank without work still takes part in the collective
        h, w = pairs[0][0]['img'].shape[-2:]
        tmpl = dict(pts=torch.zeros((0, h, w, 3)), conf=torch.zeros((0, h, w)))

    def get(d, k, proto):
        return (d[k] if local is not None else proto).to(comm_dev, torch.float32)
    p1 = _all_gather_rows(get(local['pred1'] if local else None, 'pts3d', tmpl['pts'] if tmpl else None), counts, group)
    c1 = _all_gather_rows(get(local['pred1'] if local else None, 'conf', tmpl['conf'] if tmpl else None), counts, group)
    p2 = _all_gather_rows(get(local['pred2'] if local else None, 'pts3d_in_other_view', tmpl['pts'] if tmpl else None), counts, group)
    c2 = _all_gather_rows(get(local['pred2'] if local else None, 'conf', tmpl['conf'] if tmpl else None), counts, group)
    out_dev = torch.device('cpu') if gather_device is None else torch.device(gather_device)
    # the views (images, indices) are inputs every rank already holds: rebuild them locally in global order
    view1 = collate_with_cat([a for a, b in pairs])
    view2 = collate_with_cat([b for a, b in pairs])
    return dict(view1=view1, view2=view2,
                pred1=dict(pts3d=p1.to(out_dev), conf=c1.to(out_dev)),
                pred2=dict(pts3d_in_other_view=p2.to(out_dev), conf=c2.to(out_dev)), loss=None)
