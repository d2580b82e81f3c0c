"""inference(): batched pairwise forward (API mirror of dust3r/inference.py:14-78).

Same contract as the reference: takes the `make_pairs` list, returns {'view1','view2','pred1','pred2','loss'}
with every tensor ON CPU, concatenated over pairs in input order (lists when image sizes are mixed).
Differences are internal: each batch is one C-ABI call (d3r_forward_pairs), device->host copies of
the predictions go through pinned staging buffers and overlap the next batch's compute, and
`keep_on_device=True` (extension) skips the host round trip for callers that feed global_aligner next
(SURVEY §8f rank 2)."""
from __future__ import annotations

import os

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


_POOL = None
_TRACE = None   # set to a list to collect (label, perf_counter) host timestamps of the pipeline (diagnostics)


def _mark(label):
    if _TRACE is not None:
        import time
        _TRACE.append((label, time.perf_counter()))


def _copy_pool():
    global _POOL
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=max(1, min(8, (os.cpu_count() or 2) // 2)))
    return _POOL


def _fill_pinned(dst, tensors, row0, wait=True):
    """Copy each view's image rows into the pinned staging tensor `dst` starting at row `row0` (memcpy on a small
    thread pool: Tensor.copy_ releases the GIL, one thread saturates only ~10 GB/s of host bandwidth).
    Returns (next row, futures); with wait=False the copies are left running in the background."""
    jobs, r = [], row0
    for t in tensors:
        k = int(t.shape[0])
        jobs.append((dst[r:r + k], t))
        r += k
    futs = [_copy_pool().submit(d.copy_, t) for d, t in jobs]
    if wait:
        for f in futs:
            f.result()
        futs = []
    return r, futs


def _uploadable(t, dev):
    """Sources the copy stream can read directly: pinned host memory, or an image that already lives on the target GPU
    (load_images(..., device=...): resized and normalised there)."""
    if not t.is_cuda:
        return t.is_pinned()
    return t.device.index == (dev.index if dev.index is not None else torch.cuda.current_device())


def _micro_batch(batch_size):
    """Pairs per fused forward call.  A user batch of >= 16 pairs is run as two halves so that the host-side
    staging + H2D of one half and the D2H of the other overlap the GPU compute (per-pair results do not depend on
    the batch they are computed in); halves stay even so symmetrised (a,b),(b,a) neighbours are kept together."""
    if batch_size < 16:
        return batch_size
    half = (batch_size + 1) // 2
    return half + (half & 1)


@torch.no_grad()
def inference(pairs, model, device, batch_size=8, verbose=True, keep_on_device=False, return_images=True):
    """inference.py:55-72.  Returns {'view1','view2','pred1','pred2','loss'}; tensors on CPU (pinned) unless
    keep_on_device.  Software pipeline over micro-batches: images are gathered into pinned host memory (which is
    also the returned, collated view; return_images=False -- extension -- leaves 'img' out of the returned views and skips
    that copy where the upload does not need it), uploaded on a copy stream, run through one fused forward call, and the
    predictions are copied D2H on a second side stream into the final (whole pair list) pinned output -- the
    upload of batch k+1 and the download of batch k-1 overlap the compute of batch k."""
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

    _mark('begin')
    n = len(pairs)
    views = ([a for a, b in pairs], [b for a, b in pairs])
    rows = [sum(int(v['img'].shape[0]) for v in vs) for vs in views]
    assert rows[0] == rows[1], 'both views of a pair must hold the same number of images'
    proto = [vs[0]['img'] for vs in views]
    # pinned staging = the collated 'img' of the returned views; device copies of the whole pair list (a few MB / pair)
    img_pin = None    # allocated below, unless the caller does not want the images back and the upload does not stage through it
    img_dev = None    # device copies of both views: only the non-deduplicated path needs them (allocated there)
    meta_all = [{key: collate_with_cat([v[key] for v in vs]) for key in vs[0] if key != 'img'} for vs in views]
    _mark('alloc+meta')
    outs = None
    main = torch.cuda.current_stream(dev)
    up, side = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    up.wait_stream(main)
    mb = _micro_batch(batch_size)
    r0 = 0
    pending = []
    # pair lists from make_pairs reuse the same image dict in many pairs (n images -> up to n(n-1) pairs): upload every
    # distinct image once and let each batch address the ones it needs through index maps -- the encoder then runs once
    # per distinct image of a batch instead of twice per pair (its output for an image does not depend on what else
    # is in the batch, so results are unchanged)
    uniq_dev, gidx = None, ([], [])
    if hasattr(model, 'forward_indexed') and all(int(v['img'].shape[0]) == 1 for vs in views for v in vs):
        uniq, order = {}, []
        for k in range(2):
            for v in views[k]:
                t = v['img']
                key = (t.data_ptr(), tuple(t.shape), tuple(t.stride()))
                if key not in uniq:
                    uniq[key] = len(order)
                    order.append(t)
                gidx[k].append(uniq[key])
        if len(order) < 2 * n:
            uniq_dev = torch.empty((len(order),) + tuple(order[0].shape[1:]), dtype=order[0].dtype, device=dev)
            up.wait_stream(main)
            with torch.cuda.stream(up):
                if all(_uploadable(t, dev) for t in order):
                    for j, t in enumerate(order):
                        uniq_dev[j:j + 1].copy_(t, non_blocking=True)
                else:
                    stage = torch.empty(uniq_dev.shape, dtype=uniq_dev.dtype, pin_memory=True)
                    _fill_pinned(stage, order, 0, wait=True)
                    uniq_dev.copy_(stage, non_blocking=True)
            ev_uniq = torch.cuda.Event()
            ev_uniq.record(up)
            main.wait_event(ev_uniq)
    all_pinned = all(_uploadable(v['img'], dev) for vs in views for v in vs)
    if return_images or not (uniq_dev is not None or all_pinned):
        img_pin = [torch.empty((rows[k],) + tuple(proto[k].shape[1:]), dtype=proto[k].dtype, pin_memory=True) for k in range(2)]
    for i in tqdm.trange(0, n, mb, disable=not verbose):
        chunk = (views[0][i:i + mb], views[1][i:i + mb])
        r1 = r0
        srcs = [[v['img'] for v in chunk[k]] for k in range(2)]
        direct = all(_uploadable(t, dev) for ts in srcs for t in ts)
        indexed = uniq_dev is not None
        for k in range(2):
            # sources already in pinned memory are uploaded straight from where they are; the collated copy that the
            # caller gets back is then filled in the background, off the critical path
            if img_pin is None:
                r1 = r0 + sum(int(t.shape[0]) for t in srcs[k])
                continue
            r1, futs = _fill_pinned(img_pin[k], srcs[k], r0, wait=not (direct or indexed))
            pending.extend(futs)
        _mark('fill')
        if indexed:
            g1, g2 = gidx[0][i:i + mb], gidx[1][i:i + mb]
            ids = sorted(set(g1) | set(g2))
            loc = {g: j for j, g in enumerate(ids)}
            sel = uniq_dev if len(ids) == uniq_dev.shape[0] else uniq_dev.index_select(0, torch.tensor(ids, device=dev))
            _mark('h2d+meta')
            pred1, pred2 = model.forward_indexed(sel, [loc[g] for g in g1], [loc[g] for g in g2])
        else:
            # device staging: two micro-batch sized buffers per view, used alternately (the reference holds one batch on the
            # GPU at a time; a whole-pair-list copy would grow by 4.7 MB per pair at 512x384).  A buffer is reused two
            # micro-batches later: the upload stream first waits for the forward that last read it.
            if img_dev is None:
                per_item = [int(v['img'].shape[0]) for v in views[0]]
                cap = max(sum(per_item[c:c + mb]) for c in range(0, n, mb))
                img_dev = [[torch.empty((cap,) + tuple(proto[k].shape[1:]), dtype=proto[k].dtype, device=dev) for k in range(2)]
                           for _ in range(2)]
                dev_free = [None, None]
            slot = (i // mb) & 1
            nrow = r1 - r0
            with torch.cuda.stream(up):
                if dev_free[slot] is not None:
                    up.wait_event(dev_free[slot])
                for k in range(2):
                    if direct:
                        r = 0
                        for t in srcs[k]:
                            img_dev[slot][k][r:r + int(t.shape[0])].copy_(t, non_blocking=True)
                            r += int(t.shape[0])
                    else:
                        img_dev[slot][k][:nrow].copy_(img_pin[k][r0:r1], non_blocking=True)
            ev_up = torch.cuda.Event()
            ev_up.record(up)
            main.wait_event(ev_up)
            d = [dict({key: collate_with_cat([v[key] for v in chunk[k]]) for key in chunk[k][0] if key != 'img'},
                      img=img_dev[slot][k][:nrow]) for k in range(2)]
            _mark('h2d+meta')
            pred1, pred2 = model(d[0], d[1])
            dev_free[slot] = torch.cuda.Event()
            dev_free[slot].record(main)
        _mark('forward-enqueued')
        flat = {('pred1', k): v for k, v in pred1.items()}
        flat.update({('pred2', k): v for k, v in pred2.items()})
        if outs is None:
            if keep_on_device:
                outs = {key: torch.empty((rows[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=dev) for key, t in flat.items()}
            else:
                outs = {key: torch.empty((rows[0],) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True) for key, t in flat.items()}
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(side):
            side.wait_event(ev)
            for key, t in flat.items():
                outs[key][r0:r1].copy_(t, non_blocking=True)
                t.record_stream(side)
        _mark('d2h-enqueued')
        r0 = r1
    side.synchronize()
    for f in pending:
        f.result()
    _mark('synced')
    main.wait_stream(up)
    if return_images:
        res = dict(view1=dict(meta_all[0], img=img_pin[0]), view2=dict(meta_all[1], img=img_pin[1]), pred1={}, pred2={}, loss=None)
    else:
        res = dict(view1=dict(meta_all[0]), view2=dict(meta_all[1]), pred1={}, pred2={}, loss=None)
    for (which, k), t in outs.items():
        res[which][k] = t
    return res
