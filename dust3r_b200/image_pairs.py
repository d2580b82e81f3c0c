"""Pair-graph construction: which image pairs go through the network (dust3r/image_pairs.py:12-104).

Pair ORDER is part of the contract -- edges index every downstream buffer -- and for the windowed graphs the
reference's order is whatever iterating a Python `set` of int tuples gives, a deterministic function of the insertion
sequence.  The windowed graphs therefore insert the same tuples in the same sequence into a set here, and
tests/golden/make_pairs.npz pins the resulting (idx1, idx2) lists against the reference for n in {2, 3, 8, 50} x every
graph x symmetrize x prefilter."""
from __future__ import annotations

import numpy as np
import torch


def _suffix_int(spec, default):
    """'swin-5-noncyclic' -> 5, 'oneref' -> default."""
    parts = spec.split('-')
    try:
        return int(parts[1])
    except (IndexError, ValueError):
        return default


def _windowed(n, offsets, cyclic, self_pairs):
    """Undirected pairs {i, i + o} for every image i and offset o, collected through a set (see module docstring).
    A cyclic window longer than the sequence wraps onto the image itself: the reference keeps such (i, i) pairs for
    'swin' and drops them for 'logwin' -- `self_pairs` says which."""
    bag = set()
    for i in range(n):
        for o in offsets:
            j = (i + o) % n if cyclic else i + o
            if 0 <= j < n and (self_pairs or j != i):
                bag.add((min(i, j), max(i, j)))
    return list(bag)


def _graph_edges(n, spec):
    """(i, j) image-index pairs of the named scene graph; unknown names give no pairs (as the reference does)."""
    kind = spec.split('-')[0]
    cyclic = not spec.endswith('noncyclic')
    if kind == 'complete':
        return [(i, j) for i in range(n) for j in range(i)]
    if kind == 'swin':          # sliding window: every image with its next `size` neighbours
        size = _suffix_int(spec, 3)
        return _windowed(n, range(1, size + 1), cyclic, self_pairs=True)
    if kind == 'logwin':        # neighbours at distances 1, 2, 4, ... on both sides
        size = _suffix_int(spec, 3)
        steps = [2 ** k for k in range(size)]
        return _windowed(n, [-s for s in steps] + steps, cyclic, self_pairs=False)
    if kind == 'oneref':        # a star around one reference image
        ref = _suffix_int(spec, 0)
        return [(ref, j) for j in range(n) if j != ref]
    return []


def make_pairs(imgs, scene_graph='complete', prefilter=None, symmetrize=True):
    """List of (view_i, view_j) dict pairs; symmetrize appends every mirrored pair AFTER the originals; prefilter
    'seqN' / 'cycN' keeps pairs at most N frames apart (cyclically for 'cyc')."""
    pairs = [(imgs[i], imgs[j]) for i, j in _graph_edges(len(imgs), scene_graph)]
    if symmetrize:
        pairs = pairs + [(second, first) for first, second in pairs]
    if isinstance(prefilter, str):
        for tag, cyclic in (('seq', False), ('cyc', True)):
            if prefilter.startswith(tag):
                pairs = filter_pairs_seq(pairs, int(prefilter[len(tag):]), cyclic=cyclic)
    return pairs


def sel(x, kept):
    """Keep the entries `kept` of every tensor / array / list found in (nested dicts of) x."""
    if isinstance(x, dict):
        return {key: sel(val, kept) for key, val in x.items()}
    if isinstance(x, (torch.Tensor, np.ndarray)):
        return x[kept]
    if isinstance(x, (tuple, list)):
        return type(x)(x[k] for k in kept)
    return None


def _close_in_sequence(edges, max_gap, cyclic=False):
    """Positions of the edges whose two frame indices are at most `max_gap` apart (on a ring of n frames if cyclic)."""
    n = 1 + max(max(edge) for edge in edges)
    keep = []
    for pos, (i, j) in enumerate(edges):
        gap = abs(i - j)
        if cyclic:
            gap = min(gap, abs(i + n - j), abs(i - n - j))     # going round the ring either way
        if gap <= max_gap:
            keep.append(pos)
    return keep


def filter_pairs_seq(pairs, seq_dis_thr, cyclic=False):
    edges = [(first['idx'], second['idx']) for first, second in pairs]
    return [pairs[pos] for pos in _close_in_sequence(edges, seq_dis_thr, cyclic=cyclic)]


def filter_edges_seq(view1, view2, pred1, pred2, seq_dis_thr, cyclic=False):
    """Same filter on the collated output of inference()."""
    edges = [(int(i), int(j)) for i, j in zip(view1['idx'], view2['idx'])]
    keep = _close_in_sequence(edges, seq_dis_thr, cyclic=cyclic)
    print(f'>> Filtering edges more than {seq_dis_thr} frames apart: kept {len(keep)}/{len(edges)} edges')
    return sel(view1, keep), sel(view2, keep), sel(pred1, keep), sel(pred2, keep)
