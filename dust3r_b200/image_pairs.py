"""Pair-graph construction (API + bit-exact index behaviour of dust3r/image_pairs.py:12-104).

Pair ORDER is part of the contract (edges index every downstream buffer).  The sliding / log windows
therefore go through a Python `set` of int tuples exactly as the reference does (its iteration order is
a deterministic function of the insertion sequence for int tuples), and tests/golden/make_pairs.npz
pins the produced (idx1, idx2) lists against the reference for n in {2,3,8,50}."""
from __future__ import annotations

import numpy as np
import torch


def _window_size(scene_graph, default=3):
    try:
        return int(scene_graph.split('-')[1])
    except Exception:
        return default


def _pair_ids(n, scene_graph):
    ids = []
    if scene_graph == 'complete':
        ids = [(i, j) for i in range(n) for j in range(i)]
    elif scene_graph.startswith('swin'):
        cyclic = not scene_graph.endswith('noncyclic')
        win = _window_size(scene_graph)
        seen = set()
        for i in range(n):
            for off in range(1, win + 1):
                j = i + off
                if cyclic:
                    j %= n
                if j >= n:
                    continue
                seen.add((i, j) if i < j else (j, i))
        ids = list(seen)
    elif scene_graph.startswith('logwin'):
        cyclic = not scene_graph.endswith('noncyclic')
        win = _window_size(scene_graph)
        offsets = [2 ** k for k in range(win)]
        seen = set()
        for i in range(n):
            for j in [i - o for o in offsets] + [i + o for o in offsets]:
                if cyclic:
                    j %= n
                if j < 0 or j >= n or j == i:
                    continue
                seen.add((i, j) if i < j else (j, i))
        ids = list(seen)
    elif scene_graph.startswith('oneref'):
        ref = int(scene_graph.split('-')[1]) if '-' in scene_graph else 0
        ids = [(ref, j) for j in range(n) if j != ref]
    return ids


def make_pairs(imgs, scene_graph='complete', prefilter=None, symmetrize=True):
    pairs = [(imgs[i], imgs[j]) for i, j in _pair_ids(len(imgs), scene_graph)]
    if symmetrize:
        pairs += [(b, a) for a, b in pairs]
    if isinstance(prefilter, str) and prefilter.startswith('seq'):
        pairs = filter_pairs_seq(pairs, int(prefilter[3:]))
    if isinstance(prefilter, str) and prefilter.startswith('cyc'):
        pairs = filter_pairs_seq(pairs, int(prefilter[3:]), cyclic=True)
    return pairs


def sel(x, kept):
    if isinstance(x, dict):
        return {k: sel(v, kept) for k, v in x.items()}
    if isinstance(x, (torch.Tensor, np.ndarray)):
        return x[kept]
    if isinstance(x, (tuple, list)):
        return type(x)([x[k] for k in kept])


def _filter_edges_seq(edges, seq_dis_thr, cyclic=False):
    n = max(max(e) for e in edges) + 1
    kept = []
    for e, (i, j) in enumerate(edges):
        dis = abs(i - j)
        if cyclic:
            dis = min(dis, abs(i + n - j), abs(i - n - j))
        if dis <= seq_dis_thr:
            kept.append(e)
    return kept


def filter_pairs_seq(pairs, seq_dis_thr, cyclic=False):
    edges = [(a['idx'], b['idx']) for a, b in pairs]
    return [pairs[i] for i in _filter_edges_seq(edges, seq_dis_thr, cyclic=cyclic)]


def filter_edges_seq(view1, view2, pred1, pred2, seq_dis_thr, cyclic=False):
    edges = [(int(i), int(j)) for i, j in zip(view1['idx'], view2['idx'])]
    kept = _filter_edges_seq(edges, seq_dis_thr, cyclic=cyclic)
    print(f'>> Filtering edges more than {seq_dis_thr} frames apart: kept {len(kept)}/{len(edges)} edges')
    return sel(view1, kept), sel(view2, kept), sel(pred1, kept), sel(pred2, kept)
