"""Host logic of the alignment engine (dust3r_b200/cloud_opt/engine.py) without a GPU: the C library is replaced by a RECORDING
stand-in and the torch.cuda entry points the engine touches by no-ops, so the engine is built on CPU tensors and everything it
would hand to the kernels can be inspected -- the packing table (pointers, observation offsets, loss coefficients), the CSR
image -> entry tables, the learning-rate / Adam schedule rows, the streaming work items and the algorithmic byte count that
bench.py divides by the kernel time (32*E*P + 24*n*P).  No kernel runs; the numerics are the `-m gpu` tests' job."""
import contextlib
import ctypes as C
import math
import types

import numpy as np
import pytest
import torch


class _RecordingLib:
    """Every d3r_* entry point returns 0 and records its arguments; the size queries answer like the real library."""
    CONSTS = {'d3r_align_stream_slots_per_item': 3, 'd3r_align_stream_warps_per_cta': 8, 'd3r_align_stream_max_window': 8,
              'd3r_sizeof_align_item': 64, 'd3r_align_chunk_pixels': 2048, 'd3r_sizeof_pack_entry': 32}

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def fn(*args):
            self.calls.append((name, args))
            if name == 'd3r_align_workspace_floats':
                return 4096
            return self.CONSTS.get(name, 0)
        return fn


@pytest.fixture()
def fake_cuda(monkeypatch):
    from dust3r_b200 import _lib
    lib = _RecordingLib()
    cpu = torch.device('cpu')
    monkeypatch.setattr(_lib, 'require_cuda_device', lambda d: cpu)
    monkeypatch.setattr(_lib, 'get_lib', lambda: lib)
    monkeypatch.setattr(_lib, 'check', lambda rc: None)
    monkeypatch.setattr(torch.cuda, 'device', lambda d: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, 'current_stream', lambda d=None: types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None))
    monkeypatch.setattr(torch.cuda, 'get_device_properties', lambda d: types.SimpleNamespace(multi_processor_count=148))
    return lib


def _engine(lib, n, edges, shapes, variant, alias=False, **kw):
    from dust3r_b200.cloud_opt.engine import AlignEngine
    g = torch.Generator().manual_seed(0)
    if alias:      # every edge shares one buffer per side: only pointers and sizes matter to the host logic
        h, w = shapes[0]
        p, c = torch.zeros((h, w, 3)), torch.ones((h, w))
        pred_i = pred_j = [p] * len(edges)
        conf_i = conf_j = [c] * len(edges)
    else:
        pred_i = [torch.randn(shapes[i] + (3,), generator=g) for i, j in edges]
        pred_j = [torch.randn(shapes[j] + (3,), generator=g) for i, j in edges]
        conf_i = [1 + torch.rand(shapes[i], generator=g) for i, j in edges]
        conf_j = [1 + torch.rand(shapes[j], generator=g) for i, j in edges]
    eng = AlignEngine(edges, shapes, pred_i, pred_j, conf_i, conf_j, 'cpu', variant=variant, **kw)
    return eng, (pred_i, pred_j, conf_i, conf_j)


def _pack_table(lib):
    from dust3r_b200.cloud_opt.engine import PACK_ENTRY
    name, args = [c for c in lib.calls if c[0] == 'd3r_align_pack_entries'][-1]
    ptr, n_entries, max_area, conf_mode, stream_flag = args[:5]
    buf = (C.c_char * (n_entries * PACK_ENTRY.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=PACK_ENTRY).copy(), max_area, conf_mode, stream_flag


def test_algorithmic_bytes_of_the_benchmark_configs(fake_cuda):
    """SURVEY §8d / BASELINE configs 3 and 5: the numerator of cloud_opt.roofline.  Config 3 at full size; the config-5 graph
    (50 views, 1225 pairs) at 1/64 of the pixels (its observation buffer would be 7.7 GB) -- the count is linear in P."""
    for n, (H, W), scale, want in ((8, (384, 512), 1, 213_909_504), (50, (48, 64), 64, 7_942_963_200)):
        edges = [(i, j) for i in range(n) for j in range(i)]
        eng, _ = _engine(fake_cuda, n, edges, [(H, W)] * n, 'stacked', alias=True, pix_stride=H * W)
        E, P = len(edges), H * W
        assert eng.algorithmic_bytes_per_iter() == 32 * E * P + 24 * n * P
        assert eng.algorithmic_bytes_per_iter() * scale == want
        assert eng.kernel == 'stream' and eng.total_obs == 2 * E * P          # multiples of 64 pixels: no slot padding
        assert eng.obs.numel() == 8 * E * P


@pytest.mark.parametrize('variant', ['stacked', 'per_edge'])
@pytest.mark.parametrize('kernel', ['stream', 'general'])
def test_packing_table_and_entry_csr(fake_cuda, variant, kernel):
    from dust3r_b200.cloud_opt.engine import SLOT_PX
    shapes = [(24, 32), (32, 24), (16, 48), (24, 32)]
    n = len(shapes)
    edges = [(0, 1), (1, 0), (2, 0), (3, 2), (1, 3)]
    eng, (pred_i, pred_j, conf_i, conf_j) = _engine(fake_cuda, n, edges, shapes, variant, kernel=kernel, conf_mode='log')
    E = len(edges)
    areas = [h * w for h, w in shapes]
    table, max_area, conf_mode, stream_flag = _pack_table(fake_cuda)
    assert len(table) == 2 * E and max_area == max(areas) and conf_mode == 1 and stream_flag == (1 if kernel == 'stream' else 0)
    ent_ptr, ent_edge = eng._ent_ptr.numpy(), eng._ent_edge.numpy()
    edge_ent, ent_coef, ent_off = eng._edge_ent.numpy(), eng._ent_coef.numpy(), eng._ent_obs_off.numpy()
    seen = set()
    off = 0
    tot = [sum(areas[i] for i, j in edges), sum(areas[j] for i, j in edges)]
    for img in range(n):
        ents = range(ent_ptr[img], ent_ptr[img + 1])
        degree = sum((i == img) + (j == img) for i, j in edges)
        assert len(ents) == degree
        for k in ents:
            e = int(ent_edge[k])
            side = 0 if edge_ent[e, 0] == k else 1
            assert edge_ent[e, side] == k and edges[e][side] == img and (e, side) not in seen
            seen.add((e, side))
            src_p = (pred_i if side == 0 else pred_j)[e]
            src_c = (conf_i if side == 0 else conf_j)[e]
            row = table[k]
            assert int(row['pts']) == src_p.data_ptr() and int(row['conf']) == src_c.data_ptr()      # fp32 contiguous inputs are read in place
            assert int(row['area']) == areas[img] and int(row['obs_off']) == off == int(ent_off[k])
            want = 1.0 / tot[side] if variant == 'stacked' else 1.0 / (areas[img] * E)
            assert row['coef'] == np.float32(want) == ent_coef[k]
            off += ((areas[img] + SLOT_PX - 1) // SLOT_PX) * SLOT_PX if kernel == 'stream' else areas[img]
    assert seen == {(e, s) for e in range(E) for s in (0, 1)} and eng.total_obs == off
    assert eng.obs_px == sum(areas[i] + areas[j] for i, j in edges)
    assert eng.max_deg == int(np.diff(ent_ptr).max())
    # the general kernel's CTA table: every image cut into chunks of chunk_px pixels, whole number of CTAs per image
    chunk_ptr, chunk_img = eng._chunk_ptr.numpy(), eng._chunk_img.numpy()
    assert 0 < eng.chunk_px <= 2048 and eng.n_chunks == len(chunk_img) == chunk_ptr[-1]
    for img in range(n):
        assert chunk_ptr[img + 1] - chunk_ptr[img] == math.ceil(areas[img] / eng.chunk_px)
        assert (chunk_img[chunk_ptr[img]:chunk_ptr[img + 1]] == img).all()
    if kernel == 'stream':
        assert eng.n_items > 0 and eng.stream_grid > 0 and eng.stream_ppt == 3 and 1 <= eng.stream_window <= 8
        assert eng._items_rev is not None and eng._items_rev.shape == eng._items.shape
    else:
        assert eng.n_items == 0 and eng._items is None


def test_odd_shapes_fall_back_to_the_general_kernel(fake_cuda):
    eng, _ = _engine(fake_cuda, 2, [(0, 1), (1, 0)], [(5, 7), (9, 3)], 'per_edge')
    assert eng.kernel == 'general'
    with pytest.raises(ValueError):
        _engine(fake_cuda, 2, [(0, 1), (1, 0)], [(5, 7), (9, 3)], 'per_edge', kernel='stream')
    with pytest.raises(ValueError):
        _engine(fake_cuda, 2, [(0, 1)], [(8, 8), (8, 8)], 'stacked', conf_mode='nope')


@pytest.mark.parametrize('schedule', ['cosine', 'linear'])
def test_schedule_rows_are_the_reference_learning_rates_and_adam_corrections(schedule):
    """base_opt.py:352-360 (lr of iteration `it` of `niter`) and torch.optim.Adam's bias corrections with betas (0.9, 0.9):
    step = lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)."""
    from dust3r_b200.cloud_opt.engine import AlignEngine
    niter, lr, lr_min = 37, 0.01, 1e-6
    rows = AlignEngine.make_schedule(niter, lr, schedule, lr_min)
    assert rows.shape == (niter, 4) and rows.dtype == np.float32
    for it in range(niter):
        t = it / niter
        cur = lr_min + (lr - lr_min) * (1 + math.cos(t * math.pi)) / 2 if schedule == 'cosine' else (1 - t) * lr + t * lr_min
        step = it + 1
        want = (cur, cur / (1 - 0.9 ** step), math.sqrt(1 - 0.9 ** step), 0.0)
        assert np.allclose(rows[it], np.float32(want), rtol=1e-7, atol=0), (it, rows[it], want)
    with pytest.raises(ValueError):
        AlignEngine.make_schedule(3, lr, 'step')
    # an adaptive torch Adam on a scalar reproduces the same per-step scale
    p = torch.nn.Parameter(torch.tensor([1.0]))
    opt = torch.optim.Adam([p], lr=1.0, betas=(0.9, 0.9), eps=0.0)
    for it in range(3):
        opt.zero_grad()
        (p * 2.0).sum().backward()
        before = p.detach().clone()
        opt.param_groups[0]['lr'] = float(rows[it, 0])
        opt.step()
        # constant gradient g: m_hat = g, v_hat = g^2 -> update = lr
        assert torch.allclose(before - p.detach(), torch.tensor([float(rows[it, 0])]), rtol=1e-5)


def test_small_parameter_layout_offsets(fake_cuda):
    eng, _ = _engine(fake_cuda, 3, [(0, 1), (1, 2), (2, 0), (0, 2)], [(8, 8)] * 3, 'stacked')
    n, E = 3, 4
    o = eng._offsets()
    assert o == dict(poses=0, focals=7 * n, pp=9 * n, pw=11 * n, adapt=11 * n + 8 * E)
    assert eng.n_small == 11 * n + 10 * E == eng.small.numel() == eng.small_m.numel() == eng.small_trainable.numel()
