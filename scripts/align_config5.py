"""BASELINE configs[4] alignment leg on one GPU: 50 views -> 1225 pairs (symmetrize=False) at 512x384, predictions
synthesised directly in HBM (as after the all-gather), ModularPointCloudOptimizer (and PointCloudOptimizer for
comparison).  Prints one JSON line per run: iterations/s, algorithmic GB/s, fraction of the measured HBM peak."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from dust3r_b200.cloud_opt import global_aligner, GlobalAlignerMode


def synth_on_device(n, edges, H, W, dev, seed=0):
    g = torch.Generator(device=dev).manual_seed(seed)
    E = len(edges)
    off = torch.tensor([0.0, 0.0, 3.0], device=dev)
    ts = torch.from_numpy(np.int32([[H, W]] * E))
    mk = lambda: torch.randn((E, H, W, 3), generator=g, device=dev) + off
    cf = lambda: 1 + 5 * torch.rand((E, H, W), generator=g, device=dev)
    return dict(view1=dict(idx=[int(i) for i, j in edges], instance=[str(i) for i, j in edges], true_shape=ts),
                view2=dict(idx=[int(j) for i, j in edges], instance=[str(j) for i, j in edges], true_shape=ts),
                pred1=dict(pts3d=mk(), conf=cf()), pred2=dict(pts3d_in_other_view=mk(), conf=cf()), loss=None)


def run(n, mode, sym=False, H=384, W=512, niter=100, warm=10):
    dev = torch.device('cuda')
    edges = [(i, j) for i in range(n) for j in range(i)]
    if sym:
        edges = edges + [(j, i) for i, j in edges]
    out = synth_on_device(n, edges, H, W, dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    torch.manual_seed(0)
    net = global_aligner(out, dev, mode=GlobalAlignerMode[mode], verbose=False)
    eng = net._get_engine()
    net._engine_push(eng)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    eng.run(warm)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); losses = eng.run(niter); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    by = eng.algorithmic_bytes_per_iter()
    pk = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'] \
        if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')) else 6650.0
    gbs = by / (ms / niter) / 1e6
    print(json.dumps(dict(n=n, E=len(edges), mode=mode, niter=niter, build_s=round(t_build, 2), ms_per_iter=ms / niter,
                          iters_per_s=niter / ms * 1e3, alg_bytes=by, GBps=gbs, frac_hbm=gbs / pk, loss0=float(losses[0]),
                          lossN=float(losses[-1]), mem_GB=round(torch.cuda.max_memory_allocated() / 1e9, 1))), flush=True)
    del net, eng, out
    torch.cuda.empty_cache()


if __name__ == '__main__':
    which = sys.argv[1:] or ['c3', 'n24', 'c5']
    if 'c3' in which:
        run(8, 'PointCloudOptimizer', niter=300)
        run(8, 'ModularPointCloudOptimizer', niter=300)
    if 'n24' in which:
        run(24, 'PointCloudOptimizer', niter=100)
    if 'c5' in which:
        run(50, 'ModularPointCloudOptimizer', niter=100)
        run(50, 'PointCloudOptimizer', niter=100)
    if 'c5sym' in which:
        run(50, 'ModularPointCloudOptimizer', sym=True, niter=50)
