"""Quick device-side timing of the fused alignment step (used while iterating; bench.py is the contract)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dust3r_b200.utils.synth import synth_pair_predictions
from dust3r_b200.cloud_opt import global_aligner

def run(n, sym, H=384, W=512, niter=300):
    edges = [(i, j) for i in range(n) for j in range(i)]
    if sym: edges = edges + [(j, i) for i, j in edges]
    out = synth_pair_predictions(n, edges, H, W, seed=0)
    torch.manual_seed(0)
    net = global_aligner(out, 'cuda', verbose=False)
    eng = net._get_engine(); pull = net._engine_push(eng)
    eng.run(20)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); losses = eng.run(niter); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    by = eng.algorithmic_bytes_per_iter()
    print(json.dumps(dict(n=n, E=len(edges), niter=niter, ms_per_iter=ms / niter, iters_per_s=niter / ms * 1e3,
                          alg_bytes=by, GBps=by / (ms / niter) / 1e6, loss0=float(losses[0]), lossN=float(losses[-1]))))

if __name__ == '__main__':
    run(8, False); run(8, True)
    run(24, False, niter=100)
