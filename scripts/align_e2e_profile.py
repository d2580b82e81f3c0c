"""cProfile of the public alignment API (global_aligner + compute_global_alignment) on the bench workload."""
import sys, os, time, cProfile, pstats, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dust3r_b200.utils.synth import synth_pair_predictions
from dust3r_b200.cloud_opt import global_aligner
n, H, W = 8, 384, 512
edges = [(i, j) for i in range(n) for j in range(i)]
out = synth_pair_predictions(n, edges, H, W, seed=0)
for side in ('pred1', 'pred2'):
    out[side] = {k: (v.pin_memory() if torch.is_tensor(v) else v) for k, v in out[side].items()}
def run():
    torch.manual_seed(0)
    net = global_aligner(out, 'cuda', verbose=False)
    t1 = time.perf_counter()
    loss = net.compute_global_alignment(init=None, niter=300, schedule='cosine', lr=0.01)
    torch.cuda.synchronize()
    return t1, loss
run()
torch.cuda.synchronize()
t0 = time.perf_counter(); t1, loss = run(); t2 = time.perf_counter()
print(f'global_aligner {1e3 * (t1 - t0):.1f} ms, compute_global_alignment {1e3 * (t2 - t1):.1f} ms, loss {loss:.4f}')
pr = cProfile.Profile(); pr.enable(); run(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28); print(s.getvalue()[:5000])
