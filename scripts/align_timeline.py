"""CTA timeline of one alignment iteration (debug aid): prints wave structure and the tail."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dust3r_b200 import _lib
from dust3r_b200.utils.synth import synth_pair_predictions
from dust3r_b200.cloud_opt import global_aligner
n = 8
edges = [(i, j) for i in range(n) for j in range(i)]
out = synth_pair_predictions(n, edges, 384, 512, seed=0)
net = global_aligner(out, 'cuda', verbose=False)
eng = net._get_engine(); net._engine_push(eng)
eng.run(10); torch.cuda.synchronize()
lib = _lib.get_lib()
lib.d3r_align_set_debug.argtypes = [C.c_void_p]
buf = torch.zeros((eng.n_chunks + 4, 4), dtype=torch.int64, device='cuda')
lib.d3r_align_set_debug(buf.data_ptr())
eng.run(1, reset_adam=False); torch.cuda.synchronize()
lib.d3r_align_set_debug(None)
traw = buf.cpu().numpy().astype(np.float64)
t = traw[:eng.n_chunks]
steps = traw[eng.n_chunks:].reshape(-1)[:6]
t0 = t[:, 0].min()
start, main_end, ex, tail = (t[:, 0] - t0) / 1e3, (t[:, 1] - t0) / 1e3, (t[:, 2] - t0) / 1e3, (t[:, 3] - t0) / 1e3
print('ctas', len(t), 'chunk_px', eng.chunk_px)
print('start  us: min %.1f p50 %.1f max %.1f' % (start.min(), np.median(start), start.max()))
print('main   us: dur p50 %.1f min %.1f max %.1f' % (np.median(main_end - start), (main_end - start).min(), (main_end - start).max()))
print('main_end max %.1f ; exit max %.1f' % (main_end.max(), ex.max()))
last = np.argmax(tail)
print('tail cta %d: main_end %.1f ticket-exit %.1f small_step_end %.1f' % (last, main_end[last], ex[last], tail[last]))
h = np.histogram(start, bins=12)
print('start hist', h[0].tolist(), [round(x, 1) for x in h[1].tolist()])
print('small-step stamps (us):', [round((x - t0) / 1e3, 1) for x in steps.tolist()])
