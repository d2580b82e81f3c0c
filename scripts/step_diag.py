"""Step-time diagnostics: host-side launch time vs device time per forward step, drift over consecutive steps."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from bench import build_model, H, W
from dust3r_b200 import _lib
B = 32
net, cfg = build_model(torch.device('cuda:0'))
packed = net.repack()
imgs = torch.rand((2 * B, 3, H, W), device='cuda') * 2 - 1
idx1, idx2 = np.arange(B, dtype=np.int32), B + np.arange(B, dtype=np.int32)
for _ in range(3):
    packed.forward(imgs, idx1, idx2, B, H, W)
torch.cuda.synchronize()
n = 12
ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
host = []
ev[0].record()
for i in range(n):
    t0 = time.perf_counter()
    packed.forward(imgs, idx1, idx2, B, H, W)
    host.append((time.perf_counter() - t0) * 1e3)
    ev[i + 1].record()
torch.cuda.synchronize()
dev = [ev[i].elapsed_time(ev[i + 1]) for i in range(n)]
print(json.dumps(dict(kind='step_diag', host_ms=[round(h, 2) for h in host], dev_ms=[round(d, 2) for d in dev])))
# one step at a time with a sync + short sleep in between (cool GPU)
cool = []
for i in range(4):
    time.sleep(0.5)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); packed.forward(imgs, idx1, idx2, B, H, W); e1.record(); torch.cuda.synchronize()
    cool.append(round(e0.elapsed_time(e1), 2))
print(json.dumps(dict(kind='step_diag_cool', dev_ms=cool)))
