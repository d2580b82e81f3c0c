"""Host-side timeline of inference() (pinned staging / H2D / forward / D2H pipeline) for the bench e2e workload."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import build_model, synth_pairs_host
from dust3r_b200 import inference as inf
dev = torch.device('cuda:0')
net, cfg = build_model(dev)
pairs = synth_pairs_host(32, seed=99, pin=True)
keep = []
for it in range(6):
    inf._TRACE = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = inf.inference(pairs, net, dev, batch_size=32, verbose=False)
    _ = float(out['pred1']['conf'][0, 0, 0])
    t1 = time.perf_counter()
    tr = inf._TRACE
    deltas = [(tr[i][0], round((tr[i][1] - tr[i - 1][1]) * 1e3, 2)) for i in range(1, len(tr))]
    print(json.dumps(dict(call=it, total_ms=round((t1 - t0) * 1e3, 2), phases=deltas)), flush=True)
    if it < 2:
        keep.append(out)    # the first results stay alive -> later calls cannot recycle their pinned buffers
