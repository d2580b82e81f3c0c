"""Timeline of one traced CTA of the tcgen05 attention kernel (debug aid)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from dust3r_b200 import _lib
lib = _lib.get_lib()
lib.d3r_attention_set_debug.argtypes = [C.c_void_p]
B, Hh, N = 64, 16, 768
q = torch.randn((B, N, 3, Hh, 64), device='cuda').bfloat16()
out = torch.empty((B, N, Hh * 64), device='cuda', dtype=torch.bfloat16)
ld = 3 * Hh * 64
def run():
    _lib.check(lib.d3r_attention_hd64(q.data_ptr(), ld, q.data_ptr() + Hh * 64 * 2, ld, q.data_ptr() + 2 * Hh * 64 * 2, ld,
                                      out.data_ptr(), Hh * 64, B, Hh, N, N, 0.125, _lib.stream_ptr()))
for _ in range(3): run()
torch.cuda.synchronize()
buf = torch.zeros((B, 64), dtype=torch.int64, device='cuda')
lib.d3r_attention_set_debug(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
lib.d3r_attention_set_debug(None)
print('kernel ms', e0.elapsed_time(e1))
t = buf.cpu().numpy().astype(np.float64)
for z in (20, 40):
    sm, mm = t[z, :32], t[z, 32:]
    t0 = sm[0]
    print(f'--- image {z}: CTA start 0, end {(sm[31]-t0)/1e3:.2f} us')
    clk = sm[16:25]
    print('  block-3 softmax cycle deltas [wait_s, ld0, exp0, exp1, (o_done wait), (rescale), STS, fence+arrive]:', [int(clk[k+1]-clk[k]) for k in range(8)])
    for j in range(2):
        a = [(sm[1 + j * 5 + k] - t0) / 1e3 for k in range(5)]
        m = [(mm[1 + j * 3 + k] - t0) / 1e3 for k in range(3)]
        print(f' blk {j}: softmax wait_s {a[0]:.2f} got_s {a[1]:.2f} exps_done {a[2]:.2f} o_done {a[3]:.2f} p_ready {a[4]:.2f} | mma S_issued {m[0]:.2f} p_seen {m[1]:.2f} PV_issued {m[2]:.2f}')
