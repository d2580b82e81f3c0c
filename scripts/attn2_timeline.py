"""Cycle stamps of one persistent CTA of the split-row attention kernel (debug aid)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dust3r_b200 import _lib
lib = _lib.get_lib()
abl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 3
lib.d3r_set_attention_impl(impl + 10 * abl)
B, Hh, N = 64, 16, 768
ld = 3 * Hh * 64
qkv = torch.randn((B, N, ld), device='cuda').bfloat16()
out = torch.empty((B, N, Hh * 64), device='cuda', dtype=torch.bfloat16)
dbg = torch.zeros(4 * 16 + 64 * 64, dtype=torch.int64, device='cuda')
def f():
    _lib.check(lib.d3r_attention_hd64(qkv.data_ptr(), ld, qkv.data_ptr() + Hh * 128, ld, qkv.data_ptr() + Hh * 256, ld,
                                      out.data_ptr(), Hh * 64, B, Hh, N, N, 0.125, _lib.stream_ptr()))
for _ in range(3): f()
torch.cuda.synchronize()
lib.d3r_attention_set_debug.argtypes = [__import__('ctypes').c_void_p]
lib.d3r_attention_set_debug(dbg.data_ptr())
f(); torch.cuda.synchronize()
lib.d3r_attention_set_debug(None)
d = dbg[:64].cpu().numpy().reshape(4, 16)
t0 = d[d > 0].min()
names = ['sm:wait_s', 'sm:got_s', 'sm:exps_done', 'sm:o_done_seen', 'sm:P_stored', 'sm:xchg_done', 'sm:p_arrived', 'sm:epilogue_done',   # impl 3: 4 = P st issued, 5 = st complete
        
         'mma:S_begin', 'mma:k_full', 'mma:s_free', 'mma:S_issued', 'mma:wait_p', 'mma:p_seen', 'mma:v_full', 'mma:PV_issued']
print('ablation', abl)
for gi in range(4):
    ev = sorted((int(d[gi, k] - t0), names[k]) for k in range(16) if d[gi, k] > 0)
    print('block', (10 if impl == 3 else 6) + gi, ' '.join(f'{n}@{t}' for t, n in ev))
