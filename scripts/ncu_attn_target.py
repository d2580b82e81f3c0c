"""ncu target: a few launches of the attention kernel on the encoder shape.  Usage: ncu ... python scripts/ncu_attn_target.py [impl]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dust3r_b200 import _lib
lib = _lib.get_lib()
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 3
lib.d3r_set_attention_impl(impl)
B, Hh, N = 64, 16, 768
ld = 3 * Hh * 64
qkv = torch.randn((B, N, ld), device='cuda').bfloat16()
out = torch.empty((B, N, Hh * 64), device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    _lib.check(lib.d3r_attention_hd64(qkv.data_ptr(), ld, qkv.data_ptr() + Hh * 128, ld, qkv.data_ptr() + Hh * 256, ld,
                                      out.data_ptr(), Hh * 64, B, Hh, N, N, 0.125, _lib.stream_ptr()))
torch.cuda.synchronize()
