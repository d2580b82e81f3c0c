"""ncu target for round planning: isolated launches of the kernels that matter, at the model's shapes.
Usage: ncu --set full ... -k regex:"gemm|attention|upsample|layernorm" python scripts/ncu_kernels_target.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dust3r_b200 import _lib
lib = _lib.get_lib()
# GEMM shapes (M, N, K, flags): enc proj (resid), dec proj (resid), dec qkv (rope needs tables -> plain bias here), dec fc1 GELU
for (M, N, K, fl) in [(49152, 1024, 1024, 0x11), (24576, 768, 768, 0x11), (24576, 2304, 768, 0x1), (24576, 3072, 768, 0x3),
                      (49152, 4096, 1024, 0x3), (49152, 1024, 4096, 0x11)]:
    A = torch.randn((M, K), device='cuda').bfloat16(); B = torch.randn((N, K), device='cuda').bfloat16()
    bias = torch.randn((N,), device='cuda')
    out = torch.zeros((M, N), device='cuda', dtype=torch.float32 if fl & 0x18 else torch.bfloat16)
    for _ in range(2):
        _lib.check(lib.d3r_gemm_bf16(A.data_ptr(), B.data_ptr(), out.data_ptr(), bias.data_ptr(), None, None, M, N, K, N, fl, None, None, 0, 0, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
    del A, B, out
# attention (encoder shape)
Bn, Hh, N = 64, 16, 768
ld = 3 * Hh * 64
qkv = torch.randn((Bn, N, ld), device='cuda').bfloat16()
o = torch.empty((Bn, N, Hh * 64), device='cuda', dtype=torch.bfloat16)
for _ in range(2):
    _lib.check(lib.d3r_attention_hd64(qkv.data_ptr(), ld, qkv.data_ptr() + Hh * 128, ld, qkv.data_ptr() + Hh * 256, ld, o.data_ptr(), Hh * 64,
                                      Bn, Hh, N, N, 0.125, _lib.stream_ptr()))
torch.cuda.synchronize()
