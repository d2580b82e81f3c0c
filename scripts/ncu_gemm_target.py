"""ncu target: a few launches of chosen GEMM epilogue variants.  Usage: ncu ... python scripts/ncu_gemm_target.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dust3r_b200 import _lib
lib = _lib.get_lib()
for (M, N, K, fl) in [(49152, 1024, 1024, 0x11), (49152, 4096, 1024, 0x3), (49152, 1024, 1024, 0x1), (24576, 768, 768, 0x11)]:
    A = torch.randn((M, K), device='cuda').bfloat16(); B = torch.randn((N, K), device='cuda').bfloat16()
    bias = torch.randn((N,), device='cuda')
    out = torch.zeros((M, N), device='cuda', dtype=torch.float32 if fl & 0x18 else torch.bfloat16)
    for _ in range(2):
        _lib.check(lib.d3r_gemm_bf16(A.data_ptr(), B.data_ptr(), out.data_ptr(), bias.data_ptr(), None, None, M, N, K, N, fl, None, None, 0, 0, 0, _lib.stream_ptr()))
    torch.cuda.synchronize()
