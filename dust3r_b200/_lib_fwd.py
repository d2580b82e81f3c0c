"""ctypes prototypes of the forward-path entry points (include/dust3r_b200.h, "Path 1")."""
import ctypes as C

F_BIAS, F_GELU, F_RELU, F_OUT_F32, F_RESID_INPLACE = 1, 2, 4, 8, 16
F_ADD0, F_ADD1, F_OUT2_RELU, F_ROPE, F_OUT2_BF16 = 32, 64, 128, 256, 2048


def declare(lib):
    vp, i32, i64, u32, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
    lib.d3r_gemm_bf16.restype = C.c_int
    lib.d3r_gemm_bf16.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, i64, u32, vp, vp, i32, i32, i32, vp]
    lib.d3r_conv3x3_bf16.restype = C.c_int
    lib.d3r_conv3x3_bf16.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, u32, vp]
