// Fused softmax(q k^T / sqrt(d)) v for head dim 64 (all DUSt3R heads), bf16 in / bf16 out, fp32 softmax
// and accumulation.  Replaces the materialised (B,H,N,N) fp32 attention matrix of
// croco/models/blocks.py:105-109 / :161-165 with an on-chip streaming softmax.
//
// v1 kernel: 64 queries per CTA (4 warps x 16 rows), keys/values streamed in 64-row tiles through a
// double-buffered, XOR-swizzled shared-memory ring (cp.async), tensor-core math via
// mma.sync.m16n8k16 (legacy HMMA path).  A tcgen05/TMEM version replaces it in a later round; this one
// pins the numerics and the interface.
#include "d3r_common.cuh"
#include "elementwise.h"
#include "prof.h"
#include "pdl.cuh"
#include <cuda_bf16.h>

namespace d3r {
namespace attn {

constexpr int BM = 64, BN = 64, D = 64, kThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// tile of 64 rows x 64 bf16 (128 B per row = 8 chunks of 16 B), chunk index XOR-swizzled by (row & 7)
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__device__ __forceinline__ void load_tile(uint32_t smem_base, const __nv_bfloat16* g, long long ld, int row0, int nrows, int tid) {
  // 64 rows x 8 chunks = 512 x 16 B; 128 threads -> 4 each
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * kThreads;
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (row0 + r) < nrows;
    const __nv_bfloat16* src = g + (long long)(ok ? row0 + r : 0) * ld + c * 8;
    cp_async16(smem_base + tile_off(r, c), src, ok);
  }
}

__global__ void __launch_bounds__(kThreads) attention_kernel(const __nv_bfloat16* __restrict__ q, long long ldq,
                                                             const __nv_bfloat16* __restrict__ k, long long ldk,
                                                             const __nv_bfloat16* __restrict__ v, long long ldv,
                                                             __nv_bfloat16* __restrict__ out, long long ldo, int Nq, int Nk,
                                                             float scale_log2) {
  pdl::sync_with_predecessor();   // PDL: nothing above touches memory produced by other kernels
  __shared__ __align__(128) uint8_t s_q[BM * 128];
  __shared__ __align__(128) uint8_t s_k[2][BN * 128];
  __shared__ __align__(128) uint8_t s_v[2][BN * 128];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* qg = q + (long long)b * Nq * ldq + h * D;
  const __nv_bfloat16* kg = k + (long long)b * Nk * ldk + h * D;
  const __nv_bfloat16* vg = v + (long long)b * Nk * ldv + h * D;

  load_tile(smem_u32(s_q), qg, ldq, q0, Nq, tid);
  load_tile(smem_u32(s_k[0]), kg, ldk, 0, Nk, tid);
  load_tile(smem_u32(s_v[0]), vg, ldv, 0, Nk, tid);
  cp_commit();
  const int nblk = (Nk + BN - 1) / BN;

  uint32_t qf[4][4];  // A fragments of this warp's 16 query rows, 4 k-steps of 16
  float o[8][4];      // 16 x 64 output accumulator (8 n-tiles of 8)
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j][0] = o[j][1] = o[j][2] = o[j][3] = 0.f;
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;  // rows lane/4 and lane/4 + 8

  for (int blk = 0; blk < nblk; ++blk) {
    const int buf = blk & 1;
    if (blk + 1 < nblk) {
      load_tile(smem_u32(s_k[buf ^ 1]), kg, ldk, (blk + 1) * BN, Nk, tid);
      load_tile(smem_u32(s_v[buf ^ 1]), vg, ldv, (blk + 1) * BN, Nk, tid);
      cp_commit();
      cp_wait<1>();
    } else {
      cp_wait<0>();
    }
    __syncthreads();
    if (blk == 0) {
      const uint32_t sq = smem_u32(s_q);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 15), c = ks * 2 + (lane >> 4);
        ldsm_x4(sq + tile_off(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
    const uint32_t sk = smem_u32(s_k[buf]);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // pairs of 8-key n-tiles
        const int r = jp * 16 + (lane & 7) + ((lane >> 4) << 3);
        const int c = ks * 2 + ((lane >> 3) & 1);
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sk + tile_off(r, c), b0, b1, b2, b3);
        mma16816(s[2 * jp], qf[ks], b0, b1);
        mma16816(s[2 * jp + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask keys beyond Nk, online softmax ----
    const int key0 = blk * BN + (lane & 3) * 2;
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int kk = key0 + j * 8;
      if (kk >= Nk) { s[j][0] = -INFINITY; s[j][2] = -INFINITY; }
      if (kk + 1 >= Nk) { s[j][1] = -INFINITY; s[j][3] = -INFINITY; }
      mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
      mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float corr0 = exp2f((m0 - mx0) * scale_log2), corr1 = exp2f((m1 - mx1) * scale_log2);
    m0 = mx0; m1 = mx1;
    const float ms0 = mx0 * scale_log2, ms1 = mx1 * scale_log2;
    float rs0 = 0.f, rs1 = 0.f;
    uint32_t pf[4][4];  // P as A fragments for 4 k-steps of 16 keys
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float p0 = exp2f(s[j][0] * scale_log2 - ms0), p1 = exp2f(s[j][1] * scale_log2 - ms0);
      const float p2 = exp2f(s[j][2] * scale_log2 - ms1), p3 = exp2f(s[j][3] * scale_log2 - ms1);
      rs0 += p0 + p1;
      rs1 += p2 + p3;
      pf[j >> 1][(j & 1) * 2 + 0] = pack2(p0, p1);
      pf[j >> 1][(j & 1) * 2 + 1] = pack2(p2, p3);
    }
    l0 = l0 * corr0 + rs0;
    l1 = l1 * corr1 + rs1;
#pragma unroll
    for (int j = 0; j < 8; ++j) { o[j][0] *= corr0; o[j][1] *= corr0; o[j][2] *= corr1; o[j][3] *= corr1; }
    // ---- O += P V ----
    const uint32_t sv = smem_u32(s_v[buf]);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {      // 16 keys
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {    // pairs of 8-wide d tiles
        const int r = ks * 16 + (lane & 15);
        const int c = jp * 2 + (lane >> 4);
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sv + tile_off(r, c), b0, b1, b2, b3);
        mma16816(o[2 * jp], pf[ks], b0, b1);
        mma16816(o[2 * jp + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();
  }
  // ---- normalise and store ----
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float i0 = 1.f / l0, i1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + (lane >> 2), r1 = r0 + 8;
  __nv_bfloat16* og = out + (long long)b * Nq * ldo + h * D + (lane & 3) * 2;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (r0 < Nq) *reinterpret_cast<uint32_t*>(og + (long long)r0 * ldo + j * 8) = pack2(o[j][0] * i0, o[j][1] * i0);
    if (r1 < Nq) *reinterpret_cast<uint32_t*>(og + (long long)r1 * ldo + j * 8) = pack2(o[j][2] * i1, o[j][3] * i1);
  }
}

static int g_impl = 3;  // 3 (default): tcgen05 split-row kernel with P in TMEM (attention_tc3.cu), 2: tcgen05 split-row kernel with P through
                        // shared memory (attention_tc2.cu), 1: tcgen05 row-per-thread kernel (attention_tc.cu), 0: mma.sync
                        // streaming kernel (A/B reference)
void set_tc2_ablation(int a);
void set_tc3_ablation(int a);
int attention_hd64_tc3(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                       long long ldo, int B, int heads, int Nq, int Nk, float scale, cudaStream_t st);
void set_tc_occupancy_pad(int bytes);
// impl 1 + 10k (debug): k x 16 KB of extra shared memory per CTA
void set_impl(int impl) {
  g_impl = impl % 10;
  if (g_impl == 1) { set_tc_occupancy_pad((impl / 10) * 16 * 1024); set_tc2_ablation(0); }
  else if (g_impl == 3) { set_tc_occupancy_pad(0); set_tc3_ablation(impl / 10); }
  else { set_tc_occupancy_pad(0); set_tc2_ablation(impl / 10); }
}   // impl 12/22/32: timing ablations of impl 2 (debug)

int attention_hd64(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                   long long ldo, int B, int heads, int Nq, int Nk, float scale, cudaStream_t st) {
  if (g_impl == 3) return attention_hd64_tc3(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, scale, st);
  if (g_impl == 2) return attention_hd64_tc2(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, scale, st);
  if (g_impl == 1) return attention_hd64_tc(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, scale, st);
  D3R_CHECK_ARG(q && k && v && out && B > 0 && heads > 0 && Nq > 0 && Nk > 0, "attention: bad arguments");
  D3R_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0, "attention: row strides must keep 16-byte alignment");
  dim3 grid((Nq + BM - 1) / BM, heads, B);
  prof::Scope scope("attention_hd64", st, 4.0 * double(B) * heads * double(Nq) * double(Nk) * 64.0);
  D3R_CUDA(pdl::launch(attention_kernel, grid, dim3(kThreads), 0, st, (const __nv_bfloat16*)q, ldq, (const __nv_bfloat16*)k, ldk,
                       (const __nv_bfloat16*)v, ldv, (__nv_bfloat16*)out, ldo, Nq, Nk, scale * 1.4426950408889634f));
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}

}  // namespace attn
}  // namespace d3r

namespace d3r { namespace attn { int attention_set_debug(void* dev_buf); } }
extern "C" int d3r_attention_set_debug(void* dev_buf) { return d3r::attn::attention_set_debug(dev_buf); }

extern "C" void d3r_set_attention_impl(int32_t impl) { d3r::attn::set_impl(impl); }

extern "C" int d3r_attention_hd64(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                                  int64_t ldo, int32_t B, int32_t heads, int32_t Nq, int32_t Nk, float scale, void* stream) {
  return d3r::attn::attention_hd64(q, ldq, k, ldk, v, ldv, out, ldo, B, heads, Nq, Nk, scale, (cudaStream_t)stream);
}
