// tcgen05 / TMEM fused attention for head dim 64 (sm_100a).
//
//   O = softmax(Q K^T * scale) V      per (image b, head h), bf16 in/out, fp32 softmax + accumulation
//
// CTA = (b, h, 128-query tile), 192 threads, 2 CTAs co-resident per SM (112 KB smem, 256 TMEM columns each)
// so one CTA's softmax (MUFU/FMA bound) overlaps the other's tensor-core work:
//   warp 0      TMA producer : Q tile once, then K_j / V_j (128 keys x 64) into 2-deep 128B-swizzled rings
//   warp 1      MMA issuer   : S = Q K_j^T  (UMMA 128x128x16 x4, K-major A and B)  -> TMEM cols [0,128)
//                              O += P V_j   (UMMA 128x64x16 x8, A = P from smem, B = V_j MN-major) -> cols [128,192)
//   warps 2..5  softmax      : thread = query row.  one tcgen05.ld pass pulls the S row into registers (S is
//                              released at once), max / exp2 / row sum -> P (bf16) stored to smem in the canonical
//                              K-major SW128 layout; O in TMEM is rescaled lazily (only when the running max grew
//                              by more than 2^8); final 1/l normalisation + store.
// Keys beyond Nk are masked to -inf (TMA zero-fills rows past the image because the tensor maps are 3-D
// {cols, tokens, images}).
#include "d3r_common.cuh"
#include "sm100_ptx.cuh"
#include "elementwise.h"
#include "prof.h"
#include "attention_tc_common.cuh"
#include "pdl.cuh"
#include <type_traits>

namespace d3r {
namespace attn {

namespace tc {

constexpr int BQ = 128, BK = 64, D = 64;   // 64-key blocks: 64 KB smem + 128 TMEM columns per CTA -> 3 CTAs / SM
constexpr int kThreads = 192;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB: 128 rows x 128 B (Q tile, one P atom)
constexpr int kKVBytes = BK * 64 * 2;     // one K or V block: BK rows x 128 B
constexpr int kCTAsPerSM = (BK == 64) ? 3 : 2;
constexpr int kSmemBytes = kTileBytes /*Q*/ + 2 * kKVBytes /*K ring*/ + 2 * kKVBytes /*V ring*/ +
                           (BK / 64) * kTileBytes /*P: one SW128 atom per 64 keys*/ + 128 /*barriers*/;
constexpr int kTmemCols = (BK + 64 <= 128) ? 128 : 256;
constexpr uint32_t kColS = 0, kColO = BK;

using namespace tcc;

// optional timeline instrumentation (debug aid)
__device__ unsigned long long* g_attn_dbg = nullptr;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ATT_STAMP(slot) do { if (dbg && (slot) < 16) dbg[(slot)] = gtime(); } while (0)
#define ATT_CLK(k) do { if (dbg && threadIdx.x == 64 && j == 3) dbg[16 + (k)] = (unsigned long long)clock64(); } while (0)

__global__ void __launch_bounds__(kThreads, kCTAsPerSM)
attention_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, long long ldo, int Nq, int Nk,
                    int q_col0, int k_col0, int v_col0, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzle atoms need 1024-byte alignment
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + kTileBytes;
  uint8_t* s_v = s_k + 2 * kKVBytes;
  uint8_t* s_p = s_v + 2 * kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + (BK / 64) * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_ready = bars + 9;   // S_j in TMEM
  uint64_t* s_free = bars + 10;   // softmax finished reading S_j
  uint64_t* p_ready = bars + 11;  // P_j in smem (+ O rescaled)
  uint64_t* o_done = bars + 12;   // P V_j accumulated
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BQ, h = blockIdx.y, b = blockIdx.z;
  const int nblk = (Nk + BK - 1) / BK;
  // 64 stamps per traced CTA: [0..31] softmax thread (row 0 of warp 2), [32..63] MMA thread
  unsigned long long* dbg = nullptr;
  if (g_attn_dbg && blockIdx.x == 2 && blockIdx.y == 3) {
    if (threadIdx.x == 64) dbg = g_attn_dbg + size_t(blockIdx.z) * 64;
    if (threadIdx.x == 32) dbg = g_attn_dbg + size_t(blockIdx.z) * 64 + 32;
  }
  ATT_STAMP(0);

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_k);
    ptx::prefetch_tmap(&tmap_v);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(ptx::smem_u32(q_full), 1);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&k_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&k_empty[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&v_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&v_empty[s]), 1);
    }
    ptx::mbar_init(ptx::smem_u32(s_ready), 1);
    ptx::mbar_init(ptx::smem_u32(s_free), 4);
    ptx::mbar_init(ptx::smem_u32(p_ready), 4);
    ptx::mbar_init(ptx::smem_u32(o_done), 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl::sync_with_predecessor();   // set-up done; from here on the kernel reads q/k/v written by its predecessor

  if (warp == 0) {
    // ================= TMA producer =================
    if (ptx::elect_one()) {
      ptx::mbar_arrive_expect_tx(ptx::smem_u32(q_full), kTileBytes);
      ptx::tma_load_3d(ptx::smem_u32(s_q), &tmap_q, ptx::smem_u32(q_full), q_col0 + h * D, q0, b);
      for (int j = 0; j < nblk; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        ptx::mbar_wait(ptx::smem_u32(&k_empty[st]), ph ^ 1);
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&k_full[st]), kKVBytes);
        ptx::tma_load_3d(ptx::smem_u32(s_k + st * kKVBytes), &tmap_k, ptx::smem_u32(&k_full[st]), k_col0 + h * D, j * BK, b);
        ptx::mbar_wait(ptx::smem_u32(&v_empty[st]), ph ^ 1);
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(&v_full[st]), kKVBytes);
        ptx::tma_load_3d(ptx::smem_u32(s_v + st * kKVBytes), &tmap_v, ptx::smem_u32(&v_full[st]), v_col0 + h * D, j * BK, b);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, BK, 0, 0);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(128, 64, 0, 1);  // B (=V) is MN-major
      const uint32_t d_s = tmem_base + kColS, d_o = tmem_base + kColO;
      const uint64_t dq = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(s_q));
      auto issue_s = [&](int j) {
        const int st = j & 1;
        ptx::mbar_wait(ptx::smem_u32(&k_full[st]), (j >> 1) & 1);
        if (j > 0) ptx::mbar_wait(ptx::smem_u32(s_free), (j - 1) & 1);  // softmax has drained S_{j-1}
        ptx::tc_fence_after();
        const uint64_t dk = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(s_k + st * kKVBytes));
#pragma unroll
        for (int k = 0; k < D / 16; ++k) ptx::umma_bf16_ss(d_s, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k ? 1u : 0u);
        ptx::umma_commit(ptx::smem_u32(&k_empty[st]));
        ptx::umma_commit(ptx::smem_u32(s_ready));
        ATT_STAMP(1 + j * 3);
      };
      ptx::mbar_wait(ptx::smem_u32(q_full), 0);
      issue_s(0);
      for (int j = 0; j < nblk; ++j) {
        if (j + 1 < nblk) issue_s(j + 1);
        const int st = j & 1;
        ptx::mbar_wait(ptx::smem_u32(p_ready), j & 1);
        ATT_STAMP(2 + j * 3);
        ptx::mbar_wait(ptx::smem_u32(&v_full[st]), (j >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t pv = ptx::smem_u32(s_v + st * kKVBytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          const uint64_t dp = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(s_p) + (k >> 2) * kTileBytes) + uint64_t(2 * (k & 3));
          const uint64_t dv = umma_desc_mnmajor_sw128(pv + k * 16 * 128);
          ptx::umma_bf16_ss(d_o, dp, dv, idesc_o, (j | k) ? 1u : 0u);
        }
        ptx::umma_commit(ptx::smem_u32(&v_empty[st]));
        ptx::umma_commit(ptx::smem_u32(o_done));
        ATT_STAMP(3 + j * 3);
      }
    }
  } else {
    // ================= softmax / correction / epilogue warps =================
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + kColS, t_o = tmem_base + lane_addr + kColO;
    // m_ref: the exponent reference baked into P, O and l.  Any fixed reference gives the same softmax after the
    // final division, so blocks j >= 1 exponentiate against the reference they inherit (no separate max pass on
    // their critical path) while tracking the block max alongside the MUFU-bound exp loop.  Only when a row max
    // outgrows the reference by more than 2^kLazy is the reference moved: O (TMEM) and l are multiplied by the
    // correction before the next P V accumulates (p <= 2^kLazy in between, harmless in fp32 / bf16).
    constexpr float kLazy = 8.f;
    float m_ref = -INFINITY, l = 0.f, corr_pending = 1.f;
    bool pending = false;
    for (int j = 0; j < nblk; ++j) {
      const int nvalid = min(BK, Nk - j * BK);
      ATT_STAMP(1 + j * 5);
      ATT_CLK(0);
      ptx::mbar_wait(ptx::smem_u32(s_ready), j & 1);
      ptx::tc_fence_after();
      ATT_STAMP(2 + j * 5);
      float rs0 = 0.f, rs1 = 0.f, bm0 = -INFINITY, bm1 = -INFINITY;
      uint32_t packed[BK / 2];
      float ms;
      // p = exp2(s*c - m_ref*c), row sum, running block max, bf16 pack -- all in registers: the MUFU-bound part of
      // the block does not depend on the previous P V, so it runs while that MMA is still in flight.
      // `masked` is a compile-time tag: only the ragged last key block pays for the per-element masking
      auto exp_chunk_t = [&](const uint32_t* r, int c, auto masked) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
          if constexpr (decltype(masked)::value) {   // masked keys contribute p = 0
            if (c * 32 + i >= nvalid) s0 = -INFINITY;
            if (c * 32 + i + 1 >= nvalid) s1 = -INFINITY;
          }
          const float p0 = fast_exp2(fmaf(s0, scale_log2, -ms));
          const float p1 = fast_exp2(fmaf(s1, scale_log2, -ms));
          rs0 += p0;
          rs1 += p1;
          bm0 = fmaxf(bm0, s0);
          bm1 = fmaxf(bm1, s1);
          packed[c * 16 + (i >> 1)] = pack2(p0, p1);
        }
      };
      const bool ragged = nvalid < BK;   // warp-uniform
      auto exp_chunk = [&](const uint32_t* r, int c) {
        if (ragged) exp_chunk_t(r, c, std::true_type{});
        else exp_chunk_t(r, c, std::false_type{});
      };
      constexpr int kChunks = BK / 32;
      if (j == 0) {
        // first block: the reference is this block's row max, so the whole row is needed before any exp
        uint32_t sr[BK];
#pragma unroll
        for (int c = 0; c < kChunks; ++c) ptx::tmem_ld_32x32b_x32(t_s + c * 32, sr + c * 32);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(s_free));
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < BK; ++i)
          if (i < nvalid) mx = fmaxf(mx, __uint_as_float(sr[i]));
        m_ref = mx;
        ms = m_ref * scale_log2;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) exp_chunk(sr + c * 32, c);
      } else {
        // later blocks exponentiate against the inherited reference: TMEM reads (64 B/clk/SM, as scarce as the
        // MUFU) are software-pipelined chunk by chunk under the exp work instead of being a serial pre-pass
        ms = m_ref * scale_log2;
        uint32_t rbuf[2][32];
        ATT_CLK(1);
        ptx::tmem_ld_32x32b_x32(t_s + 0, rbuf[0]);
        ptx::tmem_ld_wait();
        ATT_CLK(2);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          if (c + 1 < kChunks) ptx::tmem_ld_32x32b_x32(t_s + (c + 1) * 32, rbuf[(c + 1) & 1]);
          if (c + 1 == kChunks) {
            // S_j fully read -> the MMA warp may overwrite it with S_{j+1}
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(s_free));
          }
          exp_chunk(rbuf[c & 1], c);
          ATT_CLK(3 + c);
          if (c + 1 < kChunks) ptx::tmem_ld_wait();
        }
      }
      ATT_STAMP(3 + j * 5);
      ATT_CLK(5);
      // the previous P V must have consumed P (smem) and updated O before we touch either
      if (j > 0) {
        ptx::mbar_wait(ptx::smem_u32(o_done), (j - 1) & 1);
        ptx::tc_fence_after();
      }
      ATT_STAMP(4 + j * 5);
      ATT_CLK(6);
      if (pending) {   // warp-uniform
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_o + c * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr_pending);
          tmem_st_32x32b_x32(t_o + c * 32, r);
        }
        tmem_st_wait();
        pending = false;
      }
      // swizzled store of P (K-major SW128: 16-byte chunk index ^ (row & 7)); one 16 KB atom per 64 keys
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint8_t* atom = s_p + (c >> 1) * kTileBytes + row * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int chunk = (c & 1) * 4 + q;
          *reinterpret_cast<uint4*>(atom + ((chunk ^ (row & 7)) << 4)) =
              make_uint4(packed[c * 16 + 4 * q], packed[c * 16 + 4 * q + 1], packed[c * 16 + 4 * q + 2], packed[c * 16 + 4 * q + 3]);
        }
      }
      ATT_CLK(7);
      l += rs0 + rs1;
      // make the generic-proxy smem writes of P visible to the tensor core (async proxy), then publish
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(p_ready));
      ATT_CLK(8);
      ATT_STAMP(5 + j * 5);
      // move the reference if some row of this warp outgrew it (decision is warp-uniform so the TMEM round trip
      // above stays convergent); applied before the next block's P V
      const float bmax = fmaxf(bm0, bm1);
      if (__any_sync(0xffffffffu, (bmax - m_ref) * scale_log2 > kLazy)) {
        const float m_new = fmaxf(m_ref, bmax);
        corr_pending = fast_exp2((m_ref - m_new) * scale_log2);
        l *= corr_pending;
        m_ref = m_new;
        pending = true;
      }
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    ptx::mbar_wait(ptx::smem_u32(o_done), (nblk - 1) & 1);
    ptx::tc_fence_after();
    const float inv = (pending ? corr_pending : 1.f) / l;   // a correction decided after the last block still applies to O
    const int qrow = q0 + row;
    __nv_bfloat16* orow = out + ((long long)b * Nq + qrow) * ldo + h * D;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(t_o + c * 32, r);
      ptx::tmem_ld_wait();
      if (qrow < Nq) {
        uint4* o4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          o4[q] = make_uint4(pack2(__uint_as_float(r[8 * q]) * inv, __uint_as_float(r[8 * q + 1]) * inv),
                             pack2(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv),
                             pack2(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv),
                             pack2(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv));
      }
    }
  }
  ATT_STAMP(31);
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc

// q/k/v may be column slices of wider matrices (fused qkv / kv buffers): the map covers the whole matrix that
// starts at the 16-byte aligned `*_base` pointer, the head column offset is added in the kernel.
int attention_tc2_set_debug(void* dev_buf);
int attention_tc3_set_debug(void* dev_buf);
int attention_set_debug(void* dev_buf) {
  if (int rc = attention_tc2_set_debug(dev_buf)) return rc;
  if (int rc = attention_tc3_set_debug(dev_buf)) return rc;
  unsigned long long* p = reinterpret_cast<unsigned long long*>(dev_buf);
  D3R_CUDA(cudaMemcpyToSymbol(tc::g_attn_dbg, &p, sizeof(p)));
  return D3R_OK;
}

// debug: extra dynamic shared memory per CTA to lower the number of co-resident CTAs (occupancy scaling experiments)
static int g_occupancy_pad = 0;
void set_tc_occupancy_pad(int bytes) { g_occupancy_pad = bytes; }

int attention_hd64_tc(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                      long long ldo, int B, int heads, int Nq, int Nk, float scale, cudaStream_t st) {
  D3R_CHECK_ARG(q && k && v && out && B > 0 && heads > 0 && Nq > 0 && Nk > 0, "attention: bad arguments");
  D3R_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention: row strides must be multiples of 8");
  D3R_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "attention: pointers must be 16-byte aligned");
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = tc::make_map(&mq, q, ldq, heads * 64, Nq, B, tc::BQ))) return rc;
  if ((rc = tc::make_map(&mk, k, ldk, heads * 64, Nk, B, tc::BK))) return rc;
  if ((rc = tc::make_map(&mv, v, ldv, heads * 64, Nk, B, tc::BK))) return rc;
  static bool attr = false;
  if (!attr) {
    D3R_CUDA(cudaFuncSetAttribute(tc::attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    // two CTAs per SM only fit with the maximum shared-memory carve-out
    D3R_CUDA(cudaFuncSetAttribute(tc::attention_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  dim3 grid((Nq + tc::BQ - 1) / tc::BQ, heads, B);
  prof::Scope scope("attention_tcgen05", st, 4.0 * double(B) * heads * double(Nq) * double(Nk) * 64.0);
  D3R_CUDA(pdl::launch(tc::attention_tc_kernel, grid, dim3(tc::kThreads), size_t(tc::kSmemBytes + g_occupancy_pad), st, mq, mk, mv,
                       (__nv_bfloat16*)out, ldo, Nq, Nk, 0, 0, 0, scale * 1.4426950408889634f));
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}

}  // namespace attn
}  // namespace d3r
