// tcgen05 / TMEM fused attention for head dim 64 (sm_100a), split-row variant.
//
//   O = softmax(Q K^T * scale) V      per (image b, head h), bf16 in/out, fp32 softmax + accumulation
//
// (TMA producer warp, single-thread MMA issuer, S and O in TMEM, P through
// 128B-swizzled shared memory) with the softmax work of one 128-query tile spread over EIGHT warps instead of four:
// the two warps that own a TMEM lane quarter (32 query rows) each take half of the 128 keys of a block, so the
// per-block serial chain of a thread (TMEM load -> 64 exp2 -> P store -> fence -> barrier) is half as long for the
// same number of MMA / barrier round trips, and the MUFU pipe -- the unit that bounds hd-64 attention -- sees 16
// instead of 12 softmax warps per SM (2 CTAs x 320 threads).
//
// CTAs are PERSISTENT (2 per SM): each loops over (image, head, 128-query tile) work items, so barrier set-up, TMEM
// allocation and the first Q / K round trip are paid once per CTA instead of once per tile, and the next tile's Q
// and K/V loads are already in flight while the current tile's last block and epilogue finish.
//
//   warp 0      TMA producer : per tile Q, then K_j / V_j (128 keys x 64) into 2-deep rings
//   warp 1      MMA issuer   : S = Q K_j^T (UMMA 128x128x16 x4) -> TMEM cols [0,128)
//                              O += P V_j  (UMMA 128x64x16 x8, A = P from smem, B = V_j MN-major) -> cols [128,192)
//   warps 2..9  softmax      : warp w owns rows 32*(w%4).. and keys 64*half.. (half = (w-2)/4) of every block.
//                              The exponent reference m_ref of a row must be the same in both halves: the block
//                              maxima are exchanged through shared memory (bf16, one named barrier per lane quarter
//                              and block); both warps then take identical decisions.  O (TMEM) is rescaled lazily
//                              (only when a row max outgrew the reference by 2^8), each half scaling 32 of the 64
//                              columns; row sums are kept per half and added in the epilogue.
#include "d3r_common.cuh"
#include "sm100_ptx.cuh"
#include "elementwise.h"
#include "prof.h"
#include "attention_tc_common.cuh"
#include "pdl.cuh"
#include <type_traits>

namespace d3r {
namespace attn {

namespace tc2 {

using namespace tcc;

constexpr int BQ = 128, BK = 128, D = 64;
constexpr int kSoftmaxWarps = 8;
constexpr int kThreads = 64 + 32 * kSoftmaxWarps;
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB: 128 rows x 128 B (Q tile, one P atom, one K / V block)
constexpr int kKVBytes = BK * 64 * 2;
constexpr int kSmemBytes = kTileBytes /*Q*/ + 2 * kKVBytes /*K ring*/ + 2 * kKVBytes /*V ring*/ + 2 * kTileBytes /*P*/ +
                           128 /*barriers*/;
static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM must fit the 228 KB of shared memory");
constexpr int kTmemCols = 256;
constexpr uint32_t kColS = 0, kColO = BK;
// the two warps of a lane quarter exchange per-row scalars through spare TMEM columns of their own lanes:
// [kColX + 2*parity + half] block maxima (delayed exchange), [kColL + half] row sums, [kColM + half] first-block maxima
constexpr uint32_t kColX = BK + 64, kColL = kColX + 4, kColM = kColX + 6;

__device__ __forceinline__ void pair_barrier(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}
// optional cycle stamps of one CTA (debug aid): [4 key blocks][16]: slots 0..7 softmax thread (warp 2 lane 0), 8..15 MMA thread
__device__ unsigned long long* g_attn2_dbg = nullptr;
#define A2_STAMP(gg, slot) do { if (dbg && (gg) >= 6u && (gg) < 10u) dbg[((gg) - 6u) * 16 + (slot)] = (unsigned long long)clock64(); } while (0)

__device__ __forceinline__ void tmem_st_x1(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ float tmem_ld_x1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return __uint_as_float(v);
}
// 2^x on the FMA pipe (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], rel. error 7.5e-5, far below the bf16
// rounding of P): a fraction of the exponentials is computed this way so that the MUFU pipe, which bounds hd-64
// attention, is not the only unit producing them
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -126.f);
  const float t = x + 12582912.f;            // 1.5 * 2^23: the low mantissa bits now hold round(x)
  const float r = x - (t - 12582912.f);      // in [-0.5, 0.5]
  float p = fmaf(0.0551716648f, r, 0.2426111549f);
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// ABL: timing ablations (debug; results are wrong for ABL != 0): 1 = no exp2 (raw bits packed), 2 = also no P store,
// 3 = also no TMEM loads of S (pure barrier / MMA / TMA skeleton)
// POLY: every POLY-th pair of exponentials goes to the FMA pipe (0 = none)
template <int ABL, int POLY>
__global__ void __launch_bounds__(kThreads, 2)
attention_tc2_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, long long ldo, int Nq, int Nk,
                     int heads, int total_tiles, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzle atoms need 1024-byte alignment
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + kTileBytes;
  uint8_t* s_v = s_k + 2 * kKVBytes;
  uint8_t* s_p = s_v + 2 * kKVBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_p + 2 * kTileBytes);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_ready = bars + 9;   // S_j in TMEM
  uint64_t* s_free = bars + 10;   // softmax finished reading S_j
  uint64_t* p_ready = bars + 11;  // P_j in smem (+ O rescaled)
  uint64_t* o_done = bars + 12;   // P V_j accumulated
  uint64_t* q_empty = bars + 13;  // every S MMA of the tile has read Q
  uint64_t* o_free = bars + 14;   // the epilogue has read O out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq_tiles = (Nq + BQ - 1) / BQ;
  const int nblk = (Nk + BK - 1) / BK;
  unsigned long long* dbg = (g_attn2_dbg && blockIdx.x == 5 && (threadIdx.x == 64 || threadIdx.x == 32)) ? g_attn2_dbg : nullptr;
  // work item t -> (query tile, head, image); neighbouring CTAs work on the same (image, head) at the same time,
  // so its K / V are fetched from HBM once and then hit in L2
  auto tile_coords = [&](int t, int& q0, int& h, int& b) {
    q0 = (t % nq_tiles) * BQ;
    h = (t / nq_tiles) % heads;
    b = t / (nq_tiles * heads);
  };

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
    ptx::mbar_init(ptx::smem_u32(s_free), kSoftmaxWarps);
    ptx::mbar_init(ptx::smem_u32(p_ready), kSoftmaxWarps);
    ptx::mbar_init(ptx::smem_u32(o_done), 1);
    ptx::mbar_init(ptx::smem_u32(q_empty), 1);
    ptx::mbar_init(ptx::smem_u32(o_free), kSoftmaxWarps);
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
      uint32_t g = 0, it = 0;   // global key-block counter / tile counter of this CTA (barrier phases run on across tiles)
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it) {
        int q0, h, b;
        tile_coords(t, q0, h, b);
        ptx::mbar_wait(ptx::smem_u32(q_empty), (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(q_full), kTileBytes);
        ptx::tma_load_3d(ptx::smem_u32(s_q), &tmap_q, ptx::smem_u32(q_full), h * D, q0, b);
        for (int j = 0; j < nblk; ++j, ++g) {
          const int st = g & 1;
          const uint32_t ph = (g >> 1) & 1;
          ptx::mbar_wait(ptx::smem_u32(&k_empty[st]), ph ^ 1);
          ptx::mbar_arrive_expect_tx(ptx::smem_u32(&k_full[st]), kKVBytes);
          ptx::tma_load_3d(ptx::smem_u32(s_k + st * kKVBytes), &tmap_k, ptx::smem_u32(&k_full[st]), h * D, j * BK, b);
          ptx::mbar_wait(ptx::smem_u32(&v_empty[st]), ph ^ 1);
          ptx::mbar_arrive_expect_tx(ptx::smem_u32(&v_full[st]), kKVBytes);
          ptx::tma_load_3d(ptx::smem_u32(s_v + st * kKVBytes), &tmap_v, ptx::smem_u32(&v_full[st]), h * D, j * BK, b);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, BK, 0, 0);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(128, 64, 0, 1);  // B (=V) is MN-major
      const uint32_t d_s = tmem_base + kColS, d_o = tmem_base + kColO;
      const uint64_t dq = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(s_q));
      auto issue_s = [&](uint32_t g) {   // g: global key-block index of this CTA
        const int st = g & 1;
        A2_STAMP(g, 8);
        ptx::mbar_wait(ptx::smem_u32(&k_full[st]), (g >> 1) & 1);
        A2_STAMP(g, 9);
        if (g > 0) ptx::mbar_wait(ptx::smem_u32(s_free), (g - 1) & 1);  // softmax has drained the previous S
        A2_STAMP(g, 10);
        ptx::tc_fence_after();
        const uint64_t dk = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(s_k + st * kKVBytes));
#pragma unroll
        for (int k = 0; k < D / 16; ++k) ptx::umma_bf16_ss(d_s, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k ? 1u : 0u);
        ptx::umma_commit(ptx::smem_u32(&k_empty[st]));
        ptx::umma_commit(ptx::smem_u32(s_ready));
        A2_STAMP(g, 11);
      };
      uint32_t g0 = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it, g0 += nblk) {
        ptx::mbar_wait(ptx::smem_u32(q_full), it & 1);
        issue_s(g0);
        if (nblk == 1) ptx::umma_commit(ptx::smem_u32(q_empty));
        for (int j = 0; j < nblk; ++j) {
          const uint32_t g = g0 + j;
          if (j + 1 < nblk) {
            issue_s(g + 1);
            if (j + 2 == nblk) ptx::umma_commit(ptx::smem_u32(q_empty));   // last S of the tile: Q may be replaced
          }
          const int st = g & 1;
          A2_STAMP(g, 12);
          ptx::mbar_wait(ptx::smem_u32(p_ready), g & 1);
          A2_STAMP(g, 13);
          ptx::mbar_wait(ptx::smem_u32(&v_full[st]), (g >> 1) & 1);
          A2_STAMP(g, 14);
          if (j == 0 && it > 0) ptx::mbar_wait(ptx::smem_u32(o_free), (it - 1) & 1);   // previous tile's O has been read
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
          A2_STAMP(g, 15);
        }
      }
    }
  } else {
    // ================= softmax / correction / epilogue warps =================
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;        // which 64 keys of a block / which 32 columns of O
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + kColS + half * 64, t_o = tmem_base + lane_addr + kColO + half * 32;
    const uint32_t t_x = tmem_base + lane_addr + kColX, t_l = tmem_base + lane_addr + kColL, t_m = tmem_base + lane_addr + kColM;
    // m_ref: the exponent reference baked into P, O and l (identical in both halves of a row).  Blocks j >= 1
    // exponentiate against the reference they inherit; it moves only when a row max outgrew it by more than 2^kLazy.
    constexpr float kLazy = 8.f;
    uint32_t g0 = 0, it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it, g0 += nblk) {
    int q0, h, b;
    tile_coords(t, q0, h, b);
    float m_ref = -INFINITY, l = 0.f, corr_pending = 1.f, own_prev = -INFINITY;
    bool pending = false;
    for (int j = 0; j < nblk; ++j) {
      const uint32_t g = g0 + j;
      const int nvalid = min(BK, Nk - j * BK) - half * 64;   // valid keys among this warp's 64 (may be <= 0)
      A2_STAMP(g, 0);
      ptx::mbar_wait(ptx::smem_u32(s_ready), g & 1);
      ptx::tc_fence_after();
      A2_STAMP(g, 1);
      float rs0 = 0.f, rs1 = 0.f, bm0 = -INFINITY, bm1 = -INFINITY;
      uint32_t packed[32];
      float ms;
      auto exp_chunk_t = [&](const uint32_t* r, int c, auto masked) {
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
          if constexpr (decltype(masked)::value) {   // masked keys contribute p = 0
            if (c * 32 + i >= nvalid) s0 = -INFINITY;
            if (c * 32 + i + 1 >= nvalid) s1 = -INFINITY;
          }
          const bool on_fma = POLY > 0 && !decltype(masked)::value && ((i >> 1) % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1;
          const float x0 = fmaf(s0, scale_log2, -ms), x1 = fmaf(s1, scale_log2, -ms);
          const float p0 = (ABL >= 1) ? x0 : (on_fma ? poly_exp2(x0) : fast_exp2(x0));
          const float p1 = (ABL >= 1) ? x1 : (on_fma ? poly_exp2(x1) : fast_exp2(x1));
          rs0 += p0;
          rs1 += p1;
          bm0 = fmaxf(bm0, s0);
          bm1 = fmaxf(bm1, s1);
          packed[c * 16 + (i >> 1)] = pack2(p0, p1);
        }
      };
      const bool ragged = nvalid < 64;   // warp-uniform
      auto exp_chunk = [&](const uint32_t* r, int c) {
        if (ragged) exp_chunk_t(r, c, std::true_type{});
        else exp_chunk_t(r, c, std::false_type{});
      };
      if (j == 0) {
        // first block: the reference is the row max over BOTH halves, so the whole row is needed before any exp
        uint32_t sr[64];
        ptx::tmem_ld_32x32b_x32(t_s, sr);
        ptx::tmem_ld_32x32b_x32(t_s + 32, sr + 32);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_relaxed(ptx::smem_u32(s_free));   // publishes only completed TMEM loads
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 64; ++i)
          if (i < nvalid) mx = fmaxf(mx, __uint_as_float(sr[i]));
        tmem_st_x1(t_m + half, mx);
        tmem_st_wait();
        ptx::tc_fence_before();
        pair_barrier(quarter);
        ptx::tc_fence_after();
        m_ref = fmaxf(mx, tmem_ld_x1(t_m + (half ^ 1)));
        ptx::tmem_ld_wait();
        if (m_ref == -INFINITY) m_ref = 0.f;   // cannot happen for Nk >= 1; keeps exp2(-inf - -inf) out of the row
        ms = m_ref * scale_log2;
        exp_chunk(sr, 0);
        exp_chunk(sr + 32, 1);
      } else {
        ms = m_ref * scale_log2;
        uint32_t rbuf[2][32];
        if (ABL >= 3) {
#pragma unroll
          for (int i = 0; i < 32; ++i) { rbuf[0][i] = __float_as_uint(float(i) * 0.01f); rbuf[1][i] = rbuf[0][i]; }
        } else {
          ptx::tmem_ld_32x32b_x32(t_s, rbuf[0]);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(t_s + 32, rbuf[1]);
        }
        exp_chunk(rbuf[0], 0);
        if (ABL < 3) ptx::tmem_ld_wait();
        // S_j fully read (by this warp) -> after all eight arrivals the MMA warp may overwrite it with S_{j+1}
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_relaxed(ptx::smem_u32(s_free));   // publishes only completed TMEM loads
        exp_chunk(rbuf[1], 1);
      }
      A2_STAMP(g, 2);
      // the previous P V must have consumed P (smem) and updated O before we touch either (the last P V of the
      // previous tile was waited for in that tile's epilogue)
      if (j > 0) {
        ptx::mbar_wait(ptx::smem_u32(o_done), (g - 1) & 1);
        ptx::tc_fence_after();
      }
      A2_STAMP(g, 3);
      // the partner's maximum of the PREVIOUS block (published before its p_ready arrival, hence before the o_done
      // just waited for) -- read asynchronously under the P store
      float peer_prev = -INFINITY;
      if (j > 0) peer_prev = tmem_ld_x1(t_x + 2 * ((j - 1) & 1) + (half ^ 1));
      if (pending) {   // warp-uniform, and the same decision in the partner warp; decided one block ago
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(t_o, r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr_pending);
        tmem_st_32x32b_x32(t_o, r);
        tmem_st_wait();
        l *= corr_pending;
        pending = false;
      }
      // swizzled store of P into this half's atom (K-major SW128: 16-byte chunk index ^ (row & 7))
      if (ABL < 2) {
        uint8_t* atom = s_p + half * kTileBytes + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(atom + ((q ^ (row & 7)) << 4)) =
              make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
      }
      A2_STAMP(g, 4);
      l += rs0 + rs1;
      if (ABL >= 2) l += __uint_as_float(packed[lane & 31]) * 1e-30f;   // keep the packed values alive
      // reference update, one block late and without a rendezvous: both halves look at the same two numbers (the
      // two block maxima of block j-1) and therefore move the reference identically.  It takes effect for the
      // exponentials of block j+1; O and l are rescaled there, after P V_j (computed against the old reference)
      // has been accumulated.  In between p <= 2^(kLazy + growth of two blocks): harmless in fp32 / bf16.
      if (j > 0) {
        ptx::tmem_ld_wait();
        const float rowmax = fmaxf(own_prev, peer_prev);
        if (__any_sync(0xffffffffu, (rowmax - m_ref) * scale_log2 > kLazy)) {
          const float m_new = fmaxf(m_ref, rowmax);
          corr_pending = fast_exp2((m_ref - m_new) * scale_log2);
          m_ref = m_new;
          pending = true;
        }
      }
      own_prev = fmaxf(bm0, bm1);
      tmem_st_x1(t_x + 2 * (j & 1) + half, own_prev);
      tmem_st_wait();
      A2_STAMP(g, 5);
      // make the generic-proxy smem writes of P visible to the tensor core (async proxy), then publish
      ptx::fence_proxy_async();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(p_ready));
      A2_STAMP(g, 6);
    }
    // ---- epilogue: O / l -> bf16 -> global (each half: 32 of the 64 columns) ----
    ptx::mbar_wait(ptx::smem_u32(o_done), (g0 + nblk - 1) & 1);
    ptx::tc_fence_after();
    uint32_t r[32];
    ptx::tmem_ld_32x32b_x32(t_o, r);
    // O, l and (if a reference move is still pending) both halves share the same reference: no correction needed
    tmem_st_x1(t_l + half, l);
    tmem_st_wait();
    ptx::tc_fence_before();
    pair_barrier(quarter);
    ptx::tc_fence_after();
    const float l_peer = tmem_ld_x1(t_l + (half ^ 1));
    ptx::tmem_ld_wait();
    const float inv = 1.f / (l + l_peer);
    // O is in registers: the MMA warp may start the next tile's P V
    ptx::tc_fence_before();
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive_relaxed(ptx::smem_u32(o_free));
    const int qrow = q0 + row;
    if (qrow < Nq) {
      uint4* o4 = reinterpret_cast<uint4*>(out + ((long long)b * Nq + qrow) * ldo + h * D + half * 32);
#pragma unroll
      for (int q = 0; q < 4; ++q)
        o4[q] = make_uint4(pack2(__uint_as_float(r[8 * q]) * inv, __uint_as_float(r[8 * q + 1]) * inv),
                           pack2(__uint_as_float(r[8 * q + 2]) * inv, __uint_as_float(r[8 * q + 3]) * inv),
                           pack2(__uint_as_float(r[8 * q + 4]) * inv, __uint_as_float(r[8 * q + 5]) * inv),
                           pack2(__uint_as_float(r[8 * q + 6]) * inv, __uint_as_float(r[8 * q + 7]) * inv));
    }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc2

int attention_tc2_set_debug(void* dev_buf) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(dev_buf);
  D3R_CUDA(cudaMemcpyToSymbol(tc2::g_attn2_dbg, &p, sizeof(p)));
  return D3R_OK;
}

constexpr int kDefaultPoly = 0;   // measured: any share of FMA-pipe exponentials is slower here (issue slots, not the MUFU, run out first)
static int g_ablation = 0;
void set_tc2_ablation(int a) { g_ablation = a; }

template <int ABL, int POLY>
static int launch_tc2(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, void* out, long long ldo, int B, int heads,
                      int Nq, int Nk, float scale, cudaStream_t st);

int attention_hd64_tc2(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                       long long ldo, int B, int heads, int Nq, int Nk, float scale, cudaStream_t st) {
  D3R_CHECK_ARG(q && k && v && out && B > 0 && heads > 0 && Nq > 0 && Nk > 0, "attention: bad arguments");
  D3R_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention: row strides must be multiples of 8");
  D3R_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "attention: pointers must be 16-byte aligned");
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = tcc::make_map(&mq, q, ldq, heads * 64, Nq, B, tc2::BQ))) return rc;
  if ((rc = tcc::make_map(&mk, k, ldk, heads * 64, Nk, B, tc2::BK))) return rc;
  if ((rc = tcc::make_map(&mv, v, ldv, heads * 64, Nk, B, tc2::BK))) return rc;
  switch (g_ablation) {   // debug encodings: 1..3 timing ablations; 4..7 share of exponentials on the FMA pipe
    case 1: return launch_tc2<1, 0>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 2: return launch_tc2<2, 0>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 3: return launch_tc2<3, 0>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 4: return launch_tc2<0, 0>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 5: return launch_tc2<0, 8>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 6: return launch_tc2<0, 3>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 7: return launch_tc2<0, 2>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    default: return launch_tc2<0, kDefaultPoly>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
  }
}

template <int ABL, int POLY>
static int launch_tc2(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, void* out, long long ldo, int B, int heads,
                      int Nq, int Nk, float scale, cudaStream_t st) {
  static unsigned long long attr_devices = 0;
  if (first_launch_on_this_device(attr_devices)) {
    D3R_CUDA(cudaFuncSetAttribute(tc2::attention_tc2_kernel<ABL, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::kSmemBytes));
    // two CTAs per SM only fit with the maximum shared-memory carve-out
    D3R_CUDA(cudaFuncSetAttribute(tc2::attention_tc2_kernel<ABL, POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  const int total_tiles = ((Nq + tc2::BQ - 1) / tc2::BQ) * heads * B;
  const int slots = 2 * num_sms();   // two persistent CTAs per SM
  // equal number of tiles per CTA where possible: a grid of `slots` CTAs would leave a ragged last round
  const int rounds = (total_tiles + slots - 1) / slots;
  const int grid = (total_tiles + rounds - 1) / rounds;
  prof::Scope scope("attention_tcgen05_split", st, 4.0 * double(B) * heads * double(Nq) * double(Nk) * 64.0);
  D3R_CUDA(pdl::launch(tc2::attention_tc2_kernel<ABL, POLY>, dim3(grid), dim3(tc2::kThreads), size_t(tc2::kSmemBytes), st, mq, mk, mv,
                       (__nv_bfloat16*)out, ldo, Nq, Nk, heads, total_tiles, scale * 1.4426950408889634f));
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}

}  // namespace attn
}  // namespace d3r
