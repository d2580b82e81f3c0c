// tcgen05 / TMEM fused attention for head dim 64 (sm_100a): split-row softmax, P kept in TENSOR MEMORY.
//
//   O = softmax(Q K^T * scale) V      per (image b, head h), bf16 in/out, fp32 accumulation
//   reference: croco/models/blocks.py:94-112 (Attention), 147-169 (CrossAttention)
//
// Same CTA structure as attention_tc2.cu (persistent CTAs, 2 per SM; TMA producer warp, single-thread MMA issuer, eight
// softmax warps: the two warps of a TMEM lane quarter split the 128 keys of a block), with the two costs that bounded it
// removed:
//   * P never touches shared memory.  The softmax warps write bf16 P_j straight into 64 TMEM columns (tcgen05.st) and
//     O += P_j V_j is issued in the A-from-TMEM form of tcgen05.mma: no 32 KB swizzled store + fence.proxy.async per block,
//     and the PV MMA reads 16 KB of shared memory (V) instead of 48 KB.  TMEM: S [0,128) | P [128,192) | O [192,256).
//   * the per-key instruction count is cut from ~4.5 to 3: the exponent argument of a key PAIR comes from one packed FFMA2,
//     the row sum from one packed FADD2, the running maximum from one packed HMNMX2 on the bf16 pair that goes to TMEM;
//     TMEM reads of S are software-pipelined against the exponentials, 16 keys at a time; every 4th key pair takes its two
//     exponentials on the FMA pipe (Cody-Waite + degree-3 polynomial, rel. error 7.5e-5), because the MUFU is the busiest unit
//     of the kernel (54 % active against 42 % of the issue slots).  (ex2.approx.ftz.bf16x2 was
//     measured: same MUFU time per key as fp32 -- 16.5 / clk / SM either way -- and 2-3 % error on the dominant keys once
//     the lazy reference lets the exponent grow to +8; the exponentials stay fp32.)
// Softmax bookkeeping lives in the log2 domain: t = s * scale * log2(e), reference `ms`, x = t - ms, p = 2^x.  The per-row
// scalars the two halves exchange (first-block maximum, block maxima one block late, row sums) go through 3 KB of shared
// memory, ordered by the same barriers as in attention_tc2.cu; the freed shared memory makes the K ring three deep.
#include "d3r_common.cuh"
#include "sm100_ptx.cuh"
#include "elementwise.h"
#include "prof.h"
#include "attention_tc_common.cuh"
#include "pdl.cuh"
#include <type_traits>

namespace d3r {
namespace attn {

namespace tc3 {

using namespace tcc;

constexpr int BQ = 128, BK = 128, D = 64;
constexpr int kSoftmaxWarps = 8;
constexpr int kThreads = 64 + 32 * kSoftmaxWarps;
constexpr int kTileBytes = 128 * 64 * 2;     // 16 KB: 128 rows x 128 B (Q tile, one K / V block)
constexpr int kKVBytes = BK * 64 * 2;
constexpr int kKStages = 3, kVStages = 2;
constexpr int kXchFloats = 3 * 2 * 128;      // [slot][half][row]: slots 0/1 block maxima (by block parity), slot 2 first max / row sum
constexpr int kSmemBytes = kTileBytes /*Q*/ + kKStages * kKVBytes + kVStages * kKVBytes + kXchFloats * 4 + 256 /*barriers*/;
static_assert(2 * (kSmemBytes + 1024) <= 233472, "two CTAs per SM must fit the 228 KB of shared memory");
constexpr int kTmemCols = 256;
constexpr uint32_t kColS = 0, kColP = BK, kColO = BK + 64;

typedef unsigned long long f2;               // two packed fp32
__device__ __forceinline__ f2 pk2(float lo, float hi) { f2 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d; }
__device__ __forceinline__ void upk2(f2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f2 fadd2(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 fsub2(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 ffma2(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
// {hi, lo} fp32 -> packed bf16x2 (lo in the low half: the even key of a pair)
__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) {
  uint32_t d;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t d; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(d) : "r"(x)); return d; }
__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) { uint32_t d; asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t a, uint32_t b) { uint32_t d; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ float bf_lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf_hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void pair_barrier(int quarter) {
  asm volatile("bar.sync %0, 64;" ::"r"(quarter + 1) : "memory");
}

// optional cycle stamps of one CTA (debug aid): [4 key blocks][16]: slots 0..7 softmax thread (warp 2 lane 0), 8..15 MMA thread
__device__ unsigned long long* g_attn3_dbg = nullptr;
#define A3_STAMP(gg, slot) do { if (dbg && (gg) >= 10u && (gg) < 14u) dbg[((gg) - 10u) * 16 + (slot)] = (unsigned long long)clock64(); } while (0)

// 2^x for a key pair on the FMA pipe (Cody-Waite split + degree-3 minimax on [-0.5, 0.5], rel. error 7.5e-5 -- far below the bf16
// rounding of P): offloads a share of the exponentials from the MUFU, the busiest unit of this kernel (54 % vs 42 % issue slots)
__device__ __forceinline__ f2 poly_exp2_pair(f2 x) {
  float x0, x1;
  upk2(x, x0, x1);
  x = pk2(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const f2 magic = pk2(12582912.f, 12582912.f);           // 1.5 * 2^23: the low mantissa bits of t hold round(x)
  const f2 t = fadd2(x, magic);
  const f2 r = fadd2(x, fsub2(magic, t));                   // x - (t - magic) in [-0.5, 0.5]
  f2 p = ffma2(pk2(0.0551716648f, 0.0551716648f), r, pk2(0.2426111549f, 0.2426111549f));
  p = ffma2(p, r, pk2(0.6932609677f, 0.6932609677f));
  p = ffma2(p, r, pk2(0.9999280572f, 0.9999280572f));
  float p0, p1, t0, t1;
  upk2(p, p0, p1);
  upk2(t, t0, t1);
  return pk2(__int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23)), __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23)));
}

template <int ABL, int POLY>   // ABL: timing ablation (debug; results are wrong for ABL != 0): 1 = no exponentials; POLY: every POLY-th pair on the FMA pipe (0 = none)
__global__ void __launch_bounds__(kThreads, 2)
attention_tc3_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                     const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, long long ldo, int Nq, int Nk,
                     int heads, int total_tiles, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t smem[];   // 128B-swizzle atoms need 1024-byte alignment
  if ((ptx::smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* s_q = smem;
  uint8_t* s_k = s_q + kTileBytes;
  uint8_t* s_v = s_k + kKStages * kKVBytes;
  float* s_x = reinterpret_cast<float*>(s_v + kVStages * kKVBytes);   // [3][2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_x + kXchFloats);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [3]
  uint64_t* k_empty = bars + 4;    // [3]
  uint64_t* v_full = bars + 7;     // [2]
  uint64_t* v_empty = bars + 9;    // [2]
  uint64_t* s_ready = bars + 11;   // S_j in TMEM
  uint64_t* s_free = bars + 12;    // softmax finished reading S_j
  uint64_t* p_ready = bars + 13;   // P_j in TMEM (+ O rescaled)
  uint64_t* o_done = bars + 14;    // P V_j accumulated
  uint64_t* q_empty = bars + 15;   // every S MMA of the tile has read Q
  uint64_t* o_free = bars + 16;    // the epilogue has read O out of TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nq_tiles = (Nq + BQ - 1) / BQ;
  const int nblk = (Nk + BK - 1) / BK;
  unsigned long long* dbg = (g_attn3_dbg && blockIdx.x == 5 && (threadIdx.x == 64 || threadIdx.x == 32)) ? g_attn3_dbg : nullptr;
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
    for (int s = 0; s < kKStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&k_full[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&k_empty[s]), 1);
    }
    for (int s = 0; s < kVStages; ++s) {
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
        ptx::mbar_wait_hint(ptx::smem_u32(q_empty), (it & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(ptx::smem_u32(q_full), kTileBytes);
        ptx::tma_load_3d(ptx::smem_u32(s_q), &tmap_q, ptx::smem_u32(q_full), h * D, q0, b);
        for (int j = 0; j < nblk; ++j, ++g) {
          const uint32_t ks = g % kKStages, kph = (g / kKStages) & 1;
          const uint32_t vs = g % kVStages, vph = (g / kVStages) & 1;
          ptx::mbar_wait_hint(ptx::smem_u32(&k_empty[ks]), kph ^ 1);
          ptx::mbar_arrive_expect_tx(ptx::smem_u32(&k_full[ks]), kKVBytes);
          ptx::tma_load_3d(ptx::smem_u32(s_k + ks * kKVBytes), &tmap_k, ptx::smem_u32(&k_full[ks]), h * D, j * BK, b);
          ptx::mbar_wait_hint(ptx::smem_u32(&v_empty[vs]), vph ^ 1);
          ptx::mbar_arrive_expect_tx(ptx::smem_u32(&v_full[vs]), kKVBytes);
          ptx::tma_load_3d(ptx::smem_u32(s_v + vs * kKVBytes), &tmap_v, ptx::smem_u32(&v_full[vs]), h * D, j * BK, b);
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc_s = ptx::umma_idesc_bf16(128, BK, 0, 0);
      constexpr uint32_t idesc_o = ptx::umma_idesc_bf16(128, 64, 0, 1);  // A = P from TMEM (K-major), B = V MN-major
      const uint32_t d_s = tmem_base + kColS, d_o = tmem_base + kColO, a_p = tmem_base + kColP;
      const uint64_t dq = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(s_q));
      auto issue_s = [&](uint32_t g) {   // g: global key-block index of this CTA
        const uint32_t ks = g % kKStages;
        A3_STAMP(g, 8);
        ptx::mbar_wait_hint(ptx::smem_u32(&k_full[ks]), (g / kKStages) & 1);
        A3_STAMP(g, 9);
        if (g > 0) ptx::mbar_wait_hint(ptx::smem_u32(s_free), (g - 1) & 1);  // softmax has drained the previous S
        A3_STAMP(g, 10);
        ptx::tc_fence_after();
        const uint64_t dk = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(s_k + ks * kKVBytes));
#pragma unroll
        for (int k = 0; k < D / 16; ++k) ptx::umma_bf16_ss(d_s, dq + uint64_t(2 * k), dk + uint64_t(2 * k), idesc_s, k ? 1u : 0u);
        ptx::umma_commit(ptx::smem_u32(&k_empty[ks]));
        ptx::umma_commit(ptx::smem_u32(s_ready));
        A3_STAMP(g, 11);
      };
      uint32_t g0 = 0, it = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it, g0 += nblk) {
        ptx::mbar_wait_hint(ptx::smem_u32(q_full), it & 1);
        issue_s(g0);
        if (nblk == 1) ptx::umma_commit(ptx::smem_u32(q_empty));
        for (int j = 0; j < nblk; ++j) {
          const uint32_t g = g0 + j;
          if (j + 1 < nblk) {
            issue_s(g + 1);
            if (j + 2 == nblk) ptx::umma_commit(ptx::smem_u32(q_empty));   // last S of the tile: Q may be replaced
          }
          const uint32_t vs = g % kVStages;
          A3_STAMP(g, 12);
          ptx::mbar_wait_hint(ptx::smem_u32(p_ready), g & 1);
          A3_STAMP(g, 13);
          ptx::mbar_wait_hint(ptx::smem_u32(&v_full[vs]), (g / kVStages) & 1);
          A3_STAMP(g, 14);
          if (j == 0 && it > 0) ptx::mbar_wait_hint(ptx::smem_u32(o_free), (it - 1) & 1);   // previous tile's O has been read
          ptx::tc_fence_after();
          const uint32_t pv = ptx::smem_u32(s_v + vs * kKVBytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // A: rows = TMEM lanes, 16 keys = 8 columns of packed bf16 pairs; B: 16 keys x 64 dims of V, MN-major
            const uint64_t dv = umma_desc_mnmajor_sw128(pv + k * 16 * 128);
            ptx::umma_bf16_ts(d_o, a_p + uint32_t(8 * k), dv, idesc_o, (j | k) ? 1u : 0u);
          }
          ptx::umma_commit(ptx::smem_u32(&v_empty[vs]));
          ptx::umma_commit(ptx::smem_u32(o_done));
          A3_STAMP(g, 15);
        }
      }
    }
  } else {
    // ================= softmax / correction / epilogue warps =================
    const int quarter = warp & 3;            // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;        // which 64 keys of a block / which 32 columns of O
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = uint32_t(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_addr + kColS + half * 64, t_p = tmem_base + lane_addr + kColP + half * 32,
                   t_o = tmem_base + lane_addr + kColO + half * 32;
    float* x_own = s_x + half * 128 + row;           // + slot * 256
    float* x_peer = s_x + (half ^ 1) * 128 + row;
    // ms: the exponent reference (log2 domain) baked into P, O and l, identical in both halves of a row.  Blocks j >= 1
    // exponentiate against the reference they inherit; it moves only when a row max outgrew it by more than kLazy.
    constexpr float kLazy = 8.f;
    uint32_t g0 = 0, it = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++it, g0 += nblk) {
      int q0, h, b;
      tile_coords(t, q0, h, b);
      float ms = 0.f, l = 0.f, corr_pending = 1.f, own_prev = -INFINITY;
      bool pending = false;
      for (int j = 0; j < nblk; ++j) {
        const uint32_t g = g0 + j;
        const int nvalid = min(BK, Nk - j * BK) - half * 64;   // valid keys among this warp's 64 (may be <= 0)
        A3_STAMP(g, 0);
        ptx::mbar_wait(ptx::smem_u32(s_ready), g & 1);
        ptx::tc_fence_after();
        A3_STAMP(g, 1);
        uint32_t packed[32];
        uint32_t pmax = 0u;            // (+0, +0): running maximum of the bf16 probabilities (positive bf16 order like floats)
        f2 lsum2 = 0ull;               // packed row-sum accumulator (even keys | odd keys)
        const bool ragged = nvalid < 64;   // warp-uniform
        float ms_blk;
        // p = 2^(s * scale_log2 - ms): exponent pairs from one packed FFMA2, fp32 MUFU exponentials (a bf16 exponent would
        // cost the dominant keys 2-3 % once the lazy reference lets x grow to +8), packed row sums, one pack to bf16
        auto exp_chunk_t = [&](const uint32_t* r, int c0, int n, float ms_, auto masked) {   // n keys from key c0 of this warp's 64
          const f2 sc2 = pk2(scale_log2, scale_log2), nms2 = pk2(-ms_, -ms_);
#pragma unroll
          for (int i = 0; i < n / 2; ++i) {
            const f2 x = ffma2(pk2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])), sc2, nms2);
            float x0, x1;
            upk2(x, x0, x1);
            if constexpr (decltype(masked)::value) {   // masked keys contribute p = 0
              if (c0 + 2 * i >= nvalid) x0 = -INFINITY;
              if (c0 + 2 * i + 1 >= nvalid) x1 = -INFINITY;
            }
            float p0, p1;
            if (POLY > 0 && !decltype(masked)::value && (i % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
              upk2(poly_exp2_pair(x), p0, p1);
            } else {
              p0 = (ABL >= 1) ? x0 : fast_exp2(x0);
              p1 = (ABL >= 1) ? x1 : fast_exp2(x1);
            }
            lsum2 = fadd2(lsum2, pk2(p0, p1));
            const uint32_t pb = cvt_bf16x2(p1, p0);
            pmax = max_bf16x2(pmax, pb);
            packed[c0 / 2 + i] = pb;
          }
        };
        // the ragged (last, partially filled) key block takes its own copy of the loop: folded into one, the compiler
        // predicates every key of every block (ISETP + FSEL per key: 23 % of all instructions executed)
        auto exp_chunk = [&](const uint32_t* r, int c0, int n, float ms_) {
          if (ragged) exp_chunk_t(r, c0, n, ms_, std::true_type{});
          else exp_chunk_t(r, c0, n, ms_, std::false_type{});
        };
        // TMEM reads and exponentials (MUFU) are the two long poles of a block: software pipeline them inside the warp, 16
        // keys at a time (tcgen05.ld is asynchronous until tcgen05.wait::ld)
        uint32_t rb[2][16];
        tmem_ld_32x32b_x16(t_s, rb[0]);
        ptx::tmem_ld_wait();
        if (j == 0) {
          // first block: both halves of a row must agree on a reference before any exponential.  The maximum over each
          // half's FIRST 16 keys is enough: whatever the remaining keys add is handled like growth in later blocks (lazy
          // reference move), and the whole block can run through the same pipelined loop as the others.
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (i < nvalid) mx = fmaxf(mx, __uint_as_float(rb[0][i]));
          x_own[2 * 256] = mx;
          pair_barrier(quarter);
          mx = fmaxf(mx, x_peer[2 * 256]);
          ms = (mx == -INFINITY) ? 0.f : mx * scale_log2;   // -inf: no valid key among the 2 x 16 (Nk < 16): any reference works
        }
        ms_blk = ms;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (c < 3) tmem_ld_32x32b_x16(t_s + 16 * (c + 1), rb[(c + 1) & 1]);
          exp_chunk(rb[c & 1], 16 * c, 16, ms_blk);
          if (c < 3) ptx::tmem_ld_wait();
        }
        // S_j fully read (by this warp) -> after all eight arrivals the MMA warp may overwrite it with S_{j+1}
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_relaxed(ptx::smem_u32(s_free));   // publishes only completed TMEM loads
        A3_STAMP(g, 2);
        float rs0, rs1;
        upk2(lsum2, rs0, rs1);
        const float rs = rs0 + rs1;
        // the previous P V must have consumed P_{j-1} and updated O before we touch either (the last P V of the
        // previous tile was waited for in that tile's epilogue)
        if (j > 0) {
          ptx::mbar_wait(ptx::smem_u32(o_done), (g - 1) & 1);
          ptx::tc_fence_after();
        }
        A3_STAMP(g, 3);
        // the partner's maximum of the PREVIOUS block (stored before its p_ready arrival, hence before the o_done
        // just waited for)
        float peer_prev = -INFINITY;
        if (j > 0) peer_prev = x_peer[((j - 1) & 1) * 256];
        if (pending) {   // warp-uniform, and the same decision in the partner warp; decided one block ago
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(t_o, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr_pending);
          tmem_st_32x32b_x32(t_o, r);
          l *= corr_pending;
          pending = false;
        }
        tmem_st_32x32b_x32(t_p, packed);   // 64 keys of this row = 32 columns of packed pairs
        l += rs;
        // reference update, one block late and without a rendezvous: both halves look at the same two numbers (the two
        // block maxima of block j-1, absolute in the log2 domain) and therefore move the reference identically.  It
        // takes effect for the exponentials of block j+1; O and l are rescaled there, after P V_j (computed against the
        // old reference) has been accumulated.  In between p <= 2^(kLazy + growth of two blocks): harmless in fp32 / bf16.
        if (j > 0) {
          const float rowmax = fmaxf(own_prev, peer_prev);
          if (__any_sync(0xffffffffu, rowmax - ms > kLazy)) {
            const float ms_new = fmaxf(ms, rowmax);
            corr_pending = fast_exp2(ms - ms_new);
            // the exponentials of THIS block were already taken against the old reference: switch afterwards
            ms = ms_new;
            pending = true;
          }
        }
        own_prev = __log2f(fmaxf(bf_lo(pmax), bf_hi(pmax))) + ms_blk;   // absolute (log2 domain) block maximum; -inf if masked
        x_own[(j & 1) * 256] = own_prev;
        A3_STAMP(g, 4);
        tmem_st_wait();
        A3_STAMP(g, 5);
        // P_j (and a rescaled O) are in TMEM: publish (the release also covers the shared-memory store above)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(p_ready));
        A3_STAMP(g, 6);
      }
      // ---- epilogue: O / l -> bf16 -> global (each half: 32 of the 64 columns) ----
      ptx::mbar_wait(ptx::smem_u32(o_done), (g0 + nblk - 1) & 1);
      ptx::tc_fence_after();
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(t_o, r);
      // O, l and (if a reference move is still pending) both halves share the same reference: no correction needed
      x_own[2 * 256] = l;
      ptx::tmem_ld_wait();
      pair_barrier(quarter);
      const float inv = 1.f / (l + x_peer[2 * 256]);
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
      // the slot-2 exchange cell is reused by the next tile's first-block maximum: both halves must have read it
      pair_barrier(quarter);
      A3_STAMP(g0 + nblk - 1, 7);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace tc3

int attention_tc3_set_debug(void* dev_buf) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(dev_buf);
  D3R_CUDA(cudaMemcpyToSymbol(tc3::g_attn3_dbg, &p, sizeof(p)));
  return D3R_OK;
}

constexpr int kDefaultPoly = 4;   // measured (64,16,768,768): none 0.272 ms, every 4th pair 0.243, every 3rd 0.245, every 2nd 0.261
static int g_tc3_ablation = 0;
void set_tc3_ablation(int a) { g_tc3_ablation = a; }

template <int ABL, int POLY>
static int launch_tc3(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, void* out, long long ldo, int B, int heads,
                      int Nq, int Nk, float scale, cudaStream_t st) {
  static unsigned long long attr_devices = 0;
  if (first_launch_on_this_device(attr_devices)) {
    D3R_CUDA(cudaFuncSetAttribute(tc3::attention_tc3_kernel<ABL, POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::kSmemBytes));
    // two CTAs per SM only fit with the maximum shared-memory carve-out
    D3R_CUDA(cudaFuncSetAttribute(tc3::attention_tc3_kernel<ABL, POLY>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  const int total_tiles = ((Nq + tc3::BQ - 1) / tc3::BQ) * heads * B;
  const int slots = 2 * num_sms();   // two persistent CTAs per SM
  // equal number of tiles per CTA where possible: a grid of `slots` CTAs would leave a ragged last round
  const int rounds = (total_tiles + slots - 1) / slots;
  const int grid = (total_tiles + rounds - 1) / rounds;
  prof::Scope scope("attention_tcgen05_tmemP", st, 4.0 * double(B) * heads * double(Nq) * double(Nk) * 64.0);
  D3R_CUDA(pdl::launch(tc3::attention_tc3_kernel<ABL, POLY>, dim3(grid), dim3(tc3::kThreads), size_t(tc3::kSmemBytes), st, mq, mk, mv,
                       (__nv_bfloat16*)out, ldo, Nq, Nk, heads, total_tiles, scale * 1.4426950408889634f));
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}

int attention_hd64_tc3(const void* q, long long ldq, const void* k, long long ldk, const void* v, long long ldv, void* out,
                       long long ldo, int B, int heads, int Nq, int Nk, float scale, cudaStream_t st) {
  D3R_CHECK_ARG(q && k && v && out && B > 0 && heads > 0 && Nq > 0 && Nk > 0, "attention: bad arguments");
  D3R_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "attention: row strides must be multiples of 8");
  D3R_CHECK_ARG(((uintptr_t)q & 15) == 0 && ((uintptr_t)k & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "attention: pointers must be 16-byte aligned");
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = tcc::make_map(&mq, q, ldq, heads * 64, Nq, B, tc3::BQ))) return rc;
  if ((rc = tcc::make_map(&mk, k, ldk, heads * 64, Nk, B, tc3::BK))) return rc;
  if ((rc = tcc::make_map(&mv, v, ldv, heads * 64, Nk, B, tc3::BK))) return rc;
  switch (g_tc3_ablation) {   // debug encodings (d3r_set_attention_impl(3 + 10 k)): 1 = no exponentials; 2 / 3 / 4 = every 4th / 3rd / 2nd pair on the FMA pipe
    case 1: return launch_tc3<1, 0>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 2: return launch_tc3<0, 4>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 3: return launch_tc3<0, 3>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    case 4: return launch_tc3<0, 2>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
    default: return launch_tc3<0, kDefaultPoly>(mq, mk, mv, out, ldo, B, heads, Nq, Nk, scale, st);
  }
}

}  // namespace attn
}  // namespace d3r
