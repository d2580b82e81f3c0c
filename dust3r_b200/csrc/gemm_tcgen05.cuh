// Persistent warp-specialised bf16 GEMM / implicit-GEMM convolution on tcgen05 (sm_100a).
//
//   D[M,N] = A[M,K] * B[N,K]^T   (both operands K-major, fp32 accumulation in TMEM)
//
// Roles (320 threads, 1 CTA / SM, grid = min(#tiles, #SMs), static round-robin tile schedule with the
// N index fastest so the CTAs of one wave share A rows through L2):
//   warp 0      TMA producer   : cp.async.bulk.tensor -> 128B-swizzled smem ring (kStages deep)
//   warp 1      MMA issuer     : one elected lane issues tcgen05.mma (UMMA 128 x BLOCK_N x 16),
//                                tcgen05.commit releases smem slots / publishes the accumulator
//   warps 2..9  epilogue       : tcgen05.ld (32 lanes x 32 columns per instruction) -> fused epilogue
//                                -> global.  Two TMEM accumulator buffers overlap epilogue(i) with
//                                the mainloop of tile i+1.
//
// A-operand sources: plain row-major [M,K] (2D tensor map), or NHWC activations walked as a 3x3
// convolution (4D tensor map, one box per filter tap; TMA out-of-bounds zero fill = zero padding).
//
// Fused epilogues (runtime flags): +bias, GELU(erf), ReLU, bf16/f32 store, in-place fp32 residual
// accumulate, up to two bf16 addends, dual (raw + ReLU) store, 2D RoPE on q/k columns (replaces
// croco/models/curope), k==stride transposed-conv pixel scatter, and the DPT "head.4 1x1 conv +
// postprocess" tail (dust3r/heads/postprocess.py) evaluated straight from the accumulator.
#pragma once
#include "d3r_common.cuh"
#include "sm100_ptx.cuh"
#include "pdl.cuh"

namespace d3r {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int kNumEpilogueWarps = 8;  // two warps per TMEM lane quarter, each draining half of the columns
constexpr int kNumThreads = 64 + 32 * kNumEpilogueWarps;

enum : uint32_t {
  F_BIAS = 1u << 0,
  F_GELU = 1u << 1,
  F_RELU = 1u << 2,
  F_OUT_F32 = 1u << 3,        // store fp32 instead of bf16
  F_RESID_INPLACE = 1u << 4,  // out (fp32) += result  (residual stream update)
  F_ADD0 = 1u << 5,           // + bf16 addend 0 (same indexing as out)
  F_ADD1 = 1u << 6,           // + bf16 addend 1
  F_OUT2_RELU = 1u << 7,      // also store relu(result) as bf16 to out2
  F_ROPE = 1u << 8,           // rotate columns < rope_cols (q / k heads of 64)
  F_CONVT = 1u << 9,          // kernel==stride transposed conv scatter
  F_HEAD_FINAL = 1u << 10,    // relu -> 1x1 conv to 4 ch -> pts3d/conf postprocess
  F_OUT2_BF16 = 1u << 11,     // also store the result as bf16 to out2 (used with fp32 primary)
};

struct Params {
  int M, N, K;
  int num_kb;
  int mode;  // 0 plain, 1 conv3x3 (stride 1, pad 1)
  uint32_t flags;
  void* out;
  void* out2;
  const void* add0;
  const void* add1;
  const float* bias;
  long long ldo;
  // RoPE
  const float* rope_cos;  // [max_pos][16]
  const float* rope_sin;
  int rope_cols, tokens_per_img, grid_w;
  // conv geometry
  int cB, cH, cW, cin_blocks, tile_w, tile_h, tiles_x, tiles_y;
  // transposed conv (k == stride): A rows are input pixels (b, iy, ix)
  int tk, th_in, tw_in, tCout;
  // head tail
  const float* w4;  // [4][128]
  const float* b4;  // [4]
  float* pts3d;     // (B,H,W,3)
  float* conf;      // (B,H,W)
  int depth_mode;   // 0 linear, 1 square, 2 exp
  int conf_mode;    // 0 none, 1 exp, 2 sigmoid
  float conf_min, conf_max;
};

// Epilogue specialisations (compile-time): the runtime flag word is masked with the set a specialisation supports,
// so the other epilogue branches are dead code and do not cost registers or instruction cache.
//   EPI_GENERIC  every flag (conv, transposed conv, DPT fusions, head tail, ...)
//   EPI_RESID    out(f32) += acc + bias  through shared memory and a TMA reduce-add (cp.reduce.async.bulk.tensor):
//                the residual stream is never loaded by the SM and is updated in whole 128-byte lines
//   EPI_ACT      bias / GELU / ReLU -> bf16
//   EPI_ROPE     bias + 2D RoPE -> bf16 (q,k,v projections)
enum : int { EPI_GENERIC = 0, EPI_RESID = 1, EPI_ACT = 2, EPI_ROPE = 3 };
template <int EPI> struct EpiMask { static constexpr uint32_t value = 0xFFFFFFFFu; };
template <> struct EpiMask<EPI_RESID> { static constexpr uint32_t value = F_BIAS | F_RESID_INPLACE; };
template <> struct EpiMask<EPI_ACT> { static constexpr uint32_t value = F_BIAS | F_GELU | F_RELU; };
template <> struct EpiMask<EPI_ROPE> { static constexpr uint32_t value = F_BIAS | F_ROPE; };
__host__ __device__ inline int pick_epi(int mode, uint32_t flags) {
  if (mode != 0) return EPI_GENERIC;
  if ((flags & ~F_BIAS) == F_RESID_INPLACE) return EPI_RESID;
  if ((flags & ~EpiMask<EPI_ACT>::value) == 0) return EPI_ACT;
  if ((flags & ~EpiMask<EPI_ROPE>::value) == 0) return EPI_ROPE;
  return EPI_GENERIC;
}
constexpr int kOutStageBytes = 32 * 32 * 4;   // one 32-row x 32-column fp32 box per epilogue warp (EPI_RESID)

template <int BLOCK_N, int EPI = EPI_GENERIC>
struct Cfg {
  static constexpr int kStageA = BLOCK_M * BLOCK_K * 2;
  static constexpr int kStageB = BLOCK_N * BLOCK_K * 2;
  static constexpr int kOutStage = (EPI == EPI_RESID) ? kNumEpilogueWarps * kOutStageBytes : 0;
  static constexpr int kStages = ((BLOCK_N >= 256) ? 4 : ((BLOCK_N >= 128) ? 6 : 8)) - ((EPI == EPI_RESID) ? 1 : 0);
  static constexpr int kTmemCols = (2 * BLOCK_N <= 32) ? 32 : (2 * BLOCK_N <= 64) ? 64 : (2 * BLOCK_N <= 128) ? 128
                                   : (2 * BLOCK_N <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * (kStageA + kStageB) + kOutStage + 1024 /*align*/ + 256 /*barriers*/ + 4 * 128 * 4 + 64 +
                                    kNumEpilogueWarps * 128 * 4 /*bias slices*/;
};

// exact-erf GELU (nn.GELU default).  erf by Abramowitz-Stegun 7.1.26 (|abs err| <= 1.5e-7, far below the
// bf16 rounding of the stored activation).  Written for the epilogue's instruction budget (14 per element, two
// of them MUFU): with u = |x| sqrt(log2(e)/2),  exp(-x^2/2) = 2^(-u u);  h = x/2 poly(t) t 2^(-u u) (the 1/2 is
// folded into the coefficients) is x/2 (1 - erf(|x|/sqrt2)), and  gelu(x) = max(x, 0) - |h|  on both sides of 0.
__device__ __forceinline__ float gelu_erf(float x) {
  constexpr float kU = 0.84932180028801904f;            // sqrt(log2(e) / 2)
  constexpr float kP = 0.3275911f * 0.70710678118654752f / kU;
  const float u = fabsf(x) * kU;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(kP, u, 1.f)));
  float poly = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
  poly = fmaf(poly, t, 0.5f * 1.421413741f);
  poly = fmaf(poly, t, 0.5f * -0.284496736f);
  poly = fmaf(poly, t, 0.5f * 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-u * u));
  const float h = x * (poly * t * e);
  return fmaxf(x, 0.f) - fabsf(h);
}

// 256-bit global accesses (sm_100: LDG/STG.E.ENL2.256): a lane moves one whole 32-byte sector per instruction,
// so the row-per-thread epilogue writes full sectors instead of two 16-byte halves
__device__ __forceinline__ void st256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]),
               "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void ld256(const void* p, uint32_t* r) {
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]),
               "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}
__device__ __forceinline__ void ld256_nc(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]),
               "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const __grid_constant__ CUtensorMap tmap_o, const Params p) {
  using C = Cfg<BLOCK_N, EPI>;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kStageA;
  uint8_t* smem_out = smem + C::kStages * (C::kStageA + C::kStageB);   // EPI_RESID: [kNumEpilogueWarps][32 rows][128 B], 128B-swizzled
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + C::kOutStage);
  uint64_t* full_bar = bars;                      // [kStages]
  uint64_t* empty_bar = bars + C::kStages;        // [kStages]
  uint64_t* tfull_bar = bars + 2 * C::kStages;    // [2]
  uint64_t* tempty_bar = bars + 2 * C::kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);
  float* s_w4 = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 256);  // [4][128] + [4]
  float* s_bias = s_w4 + 4 * 128 + 16;                                             // [kNumEpilogueWarps][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int m_tiles = (p.mode == 1) ? p.cB * p.tiles_y * p.tiles_x : (p.M + BLOCK_M - 1) / BLOCK_M;
  const int total_tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C::kStages; ++s) {
      ptx::mbar_init(ptx::smem_u32(&full_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&empty_bar[s]), 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(ptx::smem_u32(&tfull_bar[s]), 1);
      ptx::mbar_init(ptx::smem_u32(&tempty_bar[s]), kNumEpilogueWarps);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    __syncwarp();
    ptx::tmem_alloc(ptx::smem_u32(tmem_slot), C::kTmemCols);
    ptx::tmem_relinquish();
  }
  if ((p.flags & F_HEAD_FINAL) && threadIdx.x >= 64) {
    for (int i = threadIdx.x - 64; i < 4 * 128 + 4; i += 32 * kNumEpilogueWarps) s_w4[i] = (i < 512) ? p.w4[i] : p.b4[i - 512];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl::sync_with_predecessor();   // set-up done (weights-only reads so far); A / residual / addends come from other kernels

  if (warp == 0) {
    // ================= TMA producer =================
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int tn = tile % n_tiles, tm = tile / n_tiles;
        int cb = 0, cy0 = 0, cx0 = 0;
        if (p.mode == 1) {
          cb = tm / (p.tiles_y * p.tiles_x);
          const int r = tm - cb * (p.tiles_y * p.tiles_x);
          cy0 = (r / p.tiles_x) * p.tile_h;
          cx0 = (r % p.tiles_x) * p.tile_w;
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&empty_bar[stage]), phase ^ 1);
          const uint32_t fb = ptx::smem_u32(&full_bar[stage]);
          ptx::mbar_arrive_expect_tx(fb, C::kStageA + C::kStageB);
          const uint32_t sa = ptx::smem_u32(smem_a + stage * C::kStageA);
          const uint32_t sb = ptx::smem_u32(smem_b + stage * C::kStageB);
          if (p.mode == 1) {
            const int tap = kb / p.cin_blocks, cblk = kb - tap * p.cin_blocks;
            const int dy = tap / 3, dx = tap - dy * 3;
            ptx::tma_load_4d(sa, &tmap_a, fb, cblk * BLOCK_K, cx0 + dx - 1, cy0 + dy - 1, cb);
            ptx::tma_load_3d(sb, &tmap_b, fb, cblk * BLOCK_K, tap, tn * BLOCK_N);
          } else {
            ptx::tma_load_2d(sa, &tmap_a, fb, kb * BLOCK_K, tm * BLOCK_M);
            ptx::tma_load_3d(sb, &tmap_b, fb, kb * BLOCK_K, 0, tn * BLOCK_N);
          }
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        ptx::mbar_wait(ptx::smem_u32(&tempty_bar[acc]), acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          ptx::mbar_wait(ptx::smem_u32(&full_bar[stage]), phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_a + stage * C::kStageA));
          const uint64_t db = ptx::umma_desc_kmajor_sw128(ptx::smem_u32(smem_b + stage * C::kStageB));
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // advance 32 B (=UMMA_K bf16) inside the 128 B swizzle row: +2 in 16-byte units
            ptx::umma_bf16_ss(d_tmem, da + uint64_t(2 * k), db + uint64_t(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(ptx::smem_u32(&empty_bar[stage]));
          if (++stage == C::kStages) { stage = 0; phase ^= 1; }
        }
        ptx::umma_commit(ptx::smem_u32(&tfull_bar[acc]));
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ================= epilogue warps =================
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row_in_tile = quarter * 32 + lane;
    // column split between the two warps of a quarter; the head tail needs whole rows -> first warp only
    constexpr int kChunks = BLOCK_N / 32;
    const int half = (warp - 2) >> 2;
    const bool whole_row = (p.flags & F_HEAD_FINAL) != 0;
    const int ch_begin = whole_row ? 0 : half * (kChunks / 2);
    const int ch_end = whole_row ? (half == 0 ? kChunks : 0) : (half + 1) * (kChunks / 2);
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t flags = p.flags & EpiMask<EPI>::value;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int tn = tile % n_tiles, tm = tile / n_tiles;
      // ---- where does this thread's row live in the output? ----
      bool valid;
      long long row_off;        // element offset of column 0 of this row in out (plain / conv)
      int tok = 0;              // token index inside its image (RoPE)
      long long pix = 0;        // output pixel index (head tail)
      int ct_b = 0, ct_iy = 0, ct_ix = 0;
      if (p.mode == 1) {
        const int cb = tm / (p.tiles_y * p.tiles_x);
        const int r = tm - cb * (p.tiles_y * p.tiles_x);
        const int y = (r / p.tiles_x) * p.tile_h + row_in_tile / p.tile_w;
        const int x = (r % p.tiles_x) * p.tile_w + row_in_tile % p.tile_w;
        valid = (y < p.cH) && (x < p.cW);
        pix = (long long)(cb * p.cH + y) * p.cW + x;
        row_off = pix * p.ldo;
      } else {
        const int row = tm * BLOCK_M + row_in_tile;
        valid = row < p.M;
        row_off = (long long)row * p.ldo;
        if (flags & F_ROPE) tok = row % p.tokens_per_img;
        if (flags & F_CONVT) {
          ct_b = row / (p.th_in * p.tw_in);
          const int r = row - ct_b * (p.th_in * p.tw_in);
          ct_iy = r / p.tw_in;
          ct_ix = r - ct_iy * p.tw_in;
        }
      }
      float head_acc[4] = {0.f, 0.f, 0.f, 0.f};

      // ---- work that does not need the accumulator, issued before waiting for it ----
      // (a) this warp's bias slice -> shared memory (one coalesced load per tile instead of 8 dependent
      //     broadcast loads per 32-column chunk on the epilogue's critical path)
      float* s_bias_w = s_bias + (warp - 2) * 128;
      if (flags & F_BIAS) {
        __syncwarp();
        for (int i = lane * 4; i < (ch_end - ch_begin) * 32; i += 128) {
          const int c = tn * BLOCK_N + ch_begin * 32 + i;
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < p.N) b = __ldg(reinterpret_cast<const float4*>(p.bias + ((flags & F_CONVT) ? (c % p.tCout) : c)));
          *reinterpret_cast<float4*>(s_bias_w + i) = b;
        }
        __syncwarp();
      }
      ptx::mbar_wait(ptx::smem_u32(&tfull_bar[acc]), acc_phase);
      ptx::tc_fence_after();
      // one 32-column chunk; false = past the last column (warp-uniform)
      auto chunk_body = [&](const int ch) -> bool {
        const int col0 = tn * BLOCK_N + ch * 32;
        if (col0 >= p.N) return false;
        uint32_t raw[32];
        ptx::tmem_ld_32x32b_x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BLOCK_N + ch * 32), raw);
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[j]);
        if (flags & F_BIAS) {
          const float4* b4p = reinterpret_cast<const float4*>(s_bias_w + (ch - ch_begin) * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = b4p[j];
            v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        if constexpr (EPI == EPI_RESID) {
          // thread = row: 8 x 16 B into this warp's 128B-swizzled staging box, then one TMA reduce-add of the
          // 32 x 32 box onto the fp32 residual stream (rows past M are clipped by the tensor map)
          uint8_t* stg = smem_out + (warp - 2) * kOutStageBytes;
          if (lane == 0) ptx::bulk_wait_read0();   // the previous box has left shared memory
          __syncwarp();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
          ptx::fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_reduce_add_2d(&tmap_o, ptx::smem_u32(stg), col0, tm * BLOCK_M + quarter * 32);
            ptx::bulk_commit();
          }
          return true;
        }
        if (flags & F_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if ((flags & F_ROPE) && col0 < p.rope_cols) {
          // chunk = one half (y or x) of a 64-wide head: pairs (k, k+16), angle = pos * base^(-k/16)
          const int pos = ((col0 & 63) < 32) ? (tok / p.grid_w) : (tok % p.grid_w);
          const float4* c4 = reinterpret_cast<const float4*>(p.rope_cos + pos * 16);
          const float4* s4 = reinterpret_cast<const float4*>(p.rope_sin + pos * 16);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float4 c = __ldg(c4 + j), s = __ldg(s4 + j);
            const float cc[4] = {c.x, c.y, c.z, c.w}, ss[4] = {s.x, s.y, s.z, s.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float u = v[4 * j + t], w = v[4 * j + t + 16];
              v[4 * j + t] = u * cc[t] - w * ss[t];
              v[4 * j + t + 16] = w * cc[t] + u * ss[t];
            }
          }
        }
        if (flags & F_HEAD_FINAL) {
          // relu(conv) . w4  accumulated across the 4 chunks of the 128-channel row
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float r = fmaxf(v[j], 0.f);
            const int c = ch * 32 + j;
            head_acc[0] += r * s_w4[0 * 128 + c];
            head_acc[1] += r * s_w4[1 * 128 + c];
            head_acc[2] += r * s_w4[2 * 128 + c];
            head_acc[3] += r * s_w4[3 * 128 + c];
          }
          return true;
        }
        if (!valid) return true;
        long long off;
        if (flags & F_CONVT) {
          const int kk = col0 / p.tCout, co = col0 - kk * p.tCout;
          const int ky = kk / p.tk, kx = kk - ky * p.tk;
          off = ((long long)(ct_b * p.th_in * p.tk + ct_iy * p.tk + ky) * (p.tw_in * p.tk) + ct_ix * p.tk + kx) * p.tCout + co;
        } else {
          off = row_off + col0;
        }
        if (flags & F_ADD0) {
          const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(p.add0) + off;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t w[8];
            ld256_nc(a + 16 * j, w);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[t]);
              v[16 * j + 2 * t] += __low2float(h);
              v[16 * j + 2 * t + 1] += __high2float(h);
            }
          }
        }
        if (flags & F_ADD1) {
          const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(p.add1) + off;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t w[8];
            ld256_nc(a + 16 * j, w);
#pragma unroll
            for (int t = 0; t < 8; ++t) {
              const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[t]);
              v[16 * j + 2 * t] += __low2float(h);
              v[16 * j + 2 * t + 1] += __high2float(h);
            }
          }
        }
        if (flags & F_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (flags & (F_OUT_F32 | F_RESID_INPLACE)) {
          float* o = reinterpret_cast<float*>(p.out) + off;
          if (flags & F_RESID_INPLACE) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[8];
              ld256(o + 8 * j, w);
#pragma unroll
              for (int t = 0; t < 8; ++t) v[8 * j + t] += __uint_as_float(w[t]);
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t w[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w[t] = __float_as_uint(v[8 * j + t]);
            st256(o + 8 * j, w);
          }
          if (flags & F_OUT2_BF16) {
            __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.out2) + off;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint32_t w[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) w[t] = pack_bf16(v[16 * j + 2 * t], v[16 * j + 2 * t + 1]);
              st256(o2 + 16 * j, w);
            }
          }
        } else {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + off;
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            uint32_t w[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) w[t] = pack_bf16(v[16 * j + 2 * t], v[16 * j + 2 * t + 1]);
            st256(o + 16 * j, w);
          }
          if (flags & F_OUT2_RELU) {
            __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.out2) + off;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              uint32_t w[8];
#pragma unroll
              for (int t = 0; t < 8; ++t) w[t] = pack_bf16(fmaxf(v[16 * j + 2 * t], 0.f), fmaxf(v[16 * j + 2 * t + 1], 0.f));
              st256(o2 + 16 * j, w);
            }
          }
        }
        return true;
      };
#pragma unroll 1
      for (int ch = ch_begin; ch < ch_end; ++ch)
        if (!chunk_body(ch)) break;
      // accumulator drained: hand the TMEM buffer back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_relaxed(ptx::smem_u32(&tempty_bar[acc]));
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }

      if ((flags & F_HEAD_FINAL) && valid && half == 0) {
        // dust3r/heads/postprocess.py: pts3d = xyz/|xyz| * f(|xyz|), conf = vmin + exp(x) (clipped)
        const float x = head_acc[0] + s_w4[512 + 0], y = head_acc[1] + s_w4[512 + 1], z = head_acc[2] + s_w4[512 + 2];
        float ox = x, oy = y, oz = z;
        if (p.depth_mode != 0) {
          const float d = sqrtf(x * x + y * y + z * z);
          const float dc = fmaxf(d, 1e-8f);
          const float s = (p.depth_mode == 2) ? expm1f(d) : d * d;
          ox = x / dc * s; oy = y / dc * s; oz = z / dc * s;
        }
        float* o = p.pts3d + pix * 3;
        o[0] = ox; o[1] = oy; o[2] = oz;
        if (p.conf_mode != 0) {
          const float c = head_acc[3] + s_w4[512 + 3];
          float r;
          if (p.conf_mode == 1) r = p.conf_min + fminf(expf(c), p.conf_max - p.conf_min);
          else r = (p.conf_max - p.conf_min) * (1.f / (1.f + expf(-c))) + p.conf_min;
          p.conf[pix] = r;
        }
      }
    }
  }

  if constexpr (EPI == EPI_RESID) {
    if (warp >= 2 && lane == 0) ptx::bulk_wait0();   // staged boxes fully written before shared memory goes away
  }
  // ---- teardown ----
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

}  // namespace gemm
}  // namespace d3r
