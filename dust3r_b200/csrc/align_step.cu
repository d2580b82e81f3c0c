// Fused global-alignment step for B200 (sm_100a): one launch per Adam iteration.
//
// Replaces, per iteration, the ~dozens of elementwise/bmm/gather kernels + autograd + foreach-Adam
// of the reference loop (dust3r/cloud_opt/base_opt.py:352-366 driving optimizer.py:188-201 or
// base_opt.py:246-273).  HBM-bound fp32: every observation float4 (pred.xyz, weight) is read once,
// every log-depth and its two Adam moments are read and written once  -> 32*E*P + 24*n*P bytes
// per iteration.  No tensor cores (K=3 contractions).
//
// Work decomposition
//   CTA  = (image i, chunk of kChunk pixels).  Each thread owns kPPT pixels, keeps their world
//          points X and the accumulated dL/dX in registers and loops over the entries
//          (edge, side) incident to image i, streaming the entry's float4 observations.
//   per entry the CTA reduces 13 sums (sum g (x) q, sum g, loss) with warp shuffles -> smem ->
//          one partial row in global memory (deterministic: no float atomics anywhere).
//   after the entry loop the thread has dL/dX for its pixels: depth gradient + Adam in place.
//   last CTA of an image (atomic ticket) reduces that image's partial rows; last CTA of the grid
//          turns the sums into pose/focal/pp/pairwise-pose gradients (quaternion-normalise,
//          signed_expm1 and the mean-coupled scale backward done analytically), applies Adam and
//          writes the transforms of the next iteration.
#include "d3r_common.cuh"
#include "prof.h"

namespace d3r {
namespace align {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kPPT = 8;                    // pixels per thread
constexpr int kChunk = kThreads * kPPT;    // pixels per CTA
constexpr int kEntVals = 13;               // 9 (g (x) q) + 3 (g) + 1 (loss)
constexpr int kImgVals = 12;               // 9 (G (x) c) + 3 (G)
constexpr int kEdgeT = 12;                 // M = s*R*diag(adapt) (9) + s*T (3)
constexpr int kImgT = 16;                  // R (9) T (3) 1/fx 1/fy cx cy

struct Workspace {
  float* edgeT;          // [E][12]
  float* imgT;           // [n][16]
  long long* ent_acc;    // [2E][13]  fixed-point (2^44) accumulators, zero between launches
  long long* img_acc;    // [n][12]
  float* g_edge;         // [E][10]   small-step scratch: pairwise-pose / adaptor gradients
  float* g_img;          // [n][11]   pose (7) / focal (2) / pp (2) gradients
  int* flags;            // [4]       [0] = fixed-point overflow seen
};

__host__ __device__ inline int64_t align4(int64_t x) { return (x + 3) & ~int64_t(3); }

__host__ __device__ inline Workspace carve(float* ws, int n, int E) {
  Workspace w;
  int64_t o = 0;
  w.edgeT = ws + o;    o += align4(int64_t(E) * kEdgeT);
  w.imgT = ws + o;     o += align4(int64_t(n) * kImgT);
  w.ent_acc = reinterpret_cast<long long*>(ws + o); o += align4(int64_t(2) * E * kEntVals * 2);
  w.img_acc = reinterpret_cast<long long*>(ws + o); o += align4(int64_t(n) * kImgVals * 2);
  w.g_edge = ws + o;   o += align4(int64_t(E) * 10);
  w.g_img = ws + o;    o += align4(int64_t(n) * 11);
  w.flags = reinterpret_cast<int*>(ws + o); o += 4;
  return w;
}

inline int64_t workspace_floats(int n, int E) {
  return align4(int64_t(E) * kEdgeT) + align4(int64_t(n) * kImgT) + align4(int64_t(2) * E * kEntVals * 2) +
         align4(int64_t(n) * kImgVals * 2) + align4(int64_t(E) * 10) + align4(int64_t(n) * 11) + 4;
}

// Order-independent (hence deterministic) cross-CTA accumulation: every CTA contributes its exactly-ordered fp32
// partial sum as a 2^44 fixed-point integer through a 64-bit integer atomic.  Resolution 5.7e-14, range +-5e5.
constexpr double kFixScale = 17592186044416.0;        // 2^44
constexpr double kFixInv = 1.0 / 17592186044416.0;
__device__ __forceinline__ void fix_add(long long* dst, float x, int* overflow_flag) {
  if (!(fabsf(x) < 262144.f)) *overflow_flag = 1;     // also catches NaN / Inf
  const long long q = __double2ll_rn(double(x) * kFixScale);
  atomicAdd(reinterpret_cast<unsigned long long*>(dst), static_cast<unsigned long long>(q));
}
__device__ __forceinline__ float fix_get(const long long* src) {
  const long long q = __ldcg(src);
  return float(double(q) * kFixInv);
}

// offsets inside the `small` parameter buffer
struct SmallLayout {
  int poses, focals, pp, pw, adapt, total;
  __host__ __device__ SmallLayout(int n, int E) {
    poses = 0; focals = n * 7; pp = focals + n * 2; pw = pp + n * 2; adapt = pw + E * 8; total = adapt + E * 2;
  }
};

__device__ __forceinline__ float signed_expm1f(float x) {
  float s = (x > 0.f) - (x < 0.f);
  return s * expm1f(fabsf(x));
}

// unit quaternion (x,y,z,w) -> rotation, row-major R[a*3+b]
__device__ __forceinline__ void quat_to_R(const float* q, float* R, float* qhat, float* nrm) {
  float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  float x = q[0] / n, y = q[1] / n, z = q[2] / n, w = q[3] / n;
  R[0] = 1.f - 2.f * (y * y + z * z); R[1] = 2.f * (x * y - w * z);       R[2] = 2.f * (x * z + w * y);
  R[3] = 2.f * (x * y + w * z);       R[4] = 1.f - 2.f * (x * x + z * z); R[5] = 2.f * (y * z - w * x);
  R[6] = 2.f * (x * z - w * y);       R[7] = 2.f * (y * z + w * x);       R[8] = 1.f - 2.f * (x * x + y * y);
  if (qhat) { qhat[0] = x; qhat[1] = y; qhat[2] = z; qhat[3] = w; }
  if (nrm) *nrm = n;
}

// dL/dq (raw, un-normalised) from D = dL/dR
__device__ __forceinline__ void quat_backward(const float* D, const float* qh, float n, float* gq) {
  float x = qh[0], y = qh[1], z = qh[2], w = qh[3];
  float gx = 2.f * (y * (D[1] + D[3]) + z * (D[2] + D[6]) - 2.f * x * (D[4] + D[8]) + w * (D[7] - D[5]));
  float gy = 2.f * (x * (D[1] + D[3]) + z * (D[5] + D[7]) - 2.f * y * (D[0] + D[8]) + w * (D[2] - D[6]));
  float gz = 2.f * (x * (D[2] + D[6]) + y * (D[5] + D[7]) - 2.f * z * (D[0] + D[4]) + w * (D[3] - D[1]));
  float gw = 2.f * (x * (D[7] - D[5]) + y * (D[2] - D[6]) + z * (D[3] - D[1]));
  float dot = gx * x + gy * y + gz * z + gw * w;
  gq[0] = (gx - x * dot) / n; gq[1] = (gy - y * dot) / n; gq[2] = (gz - z * dot) / n; gq[3] = (gw - w * dot) / n;
}

// torch.optim.Adam single-tensor math (lerp for exp_avg; mul+addcmul for exp_avg_sq; addcdiv)
__device__ __forceinline__ float adam_update(float p, float g, float& m, float& v, float beta1, float beta2,
                                             float step_size, float bc2_sqrt, float eps) {
  m = m + (1.f - beta1) * (g - m);
  v = v * beta2 + (1.f - beta2) * g * g;
  float denom = sqrtf(v) / bc2_sqrt + eps;
  return p - step_size * (m / denom);
}

// optional timeline instrumentation (debug aid): 4 x uint64 globaltimer stamps per CTA
__device__ unsigned long long* g_align_dbg = nullptr;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

#define D3R_TSTAMP(i) do { if (g_align_dbg && threadIdx.x == 0) g_align_dbg[4 * size_t(D.n_chunks) + (i)] = gtime(); } while (0)

// ---- derived transforms / small-parameter step (run by ONE CTA while the rest of the chip idles) ----------
// Written for latency: independent global loads are issued together, block reductions cost one barrier (every
// thread re-adds the 8 warp partials itself), pointers are __restrict__, and the edge work (low thread ids)
// and image work (high thread ids) run concurrently on different warps.
__device__ __forceinline__ float block_sum8(float v, float* slot /* 8 floats */) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) slot[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) t += slot[w];
  return t;
}

struct EdgeGeom {
  float R[9], T[3], ad[3], s, qh[4], qn;
};
__device__ __forceinline__ void edge_geom(const float* p8, float a0, float a1, const d3r_align_desc& D, float mean_sigma,
                                          float log_base, EdgeGeom& g) {
  quat_to_R(p8, g.R, g.qh, &g.qn);
  g.s = expf(p8[7]);
  if (D.norm_pw_scale) g.s *= expf(log_base - mean_sigma);
  g.ad[0] = a0; g.ad[1] = a0; g.ad[2] = a1;
  if (D.norm_pw_scale) {
    const float mu = (a0 + a0 + a1) / 3.f;
    g.ad[0] -= mu; g.ad[1] -= mu; g.ad[2] -= mu;
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) g.ad[b] = expf(g.ad[b] / D.pw_break);
#pragma unroll
  for (int a = 0; a < 3; ++a) g.T[a] = signed_expm1f(p8[4 + a]);
}

// image i is handled by thread (blockDim-1-i) so that images and edges land on different warps
__device__ __forceinline__ int img_of_thread(int it) { return int(blockDim.x) - 1 - int(threadIdx.x) + it * int(blockDim.x); }

__device__ void compute_transforms(const d3r_align_desc& D, const Workspace& ws, float* s_red) {
  const int n = D.n_imgs, E = D.n_edges;
  const SmallLayout L(n, E);
  const float* __restrict__ sm = D.small;
  float part = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) part += sm[L.pw + e * 8 + 7];
  const float mean_sigma = block_sum8(part, s_red) / float(E);   // base_opt.py:178-184
  const float log_base = logf(D.base_scale);
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float p8[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) p8[k] = sm[L.pw + e * 8 + k];
    const float a0 = sm[L.adapt + e * 2 + 0], a1 = sm[L.adapt + e * 2 + 1];
    EdgeGeom g;
    edge_geom(p8, a0, a1, D, mean_sigma, log_base, g);
    float* o = ws.edgeT + e * kEdgeT;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b) o[a * 3 + b] = g.s * g.R[a * 3 + b] * g.ad[b];
#pragma unroll
    for (int a = 0; a < 3; ++a) o[9 + a] = g.s * g.T[a];
  }
  for (int r = 0, i = img_of_thread(0); i < n; i = img_of_thread(++r)) {
    float p7[7];
#pragma unroll
    for (int k = 0; k < 7; ++k) p7[k] = sm[L.poses + i * 7 + k];
    const float f0 = sm[L.focals + i * 2 + 0], f1 = sm[L.focals + i * 2 + 1];
    const float pp0 = sm[L.pp + i * 2 + 0], pp1 = sm[L.pp + i * 2 + 1];
    const int Hh = D.img_hw[i * 2 + 0], Ww = D.img_hw[i * 2 + 1];
    float* o = ws.imgT + i * kImgT;
    quat_to_R(p7, o, nullptr, nullptr);
#pragma unroll
    for (int a = 0; a < 3; ++a) o[9 + a] = signed_expm1f(p7[4 + a]);
    o[12] = 1.f / expf(f0 / D.focal_break);
    o[13] = 1.f / expf(f1 / D.focal_break);
    o[14] = 0.5f * float(Ww) + 10.f * pp0;
    o[15] = 0.5f * float(Hh) + 10.f * pp1;
  }
}

__global__ void __launch_bounds__(kThreads) prepare_kernel(const __grid_constant__ d3r_align_desc D) {
  __shared__ float s_red[40];
  Workspace ws = carve(D.workspace, D.n_imgs, D.n_edges);
  compute_transforms(D, ws, s_red);
}

// backward through the small parameters + Adam; run by the last CTA of the grid
__device__ void small_param_step(const d3r_align_desc& D, const Workspace& ws, int it, float* s_red) {
  const int n = D.n_imgs, E = D.n_edges;
  const SmallLayout L(n, E);
  float* __restrict__ sm = D.small;
  float* __restrict__ am = D.small_m;
  float* __restrict__ av = D.small_v;
  const uint8_t* __restrict__ tr = D.small_trainable;
  const float step_size = D.sched[it * 4 + 1], bc2s = D.sched[it * 4 + 2];
  const float b1 = D.beta1, b2 = D.beta2, eps = D.adam_eps;

  // phase 0: loss (fixed order over entries) and mean log-scale, one barrier
  float lpart = 0.f, spart = 0.f;
  for (int k = threadIdx.x; k < 2 * E; k += blockDim.x) lpart += fix_get(ws.ent_acc + k * kEntVals + 12);
  for (int e = threadIdx.x; e < E; e += blockDim.x) spart += sm[L.pw + e * 8 + 7];
  lpart = warp_sum(lpart);
  spart = warp_sum(spart);
  if ((threadIdx.x & 31) == 0) { s_red[threadIdx.x >> 5] = lpart; s_red[8 + (threadIdx.x >> 5)] = spart; }
  __syncthreads();
  float loss = 0.f, mean_sigma = 0.f;
#pragma unroll
  for (int w = 0; w < kWarps; ++w) { loss += s_red[w]; mean_sigma += s_red[8 + w]; }
  mean_sigma /= float(E);
  if (threadIdx.x == 0) D.loss_out[it] = loss;
  const float log_base = logf(D.base_scale);
  D3R_TSTAMP(0);

  // phase 1: gradients -> scratch.  edges on low thread ids, images on high thread ids (different warps).
  float coupl = 0.f;
  if (!D.eval_only) {
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
      float p8[8], si[12], sj[12];
      const int ei = D.edge_ent[e * 2 + 0], ej = D.edge_ent[e * 2 + 1];
#pragma unroll
      for (int k = 0; k < 8; ++k) p8[k] = sm[L.pw + e * 8 + k];
      const float a0 = sm[L.adapt + e * 2 + 0], a1 = sm[L.adapt + e * 2 + 1];
#pragma unroll
      for (int k = 0; k < 12; ++k) { si[k] = fix_get(ws.ent_acc + ei * kEntVals + k); sj[k] = fix_get(ws.ent_acc + ej * kEntVals + k); }
      EdgeGeom g;
      edge_geom(p8, a0, a1, D, mean_sigma, log_base, g);
      float dM[9], dt[3];
#pragma unroll
      for (int k = 0; k < 9; ++k) dM[k] = -(si[k] + sj[k]);   // dL/dM_ab = -sum g_a q_b
#pragma unroll
      for (int a = 0; a < 3; ++a) dt[a] = -(si[9 + a] + sj[9 + a]);
      float dLds = 0.f, dR[9], dad[3] = {0.f, 0.f, 0.f};
#pragma unroll
      for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          dLds += dM[a * 3 + b] * g.R[a * 3 + b] * g.ad[b];
          dR[a * 3 + b] = dM[a * 3 + b] * g.s * g.ad[b];
          dad[b] += dM[a * 3 + b] * g.s * g.R[a * 3 + b];
        }
        dLds += dt[a] * g.T[a];
      }
      float gr[10];
      quat_backward(dR, g.qh, g.qn, gr);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float t = p8[4 + a];
        const float sg = (t > 0.f) - (t < 0.f);
        gr[4 + a] = dt[a] * g.s * sg * sg * expf(fabsf(t));
      }
      gr[7] = dLds * g.s;             // the mean-coupling term is subtracted in phase 2
      coupl += dLds * g.s;
      // adaptors (base_opt.py:143-148): adapt3 = exp((cat(a0,a0,a1) - mean)/pw_break)
      float gad[3];
#pragma unroll
      for (int b = 0; b < 3; ++b) gad[b] = dad[b] * g.ad[b] / D.pw_break;
      if (D.norm_pw_scale) { const float mu = (gad[0] + gad[1] + gad[2]) / 3.f; gad[0] -= mu; gad[1] -= mu; gad[2] -= mu; }
      gr[8] = gad[0] + gad[1];
      gr[9] = gad[2];
#pragma unroll
      for (int k = 0; k < 10; ++k) ws.g_edge[e * 10 + k] = gr[k];
    }
    for (int r = 0, i = img_of_thread(0); i < n; i = img_of_thread(++r)) {
      float q7[7], S[12];
#pragma unroll
      for (int k = 0; k < 7; ++k) q7[k] = sm[L.poses + i * 7 + k];
      const float f0 = sm[L.focals + i * 2 + 0], f1 = sm[L.focals + i * 2 + 1];
#pragma unroll
      for (int k = 0; k < 12; ++k) S[k] = fix_get(ws.img_acc + i * kImgVals + k);   // S[a*3+b] = sum G_a c_b ; S[9+a] = sum G_a
      float R[9], qh[4], qn;
      quat_to_R(q7, R, qh, &qn);
      float gr[11];
      quat_backward(S, qh, qn, gr);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float t = q7[4 + a];
        const float sg = (t > 0.f) - (t < 0.f);
        gr[4 + a] = S[9 + a] * sg * sg * expf(fabsf(t));
      }
      // focals: c_x = d (u-cx)/fx, fx = exp(phi/focal_break);  pp: cx = W/2 + 10 pp_x
      float gfx = 0.f, gfy = 0.f, gpx = 0.f, gpy = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        gfx += R[a * 3 + 0] * S[a * 3 + 0];
        gfy += R[a * 3 + 1] * S[a * 3 + 1];
        gpx += R[a * 3 + 0] * S[a * 3 + 2];
        gpy += R[a * 3 + 1] * S[a * 3 + 2];
      }
      gfx = -gfx / D.focal_break;
      gfy = -gfy / D.focal_break;
      // one shared focal: both slots get the summed gradient and evolve identically
      gr[7] = D.tied_focal ? gfx + gfy : gfx;
      gr[8] = D.tied_focal ? gfx + gfy : gfy;
      gr[9] = -10.f * gpx / expf(f0 / D.focal_break);
      gr[10] = -10.f * gpy / expf(f1 / D.focal_break);
#pragma unroll
      for (int k = 0; k < 11; ++k) ws.g_img[i * 11 + k] = gr[k];
    }
  }
  D3R_TSTAMP(1);
  // zero the accumulators for the next launch (everything has been read above; barrier inside block_sum8)
  const float coupling = block_sum8(coupl, s_red + 16) / float(E);
  for (int k = threadIdx.x; k < 2 * E * kEntVals; k += blockDim.x) ws.ent_acc[k] = 0;
  for (int k = threadIdx.x; k < n * kImgVals; k += blockDim.x) ws.img_acc[k] = 0;
  if (D.eval_only) return;
  D3R_TSTAMP(2);

  // phase 2: Adam over the flat parameter vector  [poses 7n | focals 2n | pp 2n | pw 8E | adapt 2E]
  for (int idx = threadIdx.x; idx < L.total; idx += blockDim.x) {
    if (!tr[idx]) continue;
    float g;
    if (idx < L.focals) {
      const int i = idx / 7;
      g = ws.g_img[i * 11 + (idx - i * 7)];
    } else if (idx < L.pp) {
      const int r = idx - L.focals;
      g = ws.g_img[(r >> 1) * 11 + 7 + (r & 1)];
    } else if (idx < L.pw) {
      const int r = idx - L.pp;
      g = ws.g_img[(r >> 1) * 11 + 9 + (r & 1)];
    } else if (idx < L.adapt) {
      const int r = idx - L.pw;
      g = ws.g_edge[(r >> 3) * 10 + (r & 7)];
      if ((r & 7) == 7 && D.norm_pw_scale) g -= coupling;
    } else {
      const int r = idx - L.adapt;
      g = ws.g_edge[(r >> 1) * 10 + 8 + (r & 1)];
    }
    float m = am[idx], v = av[idx];
    sm[idx] = adam_update(sm[idx], g, m, v, b1, b2, step_size, bc2s, eps);
    am[idx] = m;
    av[idx] = v;
  }
  D3R_TSTAMP(3);
  __syncthreads();
  compute_transforms(D, ws, s_red + 24);
  D3R_TSTAMP(4);
}

// ---- the per-iteration kernel ---------------------------------------------------------------
// 8 warps, all computing; lane 0 of warp 0 doubles as the producer: each incident entry's float4 observations
// for this CTA's pixel chunk are streamed into a 3-stage shared-memory ring with cp.async.bulk (TMA 1-D copies,
// mbarrier completion), refilled as soon as a stage is drained, so up to 2 x 84 KB of reads are in flight per
// SM independent of the warps' compute progress.
constexpr int kStages = 3;
constexpr int kEntTile = 16;               // entries whose per-warp partial sums are staged in smem at a time
constexpr int kRedVals = 16;               // 13 padded to 16 for the halving butterfly

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// Sum 16 per-lane values across the warp with 16 shuffles (instead of 16 x 5): at every halving step a lane
// keeps one half of its values and trades the other half with its partner.  On return lane l holds the
// warp-wide total of value index (l >> 1) & 15 in v[0].
__device__ __forceinline__ float butterfly16(float (&v)[kRedVals], int lane) {
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = up ? v[i] : v[i + 8];
      const float keep = up ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = up ? v[i] : v[i + 4];
      const float keep = up ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = up ? v[i] : v[i + 2];
      const float keep = up ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1];
    const float keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
  return v[0];
}

template <bool kL2, int PPT>
__global__ void __launch_bounds__(kThreads, 2)
align_iter_kernel(const __grid_constant__ d3r_align_desc D, int it) {
  constexpr int kSlots = PPT * kThreads;     // pixel slots of this instantiation (>= chunk_px)
  extern __shared__ __align__(128) uint8_t s_dyn[];
  float4* s_obs = reinterpret_cast<float4*>(s_dyn);                                   // [kStages][kSlots]
  float* s_ent = reinterpret_cast<float*>(s_dyn + kStages * kSlots * sizeof(float4));  // [kEntTile][kWarps][13]
  __shared__ float s_img[kWarps * kImgVals];
  __shared__ float s_red[40];
  __shared__ int s_flag;
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];

  const Workspace ws = carve(D.workspace, D.n_imgs, D.n_edges);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x;
  unsigned long long* dbg = g_align_dbg ? g_align_dbg + 4 * size_t(chunk) : nullptr;
  if (dbg && tid == 0) dbg[0] = gtime();
  const int img = D.chunk_img[chunk];
  const int lc = chunk - D.img_chunk_ptr[img];
  const int H = D.img_hw[img * 2 + 0], W = D.img_hw[img * 2 + 1];
  const int P = H * W;
  const int pbase = lc * D.chunk_px;                       // first pixel of this CTA
  const int npx = min(D.chunk_px, P - pbase);              // 1 .. kSlots
  const int64_t poff = D.img_pix_off[img];
  const int e0 = D.img_ent_ptr[img], e1 = D.img_ent_ptr[img + 1];
  const int deg = e1 - e0;
  const float4* obs_base = reinterpret_cast<const float4*>(D.obs);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(smem_u32(&s_full[s]), 1);
      mbar_init(smem_u32(&s_empty[s]), kWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // slots past the chunk are never written by the bulk copies: zero them once (weight 0 -> no contribution)
  for (int s = 0; s < kStages; ++s)
    for (int q = npx + tid; q < kSlots; q += kThreads) s_obs[s * kSlots + q] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();

  const uint32_t stage_bytes = uint32_t(npx) * 16u;
  auto produce = [&](int k) {   // called by one thread
    const int s = k % kStages;
    mbar_wait(smem_u32(&s_empty[s]), ((k / kStages) & 1) ^ 1);
    mbar_expect_tx(smem_u32(&s_full[s]), stage_bytes);
    bulk_g2s(smem_u32(s_obs + s * kSlots), obs_base + D.ent_obs_off[e0 + k] + pbase, stage_bytes, smem_u32(&s_full[s]));
  };
  if (tid == 0)
    for (int k = 0; k < min(kStages, deg); ++k) produce(k);

  // Programmatic dependent launch: everything above touches only per-problem constants (index tables, the
  // observation slabs) -- this iteration's CTAs were allowed to start it while the previous iteration's last CTA
  // was still in its small-parameter step.  Everything below reads what that step (and the previous depth update)
  // wrote, so wait here for the previous grid to complete and flush.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // ... and allow the NEXT iteration's CTAs to be scheduled as soon as every CTA of this grid has got this far: they
  // take the slots of this grid's last wave as its CTAs retire, and block at their own griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const float* iT = ws.imgT + img * kImgT;
  float R[9], T[3];
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = iT[k];
#pragma unroll
  for (int k = 0; k < 3; ++k) T[k] = iT[9 + k];
  const float ifx = iT[12], ify = iT[13], cx = iT[14], cy = iT[15];

  float X[PPT][3], G[PPT][3];
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int q = tid + k * kThreads;
    const int p = pbase + q;
    const float ld = (q < npx) ? D.logd[poff + p] : 0.f;
    const float d = expf(ld);
    const int v = p / W, u = p - v * W;
    const float c0 = d * (float(u) - cx) * ifx, c1 = d * (float(v) - cy) * ify;
    X[k][0] = R[0] * c0 + R[1] * c1 + R[2] * d + T[0];
    X[k][1] = R[3] * c0 + R[4] * c1 + R[5] * d + T[1];
    X[k][2] = R[6] * c0 + R[7] * c1 + R[8] * d + T[2];
    G[k][0] = G[k][1] = G[k][2] = 0.f;
  }

  __shared__ __align__(16) float s_T[kEntTile][16];   // per-entry transform M (9), t (3), coef: staged per tile
  int stage = 0;
  uint32_t stage_phase = 0;
  for (int k0 = 0; k0 < deg; k0 += kEntTile) {
    const int kend = min(deg, k0 + kEntTile);
    // stage this tile's edge transforms in shared memory (one coalesced pass instead of 13 dependent LDGs per
    // entry per thread sitting on the critical path of every entry)
    if (tid < (kend - k0) * 13) {
      const int k = tid / 13, v = tid - k * 13;
      const int ent = e0 + k0 + k;
      s_T[k][v] = (v < 12) ? ws.edgeT[D.ent_edge[ent] * kEdgeT + v] : D.ent_coef[ent];
    }
    __syncthreads();
    for (int kk = k0; kk < kend; ++kk) {
      const float4 t0 = *reinterpret_cast<const float4*>(&s_T[kk - k0][0]);
      const float4 t1 = *reinterpret_cast<const float4*>(&s_T[kk - k0][4]);
      const float4 t2 = *reinterpret_cast<const float4*>(&s_T[kk - k0][8]);
      const float coef = s_T[kk - k0][12];
      const float M[9] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x};
      const float t[3] = {t2.y, t2.z, t2.w};
      const int s = stage;
      mbar_wait(smem_u32(&s_full[s]), stage_phase);
      if (++stage == kStages) { stage = 0; stage_phase ^= 1; }
      const float4* so = s_obs + s * kSlots;
      float acc[kRedVals];
#pragma unroll
      for (int k = 0; k < kRedVals; ++k) acc[k] = 0.f;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        const float4 o = so[tid + k * kThreads];
        const float qx = o.x, qy = o.y, qz = o.z;
        const float lw = coef * o.w;
        // r = X - (M q + t), written as FMA chains (the compiler may not re-associate fp32)
        const float r0 = X[k][0] - fmaf(M[0], qx, fmaf(M[1], qy, fmaf(M[2], qz, t[0])));
        const float r1 = X[k][1] - fmaf(M[3], qx, fmaf(M[4], qy, fmaf(M[5], qz, t[1])));
        const float r2 = X[k][2] - fmaf(M[6], qx, fmaf(M[7], qy, fmaf(M[8], qz, t[2])));
        const float rho2 = fmaf(r0, r0, fmaf(r1, r1, r2 * r2));
        float gs;
        if (kL2) {
          acc[12] = fmaf(lw, rho2, acc[12]);
          gs = 2.f * lw;
        } else {
          // torch's norm backward yields 0 at ||r|| == 0
          const float inv = rho2 > 0.f ? rsqrtf(rho2) : 0.f;
          acc[12] = fmaf(lw, rho2 * inv, acc[12]);
          gs = lw * inv;
        }
        const float g0 = gs * r0, g1 = gs * r1, g2 = gs * r2;
        G[k][0] += g0; G[k][1] += g1; G[k][2] += g2;
        acc[0] = fmaf(g0, qx, acc[0]); acc[1] = fmaf(g0, qy, acc[1]); acc[2] = fmaf(g0, qz, acc[2]);
        acc[3] = fmaf(g1, qx, acc[3]); acc[4] = fmaf(g1, qy, acc[4]); acc[5] = fmaf(g1, qz, acc[5]);
        acc[6] = fmaf(g2, qx, acc[6]); acc[7] = fmaf(g2, qy, acc[7]); acc[8] = fmaf(g2, qz, acc[8]);
        acc[9] += g0; acc[10] += g1; acc[11] += g2;
      }
      // this warp is done with the stage: hand it back, and (thread 0) refill it with entry kk + kStages
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));
      if (tid == 0 && kk + kStages < deg) produce(kk + kStages);
      const float tot = butterfly16(acc, lane);
      const int vi = (lane >> 1) & 15;
      if (!(lane & 1) && vi < kEntVals) s_ent[((kk - k0) * kWarps + warp) * kEntVals + vi] = tot;
    }
    // cross-warp sums of this tile of entries (fixed order) -> deterministic fixed-point accumulation
    __syncthreads();
    for (int idx = tid; idx < (kend - k0) * kEntVals; idx += kThreads) {
      const int k = idx / kEntVals, v = idx - k * kEntVals;
      float sacc = 0.f;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) sacc += s_ent[(k * kWarps + w) * kEntVals + v];
      fix_add(ws.ent_acc + int64_t(e0 + k0 + k) * kEntVals + v, sacc, ws.flags);
    }
    __syncthreads();
  }

  // depth gradient + Adam (in place), per-image pose/focal sums
  float S[kRedVals];
#pragma unroll
  for (int k = 0; k < kRedVals; ++k) S[k] = 0.f;
  if (!D.eval_only) {
    const float step_size = D.sched[it * 4 + 1], bc2s = D.sched[it * 4 + 2];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
      const int q = tid + k * kThreads;
      if (q < npx) {
        const int p = pbase + q;
        const float ld = D.logd[poff + p];
        const float d = expf(ld);
        const int v = p / W, u = p - v * W;
        const float c0 = d * (float(u) - cx) * ifx, c1 = d * (float(v) - cy) * ify;
        // dX/dlogd = R c  (c is linear in d)
        const float gd = G[k][0] * (X[k][0] - T[0]) + G[k][1] * (X[k][1] - T[1]) + G[k][2] * (X[k][2] - T[2]);
        float m = D.logd_m[poff + p], vv = D.logd_v[poff + p];
        const float nld = adam_update(ld, gd, m, vv, D.beta1, D.beta2, step_size, bc2s, D.adam_eps);
        D.logd[poff + p] = nld;
        D.logd_m[poff + p] = m;
        D.logd_v[poff + p] = vv;
        S[0] += G[k][0] * c0; S[1] += G[k][0] * c1; S[2] += G[k][0] * d;
        S[3] += G[k][1] * c0; S[4] += G[k][1] * c1; S[5] += G[k][1] * d;
        S[6] += G[k][2] * c0; S[7] += G[k][2] * c1; S[8] += G[k][2] * d;
        S[9] += G[k][0]; S[10] += G[k][1]; S[11] += G[k][2];
      }
    }
  }
  {
    const float tot = butterfly16(S, lane);
    const int vi = (lane >> 1) & 15;
    if (!(lane & 1) && vi < kImgVals) s_img[warp * kImgVals + vi] = tot;
  }
  __syncthreads();
  if (tid < kImgVals) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += s_img[w * kImgVals + tid];
    fix_add(ws.img_acc + int64_t(img) * kImgVals + tid, s, ws.flags);
  }
  if (dbg && tid == 0) dbg[1] = gtime();

  // ---- grid ticket: the last CTA to finish runs the small-parameter step ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(D.counters, 1) == int(gridDim.x) - 1);
  __syncthreads();
  if (!s_flag) {
    if (dbg && tid == 0) dbg[2] = gtime();
    return;
  }
  __threadfence();
  if (tid == 0) D.counters[0] = 0;   // re-arm for the next launch
  if (dbg && tid == 0) dbg[2] = gtime();
  small_param_step(D, ws, it, s_red);
  if (dbg && tid == 0) dbg[3] = gtime();
}

__global__ void __launch_bounds__(kThreads) pts3d_kernel(const __grid_constant__ d3r_align_desc D, float* out) {
  const Workspace ws = carve(D.workspace, D.n_imgs, D.n_edges);
  const int chunk = blockIdx.x;
  const int img = D.chunk_img[chunk];
  const int lc = chunk - D.img_chunk_ptr[img];
  const int H = D.img_hw[img * 2 + 0], W = D.img_hw[img * 2 + 1];
  const int P = H * W;
  const int64_t poff = D.img_pix_off[img];
  const float* iT = ws.imgT + img * kImgT;
  const int pend = min(P, (lc + 1) * D.chunk_px);
  for (int p = lc * D.chunk_px + threadIdx.x; p < pend; p += kThreads) {
    const float d = expf(D.logd[poff + p]);
    const int v = p / W, u = p - v * W;
    const float c0 = d * (float(u) - iT[14]) * iT[12], c1 = d * (float(v) - iT[15]) * iT[13];
    float* o = out + (poff + p) * 3;
    o[0] = iT[0] * c0 + iT[1] * c1 + iT[2] * d + iT[9];
    o[1] = iT[3] * c0 + iT[4] * c1 + iT[5] * d + iT[10];
    o[2] = iT[6] * c0 + iT[7] * c1 + iT[8] * d + iT[11];
  }
}

__global__ void pack_obs_kernel(const float* __restrict__ pts, const float* __restrict__ w, float4* __restrict__ obs,
                                int64_t n) {
  int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) obs[i] = make_float4(pts[i * 3 + 0], pts[i * 3 + 1], pts[i * 3 + 2], w[i]);
}

}  // namespace align
}  // namespace d3r

using namespace d3r;
using namespace d3r::align;

extern "C" int d3r_align_set_debug(void* dev_buf) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(dev_buf);
  D3R_CUDA(cudaMemcpyToSymbol(g_align_dbg, &p, sizeof(p)));
  return D3R_OK;
}

extern "C" int d3r_align_chunk_pixels(void) { return kChunk; }
extern "C" int d3r_sizeof_align_desc(void) { return (int)sizeof(d3r_align_desc); }

extern "C" int64_t d3r_align_workspace_floats(int32_t n_imgs, int32_t n_edges, int32_t n_chunks, int32_t max_chunks) {
  (void)n_chunks; (void)max_chunks;   // kept in the signature for ABI stability; accumulators are per entry now
  return workspace_floats(n_imgs, n_edges);
}

static int validate(const d3r_align_desc* d) {
  D3R_CHECK_ARG(d != nullptr, "d3r_align: null descriptor");
  D3R_CHECK_ARG(d->n_imgs > 0 && d->n_edges > 0 && d->n_entries == 2 * d->n_edges, "d3r_align: bad sizes");
  D3R_CHECK_ARG(d->n_chunks > 0 && d->max_chunks > 0 && d->max_deg > 0, "d3r_align: bad chunking");
  D3R_CHECK_ARG(d->obs && d->logd && d->logd_m && d->logd_v && d->small && d->small_m && d->small_v &&
                    d->small_trainable && d->workspace && d->sched && d->loss_out && d->counters,
                "d3r_align: null buffer");
  D3R_CHECK_ARG(d->chunk_px > 0 && d->chunk_px <= kChunk, "d3r_align: chunk_px=%d must be in [1, %d]", d->chunk_px, kChunk);
  return D3R_OK;
}

extern "C" int d3r_align_prepare(const d3r_align_desc* desc, void* stream) {
  int rc = validate(desc);
  if (rc) return rc;
  prepare_kernel<<<1, kThreads, 0, (cudaStream_t)stream>>>(*desc);
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}

template <bool kL2, int PPT>
static int launch_iters(const d3r_align_desc* desc, int it_begin, int it_end, cudaStream_t st) {
  const size_t smem = size_t(kStages) * PPT * kThreads * sizeof(float4) + size_t(kEntTile) * kWarps * kEntVals * sizeof(float);
  static bool attr = false;
  if (!attr) {
    D3R_CUDA(cudaFuncSetAttribute(align_iter_kernel<kL2, PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // two CTAs per SM need the maximum shared-memory carve-out
    D3R_CUDA(cudaFuncSetAttribute(align_iter_kernel<kL2, PPT>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)desc->n_chunks);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;   // overlap a launch's prologue with its predecessor's tail
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  for (int it = it_begin; it < it_end; ++it) D3R_CUDA(cudaLaunchKernelEx(&cfg, align_iter_kernel<kL2, PPT>, *desc, it));
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}

template <bool kL2>
static int launch_ppt(const d3r_align_desc* desc, int it_begin, int it_end, cudaStream_t st) {
  const int ppt = (desc->chunk_px + kThreads - 1) / kThreads;   // pixel slots per thread this problem needs
  if (ppt <= 4) return launch_iters<kL2, 4>(desc, it_begin, it_end, st);
  if (ppt == 5) return launch_iters<kL2, 5>(desc, it_begin, it_end, st);
  if (ppt == 6) return launch_iters<kL2, 6>(desc, it_begin, it_end, st);
  if (ppt == 7) return launch_iters<kL2, 7>(desc, it_begin, it_end, st);
  return launch_iters<kL2, 8>(desc, it_begin, it_end, st);
}

extern "C" int d3r_align_run(const d3r_align_desc* desc, int32_t it_begin, int32_t it_end, void* stream) {
  int rc = validate(desc);
  if (rc) return rc;
  D3R_CHECK_ARG(it_begin >= 0 && it_end >= it_begin, "d3r_align_run: bad iteration range");
  prof::Scope scope("align_iter", (cudaStream_t)stream, 0.0, 0.0, it_end - it_begin);
  return desc->dist_l2 ? launch_ppt<true>(desc, it_begin, it_end, (cudaStream_t)stream)
                       : launch_ppt<false>(desc, it_begin, it_end, (cudaStream_t)stream);
}

/* 1 if a fixed-point accumulator overflowed (|partial sum| >= 2^18) since the flag was last cleared. */
extern "C" int d3r_align_overflow_flag(const d3r_align_desc* desc, int32_t* host_out, void* stream) {
  int rc = validate(desc);
  if (rc) return rc;
  const Workspace ws = carve(desc->workspace, desc->n_imgs, desc->n_edges);
  D3R_CUDA(cudaMemcpyAsync(host_out, ws.flags, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  D3R_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  return D3R_OK;
}

extern "C" int d3r_align_pts3d(const d3r_align_desc* desc, float* out_dev, void* stream) {
  int rc = validate(desc);
  if (rc) return rc;
  D3R_CHECK_ARG(out_dev != nullptr, "d3r_align_pts3d: null output");
  pts3d_kernel<<<desc->n_chunks, kThreads, 0, (cudaStream_t)stream>>>(*desc, out_dev);
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}

extern "C" int d3r_align_pack_obs(const float* pts_dev, const float* weight_dev, void* obs_dev, int64_t obs_off,
                                  int64_t n_pix, void* stream) {
  D3R_CHECK_ARG(pts_dev && weight_dev && obs_dev && n_pix >= 0, "d3r_align_pack_obs: bad arguments");
  if (n_pix == 0) return D3R_OK;
  const int threads = 256;
  const int64_t blocks = (n_pix + threads - 1) / threads;
  pack_obs_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(pts_dev, weight_dev,
                                                                        reinterpret_cast<float4*>(obs_dev) + obs_off, n_pix);
  D3R_LAUNCH_CHECK();
  return D3R_OK;
}
