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
  float* edgeT;     // [E][12]
  float* imgT;      // [n][16]
  float* ent_part;  // [2E][max_chunks][13]
  float* img_part;  // [n_chunks][12]
  float* ent_sum;   // [2E][13]
  float* img_sum;   // [n][12]
};

__host__ __device__ inline int64_t align4(int64_t x) { return (x + 3) & ~int64_t(3); }

__host__ __device__ inline Workspace carve(float* ws, int n, int E, int n_chunks, int max_chunks) {
  Workspace w;
  int64_t o = 0;
  w.edgeT = ws + o;    o += align4(int64_t(E) * kEdgeT);
  w.imgT = ws + o;     o += align4(int64_t(n) * kImgT);
  w.ent_part = ws + o; o += align4(int64_t(2) * E * max_chunks * kEntVals);
  w.img_part = ws + o; o += align4(int64_t(n_chunks) * kImgVals);
  w.ent_sum = ws + o;  o += align4(int64_t(2) * E * kEntVals);
  w.img_sum = ws + o;  o += align4(int64_t(n) * kImgVals);
  return w;
}

inline int64_t workspace_floats(int n, int E, int n_chunks, int max_chunks) {
  return align4(int64_t(E) * kEdgeT) + align4(int64_t(n) * kImgT) +
         align4(int64_t(2) * E * max_chunks * kEntVals) + align4(int64_t(n_chunks) * kImgVals) +
         align4(int64_t(2) * E * kEntVals) + align4(int64_t(n) * kImgVals);
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

// ---- derived transforms (run by one CTA) -------------------------------------------------------
__device__ void compute_transforms(const d3r_align_desc& D, const Workspace& ws, float* s_red) {
  const int n = D.n_imgs, E = D.n_edges;
  const SmallLayout L(n, E);
  const float* sm = D.small;
  // mean of the pairwise log-scales (base_opt.py:178-184)
  float part = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) part += sm[L.pw + e * 8 + 7];
  part = warp_sum(part);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    s_red[32] = t / float(E);
  }
  __syncthreads();
  const float mean_sigma = s_red[32];
  const float log_base = logf(D.base_scale);
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float* p = sm + L.pw + e * 8;
    float R[9];
    quat_to_R(p, R, nullptr, nullptr);
    float s = expf(p[7]);
    if (D.norm_pw_scale) s = s * expf(log_base - mean_sigma);
    float a0 = sm[L.adapt + e * 2 + 0], a1 = sm[L.adapt + e * 2 + 1];
    float ad[3] = {a0, a0, a1};
    if (D.norm_pw_scale) {
      float mu = (a0 + a0 + a1) / 3.f;
      ad[0] -= mu; ad[1] -= mu; ad[2] -= mu;
    }
    for (int b = 0; b < 3; ++b) ad[b] = expf(ad[b] / D.pw_break);
    float* o = ws.edgeT + e * kEdgeT;
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b) o[a * 3 + b] = s * R[a * 3 + b] * ad[b];
    for (int a = 0; a < 3; ++a) o[9 + a] = s * signed_expm1f(p[4 + a]);
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float* p = sm + L.poses + i * 7;
    float* o = ws.imgT + i * kImgT;
    quat_to_R(p, o, nullptr, nullptr);
    for (int a = 0; a < 3; ++a) o[9 + a] = signed_expm1f(p[4 + a]);
    float fx = expf(sm[L.focals + i * 2 + 0] / D.focal_break);
    float fy = expf(sm[L.focals + i * 2 + 1] / D.focal_break);
    o[12] = 1.f / fx;
    o[13] = 1.f / fy;
    o[14] = 0.5f * float(D.img_hw[i * 2 + 1]) + 10.f * sm[L.pp + i * 2 + 0];
    o[15] = 0.5f * float(D.img_hw[i * 2 + 0]) + 10.f * sm[L.pp + i * 2 + 1];
  }
}

__global__ void __launch_bounds__(kThreads) prepare_kernel(const __grid_constant__ d3r_align_desc D) {
  __shared__ float s_red[40];
  Workspace ws = carve(D.workspace, D.n_imgs, D.n_edges, D.n_chunks, D.max_chunks);
  compute_transforms(D, ws, s_red);
}

// ---- small-parameter backward + Adam (run by the last CTA of the grid) -------------------------
__device__ void small_param_step(const d3r_align_desc& D, const Workspace& ws, int it, float* s_red) {
  const int n = D.n_imgs, E = D.n_edges;
  const SmallLayout L(n, E);
  float* sm = D.small;
  float* am = D.small_m;
  float* av = D.small_v;
  const uint8_t* tr = D.small_trainable;
  const float step_size = D.sched[it * 4 + 1], bc2s = D.sched[it * 4 + 2];
  const float b1 = D.beta1, b2 = D.beta2, eps = D.adam_eps;

  // loss = sum over entries (coefficients already folded in), fixed order
  float lpart = 0.f;
  for (int k = threadIdx.x; k < 2 * E; k += blockDim.x) lpart += __ldcg(ws.ent_sum + k * kEntVals + 12);
  lpart = warp_sum(lpart);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = lpart;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    D.loss_out[it] = t;
  }
  __syncthreads();
  if (D.eval_only) return;

  // pass 1 over edges: dL/ds_e * s_e summed (the mean-coupling term of get_pw_norm_scale_factor)
  float mean_sigma = 0.f;
  {
    float part = 0.f;
    for (int e = threadIdx.x; e < E; e += blockDim.x) part += sm[L.pw + e * 8 + 7];
    part = warp_sum(part);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
      s_red[32] = t / float(E);
    }
    __syncthreads();
    mean_sigma = s_red[32];
    __syncthreads();
  }
  const float log_base = logf(D.base_scale);

  // per-edge gradient of (sigma_e) before the coupling, kept in registers across the two passes
  // (each thread revisits the same edges in the same order)
  float coupl = 0.f;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float* p = sm + L.pw + e * 8;
    float R[9];
    quat_to_R(p, R, nullptr, nullptr);
    float s = expf(p[7]);
    if (D.norm_pw_scale) s = s * expf(log_base - mean_sigma);
    float a0 = sm[L.adapt + e * 2 + 0], a1 = sm[L.adapt + e * 2 + 1];
    float ad[3] = {a0, a0, a1};
    if (D.norm_pw_scale) { float mu = (a0 + a0 + a1) / 3.f; ad[0] -= mu; ad[1] -= mu; ad[2] -= mu; }
    for (int b = 0; b < 3; ++b) ad[b] = expf(ad[b] / D.pw_break);
    const float* si = ws.ent_sum + D.edge_ent[e * 2 + 0] * kEntVals;
    const float* sj = ws.ent_sum + D.edge_ent[e * 2 + 1] * kEntVals;
    float T[3];
    for (int a = 0; a < 3; ++a) T[a] = signed_expm1f(p[4 + a]);
    // dL/dM_ab = -sum g_a q_b ; dL/dt'_a = -sum g_a
    float dLds = 0.f;
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) dLds -= (__ldcg(si + a * 3 + b) + __ldcg(sj + a * 3 + b)) * R[a * 3 + b] * ad[b];
      dLds -= (__ldcg(si + 9 + a) + __ldcg(sj + 9 + a)) * T[a];
    }
    coupl += dLds * s;
  }
  coupl = warp_sum(coupl);
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = coupl;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_red[w];
    s_red[33] = t / float(E);
  }
  __syncthreads();
  const float coupling = D.norm_pw_scale ? s_red[33] : 0.f;

  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float* p = sm + L.pw + e * 8;
    float R[9], qh[4], qn;
    quat_to_R(p, R, qh, &qn);
    float s = expf(p[7]);
    if (D.norm_pw_scale) s = s * expf(log_base - mean_sigma);
    float a0 = sm[L.adapt + e * 2 + 0], a1 = sm[L.adapt + e * 2 + 1];
    float ad[3] = {a0, a0, a1};
    if (D.norm_pw_scale) { float mu = (a0 + a0 + a1) / 3.f; ad[0] -= mu; ad[1] -= mu; ad[2] -= mu; }
    for (int b = 0; b < 3; ++b) ad[b] = expf(ad[b] / D.pw_break);
    const float* si = ws.ent_sum + D.edge_ent[e * 2 + 0] * kEntVals;
    const float* sj = ws.ent_sum + D.edge_ent[e * 2 + 1] * kEntVals;
    float dM[9], dt[3], T[3];
    for (int k = 0; k < 9; ++k) dM[k] = -(__ldcg(si + k) + __ldcg(sj + k));
    for (int a = 0; a < 3; ++a) { dt[a] = -(__ldcg(si + 9 + a) + __ldcg(sj + 9 + a)); T[a] = signed_expm1f(p[4 + a]); }
    float dLds = 0.f, dR[9], dad[3] = {0.f, 0.f, 0.f};
    for (int a = 0; a < 3; ++a) {
      for (int b = 0; b < 3; ++b) {
        dLds += dM[a * 3 + b] * R[a * 3 + b] * ad[b];
        dR[a * 3 + b] = dM[a * 3 + b] * s * ad[b];
        dad[b] += dM[a * 3 + b] * s * R[a * 3 + b];
      }
      dLds += dt[a] * T[a];
    }
    float g[8];
    quat_backward(dR, qh, qn, g);
    for (int a = 0; a < 3; ++a) {
      float t = p[4 + a];
      float sg = (t > 0.f) - (t < 0.f);
      g[4 + a] = dt[a] * s * sg * sg * expf(fabsf(t));
    }
    g[7] = dLds * s - coupling;
    // adaptors (base_opt.py:143-148): adapt3 = exp((cat(a0,a0,a1) - mean)/pw_break)
    float gad[3];
    for (int b = 0; b < 3; ++b) gad[b] = dad[b] * ad[b] / D.pw_break;
    if (D.norm_pw_scale) { float mu = (gad[0] + gad[1] + gad[2]) / 3.f; gad[0] -= mu; gad[1] -= mu; gad[2] -= mu; }
    float ga[2] = {gad[0] + gad[1], gad[2]};
    for (int k = 0; k < 8; ++k) {
      int idx = L.pw + e * 8 + k;
      if (tr[idx]) sm[idx] = adam_update(sm[idx], g[k], am[idx], av[idx], b1, b2, step_size, bc2s, eps);
    }
    for (int k = 0; k < 2; ++k) {
      int idx = L.adapt + e * 2 + k;
      if (tr[idx]) sm[idx] = adam_update(sm[idx], ga[k], am[idx], av[idx], b1, b2, step_size, bc2s, eps);
    }
  }

  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    float* p = sm + L.poses + i * 7;
    float R[9], qh[4], qn;
    quat_to_R(p, R, qh, &qn);
    const float* S = ws.img_sum + i * kImgVals;  // S[a*3+b] = sum G_a c_b ; S[9+a] = sum G_a
    float dR[9];
    for (int k = 0; k < 9; ++k) dR[k] = __ldcg(S + k);
    float g[7];
    quat_backward(dR, qh, qn, g);
    for (int a = 0; a < 3; ++a) {
      float t = p[4 + a];
      float sg = (t > 0.f) - (t < 0.f);
      g[4 + a] = __ldcg(S + 9 + a) * sg * sg * expf(fabsf(t));
    }
    for (int k = 0; k < 7; ++k) {
      int idx = L.poses + i * 7 + k;
      if (tr[idx]) sm[idx] = adam_update(sm[idx], g[k], am[idx], av[idx], b1, b2, step_size, bc2s, eps);
    }
    // focals: c_x = d (u-cx)/fx, fx = exp(phi/focal_break)
    float gfx = 0.f, gfy = 0.f, gpx = 0.f, gpy = 0.f;
    for (int a = 0; a < 3; ++a) {
      gfx += R[a * 3 + 0] * dR[a * 3 + 0];
      gfy += R[a * 3 + 1] * dR[a * 3 + 1];
      gpx += R[a * 3 + 0] * dR[a * 3 + 2];
      gpy += R[a * 3 + 1] * dR[a * 3 + 2];
    }
    gfx = -gfx / D.focal_break;
    gfy = -gfy / D.focal_break;
    float fx = expf(sm[L.focals + i * 2 + 0] / D.focal_break);
    float fy = expf(sm[L.focals + i * 2 + 1] / D.focal_break);
    gpx = -10.f * gpx / fx;
    gpy = -10.f * gpy / fy;
    if (D.tied_focal) {
      int idx = L.focals + i * 2;
      if (tr[idx]) {
        float v = adam_update(sm[idx], gfx + gfy, am[idx], av[idx], b1, b2, step_size, bc2s, eps);
        sm[idx] = v; sm[idx + 1] = v;
      }
    } else {
      int idx = L.focals + i * 2;
      if (tr[idx]) sm[idx] = adam_update(sm[idx], gfx, am[idx], av[idx], b1, b2, step_size, bc2s, eps);
      if (tr[idx + 1]) sm[idx + 1] = adam_update(sm[idx + 1], gfy, am[idx + 1], av[idx + 1], b1, b2, step_size, bc2s, eps);
    }
    {
      int idx = L.pp + i * 2;
      if (tr[idx]) sm[idx] = adam_update(sm[idx], gpx, am[idx], av[idx], b1, b2, step_size, bc2s, eps);
      if (tr[idx + 1]) sm[idx + 1] = adam_update(sm[idx + 1], gpy, am[idx + 1], av[idx + 1], b1, b2, step_size, bc2s, eps);
    }
  }
  __threadfence_block();
  __syncthreads();
  compute_transforms(D, ws, s_red);
}

// ---- the per-iteration kernel ---------------------------------------------------------------
// 8 warps, all computing; lane 0 of warp 0 doubles as the producer: each incident entry's float4 observations
// for this CTA's pixel chunk are streamed into a 3-stage shared-memory ring with cp.async.bulk (TMA 1-D copies,
// mbarrier completion), refilled as soon as a stage is drained, so up to 2 x 84 KB of reads are in flight per
// SM independent of the warps' compute progress.
constexpr int kStages = 3;
constexpr int kEntTile = 16;               // entries whose per-warp partial sums are staged in smem at a time
constexpr int kThreadsIter = kThreads;      // 8 warps (warp allocation granularity is 4: a 9th warp would cost a whole CTA/SM)
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

template <bool kL2>
__global__ void __launch_bounds__(kThreadsIter, 2)
align_iter_kernel(const __grid_constant__ d3r_align_desc D, int it) {
  extern __shared__ __align__(128) uint8_t s_dyn[];
  float4* s_obs = reinterpret_cast<float4*>(s_dyn);                                  // [kStages][kChunk]
  float* s_ent = reinterpret_cast<float*>(s_dyn + kStages * kChunk * sizeof(float4)); // [kEntTile][kWarps][13]
  __shared__ float s_img[kWarps * kImgVals];
  __shared__ float s_red[40];
  __shared__ int s_flag;
  __shared__ __align__(8) uint64_t s_full[kStages], s_empty[kStages];

  const Workspace ws = carve(D.workspace, D.n_imgs, D.n_edges, D.n_chunks, D.max_chunks);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunk = blockIdx.x;
  const int img = D.chunk_img[chunk];
  const int lc = chunk - D.img_chunk_ptr[img];
  const int H = D.img_hw[img * 2 + 0], W = D.img_hw[img * 2 + 1];
  const int P = H * W;
  const int pbase = lc * D.chunk_px;                       // first pixel of this CTA
  const int npx = min(D.chunk_px, P - pbase);              // >= 1
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
  __syncthreads();

  const uint32_t stage_bytes = uint32_t(npx) * 16u;
  auto produce = [&](int k) {   // called by one thread
    const int s = k % kStages;
    mbar_wait(smem_u32(&s_empty[s]), ((k / kStages) & 1) ^ 1);
    mbar_expect_tx(smem_u32(&s_full[s]), stage_bytes);
    bulk_g2s(smem_u32(s_obs + s * kChunk), obs_base + D.ent_obs_off[e0 + k] + pbase, stage_bytes, smem_u32(&s_full[s]));
  };
  if (tid == 0)
    for (int k = 0; k < min(kStages, deg); ++k) produce(k);
  {
    // ================= compute warps =================
    const float* iT = ws.imgT + img * kImgT;
    float R[9], T[3];
#pragma unroll
    for (int k = 0; k < 9; ++k) R[k] = iT[k];
#pragma unroll
    for (int k = 0; k < 3; ++k) T[k] = iT[9 + k];
    const float ifx = iT[12], ify = iT[13], cx = iT[14], cy = iT[15];

    float X[kPPT][3], G[kPPT][3];
#pragma unroll
    for (int k = 0; k < kPPT; ++k) {
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

    for (int k0 = 0; k0 < deg; k0 += kEntTile) {
      const int kend = min(deg, k0 + kEntTile);
      for (int kk = k0; kk < kend; ++kk) {
        const int ent = e0 + kk;
        const float* eT = ws.edgeT + D.ent_edge[ent] * kEdgeT;
        float M[9], t[3];
#pragma unroll
        for (int k = 0; k < 9; ++k) M[k] = eT[k];
#pragma unroll
        for (int k = 0; k < 3; ++k) t[k] = eT[9 + k];
        const float coef = D.ent_coef[ent];
        const int s = kk % kStages;
        mbar_wait(smem_u32(&s_full[s]), (kk / kStages) & 1);
        const float4* so = s_obs + s * kChunk;
        float acc[kRedVals];
#pragma unroll
        for (int k = 0; k < kRedVals; ++k) acc[k] = 0.f;
#pragma unroll
        for (int k = 0; k < kPPT; ++k) {
          const int q = tid + k * kThreads;
          float4 o = so[q];
          if (q >= npx) o = make_float4(0.f, 0.f, 0.f, 0.f);   // stale smem beyond the chunk: force weight 0
          const float qx = o.x, qy = o.y, qz = o.z;
          const float lw = coef * o.w;
          const float r0 = X[k][0] - (M[0] * qx + M[1] * qy + M[2] * qz + t[0]);
          const float r1 = X[k][1] - (M[3] * qx + M[4] * qy + M[5] * qz + t[1]);
          const float r2 = X[k][2] - (M[6] * qx + M[7] * qy + M[8] * qz + t[2]);
          const float rho2 = r0 * r0 + r1 * r1 + r2 * r2;
          float gs;
          if (kL2) {
            acc[12] += lw * rho2;
            gs = 2.f * lw;
          } else {
            // torch's norm backward yields 0 at ||r|| == 0
            const float inv = rho2 > 0.f ? rsqrtf(rho2) : 0.f;
            acc[12] += lw * (rho2 * inv);
            gs = lw * inv;
          }
          const float g0 = gs * r0, g1 = gs * r1, g2 = gs * r2;
          G[k][0] += g0; G[k][1] += g1; G[k][2] += g2;
          acc[0] += g0 * qx; acc[1] += g0 * qy; acc[2] += g0 * qz;
          acc[3] += g1 * qx; acc[4] += g1 * qy; acc[5] += g1 * qz;
          acc[6] += g2 * qx; acc[7] += g2 * qy; acc[8] += g2 * qz;
          acc[9] += g0; acc[10] += g1; acc[11] += g2;
        }
        // this warp is done with the stage: hand it back to the producer
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_empty[s]));
        if (tid == 0 && kk + kStages < deg) produce(kk + kStages);   // refill the stage just drained
        const float tot = butterfly16(acc, lane);
        const int vi = (lane >> 1) & 15;
        if (!(lane & 1) && vi < kEntVals) s_ent[((kk - k0) * kWarps + warp) * kEntVals + vi] = tot;
      }
      // cross-warp sums of this tile of entries -> partial rows in global memory (fixed order)
      asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
      for (int idx = tid; idx < (kend - k0) * kEntVals; idx += kThreads) {
        const int k = idx / kEntVals, v = idx - k * kEntVals;
        float sacc = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) sacc += s_ent[(k * kWarps + w) * kEntVals + v];
        __stcg(ws.ent_part + (int64_t(e0 + k0 + k) * D.max_chunks + lc) * kEntVals + v, sacc);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
    }

    // depth gradient + Adam (in place), per-image pose/focal sums
    float S[kRedVals];
#pragma unroll
    for (int k = 0; k < kRedVals; ++k) S[k] = 0.f;
    if (!D.eval_only) {
      const float step_size = D.sched[it * 4 + 1], bc2s = D.sched[it * 4 + 2];
#pragma unroll
      for (int k = 0; k < kPPT; ++k) {
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
  }
  __syncthreads();
  if (tid < kImgVals) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += s_img[w * kImgVals + tid];
    __stcg(ws.img_part + int64_t(chunk) * kImgVals + tid, s);
  }

  // ---- ticket 1: last CTA of this image reduces the image's partial rows ----
  __threadfence();
  __syncthreads();
  const int nchunk_img = D.img_chunk_ptr[img + 1] - D.img_chunk_ptr[img];
  if (tid == 0) s_flag = (atomicAdd(D.counters + img, 1) == nchunk_img - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  // one warp per (entry, value) row: lanes take chunks round-robin (independent loads in flight), then a fixed
  // shuffle tree -> deterministic and ~100x shorter than a serial walk over the chunks
  for (int task = warp; task < deg * kEntVals + kImgVals; task += kWarps) {
    const float* src;
    float* dst;
    int stride;
    if (task < deg * kEntVals) {
      const int k = task / kEntVals, v = task - k * kEntVals;
      src = ws.ent_part + int64_t(e0 + k) * D.max_chunks * kEntVals + v;
      dst = ws.ent_sum + (e0 + k) * kEntVals + v;
      stride = kEntVals;
    } else {
      const int v = task - deg * kEntVals;
      src = ws.img_part + int64_t(D.img_chunk_ptr[img]) * kImgVals + v;
      dst = ws.img_sum + img * kImgVals + v;
      stride = kImgVals;
    }
    float sacc = 0.f;
    for (int c0 = 0; c0 < nchunk_img; c0 += 128) {   // 4 independent loads per lane in flight
      const int ca = c0 + lane, cb = ca + 32, cc = ca + 64, cd = ca + 96;
      const float va = ca < nchunk_img ? __ldcg(src + ca * stride) : 0.f;
      const float vb = cb < nchunk_img ? __ldcg(src + cb * stride) : 0.f;
      const float vc = cc < nchunk_img ? __ldcg(src + cc * stride) : 0.f;
      const float vd = cd < nchunk_img ? __ldcg(src + cd * stride) : 0.f;
      sacc += (va + vb) + (vc + vd);
    }
    sacc = warp_sum(sacc);
    if (lane == 0) __stcg(dst, sacc);
  }
  if (tid == 0) D.counters[img] = 0;  // re-arm for the next launch

  // ---- ticket 2: last image finisher updates the small parameters ----
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(D.counters + D.n_imgs, 1) == D.n_imgs - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  if (tid == 0) D.counters[D.n_imgs] = 0;
  small_param_step(D, ws, it, s_red);
}

__global__ void __launch_bounds__(kThreads) pts3d_kernel(const __grid_constant__ d3r_align_desc D, float* out) {
  const Workspace ws = carve(D.workspace, D.n_imgs, D.n_edges, D.n_chunks, D.max_chunks);
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

extern "C" int d3r_align_chunk_pixels(void) { return kChunk; }
extern "C" int d3r_sizeof_align_desc(void) { return (int)sizeof(d3r_align_desc); }

extern "C" int64_t d3r_align_workspace_floats(int32_t n_imgs, int32_t n_edges, int32_t n_chunks, int32_t max_chunks) {
  return workspace_floats(n_imgs, n_edges, n_chunks, max_chunks);
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

extern "C" int d3r_align_run(const d3r_align_desc* desc, int32_t it_begin, int32_t it_end, void* stream) {
  int rc = validate(desc);
  if (rc) return rc;
  D3R_CHECK_ARG(it_begin >= 0 && it_end >= it_begin, "d3r_align_run: bad iteration range");
  const size_t smem = size_t(kStages) * kChunk * sizeof(float4) + size_t(kEntTile) * kWarps * kEntVals * sizeof(float);
  if (desc->dist_l2) {
    D3R_CUDA(cudaFuncSetAttribute(align_iter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    D3R_CUDA(cudaFuncSetAttribute(align_iter_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  } else {
    D3R_CUDA(cudaFuncSetAttribute(align_iter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    D3R_CUDA(cudaFuncSetAttribute(align_iter_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  }
  prof::Scope scope("align_iter", (cudaStream_t)stream, 0.0, 0.0, it_end - it_begin);
  for (int it = it_begin; it < it_end; ++it) {
    if (desc->dist_l2)
      align_iter_kernel<true><<<desc->n_chunks, kThreadsIter, smem, (cudaStream_t)stream>>>(*desc, it);
    else
      align_iter_kernel<false><<<desc->n_chunks, kThreadsIter, smem, (cudaStream_t)stream>>>(*desc, it);
  }
  D3R_LAUNCH_CHECK();
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
