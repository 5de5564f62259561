// Triplet attention core, tensor-core families: shared host side (sm_100a, bf16 / fp16, head dim 16, N <= 64).
//
// Replaces lib/tgt/layers/triplet.py:213-227 and 232-246 (reference): per (graph b, head h, direction) an
// N x N x d attention for every junction atom j, with a bias/gate/mask tile that does NOT depend on j.
// This file holds what the kernel families share:
//   * the prep kernel that transposes the [R, H] bias / gate columns of the projection into [b,dir,h,i,k] tiles
//     (adding the mask, applying the sigmoid) and the post kernel that scatters dE / dG back; both live in the
//     caller-provided workspace, and the forward's tiles are reused by the backward;
//   * workspace sizing and the dispatch between the tcgen05 / TMEM family (triplet_tc.cu), the TMA-staged mma.sync family
//     (triplet_tma.cu) and the fused projection + attention forward (triplet_fused.cu) by kernel policy.
// The cp.async-staged mma.sync kernels that lived here in round 1 were removed (slower than both remaining families).
#include "triplet_common.cuh"
#include <stdlib.h>

namespace tgt {

// ------------------------------------------------------------------------------------------------ prep / post
// ws_e : [B,2,H,64,64] f32   (E + mask) * log2(e), clamped to >= NEG_BIG (masked keys stay finite, so a fully
//        masked row is a uniform softmax over its N keys exactly like the reference); tile padding = -inf
// ws_g : [B,2,H,64,64] f16   sigmoid(G + mask) (1 when ungated); padding = 0
// tile index (i = query, k = key); the projection row holding entry (i,k) is (b,i,k) inward, (b,k,i) outward.
template <typename T>
__global__ void __launch_bounds__(256)
tri_prep_bias_gate(const tgt_triplet_attn_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                   float *__restrict__ ws_e, __half *__restrict__ ws_g) {
  extern __shared__ float sm[];                 // [H][65] x 2
  const int N = D.N, H = D.H;
  const int i = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  float *se = sm, *sg = sm + H * 65;
  const int off_e = D.off_e[dir], off_g = D.off_g[dir];
  for (int idx = threadIdx.x; idx < TN * H; idx += blockDim.x) {
    const int k = idx / H, h = idx - k * H;
    float ev = -INFINITY, gv = 0.f;      // tile padding (i or k >= N): excluded from every softmax
    if (i < N && k < N) {
      const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
      const float mk = mask[brow];
      float e = mk;
      if (off_e >= 0) e += to_f(proj[brow * D.ld + off_e + h]);
      ev = fmaxf(e * LOG2E, NEG_BIG);
      gv = off_g >= 0 ? 1.f / (1.f + __expf(-(to_f(proj[brow * D.ld + off_g + h]) + mk))) : 1.f;
    }
    se[h * 65 + k] = ev;
    sg[h * 65 + k] = gv;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TN * H; idx += blockDim.x) {
    const int h = idx / TN, k = idx - h * TN;
    const int64_t o = ((((int64_t)(b * 2 + dir) * H + h) * TN) + i) * TN + k;
    ws_e[o] = se[h * 65 + k];
    ws_g[o] = __float2half_rn(sg[h * 65 + k]);
  }
}

// scatter dE / dG tiles ([B,2,H,64,64] f32) back into the projection-gradient columns
template <typename T>
__global__ void __launch_bounds__(256)
tri_post_bias_gate(const tgt_triplet_attn_desc D, const float *__restrict__ ws_de, const float *__restrict__ ws_dg,
                   T *__restrict__ dproj) {
  extern __shared__ float sm[];
  const int N = D.N, H = D.H;
  const int i = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  float *se = sm, *sg = sm + H * 65;
  const int off_e = D.off_e[dir], off_g = D.off_g[dir];
  if (off_e < 0 && off_g < 0) return;
  for (int idx = threadIdx.x; idx < TN * H; idx += blockDim.x) {
    const int h = idx / TN, k = idx - h * TN;
    const int64_t o = ((((int64_t)(b * 2 + dir) * H + h) * TN) + i) * TN + k;
    se[h * 65 + k] = ws_de[o];
    sg[h * 65 + k] = ws_dg[o];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < TN * H; idx += blockDim.x) {
    const int k = idx / H, h = idx - k * H;
    if (k >= N) continue;
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    if (off_e >= 0) dproj[brow * D.ld + off_e + h] = from_f<T>(se[h * 65 + k]);
    if (off_g >= 0) dproj[brow * D.ld + off_g + h] = from_f<T>(sg[h * 65 + k]);
  }
}

// ---- H = 16 fast paths of the two kernels above: the E (and G) values of one edge row are 16 contiguous 16-bit channels =
// two 16-byte vectors, so a thread moves 8 heads per load / store instead of 2 bytes, and the tile side moves float4 / uint4.
// Thread t of 256: row k = t / 4, part = t % 4: parts 0,1 = heads 0-7 / 8-15 of E, parts 2,3 = the same of G.
template <typename T>
__global__ void __launch_bounds__(256)
tri_prep_bias_gate_h16(const tgt_triplet_attn_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                       float *__restrict__ ws_e, __half *__restrict__ ws_g) {
  __shared__ float se[16 * 65], sg[16 * 65];
  const int N = D.N;
  const int i = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int off_e = D.off_e[dir], off_g = D.off_g[dir];
  const int k = threadIdx.x >> 2, part = threadIdx.x & 3;
  const bool is_g = part >= 2;
  const int h0 = (part & 1) * 8;
  float v[8];
  if (i < N && k < N) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    const float mk = mask[brow];
    const int off = is_g ? off_g : off_e;
    float x[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) x[q] = 0.f;
    if (off >= 0) load_vec<T, 8>(proj + brow * D.ld + off + h0, x);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (is_g) v[q] = off_g >= 0 ? 1.f / (1.f + __expf(-(x[q] + mk))) : 1.f;
      else v[q] = fmaxf((x[q] + mk) * LOG2E, NEG_BIG);
    }
  } else {                              // tile padding (i or k >= N): excluded from every softmax
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = is_g ? 0.f : -INFINITY;
  }
  float *dst = is_g ? sg : se;
#pragma unroll
  for (int q = 0; q < 8; ++q) dst[(h0 + q) * 65 + k] = v[q];
  __syncthreads();
  // tile rows: (h, i, :) = 64 fp32 (256 B) / 64 halfs (128 B)
  const int64_t base = (((int64_t)(b * 2 + dir) * 16) * TN + i) * TN;        // + h * TN * TN + k
  {
    const int h = threadIdx.x >> 4, k4 = (threadIdx.x & 15) * 4;
    const float *r = se + h * 65 + k4;
    *reinterpret_cast<float4 *>(ws_e + base + (int64_t)h * TN * TN + k4) = make_float4(r[0], r[1], r[2], r[3]);
  }
  if (threadIdx.x < 128) {
    const int h = threadIdx.x >> 3, k8 = (threadIdx.x & 7) * 8;
    const float *r = sg + h * 65 + k8;
    __half2 p[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) p[q] = __floats2half2_rn(r[2 * q], r[2 * q + 1]);
    *reinterpret_cast<uint4 *>(ws_g + base + (int64_t)h * TN * TN + k8) = *reinterpret_cast<uint4 *>(p);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
tri_post_bias_gate_h16(const tgt_triplet_attn_desc D, const float *__restrict__ ws_de, const float *__restrict__ ws_dg,
                       T *__restrict__ dproj) {
  __shared__ float se[16 * 65], sg[16 * 65];
  const int N = D.N;
  const int i = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int off_e = D.off_e[dir], off_g = D.off_g[dir];
  if (off_e < 0 && off_g < 0) return;
  const int64_t base = (((int64_t)(b * 2 + dir) * 16) * TN + i) * TN;
  {
    const int h = threadIdx.x >> 4, k4 = (threadIdx.x & 15) * 4;
    const float4 e = *reinterpret_cast<const float4 *>(ws_de + base + (int64_t)h * TN * TN + k4);
    const float4 g = *reinterpret_cast<const float4 *>(ws_dg + base + (int64_t)h * TN * TN + k4);
    float *r = se + h * 65 + k4, *t = sg + h * 65 + k4;
    r[0] = e.x; r[1] = e.y; r[2] = e.z; r[3] = e.w;
    t[0] = g.x; t[1] = g.y; t[2] = g.z; t[3] = g.w;
  }
  __syncthreads();
  const int k = threadIdx.x >> 2, part = threadIdx.x & 3;
  const bool is_g = part >= 2;
  const int h0 = (part & 1) * 8;
  const int off = is_g ? off_g : off_e;
  if (k < N && off >= 0) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    const float *src = is_g ? sg : se;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = src[(h0 + q) * 65 + k];
    store_vec<T, 8>(dproj + brow * D.ld + off + h0, v);
  }
}

static bool bias_gate_h16_ok(const tgt_triplet_attn_desc &D) {
  if (D.H != 16 || (D.ld % 8)) return false;
  for (int dir = 0; dir < 2; ++dir)
    if ((D.off_e[dir] >= 0 && D.off_e[dir] % 8) || (D.off_g[dir] >= 0 && D.off_g[dir] % 8)) return false;
  return true;
}


// ------------------------------------------------------------------------------------------------ host
static size_t tile_elems(const tgt_triplet_attn_desc &D) { return (size_t)D.B * 2 * D.H * TN * TN; }

size_t triplet_attn_mma_workspace(const tgt_triplet_attn_desc &D, int backward) {
  const size_t n = tile_elems(D);
  size_t bytes = n * sizeof(float) + n * sizeof(__half);
  if (backward) bytes += 2 * n * sizeof(float);
  return bytes + 1024 + 256;          // + alignment slack + the work-item counter of the persistent tcgen05 backward
}

bool triplet_attn_mma_supported(const tgt_triplet_attn_desc &D) {
  if (D.dtype != TGT_BF16 && D.dtype != TGT_F16) return false;
  if (D.d != HD || D.N > TN) return false;
  if (D.ld % 8) return false;
  for (int dir = 0; dir < 2; ++dir)
    if ((D.off_q[dir] % 8) || (D.off_k[dir] % 8) || (D.off_v[dir] % 8)) return false;
  return true;
}

// triplet_tma.cu
bool triplet_attn_tma_available();
int triplet_attn_fwd_tma_launch(const tgt_triplet_attn_desc &D, const void *proj, void *va, float *stats,
                                const float *ws_e, const __half *ws_g, cudaStream_t st);
int triplet_attn_bwd_tma_launch(const tgt_triplet_attn_desc &D, const void *proj, const void *dva, const float *stats,
                                void *dproj, const float *ws_e, const __half *ws_g, float *ws_de, float *ws_dg, float *dbias,
                                cudaStream_t st);
bool triplet_attn_bwd_tma_has_bias();
// triplet_fused.cu
bool triplet_attn_fused_supported(const tgt_triplet_attn_desc &D, int We);
int triplet_attn_fused_launch(const tgt_triplet_attn_desc &D, int We, const void *x, int64_t ldx, const float *mean,
                              const float *rstd, const void *wf, const float *wcolsum, const float *wbias, void *va,
                              float *stats, const float *ws_e, const __half *ws_g, cudaStream_t st);
// triplet_tc.cu
bool triplet_attn_tc_supported(const tgt_triplet_attn_desc &D);
int triplet_attn_fwd_tc_launch(const tgt_triplet_attn_desc &D, const void *proj, void *va, float *va_f32, float *stats,
                               const float *ws_e, const __half *ws_g, cudaStream_t st);
int triplet_attn_bwd_tc_launch(const tgt_triplet_attn_desc &D, const void *proj, const void *va, const void *dva,
                               const float *stats, void *dproj, const float *ws_e, const __half *ws_g, float *ws_de,
                               float *ws_dg, int *counter, cudaStream_t st);
// every tensor-core policy stages with TMA (the cp.async-staged kernels of round 1, policy 2, were removed in round 2:
// slower than the TMA family and no longer needed as a cross-check now that there are two independent tensor-core families)
static bool use_tma() { return g_policy.load() != 1 && triplet_attn_tma_available(); }
// tcgen05 / TMEM core (triplet_tc.cu): policy 4 runs it forward and backward.  Under the default policy 0 the forward
// runs it too (1.01 ms against 1.06 ms for the TMA-staged mma.sync kernel at config 3; both write the same log-sum-exp
// statistics, so the backward family is independent); the tcgen05 backward (2.40 ms against 2.47 ms) becomes the
// default only together with a weight-gradient GEMM that supplies the projection's bias gradient, which the mma.sync
// backward produces as a by-product and the tcgen05 one does not (TGT_TRI_TC_BWD=1 forces it).
static bool tc_bwd_default() {
  static const int v = [] { const char *e = getenv("TGT_TRI_TC_BWD"); return e ? atoi(e) : 0; }();
  return v != 0;
}
static bool use_tc(const tgt_triplet_attn_desc &D, bool backward = false) {
  const int p = g_policy.load();
  const bool want = p == 4 || ((p == 0 || p == 3) && (!backward || tc_bwd_default()));
  return want && triplet_attn_tc_supported(D);
}

struct Ws {
  float *e; __half *g; float *de; float *dg;
  int *counter;
};
static Ws carve(const tgt_triplet_attn_desc &D, void *ws) {
  const size_t n = tile_elems(D);
  uintptr_t p = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  Ws w;
  w.e = (float *)p;
  w.de = w.e + n;
  w.dg = w.de + n;
  w.g = (__half *)(w.dg + n);
  w.counter = (int *)(((uintptr_t)(w.g + n) + 15) & ~(uintptr_t)15);
  return w;
}
static Ws carve_fwd(const tgt_triplet_attn_desc &D, void *ws) {
  const size_t n = tile_elems(D);
  uintptr_t p = ((uintptr_t)ws + 255) & ~(uintptr_t)255;
  Ws w;
  w.e = (float *)p;
  w.g = (__half *)(w.e + n);
  w.de = w.dg = nullptr;
  w.counter = nullptr;
  return w;
}

template <typename T>
static int fwd_impl(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, void *va, float *stats,
                    void *ws, cudaStream_t st) {
  const Ws w = carve_fwd(D, ws);
  const size_t psm = 2 * (size_t)D.H * 65 * sizeof(float);
  if (bias_gate_h16_ok(D)) tri_prep_bias_gate_h16<T><<<dim3(TN, 2, D.B), 256, 0, st>>>(D, (const T *)proj, mask, w.e, w.g);
  else tri_prep_bias_gate<T><<<dim3(TN, 2, D.B), 256, psm, st>>>(D, (const T *)proj, mask, w.e, w.g);
  if (int e = check_launch("tri_prep_bias_gate")) return e;
  if (use_tc(D)) return triplet_attn_fwd_tc_launch(D, proj, va, nullptr, stats, w.e, w.g, st);
  if (use_tma()) return triplet_attn_fwd_tma_launch(D, proj, va, stats, w.e, w.g, st);
  return fail("triplet_attn_fwd: cuTensorMapEncodeTiled is not available from the driver (TMA is required)");
}

template <typename T>
static int bwd_impl(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, const void *va, const void *dva,
                    const float *stats, void *dproj, void *ws, const void *fwd_ws, float *dbias, cudaStream_t st) {
  if (dbias && (use_tc(D, true) || !(use_tma() && triplet_attn_bwd_tma_has_bias())))
    return fail("triplet_attn_bwd: the projection-bias by-product is not available with this kernel family");
  Ws w = carve(D, ws);
  const size_t psm = 2 * (size_t)D.H * 65 * sizeof(float);
  if (fwd_ws) {
    // bias / gate tiles of the forward call, kept alive by the caller: identical to what prep would recompute
    const Ws wf = carve_fwd(D, const_cast<void *>(fwd_ws));
    w.e = wf.e;
    w.g = wf.g;
  } else {
    if (bias_gate_h16_ok(D)) tri_prep_bias_gate_h16<T><<<dim3(TN, 2, D.B), 256, 0, st>>>(D, (const T *)proj, mask, w.e, w.g);
    else tri_prep_bias_gate<T><<<dim3(TN, 2, D.B), 256, psm, st>>>(D, (const T *)proj, mask, w.e, w.g);
    if (int e = check_launch("tri_prep_bias_gate")) return e;
  }
  if (use_tc(D, true)) {
    if (dbias) return fail("triplet_attn_bwd: the projection-bias by-product is not produced by the tcgen05 kernel");
    if (int e = triplet_attn_bwd_tc_launch(D, proj, va, dva, stats, dproj, w.e, w.g, w.de, w.dg, w.counter, st)) return e;
    if (bias_gate_h16_ok(D)) tri_post_bias_gate_h16<T><<<dim3(D.N, 2, D.B), 256, 0, st>>>(D, w.de, w.dg, (T *)dproj);
    else tri_post_bias_gate<T><<<dim3(D.N, 2, D.B), 256, psm, st>>>(D, w.de, w.dg, (T *)dproj);
    return check_launch("tri_post_bias_gate");
  }
  if (use_tma()) {
    if (int e = triplet_attn_bwd_tma_launch(D, proj, dva, stats, dproj, w.e, w.g, w.de, w.dg, dbias, st)) return e;
    if (bias_gate_h16_ok(D)) tri_post_bias_gate_h16<T><<<dim3(D.N, 2, D.B), 256, 0, st>>>(D, w.de, w.dg, (T *)dproj);
    else tri_post_bias_gate<T><<<dim3(D.N, 2, D.B), 256, psm, st>>>(D, w.de, w.dg, (T *)dproj);
    return check_launch("tri_post_bias_gate");
  }
  return fail("triplet_attn_bwd: cuTensorMapEncodeTiled is not available from the driver (TMA is required)");
}

// fused forward: bias / gate tiles from the [R, D.ld] E|G projection `proj_eg` (columns D.off_e / D.off_g), then the
// projection + attention kernel of triplet_fused.cu on the raw edge rows x
int triplet_attn_fused_fwd(const tgt_triplet_attn_desc &D, int We, const void *x, int64_t ldx, const float *mean,
                           const float *rstd, const void *wf, const float *wcolsum, const float *wbias,
                           const void *proj_eg, const float *mask, void *va, float *stats, void *ws, size_t ws_bytes,
                           cudaStream_t st) {
  if (!triplet_attn_fused_supported(D, We)) return fail("triplet_attn_fused_fwd: unsupported shape / dtype");
  if (!ws || ws_bytes < triplet_attn_mma_workspace(D, 0))
    return fail("triplet_attn_fused_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, triplet_attn_mma_workspace(D, 0));
  if (((uintptr_t)x | (uintptr_t)va | (uintptr_t)wf) & 15) return fail("triplet_attn_fused_fwd: x / wf / va must be 16-byte aligned");
  if (ldx % 8) return fail("triplet_attn_fused_fwd: ldx must be a multiple of 8");
  if (D.B > 65535) return fail("triplet_attn_fused_fwd: B > 65535 unsupported");
  const Ws w = carve_fwd(D, ws);
  const size_t psm = 2 * (size_t)D.H * 65 * sizeof(float);
  if (D.dtype == TGT_BF16)
    tri_prep_bias_gate<__nv_bfloat16><<<dim3(TN, 2, D.B), 256, psm, st>>>(D, (const __nv_bfloat16 *)proj_eg, mask, w.e, w.g);
  else
    tri_prep_bias_gate<__half><<<dim3(TN, 2, D.B), 256, psm, st>>>(D, (const __half *)proj_eg, mask, w.e, w.g);
  if (int e = check_launch("tri_prep_bias_gate")) return e;
  return triplet_attn_fused_launch(D, We, x, ldx, mean, rstd, wf, wcolsum, wbias, va, stats, w.e, w.g, st);
}

int triplet_attn_fwd_mma(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, void *va, float *stats,
                         void *ws, size_t ws_bytes, cudaStream_t st) {
  if (!ws || ws_bytes < triplet_attn_mma_workspace(D, 0))
    return fail("triplet_attn_fwd: workspace too small (%zu < %zu bytes)", ws_bytes, triplet_attn_mma_workspace(D, 0));
  if (((uintptr_t)proj | (uintptr_t)va) & 15) return fail("triplet_attn_fwd: proj / va must be 16-byte aligned");
  if (D.B > 65535) return fail("triplet_attn_fwd: B > 65535 unsupported");
  if (D.dtype == TGT_BF16) return fwd_impl<__nv_bfloat16>(D, proj, mask, va, stats, ws, st);
  return fwd_impl<__half>(D, proj, mask, va, stats, ws, st);
}

int triplet_attn_bwd_mma(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, const void *va,
                         const void *dva, const float *stats, void *dproj, void *ws, size_t ws_bytes,
                         const void *fwd_ws, float *dbias, cudaStream_t st) {
  if (!ws || ws_bytes < triplet_attn_mma_workspace(D, 1))
    return fail("triplet_attn_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, triplet_attn_mma_workspace(D, 1));
  if (((uintptr_t)proj | (uintptr_t)dva | (uintptr_t)dproj) & 15)
    return fail("triplet_attn_bwd: proj / dva / dproj must be 16-byte aligned");
  if (D.B > 65535) return fail("triplet_attn_bwd: B > 65535 unsupported");
  if (D.dtype == TGT_BF16) return bwd_impl<__nv_bfloat16>(D, proj, mask, va, dva, stats, dproj, ws, fwd_ws, dbias, st);
  return bwd_impl<__half>(D, proj, mask, va, dva, stats, dproj, ws, fwd_ws, dbias, st);
}

int triplet_attn_fwd_tc_f32out(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, float *va_f32,
                               float *stats, void *ws, size_t ws_bytes, cudaStream_t st) {
  if (!triplet_attn_mma_supported(D) || !triplet_attn_tc_supported(D))
    return fail("triplet_attn_fwd_f32out: shape / dtype not supported by the tcgen05 kernel");
  if (!ws || ws_bytes < triplet_attn_mma_workspace(D, 0))
    return fail("triplet_attn_fwd_f32out: workspace too small (%zu < %zu bytes)", ws_bytes, triplet_attn_mma_workspace(D, 0));
  if (((uintptr_t)proj | (uintptr_t)va_f32) & 15) return fail("triplet_attn_fwd_f32out: proj / va must be 16-byte aligned");
  const Ws w = carve_fwd(D, ws);
  const size_t psm = 2 * (size_t)D.H * 65 * sizeof(float);
  if (D.dtype == TGT_BF16)
    tri_prep_bias_gate<__nv_bfloat16><<<dim3(TN, 2, D.B), 256, psm, st>>>(D, (const __nv_bfloat16 *)proj, mask, w.e, w.g);
  else
    tri_prep_bias_gate<__half><<<dim3(TN, 2, D.B), 256, psm, st>>>(D, (const __half *)proj, mask, w.e, w.g);
  if (int e = check_launch("tri_prep_bias_gate")) return e;
  return triplet_attn_fwd_tc_launch(D, proj, nullptr, va_f32, stats, w.e, w.g, st);
}

bool triplet_attn_bwd_bias_available(const tgt_triplet_attn_desc &D) {
  return triplet_attn_mma_supported(D) && !use_tc(D, true) && use_tma() && triplet_attn_bwd_tma_has_bias();
}

}  // namespace tgt
