// Row-wise HBM-bound kernels: LayerNorm fwd/bwd over edge rows, GELU+dropout, scaled residual.
// All are one-pass streaming kernels: 16-byte vector loads/stores, one warp per row for LN
// (W = 256 bf16 -> exactly one 16B vector per lane), grid-stride over rows so the grid can be
// sized to the 148 SMs.
#include <mutex>
#include "common.cuh"
#include <type_traits>

namespace tgt {

constexpr int LN_MAX_ITERS = 8;   // supports W up to 8*32*(16/sizeof(T)) channels

template <typename XT, typename YT>
__global__ void __launch_bounds__(256)
ln_fwd_kernel(const XT *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
              YT *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd,
              int64_t rows, int W, float eps, int64_t ldy) {
  constexpr int V = 4;                       // process 4 channels per lane per iteration (16B fp32 / 8B bf16)
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int iters = (W + 32 * V - 1) / (32 * V);
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const XT *xr = x + r * W;
    float v[LN_MAX_ITERS][V];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < LN_MAX_ITERS; ++it) {
      if (it < iters) {
        const int c = (it * 32 + lane) * V;
        if (c < W) {
          if constexpr (sizeof(XT) == 4) {
            float4 t = *reinterpret_cast<const float4 *>(xr + c);
            v[it][0] = t.x; v[it][1] = t.y; v[it][2] = t.z; v[it][3] = t.w;
          } else {
            uint2 raw = *reinterpret_cast<const uint2 *>(xr + c);
            const XT *e = reinterpret_cast<const XT *>(&raw);
#pragma unroll
            for (int q = 0; q < V; ++q) v[it][q] = to_f(e[q]);
          }
#pragma unroll
          for (int q = 0; q < V; ++q) s += v[it][q];
        } else {
#pragma unroll
          for (int q = 0; q < V; ++q) v[it][q] = 0.f;
        }
      }
    }
    const float mu = warp_sum(s) / (float)W;
    float ss = 0.f;
#pragma unroll
    for (int it = 0; it < LN_MAX_ITERS; ++it) {
      if (it < iters) {
        const int c = (it * 32 + lane) * V;
        if (c < W) {
#pragma unroll
          for (int q = 0; q < V; ++q) { const float dlt = v[it][q] - mu; ss += dlt * dlt; }
        }
      }
    }
    const float rs = rsqrtf(warp_sum(ss) / (float)W + eps);
    if (lane == 0) { mean[r] = mu; rstd[r] = rs; }
    YT *yr = y + r * ldy;
    if (lane < ldy - W) yr[W + lane] = from_f<YT>(lane == 0 ? 1.f : 0.f);      // augmentation columns [1, 0, ...]
#pragma unroll
    for (int it = 0; it < LN_MAX_ITERS; ++it) {
      if (it < iters) {
        const int c = (it * 32 + lane) * V;
        if (c < W) {
          const float4 g = *reinterpret_cast<const float4 *>(gamma + c);
          const float4 b = *reinterpret_cast<const float4 *>(beta + c);
          float o[V];
          o[0] = (v[it][0] - mu) * rs * g.x + b.x;
          o[1] = (v[it][1] - mu) * rs * g.y + b.y;
          o[2] = (v[it][2] - mu) * rs * g.z + b.z;
          o[3] = (v[it][3] - mu) * rs * g.w + b.w;
          if constexpr (sizeof(YT) == 4) {
            *reinterpret_cast<float4 *>(yr + c) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
            uint2 raw;
            YT *e = reinterpret_cast<YT *>(&raw);
#pragma unroll
            for (int q = 0; q < V; ++q) e[q] = from_f<YT>(o[q]);
            *reinterpret_cast<uint2 *>(yr + c) = raw;
          }
        }
      }
    }
  }
}

// dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)) (+ dres);  dgamma += dy*xhat; dbeta += dy
template <typename XT, typename YT>
__global__ void __launch_bounds__(256)
ln_bwd_kernel(const YT *__restrict__ dy, const XT *__restrict__ x, const float *__restrict__ gamma,
              const float *__restrict__ mean, const float *__restrict__ rstd,
              const XT *__restrict__ dres, XT *__restrict__ dx,
              float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t rows, int W) {
  constexpr int V = 4;
  extern __shared__ float sm[];              // [2][W] block partials
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const int iters = (W + 32 * V - 1) / (32 * V);
  for (int c = threadIdx.x; c < 2 * W; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  float ag[LN_MAX_ITERS][V], ab[LN_MAX_ITERS][V];
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it)
#pragma unroll
    for (int q = 0; q < V; ++q) { ag[it][q] = 0.f; ab[it][q] = 0.f; }

  for (int64_t r = warp0; r < rows; r += nwarps) {
    const XT *xr = x + r * W;
    const YT *dyr = dy + r * W;
    const float mu = mean[r], rs = rstd[r];
    float xh[LN_MAX_ITERS][V], gd[LN_MAX_ITERS][V];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int it = 0; it < LN_MAX_ITERS; ++it) {
      if (it < iters) {
        const int c = (it * 32 + lane) * V;
        if (c < W) {
          float xv[V], dv[V];
          if constexpr (sizeof(XT) == 4) {
            float4 t = *reinterpret_cast<const float4 *>(xr + c);
            xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
          } else {
            uint2 raw = *reinterpret_cast<const uint2 *>(xr + c);
            const XT *e = reinterpret_cast<const XT *>(&raw);
#pragma unroll
            for (int q = 0; q < V; ++q) xv[q] = to_f(e[q]);
          }
          if constexpr (sizeof(YT) == 4) {
            float4 t = *reinterpret_cast<const float4 *>(dyr + c);
            dv[0] = t.x; dv[1] = t.y; dv[2] = t.z; dv[3] = t.w;
          } else {
            uint2 raw = *reinterpret_cast<const uint2 *>(dyr + c);
            const YT *e = reinterpret_cast<const YT *>(&raw);
#pragma unroll
            for (int q = 0; q < V; ++q) dv[q] = to_f(e[q]);
          }
          const float4 g = *reinterpret_cast<const float4 *>(gamma + c);
          const float gg[V] = {g.x, g.y, g.z, g.w};
#pragma unroll
          for (int q = 0; q < V; ++q) {
            xh[it][q] = (xv[q] - mu) * rs;
            gd[it][q] = gg[q] * dv[q];
            s1 += gd[it][q];
            s2 += gd[it][q] * xh[it][q];
            ag[it][q] += dv[q] * xh[it][q];
            ab[it][q] += dv[q];
          }
        }
      }
    }
    s1 = warp_sum(s1) / (float)W;
    s2 = warp_sum(s2) / (float)W;
    XT *dxr = dx + r * W;
#pragma unroll
    for (int it = 0; it < LN_MAX_ITERS; ++it) {
      if (it < iters) {
        const int c = (it * 32 + lane) * V;
        if (c < W) {
          float o[V];
#pragma unroll
          for (int q = 0; q < V; ++q) o[q] = rs * (gd[it][q] - s1 - xh[it][q] * s2);
          if (dres != nullptr) {
            if constexpr (sizeof(XT) == 4) {
              float4 t = *reinterpret_cast<const float4 *>(dres + r * W + c);
              o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
            } else {
              uint2 raw = *reinterpret_cast<const uint2 *>(dres + r * W + c);
              const XT *e = reinterpret_cast<const XT *>(&raw);
#pragma unroll
              for (int q = 0; q < V; ++q) o[q] += to_f(e[q]);
            }
          }
          if constexpr (sizeof(XT) == 4) {
            *reinterpret_cast<float4 *>(dxr + c) = make_float4(o[0], o[1], o[2], o[3]);
          } else {
            uint2 raw;
            XT *e = reinterpret_cast<XT *>(&raw);
#pragma unroll
            for (int q = 0; q < V; ++q) e[q] = from_f<XT>(o[q]);
            *reinterpret_cast<uint2 *>(dxr + c) = raw;
          }
        }
      }
    }
  }
  // block reduce the per-lane channel partials through shared memory, then one atomic per channel
#pragma unroll
  for (int it = 0; it < LN_MAX_ITERS; ++it) {
    if (it < iters) {
      const int c = (it * 32 + lane) * V;
      if (c < W) {
#pragma unroll
        for (int q = 0; q < V; ++q) {
          atomicAdd(&sm[c + q], ag[it][q]);
          atomicAdd(&sm[W + c + q], ab[it][q]);
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    atomicAdd(&dgamma[c], sm[c]);
    atomicAdd(&dbeta[c], sm[W + c]);
  }
}


// ------------------------------------------------------------------ W = 256, 16-bit in/out fast path
// One warp per row, one 16-byte vector per lane, RU rows in flight per warp so that enough bytes are
// outstanding per SM to cover the HBM latency (Little: ~45 KB/SM at 6.5 TB/s).
template <typename T> struct V8 {
  typedef uint4 Raw;
  static __device__ __forceinline__ Raw load(const T *p) { return *reinterpret_cast<const uint4 *>(p); }
  static __device__ __forceinline__ void store(T *p, const float (&o)[8]) { *reinterpret_cast<uint4 *>(p) = pack(o); }
  static __device__ __forceinline__ void unpack(const uint4 &raw, float (&o)[8]) {
    const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
    for (int q = 0; q < 8; ++q) o[q] = to_f(e[q]);
  }
  static __device__ __forceinline__ uint4 pack(const float (&o)[8]) {
    uint4 raw;
    T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
    for (int q = 0; q < 8; ++q) e[q] = from_f<T>(o[q]);
    return raw;
  }
};

// fp32 rows (the embedding output that enters the first layer): 8 elements = two 16-byte vectors
template <> struct V8<float> {
  struct Raw { float4 a, b; };
  static __device__ __forceinline__ Raw load(const float *p) {
    Raw r;
    r.a = *reinterpret_cast<const float4 *>(p);
    r.b = *reinterpret_cast<const float4 *>(p + 4);
    return r;
  }
  static __device__ __forceinline__ void unpack(const Raw &r, float (&o)[8]) {
    o[0] = r.a.x; o[1] = r.a.y; o[2] = r.a.z; o[3] = r.a.w; o[4] = r.b.x; o[5] = r.b.y; o[6] = r.b.z; o[7] = r.b.w;
  }
  static __device__ __forceinline__ void store(float *p, const float (&o)[8]) {
    *reinterpret_cast<float4 *>(p) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4 *>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
};

template <typename XT, typename T, int RU>
__global__ void __launch_bounds__(256)
ln_fwd_w256(const XT *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ beta,
            T *__restrict__ y, float *__restrict__ mean, float *__restrict__ rstd, int64_t rows, float eps, int64_t ldy) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float g[8], bt[8];
  {
    const float4 g0 = *reinterpret_cast<const float4 *>(gamma + lane * 8), g1 = *reinterpret_cast<const float4 *>(gamma + lane * 8 + 4);
    const float4 b0 = *reinterpret_cast<const float4 *>(beta + lane * 8), b1 = *reinterpret_cast<const float4 *>(beta + lane * 8 + 4);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
    bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
  }
  for (int64_t r0 = warp0 * RU; r0 < rows; r0 += nwarps * RU) {
    typename V8<XT>::Raw raw[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u)
      if (r0 + u < rows) raw[u] = V8<XT>::load(x + (r0 + u) * 256 + lane * 8);
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      if (r0 + u >= rows) continue;
      float v[8];
      V8<XT>::unpack(raw[u], v);
      float s = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) s += v[q];
      const float mu = warp_sum(s) * (1.f / 256.f);
      float ss = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) { v[q] -= mu; ss += v[q] * v[q]; }
      const float rs = rsqrtf(warp_sum(ss) * (1.f / 256.f) + eps);
      if (lane == 0) { mean[r0 + u] = mu; rstd[r0 + u] = rs; }
#pragma unroll
      for (int q = 0; q < 8; ++q) v[q] = v[q] * rs * g[q] + bt[q];
      *reinterpret_cast<uint4 *>(y + (r0 + u) * ldy + lane * 8) = V8<T>::pack(v);
      if (ldy > 256 && lane == 0) {                  // augmentation columns [1, 0, 0, 0, 0, 0, 0, 0]
        const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        *reinterpret_cast<uint4 *>(y + (r0 + u) * ldy + 256) = V8<T>::pack(one);
      }
    }
  }
}

// EMIT_Y: also write y = LN(x) (the normalised rows are in registers anyway) as [rows, ldy] with the augmentation columns
// -- the weight-gradient GEMM of the Linear behind the LayerNorm then needs no separate LayerNorm recompute pass.
// COLSUM: also accumulate  dres_colsum[c] += sum_r w_r * dres[r, c],  w_r = res_row_scale[r / rows_per_scale] (1 if null):
// the residual gradient dres IS the upstream gradient of the module whose LayerNorm this is, so this is the bias gradient
// of the module's output projection (DropPath-weighted) -- it costs no pass of its own.
//
// The kernel moves up to five edge-sized streams (dy, x, dres in; dx, y out) and nothing else, so what matters is bytes in
// flight: each warp owns a private shared-memory ring of LNB_STAGES stages x RU rows x (x | dy | dres) filled by cp.async
// (every lane copies and later reads back its own 16-byte pieces, so no barrier is needed), i.e. ~9 KB per warp and
// ~144 KB per SM requested ahead of the arithmetic instead of the 48 KB that register staging allowed.
constexpr int LNB_STAGES = 3;
template <typename XT> __host__ __device__ constexpr int lnb_stage_bytes(int ru) { return ru * (2 * 256 * (int)sizeof(XT) + 512); }

__device__ __forceinline__ void lnb_cp16(uint32_t dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src));
}

template <typename XT, typename T, int RU, bool EMIT_Y = false, bool COLSUM = false>
__global__ void __launch_bounds__(256)
ln_bwd_w256(const T *__restrict__ dy, const XT *__restrict__ x, const float *__restrict__ gamma,
            const float *__restrict__ mean, const float *__restrict__ rstd, const XT *__restrict__ dres,
            XT *__restrict__ dx, float *__restrict__ dgamma, float *__restrict__ dbeta, int64_t rows,
            const float *__restrict__ beta = nullptr, T *__restrict__ y = nullptr, int64_t ldy = 0,
            const float *__restrict__ res_row_scale = nullptr, int64_t rows_per_scale = 1,
            float *__restrict__ dres_colsum = nullptr) {
  __shared__ float sm[(COLSUM ? 3 : 2) * 256];
  extern __shared__ __align__(16) unsigned char ring_raw[];
  constexpr int XB = 256 * (int)sizeof(XT);                   // bytes of an x / dres row
  constexpr int XV = XB / 512;                                // 16-byte pieces per lane of such a row (1 or 2)
  constexpr int ROWB = 2 * XB + 512;                          // x | dres | dy of one row
  constexpr int STAGE = RU * ROWB;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + wib;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(ring_raw) + (uint32_t)wib * (LNB_STAGES * STAGE);
  unsigned char *ring_gen = ring_raw + wib * (LNB_STAGES * STAGE);
  for (int c = threadIdx.x; c < (COLSUM ? 768 : 512); c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  float g[8], ag[8], ab[8], ar[COLSUM ? 8 : 1];
  if constexpr (COLSUM) {
#pragma unroll
    for (int q = 0; q < 8; ++q) ar[q] = 0.f;
  }
  {
    const float4 g0 = *reinterpret_cast<const float4 *>(gamma + lane * 8), g1 = *reinterpret_cast<const float4 *>(gamma + lane * 8 + 4);
    g[0] = g0.x; g[1] = g0.y; g[2] = g0.z; g[3] = g0.w; g[4] = g1.x; g[5] = g1.y; g[6] = g1.z; g[7] = g1.w;
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) ag[q] = ab[q] = 0.f;
  float bt[8];
  if constexpr (EMIT_Y) {
    const float4 b0 = *reinterpret_cast<const float4 *>(beta + lane * 8), b1 = *reinterpret_cast<const float4 *>(beta + lane * 8 + 4);
    bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
  }
  // stage `slot` <- rows r0 .. r0+RU-1: this lane's 8 channels of x, dres and dy (lane * 8 elements into each row)
  auto issue = [&](int64_t r0, int slot) {
    if (r0 < rows) {
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        if (r0 + u < rows) {
          const uint32_t d = ring + slot * STAGE + u * ROWB;
          const unsigned char *xs = reinterpret_cast<const unsigned char *>(x + (r0 + u) * 256 + lane * 8);
#pragma unroll
          for (int v = 0; v < XV; ++v) lnb_cp16(d + lane * (16 * XV) + v * 16, xs + v * 16);
          if (dres) {
            const unsigned char *rs_ = reinterpret_cast<const unsigned char *>(dres + (r0 + u) * 256 + lane * 8);
#pragma unroll
            for (int v = 0; v < XV; ++v) lnb_cp16(d + XB + lane * (16 * XV) + v * 16, rs_ + v * 16);
          }
          lnb_cp16(d + 2 * XB + lane * 16, dy + (r0 + u) * 256 + lane * 8);
        }
      }
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  const int64_t step = nwarps * RU;
#pragma unroll
  for (int s = 0; s < LNB_STAGES - 1; ++s) issue(warp0 * RU + s * step, s);
  int slot = 0;
  for (int64_t r0 = warp0 * RU; r0 < rows; r0 += step) {
    int nslot = slot + LNB_STAGES - 1;
    if (nslot >= LNB_STAGES) nslot -= LNB_STAGES;
    issue(r0 + (LNB_STAGES - 1) * step, nslot);               // (the slot processed in the previous iteration)
    asm volatile("cp.async.wait_group %0;\n" ::"n"(LNB_STAGES - 1));
    float mu[RU], rs[RU];
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      if (r0 + u < rows) {
        mu[u] = mean[r0 + u];
        rs[u] = rstd[r0 + u];
      }
    }
#pragma unroll
    for (int u = 0; u < RU; ++u) {
      if (r0 + u >= rows) continue;
      const unsigned char *st = ring_gen + slot * STAGE + u * ROWB;
      float xv[8], dv[8];
      V8<XT>::unpack(V8<XT>::load(reinterpret_cast<const XT *>(st + lane * (16 * XV))), xv);
      V8<T>::unpack(*reinterpret_cast<const uint4 *>(st + 2 * XB + lane * 16), dv);
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        xv[q] = (xv[q] - mu[u]) * rs[u];          // xhat
        ag[q] += dv[q] * xv[q];
        ab[q] += dv[q];
        dv[q] *= g[q];                            // g * dy
        s1 += dv[q];
        s2 += dv[q] * xv[q];
      }
      if constexpr (EMIT_Y) {                     // same expression as ln_fwd_w256: bit-identical to the recompute
        float yv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) yv[q] = xv[q] * g[q] + bt[q];
        *reinterpret_cast<uint4 *>(y + (r0 + u) * ldy + lane * 8) = V8<T>::pack(yv);
        if (ldy > 256 && lane == 0) {
          const float one[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          *reinterpret_cast<uint4 *>(y + (r0 + u) * ldy + 256) = V8<T>::pack(one);
        }
      }
      s1 = warp_sum(s1) * (1.f / 256.f);
      s2 = warp_sum(s2) * (1.f / 256.f);
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = rs[u] * (dv[q] - s1 - xv[q] * s2);
      if (dres) {
        float rv[8];
        V8<XT>::unpack(V8<XT>::load(reinterpret_cast<const XT *>(st + XB + lane * (16 * XV))), rv);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] += rv[q];
        if constexpr (COLSUM) {
          const float w = res_row_scale ? res_row_scale[(r0 + u) / rows_per_scale] : 1.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) ar[q] = fmaf(w, rv[q], ar[q]);
        }
      }
      V8<XT>::store(dx + (r0 + u) * 256 + lane * 8, o);
    }
    slot = slot + 1 == LNB_STAGES ? 0 : slot + 1;
  }
  asm volatile("cp.async.wait_group 0;\n" ::);
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    atomicAdd(&sm[lane * 8 + q], ag[q]);
    atomicAdd(&sm[256 + lane * 8 + q], ab[q]);
    if constexpr (COLSUM) atomicAdd(&sm[512 + lane * 8 + q], ar[q]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 256; c += blockDim.x) {
    atomicAdd(&dgamma[c], sm[c]);
    atomicAdd(&dbeta[c], sm[256 + c]);
    if constexpr (COLSUM) atomicAdd(&dres_colsum[c], sm[512 + c]);
  }
}

// ------------------------------------------------------------------ W = 256*NV (node rows, W = 768), 16-bit in/out
// The node-side tensors have only B*N rows (16 K at config 3), far too few to hide the memory latency by occupancy: one
// warp per row, NV 16-byte vectors of dy / x / dres per lane, and the NEXT row's vectors are requested before the current
// row is reduced.  One persistent CTA per SM, so the per-CTA reduction of dgamma / dbeta (shared-memory atomics, then
// 2*W global atomics) is paid 148 times, not once per 32 rows.
template <typename T, int NV>
__global__ void __launch_bounds__(256)
ln_bwd_wide(const T *__restrict__ dy, const T *__restrict__ x, const float *__restrict__ gamma, const float *__restrict__ mean,
            const float *__restrict__ rstd, const T *__restrict__ dres, T *__restrict__ dx, float *__restrict__ dgamma,
            float *__restrict__ dbeta, int64_t rows) {
  constexpr int W = 256 * NV;
  __shared__ float sm[2 * W];
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int c = threadIdx.x; c < 2 * W; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  float g[NV][8], ag[NV][8], ab[NV][8];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const float *gp = gamma + (v * 32 + lane) * 8;
    const float4 g0 = *reinterpret_cast<const float4 *>(gp), g1 = *reinterpret_cast<const float4 *>(gp + 4);
    g[v][0] = g0.x; g[v][1] = g0.y; g[v][2] = g0.z; g[v][3] = g0.w; g[v][4] = g1.x; g[v][5] = g1.y; g[v][6] = g1.z; g[v][7] = g1.w;
#pragma unroll
    for (int q = 0; q < 8; ++q) ag[v][q] = ab[v][q] = 0.f;
  }
  uint4 nx[NV], nd[NV], nr[NV];
  float nmu = 0.f, nrs = 0.f;
  auto fetch = [&](int64_t r) {
    if (r < rows) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int64_t o = r * W + (v * 32 + lane) * 8;
        nx[v] = *reinterpret_cast<const uint4 *>(x + o);
        nd[v] = *reinterpret_cast<const uint4 *>(dy + o);
        if (dres) nr[v] = *reinterpret_cast<const uint4 *>(dres + o);
      }
      nmu = mean[r];
      nrs = rstd[r];
    }
  };
  fetch(warp0);
  for (int64_t r = warp0; r < rows; r += nwarps) {
    uint4 cx[NV], cd[NV], cr[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) { cx[v] = nx[v]; cd[v] = nd[v]; cr[v] = nr[v]; }
    const float mu = nmu, rs = nrs;
    fetch(r + nwarps);
    float xh[NV][8], gd[NV][8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float dv[8];
      V8<T>::unpack(cx[v], xh[v]);
      V8<T>::unpack(cd[v], dv);
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        xh[v][q] = (xh[v][q] - mu) * rs;
        ag[v][q] += dv[q] * xh[v][q];
        ab[v][q] += dv[q];
        gd[v][q] = dv[q] * g[v][q];
        s1 += gd[v][q];
        s2 += gd[v][q] * xh[v][q];
      }
    }
    s1 = warp_sum(s1) * (1.f / W);
    s2 = warp_sum(s2) * (1.f / W);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float o[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) o[q] = rs * (gd[v][q] - s1 - xh[v][q] * s2);
      if (dres) {
        float rv[8];
        V8<T>::unpack(cr[v], rv);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] += rv[q];
      }
      V8<T>::store(dx + r * W + (v * 32 + lane) * 8, o);
    }
  }
#pragma unroll
  for (int v = 0; v < NV; ++v)
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      atomicAdd(&sm[(v * 32 + lane) * 8 + q], ag[v][q]);
      atomicAdd(&sm[W + (v * 32 + lane) * 8 + q], ab[v][q]);
    }
  __syncthreads();
  for (int c = threadIdx.x; c < W; c += blockDim.x) {
    atomicAdd(&dgamma[c], sm[c]);
    atomicAdd(&dbeta[c], sm[W + c]);
  }
}

// ------------------------------------------------------------------ cross-entropy over distance bins, one warp per row
// F.cross_entropy(logits, target, reduction='none') of DiscreteDistLoss (training_schemes/pcqm/commons.py:26-48) on the
// [R, nbins] logits of the distance head, R = B*N*N.  The torch path converts the 16-bit logits to fp32 (2 passes), keeps
// fp32 log-probabilities for backward and runs log-softmax backward over them: 8 GB of traffic for 1 GB of logits.  Here
// the forward reads the logits once (row max, log-sum-exp in fp32; xent and lse are [R] vectors) and the backward writes
// d logits = (softmax - onehot) * g_r straight from the logits; rows whose upstream gradient is zero (masked pairs) are
// neither read nor exponentiated.  NCH = nbins / 256 chunks of 8 columns per lane.
template <typename T, int NCH>
__global__ void __launch_bounds__(256)
xent_rows_fwd(const T *__restrict__ logits, int64_t ld, const int64_t *__restrict__ target, float *__restrict__ xent,
              float *__restrict__ lse, int64_t rows, int nbins) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp0; r < rows; r += nwarps) {
    typename V8<T>::Raw raw[NCH];
    const T *row = logits + r * ld;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if ((c * 32 + lane) * 8 < nbins) raw[c] = V8<T>::load(row + (c * 32 + lane) * 8);
    const int64_t t = target[r];
    float v[NCH][8], mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if ((c * 32 + lane) * 8 < nbins) {
        V8<T>::unpack(raw[c], v[c]);
#pragma unroll
        for (int q = 0; q < 8; ++q) mx = fmaxf(mx, v[c][q]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f, at = 0.f;
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if ((c * 32 + lane) * 8 < nbins) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          sum += __expf(v[c][q] - mx);
          if ((c * 32 + lane) * 8 + q == t) at = v[c][q];
        }
      }
    }
    sum = warp_sum(sum);
    at = warp_sum(at);                         // exactly one lane holds the target column
    if (lane == 0) {
      const float l = mx + __logf(sum);
      lse[r] = l;
      xent[r] = (t >= 0 && t < nbins) ? l - at : 0.f;
    }
  }
}

template <typename T, int NCH>
__global__ void __launch_bounds__(256)
xent_rows_bwd(const T *__restrict__ logits, int64_t ld, const int64_t *__restrict__ target, const float *__restrict__ lse,
              const float *__restrict__ gx, T *__restrict__ dlogits, int64_t rows, int nbins) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float g = gx[r];
    T *out = dlogits + r * (int64_t)nbins;
    if (g == 0.f) {                            // masked pair: no read, no exp
      const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < NCH; ++c)
        if ((c * 32 + lane) * 8 < nbins) V8<T>::store(out + (c * 32 + lane) * 8, z);
      continue;
    }
    typename V8<T>::Raw raw[NCH];
    const T *row = logits + r * ld;
#pragma unroll
    for (int c = 0; c < NCH; ++c)
      if ((c * 32 + lane) * 8 < nbins) raw[c] = V8<T>::load(row + (c * 32 + lane) * 8);
    const int64_t t = target[r];
    const float l = lse[r];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
      if ((c * 32 + lane) * 8 < nbins) {
        float v[8], o[8];
        V8<T>::unpack(raw[c], v);
#pragma unroll
        for (int q = 0; q < 8; ++q) o[q] = (__expf(v[q] - l) - ((c * 32 + lane) * 8 + q == t ? 1.f : 0.f)) * g;
        V8<T>::store(out + (c * 32 + lane) * 8, o);
      }
    }
  }
}

// ------------------------------------------------------------------ gelu + dropout
__device__ __forceinline__ float gelu_f(float u) { return 0.5f * u * (1.f + erff(u * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad(float u) {
  const float cdf = 0.5f * (1.f + erff(u * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(256)
gelu_dropout_kernel(const T *__restrict__ u, const T *__restrict__ dy, T *__restrict__ out, int64_t n,
                    float p_drop, uint64_t seed, const uint64_t *__restrict__ seed_ptr) {
  constexpr int NV = 16 / sizeof(T);
  const DropCfg dc = make_drop_cfg(p_drop, seed_ptr ? *seed_ptr : seed);
  const int64_t nvec = n / NV;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
    float a[NV], o[NV];
    load_vec<T, NV>(u + i * NV, a);
    float g[NV];
    if constexpr (BWD) load_vec<T, NV>(dy + i * NV, g);
#pragma unroll
    for (int q = 0; q < NV; q += 2) {
      float m0 = 1.f, m1 = 1.f;
      if (p_drop > 0.f) drop_mask_pair(dc, (uint64_t)(i * NV + q) >> 1, m0, m1);
      if constexpr (BWD) {
        o[q] = g[q] * m0 * gelu_grad(a[q]);
        o[q + 1] = g[q + 1] * m1 * gelu_grad(a[q + 1]);
      } else {
        o[q] = gelu_f(a[q]) * m0;
        o[q + 1] = gelu_f(a[q + 1]) * m1;
      }
    }
    store_vec<T, NV>(out + i * NV, o);
  }
  // tail
  if (blockIdx.x == 0) {
    for (int64_t e = nvec * NV + threadIdx.x; e < n; e += blockDim.x) {
      const float m = p_drop > 0.f ? drop_mask_one(dc, (uint64_t)e) : 1.f;
      const float a = to_f(u[e]);
      out[e] = from_f<T>(BWD ? to_f(dy[e]) * m * gelu_grad(a) : gelu_f(a) * m);
    }
  }
}

// ------------------------------------------------------------------ out = res + scale[b]*x
template <typename T, typename RT>
__global__ void __launch_bounds__(256)
scaled_residual_kernel(const T *__restrict__ x, const RT *__restrict__ res, const float *__restrict__ scale,
                       T *__restrict__ out, int64_t B, int64_t inner) {
  constexpr int NV = 4;
  const int64_t per = inner / NV;
  const int64_t total = B * per;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t b = i / per;
    const float s = scale ? scale[b] : 1.f;
    const int64_t off = i * NV;
    float xv[NV], rv[NV];
    if constexpr (sizeof(T) == 4) {
      float4 t = *reinterpret_cast<const float4 *>(x + off);
      xv[0] = t.x; xv[1] = t.y; xv[2] = t.z; xv[3] = t.w;
    } else {
      uint2 raw = *reinterpret_cast<const uint2 *>(x + off);
      const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
      for (int q = 0; q < NV; ++q) xv[q] = to_f(e[q]);
    }
    if (res == nullptr) {
      rv[0] = rv[1] = rv[2] = rv[3] = 0.f;
    } else if constexpr (sizeof(RT) == 4) {
      float4 t = *reinterpret_cast<const float4 *>(res + off);
      rv[0] = t.x; rv[1] = t.y; rv[2] = t.z; rv[3] = t.w;
    } else {
      uint2 raw = *reinterpret_cast<const uint2 *>(res + off);
      const RT *e = reinterpret_cast<const RT *>(&raw);
#pragma unroll
      for (int q = 0; q < NV; ++q) rv[q] = to_f(e[q]);
    }
    float o[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) o[q] = rv[q] + s * xv[q];
    if constexpr (sizeof(T) == 4) {
      *reinterpret_cast<float4 *>(out + off) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      uint2 raw;
      T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
      for (int q = 0; q < NV; ++q) e[q] = from_f<T>(o[q]);
      *reinterpret_cast<uint2 *>(out + off) = raw;
    }
  }
}

// ------------------------------------------------------------------ Gaussian basis of the 3-D distance embedding
// out[r,k] = exp(-0.5 ((x_r - mu_k) / sd_k)^2) / (sqrt(2*3.14159) sd_k)      (reference models/pcqm/layers.py:25-48)
// One warp per row, K <= 128 kernels (4 per lane).  Backward recomputes the basis from x (nothing O(R*K) is saved) and
// reduces d mu / d sd over the rows in registers -> shared memory -> one atomic per kernel and CTA.
constexpr float GB_NORM = 2.5066272622f;      // (2 * 3.14159) ** 0.5, the reference's constant

template <typename T>
__global__ void __launch_bounds__(256)
gaussian_basis_fwd_kernel(const float *__restrict__ x, const float *__restrict__ mu, const float *__restrict__ sd,
                          T *__restrict__ out, int64_t rows, int K) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  float m[4], is[4], nrm[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int k = lane * 4 + q;
    m[q] = k < K ? mu[k] : 0.f;
    const float s_ = k < K ? sd[k] : 1.f;
    is[q] = 1.f / s_;
    nrm[q] = 1.f / (GB_NORM * s_);
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float xr = x[r];
    float o[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float z = (xr - m[q]) * is[q];
      o[q] = __expf(-0.5f * z * z) * nrm[q];
    }
    if (lane * 4 + 3 < K) {
      if constexpr (sizeof(T) == 4) {
        *reinterpret_cast<float4 *>(out + r * K + lane * 4) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        T v[4] = {from_f<T>(o[0]), from_f<T>(o[1]), from_f<T>(o[2]), from_f<T>(o[3])};
        *reinterpret_cast<uint2 *>(out + r * K + lane * 4) = *reinterpret_cast<uint2 *>(v);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (lane * 4 + q < K) out[r * K + lane * 4 + q] = from_f<T>(o[q]);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
gaussian_basis_bwd_kernel(const float *__restrict__ x, const float *__restrict__ mu, const float *__restrict__ sd,
                          const T *__restrict__ dout, float *__restrict__ dx, float *__restrict__ dmu,
                          float *__restrict__ dsd, int64_t rows, int K) {
  __shared__ float sm[2 * 128];
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int c = threadIdx.x; c < 256; c += blockDim.x) sm[c] = 0.f;
  __syncthreads();
  float m[4], is[4], nrm[4], am[4], as[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int k = lane * 4 + q;
    m[q] = k < K ? mu[k] : 0.f;
    const float s_ = k < K ? sd[k] : 1.f;
    is[q] = 1.f / s_;
    nrm[q] = k < K ? 1.f / (GB_NORM * s_) : 0.f;
    am[q] = as[q] = 0.f;
  }
  for (int64_t r = warp0; r < rows; r += nwarps) {
    const float xr = x[r];
    float g[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane * 4 + 3 < K) {
      if constexpr (sizeof(T) == 4) {
        const float4 v = *reinterpret_cast<const float4 *>(dout + r * K + lane * 4);
        g[0] = v.x; g[1] = v.y; g[2] = v.z; g[3] = v.w;
      } else {
        const uint2 raw = *reinterpret_cast<const uint2 *>(dout + r * K + lane * 4);
        const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
        for (int q = 0; q < 4; ++q) g[q] = to_f(e[q]);
      }
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (lane * 4 + q < K) g[q] = to_f(dout[r * K + lane * 4 + q]);
    }
    float dxr = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float z = (xr - m[q]) * is[q];
      const float go = g[q] * __expf(-0.5f * z * z) * nrm[q];      // dout * out
      const float t = go * z * is[q];                               // dout * out * z / sd
      dxr -= t;
      am[q] += t;
      as[q] += go * (z * z - 1.f) * is[q];
    }
    dxr = warp_sum(dxr);
    if (lane == 0) dx[r] = dxr;
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    atomicAdd(&sm[lane * 4 + q], am[q]);
    atomicAdd(&sm[128 + lane * 4 + q], as[q]);
  }
  __syncthreads();
  for (int c = threadIdx.x; c < K; c += blockDim.x) {
    atomicAdd(&dmu[c], sm[c]);
    atomicAdd(&dsd[c], sm[128 + c]);
  }
}

static int grid_for(int64_t work_items, int per_block) {
  int64_t g = (work_items + per_block - 1) / per_block;
  const int64_t cap = 148 * 16;       // a few resident CTAs per SM, multiple of the SM count
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <typename XT, typename YT>
static int ln_fwd_launch(const void *x, const float *gamma, const float *beta, void *y, float *mean,
                         float *rstd, int64_t rows, int W, float eps, int64_t ldy, cudaStream_t st) {
  if constexpr (sizeof(YT) == 2 && (std::is_same<XT, YT>::value || std::is_same<XT, float>::value)) {
    if (W == 256 && (ldy == 256 || ldy == 264) && (((uintptr_t)x | (uintptr_t)y) & 15) == 0) {
      ln_fwd_w256<XT, YT, 4><<<grid_for(rows, 8 * 4), 256, 0, st>>>((const XT *)x, gamma, beta, (YT *)y, mean, rstd, rows, eps,
                                                                 ldy);
      return check_launch("ln_fwd_w256");
    }
  }
  ln_fwd_kernel<XT, YT><<<grid_for(rows, 8), 256, 0, st>>>((const XT *)x, gamma, beta, (YT *)y, mean, rstd,
                                                            rows, W, eps, ldy);
  return check_launch("ln_fwd_kernel");
}
template <typename XT, typename YT>
static int ln_bwd_launch(const void *dy, const void *x, const float *gamma, const float *mean,
                         const float *rstd, const void *dres, void *dx, float *dgamma, float *dbeta,
                         int64_t rows, int W, cudaStream_t st, const float *beta = nullptr, void *y = nullptr,
                         int64_t ldy = 0, const float *res_row_scale = nullptr, int64_t rows_per_scale = 1,
                         float *dres_colsum = nullptr) {
  if constexpr (sizeof(YT) == 2 && (std::is_same<XT, YT>::value || std::is_same<XT, float>::value)) {
    if (W == 256 && (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dres | (uintptr_t)y) & 15) == 0) {
      int g = grid_for(rows, 8 * 2 * 8);
      if (g > 148 * 8) g = 148 * 8;
      constexpr int RING = 8 * LNB_STAGES * lnb_stage_bytes<XT>(2);       // 8 warps x 3 stages x 2 rows: 72 KB (16-bit x)
#define LNB_LAUNCH(EY, CS, ...)                                                                                          \
  do {                                                                                                                   \
    static std::once_flag once;                                                                                          \
    std::call_once(once, [] {                                                                                            \
      cudaFuncSetAttribute(ln_bwd_w256<XT, YT, 2, EY, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, RING);              \
    });                                                                                                                  \
    ln_bwd_w256<XT, YT, 2, EY, CS><<<g, 256, RING, st>>>(__VA_ARGS__);                                                    \
  } while (0)
      if (dres_colsum && (!dres || rows_per_scale <= 0)) return fail("layernorm_bwd_y: dres_colsum needs dres and rows_per_scale > 0");
      if (y) {
        if (!beta || (ldy != 256 && ldy != 264)) return fail("layernorm_bwd_y: beta is required and ldy must be 256 or 264");
        if (dres_colsum)
          LNB_LAUNCH(true, true, (const YT *)dy, (const XT *)x, gamma, mean, rstd, (const XT *)dres, (XT *)dx, dgamma, dbeta, rows,
                     beta, (YT *)y, ldy, res_row_scale, rows_per_scale, dres_colsum);
        else
          LNB_LAUNCH(true, false, (const YT *)dy, (const XT *)x, gamma, mean, rstd, (const XT *)dres, (XT *)dx, dgamma, dbeta, rows,
                     beta, (YT *)y, ldy);
        return check_launch("ln_bwd_w256_y");
      }
      if (dres_colsum) {
        LNB_LAUNCH(false, true, (const YT *)dy, (const XT *)x, gamma, mean, rstd, (const XT *)dres, (XT *)dx, dgamma, dbeta, rows,
                   nullptr, nullptr, 0, res_row_scale, rows_per_scale, dres_colsum);
        return check_launch("ln_bwd_w256_cs");
      }
      LNB_LAUNCH(false, false, (const YT *)dy, (const XT *)x, gamma, mean, rstd, (const XT *)dres, (XT *)dx, dgamma, dbeta, rows);
#undef LNB_LAUNCH
      return check_launch("ln_bwd_w256");
    }
  }
  if (y || dres_colsum) return fail("layernorm_bwd_y: only the W = 256, 16-bit gradient fast path can emit y / dres_colsum");
  if constexpr (sizeof(YT) == 2 && std::is_same<XT, YT>::value) {
    if (W == 768 && (((uintptr_t)x | (uintptr_t)dy | (uintptr_t)dx | (uintptr_t)dres) & 15) == 0) {
      const int g = (int)std::min<int64_t>((rows + 7) / 8, 148);
      ln_bwd_wide<YT, 3><<<g, 256, 0, st>>>((const YT *)dy, (const YT *)x, gamma, mean, rstd, (const YT *)dres, (YT *)dx, dgamma,
                                          dbeta, rows);
      return check_launch("ln_bwd_wide");
    }
  }
  int g = grid_for(rows, 8 * 4);            // few rows per warp: node-side tensors have only B*N rows
  if (g > 148 * 4) g = 148 * 4;
  ln_bwd_kernel<XT, YT><<<g, 256, 2 * W * sizeof(float), st>>>((const YT *)dy, (const XT *)x, gamma, mean, rstd,
                                                               (const XT *)dres, (XT *)dx, dgamma, dbeta, rows, W);
  return check_launch("ln_bwd_kernel");
}

}  // namespace tgt

using namespace tgt;

#define LN_COMBOS(X)                                    \
  X(TGT_F32, TGT_F32, float, float)                     \
  X(TGT_F32, TGT_BF16, float, __nv_bfloat16)            \
  X(TGT_BF16, TGT_BF16, __nv_bfloat16, __nv_bfloat16)   \
  X(TGT_F32, TGT_F16, float, __half)                    \
  X(TGT_F16, TGT_F16, __half, __half)

extern "C" int tgt_layernorm_fwd(const void *x, const float *gamma, const float *beta, void *y, float *mean,
                                 float *rstd, int64_t rows, int W, int64_t ldy, float eps, int x_dtype, int y_dtype,
                                 void *stream) {
  if (W % 4 != 0 || W > LN_MAX_ITERS * 128) return fail("layernorm: W=%d unsupported (need W%%4==0, W<=%d)", W, LN_MAX_ITERS * 128);
  if (ldy < W || ldy > W + 32 || ldy % 4) return fail("layernorm: ldy=%lld must be in [W, W+32] and a multiple of 4", (long long)ldy);
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define X(xc, yc, XT, YT) \
  if (x_dtype == xc && y_dtype == yc) return ln_fwd_launch<XT, YT>(x, gamma, beta, y, mean, rstd, rows, W, eps, ldy, st);
  LN_COMBOS(X)
#undef X
  return fail("layernorm_fwd: unsupported dtype combination x=%d y=%d", x_dtype, y_dtype);
}

extern "C" int tgt_layernorm_bwd(const void *dy, const void *x, const float *gamma, const float *mean,
                                 const float *rstd, const void *dres, void *dx, float *dgamma, float *dbeta,
                                 int64_t rows, int W, int x_dtype, int y_dtype, void *stream) {
  if (W % 4 != 0 || W > LN_MAX_ITERS * 128) return fail("layernorm: W=%d unsupported", W);
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define X(xc, yc, XT, YT) \
  if (x_dtype == xc && y_dtype == yc) return ln_bwd_launch<XT, YT>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, W, st);
  LN_COMBOS(X)
#undef X
  return fail("layernorm_bwd: unsupported dtype combination x=%d y=%d", x_dtype, y_dtype);
}

extern "C" int tgt_layernorm_bwd_y(const void *dy, const void *x, const float *gamma, const float *beta, const float *mean,
                                   const float *rstd, const void *dres, void *dx, float *dgamma, float *dbeta, void *y,
                                   int64_t ldy, int64_t rows, int W, int x_dtype, int y_dtype, const float *res_row_scale,
                                   int64_t rows_per_scale, float *dres_colsum, void *stream) {
  if (!y && !dres_colsum) return fail("layernorm_bwd_y: nothing extra to emit (use tgt_layernorm_bwd)");
  if (y && !beta) return fail("layernorm_bwd_y: beta is required to emit y");
  if (rows <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
#define X(xc, yc, XT, YT)                                                                                         \
  if (x_dtype == xc && y_dtype == yc)                                                                             \
    return ln_bwd_launch<XT, YT>(dy, x, gamma, mean, rstd, dres, dx, dgamma, dbeta, rows, W, st, beta, y, ldy,    \
                                 res_row_scale, rows_per_scale, dres_colsum);
  LN_COMBOS(X)
#undef X
  return fail("layernorm_bwd_y: unsupported dtype combination x=%d y=%d", x_dtype, y_dtype);
}

extern "C" int tgt_gelu_dropout_fwd(const void *u, void *y, int64_t n, float p_drop, uint64_t seed, int dtype,
                                    void *stream) {
  if (n <= 0) return 0;
  if (p_drop < 0.f || p_drop >= 1.f) return fail("gelu_dropout: p_drop=%f out of range", p_drop);
  cudaStream_t st = (cudaStream_t)stream;
  TGT_DISPATCH_DTYPE(dtype, T, {
    gelu_dropout_kernel<T, false><<<grid_for(n / (16 / sizeof(T)) + 1, 256 * 4), 256, 0, st>>>(
        (const T *)u, nullptr, (T *)y, n, p_drop, seed, nullptr);
  });
  return check_launch("gelu_dropout_fwd");
}
extern "C" int tgt_gelu_dropout_fwd_dseed(const void *u, void *y, int64_t n, float p_drop, const uint64_t *seed_ptr,
                                          int dtype, void *stream) {
  if (n <= 0) return 0;
  if (p_drop < 0.f || p_drop >= 1.f) return fail("gelu_dropout: p_drop=%f out of range", p_drop);
  if (!seed_ptr) return fail("gelu_dropout_fwd_dseed: null seed pointer");
  cudaStream_t st = (cudaStream_t)stream;
  TGT_DISPATCH_DTYPE(dtype, T, {
    gelu_dropout_kernel<T, false><<<grid_for(n / (16 / sizeof(T)) + 1, 256 * 4), 256, 0, st>>>(
        (const T *)u, nullptr, (T *)y, n, p_drop, 0, seed_ptr);
  });
  return check_launch("gelu_dropout_fwd_dseed");
}
extern "C" int tgt_gelu_dropout_bwd(const void *u, const void *dy, void *du, int64_t n, float p_drop, uint64_t seed,
                                    int dtype, void *stream) {
  if (n <= 0) return 0;
  if (p_drop < 0.f || p_drop >= 1.f) return fail("gelu_dropout: p_drop=%f out of range", p_drop);
  cudaStream_t st = (cudaStream_t)stream;
  TGT_DISPATCH_DTYPE(dtype, T, {
    gelu_dropout_kernel<T, true><<<grid_for(n / (16 / sizeof(T)) + 1, 256 * 4), 256, 0, st>>>(
        (const T *)u, (const T *)dy, (T *)du, n, p_drop, seed, nullptr);
  });
  return check_launch("gelu_dropout_bwd");
}

extern "C" int tgt_scaled_residual(const void *x, const void *res, const float *scale, void *out, int64_t B,
                                   int64_t inner, int dtype, int res_dtype, void *stream) {
  if (B <= 0 || inner <= 0) return 0;
  if (inner % 4 != 0) return fail("scaled_residual: inner=%lld must be a multiple of 4", (long long)inner);
  cudaStream_t st = (cudaStream_t)stream;
  const int g = grid_for(B * inner / 4, 256 * 4);
  if (dtype == res_dtype) {
    TGT_DISPATCH_DTYPE(dtype, T, {
      scaled_residual_kernel<T, T><<<g, 256, 0, st>>>((const T *)x, (const T *)res, scale, (T *)out, B, inner);
    });
  } else if (res_dtype == TGT_F32) {
    TGT_DISPATCH_DTYPE(dtype, T, {
      scaled_residual_kernel<T, float><<<g, 256, 0, st>>>((const T *)x, (const float *)res, scale, (T *)out, B, inner);
    });
  } else {
    return fail("scaled_residual: unsupported dtype combination %d/%d", dtype, res_dtype);
  }
  return check_launch("scaled_residual");
}

template <typename T>
static int xent_launch(bool bwd, const void *logits, int64_t ld, const int64_t *target, float *xent, float *lse, const float *gx,
                       void *dlogits, int64_t rows, int nbins, cudaStream_t st) {
  const int nch = (nbins + 255) / 256;
  const int g = grid_for(rows, 8 * 4);
#define X(N)                                                                                                             \
  do {                                                                                                                   \
    if (bwd)                                                                                                             \
      xent_rows_bwd<T, N><<<g, 256, 0, st>>>((const T *)logits, ld, target, lse, gx, (T *)dlogits, rows, nbins);           \
    else                                                                                                                 \
      xent_rows_fwd<T, N><<<g, 256, 0, st>>>((const T *)logits, ld, target, xent, lse, rows, nbins);                      \
  } while (0)
  switch (nch) {
    case 1: X(1); break;
    case 2: X(2); break;
    case 3: X(3); break;
    default: X(4); break;
  }
#undef X
  return check_launch(bwd ? "xent_rows_bwd" : "xent_rows_fwd");
}
static int xent_check(const void *logits, int64_t ld, int64_t rows, int nbins) {
  if (nbins <= 0 || nbins % 8 || nbins > 1024) return fail("xent_rows: nbins=%d must be a multiple of 8 and <= 1024", nbins);
  if (ld % 8 || ld < nbins || (reinterpret_cast<uintptr_t>(logits) & 15)) return fail("xent_rows: logits rows must be 16-byte aligned");
  (void)rows;
  return 0;
}
extern "C" int tgt_xent_rows_fwd(const void *logits, int64_t ld, const int64_t *target, float *xent, float *lse, int64_t rows,
                                 int nbins, int dtype, void *stream) {
  if (rows <= 0) return 0;
  if (!logits || !target || !xent || !lse) return fail("xent_rows_fwd: null argument");
  if (int e = xent_check(logits, ld, rows, nbins)) return e;
  TGT_DISPATCH_DTYPE(dtype, T, return xent_launch<T>(false, logits, ld, target, xent, lse, nullptr, nullptr, rows, nbins,
                                                     (cudaStream_t)stream));
  return 0;
}
extern "C" int tgt_xent_rows_bwd(const void *logits, int64_t ld, const int64_t *target, const float *lse, const float *gx,
                                 void *dlogits, int64_t rows, int nbins, int dtype, void *stream) {
  if (rows <= 0) return 0;
  if (!logits || !target || !lse || !gx || !dlogits) return fail("xent_rows_bwd: null argument");
  if (int e = xent_check(logits, ld, rows, nbins)) return e;
  if (reinterpret_cast<uintptr_t>(dlogits) & 15) return fail("xent_rows_bwd: dlogits must be 16-byte aligned");
  TGT_DISPATCH_DTYPE(dtype, T, return xent_launch<T>(true, logits, ld, target, nullptr, const_cast<float *>(lse), gx, dlogits,
                                                     rows, nbins, (cudaStream_t)stream));
  return 0;
}

extern "C" int tgt_gaussian_basis_fwd(const float *x, const float *mu, const float *sd, void *out, int64_t rows, int K,
                                      int out_dtype, void *stream) {
  if (rows <= 0) return 0;
  if (K <= 0 || K > 128 || K % 4) return fail("gaussian_basis: K=%d must be a multiple of 4 and <= 128", K);
  cudaStream_t st = (cudaStream_t)stream;
  TGT_DISPATCH_DTYPE(out_dtype, T,
                     (gaussian_basis_fwd_kernel<T><<<grid_for(rows, 8 * 8), 256, 0, st>>>(x, mu, sd, (T *)out, rows, K)));
  return check_launch("gaussian_basis_fwd");
}

extern "C" int tgt_gaussian_basis_bwd(const float *x, const float *mu, const float *sd, const void *dout, float *dx,
                                      float *dmu, float *dsd, int64_t rows, int K, int dout_dtype, void *stream) {
  if (rows <= 0) return 0;
  if (K <= 0 || K > 128 || K % 4) return fail("gaussian_basis: K=%d must be a multiple of 4 and <= 128", K);
  cudaStream_t st = (cudaStream_t)stream;
  TGT_DISPATCH_DTYPE(dout_dtype, T,
                     (gaussian_basis_bwd_kernel<T><<<grid_for(rows, 8 * 8), 256, 0, st>>>(x, mu, sd, (const T *)dout, dx,
                                                                                      dmu, dsd, rows, K)));
  return check_launch("gaussian_basis_bwd");
}
