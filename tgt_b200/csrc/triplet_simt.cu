// Generic (any dtype, any head-dim <= 64, N <= 64) triplet attention / aggregate cores on CUDA cores.
// These are the fp32 parity kernels and the shape-general path; the bf16 tensor-core kernels live in
// triplet_mma.cu.  fp32 math throughout; nothing O(N^3) is ever written to memory.
//
// Index conventions (reference lib/tgt/layers/triplet.py:205-250):
//   inward : query row (b,i,j)   key/value row (b,j,k)   bias/gate/mask row (b,i,k)
//   outward: query row (b,i,j)   key/value row (b,k,j)   bias/gate/mask row (b,k,i)
#include "common.cuh"

namespace tgt {

constexpr int SIMT_MAX_N = 64;

template <bool ACC>
__device__ __forceinline__ float sigm(float x) { return ACC ? sigmoid_acc(x) : sigmoidf_(x); }

// ------------------------------------------------------------------------------------------
// forward: one block per (j, h, b); thread i = query.  K/V rows staged in shared memory.
// ------------------------------------------------------------------------------------------
template <typename T, int DMAX>
__global__ void __launch_bounds__(SIMT_MAX_N)
tri_attn_fwd_simt(const tgt_triplet_attn_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                  T *__restrict__ va, float *__restrict__ stats) {
  extern __shared__ float sm[];
  const int N = D.N, H = D.H, d = D.d;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z, i = threadIdx.x;
  float *Ks = sm;                 // [N][d]
  float *Vs = sm + N * d;         // [N][d]
  const int64_t ld = D.ld;
  const int64_t ldva = 2 * (int64_t)H * d;
  constexpr bool ACC = sizeof(T) == 4;

  for (int dir = 0; dir < 2; ++dir) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * d; idx += blockDim.x) {
      const int k = idx / d, dd = idx - k * d;
      const int64_t row = dir == 0 ? ((int64_t)(b * N + j) * N + k) : ((int64_t)(b * N + k) * N + j);
      Ks[idx] = to_f(proj[row * ld + D.off_k[dir] + h * d + dd]);
      Vs[idx] = to_f(proj[row * ld + D.off_v[dir] + h * d + dd]);
    }
    __syncthreads();
    if (i < N) {
      const int64_t qrow = (int64_t)(b * N + i) * N + j;
      float q[DMAX], o[DMAX];
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd) {
        q[dd] = dd < d ? to_f(proj[qrow * ld + D.off_q[dir] + h * d + dd]) * D.scale : 0.f;
        o[dd] = 0.f;
      }
      float m = -INFINITY, l = 0.f;
      for (int k = 0; k < N; ++k) {
        const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
        float s = 0.f;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) s += q[dd] * Ks[k * d + dd];
        const float mk = mask[brow];
        if (D.off_e[dir] >= 0) s += to_f(proj[brow * ld + D.off_e[dir] + h]);
        s += mk;
        float g = 1.f;
        if (D.off_g[dir] >= 0) g = sigm<ACC>(to_f(proj[brow * ld + D.off_g[dir] + h]) + mk);
        const float mn = fmaxf(m, s);
        const float corr = expf(m - mn);
        const float p = expf(s - mn);
        l = l * corr + p;
        const float pg = p * g;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) o[dd] = o[dd] * corr + pg * Vs[k * d + dd];
        m = mn;
      }
      const float inv = 1.f / l;
      T *out = va + qrow * ldva + (int64_t)dir * H * d + h * d;
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd)
        if (dd < d) out[dd] = from_f<T>(o[dd] * inv);
      float *st = stats + ((((int64_t)(b * 2 + dir) * H + h) * N + j) * N + i) * 2;
      st[0] = m;
      st[1] = inv;
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: one block per (h, b), loops over direction and junction j so that dE/dG (which
// reduce over j) accumulate in shared memory without atomics.  Phase 1: thread = query i,
// phase 2: thread = key k.  Deterministic.
// ------------------------------------------------------------------------------------------
template <typename T, int DMAX>
__global__ void __launch_bounds__(SIMT_MAX_N)
tri_attn_bwd_simt(const tgt_triplet_attn_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                  const T *__restrict__ va, const T *__restrict__ dva, const float *__restrict__ stats,
                  T *__restrict__ dproj) {
  extern __shared__ float sm[];
  const int N = D.N, H = D.H, d = D.d;
  const int h = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
  const int NP = N + 1;
  float *Ks = sm;                   // [N][d]
  float *Vs = Ks + N * d;           // [N][d]
  float *Qs = Vs + N * d;           // [N][d]  (pre-scaled queries)
  float *dOs = Qs + N * d;          // [N][d]
  float *As = dOs + N * d;          // [N][NP]  A[i][k] = P*g
  float *dSs = As + N * NP;         // [N][NP]
  float *dEa = dSs + N * NP;        // [N][NP]  sum over j of dS
  float *dGa = dEa + N * NP;        // [N][NP]
  const int64_t ld = D.ld;
  const int64_t ldva = 2 * (int64_t)H * d;
  constexpr bool ACC = sizeof(T) == 4;

  for (int dir = 0; dir < 2; ++dir) {
    for (int idx = t; idx < N * NP; idx += blockDim.x) { dEa[idx] = 0.f; dGa[idx] = 0.f; }
    const bool has_e = D.off_e[dir] >= 0, has_g = D.off_g[dir] >= 0;
    for (int j = 0; j < N; ++j) {
      __syncthreads();
      for (int idx = t; idx < N * d; idx += blockDim.x) {
        const int k = idx / d, dd = idx - k * d;
        const int64_t krow = dir == 0 ? ((int64_t)(b * N + j) * N + k) : ((int64_t)(b * N + k) * N + j);
        const int64_t qrow = (int64_t)(b * N + k) * N + j;      // here k plays the role of i
        Ks[idx] = to_f(proj[krow * ld + D.off_k[dir] + h * d + dd]);
        Vs[idx] = to_f(proj[krow * ld + D.off_v[dir] + h * d + dd]);
        Qs[idx] = to_f(proj[qrow * ld + D.off_q[dir] + h * d + dd]) * D.scale;
        dOs[idx] = to_f(dva[qrow * ldva + (int64_t)dir * H * d + h * d + dd]);
      }
      __syncthreads();
      // ---- phase 1: thread = query i
      if (t < N) {
        const int i = t;
        const int64_t qrow = (int64_t)(b * N + i) * N + j;
        float q[DMAX], go[DMAX], dq[DMAX];
        float delta = 0.f;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd) {
          q[dd] = dd < d ? Qs[i * d + dd] : 0.f;
          go[dd] = dd < d ? dOs[i * d + dd] : 0.f;
          dq[dd] = 0.f;
          if (dd < d) delta += go[dd] * to_f(va[qrow * ldva + (int64_t)dir * H * d + h * d + dd]);
        }
        const float *st = stats + ((((int64_t)(b * 2 + dir) * H + h) * N + j) * N + i) * 2;
        const float m = st[0], inv = st[1];
        for (int k = 0; k < N; ++k) {
          const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
          float s = 0.f, dA = 0.f;
#pragma unroll
          for (int dd = 0; dd < DMAX; ++dd)
            if (dd < d) { s += q[dd] * Ks[k * d + dd]; dA += go[dd] * Vs[k * d + dd]; }
          const float mk = mask[brow];
          if (has_e) s += to_f(proj[brow * ld + D.off_e[dir] + h]);
          s += mk;
          float g = 1.f;
          if (has_g) g = sigm<ACC>(to_f(proj[brow * ld + D.off_g[dir] + h]) + mk);
          const float p = expf(s - m) * inv;
          const float dS = p * (dA * g - delta);
          As[i * NP + k] = p * g;
          dSs[i * NP + k] = dS;
          dEa[i * NP + k] += dS;
          dGa[i * NP + k] += g * (1.f - g) * dA * p;
#pragma unroll
          for (int dd = 0; dd < DMAX; ++dd)
            if (dd < d) dq[dd] += dS * Ks[k * d + dd];
        }
        T *o = dproj + qrow * ld + D.off_q[dir] + h * d;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) o[dd] = from_f<T>(dq[dd] * D.scale);
      }
      __syncthreads();
      // ---- phase 2: thread = key k
      if (t < N) {
        const int k = t;
        float dk[DMAX], dv[DMAX];
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd) { dk[dd] = 0.f; dv[dd] = 0.f; }
        for (int i = 0; i < N; ++i) {
          const float a = As[i * NP + k], ds = dSs[i * NP + k];
#pragma unroll
          for (int dd = 0; dd < DMAX; ++dd)
            if (dd < d) { dv[dd] += a * dOs[i * d + dd]; dk[dd] += ds * Qs[i * d + dd]; }
        }
        const int64_t krow = dir == 0 ? ((int64_t)(b * N + j) * N + k) : ((int64_t)(b * N + k) * N + j);
        T *ok = dproj + krow * ld + D.off_k[dir] + h * d;
        T *ov = dproj + krow * ld + D.off_v[dir] + h * d;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) { ok[dd] = from_f<T>(dk[dd]); ov[dd] = from_f<T>(dv[dd]); }
      }
    }
    __syncthreads();
    if (has_e || has_g) {
      for (int idx = t; idx < N * N; idx += blockDim.x) {
        const int i = idx / N, k = idx - i * N;
        const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
        if (has_e) dproj[brow * ld + D.off_e[dir] + h] = from_f<T>(dEa[i * NP + k]);
        if (has_g) dproj[brow * ld + D.off_g[dir] + h] = from_f<T>(dGa[i * NP + k]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------
// aggregate: weights kernel  aw[b,dir,h,i,k]  (query i, key k)
//   dir 0: softmax_k(E_in[b,i,k,h] (+M[b,i,k])) * sigmoid(G_in[b,i,k,h] (+M))
//   dir 1: softmax over k of E_out[b,k,i,h] (+M[b,k,i])  * sigmoid(G_out[b,k,i,h] (+M))
// one block per (h, dir, b); thread i.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(SIMT_MAX_N)
tri_aggr_weights(const tgt_triplet_aggr_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                 float *__restrict__ aw) {
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z, i = threadIdx.x;
  if (i >= N) return;
  constexpr bool ACC = sizeof(T) == 4;
  const int64_t ld = D.ld;
  const bool use_m = D.mask_dir[dir] != 0, has_g = D.off_g[dir] >= 0;
  float m = -INFINITY;
  for (int k = 0; k < N; ++k) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    float s = to_f(proj[brow * ld + D.off_e[dir] + h]);
    if (use_m) s += mask[brow];
    m = fmaxf(m, s);
  }
  float l = 0.f;
  for (int k = 0; k < N; ++k) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    float s = to_f(proj[brow * ld + D.off_e[dir] + h]);
    if (use_m) s += mask[brow];
    l += expf(s - m);
  }
  const float inv = 1.f / l;
  float *out = aw + (((int64_t)(b * 2 + dir) * H + h) * N + i) * N;
  for (int k = 0; k < N; ++k) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    const float mk = use_m ? mask[brow] : 0.f;
    const float s = to_f(proj[brow * ld + D.off_e[dir] + h]) + mk;
    float g = 1.f;
    if (has_g) g = sigm<ACC>(to_f(proj[brow * ld + D.off_g[dir] + h]) + mk);
    out[k] = expf(s - m) * inv * g;
  }
}

// aggregate apply: Va[b,i,j,dir,h,:] = sum_k aw[b,dir,h,i,k] * V[keyrow(j,k)][h,:];  block (j,h,b), thread i
template <typename T, int DMAX>
__global__ void __launch_bounds__(SIMT_MAX_N)
tri_aggr_apply(const tgt_triplet_aggr_desc D, const T *__restrict__ proj, const float *__restrict__ aw,
               T *__restrict__ va) {
  extern __shared__ float sm[];
  const int N = D.N, H = D.H, d = D.d;
  const int j = blockIdx.x, h = blockIdx.y, b = blockIdx.z, i = threadIdx.x;
  float *Vs = sm;
  const int64_t ld = D.ld, ldva = 2 * (int64_t)H * d;
  for (int dir = 0; dir < 2; ++dir) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * d; idx += blockDim.x) {
      const int k = idx / d, dd = idx - k * d;
      const int64_t row = dir == 0 ? ((int64_t)(b * N + j) * N + k) : ((int64_t)(b * N + k) * N + j);
      Vs[idx] = to_f(proj[row * ld + D.off_v[dir] + h * d + dd]);
    }
    __syncthreads();
    if (i < N) {
      float o[DMAX];
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd) o[dd] = 0.f;
      const float *a = aw + (((int64_t)(b * 2 + dir) * H + h) * N + i) * N;
      for (int k = 0; k < N; ++k) {
        const float w = a[k];
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) o[dd] += w * Vs[k * d + dd];
      }
      T *out = va + ((int64_t)(b * N + i) * N + j) * ldva + (int64_t)dir * H * d + h * d;
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd)
        if (dd < d) out[dd] = from_f<T>(o[dd]);
    }
  }
}

// aggregate backward part 1: block (h, dir, b) loops over j; thread k = key.
//   dV[keyrow(j,k)] = sum_i aw[i,k] * dO[(i,j)]        daw[i,k] = sum_j dO[(i,j)] . V[keyrow(j,k)]
template <typename T, int DMAX>
__global__ void __launch_bounds__(SIMT_MAX_N)
tri_aggr_bwd_apply(const tgt_triplet_aggr_desc D, const T *__restrict__ proj, const T *__restrict__ dva,
                   const float *__restrict__ aw, float *__restrict__ daw, T *__restrict__ dproj) {
  extern __shared__ float sm[];
  const int N = D.N, H = D.H, d = D.d;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z, k = threadIdx.x;
  const int NP = N + 1;
  float *dOs = sm;                // [N][d]
  float *Aw = dOs + N * d;        // [N][NP]
  float *dA = Aw + N * NP;        // [N][NP]
  const int64_t ld = D.ld, ldva = 2 * (int64_t)H * d;
  const float *a = aw + ((int64_t)(b * 2 + dir) * H + h) * N * N;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    const int i = idx / N, kk = idx - i * N;
    Aw[i * NP + kk] = a[idx];
    dA[i * NP + kk] = 0.f;
  }
  for (int j = 0; j < N; ++j) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < N * d; idx += blockDim.x) {
      const int i = idx / d, dd = idx - i * d;
      dOs[idx] = to_f(dva[((int64_t)(b * N + i) * N + j) * ldva + (int64_t)dir * H * d + h * d + dd]);
    }
    __syncthreads();
    if (k < N) {
      const int64_t krow = dir == 0 ? ((int64_t)(b * N + j) * N + k) : ((int64_t)(b * N + k) * N + j);
      float v[DMAX], dv[DMAX];
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd) {
        v[dd] = dd < d ? to_f(proj[krow * ld + D.off_v[dir] + h * d + dd]) : 0.f;
        dv[dd] = 0.f;
      }
      for (int i = 0; i < N; ++i) {
        const float w = Aw[i * NP + k];
        float acc = 0.f;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) { dv[dd] += w * dOs[i * d + dd]; acc += dOs[i * d + dd] * v[dd]; }
        dA[i * NP + k] += acc;
      }
      T *ov = dproj + krow * ld + D.off_v[dir] + h * d;
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd)
        if (dd < d) ov[dd] = from_f<T>(dv[dd]);
    }
  }
  __syncthreads();
  float *o = daw + ((int64_t)(b * 2 + dir) * H + h) * N * N;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    const int i = idx / N, kk = idx - i * N;
    o[idx] = dA[i * NP + kk];
  }
}

// aggregate backward part 2: softmax/gate backward per (h,dir,b), thread i (row of the softmax).
template <typename T>
__global__ void __launch_bounds__(SIMT_MAX_N)
tri_aggr_bwd_weights(const tgt_triplet_aggr_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                     const float *__restrict__ daw, T *__restrict__ dproj) {
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z, i = threadIdx.x;
  if (i >= N) return;
  constexpr bool ACC = sizeof(T) == 4;
  const int64_t ld = D.ld;
  const bool use_m = D.mask_dir[dir] != 0, has_g = D.off_g[dir] >= 0;
  const float *da = daw + (((int64_t)(b * 2 + dir) * H + h) * N + i) * N;
  float m = -INFINITY;
  for (int k = 0; k < N; ++k) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    float s = to_f(proj[brow * ld + D.off_e[dir] + h]);
    if (use_m) s += mask[brow];
    m = fmaxf(m, s);
  }
  float l = 0.f;
  for (int k = 0; k < N; ++k) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    float s = to_f(proj[brow * ld + D.off_e[dir] + h]);
    if (use_m) s += mask[brow];
    l += expf(s - m);
  }
  const float inv = 1.f / l;
  float delta = 0.f;
  for (int k = 0; k < N; ++k) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    const float mk = use_m ? mask[brow] : 0.f;
    const float p = expf(to_f(proj[brow * ld + D.off_e[dir] + h]) + mk - m) * inv;
    float g = 1.f;
    if (has_g) g = sigm<ACC>(to_f(proj[brow * ld + D.off_g[dir] + h]) + mk);
    delta += da[k] * g * p;
  }
  for (int k = 0; k < N; ++k) {
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    const float mk = use_m ? mask[brow] : 0.f;
    const float p = expf(to_f(proj[brow * ld + D.off_e[dir] + h]) + mk - m) * inv;
    float g = 1.f;
    if (has_g) g = sigm<ACC>(to_f(proj[brow * ld + D.off_g[dir] + h]) + mk);
    dproj[brow * ld + D.off_e[dir] + h] = from_f<T>(p * (da[k] * g - delta));
    if (has_g) dproj[brow * ld + D.off_g[dir] + h] = from_f<T>(da[k] * p * g * (1.f - g));
  }
}

// ------------------------------------------------------------------------------------------ host
template <typename T>
static int attn_fwd_simt(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, void *va,
                         float *stats, cudaStream_t st) {
  dim3 grid(D.N, D.H, D.B);
  const size_t smem = 2 * (size_t)D.N * D.d * sizeof(float);
#define L(DM)                                                                                            \
  tri_attn_fwd_simt<T, DM><<<grid, SIMT_MAX_N, smem, st>>>(D, (const T *)proj, mask, (T *)va, stats)
  if (D.d <= 16) L(16); else if (D.d <= 32) L(32); else L(64);
#undef L
  return check_launch("tri_attn_fwd_simt");
}

template <typename T>
static int attn_bwd_simt(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, const void *va,
                         const void *dva, const float *stats, void *dproj, cudaStream_t st) {
  dim3 grid(D.H, D.B);
  const size_t smem = (4 * (size_t)D.N * D.d + 4 * (size_t)D.N * (D.N + 1)) * sizeof(float);
#define L(DM)                                                                                              \
  do {                                                                                                     \
    TGT_CUDA_OK(cudaFuncSetAttribute(tri_attn_bwd_simt<T, DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                          \
    tri_attn_bwd_simt<T, DM><<<grid, SIMT_MAX_N, smem, st>>>(D, (const T *)proj, mask, (const T *)va,      \
                                                             (const T *)dva, stats, (T *)dproj);           \
  } while (0)
  if (D.d <= 16) L(16); else if (D.d <= 32) L(32); else L(64);
#undef L
  return check_launch("tri_attn_bwd_simt");
}

int triplet_attn_fwd_simt(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, void *va,
                          float *stats, cudaStream_t st) {
  TGT_DISPATCH_DTYPE(D.dtype, T, return attn_fwd_simt<T>(D, proj, mask, va, stats, st));
  return 0;
}
int triplet_attn_bwd_simt(const tgt_triplet_attn_desc &D, const void *proj, const float *mask, const void *va,
                          const void *dva, const float *stats, void *dproj, cudaStream_t st) {
  TGT_DISPATCH_DTYPE(D.dtype, T, return attn_bwd_simt<T>(D, proj, mask, va, dva, stats, dproj, st));
  return 0;
}

template <typename T>
static int aggr_fwd_simt(const tgt_triplet_aggr_desc &D, const void *proj, const float *mask, void *va,
                         float *aw, cudaStream_t st) {
  tri_aggr_weights<T><<<dim3(D.H, 2, D.B), SIMT_MAX_N, 0, st>>>(D, (const T *)proj, mask, aw);
  if (int e = check_launch("tri_aggr_weights")) return e;
  dim3 grid(D.N, D.H, D.B);
  const size_t smem = (size_t)D.N * D.d * sizeof(float);
#define L(DM) tri_aggr_apply<T, DM><<<grid, SIMT_MAX_N, smem, st>>>(D, (const T *)proj, aw, (T *)va)
  if (D.d <= 16) L(16); else if (D.d <= 32) L(32); else L(64);
#undef L
  return check_launch("tri_aggr_apply");
}

template <typename T>
static int aggr_bwd_simt(const tgt_triplet_aggr_desc &D, const void *proj, const float *mask, const void *dva,
                         const float *aw, float *daw, void *dproj, cudaStream_t st) {
  dim3 grid(D.H, 2, D.B);
  const size_t smem = ((size_t)D.N * D.d + 2 * (size_t)D.N * (D.N + 1)) * sizeof(float);
#define L(DM)                                                                                               \
  do {                                                                                                      \
    TGT_CUDA_OK(cudaFuncSetAttribute(tri_aggr_bwd_apply<T, DM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)smem));                                                           \
    tri_aggr_bwd_apply<T, DM><<<grid, SIMT_MAX_N, smem, st>>>(D, (const T *)proj, (const T *)dva, aw, daw,  \
                                                              (T *)dproj);                                  \
  } while (0)
  if (D.d <= 16) L(16); else if (D.d <= 32) L(32); else L(64);
#undef L
  if (int e = check_launch("tri_aggr_bwd_apply")) return e;
  tri_aggr_bwd_weights<T><<<grid, SIMT_MAX_N, 0, st>>>(D, (const T *)proj, mask, daw, (T *)dproj);
  return check_launch("tri_aggr_bwd_weights");
}

// ------------------------------------------------------------------------------------------
// aggregate weights, row-per-CTA versions: block (i, dir, b) reads the E | G values of its N edge rows as contiguous
// H-wide segments (one 32-byte sector per row and block instead of 2-byte gathers per thread), transposes through shared
// memory, and one warp per head runs the softmax over k with warp reductions.  Same arithmetic as tri_aggr_weights /
// tri_aggr_bwd_weights (which remain the generic path for H > 32).
// ------------------------------------------------------------------------------------------
constexpr int AGW_LD = SIMT_MAX_N + 1;

template <typename T>
__global__ void __launch_bounds__(256)
tri_aggr_weights_rows(const tgt_triplet_aggr_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                      float *__restrict__ aw) {
  extern __shared__ float sm[];
  const int N = D.N, H = D.H;
  const int i = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  constexpr bool ACC = sizeof(T) == 4;
  const int64_t ld = D.ld;
  const bool use_m = D.mask_dir[dir] != 0, has_g = D.off_g[dir] >= 0;
  float *se = sm, *sg = sm + H * AGW_LD;
  for (int idx = threadIdx.x; idx < N * H; idx += blockDim.x) {
    const int k = idx / H, h = idx - k * H;
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    const float mk = use_m ? mask[brow] : 0.f;
    se[h * AGW_LD + k] = to_f(proj[brow * ld + D.off_e[dir] + h]) + mk;
    sg[h * AGW_LD + k] = has_g ? sigm<ACC>(to_f(proj[brow * ld + D.off_g[dir] + h]) + mk) : 1.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int h = threadIdx.x >> 5; h < H; h += blockDim.x >> 5) {
    const float s0 = lane < N ? se[h * AGW_LD + lane] : -INFINITY;
    const float s1 = lane + 32 < N ? se[h * AGW_LD + lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(s0, s1));
    const float p0 = lane < N ? expf(s0 - m) : 0.f, p1 = lane + 32 < N ? expf(s1 - m) : 0.f;
    const float inv = 1.f / warp_sum(p0 + p1);
    float *out = aw + (((int64_t)(b * 2 + dir) * H + h) * N + i) * N;
    if (lane < N) out[lane] = p0 * inv * sg[h * AGW_LD + lane];
    if (lane + 32 < N) out[lane + 32] = p1 * inv * sg[h * AGW_LD + lane + 32];
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
tri_aggr_bwd_weights_rows(const tgt_triplet_aggr_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                          const float *__restrict__ daw, T *__restrict__ dproj) {
  extern __shared__ float sm[];
  const int N = D.N, H = D.H;
  const int i = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  constexpr bool ACC = sizeof(T) == 4;
  const int64_t ld = D.ld;
  const bool use_m = D.mask_dir[dir] != 0, has_g = D.off_g[dir] >= 0;
  float *se = sm, *sg = sm + H * AGW_LD;           // logits -> dE ; gates -> dG
  for (int idx = threadIdx.x; idx < N * H; idx += blockDim.x) {
    const int k = idx / H, h = idx - k * H;
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    const float mk = use_m ? mask[brow] : 0.f;
    se[h * AGW_LD + k] = to_f(proj[brow * ld + D.off_e[dir] + h]) + mk;
    sg[h * AGW_LD + k] = has_g ? sigm<ACC>(to_f(proj[brow * ld + D.off_g[dir] + h]) + mk) : 1.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  for (int h = threadIdx.x >> 5; h < H; h += blockDim.x >> 5) {
    const float *da = daw + (((int64_t)(b * 2 + dir) * H + h) * N + i) * N;
    const bool v0 = lane < N, v1 = lane + 32 < N;
    const float s0 = v0 ? se[h * AGW_LD + lane] : -INFINITY, s1 = v1 ? se[h * AGW_LD + lane + 32] : -INFINITY;
    const float m = warp_max(fmaxf(s0, s1));
    float p0 = v0 ? expf(s0 - m) : 0.f, p1 = v1 ? expf(s1 - m) : 0.f;
    const float inv = 1.f / warp_sum(p0 + p1);
    p0 *= inv;
    p1 *= inv;
    const float g0 = v0 ? sg[h * AGW_LD + lane] : 0.f, g1 = v1 ? sg[h * AGW_LD + lane + 32] : 0.f;
    const float d0 = v0 ? da[lane] : 0.f, d1 = v1 ? da[lane + 32] : 0.f;
    const float delta = warp_sum(d0 * g0 * p0 + d1 * g1 * p1);
    if (v0) {
      se[h * AGW_LD + lane] = p0 * (d0 * g0 - delta);
      sg[h * AGW_LD + lane] = d0 * p0 * g0 * (1.f - g0);
    }
    if (v1) {
      se[h * AGW_LD + lane + 32] = p1 * (d1 * g1 - delta);
      sg[h * AGW_LD + lane + 32] = d1 * p1 * g1 * (1.f - g1);
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * H; idx += blockDim.x) {
    const int k = idx / H, h = idx - k * H;
    const int64_t brow = dir == 0 ? ((int64_t)(b * N + i) * N + k) : ((int64_t)(b * N + k) * N + i);
    dproj[brow * ld + D.off_e[dir] + h] = from_f<T>(se[h * AGW_LD + k]);
    if (has_g) dproj[brow * ld + D.off_g[dir] + h] = from_f<T>(sg[h * AGW_LD + k]);
  }
}

// the O(N^2 H) halves alone (softmax / gate weights and their backward): the tensor-core apply kernels of triplet_tma.cu
// sit between them
int triplet_aggr_weights_simt(const tgt_triplet_aggr_desc &D, const void *proj, const float *mask, float *aw,
                              cudaStream_t st) {
  if (D.H <= 32) {
    const size_t smem = 2 * (size_t)D.H * AGW_LD * sizeof(float);
    TGT_DISPATCH_DTYPE(D.dtype, T, tri_aggr_weights_rows<T><<<dim3(D.N, 2, D.B), 256, smem, st>>>(D, (const T *)proj, mask, aw));
    return check_launch("tri_aggr_weights_rows");
  }
  TGT_DISPATCH_DTYPE(D.dtype, T,
                     tri_aggr_weights<T><<<dim3(D.H, 2, D.B), SIMT_MAX_N, 0, st>>>(D, (const T *)proj, mask, aw));
  return check_launch("tri_aggr_weights");
}
int triplet_aggr_bwd_weights_simt(const tgt_triplet_aggr_desc &D, const void *proj, const float *mask, const float *daw,
                                  void *dproj, cudaStream_t st) {
  if (D.H <= 32) {
    const size_t smem = 2 * (size_t)D.H * AGW_LD * sizeof(float);
    TGT_DISPATCH_DTYPE(D.dtype, T, tri_aggr_bwd_weights_rows<T><<<dim3(D.N, 2, D.B), 256, smem, st>>>(
                                       D, (const T *)proj, mask, daw, (T *)dproj));
    return check_launch("tri_aggr_bwd_weights_rows");
  }
  TGT_DISPATCH_DTYPE(D.dtype, T, tri_aggr_bwd_weights<T><<<dim3(D.H, 2, D.B), SIMT_MAX_N, 0, st>>>(
                                     D, (const T *)proj, mask, daw, (T *)dproj));
  return check_launch("tri_aggr_bwd_weights");
}

int triplet_aggr_fwd_simt(const tgt_triplet_aggr_desc &D, const void *proj, const float *mask, void *va,
                          float *aw, cudaStream_t st) {
  TGT_DISPATCH_DTYPE(D.dtype, T, return aggr_fwd_simt<T>(D, proj, mask, va, aw, st));
  return 0;
}
int triplet_aggr_bwd_simt(const tgt_triplet_aggr_desc &D, const void *proj, const float *mask, const void *dva,
                          const float *aw, float *daw, void *dproj, cudaStream_t st) {
  TGT_DISPATCH_DTYPE(D.dtype, T, return aggr_bwd_simt<T>(D, proj, mask, dva, aw, daw, dproj, st));
  return 0;
}

}  // namespace tgt
