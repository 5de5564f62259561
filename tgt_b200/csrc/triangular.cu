// TriangularUpdate core (reference lib/tgt/layers/triplet.py:154-170) and the sigmoid(gate)*lin output gate (172-175).
//
// Every branch of the reference is siglin(g + mask, l) = sigmoid(g + mask) * l on an O(N^2 H) tensor followed by a
// per-(graph, head) N x N x N contraction
//     Va_in [i,j] = sum_k E_in [i,k] * V_in [j,k]          (triplet.py:166, 'bikh,bjkh->bijh')
//     Va_out[i,j] = sum_k E_out[k,i] * V_out[k,j]          (triplet.py:167, 'bkih,bkjh->bijh')
// One CTA owns (head, graph): the four gated N x N operand tiles are built ONCE in shared memory (fp32) straight from
// the projection rows -- the gated tensors never exist in HBM -- and both directions are contracted from them.  The
// work is 4*B*N^3*H flops on 8*R*H inputs (H = heads, not channels): latency/HBM trivial next to the attention
// variants, so this is a plain SIMT kernel (no tensor cores: K = N <= 64 per 64x64 tile and H tiles per graph).
// Backward recomputes the gated tiles and emits d(proj) for all eight column blocks in one pass.
#include "common.cuh"

namespace tgt {

constexpr int TRI_MAX_N = 64;
constexpr int TRI_LD = TRI_MAX_N + 1;                 // +1: conflict-free column walks
constexpr int TRI_TILE = TRI_MAX_N * TRI_LD;
constexpr int TRI_THREADS = 256;

// gated operand tiles of (b, h): t[0] = V_in, t[1] = V_out, t[2] = E_in, t[3] = E_out, indexed [x][y] like the edge
// tensor; entries outside N x N are zero so the contractions may run over the full tile.
template <typename T>
__device__ __forceinline__ void tri_load_tiles(const tgt_triangular_desc &D, const T *__restrict__ proj,
                                               const float *__restrict__ mask, int b, int h, float *t) {
  const int N = D.N, H = D.H;
  for (int idx = threadIdx.x; idx < TRI_MAX_N * TRI_MAX_N; idx += TRI_THREADS) {
    const int x = idx / TRI_MAX_N, y = idx % TRI_MAX_N;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (x < N && y < N) {
      const int64_t row = ((int64_t)b * N + x) * N + y;
      const T *p = proj + row * D.ld;
      const float m = mask[row];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int off = (c < 2 ? D.off_v : D.off_e) + (c & 1) * 2 * H + h;      // gate block; lin block follows at +H
        v[c] = sigmoid_acc(to_f(p[off]) + m) * to_f(p[off + H]);
      }
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) t[c * TRI_TILE + x * TRI_LD + y] = v[c];
  }
}

template <typename T>
__global__ void __launch_bounds__(TRI_THREADS)
triangular_fwd_kernel(const tgt_triangular_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                      T *__restrict__ va) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y, N = D.N, H = D.H;
  tri_load_tiles<T>(D, proj, mask, b, h, sm);
  __syncthreads();
  const float *Vin = sm, *Vout = sm + TRI_TILE, *Ein = sm + 2 * TRI_TILE, *Eout = sm + 3 * TRI_TILE;
  // thread -> outputs (i = ti + 16 a, j = tj + 16 c), a, c < 4
  const int ti = threadIdx.x / 16, tj = threadIdx.x % 16;
  float in[4][4] = {}, out[4][4] = {};
  for (int k = 0; k < N; ++k) {
    float ei[4], vi[4], eo[4], vo[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      ei[a] = Ein[(ti + 16 * a) * TRI_LD + k];
      vi[a] = Vin[(tj + 16 * a) * TRI_LD + k];
      eo[a] = Eout[k * TRI_LD + ti + 16 * a];
      vo[a] = Vout[k * TRI_LD + tj + 16 * a];
    }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        in[a][c] = fmaf(ei[a], vi[c], in[a][c]);
        out[a][c] = fmaf(eo[a], vo[c], out[a][c]);
      }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int i = ti + 16 * a, j = tj + 16 * c;
      if (i < N && j < N) {
        T *o = va + (((int64_t)b * N + i) * N + j) * (2 * H);
        o[h] = from_f<T>(in[a][c]);
        o[H + h] = from_f<T>(out[a][c]);
      }
    }
}

// d(proj) of all eight blocks.  With X = sigmoid(g + m) * l:  dl = dX * s,  dg = dX * l * s * (1 - s).
//   dE_in [x,y] = sum_j dVa_in [x,j] V_in [j,y]        dV_in [x,y] = sum_i dVa_in [i,x] E_in [i,y]
//   dE_out[x,y] = sum_j dVa_out[y,j] V_out[x,j]        dV_out[x,y] = sum_i dVa_out[i,y] E_out[x,i]
template <typename T>
__global__ void __launch_bounds__(TRI_THREADS)
triangular_bwd_kernel(const tgt_triangular_desc D, const T *__restrict__ proj, const float *__restrict__ mask,
                      const T *__restrict__ dva, T *__restrict__ dproj) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, b = blockIdx.y, N = D.N, H = D.H;
  tri_load_tiles<T>(D, proj, mask, b, h, sm);
  float *dIn = sm + 4 * TRI_TILE, *dOut = sm + 5 * TRI_TILE;
  for (int idx = threadIdx.x; idx < TRI_MAX_N * TRI_MAX_N; idx += TRI_THREADS) {
    const int i = idx / TRI_MAX_N, j = idx % TRI_MAX_N;
    float a = 0.f, c = 0.f;
    if (i < N && j < N) {
      const T *g = dva + (((int64_t)b * N + i) * N + j) * (2 * H);
      a = to_f(g[h]);
      c = to_f(g[H + h]);
    }
    dIn[i * TRI_LD + j] = a;
    dOut[i * TRI_LD + j] = c;
  }
  __syncthreads();
  const float *Vin = sm, *Vout = sm + TRI_TILE, *Ein = sm + 2 * TRI_TILE, *Eout = sm + 3 * TRI_TILE;
  for (int idx = threadIdx.x; idx < N * N; idx += TRI_THREADS) {
    const int x = idx / N, y = idx % N;
    float dX[4] = {0.f, 0.f, 0.f, 0.f};                 // V_in, V_out, E_in, E_out  (tile order)
    for (int r = 0; r < N; ++r) {
      dX[0] = fmaf(dIn[r * TRI_LD + x], Ein[r * TRI_LD + y], dX[0]);
      dX[1] = fmaf(dOut[r * TRI_LD + y], Eout[x * TRI_LD + r], dX[1]);
      dX[2] = fmaf(dIn[x * TRI_LD + r], Vin[r * TRI_LD + y], dX[2]);
      dX[3] = fmaf(dOut[y * TRI_LD + r], Vout[x * TRI_LD + r], dX[3]);
    }
    const int64_t row = ((int64_t)b * N + x) * N + y;
    const T *p = proj + row * D.ld;
    T *q = dproj + row * D.ld;
    const float m = mask[row];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int off = (c < 2 ? D.off_v : D.off_e) + (c & 1) * 2 * H + h;
      const float l = to_f(p[off + H]);
      const float s = sigmoid_acc(to_f(p[off]) + m);
      q[off] = from_f<T>(dX[c] * l * s * (1.f - s));
      q[off + H] = from_f<T>(dX[c] * s);
    }
  }
}

// y[r, c] = sigmoid(x[r, c]) * x[r, W + c]     (triplet.py:172-175: e_g, e_l = lin_O(Va).chunk(2); siglin)
template <typename T>
__global__ void __launch_bounds__(256)
siglin_fwd_kernel(const T *__restrict__ x, T *__restrict__ y, int64_t rows, int W) {
  constexpr int NV = Vec16<T>::n;
  const int wv = W / NV;
  const int64_t total = rows * wv;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / wv;
    const int c = (int)(idx % wv) * NV;
    float g[NV], l[NV], o[NV];
    load_vec<T, NV>(x + r * 2 * W + c, g);
    load_vec<T, NV>(x + r * 2 * W + W + c, l);
#pragma unroll
    for (int i = 0; i < NV; ++i) o[i] = sigmoid_acc(g[i]) * l[i];
    store_vec<T, NV>(y + r * W + c, o);
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
siglin_bwd_kernel(const T *__restrict__ x, const T *__restrict__ dy, T *__restrict__ dx, int64_t rows, int W) {
  constexpr int NV = Vec16<T>::n;
  const int wv = W / NV;
  const int64_t total = rows * wv;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / wv;
    const int c = (int)(idx % wv) * NV;
    float g[NV], l[NV], d[NV], dg[NV], dl[NV];
    load_vec<T, NV>(x + r * 2 * W + c, g);
    load_vec<T, NV>(x + r * 2 * W + W + c, l);
    load_vec<T, NV>(dy + r * W + c, d);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float s = sigmoid_acc(g[i]);
      dg[i] = d[i] * l[i] * s * (1.f - s);
      dl[i] = d[i] * s;
    }
    store_vec<T, NV>(dx + r * 2 * W + c, dg);
    store_vec<T, NV>(dx + r * 2 * W + W + c, dl);
  }
}

template <typename T>
static int tri_fwd(const tgt_triangular_desc &D, const void *proj, const float *mask, void *va, cudaStream_t st) {
  const int smem = 4 * TRI_TILE * (int)sizeof(float);
  TGT_CUDA_OK(cudaFuncSetAttribute(triangular_fwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  KernelTimerScope ts("triangular_fwd", st);
  triangular_fwd_kernel<T><<<dim3(D.H, D.B), TRI_THREADS, smem, st>>>(D, (const T *)proj, mask, (T *)va);
  return check_launch("triangular_fwd");
}
template <typename T>
static int tri_bwd(const tgt_triangular_desc &D, const void *proj, const float *mask, const void *dva, void *dproj,
                   cudaStream_t st) {
  const int smem = 6 * TRI_TILE * (int)sizeof(float);
  TGT_CUDA_OK(cudaFuncSetAttribute(triangular_bwd_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  KernelTimerScope ts("triangular_bwd", st);
  triangular_bwd_kernel<T><<<dim3(D.H, D.B), TRI_THREADS, smem, st>>>(D, (const T *)proj, mask, (const T *)dva,
                                                                     (T *)dproj);
  return check_launch("triangular_bwd");
}

static int tri_check(const tgt_triangular_desc *D) {
  if (!D) return fail("triangular: null descriptor");
  if (D->B <= 0 || D->N <= 0 || D->H <= 0) return fail("triangular: bad shape B=%d N=%d H=%d", D->B, D->N, D->H);
  if (D->N > TRI_MAX_N) return fail("triangular: N=%d > 64 unsupported", D->N);
  if (D->B > 65535) return fail("triangular: B > 65535 unsupported");
  if (D->off_v < 0 || D->off_e < 0 || D->ld < 8 * (int64_t)D->H) return fail("triangular: bad column layout");
  return 0;
}

static int grid_for(int64_t total) {
  int64_t g = (total + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 16 ? 148 * 16 : g));
}

}  // namespace tgt

using namespace tgt;

extern "C" int tgt_triangular_fwd(const tgt_triangular_desc *D, const void *proj, const float *mask, void *va,
                                  void *stream) {
  if (int e = tri_check(D)) return e;
  if (!proj || !mask || !va) return fail("triangular_fwd: null argument");
  TGT_DISPATCH_DTYPE(D->dtype, T, return tri_fwd<T>(*D, proj, mask, va, (cudaStream_t)stream));
  return 0;
}

extern "C" int tgt_triangular_bwd(const tgt_triangular_desc *D, const void *proj, const float *mask, const void *dva,
                                  void *dproj, void *stream) {
  if (int e = tri_check(D)) return e;
  if (!proj || !mask || !dva || !dproj) return fail("triangular_bwd: null argument");
  TGT_DISPATCH_DTYPE(D->dtype, T, return tri_bwd<T>(*D, proj, mask, dva, dproj, (cudaStream_t)stream));
  return 0;
}

extern "C" int tgt_siglin_fwd(const void *x, void *y, int64_t rows, int W, int dtype, void *stream) {
  if (rows <= 0) return 0;
  if (!x || !y) return fail("siglin_fwd: null argument");
  TGT_DISPATCH_DTYPE(dtype, T, {
    if (W <= 0 || W % Vec16<T>::n) return fail("siglin: width %d must be a multiple of %d", W, Vec16<T>::n);
    siglin_fwd_kernel<T><<<grid_for(rows * (W / Vec16<T>::n)), 256, 0, (cudaStream_t)stream>>>((const T *)x, (T *)y,
                                                                                               rows, W);
    return check_launch("siglin_fwd");
  });
  return 0;
}

extern "C" int tgt_siglin_bwd(const void *x, const void *dy, void *dx, int64_t rows, int W, int dtype, void *stream) {
  if (rows <= 0) return 0;
  if (!x || !dy || !dx) return fail("siglin_bwd: null argument");
  TGT_DISPATCH_DTYPE(dtype, T, {
    if (W <= 0 || W % Vec16<T>::n) return fail("siglin: width %d must be a multiple of %d", W, Vec16<T>::n);
    siglin_bwd_kernel<T><<<grid_for(rows * (W / Vec16<T>::n)), 256, 0, (cudaStream_t)stream>>>(
        (const T *)x, (const T *)dy, (T *)dx, rows, W);
    return check_launch("siglin_bwd");
  });
  return 0;
}
