// EGT node/edge attention core (reference lib/tgt/layers/layers.py:62-78 and 121-125).
//
// Reference-native channel layout c = dd*H + h is kept: with one thread per head, every load of
// Q/K/V/E/G/H_hat is contiguous across the warp (h is the innermost index everywhere), so the
// kernel streams the [R, 2H] bias/gate tensor and writes the [R, H] logits with full coalescing.
// The block is HBM-bound on those two edge-sized tensors; K/V of one graph (N x 3Wn) stay in L1/L2.
//
//   fwd   : block (l, b), thread h.   H_hat = scale*Q.K + E ; A = softmax(H_hat+M)*sigmoid(G+M)
//           V_att = (A V) * log1p(sum_m gate)
//   bwd A : block (l, b), thread h -> dE (=dH), dG, dQ            (row-wise quantities)
//   bwd B : block (m, b), thread h -> dK, dV                      (column-wise reductions over l)
#include "common.cuh"

namespace tgt {

template <typename T, int DMAX>
__global__ void egt_fwd_kernel(const tgt_egt_desc D, const T *__restrict__ qkv, const T *__restrict__ eg,
                               const float *__restrict__ mask, const float *__restrict__ src,
                               T *__restrict__ hhat, T *__restrict__ vatt, float *__restrict__ stats) {
  const int N = D.N, H = D.H, d = D.d;
  const int l = blockIdx.x, b = blockIdx.y;
  constexpr bool ACC = sizeof(T) == 4;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const T *qp = qkv + (int64_t)(b * N + l) * D.ld_qkv;
    float q[DMAX], o[DMAX];
#pragma unroll
    for (int dd = 0; dd < DMAX; ++dd) {
      q[dd] = dd < d ? to_f(qp[dd * H + h]) * D.scale : 0.f;
      o[dd] = 0.f;
    }
    float mx = -INFINITY, lsum = 0.f, deg = 0.f;
    for (int m = 0; m < N; ++m) {
      const T *kp = qkv + (int64_t)(b * N + m) * D.ld_qkv + (int64_t)H * d;
      const int64_t row = (int64_t)(b * N + l) * N + m;
      float s = 0.f;
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd)
        if (dd < d) s += q[dd] * to_f(kp[dd * H + h]);
      s += to_f(eg[row * D.ld_eg + h]);
      hhat[row * H + h] = from_f<T>(s);
      if (D.attend) {
        float mk = mask[row];
        if (src) mk += src[b * N + m];
        const float g = ACC ? sigmoid_acc(to_f(eg[row * D.ld_eg + H + h]) + mk)
                            : sigmoidf_(to_f(eg[row * D.ld_eg + H + h]) + mk);
        s += mk;
        // keys masked to -inf (padding + source dropout, layers.py:55-59) contribute exactly 0, also as the
        // first key of the online softmax; a row whose keys are ALL -inf ends as 0 * inf = NaN like the reference
        const float mn = fmaxf(mx, s);
        const float corr = mx == -INFINITY ? 0.f : expf(mx - mn);
        const float p = s == -INFINITY ? 0.f : expf(s - mn);
        lsum = lsum * corr + p;
        const float pg = p * g;
        const T *vp = kp + (int64_t)H * d;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) o[dd] = o[dd] * corr + pg * to_f(vp[dd * H + h]);
        deg += g;
        mx = mn;
      }
    }
    if (D.attend) {
      const float inv = 1.f / lsum;
      const float sc = D.scale_degree ? log1pf(deg) : 1.f;
      T *op = vatt + (int64_t)(b * N + l) * H * d;
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd)
        if (dd < d) op[dd * H + h] = from_f<T>(o[dd] * inv * sc);
      float *st = stats + ((int64_t)(b * N + l) * H + h) * 3;
      st[0] = mx; st[1] = inv; st[2] = deg;
    }
  }
}

template <typename T, int DMAX>
__global__ void egt_bwd_rows_kernel(const tgt_egt_desc D, const T *__restrict__ qkv, const T *__restrict__ eg,
                                    const float *__restrict__ mask, const float *__restrict__ src,
                                    const float *__restrict__ stats, const T *__restrict__ dhhat,
                                    const T *__restrict__ dvatt, T *__restrict__ dqkv, T *__restrict__ deg_out) {
  const int N = D.N, H = D.H, d = D.d;
  const int l = blockIdx.x, b = blockIdx.y;
  constexpr bool ACC = sizeof(T) == 4;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const T *qp = qkv + (int64_t)(b * N + l) * D.ld_qkv;
    float q[DMAX], dq[DMAX];
#pragma unroll
    for (int dd = 0; dd < DMAX; ++dd) { q[dd] = dd < d ? to_f(qp[dd * H + h]) * D.scale : 0.f; dq[dd] = 0.f; }
    if (!D.attend) {
      // EdgeUpdate: dE = dH ; dQ = scale * dH K
      for (int m = 0; m < N; ++m) {
        const int64_t row = (int64_t)(b * N + l) * N + m;
        const float dH = dhhat ? to_f(dhhat[row * H + h]) : 0.f;
        deg_out[row * D.ld_eg + h] = from_f<T>(dH);
        const T *kp = qkv + (int64_t)(b * N + m) * D.ld_qkv + (int64_t)H * d;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) dq[dd] += dH * to_f(kp[dd * H + h]);
      }
    } else {
      const float *st = stats + ((int64_t)(b * N + l) * H + h) * 3;
      const float mx = st[0], inv = st[1], deg = st[2];
      const float sc = D.scale_degree ? log1pf(deg) : 1.f;
      const T *gp = dvatt + (int64_t)(b * N + l) * H * d;
      float go[DMAX], u[DMAX];
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd) { go[dd] = dd < d ? to_f(gp[dd * H + h]) : 0.f; u[dd] = 0.f; }
      // pass 1: U = A V (unscaled attention output)
      for (int m = 0; m < N; ++m) {
        const T *kp = qkv + (int64_t)(b * N + m) * D.ld_qkv + (int64_t)H * d;
        const T *vp = kp + (int64_t)H * d;
        const int64_t row = (int64_t)(b * N + l) * N + m;
        float s = 0.f;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) s += q[dd] * to_f(kp[dd * H + h]);
        float mk = mask[row];
        if (src) mk += src[b * N + m];
        s += to_f(eg[row * D.ld_eg + h]) + mk;
        const float g = ACC ? sigmoid_acc(to_f(eg[row * D.ld_eg + H + h]) + mk)
                            : sigmoidf_(to_f(eg[row * D.ld_eg + H + h]) + mk);
        const float a = expf(s - mx) * inv * g;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) u[dd] += a * to_f(vp[dd * H + h]);
      }
      float dsc = 0.f, delta = 0.f;       // dsc = dV_att . U ; delta = dU . U = sc * dsc
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd) dsc += go[dd] * u[dd];
      delta = dsc * sc;
      const float ddeg = D.scale_degree ? dsc / (1.f + deg) : 0.f;
      // pass 2: gradients
      for (int m = 0; m < N; ++m) {
        const T *kp = qkv + (int64_t)(b * N + m) * D.ld_qkv + (int64_t)H * d;
        const T *vp = kp + (int64_t)H * d;
        const int64_t row = (int64_t)(b * N + l) * N + m;
        float s = 0.f, dA = 0.f;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) { s += q[dd] * to_f(kp[dd * H + h]); dA += go[dd] * to_f(vp[dd * H + h]); }
        dA *= sc;
        float mk = mask[row];
        if (src) mk += src[b * N + m];
        s += to_f(eg[row * D.ld_eg + h]) + mk;
        const float g = ACC ? sigmoid_acc(to_f(eg[row * D.ld_eg + H + h]) + mk)
                            : sigmoidf_(to_f(eg[row * D.ld_eg + H + h]) + mk);
        const float p = expf(s - mx) * inv;
        float dH = p * (dA * g - delta);
        if (dhhat) dH += to_f(dhhat[row * H + h]);
        const float dg = (dA * p + ddeg) * g * (1.f - g);
        deg_out[row * D.ld_eg + h] = from_f<T>(dH);
        deg_out[row * D.ld_eg + H + h] = from_f<T>(dg);
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) dq[dd] += dH * to_f(kp[dd * H + h]);
      }
    }
    T *op = dqkv + (int64_t)(b * N + l) * D.ld_qkv;
#pragma unroll
    for (int dd = 0; dd < DMAX; ++dd)
      if (dd < d) op[dd * H + h] = from_f<T>(dq[dd] * D.scale);
  }
}

// column pass: block (m, b), thread h.  Reads dH from deg_out (written by the row pass).
template <typename T, int DMAX>
__global__ void egt_bwd_cols_kernel(const tgt_egt_desc D, const T *__restrict__ qkv, const T *__restrict__ eg,
                                    const float *__restrict__ mask, const float *__restrict__ src,
                                    const float *__restrict__ stats, const T *__restrict__ dvatt,
                                    const T *__restrict__ deg_out, T *__restrict__ dqkv) {
  const int N = D.N, H = D.H, d = D.d;
  const int m = blockIdx.x, b = blockIdx.y;
  constexpr bool ACC = sizeof(T) == 4;
  for (int h = threadIdx.x; h < H; h += blockDim.x) {
    const T *kp = qkv + (int64_t)(b * N + m) * D.ld_qkv + (int64_t)H * d;
    float kk[DMAX], dk[DMAX], dv[DMAX];
#pragma unroll
    for (int dd = 0; dd < DMAX; ++dd) { kk[dd] = dd < d ? to_f(kp[dd * H + h]) : 0.f; dk[dd] = 0.f; dv[dd] = 0.f; }
    const float smk = src ? src[b * N + m] : 0.f;
    for (int l = 0; l < N; ++l) {
      const T *qp = qkv + (int64_t)(b * N + l) * D.ld_qkv;
      const int64_t row = (int64_t)(b * N + l) * N + m;
      const float dH = to_f(deg_out[row * D.ld_eg + h]);
      float s = 0.f;
#pragma unroll
      for (int dd = 0; dd < DMAX; ++dd)
        if (dd < d) { const float qv = to_f(qp[dd * H + h]) * D.scale; s += qv * kk[dd]; dk[dd] += dH * qv; }
      if (D.attend) {
        const float *st = stats + ((int64_t)(b * N + l) * H + h) * 3;
        const float mk = mask[row] + smk;
        s += to_f(eg[row * D.ld_eg + h]) + mk;
        const float g = ACC ? sigmoid_acc(to_f(eg[row * D.ld_eg + H + h]) + mk)
                            : sigmoidf_(to_f(eg[row * D.ld_eg + H + h]) + mk);
        const float sc = D.scale_degree ? log1pf(st[2]) : 1.f;
        const float a = expf(s - st[0]) * st[1] * g * sc;
        const T *gp = dvatt + (int64_t)(b * N + l) * H * d;
#pragma unroll
        for (int dd = 0; dd < DMAX; ++dd)
          if (dd < d) dv[dd] += a * to_f(gp[dd * H + h]);
      }
    }
    T *okp = dqkv + (int64_t)(b * N + m) * D.ld_qkv + (int64_t)H * d;
#pragma unroll
    for (int dd = 0; dd < DMAX; ++dd)
      if (dd < d) {
        okp[dd * H + h] = from_f<T>(dk[dd]);
        if (D.attend) okp[(int64_t)H * d + dd * H + h] = from_f<T>(dv[dd]);
      }
  }
}

static int egt_threads(int H) {
  int t = ((H + 31) / 32) * 32;
  return t > 256 ? 256 : t;
}

template <typename T>
static int egt_fwd_t(const tgt_egt_desc &D, const void *qkv, const void *eg, const float *mask, const float *src,
                     void *hhat, void *vatt, float *stats, cudaStream_t st) {
  dim3 grid(D.N, D.B);
  const int th = egt_threads(D.H);
#define L(DM) egt_fwd_kernel<T, DM><<<grid, th, 0, st>>>(D, (const T *)qkv, (const T *)eg, mask, src, (T *)hhat, (T *)vatt, stats)
  if (D.d <= 16) L(16); else if (D.d <= 32) L(32); else L(64);
#undef L
  return check_launch("egt_fwd_kernel");
}

template <typename T>
static int egt_bwd_t(const tgt_egt_desc &D, const void *qkv, const void *eg, const float *mask, const float *src,
                     const float *stats, const void *dhhat, const void *dvatt, void *dqkv, void *deg,
                     cudaStream_t st) {
  dim3 grid(D.N, D.B);
  const int th = egt_threads(D.H);
#define L(DM)                                                                                                  \
  do {                                                                                                         \
    egt_bwd_rows_kernel<T, DM><<<grid, th, 0, st>>>(D, (const T *)qkv, (const T *)eg, mask, src, stats,        \
                                                    (const T *)dhhat, (const T *)dvatt, (T *)dqkv, (T *)deg);  \
    if (int e = check_launch("egt_bwd_rows_kernel")) return e;                                                 \
    egt_bwd_cols_kernel<T, DM><<<grid, th, 0, st>>>(D, (const T *)qkv, (const T *)eg, mask, src, stats,        \
                                                    (const T *)dvatt, (const T *)deg, (T *)dqkv);              \
  } while (0)
  if (D.d <= 16) L(16); else if (D.d <= 32) L(32); else L(64);
#undef L
  return check_launch("egt_bwd_cols_kernel");
}

// egt_fast.cu
bool egt_fast_supported(const tgt_egt_desc &);
size_t egt_fast_workspace(const tgt_egt_desc &);
int egt_fwd_fast_launch(const tgt_egt_desc &, const void *, const void *, const float *, const float *, void *, void *,
                        float *, cudaStream_t);
int egt_bwd_fast_launch(const tgt_egt_desc &, const void *, const void *, const float *, const float *, const float *,
                        const void *, const void *, const void *, void *, void *, void *, size_t, cudaStream_t);
}  // namespace tgt

using namespace tgt;

static int egt_check(const tgt_egt_desc *D) {
  if (!D) return fail("egt: null descriptor");
  if (D->B <= 0 || D->N <= 0 || D->H <= 0 || D->d <= 0) return fail("egt: bad shape B=%d N=%d H=%d d=%d", D->B, D->N, D->H, D->d);
  if (D->d > 64) return fail("egt: head dim %d > 64 unsupported", D->d);
  if (D->B > 65535) return fail("egt: B=%d > 65535 unsupported", D->B);
  return 0;
}

extern "C" size_t tgt_egt_attn_workspace_bytes(const tgt_egt_desc *D) {
  if (!D || egt_check(D)) return 0;
  return (g_policy.load() != 1 && egt_fast_supported(*D)) ? egt_fast_workspace(*D) : 0;
}

extern "C" int tgt_egt_attn_fwd(const tgt_egt_desc *D, const void *qkv, const void *eg, const float *mask,
                                const float *src_mask, void *hhat, void *vatt, float *stats, void *stream) {
  if (int e = egt_check(D)) return e;
  if (g_policy.load() != 1 && egt_fast_supported(*D))
    return egt_fwd_fast_launch(*D, qkv, eg, mask, src_mask, hhat, vatt, stats, (cudaStream_t)stream);
  TGT_DISPATCH_DTYPE(D->dtype, T, return egt_fwd_t<T>(*D, qkv, eg, mask, src_mask, hhat, vatt, stats, (cudaStream_t)stream));
  return 0;
}

extern "C" int tgt_egt_attn_bwd(const tgt_egt_desc *D, const void *qkv, const void *eg, const float *mask,
                                const float *src_mask, const float *stats, const void *vatt, const void *dhhat,
                                const void *dvatt, void *dqkv, void *deg, void *ws, size_t ws_bytes, void *stream) {
  if (int e = egt_check(D)) return e;
  if (g_policy.load() != 1 && egt_fast_supported(*D))
    return egt_bwd_fast_launch(*D, qkv, eg, mask, src_mask, stats, vatt, dhhat, dvatt, dqkv, deg, ws, ws_bytes,
                               (cudaStream_t)stream);
  TGT_DISPATCH_DTYPE(D->dtype, T, return egt_bwd_t<T>(*D, qkv, eg, mask, src_mask, stats, dhhat, dvatt, dqkv, deg, (cudaStream_t)stream));
  return 0;
}
