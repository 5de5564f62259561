// EGT node/edge attention core, 16-bit fast path (reference lib/tgt/layers/layers.py:62-78 and 121-125).
//
// The block is HBM-bound on the edge-sized tensors (E|G in, H_hat out: ~0.4 GB per layer at B=256, N=64) and
// has only 4*B*N^2*Wn FLOPs, so it runs on CUDA cores; what matters is that every edge byte moves exactly once,
// coalesced, and that K/V (resp. Q/dV_att) rows are staged once per CTA in shared memory instead of being
// re-fetched per query row.
//
// Channel layout is the reference's c = dd*H + h.  One thread owns a PAIR of adjacent heads (2hp, 2hp+1), so all
// loads of Q/K/V/E/G/H_hat are 4-byte and a warp touches 128 contiguous bytes.
//   fwd      : CTA = (tile of 8 query rows l, b), thread = (l, hp); K|V rows stream through smem in chunks of 16 keys
//   bwd rows : same mapping; one pass over the keys: dH (-> dE), dG, dQ and the scaled attention weights a*sc
//              (scratch, [R,H]) that the column pass needs
//   bwd cols : CTA = (tile of 8 keys m, b), thread = (m, hp); Q|dV_att rows stream through smem; dK, dV
#include "common.cuh"

namespace tgt {

constexpr int EF_ROWS = 8;       // rows (l or m) per CTA
constexpr int EF_BROWS = 4;      // rows per CTA of the backward row pass (register-heavy: 3 CTAs x 4 warps per SM)
constexpr int EF_CHUNK = 8;      // staged rows per shared-memory chunk
constexpr int EF_DMAX = 16;

__device__ __forceinline__ void ef_cp16(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void ef_commit_wait() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::);
}

template <typename T> struct Pair;
template <> struct Pair<__nv_bfloat16> {
  static __device__ __forceinline__ float2 unpack(uint32_t v) {
    return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&v));
  }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
  }
};
template <> struct Pair<__half> {
  static __device__ __forceinline__ float2 unpack(uint32_t v) { return __half22float2(*reinterpret_cast<__half2 *>(&v)); }
  static __device__ __forceinline__ uint32_t pack(float a, float b) {
    __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&v);
  }
};

template <typename T>
__device__ __forceinline__ uint32_t ldg32(const T *p) { return *reinterpret_cast<const uint32_t *>(p); }
template <typename T>
__device__ __forceinline__ void stg32(T *p, uint32_t v) { *reinterpret_cast<uint32_t *>(p) = v; }

// stage rows [r0, r0+EF_CHUNK) x `width` elements (16-bit) from src (row stride ld, column offset col0) into smem
template <typename T>
__device__ __forceinline__ void stage_rows(unsigned char *sm, const T *src, int64_t row_base, int r0, int nrows,
                                           int64_t ld, int col0, int width, int sm_col0, int sm_ld) {
  const int vec_per_row = width / 8;
  for (int idx = threadIdx.x; idx < EF_CHUNK * vec_per_row; idx += blockDim.x) {
    const int r = idx / vec_per_row, v = idx - r * vec_per_row;
    const bool ok = (r0 + r) < nrows;
    const T *g = src + (row_base + (ok ? r0 + r : 0)) * ld + col0 + v * 8;
    const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)((r * sm_ld + sm_col0 + v * 8) * 2);
    ef_cp16(dst, g, ok);
  }
}

// ------------------------------------------------------------------------------------------------ forward
constexpr int EF_FCHUNK = 4;     // keys per shared-memory stage of the forward (double buffered)

template <typename T, int HC, int DC, int ATT>
__global__ void __launch_bounds__(EF_ROWS * 32, 2)
egt_fwd_fast(const tgt_egt_desc D, const T *__restrict__ qkv, const T *__restrict__ eg, const float *__restrict__ mask,
             const float *__restrict__ src, T *__restrict__ hhat, T *__restrict__ vatt, float *__restrict__ stats) {
  extern __shared__ __align__(16) unsigned char sm[];           // 2 x [EF_FCHUNK][2*Wn] K|V   (or K only)
  const int N = D.N, H = HC ? HC : D.H, d = DC ? DC : D.d, Wn = H * d;   // HC/DC: compile-time shape (0 = runtime)
  constexpr int DL = DC ? DC : EF_DMAX;
  const int b = blockIdx.y, l = blockIdx.x * EF_ROWS + (threadIdx.x >> 5), hp = threadIdx.x & 31;
  const bool active = l < N && 2 * hp < H;
  const int la = l < N ? l : N - 1;
  const bool attend = ATT < 0 ? D.attend != 0 : ATT != 0;     // ATT: compile-time attend flag (-1 = runtime)
  const int kvw = attend ? 2 * Wn : Wn;
  const int stage_bytes = EF_FCHUNK * kvw * 2;

  // one thread = one query row x one PAIR of heads: the pair rides in packed f32x2 registers (fma.rn.f32x2)
  float2 qv[DL], ov[DL];
#pragma unroll
  for (int dd = 0; dd < DL; ++dd) {
    qv[dd] = ov[dd] = make_float2(0.f, 0.f);
    if (dd < d && active) {
      const float2 v = Pair<T>::unpack(ldg32(qkv + (int64_t)(b * N + la) * D.ld_qkv + dd * H + 2 * hp));
      qv[dd] = make_float2(v.x * D.scale, v.y * D.scale);
    }
  }
  float mx0 = -INFINITY, mx1 = -INFINITY, ls0 = 0.f, ls1 = 0.f, dg0 = 0.f, dg1 = 0.f;
  const int64_t erow0 = (int64_t)(b * N + la) * N;

  // stage `mc` : K|V rows of keys mc .. mc+EF_FCHUNK-1 into buffer `buf` (cp.async) + this thread's bias / gate / mask
  auto stage_kv = [&](int mc, int buf) {
    const int vec_per_row = kvw / 8;
    for (int idx = threadIdx.x; idx < EF_FCHUNK * vec_per_row; idx += blockDim.x) {
      const int r = idx / vec_per_row, v = idx - r * vec_per_row;
      const bool ok = (mc + r) < N;
      const T *g = qkv + ((int64_t)b * N + (ok ? mc + r : 0)) * D.ld_qkv + Wn + v * 8;
      ef_cp16((uint32_t)__cvta_generic_to_shared(sm) + (uint32_t)(buf * stage_bytes + (r * kvw + v * 8) * 2), g, ok);
    }
    asm volatile("cp.async.commit_group;\n" ::);
  };
  uint32_t ep[EF_FCHUNK], gp[EF_FCHUNK], epn[EF_FCHUNK], gpn[EF_FCHUNK];
  float mk[EF_FCHUNK], mkn[EF_FCHUNK];
  auto fetch_eg = [&](int mc, uint32_t(&e_)[EF_FCHUNK], uint32_t(&g_)[EF_FCHUNK], float(&m_)[EF_FCHUNK]) {
#pragma unroll
    for (int mm = 0; mm < EF_FCHUNK; ++mm) {
      const int m = mc + mm;
      e_[mm] = g_[mm] = 0u;
      m_[mm] = 0.f;
      if (active && m < N) {
        e_[mm] = ldg32(eg + (erow0 + m) * D.ld_eg + 2 * hp);
        if (attend) {
          g_[mm] = ldg32(eg + (erow0 + m) * D.ld_eg + H + 2 * hp);
          m_[mm] = mask[erow0 + m] + (src ? src[b * N + m] : 0.f);
        }
      }
    }
  };
  stage_kv(0, 0);
  fetch_eg(0, ep, gp, mk);

  for (int mc = 0, it = 0; mc < N; mc += EF_FCHUNK, ++it) {
    asm volatile("cp.async.wait_group 0;\n" ::);
    __syncthreads();                 // stage `it` landed; everyone is done with stage it-1 (the buffer refilled next)
    if (mc + EF_FCHUNK < N) {
      stage_kv(mc + EF_FCHUNK, (it + 1) & 1);
      fetch_eg(mc + EF_FCHUNK, epn, gpn, mkn);
    }
    const T *smT = reinterpret_cast<const T *>(sm + (it & 1) * stage_bytes);
    if (active) {
#pragma unroll
      for (int mm = 0; mm < EF_FCHUNK; ++mm) {
        const int m = mc + mm;
        if (m >= N) continue;
        const T *kp = smT + mm * kvw + 2 * hp;
        float2 sv = make_float2(0.f, 0.f);
#pragma unroll
        for (int dd = 0; dd < DL; ++dd)
          if (dd < d) sv = __ffma2_rn(qv[dd], Pair<T>::unpack(*reinterpret_cast<const uint32_t *>(kp + dd * H)), sv);
        const float2 ev = Pair<T>::unpack(ep[mm]);
        float s0 = sv.x + ev.x, s1 = sv.y + ev.y;
        stg32(hhat + (erow0 + m) * H + 2 * hp, Pair<T>::pack(s0, s1));
        if (attend) {
          const float2 gv = Pair<T>::unpack(gp[mm]);
          const float g0 = sigmoidf_(gv.x + mk[mm]), g1 = sigmoidf_(gv.y + mk[mm]);
          s0 += mk[mm];
          s1 += mk[mm];
          // online softmax; -inf keys (padding + source dropout) contribute exactly 0, an all -inf row ends as NaN
          const float n0 = fmaxf(mx0, s0), n1 = fmaxf(mx1, s1);
          const float c0 = mx0 == -INFINITY ? 0.f : __expf(mx0 - n0), c1 = mx1 == -INFINITY ? 0.f : __expf(mx1 - n1);
          const float p0 = s0 == -INFINITY ? 0.f : __expf(s0 - n0), p1 = s1 == -INFINITY ? 0.f : __expf(s1 - n1);
          ls0 = ls0 * c0 + p0;
          ls1 = ls1 * c1 + p1;
          const float2 cc = make_float2(c0, c1), aa = make_float2(p0 * g0, p1 * g1);
          const T *vp = kp + Wn;
#pragma unroll
          for (int dd = 0; dd < DL; ++dd)
            if (dd < d)
              ov[dd] = __ffma2_rn(ov[dd], cc, __fmul2_rn(aa, Pair<T>::unpack(*reinterpret_cast<const uint32_t *>(vp + dd * H))));
          dg0 += g0;
          dg1 += g1;
          mx0 = n0;
          mx1 = n1;
        }
      }
    }
#pragma unroll
    for (int mm = 0; mm < EF_FCHUNK; ++mm) {
      ep[mm] = epn[mm];
      gp[mm] = gpn[mm];
      mk[mm] = mkn[mm];
    }
  }
  if (active && attend) {
    const float i0 = 1.f / ls0, i1 = 1.f / ls1;
    const float sc0 = D.scale_degree ? log1pf(dg0) : 1.f, sc1 = D.scale_degree ? log1pf(dg1) : 1.f;
    T *op = vatt + (int64_t)(b * N + l) * Wn + 2 * hp;
#pragma unroll
    for (int dd = 0; dd < DL; ++dd)
      if (dd < d) stg32(op + dd * H, Pair<T>::pack(ov[dd].x * i0 * sc0, ov[dd].y * i1 * sc1));
    float *st = stats + ((int64_t)(b * N + l) * H + 2 * hp) * 3;
    st[0] = mx0; st[1] = i0; st[2] = dg0;
    st[3] = mx1; st[4] = i1; st[5] = dg1;
  }
}

// ------------------------------------------------------------------------------------------------ backward, row pass
// vatt is the forward output (U * sc): delta = sum_m dP P = dV_att . U * sc = dV_att . V_att
template <typename T, int HC, int DC, int ATT>
__global__ void __launch_bounds__(EF_BROWS * 32, 3)
egt_bwd_rows_fast(const tgt_egt_desc D, const T *__restrict__ qkv, const T *__restrict__ eg,
                  const float *__restrict__ mask, const float *__restrict__ src, const float *__restrict__ stats,
                  const T *__restrict__ vatt, const T *__restrict__ dhhat, const T *__restrict__ dvatt,
                  T *__restrict__ dqkv, T *__restrict__ deg_out, T *__restrict__ aw) {
  extern __shared__ __align__(16) unsigned char sm[];
  const int N = D.N, H = HC ? HC : D.H, d = DC ? DC : D.d, Wn = H * d;   // HC/DC: compile-time shape (0 = runtime)
  constexpr int DL = DC ? DC : EF_DMAX;
  const int b = blockIdx.y, l = blockIdx.x * EF_BROWS + (threadIdx.x >> 5), hp = threadIdx.x & 31;
  const bool active = l < N && 2 * hp < H;
  const int la = l < N ? l : N - 1;
  const bool attend = ATT < 0 ? D.attend != 0 : ATT != 0;     // ATT: compile-time attend flag (-1 = runtime)
  const int kvw = attend ? 2 * Wn : Wn;
  const T *smT = reinterpret_cast<const T *>(sm);

  // one thread = one query row x one PAIR of heads, carried in packed f32x2 registers (fma.rn.f32x2)
  float2 qv[DL], gov[DL], dqv[DL];
  float2 dsc = make_float2(0.f, 0.f);
#pragma unroll
  for (int dd = 0; dd < DL; ++dd) {
    qv[dd] = gov[dd] = dqv[dd] = make_float2(0.f, 0.f);
    if (dd < d && active) {
      const float2 v = Pair<T>::unpack(ldg32(qkv + (int64_t)(b * N + la) * D.ld_qkv + dd * H + 2 * hp));
      qv[dd] = make_float2(v.x * D.scale, v.y * D.scale);
      if (attend) {
        gov[dd] = Pair<T>::unpack(ldg32(dvatt + (int64_t)(b * N + la) * Wn + dd * H + 2 * hp));
        dsc = __ffma2_rn(gov[dd], Pair<T>::unpack(ldg32(vatt + (int64_t)(b * N + la) * Wn + dd * H + 2 * hp)), dsc);
      }
    }
  }
  const float dsc0 = dsc.x, dsc1 = dsc.y;
  float mx0 = 0.f, mx1 = 0.f, i0 = 0.f, i1 = 0.f, sc0 = 1.f, sc1 = 1.f, dd0 = 0.f, dd1 = 0.f;
  // dsc* currently hold dV_att . V_att = delta (dU . U);  ddeg = (dV_att . U) / (1 + deg) = delta / (sc (1 + deg))
  float delta0 = dsc0, delta1 = dsc1;
  if (active && attend) {
    const float *st = stats + ((int64_t)(b * N + l) * H + 2 * hp) * 3;
    mx0 = st[0]; i0 = st[1];
    mx1 = st[3]; i1 = st[4];
    if (D.scale_degree) {
      sc0 = log1pf(st[2]);
      sc1 = log1pf(st[5]);
      dd0 = sc0 > 0.f ? delta0 / (sc0 * (1.f + st[2])) : 0.f;
      dd1 = sc1 > 0.f ? delta1 / (sc1 * (1.f + st[5])) : 0.f;
    }
  }
  const int64_t erow0 = (int64_t)(b * N + la) * N;

  for (int mc = 0; mc < N; mc += EF_CHUNK) {
    __syncthreads();
    stage_rows<T>(sm, qkv, (int64_t)b * N, mc, N, D.ld_qkv, Wn, kvw, 0, kvw);
    uint32_t ep[EF_CHUNK], gp[EF_CHUNK], hp_[EF_CHUNK];
    float mk[EF_CHUNK];
#pragma unroll
    for (int mm = 0; mm < EF_CHUNK; ++mm) {
      const int m = mc + mm;
      ep[mm] = gp[mm] = hp_[mm] = 0u;
      mk[mm] = 0.f;
      if (active && m < N) {
        if (dhhat) hp_[mm] = ldg32(dhhat + (erow0 + m) * H + 2 * hp);
        if (attend) {
          ep[mm] = ldg32(eg + (erow0 + m) * D.ld_eg + 2 * hp);
          gp[mm] = ldg32(eg + (erow0 + m) * D.ld_eg + H + 2 * hp);
          mk[mm] = mask[erow0 + m] + (src ? src[b * N + m] : 0.f);
        }
      }
    }
    ef_commit_wait();
    __syncthreads();
    if (active) {
#pragma unroll
      for (int mm = 0; mm < EF_CHUNK; ++mm) {
        const int m = mc + mm;
        if (m >= N) continue;
        const T *kp = smT + mm * kvw + 2 * hp;
        const float2 dhv = Pair<T>::unpack(hp_[mm]);
        float dH0 = dhv.x, dH1 = dhv.y;
        if (attend) {
          const T *vp = kp + Wn;
          float2 sv = make_float2(0.f, 0.f), dAv = make_float2(0.f, 0.f);
#pragma unroll
          for (int dd = 0; dd < DL; ++dd)
            if (dd < d) {
              sv = __ffma2_rn(qv[dd], Pair<T>::unpack(*reinterpret_cast<const uint32_t *>(kp + dd * H)), sv);
              dAv = __ffma2_rn(gov[dd], Pair<T>::unpack(*reinterpret_cast<const uint32_t *>(vp + dd * H)), dAv);
            }
          float s0 = sv.x, s1 = sv.y;
          const float dA0 = dAv.x * sc0, dA1 = dAv.y * sc1;
          const float2 ev = Pair<T>::unpack(ep[mm]);
          const float2 gv = Pair<T>::unpack(gp[mm]);
          s0 += ev.x + mk[mm];
          s1 += ev.y + mk[mm];
          const float g0 = sigmoidf_(gv.x + mk[mm]), g1 = sigmoidf_(gv.y + mk[mm]);
          const float p0 = __expf(s0 - mx0) * i0, p1 = __expf(s1 - mx1) * i1;
          dH0 += p0 * (dA0 * g0 - delta0);
          dH1 += p1 * (dA1 * g1 - delta1);
          const float dG0 = (dA0 * p0 + dd0) * g0 * (1.f - g0), dG1 = (dA1 * p1 + dd1) * g1 * (1.f - g1);
          stg32(deg_out + (erow0 + m) * D.ld_eg + H + 2 * hp, Pair<T>::pack(dG0, dG1));
          stg32(aw + (erow0 + m) * H + 2 * hp, Pair<T>::pack(p0 * g0 * sc0, p1 * g1 * sc1));
        }
        stg32(deg_out + (erow0 + m) * D.ld_eg + 2 * hp, Pair<T>::pack(dH0, dH1));
        const float2 dHv = make_float2(dH0, dH1);
#pragma unroll
        for (int dd = 0; dd < DL; ++dd)
          if (dd < d) dqv[dd] = __ffma2_rn(dHv, Pair<T>::unpack(*reinterpret_cast<const uint32_t *>(kp + dd * H)), dqv[dd]);
      }
    }
  }
  if (active) {
    T *op = dqkv + (int64_t)(b * N + l) * D.ld_qkv + 2 * hp;
#pragma unroll
    for (int dd = 0; dd < DL; ++dd)
      if (dd < d) stg32(op + dd * H, Pair<T>::pack(dqv[dd].x * D.scale, dqv[dd].y * D.scale));
  }
}

// ------------------------------------------------------------------------------------------------ backward, column pass
// reads dH from deg_out (written by the row pass) and a*sc from aw;  dK = scale * sum_l dH q ; dV = sum_l (a sc) dV_att
template <typename T, int HC, int DC, int ATT>
__global__ void __launch_bounds__(EF_ROWS * 32, 2)
egt_bwd_cols_fast(const tgt_egt_desc D, const T *__restrict__ qkv, const T *__restrict__ dvatt,
                  const T *__restrict__ deg_out, const T *__restrict__ aw, T *__restrict__ dqkv) {
  extern __shared__ __align__(16) unsigned char sm[];           // [EF_CHUNK][2*Wn]  Q | dV_att
  const int N = D.N, H = HC ? HC : D.H, d = DC ? DC : D.d, Wn = H * d;   // HC/DC: compile-time shape (0 = runtime)
  constexpr int DL = DC ? DC : EF_DMAX;
  const int b = blockIdx.y, m = blockIdx.x * EF_ROWS + (threadIdx.x >> 5), hp = threadIdx.x & 31;
  const bool active = m < N && 2 * hp < H;
  const int ma = m < N ? m : N - 1;
  const bool attend = ATT < 0 ? D.attend != 0 : ATT != 0;     // ATT: compile-time attend flag (-1 = runtime)
  const int qw = attend ? 2 * Wn : Wn;
  const T *smT = reinterpret_cast<const T *>(sm);

  float2 dkv[DL], dvv[DL];
#pragma unroll
  for (int dd = 0; dd < DL; ++dd) dkv[dd] = dvv[dd] = make_float2(0.f, 0.f);

  for (int lc = 0; lc < N; lc += EF_CHUNK) {
    __syncthreads();
    stage_rows<T>(sm, qkv, (int64_t)b * N, lc, N, D.ld_qkv, 0, Wn, 0, qw);
    if (attend) stage_rows<T>(sm, dvatt, (int64_t)b * N, lc, N, Wn, 0, Wn, Wn, qw);
    uint32_t hp_[EF_CHUNK], ap[EF_CHUNK];
#pragma unroll
    for (int ll = 0; ll < EF_CHUNK; ++ll) {
      const int l = lc + ll;
      hp_[ll] = ap[ll] = 0u;
      if (active && l < N) {
        const int64_t row = (int64_t)(b * N + l) * N + ma;
        hp_[ll] = ldg32(deg_out + row * D.ld_eg + 2 * hp);
        if (attend) ap[ll] = ldg32(aw + row * H + 2 * hp);
      }
    }
    ef_commit_wait();
    __syncthreads();
    if (active) {
#pragma unroll
      for (int ll = 0; ll < EF_CHUNK; ++ll) {
        if (lc + ll >= N) continue;
        const T *qp = smT + ll * qw + 2 * hp;
        const float2 dh = Pair<T>::unpack(hp_[ll]);
        const float2 av = Pair<T>::unpack(ap[ll]);
#pragma unroll
        for (int dd = 0; dd < DL; ++dd)
          if (dd < d) {
            dkv[dd] = __ffma2_rn(dh, Pair<T>::unpack(*reinterpret_cast<const uint32_t *>(qp + dd * H)), dkv[dd]);
            if (attend)
              dvv[dd] = __ffma2_rn(av, Pair<T>::unpack(*reinterpret_cast<const uint32_t *>(qp + Wn + dd * H)), dvv[dd]);
          }
      }
    }
  }
  if (active) {
    T *okp = dqkv + (int64_t)(b * N + m) * D.ld_qkv + Wn + 2 * hp;
#pragma unroll
    for (int dd = 0; dd < DL; ++dd)
      if (dd < d) {
        stg32(okp + dd * H, Pair<T>::pack(dkv[dd].x * D.scale, dkv[dd].y * D.scale));
        if (attend) stg32(okp + Wn + dd * H, Pair<T>::pack(dvv[dd].x, dvv[dd].y));
      }
  }
}

// ------------------------------------------------------------------------------------------------ host
bool egt_fast_supported(const tgt_egt_desc &D) {
  if (D.dtype != TGT_BF16 && D.dtype != TGT_F16) return false;
  if (D.H % 2 || D.H > 64 || D.d > EF_DMAX) return false;
  if ((D.H * D.d) % 8 || D.ld_qkv % 8 || D.ld_eg % 2) return false;
  const size_t smem = (size_t)EF_CHUNK * 2 * D.H * D.d * 2;
  return smem <= 200 * 1024;
}

size_t egt_fast_workspace(const tgt_egt_desc &D) {
  return D.attend ? (size_t)D.B * D.N * D.N * D.H * 2 + 256 : 0;
}

template <typename T>
static int fwd_t(const tgt_egt_desc &D, const void *qkv, const void *eg, const float *mask, const float *src, void *hhat,
                 void *vatt, float *stats, cudaStream_t st) {
  const size_t smem = 2 * (size_t)EF_FCHUNK * (D.attend ? 2 : 1) * D.H * D.d * 2;
  dim3 grid((D.N + EF_ROWS - 1) / EF_ROWS, D.B);
#define L(HC, DC, AT)                                                                                                     \
  do {                                                                                                                \
    TGT_CUDA_OK(cudaFuncSetAttribute(egt_fwd_fast<T, HC, DC, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    egt_fwd_fast<T, HC, DC, AT><<<grid, EF_ROWS * 32, smem, st>>>(D, (const T *)qkv, (const T *)eg, mask, src, (T *)hhat,   \
                                                              (T *)vatt, stats);                                      \
  } while (0)
  // the shipped TGT geometry (Wn = 768, 64 heads) is fully specialised: all strides become immediates
  if (D.H == 64 && D.d == 12 && D.attend) L(64, 12, 1); else if (D.H == 64 && D.d == 12) L(64, 12, 0); else L(0, 0, -1);
#undef L
  return check_launch("egt_fwd_fast");
}

template <typename T>
static int bwd_t(const tgt_egt_desc &D, const void *qkv, const void *eg, const float *mask, const float *src,
                 const float *stats, const void *vatt, const void *dhhat, const void *dvatt, void *dqkv, void *deg,
                 void *ws, cudaStream_t st) {
  const size_t smem = (size_t)EF_CHUNK * (D.attend ? 2 : 1) * D.H * D.d * 2;
  dim3 grid((D.N + EF_ROWS - 1) / EF_ROWS, D.B);
  T *aw = (T *)(((uintptr_t)ws + 255) & ~(uintptr_t)255);
#define L(HC, DC, AT)                                                                                                          \
  do {                                                                                                                     \
    TGT_CUDA_OK(cudaFuncSetAttribute(egt_bwd_rows_fast<T, HC, DC, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    TGT_CUDA_OK(cudaFuncSetAttribute(egt_bwd_cols_fast<T, HC, DC, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    egt_bwd_rows_fast<T, HC, DC, AT><<<dim3((D.N + EF_BROWS - 1) / EF_BROWS, D.B), EF_BROWS * 32, smem, st>>>(D, (const T *)qkv, (const T *)eg, mask, src, stats,       \
                                                                   (const T *)vatt, (const T *)dhhat, (const T *)dvatt,    \
                                                                   (T *)dqkv, (T *)deg, aw);                               \
    if (int e = check_launch("egt_bwd_rows_fast")) return e;                                                               \
    egt_bwd_cols_fast<T, HC, DC, AT><<<grid, EF_ROWS * 32, smem, st>>>(D, (const T *)qkv, (const T *)dvatt, (const T *)deg, aw,  \
                                                                   (T *)dqkv);                                             \
  } while (0)
  if (D.H == 64 && D.d == 12 && D.attend) L(64, 12, 1); else if (D.H == 64 && D.d == 12) L(64, 12, 0); else L(0, 0, -1);
#undef L
  return check_launch("egt_bwd_cols_fast");
}

int egt_fwd_fast_launch(const tgt_egt_desc &D, const void *qkv, const void *eg, const float *mask, const float *src,
                        void *hhat, void *vatt, float *stats, cudaStream_t st) {
  if (((uintptr_t)qkv) & 15) return fail("egt_attn_fwd: qkv must be 16-byte aligned");
  if (D.dtype == TGT_BF16) return fwd_t<__nv_bfloat16>(D, qkv, eg, mask, src, hhat, vatt, stats, st);
  return fwd_t<__half>(D, qkv, eg, mask, src, hhat, vatt, stats, st);
}

int egt_bwd_fast_launch(const tgt_egt_desc &D, const void *qkv, const void *eg, const float *mask, const float *src,
                        const float *stats, const void *vatt, const void *dhhat, const void *dvatt, void *dqkv,
                        void *deg, void *ws, size_t ws_bytes, cudaStream_t st) {
  if (D.attend && (!ws || ws_bytes < egt_fast_workspace(D)))
    return fail("egt_attn_bwd: workspace too small (%zu < %zu bytes)", ws_bytes, egt_fast_workspace(D));
  if (D.attend && !vatt) return fail("egt_attn_bwd: the forward output vatt is required");
  if ((((uintptr_t)qkv) | ((uintptr_t)dvatt)) & 15) return fail("egt_attn_bwd: qkv / dvatt must be 16-byte aligned");
  if (D.dtype == TGT_BF16) return bwd_t<__nv_bfloat16>(D, qkv, eg, mask, src, stats, vatt, dhhat, dvatt, dqkv, deg, ws, st);
  return bwd_t<__half>(D, qkv, eg, mask, src, stats, vatt, dhhat, dvatt, dqkv, deg, ws, st);
}

}  // namespace tgt
