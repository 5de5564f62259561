// Tensor-core triplet attention core (bf16 / fp16, head dim 16, N <= 64) for sm_100a.
//
// Replaces lib/tgt/layers/triplet.py:213-227 and 232-246 (reference): per (graph b, head h, direction) an
// N x N x d attention for every junction atom j, with a bias/gate/mask tile that does NOT depend on j.
// Design (DESIGN.md "Kernels / triplet attention core"):
//   * one CTA = (h, dir, b), 4 warps x 16 query rows, loops over j.  The bias (E+mask) and gate sigmoid(G+mask)
//     tiles are loaded ONCE into registers in mma-accumulator layout and reused for all N junctions; in the
//     backward the dE / dG reductions over j accumulate in registers the same way (no atomics, deterministic).
//   * Q/K/V (and dO) 64x16 tiles stream through a 4-stage cp.async ring in shared memory; S = QK^T, PV,
//     dA = dO V^T, dQ = dS K run as mma.sync m16n8k16 with fp32 accumulation, P stays in registers
//     (accumulator -> A-fragment re-use), softmax in the exp2 domain with one MUFU per score.
//   * dK / dV reduce over the query index, i.e. across warps: dS and A are exchanged through a swizzled
//     shared-memory tile and every warp then owns 16 keys.
//   * the O(N^3 H) tensor never exists in memory.  The work per score is ~1 exp2 + ~10 FMA-class ops versus
//     0.25 tensor-core FMA-equivalents, so the kernel is bound by the SM's MUFU/FMA issue rate and by HBM, not by
//     the tensor pipe: that is why this core uses register-resident mma.sync fragments instead of tcgen05/TMEM
//     (a TMEM round trip per 64x64 score tile would add traffic without removing the exp/FMA work).
//   * a prep kernel transposes the [R, H] bias / gate columns of the projection into [b,dir,h,i,k] tiles
//     (adding the mask, applying the sigmoid), a post kernel scatters dE/dG back.  Both live in the caller-provided
//     workspace.
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

// ------------------------------------------------------------------------------------------------ forward
constexpr int FWD_STAGES = 4;
constexpr int FWD_STAGE_BYTES = 3 * TN * HD * 2;      // Q, K, V tiles: 6 KB

template <typename T>
__global__ void __launch_bounds__(128, 4)
tri_attn_fwd_mma(const tgt_triplet_attn_desc D, const T *__restrict__ proj, const float *__restrict__ ws_e,
                 const __half *__restrict__ ws_g, T *__restrict__ va, float *__restrict__ stats) {
  __shared__ __align__(128) unsigned char smem[FWD_STAGES * FWD_STAGE_BYTES];
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int m0 = warp * 16;
  const int64_t ld = D.ld, ldva = 2 * (int64_t)H * HD;
  const uint32_t sbase = smem_u32(smem);

  // bias (log2 domain) and gate fragments, constant over j
  float eb[8][4];
  uint32_t gt[8][2];
  {
    const int64_t tbase = ((int64_t)(b * 2 + dir) * H + h) * TN * TN;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const int col = nt * 8 + 2 * q;
      const float2 e0 = *reinterpret_cast<const float2 *>(ws_e + tbase + (m0 + g) * TN + col);
      const float2 e1 = *reinterpret_cast<const float2 *>(ws_e + tbase + (m0 + g + 8) * TN + col);
      eb[nt][0] = e0.x; eb[nt][1] = e0.y; eb[nt][2] = e1.x; eb[nt][3] = e1.y;
      gt[nt][0] = *reinterpret_cast<const uint32_t *>(ws_g + tbase + (m0 + g) * TN + col);
      gt[nt][1] = *reinterpret_cast<const uint32_t *>(ws_g + tbase + (m0 + g + 8) * TN + col);
    }
  }
  float c1r0, c1r1;
  fix_fully_masked_rows(eb, D.scale * LOG2E, c1r0, c1r1);

  RowMap rm{(int64_t)b * N, N, 0, dir};
  const int lrow = tid >> 1, lchunk = tid & 1;        // this thread's cp.async slot in each 64x16 tile
  const bool lvalid = lrow < N;
  const int lr = lvalid ? lrow : 0;
  const int offq = D.off_q[dir] + h * HD + lchunk * 8, offk = D.off_k[dir] + h * HD + lchunk * 8,
            offv = D.off_v[dir] + h * HD + lchunk * 8;
  const uint32_t ldst = tile_off(lrow, lchunk);

  auto issue = [&](int j) {
    if (j < N) {
      rm.j = j;
      const uint32_t st = sbase + (j % FWD_STAGES) * FWD_STAGE_BYTES;
      const int64_t kr = rm.krow(lr);
      cp_async16(st + ldst, proj + rm.qrow(lr) * ld + offq, lvalid);
      cp_async16(st + TN * HD * 2 + ldst, proj + kr * ld + offk, lvalid);
      cp_async16(st + 2 * TN * HD * 2 + ldst, proj + kr * ld + offv, lvalid);
    }
    cp_async_commit();
  };

#pragma unroll
  for (int s = 0; s < FWD_STAGES - 1; ++s) issue(s);

  for (int j = 0; j < N; ++j) {
    cp_async_wait<FWD_STAGES - 2>();
    __syncthreads();
    issue(j + FWD_STAGES - 1);
    const uint32_t st = sbase + (j % FWD_STAGES) * FWD_STAGE_BYTES;
    const uint32_t sQ = st, sK = st + TN * HD * 2, sV = st + 2 * TN * HD * 2;

    uint32_t qa[4];
    load_a_rows(qa, sQ, m0, lane);
    float s[8][4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t kb[4];
      load_b_nk(kb, sK, p * 16, lane);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        s[2 * p + u][0] = s[2 * p + u][1] = s[2 * p + u][2] = s[2 * p + u][3] = 0.f;
        Mma<T>::run(s[2 * p + u], qa, kb[2 * u], kb[2 * u + 1]);
      }
    }
    // logits in the log2 domain, row max
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = fmaf(s[nt][0], c1r0, eb[nt][0]);
      s[nt][1] = fmaf(s[nt][1], c1r0, eb[nt][1]);
      s[nt][2] = fmaf(s[nt][2], c1r1, eb[nt][2]);
      s[nt][3] = fmaf(s[nt][3], c1r1, eb[nt][3]);
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float l0 = 0.f, l1 = 0.f;
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float p0 = fast_exp2(s[nt][0] - mx0), p1 = fast_exp2(s[nt][1] - mx0);
      const float p2 = fast_exp2(s[nt][2] - mx1), p3 = fast_exp2(s[nt][3] - mx1);
      l0 += p0 + p1;
      l1 += p2 + p3;
      const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0]));
      const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1]));
      pa[nt >> 1][(nt & 1) * 2 + 0] = Mma<T>::pack(p0 * g0.x, p1 * g0.y);
      pa[nt >> 1][(nt & 1) * 2 + 1] = Mma<T>::pack(p2 * g1.x, p3 * g1.y);
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;

    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint32_t vb[4];
      load_b_kn(vb, sV, t * 16, lane);
      Mma<T>::run(o[0], pa[t], vb[0], vb[1]);
      Mma<T>::run(o[1], pa[t], vb[2], vb[3]);
    }
    rm.j = j;
    const int i0 = m0 + g, i1 = m0 + g + 8;
    if (i0 < N) {
      T *dst = va + rm.qrow(i0) * ldva + (int64_t)dir * H * HD + h * HD + 2 * q;
      *reinterpret_cast<uint32_t *>(dst) = Mma<T>::pack(o[0][0] * inv0, o[0][1] * inv0);
      *reinterpret_cast<uint32_t *>(dst + 8) = Mma<T>::pack(o[1][0] * inv0, o[1][1] * inv0);
    }
    if (i1 < N) {
      T *dst = va + rm.qrow(i1) * ldva + (int64_t)dir * H * HD + h * HD + 2 * q;
      *reinterpret_cast<uint32_t *>(dst) = Mma<T>::pack(o[0][2] * inv1, o[0][3] * inv1);
      *reinterpret_cast<uint32_t *>(dst + 8) = Mma<T>::pack(o[1][2] * inv1, o[1][3] * inv1);
    }
    if (q == 0) {          // one float per row: log2-domain log-sum-exp  (P = exp2(x - lse2))
      float *stp = stats + (((int64_t)(b * 2 + dir) * H + h) * N + j) * N;
      if (i0 < N) stp[i0] = mx0 + __log2f(l0);
      if (i1 < N) stp[i1] = mx1 + __log2f(l1);
    }
  }
  cp_async_wait<0>();
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int BWD_STAGES = 3;
constexpr int BWD_STAGE_BYTES = 4 * TN * HD * 2;      // Q, K, V, dO tiles: 8 KB
constexpr int XCH_BYTES = TN * TN * 2;                // one 64x64 16-bit exchange tile: 8 KB
constexpr int BWD_SMEM = BWD_STAGES * BWD_STAGE_BYTES + 4 * XCH_BYTES;   // 24 KB + 2 x (dS, A) = 56 KB

template <typename T, int MINB>
__global__ void __launch_bounds__(128, MINB)
tri_attn_bwd_mma(const tgt_triplet_attn_desc D, const T *__restrict__ proj, const float *__restrict__ ws_e,
                 const __half *__restrict__ ws_g, const T *__restrict__ dva, const float *__restrict__ stats,
                 T *__restrict__ dproj, float *__restrict__ ws_de, float *__restrict__ ws_dg) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int m0 = warp * 16;
  const int64_t ld = D.ld, ldva = 2 * (int64_t)H * HD;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t xbase = sbase + BWD_STAGES * BWD_STAGE_BYTES;

  float eb[8][4];
  uint32_t gt[8][2];
  const int64_t tbase = ((int64_t)(b * 2 + dir) * H + h) * TN * TN;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * q;
    const float2 e0 = *reinterpret_cast<const float2 *>(ws_e + tbase + (m0 + g) * TN + col);
    const float2 e1 = *reinterpret_cast<const float2 *>(ws_e + tbase + (m0 + g + 8) * TN + col);
    eb[nt][0] = e0.x; eb[nt][1] = e0.y; eb[nt][2] = e1.x; eb[nt][3] = e1.y;
    gt[nt][0] = *reinterpret_cast<const uint32_t *>(ws_g + tbase + (m0 + g) * TN + col);
    gt[nt][1] = *reinterpret_cast<const uint32_t *>(ws_g + tbase + (m0 + g + 8) * TN + col);
  }
  float de[8][4], dg[8][4];          // sum_j dS   and   sum_j dA * P
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int c = 0; c < 4; ++c) de[nt][c] = dg[nt][c] = 0.f;

  float c1r0, c1r1;
  fix_fully_masked_rows(eb, D.scale * LOG2E, c1r0, c1r1);
  RowMap rm{(int64_t)b * N, N, 0, dir};
  const int lrow = tid >> 1, lchunk = tid & 1;
  const bool lvalid = lrow < N;
  const int lr = lvalid ? lrow : 0;
  const int offq = D.off_q[dir] + h * HD, offk = D.off_k[dir] + h * HD, offv = D.off_v[dir] + h * HD;
  const int offo = dir * H * HD + h * HD;
  const uint32_t ldst = tile_off(lrow, lchunk);
  const int i0 = m0 + g, i1 = m0 + g + 8;
  const float *stb = stats + (((int64_t)(b * 2 + dir) * H + h) * N) * N;

  auto issue = [&](int j) {
    if (j < N) {
      rm.j = j;
      const uint32_t st = sbase + (j % BWD_STAGES) * BWD_STAGE_BYTES;
      const int64_t kr = rm.krow(lr), qr = rm.qrow(lr);
      cp_async16(st + ldst, proj + qr * ld + offq + lchunk * 8, lvalid);
      cp_async16(st + TN * HD * 2 + ldst, proj + kr * ld + offk + lchunk * 8, lvalid);
      cp_async16(st + 2 * TN * HD * 2 + ldst, proj + kr * ld + offv + lchunk * 8, lvalid);
      cp_async16(st + 3 * TN * HD * 2 + ldst, dva + qr * ldva + offo + lchunk * 8, lvalid);
    }
    cp_async_commit();
  };
  // rows beyond N (tile padding) get lse2 = +inf  ->  P = exp2(-inf) = 0 for the whole row
  auto load_stats = [&](int j, float &a, float &c) {
    a = INFINITY;
    c = INFINITY;
    if (j < N) {
      if (i0 < N) a = stb[(int64_t)j * N + i0];
      if (i1 < N) c = stb[(int64_t)j * N + i1];
    }
  };

#pragma unroll
  for (int s = 0; s < BWD_STAGES - 1; ++s) issue(s);
  float st0, st1;
  load_stats(0, st0, st1);

  for (int j = 0; j < N; ++j) {
    cp_async_wait<BWD_STAGES - 2>();
    __syncthreads();                                   // (A) stage j landed; everyone is done with iteration j-1
    issue(j + BWD_STAGES - 1);
    const float lse0 = st0, lse1 = st1;
    load_stats(j + 1, st0, st1);
    const uint32_t st = sbase + (j % BWD_STAGES) * BWD_STAGE_BYTES;
    const uint32_t sQ = st, sK = st + TN * HD * 2, sV = st + 2 * TN * HD * 2, sO = st + 3 * TN * HD * 2;
    const uint32_t xS = xbase + (j & 1) * 2 * XCH_BYTES, xA = xS + XCH_BYTES;

    uint32_t qa[4], oa[4];
    load_a_rows(qa, sQ, m0, lane);
    load_a_rows(oa, sO, m0, lane);
    float s[8][4], da[8][4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t kb[4], vb[4];
      load_b_nk(kb, sK, p * 16, lane);
      load_b_nk(vb, sV, p * 16, lane);
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int nt = 2 * p + u;
        s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        da[nt][0] = da[nt][1] = da[nt][2] = da[nt][3] = 0.f;
        Mma<T>::run(s[nt], qa, kb[2 * u], kb[2 * u + 1]);
        Mma<T>::run(da[nt], oa, vb[2 * u], vb[2 * u + 1]);
      }
    }
    // P (normalised), t = dA*g*P, delta = rowsum(t)
    float dl0 = 0.f, dl1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0]));
      const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1]));
      const float gg[4] = {g0.x, g0.y, g1.x, g1.y};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float p = fast_exp2(fmaf(s[nt][c], c < 2 ? c1r0 : c1r1, eb[nt][c]) - (c < 2 ? lse0 : lse1));
        const float dap = da[nt][c] * p;
        dg[nt][c] += dap;
        s[nt][c] = p;                      // P
        da[nt][c] = dap * gg[c];           // t = dA * g * P
      }
      dl0 += da[nt][0] + da[nt][1];
      dl1 += da[nt][2] + da[nt][3];
    }
    dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1);
    dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
    dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1);
    dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
    // dS = t - P*delta ; A = P*g ; exchange through shared memory ; dS also as A-fragments for dQ
    uint32_t dsa[4][4];
    unsigned char *xs = smem + BWD_STAGES * BWD_STAGE_BYTES + (j & 1) * 2 * XCH_BYTES;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0]));
      const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1]));
      const float d0 = fmaf(-s[nt][0], dl0, da[nt][0]), d1 = fmaf(-s[nt][1], dl0, da[nt][1]);
      const float d2 = fmaf(-s[nt][2], dl1, da[nt][2]), d3 = fmaf(-s[nt][3], dl1, da[nt][3]);
      de[nt][0] += d0; de[nt][1] += d1; de[nt][2] += d2; de[nt][3] += d3;
      const uint32_t ds01 = Mma<T>::pack(d0, d1), ds23 = Mma<T>::pack(d2, d3);
      dsa[nt >> 1][(nt & 1) * 2 + 0] = ds01;
      dsa[nt >> 1][(nt & 1) * 2 + 1] = ds23;
      const uint32_t a01 = Mma<T>::pack(s[nt][0] * g0.x, s[nt][1] * g0.y);
      const uint32_t a23 = Mma<T>::pack(s[nt][2] * g1.x, s[nt][3] * g1.y);
      *reinterpret_cast<uint32_t *>(xs + xch_off(i0, nt) + q * 4) = ds01;
      *reinterpret_cast<uint32_t *>(xs + xch_off(i1, nt) + q * 4) = ds23;
      *reinterpret_cast<uint32_t *>(xs + XCH_BYTES + xch_off(i0, nt) + q * 4) = a01;
      *reinterpret_cast<uint32_t *>(xs + XCH_BYTES + xch_off(i1, nt) + q * 4) = a23;
    }
    // dQ = scale * dS K   (rows of this warp)
    {
      float dq[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint32_t kb[4];
        load_b_kn(kb, sK, t * 16, lane);
        Mma<T>::run(dq[0], dsa[t], kb[0], kb[1]);
        Mma<T>::run(dq[1], dsa[t], kb[2], kb[3]);
      }
      rm.j = j;
      const float sc = D.scale;
      if (i0 < N) {
        T *dst = dproj + rm.qrow(i0) * ld + offq + 2 * q;
        *reinterpret_cast<uint32_t *>(dst) = Mma<T>::pack(dq[0][0] * sc, dq[0][1] * sc);
        *reinterpret_cast<uint32_t *>(dst + 8) = Mma<T>::pack(dq[1][0] * sc, dq[1][1] * sc);
      }
      if (i1 < N) {
        T *dst = dproj + rm.qrow(i1) * ld + offq + 2 * q;
        *reinterpret_cast<uint32_t *>(dst) = Mma<T>::pack(dq[0][2] * sc, dq[0][3] * sc);
        *reinterpret_cast<uint32_t *>(dst + 8) = Mma<T>::pack(dq[1][2] * sc, dq[1][3] * sc);
      }
    }
    __syncthreads();                                   // (B) dS / A tiles complete
    // dK = scale * dS^T Q ; dV = A^T dO   (this warp owns keys m0 .. m0+15)
    {
      float dk[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
      float dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        uint32_t at[4], qb[4], ob[4];
        load_a_xt(at, xS, m0, t * 16, lane);
        load_b_kn(qb, sQ, t * 16, lane);
        Mma<T>::run(dk[0], at, qb[0], qb[1]);
        Mma<T>::run(dk[1], at, qb[2], qb[3]);
        load_a_xt(at, xA, m0, t * 16, lane);
        load_b_kn(ob, sO, t * 16, lane);
        Mma<T>::run(dv[0], at, ob[0], ob[1]);
        Mma<T>::run(dv[1], at, ob[2], ob[3]);
      }
      const float sc = D.scale;
      if (i0 < N) {
        const int64_t kr = rm.krow(i0);
        T *dkp = dproj + kr * ld + offk + 2 * q, *dvp = dproj + kr * ld + offv + 2 * q;
        *reinterpret_cast<uint32_t *>(dkp) = Mma<T>::pack(dk[0][0] * sc, dk[0][1] * sc);
        *reinterpret_cast<uint32_t *>(dkp + 8) = Mma<T>::pack(dk[1][0] * sc, dk[1][1] * sc);
        *reinterpret_cast<uint32_t *>(dvp) = Mma<T>::pack(dv[0][0], dv[0][1]);
        *reinterpret_cast<uint32_t *>(dvp + 8) = Mma<T>::pack(dv[1][0], dv[1][1]);
      }
      if (i1 < N) {
        const int64_t kr = rm.krow(i1);
        T *dkp = dproj + kr * ld + offk + 2 * q, *dvp = dproj + kr * ld + offv + 2 * q;
        *reinterpret_cast<uint32_t *>(dkp) = Mma<T>::pack(dk[0][2] * sc, dk[0][3] * sc);
        *reinterpret_cast<uint32_t *>(dkp + 8) = Mma<T>::pack(dk[1][2] * sc, dk[1][3] * sc);
        *reinterpret_cast<uint32_t *>(dvp) = Mma<T>::pack(dv[0][2], dv[0][3]);
        *reinterpret_cast<uint32_t *>(dvp + 8) = Mma<T>::pack(dv[1][2], dv[1][3]);
      }
    }
  }
  cp_async_wait<0>();
  // dE = sum_j dS ; dG = g (1 - g) sum_j dA P
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int col = nt * 8 + 2 * q;
    const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0]));
    const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1]));
    *reinterpret_cast<float2 *>(ws_de + tbase + (m0 + g) * TN + col) = make_float2(de[nt][0], de[nt][1]);
    *reinterpret_cast<float2 *>(ws_de + tbase + (m0 + g + 8) * TN + col) = make_float2(de[nt][2], de[nt][3]);
    *reinterpret_cast<float2 *>(ws_dg + tbase + (m0 + g) * TN + col) =
        make_float2(dg[nt][0] * g0.x * (1.f - g0.x), dg[nt][1] * g0.y * (1.f - g0.y));
    *reinterpret_cast<float2 *>(ws_dg + tbase + (m0 + g + 8) * TN + col) =
        make_float2(dg[nt][2] * g1.x * (1.f - g1.x), dg[nt][3] * g1.y * (1.f - g1.y));
  }
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
// kernel policy 0 (default) / 3 / 4 / 5: TMA-staged kernels; policy 2: the cp.async-staged kernels of this file
static bool use_tma() { return g_policy.load() != 2 && g_policy.load() != 1 && triplet_attn_tma_available(); }
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
  KernelTimerScope ts("tri_attn_fwd_mma", st);
  tri_attn_fwd_mma<T><<<dim3(D.H, 2, D.B), 128, 0, st>>>(D, (const T *)proj, w.e, w.g, (T *)va, stats);
  return check_launch("tri_attn_fwd_mma");
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
  static const int minb = [] { const char *v = getenv("TGT_TRI_BWD_MINB"); return v ? atoi(v) : 2; }();   // tuning knob
#define L(MB)                                                                                                         \
  do {                                                                                                                \
    TGT_CUDA_OK(cudaFuncSetAttribute(tri_attn_bwd_mma<T, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM)); \
    tri_attn_bwd_mma<T, MB><<<dim3(D.H, 2, D.B), 128, BWD_SMEM, st>>>(D, (const T *)proj, w.e, w.g, (const T *)dva,    \
                                                                      stats, (T *)dproj, w.de, w.dg);                 \
  } while (0)
  {
    KernelTimerScope ts("tri_attn_bwd_mma", st);
    if (minb == 3) L(3); else L(2);
  }
#undef L
  if (int e = check_launch("tri_attn_bwd_mma")) return e;
  if (bias_gate_h16_ok(D)) tri_post_bias_gate_h16<T><<<dim3(D.N, 2, D.B), 256, 0, st>>>(D, w.de, w.dg, (T *)dproj);
  else tri_post_bias_gate<T><<<dim3(D.N, 2, D.B), 256, psm, st>>>(D, w.de, w.dg, (T *)dproj);
  return check_launch("tri_post_bias_gate");
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
