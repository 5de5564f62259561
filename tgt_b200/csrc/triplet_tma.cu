// Triplet attention core with TMA staging (bf16 / fp16, head dim 16, N <= 64) for sm_100a.
//
// Same math, CTA decomposition (one CTA = (head, direction, graph), loop over the junction atom j, bias / gate /
// dE / dG tiles resident in registers) and parity contract as triplet_mma.cu; what changes is how bytes move:
//   * the Q / K / V / dO 64x16 tiles of junction j are fetched by ONE thread with cp.async.bulk.tensor (TMA) through
//     4-D tensor maps over the [B, N, N, C] projection / gradient tensors -- box {16 ch, 1, 64, 1} is "column j"
//     (queries, and keys/values of the outward direction), box {16 ch, 64, 1, 1} is "row j" (keys/values of the inward
//     direction).  Rows >= N are out of bounds: zero-filled on load, dropped on store.  SWIZZLE_32B reproduces the
//     ldmatrix-friendly tile layout of triplet_common.cuh (chunk ^= row bit 2), completion is tracked by mbarriers;
//   * results (Va; dQ, dK, dV) go from mma accumulators to shared memory with stmatrix and leave with TMA tensor
//     stores (full 32-byte rows instead of 4-byte scattered stores); the dS / A exchange tiles use stmatrix too.
// Net effect: ~100 fewer issue slots per thread and iteration (no address arithmetic, no cp.async, no predicated
// 4-byte stores) and sector-exact HBM writes.  The tests cross-check this family against the tcgen05 / TMEM family
// (triplet_tc.cu) and the generic SIMT kernels.
#include "triplet_common.cuh"
#include <stdlib.h>

namespace tgt {

constexpr int TILE_BYTES = TN * HD * 2;                       // one 64 x 16 operand tile: 2 KB

// ------------------------------------------------------------------------------------------------ forward
constexpr int TF_STAGES = 4;
constexpr int TF_STAGE_BYTES = 3 * TILE_BYTES;                // Q, K, V
constexpr int TF_SMEM = TF_STAGES * TF_STAGE_BYTES + 2 * TILE_BYTES + 64 + 1024;   // + 2 output tiles + barriers + align

template <typename T>
__global__ void __launch_bounds__(128, 4)
tri_attn_fwd_tma(const tgt_triplet_attn_desc D, const __grid_constant__ CUtensorMap mPcol,
                 const __grid_constant__ CUtensorMap mProw, const __grid_constant__ CUtensorMap mVA,
                 const float *__restrict__ ws_e, const __half *__restrict__ ws_g, float *__restrict__ stats) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int m0 = warp * 16;
  const uint32_t sOut = sbase + TF_STAGES * TF_STAGE_BYTES;
  const uint32_t bar_full = sOut + 2 * TILE_BYTES;

  if (tid == 0) {
    tma_prefetch_desc(&mPcol);
    tma_prefetch_desc(&mProw);
    tma_prefetch_desc(&mVA);
    for (int s = 0; s < TF_STAGES; ++s) mbar_init(bar_full + s * 8, 1);
    fence_barrier_init();
  }

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

  const int cq = D.off_q[dir] + h * HD, ck = D.off_k[dir] + h * HD, cv = D.off_v[dir] + h * HD;
  const int co = dir * H * HD + h * HD;
  auto issue = [&](int j) {                   // thread 0 only
    if (j < N) {
      const uint32_t st = sbase + (j % TF_STAGES) * TF_STAGE_BYTES, bar = bar_full + (j % TF_STAGES) * 8;
      mbar_expect_tx(bar, TF_STAGE_BYTES);
      tma_load_4d(&mPcol, bar, st, cq, j, 0, b);
      if (dir == 0) {
        tma_load_4d(&mProw, bar, st + TILE_BYTES, ck, 0, j, b);
        tma_load_4d(&mProw, bar, st + 2 * TILE_BYTES, cv, 0, j, b);
      } else {
        tma_load_4d(&mPcol, bar, st + TILE_BYTES, ck, j, 0, b);
        tma_load_4d(&mPcol, bar, st + 2 * TILE_BYTES, cv, j, 0, b);
      }
    }
  };
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TF_STAGES - 1; ++s) issue(s);
  }

  for (int j = 0; j < N; ++j) {
    if (tid == 0) tma_store_wait_read();            // the output tile written two iterations ago has left smem
    mbar_wait(bar_full + (j % TF_STAGES) * 8, (uint32_t)((j / TF_STAGES) & 1));
    __syncthreads();                                // stage j landed; everyone is done with iteration j-1
    if (tid == 0) {
      issue(j + TF_STAGES - 1);
      if (j > 0) {
        tma_store_4d(&mVA, sOut + ((j - 1) & 1) * TILE_BYTES, co, j - 1, 0, b);
        tma_store_commit();
      }
    }
    const uint32_t st = sbase + (j % TF_STAGES) * TF_STAGE_BYTES;
    const uint32_t sQ = st, sK = st + TILE_BYTES, sV = st + 2 * TILE_BYTES;

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
    // logits, row max, exp, gate: packed f32x2 arithmetic (two accumulator columns of a row per instruction)
    const float2 c1p0 = make_float2(c1r0, c1r0), c1p1 = make_float2(c1r1, c1r1);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      float2 &s01 = *reinterpret_cast<float2 *>(&s[nt][0]), &s23 = *reinterpret_cast<float2 *>(&s[nt][2]);
      s01 = __ffma2_rn(s01, c1p0, *reinterpret_cast<const float2 *>(&eb[nt][0]));
      s23 = __ffma2_rn(s23, c1p1, *reinterpret_cast<const float2 *>(&eb[nt][2]));
      mx0 = fmaxf(mx0, fmaxf(s01.x, s01.y));
      mx1 = fmaxf(mx1, fmaxf(s23.x, s23.y));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float2 nm0 = make_float2(-mx0, -mx0), nm1 = make_float2(-mx1, -mx1);
    float2 lp0 = make_float2(0.f, 0.f), lp1 = make_float2(0.f, 0.f);
    uint32_t pa[4][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 x01 = __fadd2_rn(*reinterpret_cast<float2 *>(&s[nt][0]), nm0);
      const float2 x23 = __fadd2_rn(*reinterpret_cast<float2 *>(&s[nt][2]), nm1);
      const float2 p01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
      const float2 p23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
      lp0 = __fadd2_rn(lp0, p01);
      lp1 = __fadd2_rn(lp1, p23);
      const float2 w01 = __fmul2_rn(p01, __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0])));
      const float2 w23 = __fmul2_rn(p23, __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1])));
      pa[nt >> 1][(nt & 1) * 2 + 0] = Mma<T>::pack(w01.x, w01.y);
      pa[nt >> 1][(nt & 1) * 2 + 1] = Mma<T>::pack(w23.x, w23.y);
    }
    float l0 = lp0.x + lp0.y, l1 = lp1.x + lp1.y;
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
    o[0][0] *= inv0; o[0][1] *= inv0; o[1][0] *= inv0; o[1][1] *= inv0;
    o[0][2] *= inv1; o[0][3] *= inv1; o[1][2] *= inv1; o[1][3] *= inv1;
    store_c_tile<T>(sOut + (j & 1) * TILE_BYTES, m0, lane, o, 1.f);
    fence_proxy_async();
    const int i0 = m0 + g, i1 = m0 + g + 8;
    if (q == 0) {          // one float per row: log2-domain log-sum-exp  (P = exp2(x - lse2))
      float *stp = stats + (((int64_t)(b * 2 + dir) * H + h) * N + j) * N;
      if (i0 < N) stp[i0] = mx0 + __log2f(l0);
      if (i1 < N) stp[i1] = mx1 + __log2f(l1);
    }
  }
  __syncthreads();
  if (tid == 0) {
    tma_store_4d(&mVA, sOut + ((N - 1) & 1) * TILE_BYTES, co, N - 1, 0, b);
    tma_store_commit();
    tma_store_wait_all();
  }
}

// ------------------------------------------------------------------------------------------------ backward
constexpr int TB_STAGE_BYTES = 4 * TILE_BYTES;               // Q, K, V, dO
constexpr int TB_XCH_BYTES = TN * TN * 2;                    // one 64x64 16-bit exchange tile: 8 KB

// ------------------------------------------------------------------------------------------------ backward, pipelined
constexpr int TP_STAGES = 4;
constexpr int TP_SMEM = TP_STAGES * TB_STAGE_BYTES + 4 * TB_XCH_BYTES + 6 * TILE_BYTES + 64 + 1024;

// BIAS: also accumulate the column sums of everything this CTA contributes to d(proj) -- its 16 dQ / dK / dV channels and
// its dE / dG column -- into dbias[ld] (fp32, zero on entry, atomics): that is the bias gradient of the projection, so the
// weight-gradient GEMM needs no augmented [LN(x) | 1] operand (264 -> 256 columns: 1.08 -> 0.80 ms in cuBLAS).
template <typename T, bool BIAS>
__global__ void __launch_bounds__(128, 2)
tri_attn_bwd_tma_pipe(const tgt_triplet_attn_desc D, const __grid_constant__ CUtensorMap mPcol,
                 const __grid_constant__ CUtensorMap mProw, const __grid_constant__ CUtensorMap mDVA,
                 const __grid_constant__ CUtensorMap mDPcol, const __grid_constant__ CUtensorMap mDProw,
                 const float *__restrict__ ws_e, const __half *__restrict__ ws_g, const float *__restrict__ stats,
                 float *__restrict__ ws_de, float *__restrict__ ws_dg, float *__restrict__ dbias) {
  float2 csq[2], csk[2], csv[2];     // BIAS: per-thread column sums (columns 2q, 2q+1 of the two 8-column tiles), unscaled
  if constexpr (BIAS) {
    csq[0] = csq[1] = csk[0] = csk[1] = csv[0] = csv[1] = make_float2(0.f, 0.f);
  }
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int m0 = warp * 16;
  const uint32_t xbase = sbase + TP_STAGES * TB_STAGE_BYTES;
  const uint32_t sOut0 = xbase + 4 * TB_XCH_BYTES;            // 2 x (dQ, dK, dV) staging tiles
  const uint32_t bar_full = sOut0 + 6 * TILE_BYTES;

  if (tid == 0) {
    tma_prefetch_desc(&mPcol);
    tma_prefetch_desc(&mProw);
    tma_prefetch_desc(&mDVA);
    tma_prefetch_desc(&mDPcol);
    tma_prefetch_desc(&mDProw);
    for (int s = 0; s < TP_STAGES; ++s) mbar_init(bar_full + s * 8, 1);
    fence_barrier_init();
  }

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
  const int cq = D.off_q[dir] + h * HD, ck = D.off_k[dir] + h * HD, cv = D.off_v[dir] + h * HD;
  const int co = dir * H * HD + h * HD;
  const int i0 = m0 + g, i1 = m0 + g + 8;
  const float *stb = stats + (((int64_t)(b * 2 + dir) * H + h) * N) * N;

  auto issue = [&](int j) {                   // thread 0 only
    if (j < N) {
      const uint32_t st = sbase + (j % TP_STAGES) * TB_STAGE_BYTES, bar = bar_full + (j % TP_STAGES) * 8;
      mbar_expect_tx(bar, TB_STAGE_BYTES);
      tma_load_4d(&mPcol, bar, st, cq, j, 0, b);
      if (dir == 0) {
        tma_load_4d(&mProw, bar, st + TILE_BYTES, ck, 0, j, b);
        tma_load_4d(&mProw, bar, st + 2 * TILE_BYTES, cv, 0, j, b);
      } else {
        tma_load_4d(&mPcol, bar, st + TILE_BYTES, ck, j, 0, b);
        tma_load_4d(&mPcol, bar, st + 2 * TILE_BYTES, cv, j, 0, b);
      }
      tma_load_4d(&mDVA, bar, st + 3 * TILE_BYTES, co, j, 0, b);
    }
  };
  // thread 0 only: results staged by iteration `it` (buffer it & 1): dQ of junction it, dK / dV of junction it - 1
  auto store_group = [&](int it) {
    const uint32_t so = sOut0 + (it & 1) * 3 * TILE_BYTES;
    if (it < N) tma_store_4d(&mDPcol, so, cq, it, 0, b);
    if (it > 0) {
      const int j = it - 1;
      if (dir == 0) {
        tma_store_4d(&mDProw, so + TILE_BYTES, ck, 0, j, b);
        tma_store_4d(&mDProw, so + 2 * TILE_BYTES, cv, 0, j, b);
      } else {
        tma_store_4d(&mDPcol, so + TILE_BYTES, ck, j, 0, b);
        tma_store_4d(&mDPcol, so + 2 * TILE_BYTES, cv, j, 0, b);
      }
    }
    tma_store_commit();
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

  __syncthreads();
  if (tid == 0) {
    issue(0);
    issue(1);
  }
  float sa0, sa1, sb0, sb1;                  // lse of the next two junctions (fetched two iterations ahead)
  load_stats(0, sa0, sa1);
  load_stats(1, sb0, sb1);
  // running pointers to the lse of junction j + 2 (rows i0 / i1): one add per junction instead of 64-bit index arithmetic
  const float *pl0 = stb + 2 * (int64_t)N + (i0 < N ? i0 : 0), *pl1 = stb + 2 * (int64_t)N + (i1 < N ? i1 : 0);
  const bool ok0 = i0 < N, ok1 = i1 < N;

  // Software pipeline over the junctions: iteration `it` runs the key pass (dK, dV) of junction it-1 and the score pass
  // (S, dA, P, dS, A, dQ) of junction it back to back, so ONE CTA barrier per junction separates "exchange tiles written"
  // from "exchange tiles read" (the exchange and staging tiles are double-buffered), and each warp has two independent
  // instruction streams between barriers.
  for (int it = 0; it <= N; ++it) {
    if (tid == 0) {
      issue(it + 2);                                   // slot of junction it-2: its key pass ended before the last barrier
      if (it > 0) store_group(it - 1);
    }
    const uint32_t sOut = sOut0 + (it & 1) * 3 * TILE_BYTES;
    if (it > 0) {
      const int jp = it - 1;
      const uint32_t stp = sbase + (jp % TP_STAGES) * TB_STAGE_BYTES;
      const uint32_t sQp = stp, sOp = stp + 3 * TILE_BYTES;
      const uint32_t xSp = xbase + (jp & 1) * 2 * TB_XCH_BYTES, xAp = xSp + TB_XCH_BYTES;
      // dK = scale * dS^T Q ; dV = A^T dO   (this warp owns keys m0 .. m0+15)
      {
        float dk[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  #pragma unroll
        for (int t = 0; t < 4; ++t) {
          uint32_t at[4], qb[4], ob[4];
          load_a_xt(at, xSp, m0, t * 16, lane);
          load_b_kn(qb, sQp, t * 16, lane);
          Mma<T>::run(dk[0], at, qb[0], qb[1]);
          Mma<T>::run(dk[1], at, qb[2], qb[3]);
          load_a_xt(at, xAp, m0, t * 16, lane);
          load_b_kn(ob, sOp, t * 16, lane);
          Mma<T>::run(dv[0], at, ob[0], ob[1]);
          Mma<T>::run(dv[1], at, ob[2], ob[3]);
        }
        store_c_tile<T>(sOut + TILE_BYTES, m0, lane, dk, D.scale);
        store_c_tile<T>(sOut + 2 * TILE_BYTES, m0, lane, dv, 1.f);
        if constexpr (BIAS) {
#pragma unroll
          for (int u = 0; u < 2; ++u) {      // rows g and g+8 of columns (2q, 2q+1): packed f32x2 adds
            csk[u] = __fadd2_rn(csk[u], __fadd2_rn(make_float2(dk[u][0], dk[u][1]), make_float2(dk[u][2], dk[u][3])));
            csv[u] = __fadd2_rn(csv[u], __fadd2_rn(make_float2(dv[u][0], dv[u][1]), make_float2(dv[u][2], dv[u][3])));
          }
        }
      }

    }
    if (it < N) {
      const int j = it;
      mbar_wait(bar_full + (j % TP_STAGES) * 8, (uint32_t)((j / TP_STAGES) & 1));
      const float lse0 = sa0, lse1 = sa1;
      sa0 = sb0;
      sa1 = sb1;
      sb0 = sb1 = INFINITY;
      if (j + 2 < N) {
        if (ok0) sb0 = *pl0;
        if (ok1) sb1 = *pl1;
      }
      pl0 += N;
      pl1 += N;
      const uint32_t st = sbase + (j % TP_STAGES) * TB_STAGE_BYTES;
      const uint32_t sQ = st, sK = st + TILE_BYTES, sV = st + 2 * TILE_BYTES, sO = st + 3 * TILE_BYTES;
      const uint32_t xS = xbase + (j & 1) * 2 * TB_XCH_BYTES, xA = xS + TB_XCH_BYTES;

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
      // P (normalised), t = dA*g*P, delta = rowsum(t).  The element-wise math runs on packed f32x2 instructions
      // (fma/add/mul.rn.f32x2: two accumulator columns of one row per issue slot); results are identical to scalar fp32.
      const float2 c1p0 = make_float2(c1r0, c1r0), c1p1 = make_float2(c1r1, c1r1);
      const float2 nl0 = make_float2(-lse0, -lse0), nl1 = make_float2(-lse1, -lse1);
      float2 dlp0 = make_float2(0.f, 0.f), dlp1 = make_float2(0.f, 0.f);
  #pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
        const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0]));
        const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1]));
        float2 &s01 = *reinterpret_cast<float2 *>(&s[nt][0]), &s23 = *reinterpret_cast<float2 *>(&s[nt][2]);
        float2 &a01 = *reinterpret_cast<float2 *>(&da[nt][0]), &a23 = *reinterpret_cast<float2 *>(&da[nt][2]);
        float2 &e01 = *reinterpret_cast<float2 *>(&eb[nt][0]), &e23 = *reinterpret_cast<float2 *>(&eb[nt][2]);
        float2 &q01 = *reinterpret_cast<float2 *>(&dg[nt][0]), &q23 = *reinterpret_cast<float2 *>(&dg[nt][2]);
        const float2 x01 = __ffma2_rn(s01, c1p0, __fadd2_rn(e01, nl0));
        const float2 x23 = __ffma2_rn(s23, c1p1, __fadd2_rn(e23, nl1));
        const float2 p01 = make_float2(fast_exp2(x01.x), fast_exp2(x01.y));
        const float2 p23 = make_float2(fast_exp2(x23.x), fast_exp2(x23.y));
        const float2 dap01 = __fmul2_rn(a01, p01), dap23 = __fmul2_rn(a23, p23);
        q01 = __fadd2_rn(q01, dap01);
        q23 = __fadd2_rn(q23, dap23);
        s01 = p01;                             // P
        s23 = p23;
        a01 = __fmul2_rn(dap01, g0);           // t = dA * g * P
        a23 = __fmul2_rn(dap23, g1);
        dlp0 = __fadd2_rn(dlp0, a01);
        dlp1 = __fadd2_rn(dlp1, a23);
      }
      float dl0 = dlp0.x + dlp0.y, dl1 = dlp1.x + dlp1.y;
      dl0 += __shfl_xor_sync(0xffffffffu, dl0, 1);
      dl0 += __shfl_xor_sync(0xffffffffu, dl0, 2);
      dl1 += __shfl_xor_sync(0xffffffffu, dl1, 1);
      dl1 += __shfl_xor_sync(0xffffffffu, dl1, 2);
      // dS = t - P*delta ; A = P*g ; exchange through shared memory (stmatrix) ; dS also as A-fragments for dQ
      const float2 nd0 = make_float2(-dl0, -dl0), nd1 = make_float2(-dl1, -dl1);
      uint32_t dsa[4][4];
  #pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t aa[4];
  #pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int nt = 2 * np + u;
          const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0]));
          const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1]));
          const float2 p01 = *reinterpret_cast<float2 *>(&s[nt][0]), p23 = *reinterpret_cast<float2 *>(&s[nt][2]);
          const float2 d01 = __ffma2_rn(p01, nd0, *reinterpret_cast<float2 *>(&da[nt][0]));
          const float2 d23 = __ffma2_rn(p23, nd1, *reinterpret_cast<float2 *>(&da[nt][2]));
          float2 &f01 = *reinterpret_cast<float2 *>(&de[nt][0]), &f23 = *reinterpret_cast<float2 *>(&de[nt][2]);
          f01 = __fadd2_rn(f01, d01);
          f23 = __fadd2_rn(f23, d23);
          dsa[np][u * 2 + 0] = Mma<T>::pack(d01.x, d01.y);
          dsa[np][u * 2 + 1] = Mma<T>::pack(d23.x, d23.y);
          const float2 w01 = __fmul2_rn(p01, g0), w23 = __fmul2_rn(p23, g1);
          aa[u * 2 + 0] = Mma<T>::pack(w01.x, w01.y);
          aa[u * 2 + 1] = Mma<T>::pack(w23.x, w23.y);
        }
        store_xch_pair(xS, m0, lane, 2 * np, dsa[np][0], dsa[np][1], dsa[np][2], dsa[np][3]);
        store_xch_pair(xA, m0, lane, 2 * np, aa[0], aa[1], aa[2], aa[3]);
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
        store_c_tile<T>(sOut, m0, lane, dq, D.scale);
        if constexpr (BIAS) {
#pragma unroll
          for (int u = 0; u < 2; ++u)
            csq[u] = __fadd2_rn(csq[u], __fadd2_rn(make_float2(dq[u][0], dq[u][1]), make_float2(dq[u][2], dq[u][3])));
        }
      }

    }
    fence_proxy_async();
    if (tid == 0) tma_store_wait_read();               // the group issued at the top of this iteration has left its buffer
    __syncthreads();
  }
  if (tid == 0) {
    store_group(N);
    tma_store_wait_all();
  }
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
  if constexpr (BIAS) {
    // dQ / dK / dV: reduce over the 8 lanes that share q (rows g), one atomic per column and warp
    float v[12] = {csq[0].x * D.scale, csq[0].y * D.scale, csq[1].x * D.scale, csq[1].y * D.scale,
                   csk[0].x * D.scale, csk[0].y * D.scale, csk[1].x * D.scale, csk[1].y * D.scale,
                   csv[0].x, csv[0].y, csv[1].x, csv[1].y};
#pragma unroll
    for (int e = 0; e < 12; ++e) {
      v[e] += __shfl_xor_sync(0xffffffffu, v[e], 4);
      v[e] += __shfl_xor_sync(0xffffffffu, v[e], 8);
      v[e] += __shfl_xor_sync(0xffffffffu, v[e], 16);
    }
    if (g == 0) {
#pragma unroll
      for (int e = 0; e < 12; ++e) {
        const int base = e < 4 ? cq : (e < 8 ? ck : cv);
        atomicAdd(dbias + base + ((e >> 1) & 1) * 8 + 2 * q + (e & 1), v[e]);
      }
    }
    // dE / dG: the whole 64 x 64 tile of this (head, direction) sums into one column each
    float se = 0.f, sg = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][0]));
      const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gt[nt][1]));
      se += (de[nt][0] + de[nt][1]) + (de[nt][2] + de[nt][3]);
      sg += dg[nt][0] * g0.x * (1.f - g0.x) + dg[nt][1] * g0.y * (1.f - g0.y) + dg[nt][2] * g1.x * (1.f - g1.x) +
            dg[nt][3] * g1.y * (1.f - g1.y);
    }
    se = warp_sum(se);
    sg = warp_sum(sg);
    if (lane == 0) {
      if (D.off_e[dir] >= 0) atomicAdd(dbias + D.off_e[dir] + h, se);
      if (D.off_g[dir] >= 0) atomicAdd(dbias + D.off_g[dir] + h, sg);
    }
  }
}


// ------------------------------------------------------------------------------------------------ host
// 4-D map over a [B, N, N, C] 16-bit tensor (row pitch ld elements): box = 64 rows of 16 channels taken along the
// second-to-last index ("row" tiles, fixed first index) or along the first index ("column" tiles, fixed second index)
static int make_tile_map(CUtensorMap *map, const void *base, int B, int N, int C, int64_t ld, bool column, int dtype) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("triplet_attn: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2, (cuuint64_t)N * N * ld * 2};
  const cuuint32_t box[4] = {(cuuint32_t)HD, column ? 1u : (cuuint32_t)TN, column ? (cuuint32_t)TN : 1u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt = dtype == TGT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, 4, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("triplet_attn: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

bool triplet_attn_tma_available() { return encode_tiled_fn() != nullptr; }

template <typename T>
static int fwd_tma_impl(const tgt_triplet_attn_desc &D, const void *proj, void *va, float *stats, const float *ws_e,
                        const __half *ws_g, cudaStream_t st) {
  CUtensorMap mPcol, mProw, mVA;
  const int C = (int)D.ld, Cv = 2 * D.H * HD;
  if (int e = make_tile_map(&mPcol, proj, D.B, D.N, C, D.ld, true, D.dtype)) return e;
  if (int e = make_tile_map(&mProw, proj, D.B, D.N, C, D.ld, false, D.dtype)) return e;
  if (int e = make_tile_map(&mVA, va, D.B, D.N, Cv, Cv, true, D.dtype)) return e;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(tri_attn_fwd_tma<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TF_SMEM);
  });
  KernelTimerScope ts("tri_attn_fwd_tma", st);
  tri_attn_fwd_tma<T><<<dim3(D.H, 2, D.B), 128, TF_SMEM, st>>>(D, mPcol, mProw, mVA, ws_e, ws_g, stats);
  return check_launch("tri_attn_fwd_tma");
}

bool triplet_attn_bwd_tma_has_bias() { return true; }

// The software-pipelined kernel is the only mma.sync backward kept: the two-barrier, key-split (8 warps) and warp-specialised
// variants of round 1 all measured slower (profiles/r1_18_*, r1_20_*) and were removed in round 2 (git history has them).
template <typename T>
static int bwd_tma_impl(const tgt_triplet_attn_desc &D, const void *proj, const void *dva, const float *stats,
                        void *dproj, const float *ws_e, const __half *ws_g, float *ws_de, float *ws_dg, float *dbias,
                        cudaStream_t st) {
  CUtensorMap mPcol, mProw, mDVA, mDPcol, mDProw;
  const int C = (int)D.ld, Cv = 2 * D.H * HD;
  if (int e = make_tile_map(&mPcol, proj, D.B, D.N, C, D.ld, true, D.dtype)) return e;
  if (int e = make_tile_map(&mProw, proj, D.B, D.N, C, D.ld, false, D.dtype)) return e;
  if (int e = make_tile_map(&mDVA, dva, D.B, D.N, Cv, Cv, true, D.dtype)) return e;
  if (int e = make_tile_map(&mDPcol, dproj, D.B, D.N, C, D.ld, true, D.dtype)) return e;
  if (int e = make_tile_map(&mDProw, dproj, D.B, D.N, C, D.ld, false, D.dtype)) return e;
  static std::once_flag oncep;
  std::call_once(oncep, [] {
    cudaFuncSetAttribute(tri_attn_bwd_tma_pipe<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM);
    cudaFuncSetAttribute(tri_attn_bwd_tma_pipe<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TP_SMEM);
  });
  KernelTimerScope ts("tri_attn_bwd_tma", st);
  if (dbias)
    tri_attn_bwd_tma_pipe<T, true><<<dim3(D.H, 2, D.B), 128, TP_SMEM, st>>>(D, mPcol, mProw, mDVA, mDPcol, mDProw, ws_e,
                                                                          ws_g, stats, ws_de, ws_dg, dbias);
  else
    tri_attn_bwd_tma_pipe<T, false><<<dim3(D.H, 2, D.B), 128, TP_SMEM, st>>>(D, mPcol, mProw, mDVA, mDPcol, mDProw, ws_e,
                                                                           ws_g, stats, ws_de, ws_dg, nullptr);
  return check_launch("tri_attn_bwd_tma_pipe");
}

// ------------------------------------------------------------------------------------------------ TripletAggregate (TGT-Agx2)
// Va[b,i,j,dir,h,:] = sum_k A[i,k] V[keyrow(j,k)][h,:]  (triplet.py:61, 68) is the "P V" half of the attention kernel with
// weights A that do not depend on the junction j: one CTA = (head, direction, graph) holds A (64 x 64, 16-bit) as mma
// A-fragments in registers for all N junctions and streams the V tiles through the same TMA ring / stmatrix / TMA store
// path as tri_attn_fwd_tma.  Per junction a CTA moves 2 KB in and 2 KB out and issues 32 mma: purely HBM-bound.
// Backward: dV_j = A^T dO_j (A^T fragments resident) and dA += dO_j V_j^T accumulated over j in registers.
template <typename T>
__device__ __forceinline__ void load_aw_frags(uint32_t (&pa)[4][4], const float *__restrict__ a, int N, int m0, int g, int q,
                                              bool transpose) {
  // pa[t] = A-fragment of rows m0..m0+15, columns 16t..16t+15 of A (or of A^T), zero beyond N
  auto at = [&](int r, int c) -> float {
    if (r >= N || c >= N) return 0.f;
    return transpose ? a[(int64_t)c * N + r] : a[(int64_t)r * N + c];
  };
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int c0 = 16 * t + 2 * q;
    pa[t][0] = Mma<T>::pack(at(m0 + g, c0), at(m0 + g, c0 + 1));
    pa[t][1] = Mma<T>::pack(at(m0 + g + 8, c0), at(m0 + g + 8, c0 + 1));
    pa[t][2] = Mma<T>::pack(at(m0 + g, c0 + 8), at(m0 + g, c0 + 9));
    pa[t][3] = Mma<T>::pack(at(m0 + g + 8, c0 + 8), at(m0 + g + 8, c0 + 9));
  }
}

constexpr int TA_STAGES = 4;
constexpr int TA_SMEM = TA_STAGES * TILE_BYTES + 2 * TILE_BYTES + 64 + 1024;

template <typename T>
__global__ void __launch_bounds__(128, 6)
tri_aggr_fwd_tma(const tgt_triplet_aggr_desc D, const __grid_constant__ CUtensorMap mPcol,
                 const __grid_constant__ CUtensorMap mProw, const __grid_constant__ CUtensorMap mVA,
                 const float *__restrict__ aw) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int m0 = warp * 16;
  const uint32_t sOut = sbase + TA_STAGES * TILE_BYTES;
  const uint32_t bar_full = sOut + 2 * TILE_BYTES;
  if (tid == 0) {
    tma_prefetch_desc(&mPcol);
    tma_prefetch_desc(&mProw);
    tma_prefetch_desc(&mVA);
    for (int s = 0; s < TA_STAGES; ++s) mbar_init(bar_full + s * 8, 1);
    fence_barrier_init();
  }
  uint32_t pa[4][4];
  load_aw_frags<T>(pa, aw + ((int64_t)(b * 2 + dir) * H + h) * N * N, N, m0, g, q, false);
  const int cv = D.off_v[dir] + h * HD, co = dir * H * HD + h * HD;
  auto issue = [&](int j) {                   // thread 0 only
    if (j < N) {
      const uint32_t st = sbase + (j % TA_STAGES) * TILE_BYTES, bar = bar_full + (j % TA_STAGES) * 8;
      mbar_expect_tx(bar, TILE_BYTES);
      if (dir == 0) tma_load_4d(&mProw, bar, st, cv, 0, j, b);
      else tma_load_4d(&mPcol, bar, st, cv, j, 0, b);
    }
  };
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TA_STAGES - 1; ++s) issue(s);
  }
  for (int j = 0; j < N; ++j) {
    if (tid == 0) tma_store_wait_read();            // the output tile written two iterations ago has left smem
    mbar_wait(bar_full + (j % TA_STAGES) * 8, (uint32_t)((j / TA_STAGES) & 1));
    __syncthreads();                                // stage j landed; everyone is done with iteration j-1
    if (tid == 0) {
      issue(j + TA_STAGES - 1);
      if (j > 0) {
        tma_store_4d(&mVA, sOut + ((j - 1) & 1) * TILE_BYTES, co, j - 1, 0, b);
        tma_store_commit();
      }
    }
    const uint32_t sV = sbase + (j % TA_STAGES) * TILE_BYTES;
    float o[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint32_t vb[4];
      load_b_kn(vb, sV, t * 16, lane);
      Mma<T>::run(o[0], pa[t], vb[0], vb[1]);
      Mma<T>::run(o[1], pa[t], vb[2], vb[3]);
    }
    store_c_tile<T>(sOut + (j & 1) * TILE_BYTES, m0, lane, o, 1.f);
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    tma_store_4d(&mVA, sOut + ((N - 1) & 1) * TILE_BYTES, co, N - 1, 0, b);
    tma_store_commit();
    tma_store_wait_all();
  }
}

constexpr int TAB_STAGES = 4;
constexpr int TAB_STAGE_BYTES = 2 * TILE_BYTES;               // V, dO
constexpr int TAB_SMEM = TAB_STAGES * TAB_STAGE_BYTES + 2 * TILE_BYTES + 64 + 1024;

template <typename T>
__global__ void __launch_bounds__(128, 4)
tri_aggr_bwd_tma(const tgt_triplet_aggr_desc D, const __grid_constant__ CUtensorMap mPcol,
                 const __grid_constant__ CUtensorMap mProw, const __grid_constant__ CUtensorMap mDVA,
                 const __grid_constant__ CUtensorMap mDPcol, const __grid_constant__ CUtensorMap mDProw,
                 const float *__restrict__ aw, float *__restrict__ daw) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int N = D.N, H = D.H;
  const int h = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, q = lane & 3;
  const int m0 = warp * 16;
  const uint32_t sOut = sbase + TAB_STAGES * TAB_STAGE_BYTES;
  const uint32_t bar_full = sOut + 2 * TILE_BYTES;
  if (tid == 0) {
    tma_prefetch_desc(&mPcol);
    tma_prefetch_desc(&mProw);
    tma_prefetch_desc(&mDVA);
    tma_prefetch_desc(&mDPcol);
    tma_prefetch_desc(&mDProw);
    for (int s = 0; s < TAB_STAGES; ++s) mbar_init(bar_full + s * 8, 1);
    fence_barrier_init();
  }
  const int64_t abase = ((int64_t)(b * 2 + dir) * H + h) * N * N;
  uint32_t pat[4][4];                          // A^T: rows = keys m0.., columns = queries
  load_aw_frags<T>(pat, aw + abase, N, m0, g, q, true);
  float da[8][4];                              // dA[i = m0 + g (+8)][k]: sum over j of dO_j V_j^T
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) da[nt][0] = da[nt][1] = da[nt][2] = da[nt][3] = 0.f;
  const int cv = D.off_v[dir] + h * HD, co = dir * H * HD + h * HD;
  auto issue = [&](int j) {                   // thread 0 only
    if (j < N) {
      const uint32_t st = sbase + (j % TAB_STAGES) * TAB_STAGE_BYTES, bar = bar_full + (j % TAB_STAGES) * 8;
      mbar_expect_tx(bar, TAB_STAGE_BYTES);
      if (dir == 0) tma_load_4d(&mProw, bar, st, cv, 0, j, b);
      else tma_load_4d(&mPcol, bar, st, cv, j, 0, b);
      tma_load_4d(&mDVA, bar, st + TILE_BYTES, co, j, 0, b);
    }
  };
  auto store_dv = [&](int j) {                // thread 0 only
    if (dir == 0) tma_store_4d(&mDProw, sOut + (j & 1) * TILE_BYTES, cv, 0, j, b);
    else tma_store_4d(&mDPcol, sOut + (j & 1) * TILE_BYTES, cv, j, 0, b);
    tma_store_commit();
  };
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TAB_STAGES - 1; ++s) issue(s);
  }
  for (int j = 0; j < N; ++j) {
    if (tid == 0) tma_store_wait_read();
    mbar_wait(bar_full + (j % TAB_STAGES) * 8, (uint32_t)((j / TAB_STAGES) & 1));
    __syncthreads();
    if (tid == 0) {
      issue(j + TAB_STAGES - 1);
      if (j > 0) store_dv(j - 1);
    }
    const uint32_t sV = sbase + (j % TAB_STAGES) * TAB_STAGE_BYTES, sO = sV + TILE_BYTES;
    // dA += dO_j V_j^T   (rows of this warp x all 64 keys)
    uint32_t oa[4];
    load_a_rows(oa, sO, m0, lane);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      uint32_t vb[4];
      load_b_nk(vb, sV, p * 16, lane);
      Mma<T>::run(da[2 * p], oa, vb[0], vb[1]);
      Mma<T>::run(da[2 * p + 1], oa, vb[2], vb[3]);
    }
    // dV_j = A^T dO_j    (this warp owns keys m0 .. m0+15)
    float dv[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      uint32_t ob[4];
      load_b_kn(ob, sO, t * 16, lane);
      Mma<T>::run(dv[0], pat[t], ob[0], ob[1]);
      Mma<T>::run(dv[1], pat[t], ob[2], ob[3]);
    }
    store_c_tile<T>(sOut + (j & 1) * TILE_BYTES, m0, lane, dv, 1.f);
    fence_proxy_async();
  }
  __syncthreads();
  if (tid == 0) {
    store_dv(N - 1);
    tma_store_wait_all();
  }
  float *o = daw + abase;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int c = nt * 8 + 2 * q;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = m0 + g + (e >> 1) * 8, k = c + (e & 1);
      if (i < N && k < N) o[(int64_t)i * N + k] = da[nt][e];
    }
  }
}

bool triplet_aggr_tma_supported(const tgt_triplet_aggr_desc &D) {
  if (D.dtype != TGT_BF16 && D.dtype != TGT_F16) return false;
  if (D.d != HD || D.N > TN || (D.ld % 8)) return false;
  if ((D.off_v[0] % 8) || (D.off_v[1] % 8)) return false;
  return encode_tiled_fn() != nullptr;
}

template <typename T>
static int aggr_fwd_tma_impl(const tgt_triplet_aggr_desc &D, const void *proj, void *va, const float *aw, cudaStream_t st) {
  CUtensorMap mPcol, mProw, mVA;
  const int C = (int)D.ld, Cv = 2 * D.H * HD;
  if (int e = make_tile_map(&mPcol, proj, D.B, D.N, C, D.ld, true, D.dtype)) return e;
  if (int e = make_tile_map(&mProw, proj, D.B, D.N, C, D.ld, false, D.dtype)) return e;
  if (int e = make_tile_map(&mVA, va, D.B, D.N, Cv, Cv, true, D.dtype)) return e;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(tri_aggr_fwd_tma<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM); });
  KernelTimerScope ts("tri_aggr_fwd_tma", st);
  tri_aggr_fwd_tma<T><<<dim3(D.H, 2, D.B), 128, TA_SMEM, st>>>(D, mPcol, mProw, mVA, aw);
  return check_launch("tri_aggr_fwd_tma");
}

template <typename T>
static int aggr_bwd_tma_impl(const tgt_triplet_aggr_desc &D, const void *proj, const void *dva, const float *aw, float *daw,
                             void *dproj, cudaStream_t st) {
  CUtensorMap mPcol, mProw, mDVA, mDPcol, mDProw;
  const int C = (int)D.ld, Cv = 2 * D.H * HD;
  if (int e = make_tile_map(&mPcol, proj, D.B, D.N, C, D.ld, true, D.dtype)) return e;
  if (int e = make_tile_map(&mProw, proj, D.B, D.N, C, D.ld, false, D.dtype)) return e;
  if (int e = make_tile_map(&mDVA, dva, D.B, D.N, Cv, Cv, true, D.dtype)) return e;
  if (int e = make_tile_map(&mDPcol, dproj, D.B, D.N, C, D.ld, true, D.dtype)) return e;
  if (int e = make_tile_map(&mDProw, dproj, D.B, D.N, C, D.ld, false, D.dtype)) return e;
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(tri_aggr_bwd_tma<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TAB_SMEM); });
  KernelTimerScope ts("tri_aggr_bwd_tma", st);
  tri_aggr_bwd_tma<T><<<dim3(D.H, 2, D.B), 128, TAB_SMEM, st>>>(D, mPcol, mProw, mDVA, mDPcol, mDProw, aw, daw);
  return check_launch("tri_aggr_bwd_tma");
}

int triplet_aggr_fwd_tma_launch(const tgt_triplet_aggr_desc &D, const void *proj, void *va, const float *aw, cudaStream_t st) {
  if (D.dtype == TGT_BF16) return aggr_fwd_tma_impl<__nv_bfloat16>(D, proj, va, aw, st);
  return aggr_fwd_tma_impl<__half>(D, proj, va, aw, st);
}
int triplet_aggr_bwd_tma_launch(const tgt_triplet_aggr_desc &D, const void *proj, const void *dva, const float *aw, float *daw,
                                void *dproj, cudaStream_t st) {
  if (D.dtype == TGT_BF16) return aggr_bwd_tma_impl<__nv_bfloat16>(D, proj, dva, aw, daw, dproj, st);
  return aggr_bwd_tma_impl<__half>(D, proj, dva, aw, daw, dproj, st);
}

int triplet_attn_fwd_tma_launch(const tgt_triplet_attn_desc &D, const void *proj, void *va, float *stats,
                                const float *ws_e, const __half *ws_g, cudaStream_t st) {
  if (D.dtype == TGT_BF16) return fwd_tma_impl<__nv_bfloat16>(D, proj, va, stats, ws_e, ws_g, st);
  return fwd_tma_impl<__half>(D, proj, va, stats, ws_e, ws_g, st);
}

int triplet_attn_bwd_tma_launch(const tgt_triplet_attn_desc &D, const void *proj, const void *dva, const float *stats,
                                void *dproj, const float *ws_e, const __half *ws_g, float *ws_de, float *ws_dg, float *dbias,
                                cudaStream_t st) {
  if (D.dtype == TGT_BF16) return bwd_tma_impl<__nv_bfloat16>(D, proj, dva, stats, dproj, ws_e, ws_g, ws_de, ws_dg, dbias, st);
  return bwd_tma_impl<__half>(D, proj, dva, stats, dproj, ws_e, ws_g, ws_de, ws_dg, dbias, st);
}

}  // namespace tgt
