// Triplet attention core on tcgen05 tensor cores with TMEM accumulators (bf16 / fp16, head dim 16, N <= 64, even H).
//
// Reference math: lib/tgt/layers/triplet.py:213-227 (inward) and :232-246 (outward); index-level restatement in
// SURVEY.md Appendix A.  For one graph b, direction dir and junction atom j every head h runs an N x N x 16 attention
//       S = Q K^T * d^-1/2 + E + mask,  A = softmax_k(S) * sigmoid(G + mask),  Va = A V
// whose bias / gate / mask tile depends on (b, dir, h) only.
//
// CTA = (pair of heads hp, dir, b), 256 threads = two warpgroups; it loops over the junction j:
//   warp 4 (one lane)  TMA producer: per junction six [64 x 16] tiles (32-byte rows, SWIZZLE_32B) -- Q, K, V of both
//                      heads -- into a 5-stage mbarrier ring; rows >= N are out of bounds and arrive as zeros.
//   warp 5 (one lane)  tcgen05.mma issuer, M = 64 instructions.  A 64-row accumulator occupies 16 lanes of each of the
//                      four TMEM sub-partitions, so the two heads of the pair INTERLEAVE in the same columns:
//                      head 0 on lanes 32q + 0..15, head 1 on lanes 32q + 16..31 (q = i / 16):
//                        S_h [64 x 64] = Q_h K_h^T          one UMMA  (M 64, N 64, K 16), both operands K-major
//                        O_h [64 x 16] = P_h V_h            four UMMAs (M 64, N 16, K 16) with A = P_h read from TENSOR
//                                                           MEMORY and B = the V tile as an MN-major operand
//   warps 0-3          softmax warpgroup: thread t <-> TMEM lane t <-> (head (t >> 4) & 1, query i = 16 (t >> 5) + t % 16):
//                      its 64 bias values (log2 domain, mask folded in) and 64 gate values stay in registers for all N
//                      junctions; per junction tcgen05.ld of its S row (64 fp32), row max / exp2 / sum WITHOUT any
//                      shuffle, gate, round to 16 bit, tcgen05.st of P, then tcgen05.ld of the previous junction's O row,
//                      1/l scaling and one 32-byte global store.  setmaxnreg gives this warpgroup 232 registers per thread.
// TMEM (256 columns per CTA, two CTAs per SM), everything double buffered: S 2 x 64 | P 2 x 32 | O 2 x 16.
// Pipeline: the issuer keeps TWO S tiles in flight (S(j+2) is issued as soon as softmax(j) has read its buffer), the
// warpgroup does softmax(j) before epilogue(j-1), so in steady state no role waits for a round trip: S(j) completed one
// junction ago, PV(j-1) ran during softmax(j).  The softmax warps never synchronise with each other (mbarrier arrivals
// only): a variant with the producer / issuer folded into the softmax warps (128-thread CTAs, three per SM, named
// barrier per junction) measured 1.22 ms against 1.01 ms for this one.
//
// The O(N^3 H) tensor never exists; per junction the kernel reads 12 KB and writes 4 KB + 512 B of statistics.
#include <algorithm>
#include "tc_ptx.cuh"
#include "triplet_common.cuh"

namespace tgt {

constexpr int TC_THREADS = 256;
constexpr int TC_REGS_WORK = 232, TC_REGS_CTRL = 24;      // 128 * (232 + 24) = 2 * 128 * 128: both warpgroups start at 128
constexpr int TCF_STAGES = 5;
constexpr int TCF_TILE = TN * HD * 2;                     // 2 KB: one head's 64 x 16 operand tile
constexpr int TCF_STAGE_BYTES = 6 * TCF_TILE;             // Q_h0 Q_h1 K_h0 K_h1 V_h0 V_h1
constexpr int TCF_SMEM = TCF_STAGES * TCF_STAGE_BYTES + 256 + 1024;
// TMEM column map
constexpr uint32_t TCF_COL_S = 0, TCF_COL_P = 128, TCF_COL_O = 192, TCF_TMEM_COLS = 256;
constexpr uint32_t TC_LANE16 = 16u << 16;                 // lane offset of the second interleaved M = 64 tile

// two packed 16-bit values -> float2 (bf16: pure bit moves)
template <typename T> __device__ __forceinline__ float2 unpack16x2(uint32_t w);
template <> __device__ __forceinline__ float2 unpack16x2<__nv_bfloat16>(uint32_t w) {
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}
template <> __device__ __forceinline__ float2 unpack16x2<__half>(uint32_t w) {
  return __half22float2(*reinterpret_cast<const __half2 *>(&w));
}

template <typename T>
__global__ void __launch_bounds__(TC_THREADS, 2)
tri_attn_fwd_tc(const tgt_triplet_attn_desc D, const __grid_constant__ CUtensorMap mC16,
                const __grid_constant__ CUtensorMap mR16, const float *__restrict__ ws_e,
                const __half *__restrict__ ws_g, T *__restrict__ va, float *__restrict__ va_f32,
                float *__restrict__ stats) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int N = D.N, H = D.H;
  const int hp = blockIdx.x, dir = blockIdx.y, b = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t bars = sbase + TCF_STAGES * TCF_STAGE_BYTES;
  const uint32_t bar_full = bars, bar_empty = bars + 8 * TCF_STAGES;
  const uint32_t bar_s = bars + 16 * TCF_STAGES, bar_p = bar_s + 16, bar_o = bar_s + 32;   // two barriers each
  const uint32_t tmem_slot = bar_s + 48;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  if (tid == 0) {
    for (int s = 0; s < TCF_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int q = 0; q < 2; ++q) {
      mbar_init(bar_s + 8 * q, 1);
      mbar_init(bar_p + 8 * q, 4);            // one arrival per softmax warp
      mbar_init(bar_o + 8 * q, 1);
    }
    fence_barrier_init();
  }
  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&mC16);
    tma_prefetch_desc(&mR16);
  }
  if (warp == 5) tc_alloc(tmem_slot, TCF_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp >= 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TC_REGS_CTRL));
    if (warp == 4 && lane == 0) {
      // ------------------------------------------------------------------------------------------ TMA producer
      const int cq = D.off_q[dir] + hp * 2 * HD, ck = D.off_k[dir] + hp * 2 * HD, cv = D.off_v[dir] + hp * 2 * HD;
      for (int j = 0; j < N; ++j) {
        const int s = j % TCF_STAGES;
        mbar_wait(bar_empty + 8 * s, (uint32_t)(((j / TCF_STAGES) & 1) ^ 1));
        const uint32_t st = sbase + s * TCF_STAGE_BYTES, bar = bar_full + 8 * s;
        mbar_expect_tx(bar, TCF_STAGE_BYTES);
        tma_load_4d(&mC16, bar, st, cq, j, 0, b);                       // queries: column j of e, head 2hp
        tma_load_4d(&mC16, bar, st + TCF_TILE, cq + HD, j, 0, b);       //                         head 2hp + 1
        if (dir == 0) {                                                 // keys / values: row j of e
          tma_load_4d(&mR16, bar, st + 2 * TCF_TILE, ck, 0, j, b);
          tma_load_4d(&mR16, bar, st + 3 * TCF_TILE, ck + HD, 0, j, b);
          tma_load_4d(&mR16, bar, st + 4 * TCF_TILE, cv, 0, j, b);
          tma_load_4d(&mR16, bar, st + 5 * TCF_TILE, cv + HD, 0, j, b);
        } else {                                                        // keys / values: column j of e
          tma_load_4d(&mC16, bar, st + 2 * TCF_TILE, ck, j, 0, b);
          tma_load_4d(&mC16, bar, st + 3 * TCF_TILE, ck + HD, j, 0, b);
          tma_load_4d(&mC16, bar, st + 4 * TCF_TILE, cv, j, 0, b);
          tma_load_4d(&mC16, bar, st + 5 * TCF_TILE, cv + HD, j, 0, b);
        }
      }
    } else if (warp == 5 && lane == 0) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      const uint32_t idesc_s = umma_idesc_f16<T>(64, 64, false, false);
      const uint32_t idesc_o = umma_idesc_f16<T>(64, 16, false, true);
      auto issue_s = [&](int j) {                       // S(j) -> buffer j & 1
        const int s = j % TCF_STAGES;
        mbar_wait(bar_full + 8 * s, (uint32_t)((j / TCF_STAGES) & 1));
        tc_fence_after();
        const uint32_t st = sbase + s * TCF_STAGE_BYTES;
        const uint32_t tS = tmem + TCF_COL_S + (uint32_t)(j & 1) * 64u;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)                  // 64 rows x 32 B, 8-row groups 256 B apart
          tc_mma(tS + hh * TC_LANE16, umma_smem_desc(st + hh * TCF_TILE, 256, 16, UMMA_SW32),
                 umma_smem_desc(st + (2 + hh) * TCF_TILE, 256, 16, UMMA_SW32), idesc_s, 0u);
        tc_commit(bar_s + 8 * (j & 1));
      };
      issue_s(0);
      if (N > 1) issue_s(1);
      for (int j = 0; j < N; ++j) {
        mbar_wait(bar_p + 8 * (j & 1), (uint32_t)((j >> 1) & 1));     // P(j) is in tensor memory, S(j) has been read
        tc_fence_after();
        const uint32_t sV = sbase + (j % TCF_STAGES) * TCF_STAGE_BYTES + 4 * TCF_TILE;
        const uint32_t tO = tmem + TCF_COL_O + (uint32_t)(j & 1) * 16u, tP = tmem + TCF_COL_P + (uint32_t)(j & 1) * 32u;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)                // 16 keys per step: 8 packed columns of P, two 8-row groups of V
            tc_mma_ts(tO + hh * TC_LANE16, tP + hh * TC_LANE16 + 8u * ks,
                      umma_smem_desc(sV + hh * TCF_TILE + ks * 512, 256, 16, UMMA_SW32), idesc_o, (uint32_t)(ks != 0));
        tc_commit(bar_o + 8 * (j & 1));                 // O(j) complete
        tc_commit(bar_empty + 8 * (j % TCF_STAGES));    // ... and stage j's tiles are no longer read
        if (j + 2 < N) issue_s(j + 2);                  // its S buffer was read before P(j) was announced
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- softmax warpgroup
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TC_REGS_WORK));
    const int hh = (lane >> 4) & 1, i = warp * 16 + (lane & 15), h = hp * 2 + hh;      // TMEM lane = tid
    const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;
    const uint32_t tS = tmem + lane_sel + TCF_COL_S;
    const uint32_t tP = tmem + lane_sel + TCF_COL_P;
    const uint32_t tO = tmem + lane_sel + TCF_COL_O;

    // bias (log2 domain, mask folded in, -inf on tile padding) and gate of this row, constant over j
    float2 eb[32];
    uint32_t gt[32];
    {
      const int64_t tbase = (((int64_t)(b * 2 + dir) * H + h) * TN + i) * TN;
      const float4 *er = reinterpret_cast<const float4 *>(ws_e + tbase);
      const uint4 *gr = reinterpret_cast<const uint4 *>(ws_g + tbase);
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float4 v = er[c];
        eb[2 * c] = make_float2(v.x, v.y);
        eb[2 * c + 1] = make_float2(v.z, v.w);
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = gr[c];
        gt[4 * c] = v.x; gt[4 * c + 1] = v.y; gt[4 * c + 2] = v.z; gt[4 * c + 3] = v.w;
      }
    }
    // Rows whose N keys are ALL masked (padding atoms as queries): the reference's -3.4e38 mask absorbs the logits and the
    // softmax is uniform over the N keys.  Reproduce exactly: zero logit scale and bias (see fix_fully_masked_rows).
    float c1 = D.scale * LOG2E;
    {
      bool fm = true;
#pragma unroll
      for (int c = 0; c < 32; ++c) fm = fm && (eb[c].x <= -1e37f) && (eb[c].y <= -1e37f);
      if (fm) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          eb[c].x = eb[c].x == -INFINITY ? -INFINITY : 0.f;
          eb[c].y = eb[c].y == -INFINITY ? -INFINITY : 0.f;
        }
        c1 = 0.f;
      }
    }
    const float2 c1p = make_float2(c1, c1);
    const int Cv = 2 * H * HD, co = dir * H * HD + h * HD;
    float inv_prev = 0.f;

    auto epilogue = [&](int jj, float inv_l) {           // O(jj) -> Va[b, i, jj, co .. co+15]
      mbar_wait(bar_o + 8 * (jj & 1), (uint32_t)((jj >> 1) & 1));
      tc_fence_after();
      uint32_t o[16];
      tcx_ld16(tO + (uint32_t)(jj & 1) * 16u, o);
      tc_ld_wait();
      if (i < N) {
        const int64_t r = ((int64_t)(b * N + i) * N + jj) * Cv + co;
        if (va_f32) {
          float4 *dst = reinterpret_cast<float4 *>(va_f32 + r);
#pragma unroll
          for (int c = 0; c < 4; ++c)
            dst[c] = make_float4(__uint_as_float(o[4 * c]) * inv_l, __uint_as_float(o[4 * c + 1]) * inv_l,
                                 __uint_as_float(o[4 * c + 2]) * inv_l, __uint_as_float(o[4 * c + 3]) * inv_l);
        } else {
          uint32_t w[8];
#pragma unroll
          for (int c = 0; c < 8; ++c)
            w[c] = Mma<T>::pack(__uint_as_float(o[2 * c]) * inv_l, __uint_as_float(o[2 * c + 1]) * inv_l);
          uint4 *dst = reinterpret_cast<uint4 *>(va + r);
          dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
          dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
        }
      }
    };

    for (int j = 0; j < N; ++j) {
      mbar_wait(bar_s + 8 * (j & 1), (uint32_t)((j >> 1) & 1));
      tc_fence_after();
      uint32_t s[64];
      tcx_ld32(tS + (uint32_t)(j & 1) * 64u, s);
      tcx_ld32(tS + (uint32_t)(j & 1) * 64u + 32u, s + 32);
      tc_ld_wait();
      // logits in the log2 domain (packed f32x2 arithmetic), row max without shuffles: the thread owns the row
      float2 x[32];
      float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        x[c] = __ffma2_rn(make_float2(__uint_as_float(s[2 * c]), __uint_as_float(s[2 * c + 1])), c1p, eb[c]);
        x[c + 1] = __ffma2_rn(make_float2(__uint_as_float(s[2 * c + 2]), __uint_as_float(s[2 * c + 3])), c1p, eb[c + 1]);
        m0 = fmaxf(m0, x[c].x); m1 = fmaxf(m1, x[c].y); m2 = fmaxf(m2, x[c + 1].x); m3 = fmaxf(m3, x[c + 1].y);
      }
      const float mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      const float2 nm = make_float2(-mx, -mx);
      float2 l2a = make_float2(0.f, 0.f), l2b = make_float2(0.f, 0.f);
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float2 t0 = __fadd2_rn(x[c], nm), t1 = __fadd2_rn(x[c + 1], nm);
        const float2 p0 = make_float2(fast_exp2(t0.x), fast_exp2(t0.y)), p1 = make_float2(fast_exp2(t1.x), fast_exp2(t1.y));
        l2a = __fadd2_rn(l2a, p0);
        l2b = __fadd2_rn(l2b, p1);
        const float2 w0 = __fmul2_rn(p0, __half22float2(*reinterpret_cast<const __half2 *>(&gt[c])));
        const float2 w1 = __fmul2_rn(p1, __half22float2(*reinterpret_cast<const __half2 *>(&gt[c + 1])));
        pk[c] = Mma<T>::pack(w0.x, w0.y);
        pk[c + 1] = Mma<T>::pack(w1.x, w1.y);
      }
      const float l = (l2a.x + l2a.y) + (l2b.x + l2b.y);
      // P(j) -> tensor memory (A operand of the PV UMMAs); PV(j-2), the last reader of this buffer, has completed:
      // its commit was waited for in the epilogue of the previous iteration
      tcx_st32(tP + (uint32_t)(j & 1) * 32u, pk);
      tcx_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_p + 8 * (j & 1));
      if (i < N) stats[(((int64_t)(b * 2 + dir) * H + h) * N + j) * N + i] = mx + __log2f(l);   // P = exp2(x - lse2)
      if (j > 0) epilogue(j - 1, inv_prev);
      inv_prev = 1.f / l;
    }
    epilogue(N - 1, inv_prev);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    tc_dealloc(tmem, TCF_TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
// 4-D map over a [B, N, N, C] 16-bit tensor (row pitch ld elements): box = 64 rows x `ch` channels along the
// second-to-last index ("row" tiles: fixed first index) or along the first index ("column" tiles: fixed second index)
static int make_map(CUtensorMap *map, const void *base, int B, int N, int C, int64_t ld, bool column, int ch, int dtype) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("triplet_attn_tc: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2, (cuuint64_t)N * N * ld * 2};
  const cuuint32_t box[4] = {(cuuint32_t)ch, column ? 1u : (cuuint32_t)TN, column ? (cuuint32_t)TN : 1u, 1u};
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt = dtype == TGT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUtensorMapSwizzle sw = ch == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : (ch == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
  CUresult r = enc(map, dt, 4, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("triplet_attn_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

bool triplet_attn_tc_supported(const tgt_triplet_attn_desc &D) {
  if (D.d != HD || D.N > TN || D.N < 1 || (D.H & 1) || (D.dtype != TGT_BF16 && D.dtype != TGT_F16)) return false;
  if (D.ld % 8) return false;
  for (int dir = 0; dir < 2; ++dir)
    if ((D.off_q[dir] % 8) || (D.off_k[dir] % 8) || (D.off_v[dir] % 8)) return false;
  return encode_tiled_fn() != nullptr;
}

template <typename T>
static int fwd_tc_impl(const tgt_triplet_attn_desc &D, const void *proj, void *va, float *va_f32, float *stats,
                       const float *ws_e, const __half *ws_g, cudaStream_t st) {
  CUtensorMap mC16, mR16;
  const int C = (int)D.ld;
  if (int e = make_map(&mC16, proj, D.B, D.N, C, D.ld, true, 16, D.dtype)) return e;
  if (int e = make_map(&mR16, proj, D.B, D.N, C, D.ld, false, 16, D.dtype)) return e;
  TGT_CUDA_OK(cudaFuncSetAttribute(tri_attn_fwd_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCF_SMEM));
  KernelTimerScope ts("tri_attn_fwd_tc", st);
  tri_attn_fwd_tc<T><<<dim3(D.H / 2, 2, D.B), TC_THREADS, TCF_SMEM, st>>>(D, mC16, mR16, ws_e, ws_g, (T *)va,
                                                                          va_f32, stats);
  return check_launch("tri_attn_fwd_tc");
}

int triplet_attn_fwd_tc_launch(const tgt_triplet_attn_desc &D, const void *proj, void *va, float *va_f32, float *stats,
                               const float *ws_e, const __half *ws_g, cudaStream_t st) {
  if (D.dtype == TGT_BF16) return fwd_tc_impl<__nv_bfloat16>(D, proj, va, va_f32, stats, ws_e, ws_g, st);
  return fwd_tc_impl<__half>(D, proj, va, va_f32, stats, ws_e, ws_g, st);
}


// ================================================================================================ backward
// Backward of the gated attention core on tcgen05 / TMEM (autograd of triplet.py:213-227, 232-246):
//      S = Q K^T, dA = dO V^T                      (recomputed / computed per junction on the tensor cores)
//      P = exp2(S c + E - lse),  g = gate,  A = P g,  dP = dA g,  delta = rowsum(dP P),  dS = P (dP - delta)
//      dQ = c dS K,   dK = c dS^T Q,   dV = A^T dO,   dE = sum_j dS,   dG = g (1 - g) sum_j dA P
//
// One PERSISTENT CTA per SM, 384 threads = two compute warpgroups + one control warpgroup.  Each compute warpgroup owns a
// stream of work items (head h, direction, graph b) and loops over the junction j; it never synchronises with the other
// warpgroup, nor its four warps with each other (mbarrier arrivals only).
//   thread <-> (query row i = 16 q + lane % 16, key half kh = lane / 16) with q = warp % 4: the M = 64 accumulators of
//   S and dA are issued as TWO N = 32 UMMAs (keys 0-31 -> lanes 32q + 0..15, keys 32-63 -> lanes 32q + 16..31, the
//   interleaved half of each TMEM sub-partition), so a thread tcgen05.ld's 32 S and 32 dA values of ONE row, the row
//   reduction (delta) is one shuffle, and dE / dG accumulate in 2 x 32 registers for all N junctions (no atomics).
//   dS and A leave as 16-bit rows of two 64 x 64 shared-memory tiles (128-byte rows, 128B swizzle) that the tensor core
//   reads BOTH ways: K-major for dQ = dS K, MN-major (transposed, no data movement) for dK = dS^T Q and dV = A^T dO;
//   the Q / K / dO tiles TMA delivered are their MN-major B operands.  dQ | dV (lanes 0-15) and dK (lanes 16-31) come
//   back with one tcgen05.ld and leave with 32-byte global stores.
//   control warpgroup: warps 8, 9 = TMA producers (Q, K, V, dO tiles of a junction, 4-stage ring per compute warpgroup),
//   warps 10, 11 = tcgen05.mma issuers (4 + 12 UMMAs per junction).
// TMEM (512 columns = the SM): per warpgroup S|dA 2 x 64 (double buffered) + dQ/dK|dV 2 x 32.
// setmaxnreg: 240 registers per compute thread, 24 per control thread.
constexpr int TCB_THREADS = 384;
constexpr int TCB_REGS_WORK = 240, TCB_REGS_CTRL = 24;      // 2 * 128 * 240 + 128 * 24 = 64512
constexpr int TCB_STAGES = 4;
constexpr int TCB_STAGE_BYTES = 4 * TCF_TILE;               // Q, K, V, dO
constexpr int TCB_XT = TN * TN * 2;                         // 8 KB: one 64 x 64 16-bit tile (dS or A)
constexpr int TCB_GATE = 128 * 64;                          // 8 KB: 32 fp16 gate values per thread, [chunk][thread] x 16 B
constexpr int TCB_WG_BYTES = TCB_STAGES * TCB_STAGE_BYTES + 4 * TCB_XT + TCB_GATE;      // 72 KB
constexpr int TCB_SMEM = 2 * TCB_WG_BYTES + 512 + 1024;
constexpr uint32_t TCB_COLS_WG = 192, TCB_COL_OUT = 128, TCB_TMEM_COLS = 512;

template <typename T>
__global__ void __launch_bounds__(TCB_THREADS, 1)
tri_attn_bwd_tc(const tgt_triplet_attn_desc D, const __grid_constant__ CUtensorMap mPcol,
                const __grid_constant__ CUtensorMap mProw, const __grid_constant__ CUtensorMap mDVA,
                const float *__restrict__ ws_e, const __half *__restrict__ ws_g, const float *__restrict__ stats,
                T *__restrict__ dproj, float *__restrict__ ws_de, float *__restrict__ ws_dg) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int N = D.N, H = D.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int total_items = D.B * 2 * H;
  const uint32_t bars = sbase + 2 * TCB_WG_BYTES;
  const uint32_t tmem_slot = bars + 256;
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  // per-warpgroup barrier block (w = 0, 1): full[4] empty[4] s_ready[2] t_ready[2] o_ready[2]
  auto bar_full = [&](int w, int s) { return bars + w * 128 + 8 * s; };
  auto bar_empty = [&](int w, int s) { return bars + w * 128 + 32 + 8 * s; };
  auto bar_s = [&](int w, int q) { return bars + w * 128 + 64 + 8 * q; };
  auto bar_t = [&](int w, int q) { return bars + w * 128 + 80 + 8 * q; };
  auto bar_o = [&](int w, int q) { return bars + w * 128 + 96 + 8 * q; };

  if (tid == 0) {
    for (int w = 0; w < 2; ++w) {
      for (int s = 0; s < TCB_STAGES; ++s) {
        mbar_init(bar_full(w, s), 1);
        mbar_init(bar_empty(w, s), 1);
      }
      for (int q = 0; q < 2; ++q) {
        mbar_init(bar_s(w, q), 1);
        mbar_init(bar_t(w, q), 4);             // one arrival per compute warp
        mbar_init(bar_o(w, q), 1);
      }
    }
    fence_barrier_init();
    tma_prefetch_desc(&mPcol);
    tma_prefetch_desc(&mProw);
    tma_prefetch_desc(&mDVA);
  }
  if (warp == 10) tc_alloc(tmem_slot, TCB_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;

  if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TCB_REGS_CTRL));
    const int w = warp & 1;                                  // the compute warpgroup this control warp serves
    const uint32_t wg_smem = sbase + w * TCB_WG_BYTES;
    const int first = blockIdx.x * 2 + w, stride = gridDim.x * 2;
    if (warp < 10 && lane == 0) {
      // ------------------------------------------------------------------------------------------ TMA producer
      int n = 0;
      for (int t = first; t < total_items; t += stride) {
        const int b = t / (2 * H), r = t - b * 2 * H, dir = r / H, h = r - dir * H;
        const int cq = D.off_q[dir] + h * HD, ck = D.off_k[dir] + h * HD, cv = D.off_v[dir] + h * HD;
        const int co = dir * H * HD + h * HD;
        for (int j = 0; j < N; ++j, ++n) {
          const int s = n % TCB_STAGES;
          mbar_wait(bar_empty(w, s), (uint32_t)(((n / TCB_STAGES) & 1) ^ 1));
          const uint32_t st = wg_smem + s * TCB_STAGE_BYTES, bar = bar_full(w, s);
          mbar_expect_tx(bar, TCB_STAGE_BYTES);
          tma_load_4d(&mPcol, bar, st, cq, j, 0, b);
          if (dir == 0) {
            tma_load_4d(&mProw, bar, st + TCF_TILE, ck, 0, j, b);
            tma_load_4d(&mProw, bar, st + 2 * TCF_TILE, cv, 0, j, b);
          } else {
            tma_load_4d(&mPcol, bar, st + TCF_TILE, ck, j, 0, b);
            tma_load_4d(&mPcol, bar, st + 2 * TCF_TILE, cv, j, 0, b);
          }
          tma_load_4d(&mDVA, bar, st + 3 * TCF_TILE, co, j, 0, b);
        }
      }
    } else if (warp >= 10 && lane == 0) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      int items = 0;
      for (int t = first; t < total_items; t += stride) ++items;
      const int total_n = items * N;
      const uint32_t tw = tmem + (uint32_t)w * TCB_COLS_WG;
      const uint32_t sX = wg_smem + TCB_STAGES * TCB_STAGE_BYTES;          // [buf][dS | A]
      const uint32_t id_sa = umma_idesc_f16<T>(64, 32, false, false);
      const uint32_t id_dq = umma_idesc_f16<T>(64, 16, false, true);
      const uint32_t id_kv = umma_idesc_f16<T>(64, 16, true, true);
      auto issue_sa = [&](int n) {                       // S(n), dA(n) -> buffer n & 1
        const int s = n % TCB_STAGES;
        mbar_wait(bar_full(w, s), (uint32_t)((n / TCB_STAGES) & 1));
        tc_fence_after();
        const uint32_t st = wg_smem + s * TCB_STAGE_BYTES;
        const uint32_t tSA = tw + (uint32_t)(n & 1) * 64u;
        const uint64_t dQ_ = umma_smem_desc(st, 256, 16, UMMA_SW32), dO_ = umma_smem_desc(st + 3 * TCF_TILE, 256, 16, UMMA_SW32);
#pragma unroll
        for (int kh = 0; kh < 2; ++kh) {                 // keys 32 kh .. 32 kh + 31 -> lane offset 16 kh
          tc_mma(tSA + kh * TC_LANE16, dQ_, umma_smem_desc(st + TCF_TILE + kh * 1024, 256, 16, UMMA_SW32), id_sa, 0u);
          tc_mma(tSA + 32u + kh * TC_LANE16, dO_, umma_smem_desc(st + 2 * TCF_TILE + kh * 1024, 256, 16, UMMA_SW32), id_sa, 0u);
        }
        tc_commit(bar_s(w, n & 1));
      };
      if (total_n > 0) issue_sa(0);
      if (total_n > 1) issue_sa(1);
      for (int n = 0; n < total_n; ++n) {
        mbar_wait(bar_t(w, n & 1), (uint32_t)((n >> 1) & 1));      // dS(n), A(n) are in shared memory; S(n), dA(n) were read
        tc_fence_after();
        const uint32_t st = wg_smem + (n % TCB_STAGES) * TCB_STAGE_BYTES;
        const uint32_t sDS = sX + (uint32_t)(n & 1) * 2 * TCB_XT, sA = sDS + TCB_XT;
        const uint32_t tOut = tw + TCB_COL_OUT + (uint32_t)(n & 1) * 32u;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          // dQ[i, d] += dS[i, 16ks..] K[16ks.., d]        A: dS K-major (32 B per step), B: K tile MN-major
          tc_mma(tOut, umma_smem_desc(sDS + ks * 32, 1024, 16, UMMA_SW128), umma_smem_desc(st + TCF_TILE + ks * 512, 256, 16, UMMA_SW32),
                 id_dq, (uint32_t)(ks != 0));
          // dK[k, d] += dS[16ks.., k]^T Q[16ks.., d]      A: dS MN-major (two 8-row groups per step), B: Q tile MN-major
          tc_mma(tOut + TC_LANE16, umma_smem_desc(sDS + ks * 2048, 1024, 16, UMMA_SW128), umma_smem_desc(st + ks * 512, 256, 16, UMMA_SW32),
                 id_kv, (uint32_t)(ks != 0));
          // dV[k, d] += A[16ks.., k]^T dO[16ks.., d]
          tc_mma(tOut + 16u, umma_smem_desc(sA + ks * 2048, 1024, 16, UMMA_SW128),
                 umma_smem_desc(st + 3 * TCF_TILE + ks * 512, 256, 16, UMMA_SW32), id_kv, (uint32_t)(ks != 0));
        }
        tc_commit(bar_o(w, n & 1));
        tc_commit(bar_empty(w, n % TCB_STAGES));
        if (n + 2 < total_n) issue_sa(n + 2);
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- compute warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TCB_REGS_WORK));
    const int w = warp >> 2, q = warp & 3, tw_id = tid & 127;
    const int i = q * 16 + (lane & 15), kh = lane >> 4;
    const uint32_t wg_smem = sbase + w * TCB_WG_BYTES;
    const uint32_t sX = wg_smem + TCB_STAGES * TCB_STAGE_BYTES;
    const uint32_t sGate = sX + 4 * TCB_XT + (uint32_t)tw_id * 16u;                    // + chunk * 2048
    const uint32_t tl = tmem + (uint32_t)w * TCB_COLS_WG + ((uint32_t)(q * 32) << 16);
    const uint32_t row_off = (uint32_t)i * 128u;                                        // this row in a 64 x 64 tile
    const int first = blockIdx.x * 2 + w, stride = gridDim.x * 2;
    constexpr float LN2 = 0.6931471805599453f;
    const float2 l2e = make_float2(LOG2E, LOG2E);

    uint32_t ebk[16];
    float2 dE[16], dG[16];
    float sc = 0.f;
    int b = 0, dir = 0, h = 0;
    auto setup = [&](int t) {
      b = t / (2 * H);
      const int r = t - b * 2 * H;
      dir = r / H;
      h = r - dir * H;
      const int64_t tbase = (((int64_t)(b * 2 + dir) * H + h) * TN + i) * TN + 32 * kh;
      const float4 *er = reinterpret_cast<const float4 *>(ws_e + tbase);
      float4 v[8];
      bool fm = true;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        v[c] = er[c];
        fm = fm && v[c].x <= -1e37f && v[c].y <= -1e37f && v[c].z <= -1e37f && v[c].w <= -1e37f;
      }
      fm = __shfl_xor_sync(0xffffffffu, (int)fm, 16) && fm;            // the other half of the row
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        float4 u = v[c];
        if (fm) {                                                       // fully masked row: uniform softmax (see forward)
          u.x = u.x == -INFINITY ? -INFINITY : 0.f; u.y = u.y == -INFINITY ? -INFINITY : 0.f;
          u.z = u.z == -INFINITY ? -INFINITY : 0.f; u.w = u.w == -INFINITY ? -INFINITY : 0.f;
        }
        ebk[2 * c] = Mma<T>::pack(u.x * LN2, u.y * LN2);
        ebk[2 * c + 1] = Mma<T>::pack(u.z * LN2, u.w * LN2);
      }
      sc = fm ? 0.f : D.scale;
      const uint4 *gr = reinterpret_cast<const uint4 *>(ws_g + tbase);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 g4 = gr[c];
        st_shared_v4u(sGate + c * 2048, g4.x, g4.y, g4.z, g4.w);
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) dE[c] = dG[c] = make_float2(0.f, 0.f);
    };
    auto finish = [&]() {                                               // dE, dG = g (1 - g) * acc -> [B,2,H,64,64] tiles
      const int64_t tbase = (((int64_t)(b * 2 + dir) * H + h) * TN + i) * TN + 32 * kh;
      float4 *de = reinterpret_cast<float4 *>(ws_de + tbase), *dg = reinterpret_cast<float4 *>(ws_dg + tbase);
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint4 g4 = ld_shared_v4u(sGate + c4 * 2048);
        const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c = 4 * c4 + 2 * u;
          const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gw[2 * u]));
          const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gw[2 * u + 1]));
          de[2 * c4 + u] = make_float4(dE[c].x, dE[c].y, dE[c + 1].x, dE[c + 1].y);
          dg[2 * c4 + u] = make_float4(dG[c].x * g0.x * (1.f - g0.x), dG[c].y * g0.y * (1.f - g0.y),
                                       dG[c + 1].x * g1.x * (1.f - g1.x), dG[c + 1].y * g1.y * (1.f - g1.y));
        }
      }
    };
    // O(n-1) of the junction (pb, pdir, ph, pj): dQ | dV rows on lanes 0-15, dK rows on lanes 16-31
    int pb = 0, pdir = 0, ph = 0, pj = 0;
    auto epilogue = [&](int n) {
      mbar_wait(bar_o(w, n & 1), (uint32_t)((n >> 1) & 1));
      tc_fence_after();
      uint32_t o[32];
      tcx_ld32(tl + TCB_COL_OUT + (uint32_t)(n & 1) * 32u, o);
      tc_ld_wait();
      if (i < N) {
        const int64_t qrow = ((int64_t)(pb * N + i) * N + pj);
        const int64_t krow = pdir == 0 ? ((int64_t)(pb * N + pj) * N + i) : qrow;      // i plays the key index here
        const float s_ = D.scale;
        uint32_t wv[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) wv[c] = Mma<T>::pack(__uint_as_float(o[2 * c]) * s_, __uint_as_float(o[2 * c + 1]) * s_);
        if (kh == 0) {
          uint4 *dq = reinterpret_cast<uint4 *>(dproj + qrow * D.ld + D.off_q[pdir] + ph * HD);
          dq[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          dq[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
          uint32_t vv[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) vv[c] = Mma<T>::pack(__uint_as_float(o[16 + 2 * c]), __uint_as_float(o[17 + 2 * c]));
          uint4 *dv = reinterpret_cast<uint4 *>(dproj + krow * D.ld + D.off_v[pdir] + ph * HD);
          dv[0] = make_uint4(vv[0], vv[1], vv[2], vv[3]);
          dv[1] = make_uint4(vv[4], vv[5], vv[6], vv[7]);
        } else {
          uint4 *dk = reinterpret_cast<uint4 *>(dproj + krow * D.ld + D.off_k[pdir] + ph * HD);
          dk[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          dk[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
        }
      }
    };

    int t = first, j = 0;
    if (t < total_items) setup(t);
    float lse_next = 0.f;
    auto lse_addr = [&](int jj) { return stats + (((int64_t)(b * 2 + dir) * H + h) * N + jj) * N + i; };
    if (t < total_items) lse_next = i < N ? __ldg(lse_addr(0)) : INFINITY;
    for (int n = 0; t < total_items; ++n) {
      const float lse2 = lse_next;                       // log2-domain log-sum-exp of row i at junction j (+inf: padding row)
      if (j + 1 < N) lse_next = i < N ? __ldg(lse_addr(j + 1)) : INFINITY;
      mbar_wait(bar_s(w, n & 1), (uint32_t)((n >> 1) & 1));
      tc_fence_after();
      uint32_t sv[32], av[32];
      tcx_ld32(tl + (uint32_t)(n & 1) * 64u, sv);
      tcx_ld32(tl + (uint32_t)(n & 1) * 64u + 32u, av);
      tc_ld_wait();
      // pass 1: P, dG += dA P, A = P g, dP = dA g, delta = sum dP P
      const float2 scp = make_float2(sc, sc), nl = make_float2(-lse2, -lse2);
      float2 p[16], dp[16], dl = make_float2(0.f, 0.f);
      uint32_t apk[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const uint4 g4 = ld_shared_v4u(sGate + c4 * 2048);
        const uint32_t gw[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int c = 4 * c4 + u;
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * c]), __uint_as_float(sv[2 * c + 1])), scp,
                                      unpack16x2<T>(ebk[c]));
          const float2 e2 = __ffma2_rn(x, l2e, nl);
          p[c] = make_float2(fast_exp2(e2.x), fast_exp2(e2.y));
          const float2 da = make_float2(__uint_as_float(av[2 * c]), __uint_as_float(av[2 * c + 1]));
          const float2 g2 = __half22float2(*reinterpret_cast<const __half2 *>(&gw[u]));
          dG[c] = __ffma2_rn(da, p[c], dG[c]);
          dp[c] = __fmul2_rn(da, g2);
          dl = __ffma2_rn(dp[c], p[c], dl);
          const float2 a2 = __fmul2_rn(p[c], g2);
          apk[c] = Mma<T>::pack(a2.x, a2.y);
        }
      }
      float delta = dl.x + dl.y;
      delta += __shfl_xor_sync(0xffffffffu, delta, 16);
      // pass 2: dS = P (dP - delta), dE += dS
      const float2 nd = make_float2(-delta, -delta);
      uint32_t dpk[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        const float2 ds = __fmul2_rn(p[c], __fadd2_rn(dp[c], nd));
        dE[c] = __fadd2_rn(dE[c], ds);
        dpk[c] = Mma<T>::pack(ds.x, ds.y);
      }
      // this thread's 64 bytes of row i of the dS and A tiles (chunks 4 kh .. 4 kh + 3, 128B swizzle).  The tiles of
      // junction n-2 were last read by UMMAs whose completion the epilogue of the previous iteration waited for.
      {
        const uint32_t sDS = sX + (uint32_t)(n & 1) * 2 * TCB_XT + row_off, sA = sDS + TCB_XT;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = (uint32_t)(((4 * kh + c) ^ (i & 7)) << 4);
          st_shared_v4u(sDS + off, dpk[4 * c], dpk[4 * c + 1], dpk[4 * c + 2], dpk[4 * c + 3]);
          st_shared_v4u(sA + off, apk[4 * c], apk[4 * c + 1], apk[4 * c + 2], apk[4 * c + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t(w, n & 1));
      if (n > 0) epilogue(n - 1);
      pb = b; pdir = dir; ph = h; pj = j;
      if (++j == N) {
        finish();
        t += stride;
        j = 0;
        if (t < total_items) {
          setup(t);
          lse_next = i < N ? __ldg(lse_addr(0)) : INFINITY;
        }
      }
      if (t >= total_items) epilogue(n);                 // the very last junction of this warpgroup
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 10) {
    tc_fence_after();
    tc_dealloc(tmem, TCB_TMEM_COLS);
  }
}

template <typename T>
static int bwd_tc_impl(const tgt_triplet_attn_desc &D, const void *proj, const void *dva, const float *stats, void *dproj,
                       const float *ws_e, const __half *ws_g, float *ws_de, float *ws_dg, cudaStream_t st) {
  CUtensorMap mPcol, mProw, mDVA;
  const int C = (int)D.ld, Cv = 2 * D.H * HD;
  if (int e = make_map(&mPcol, proj, D.B, D.N, C, D.ld, true, 16, D.dtype)) return e;
  if (int e = make_map(&mProw, proj, D.B, D.N, C, D.ld, false, 16, D.dtype)) return e;
  if (int e = make_map(&mDVA, dva, D.B, D.N, Cv, Cv, true, 16, D.dtype)) return e;
  int dev = 0, sms = 0;
  TGT_CUDA_OK(cudaGetDevice(&dev));
  TGT_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TGT_CUDA_OK(cudaFuncSetAttribute(tri_attn_bwd_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM));
  const int items = D.B * 2 * D.H;
  const int grid = std::min(sms, (items + 1) / 2);
  KernelTimerScope ts("tri_attn_bwd_tc", st);
  tri_attn_bwd_tc<T><<<grid, TCB_THREADS, TCB_SMEM, st>>>(D, mPcol, mProw, mDVA, ws_e, ws_g, stats, (T *)dproj, ws_de, ws_dg);
  return check_launch("tri_attn_bwd_tc");
}

int triplet_attn_bwd_tc_launch(const tgt_triplet_attn_desc &D, const void *proj, const void *dva, const float *stats,
                               void *dproj, const float *ws_e, const __half *ws_g, float *ws_de, float *ws_dg,
                               cudaStream_t st) {
  if (D.dtype == TGT_BF16) return bwd_tc_impl<__nv_bfloat16>(D, proj, dva, stats, dproj, ws_e, ws_g, ws_de, ws_dg, st);
  return bwd_tc_impl<__half>(D, proj, dva, stats, dproj, ws_e, ws_g, ws_de, ws_dg, st);
}

}  // namespace tgt
