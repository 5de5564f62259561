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
  // (warp index through a shuffle: provably warp-uniform, so the single-lane roles below keep their operands in uniform
  //  registers and issue TMA / UMMA / commit without a vote loop around every instruction -- see gemm_tc.cu)
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
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
    if (warp == 4) {
      // ------------------------------------------------------------------------------------------ TMA producer
      // (the whole warp walks the loop, one elected lane issues)
      const int cq = D.off_q[dir] + hp * 2 * HD, ck = D.off_k[dir] + hp * 2 * HD, cv = D.off_v[dir] + hp * 2 * HD;
      for (int j = 0; j < N; ++j) {
        const int s = j % TCF_STAGES;
        mbar_wait(bar_empty + 8 * s, (uint32_t)(((j / TCF_STAGES) & 1) ^ 1));
        const uint32_t st = sbase + s * TCF_STAGE_BYTES, bar = bar_full + 8 * s;
        if (elect_one()) {
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
        __syncwarp();
      }
    } else if (warp == 5) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      const uint32_t idesc_s = umma_idesc_f16<T>(64, 64, false, false);
      const uint32_t idesc_o = umma_idesc_f16<T>(64, 16, false, true);
      auto issue_s = [&](int j) {                       // S(j) -> buffer j & 1
        const int s = j % TCF_STAGES;
        mbar_wait(bar_full + 8 * s, (uint32_t)((j / TCF_STAGES) & 1));
        tc_fence_after();
        const uint32_t st = sbase + s * TCF_STAGE_BYTES;
        const uint32_t tS = tmem + TCF_COL_S + (uint32_t)(j & 1) * 64u;
        if (elect_one()) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh)                  // 64 rows x 32 B, 8-row groups 256 B apart
            tc_mma(tS + hh * TC_LANE16, umma_smem_desc(st + hh * TCF_TILE, 256, 16, UMMA_SW32),
                   umma_smem_desc(st + (2 + hh) * TCF_TILE, 256, 16, UMMA_SW32), idesc_s, 0u);
          tc_commit(bar_s + 8 * (j & 1));
        }
        __syncwarp();
      };
      issue_s(0);
      if (N > 1) issue_s(1);
      for (int j = 0; j < N; ++j) {
        mbar_wait(bar_p + 8 * (j & 1), (uint32_t)((j >> 1) & 1));     // P(j) is in tensor memory, S(j) has been read
        tc_fence_after();
        const uint32_t sV = sbase + (j % TCF_STAGES) * TCF_STAGE_BYTES + 4 * TCF_TILE;
        const uint32_t tO = tmem + TCF_COL_O + (uint32_t)(j & 1) * 16u, tP = tmem + TCF_COL_P + (uint32_t)(j & 1) * 32u;
        if (elect_one()) {
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const uint64_t vd = umma_smem_desc(sV + hh * TCF_TILE, 256, 16, UMMA_SW32);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)              // 16 keys per step: 8 packed columns of P, two 8-row groups of V (+512 B)
              tc_mma_ts(tO + hh * TC_LANE16, tP + hh * TC_LANE16 + 8u * ks, vd + (uint64_t)(ks * 32), idesc_o, (uint32_t)(ks != 0));
          }
          tc_commit(bar_o + 8 * (j & 1));               // O(j) complete
          tc_commit(bar_empty + 8 * (j % TCF_STAGES));  // ... and stage j's tiles are no longer read
        }
        __syncwarp();
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
static int make_map(CUtensorMap *map, const void *base, int B, int N, int C, int64_t ld, bool column, int ch, int dtype,
                    int rows = TN) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("triplet_attn_tc: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)N, (cuuint64_t)B};
  const cuuint64_t gstride[3] = {(cuuint64_t)ld * 2, (cuuint64_t)N * ld * 2, (cuuint64_t)N * N * ld * 2};
  const cuuint32_t box[4] = {(cuuint32_t)ch, column ? 1u : (cuuint32_t)rows, column ? (cuuint32_t)rows : 1u, 1u};
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
// One PERSISTENT CTA per SM, 512 threads = four warpgroups with four roles:
//   warps 0-3, 4-7   two COMPUTE warpgroups.  Each owns a stream of work items (head h, direction, graph b), drawn in
//                    id order from a global counter, and loops over the junction j.  thread <-> (query row
//                    i = 16 q + lane % 16, key half kh = lane / 16), q = warp % 4: the M = 64 accumulators of S and dA
//                    are issued as TWO N = 32 UMMAs (keys 0-31 -> TMEM lanes 32q + 0..15, keys 32-63 -> lanes
//                    32q + 16..31, the interleaved half of each sub-partition), so a thread tcgen05.ld's 32 S and 32
//                    dA values of ONE row, the row reduction (delta) is one shuffle, dE / dG accumulate in 2 x 32
//                    registers for all N junctions (no atomics), bias and gate rows stay packed in 2 x 16 registers.
//                    delta = rowsum(dP P) is taken as dO . O from the dO / Va tiles (one pass over the scores).  dS and A leave as 16-bit rows of two 64 x 64 shared-memory tiles (128-byte
//                    rows, 128B swizzle) that the tensor core reads BOTH ways: K-major for dQ = dS K, MN-major
//                    (transposed, no data movement) for dK = dS^T Q and dV = A^T dO; the Q / K / dO tiles TMA
//                    delivered are their MN-major B operands.  The compute warps never wait for each other.
//   warps 8-11       CONTROL: 8, 9 = TMA producers (Q, K, V, dO tiles, 4-stage ring per compute warpgroup; they also draw
//                    the work items), 10, 11 = tcgen05.mma issuers (4 + 12 UMMAs per junction).
//   warps 12-15      EPILOGUE: warp 12 + q drains TMEM lane quadrant q of BOTH compute warpgroups' output accumulators
//                    (dQ | dV rows on lanes 0-15, dK rows on lanes 16-31), scales, rounds and stores them, so neither the
//                    TMEM round trip nor the global stores sit in the compute warps' instruction streams.
// TMEM (512 columns = the SM): per compute warpgroup S|dA 2 x 64 (double buffered) + dQ/dK|dV 2 x 32.
// setmaxnreg: 208 registers per compute thread, 40 per control thread, 56 per epilogue thread.
constexpr int TCB_THREADS = 512;
constexpr int TCB_REGS_WORK = 208, TCB_REGS_CTRL = 40, TCB_REGS_EPI = 56;     // 128 * (2 * 208 + 40 + 56) = 65536
constexpr int TCB_STAGES = 4;
constexpr int TCB_STAGE_BYTES = 5 * TCF_TILE;               // Q, K, V, dO, O (= the forward output Va, for delta)
constexpr int TCB_XT = TN * TN * 2;                         // 8 KB: one 64 x 64 16-bit tile (dS or A)
constexpr int TCB_WG_BYTES = TCB_STAGES * TCB_STAGE_BYTES + 4 * TCB_XT;      // 72 KB
constexpr int TCB_QS = 16;                                  // item-queue slots per warpgroup
constexpr int TCB_SMEM = 2 * TCB_WG_BYTES + 1024 + 1024;
constexpr uint32_t TCB_COLS_WG = 192, TCB_COL_OUT = 128, TCB_TMEM_COLS = 512;

// shared-memory matrix descriptor from (address >> 4) + a byte offset: the high word is a per-layout constant, so the
// issuing thread spends one add per descriptor
__host__ __device__ constexpr uint32_t umma_desc_hi(uint32_t sbo_bytes, uint32_t swizzle) { return (sbo_bytes >> 4) | (1u << 14) | (swizzle << 29); }
__device__ __forceinline__ uint64_t umma_desc_at(uint32_t base16, uint32_t off_bytes, uint32_t hi) {
  return ((uint64_t)hi << 32) | (uint64_t)(base16 + (off_bytes >> 4) + (1u << 16));
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// wait of a single-lane control role: back off between polls so that the spin does not eat the issue slots of the
// compute warps on the same SM sub-partition
__device__ __forceinline__ void mbar_wait_polite(uint32_t bar, uint32_t parity) {
  while (!mbar_test(bar, parity)) __nanosleep(40);
}

template <typename T>
__global__ void __launch_bounds__(TCB_THREADS, 1)
tri_attn_bwd_tc(const tgt_triplet_attn_desc D, const __grid_constant__ CUtensorMap mPcol,
                const __grid_constant__ CUtensorMap mProw, const __grid_constant__ CUtensorMap mDVA,
                const __grid_constant__ CUtensorMap mVA,
                const float *__restrict__ ws_e, const __half *__restrict__ ws_g, const float *__restrict__ stats,
                T *__restrict__ dproj, float *__restrict__ ws_de, float *__restrict__ ws_dg, int *__restrict__ item_counter) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int N = D.N, H = D.H;
  const int tid = threadIdx.x, lane = tid & 31, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform
  const int total_items = D.B * 2 * H;
  const uint32_t bars = sbase + 2 * TCB_WG_BYTES;
  const uint32_t tmem_slot = bars + 288;
  auto generic = [&](uint32_t saddr) { return smem_raw + (saddr - smem_u32(smem_raw)); };
  volatile uint32_t *tmem_slot_ptr = reinterpret_cast<volatile uint32_t *>(generic(tmem_slot));
  // per-warpgroup barrier block (w = 0, 1): full[4] empty[4] s_ready[2] t_ready[2] o_ready[2] o_free[2]
  auto bar_full = [&](int w, int s) { return bars + w * 128 + 8 * s; };
  auto bar_empty = [&](int w, int s) { return bars + w * 128 + 32 + 8 * s; };
  auto bar_s = [&](int w, int q) { return bars + w * 128 + 64 + 8 * q; };
  auto bar_t = [&](int w, int q) { return bars + w * 128 + 80 + 8 * q; };
  auto bar_o = [&](int w, int q) { return bars + w * 128 + 96 + 8 * q; };
  auto bar_of = [&](int w, int q) { return bars + w * 128 + 112 + 8 * q; };
  // item queue: the producer of a warpgroup draws work items (h, dir, b) from a global counter, in order, and hands
  // their ids to the issuer, the compute warps and the epilogue warps.  Items are therefore always started in id order by
  // whichever warpgroup is free -- the 16 heads x 2 directions of a graph, which share 128-byte lines of the projection,
  // run at the same time on neighbouring SMs and meet in L2 (a static round-robin lets them drift apart: 2.7x DRAM reads).
  auto q_item = [&](int w, int k) { return reinterpret_cast<volatile int *>(generic(bars + 320 + (w * TCB_QS + (k % TCB_QS)) * 4)); };
  auto q_full = [&](int w, int k) { return bars + 512 + (w * TCB_QS + (k % TCB_QS)) * 8; };

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
        mbar_init(bar_of(w, q), 4);            // one arrival per epilogue warp
      }
      for (int k = 0; k < TCB_QS; ++k) mbar_init(q_full(w, k), 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&mPcol);
    tma_prefetch_desc(&mProw);
    tma_prefetch_desc(&mDVA);
    tma_prefetch_desc(&mVA);
  }
  if (warp == 10) tc_alloc(tmem_slot, TCB_TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_ptr;
  const int oq0 = D.off_q[0], oq1 = D.off_q[1], ok0 = D.off_k[0], ok1 = D.off_k[1], ov0 = D.off_v[0], ov1 = D.off_v[1];

  if (warp >= 12) {
    // ------------------------------------------------------------------------------------------------ epilogue warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TCB_REGS_EPI));
    const int q = warp & 3, i = q * 16 + (lane & 15), kh = lane >> 4;
    const int64_t ld = D.ld;
    const float sc = D.scale;
    int n_[2] = {0, 0}, k_[2] = {0, 0}, j_[2] = {0, 0};       // per stream: junction counter, queue cursor, junction in item
    bool live[2] = {true, true};
    int64_t qoff[2] = {0, 0}, koff[2] = {0, 0}, kstep[2] = {0, 0};     // element offsets of this lane's rows at j = 0
    auto load_item = [&](int w) {                               // decode item k_[w] of stream w (or mark the stream finished)
      mbar_wait(q_full(w, k_[w]), (uint32_t)((k_[w] / TCB_QS) & 1));
      const int t = *q_item(w, k_[w]);
      if (t < 0) { live[w] = false; return; }
      const int b = t / (2 * H), r = t - b * 2 * H, dir = r / H, h = r - dir * H;
      // kh == 0: dQ row (b, i, j) and dV row; kh == 1: dK row.  i plays the key index for dK / dV: inward keys are row j
      // of e ((b, j, k)), outward keys are column j ((b, k, j))
      const int64_t qrow0 = (int64_t)(b * N + i) * N, krow0 = dir == 0 ? (int64_t)b * N * N + i : qrow0;
      qoff[w] = qrow0 * ld + (dir ? oq1 : oq0) + h * HD;
      koff[w] = krow0 * ld + (kh == 0 ? (dir ? ov1 : ov0) : (dir ? ok1 : ok0)) + h * HD;
      kstep[w] = dir == 0 ? (int64_t)N * ld : ld;
    };
    load_item(0);
    load_item(1);
    while (live[0] || live[1]) {
      bool did = false;
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        if (!live[w]) continue;
        const int n = n_[w];
        if (!mbar_test(bar_o(w, n & 1), (uint32_t)((n >> 1) & 1))) continue;
        did = true;
        tc_fence_after();
        const uint32_t tO = tmem + (uint32_t)w * TCB_COLS_WG + ((uint32_t)(q * 32) << 16) + TCB_COL_OUT + (uint32_t)(n & 1) * 32u;
        uint32_t o[16];
        tcx_ld16(tO, o);                                      // dQ (lanes 0-15) | dK (lanes 16-31)
        tc_ld_wait();
        const int j = j_[w];
        if (i < N) {
          uint32_t wv[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) wv[c] = Mma<T>::pack(__uint_as_float(o[2 * c]) * sc, __uint_as_float(o[2 * c + 1]) * sc);
          uint4 *dst = reinterpret_cast<uint4 *>(dproj + (kh == 0 ? qoff[w] + (int64_t)j * ld : koff[w] + (int64_t)j * kstep[w]));
          dst[0] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
          dst[1] = make_uint4(wv[4], wv[5], wv[6], wv[7]);
        }
        tcx_ld16(tO + 16u, o);                                // dV (lanes 0-15)
        tc_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_of(w, n & 1));         // this quadrant of the output accumulator is free again
        if (i < N && kh == 0) {
          uint32_t vv[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) vv[c] = Mma<T>::pack(__uint_as_float(o[2 * c]), __uint_as_float(o[2 * c + 1]));
          uint4 *dst = reinterpret_cast<uint4 *>(dproj + koff[w] + (int64_t)j * kstep[w]);
          dst[0] = make_uint4(vv[0], vv[1], vv[2], vv[3]);
          dst[1] = make_uint4(vv[4], vv[5], vv[6], vv[7]);
        }
        ++n_[w];
        if (++j_[w] == N) {
          j_[w] = 0;
          ++k_[w];
          load_item(w);
        }
      }
      if (!did) __nanosleep(40);
    }
  } else if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TCB_REGS_CTRL));
    const int w = warp & 1;                                  // the compute warpgroup this control warp serves
    const uint32_t wg_smem = sbase + w * TCB_WG_BYTES;
    if (warp < 10) {
      // ------------------------------------------------------------------------------------------ TMA producer
      // (the whole warp walks the loops with warp-uniform values; lane 0 draws the work items, one elected lane issues)
      auto draw = [&]() { return __shfl_sync(0xffffffffu, lane == 0 ? atomicAdd(item_counter, 1) : 0, 0); };
      int n = 0, t_next = draw();
      for (int k = 0;; ++k) {
        const int t = t_next < total_items ? t_next : -1;
        if (elect_one()) {
          *q_item(w, k) = t;
          mbar_arrive(q_full(w, k));                         // release: the id is visible to whoever waits on the slot
        }
        __syncwarp();
        if (t < 0) break;
        t_next = draw();                                     // next id: its latency hides behind this item's loads
        const int b = t / (2 * H), r = t - b * 2 * H, dir = r / H, h = r - dir * H;
        const int cq = (dir ? oq1 : oq0) + h * HD, ck = (dir ? ok1 : ok0) + h * HD, cv = (dir ? ov1 : ov0) + h * HD;
        const int co = dir * H * HD + h * HD;
        for (int j = 0; j < N; ++j, ++n) {
          const int s = n % TCB_STAGES;
          mbar_wait_polite(bar_empty(w, s), (uint32_t)(((n / TCB_STAGES) & 1) ^ 1));
          const uint32_t st = wg_smem + s * TCB_STAGE_BYTES, bar = bar_full(w, s);
          if (elect_one()) {
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
            tma_load_4d(&mVA, bar, st + 4 * TCF_TILE, co, j, 0, b);
          }
          __syncwarp();
        }
      }
    } else if (warp >= 10) {
      // ------------------------------------------------------------------------------------------ MMA issuer
      const uint32_t tw = tmem + (uint32_t)w * TCB_COLS_WG;
      const uint32_t sX16 = (wg_smem + TCB_STAGES * TCB_STAGE_BYTES) >> 4;       // [buf][dS | A], address >> 4
      const uint32_t id_sa = umma_idesc_f16<T>(64, 32, false, false);
      const uint32_t id_dq = umma_idesc_f16<T>(64, 16, false, true);
      const uint32_t id_kv = umma_idesc_f16<T>(64, 16, true, true);
      constexpr uint32_t HI32 = umma_desc_hi(256, UMMA_SW32), HI128 = umma_desc_hi(1024, UMMA_SW128);
      int ns = 0, ks = 0, js = 0;                            // S / dA tiles issued so far; queue cursor of the next one
      bool more = true;
      auto try_issue_sa = [&]() {                            // S(ns), dA(ns) -> buffer ns & 1, if that junction exists
        if (!more) return;
        if (js == 0) {
          mbar_wait_polite(q_full(w, ks), (uint32_t)((ks / TCB_QS) & 1));
          if (*q_item(w, ks) < 0) { more = false; return; }
        }
        const int s = ns % TCB_STAGES;
        mbar_wait_polite(bar_full(w, s), (uint32_t)((ns / TCB_STAGES) & 1));
        tc_fence_after();
        const uint32_t st16 = (wg_smem + s * TCB_STAGE_BYTES) >> 4;
        const uint32_t tSA = tw + (uint32_t)(ns & 1) * 64u;
        const uint64_t dQ_ = umma_desc_at(st16, 0, HI32), dO_ = umma_desc_at(st16, 3 * TCF_TILE, HI32);
        if (elect_one()) {
#pragma unroll
          for (int kh = 0; kh < 2; ++kh) {                   // keys 32 kh .. 32 kh + 31 -> lane offset 16 kh
            tc_mma(tSA + kh * TC_LANE16, dQ_, umma_desc_at(st16, TCF_TILE + kh * 1024, HI32), id_sa, 0u);
            tc_mma(tSA + 32u + kh * TC_LANE16, dO_, umma_desc_at(st16, 2 * TCF_TILE + kh * 1024, HI32), id_sa, 0u);
          }
          tc_commit(bar_s(w, ns & 1));
        }
        __syncwarp();
        ++ns;
        if (++js == N) { js = 0; ++ks; }
      };
      try_issue_sa();
      try_issue_sa();
      for (int n = 0; n < ns; ++n) {
        mbar_wait_polite(bar_t(w, n & 1), (uint32_t)((n >> 1) & 1));     // dS(n), A(n) are in shared memory; S(n), dA(n) were read
        if (n >= 2) mbar_wait_polite(bar_of(w, n & 1), (uint32_t)(((n - 2) >> 1) & 1));   // outputs of n-2 have left TMEM
        tc_fence_after();
        const uint32_t st16 = (wg_smem + (n % TCB_STAGES) * TCB_STAGE_BYTES) >> 4;
        const uint32_t dS16 = sX16 + (uint32_t)(n & 1) * (2 * TCB_XT >> 4), a16 = dS16 + (TCB_XT >> 4);
        const uint32_t tOut = tw + TCB_COL_OUT + (uint32_t)(n & 1) * 32u;
        if (elect_one()) {
#pragma unroll
        for (int ks_ = 0; ks_ < 4; ++ks_) {
          // dQ[i, d] += dS[i, 16ks..] K[16ks.., d]        A: dS K-major (32 B per step), B: K tile MN-major
          tc_mma(tOut, umma_desc_at(dS16, ks_ * 32, HI128), umma_desc_at(st16, TCF_TILE + ks_ * 512, HI32), id_dq,
                 (uint32_t)(ks_ != 0));
          // dK[k, d] += dS[16ks.., k]^T Q[16ks.., d]      A: dS MN-major (two 8-row groups per step), B: Q tile MN-major
          tc_mma(tOut + TC_LANE16, umma_desc_at(dS16, ks_ * 2048, HI128), umma_desc_at(st16, ks_ * 512, HI32), id_kv,
                 (uint32_t)(ks_ != 0));
          // dV[k, d] += A[16ks.., k]^T dO[16ks.., d]
          tc_mma(tOut + 16u, umma_desc_at(a16, ks_ * 2048, HI128), umma_desc_at(st16, 3 * TCF_TILE + ks_ * 512, HI32), id_kv,
                 (uint32_t)(ks_ != 0));
        }
        tc_commit(bar_o(w, n & 1));
        tc_commit(bar_empty(w, n % TCB_STAGES));
        }
        __syncwarp();
        try_issue_sa();
      }
    }
    __syncwarp();
  } else {
    // ---------------------------------------------------------------------------------------------- compute warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(TCB_REGS_WORK));
    const int w = warp >> 2, q = warp & 3;
    const int i = q * 16 + (lane & 15), kh = lane >> 4;
    const uint32_t wg_smem = sbase + w * TCB_WG_BYTES;
    const uint32_t sX = wg_smem + TCB_STAGES * TCB_STAGE_BYTES;
    const uint32_t tl = tmem + (uint32_t)w * TCB_COLS_WG + ((uint32_t)(q * 32) << 16);
    const uint32_t row_off = (uint32_t)i * 128u;                                        // this row in a 64 x 64 tile
    // this row in a [64 x 16] SWIZZLE_32B stage tile: 16-byte halves at row32 + half0 / half1
    const uint32_t row32 = (uint32_t)i * 32u, half0 = (uint32_t)(((i >> 2) & 1) << 4), half1 = half0 ^ 16u;
    constexpr float LN2 = 0.6931471805599453f;
    const float2 l2e = make_float2(LOG2E, LOG2E);

    // per item, constant over the junctions: E + mask (natural-log domain; exactly the 16-bit projection value, hugely
    // negative when masked) and the fp16 gate of this thread's 32 keys, packed two per register; dE / dG accumulators
    uint32_t ebk[16], gtk[16];
    float2 dE[16], dG[16];
    float sc = 0.f;
    int64_t tile_base = 0;                                   // this thread's 32 entries of the (b, dir, h) bias / gate tiles
    const float *lse_row = stats;                            // stats[b, dir, h, 0, i]
    int kq = 0;
    auto next_item = [&]() {
      mbar_wait(q_full(w, kq), (uint32_t)((kq / TCB_QS) & 1));
      const int t = *q_item(w, kq);
      ++kq;
      return t;
    };
    auto setup = [&](int t) {
      const int b = t / (2 * H), r = t - b * 2 * H, dir = r / H, h = r - dir * H;
      tile_base = (((int64_t)(b * 2 + dir) * H + h) * TN + i) * TN + 32 * kh;
      lse_row = stats + ((int64_t)(b * 2 + dir) * H + h) * N * N + i;
      const float4 *er = reinterpret_cast<const float4 *>(ws_e + tile_base);
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
      const uint4 *gr = reinterpret_cast<const uint4 *>(ws_g + tile_base);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint4 g4 = gr[c];
        gtk[4 * c] = g4.x; gtk[4 * c + 1] = g4.y; gtk[4 * c + 2] = g4.z; gtk[4 * c + 3] = g4.w;
      }
#pragma unroll
      for (int c = 0; c < 16; ++c) dE[c] = dG[c] = make_float2(0.f, 0.f);
    };
    auto finish = [&]() {                                               // dE, dG = g (1 - g) * acc -> [B,2,H,64,64] tiles
      float4 *de = reinterpret_cast<float4 *>(ws_de + tile_base), *dg = reinterpret_cast<float4 *>(ws_dg + tile_base);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float2 g0 = __half22float2(*reinterpret_cast<const __half2 *>(&gtk[2 * c]));
        const float2 g1 = __half22float2(*reinterpret_cast<const __half2 *>(&gtk[2 * c + 1]));
        de[c] = make_float4(dE[2 * c].x, dE[2 * c].y, dE[2 * c + 1].x, dE[2 * c + 1].y);
        dg[c] = make_float4(dG[2 * c].x * g0.x * (1.f - g0.x), dG[2 * c].y * g0.y * (1.f - g0.y),
                            dG[2 * c + 1].x * g1.x * (1.f - g1.x), dG[2 * c + 1].y * g1.y * (1.f - g1.y));
      }
    };

    int t = next_item(), j = 0;
    if (t >= 0) setup(t);
    float lse_next = (t >= 0 && i < N) ? __ldg(lse_row) : INFINITY;
    for (int n = 0; t >= 0; ++n) {
      const float lse2 = lse_next;                       // log2-domain log-sum-exp of row i at junction j (+inf: padding row)
      if (j + 1 < N && i < N) lse_next = __ldg(lse_row + (int64_t)(j + 1) * N);
      mbar_wait(bar_s(w, n & 1), (uint32_t)((n >> 1) & 1));     // S(n), dA(n) in tensor memory (=> the stage's tiles have landed)
      tc_fence_after();
      const uint32_t tSA = tl + (uint32_t)(n & 1) * 64u;
      uint32_t sv[16], av[16];
      tcx_ld16(tSA, sv);
      tcx_ld16(tSA + 32u, av);
      // delta[i] = sum_k dP P = sum_d dO[i, d] O[i, d]  (O = A V is the forward output): one pass over the scores suffices
      float delta;
      {
        const uint32_t st = wg_smem + (uint32_t)(n % TCB_STAGES) * TCB_STAGE_BYTES + row32;
        const uint4 d0 = ld_shared_v4u(st + 3 * TCF_TILE + half0), d1 = ld_shared_v4u(st + 3 * TCF_TILE + half1);
        const uint4 o0 = ld_shared_v4u(st + 4 * TCF_TILE + half0), o1 = ld_shared_v4u(st + 4 * TCF_TILE + half1);
        const uint32_t dw[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const uint32_t ow[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 8; ++c) acc = __ffma2_rn(unpack16x2<T>(dw[c]), unpack16x2<T>(ow[c]), acc);
        delta = acc.x + acc.y;
      }
      const float2 scp = make_float2(sc, sc), nl = make_float2(-lse2, -lse2), nd = make_float2(-delta, -delta);
      // the dS / A tiles of junction n-2 were read by UMMAs that have completed when o_ready(n-2) has
      if (n >= 2) mbar_wait(bar_o(w, n & 1), (uint32_t)(((n - 2) >> 1) & 1));
      const uint32_t sDS = sX + (uint32_t)(n & 1) * 2 * TCB_XT + row_off, sA = sDS + TCB_XT;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {                   // 16 keys per chunk
        tc_ld_wait();
        uint32_t s2[16], a2[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) { s2[c] = sv[c]; a2[c] = av[c]; }
        if (ch == 0) {                                   // the second chunk flies during the math of the first
          tcx_ld16(tSA + 16u, sv);
          tcx_ld16(tSA + 48u, av);
        }
        uint32_t dpk[8], apk[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const int cc = 8 * ch + c;
          const float2 x = __ffma2_rn(make_float2(__uint_as_float(s2[2 * c]), __uint_as_float(s2[2 * c + 1])), scp,
                                      unpack16x2<T>(ebk[cc]));
          const float2 e2 = __ffma2_rn(x, l2e, nl);
          const float2 p = make_float2(fast_exp2(e2.x), fast_exp2(e2.y));
          const float2 da = make_float2(__uint_as_float(a2[2 * c]), __uint_as_float(a2[2 * c + 1]));
          const float2 g2 = __half22float2(*reinterpret_cast<const __half2 *>(&gtk[cc]));
          dG[cc] = __ffma2_rn(da, p, dG[cc]);
          const float2 ds = __fmul2_rn(p, __ffma2_rn(da, g2, nd));          // P (dA g - delta)
          dE[cc] = __fadd2_rn(dE[cc], ds);
          const float2 aa = __fmul2_rn(p, g2);
          dpk[c] = Mma<T>::pack(ds.x, ds.y);
          apk[c] = Mma<T>::pack(aa.x, aa.y);
        }
        // this thread's 32 bytes per tile of row i (chunks 4 kh + 2 ch, + 1; 128B swizzle)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t off = (uint32_t)(((4 * kh + 2 * ch + c) ^ (i & 7)) << 4);
          st_shared_v4u(sDS + off, dpk[4 * c], dpk[4 * c + 1], dpk[4 * c + 2], dpk[4 * c + 3]);
          st_shared_v4u(sA + off, apk[4 * c], apk[4 * c + 1], apk[4 * c + 2], apk[4 * c + 3]);
        }
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_t(w, n & 1));
      if (++j == N) {
        finish();
        t = next_item();
        j = 0;
        if (t >= 0) {
          setup(t);
          lse_next = i < N ? __ldg(lse_row) : INFINITY;
        }
      }
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
static int bwd_tc_impl(const tgt_triplet_attn_desc &D, const void *proj, const void *va, const void *dva, const float *stats,
                       void *dproj, const float *ws_e, const __half *ws_g, float *ws_de, float *ws_dg, int *counter,
                       cudaStream_t st) {
  CUtensorMap mPcol, mProw, mDVA, mVA;
  const int C = (int)D.ld, Cv = 2 * D.H * HD;
  if (int e = make_map(&mPcol, proj, D.B, D.N, C, D.ld, true, 16, D.dtype)) return e;
  if (int e = make_map(&mProw, proj, D.B, D.N, C, D.ld, false, 16, D.dtype)) return e;
  if (int e = make_map(&mDVA, dva, D.B, D.N, Cv, Cv, true, 16, D.dtype)) return e;
  if (int e = make_map(&mVA, va, D.B, D.N, Cv, Cv, true, 16, D.dtype)) return e;
  int dev = 0, sms = 0;
  TGT_CUDA_OK(cudaGetDevice(&dev));
  TGT_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TGT_CUDA_OK(cudaFuncSetAttribute(tri_attn_bwd_tc<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, TCB_SMEM));
  const int items = D.B * 2 * D.H;
  const int grid = std::min(sms, (items + 1) / 2);
  TGT_CUDA_OK(cudaMemsetAsync(counter, 0, sizeof(int), st));        // work-item counter of the persistent CTAs
  KernelTimerScope ts("tri_attn_bwd_tc", st);
  tri_attn_bwd_tc<T><<<grid, TCB_THREADS, TCB_SMEM, st>>>(D, mPcol, mProw, mDVA, mVA, ws_e, ws_g, stats, (T *)dproj, ws_de, ws_dg,
                                                          counter);
  return check_launch("tri_attn_bwd_tc");
}

int triplet_attn_bwd_tc_launch(const tgt_triplet_attn_desc &D, const void *proj, const void *va, const void *dva,
                               const float *stats, void *dproj, const float *ws_e, const __half *ws_g, float *ws_de,
                               float *ws_dg, int *counter, cudaStream_t st) {
  if (D.dtype == TGT_BF16)
    return bwd_tc_impl<__nv_bfloat16>(D, proj, va, dva, stats, dproj, ws_e, ws_g, ws_de, ws_dg, counter, st);
  return bwd_tc_impl<__half>(D, proj, va, dva, stats, dproj, ws_e, ws_g, ws_de, ws_dg, counter, st);
}

}  // namespace tgt
