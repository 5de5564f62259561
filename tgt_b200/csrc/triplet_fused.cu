// Fused triplet-attention forward for the shipped geometry (head dim 16, edge width <= 256, N <= 64), sm_100a:
//
//     LayerNorm -> lin_QKV_in / lin_QKV_out  ->  per-junction N x N x d attention (bias, mask, gate)  ->  Va
//     (reference lib/tgt/layers/triplet.py:207-246)
//
// in ONE kernel, so that the [R, 6*We] projection tensor (3.2 GB at B=256, N=64 -- 75% of the bytes the un-fused
// forward moves) never exists in HBM.  One CTA = (graph b, pair of heads); it loops over pairs of junction atoms j:
//
//   warp 0    lane 0 = TMA producer: the LN-folded weight rows of its two heads (192 x We, once) and, per junction pair, the raw
//             edge rows as 128-row tiles: "column" tile rows (i, j0..j0+1) and "row" tile rows (j0..j0+1, k), 64-channel
//             K-blocks, 128B swizzle, 4-stage mbarrier ring, 4-D tensor maps (rows >= N zero-filled)
//   warp 0    lane 1 = tcgen05.mma issuer: D_col[128 x 128] = Xcol * [Q_in|Q_out|K_out|V_out]^T and
//             D_row[128 x 64] = Xrow * [K_in|V_in]^T into one of two TMEM accumulator sets
//   warps 4-19  four attention warpgroups (setmaxnreg moves the control warpgroup's registers to them), one per (head, direction): tcgen05.ld their 3 x 16 accumulator columns, apply
//             the LayerNorm fold (rstd_r * (acc - mean_r * colsum_c) + bias'_c, the epilogue of gemm_tc.cu), round to
//             16 bit into ldmatrix-ready 64 x 16 tiles, then run the same register-resident mma.sync attention as
//             triplet_tma.cu (bias / gate tiles of the (head, direction) pinned in registers for all N junctions) and
//             write Va with stmatrix + TMA tensor stores and the per-row log-sum-exp for backward.
//
// The tensor pipe works on junction pair t+1 while the attention warpgroups are busy with pair t.
// Bias / gate tiles come from a small LN-folded GEMM (E|G columns only) + tri_prep_bias_gate, as in the un-fused path.
#include "triplet_common.cuh"
#include <stdlib.h>
#include <algorithm>

namespace tgt {

namespace fused {
constexpr int HPC = 2;                          // heads per CTA
constexpr int NCOL = 4 * HPC * HD;              // 128 accumulator columns fed by the column tile
constexpr int NROW = 2 * HPC * HD;              // 64 accumulator columns fed by the row tile
constexpr int NW = NCOL + NROW;                 // 192 weight rows per head pair
constexpr int STAGES = 4;
constexpr int STAGE_BYTES = 128 * 128;          // one 128-row x 64-channel K-block
constexpr int KB_MAX = 4;                       // edge width <= 256
constexpr int WG = 4;                           // attention warpgroups
constexpr int THREADS = 128 + WG * 128;         // 640: control warpgroup (warp 0: lane 0 = TMA producer, lane 1 = MMA issuer;
                                                //      warps 1-3 idle) + 4 attention warpgroups
constexpr int REGS_CTRL = 24, REGS_ATTN = 112;  // setmaxnreg: 128*24 + 512*112 = 60416 <= 65536
constexpr int TILE_B = TN * HD * 2;             // 2 KB operand tile
constexpr int TILES_PER_WG = 2 * 3;             // (jj, {Q,K,V})
constexpr int SMEM_W = KB_MAX * NW * 128;       // 96 KB
constexpr int SMEM_TILES = WG * TILES_PER_WG * TILE_B;   // 48 KB
constexpr int SMEM_VEC = 2 * NW * 4;            // colsum, bias
constexpr int SMEM_TOTAL = SMEM_W + STAGES * STAGE_BYTES + SMEM_TILES + SMEM_VEC + 256 + 1024;
constexpr int TMEM_COLS = 512;                  // 2 x 192 accumulator columns
}  // namespace fused

// K-major operand tile with 128-byte rows, 128B swizzle (same descriptor as gemm_tc.cu)
__device__ __forceinline__ uint64_t fused_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <typename T> __device__ __forceinline__ uint32_t fused_idesc(int n) {
  const uint32_t fmt = DT<T>::code == TGT_BF16 ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void fused_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

template <typename T, int CL>
__global__ void __launch_bounds__(fused::THREADS, 1)
tri_fused_fwd(const tgt_triplet_attn_desc D, const int We, const __grid_constant__ CUtensorMap mXcol,
              const __grid_constant__ CUtensorMap mXrow, const __grid_constant__ CUtensorMap mWc,
              const __grid_constant__ CUtensorMap mWr, const __grid_constant__ CUtensorMap mVA,
              const float *__restrict__ row_mean, const float *__restrict__ row_rstd,
              const float *__restrict__ wcolsum, const float *__restrict__ wbias, const float *__restrict__ ws_e,
              const __half *__restrict__ ws_g, float *__restrict__ stats, const int dbg) {
  using namespace fused;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t sbase = (smem_u32(smem_raw) + 1023u) & ~1023u;
  unsigned char *sgen = smem_raw + (sbase - smem_u32(smem_raw));
  const int N = D.N, H = D.H;
  const int hp = blockIdx.x, b = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int kblocks = (We + 63) / 64;
  const int npairs = (N + 1) / 2;

  const uint32_t sWc = sbase;                                   // kblocks x [128 rows x 128 B]
  const uint32_t sWr = sWc + KB_MAX * NCOL * 128;               // kblocks x [ 64 rows x 128 B]
  const uint32_t sA = sbase + SMEM_W;
  const uint32_t sT = sA + STAGES * STAGE_BYTES;
  const uint32_t sVec = sT + SMEM_TILES;
  const uint32_t sBar = sVec + SMEM_VEC;
  float *vec_cs = reinterpret_cast<float *>(sgen + (sVec - sbase));
  float *vec_bs = vec_cs + NW;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * STAGES, bar_w = sBar + 16 * STAGES;
  const uint32_t bar_tfull = bar_w + 8, bar_tempty = bar_tfull + 16, tmem_slot = bar_tempty + 16;
  volatile uint32_t *tmem_slot_gen = reinterpret_cast<volatile uint32_t *>(sgen + (tmem_slot - sbase));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mXcol);
    tma_prefetch_desc(&mXrow);
    tma_prefetch_desc(&mWc);
    tma_prefetch_desc(&mWr);
    tma_prefetch_desc(&mVA);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bar_full + s * 8, 1);
      mbar_init(bar_empty + s * 8, CL);               // every CTA of the cluster must have drained the stage
    }
    mbar_init(bar_w, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + i * 8, 1);
      mbar_init(bar_tempty + i * 8, WG * 4);          // one arrival per attention warp
    }
    fence_barrier_init();
  }
  if (warp == 0) tc_alloc(tmem_slot, TMEM_COLS);
  // LN-fold vectors of this head pair (weight rows are already in panel order)
  for (int c = tid; c < NW; c += THREADS) {
    vec_cs[c] = wcolsum[hp * NW + c];
    vec_bs[c] = wbias[hp * NW + c];
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                    // barriers of every CTA exist before any remote arrive / multicast
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0u;
  constexpr uint16_t CMASK = (uint16_t)((1u << CL) - 1u);
  constexpr int SLICE_ROWS = 128 / CL;               // each CTA fetches 1/CL of every activation tile and multicasts it

  if (warp < 4) {
    // the control warpgroup hands its registers to the attention warpgroups
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_CTRL));
  }
  if (warp >= 1 && warp < 4) {
    // idle warps of the control warpgroup
  } else if (warp == 0) {
    // ------------------------------------------------------------------------------------------ TMA producer (lane 0)
    if (lane == 0) {
      mbar_expect_tx(bar_w, (uint32_t)kblocks * NW * 128);
      for (int kb = 0; kb < kblocks; ++kb) {
        tma_load_2d(&mWc, bar_w, sWc + kb * NCOL * 128, kb * 64, hp * NW);
        tma_load_2d(&mWr, bar_w, sWr + kb * NROW * 128, kb * 64, hp * NW + NCOL);
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < npairs; ++t) {
        for (int half = 0; half < 2; ++half) {
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(bar_empty + stage * 8, phase ^ 1);
            mbar_expect_tx(bar_full + stage * 8, STAGE_BYTES);
            const CUtensorMap *mp = half == 0 ? &mXcol : &mXrow;
            if (CL == 1) {
              tma_load_4d(mp, bar_full + stage * 8, sA + stage * STAGE_BYTES, kb * 64, 0, 2 * t, b);
            } else {
              // tile rows are (jj, row): slice `crank` = rows [crank*SLICE_ROWS, +SLICE_ROWS) of the 128
              const int r0 = (int)crank * SLICE_ROWS;
              tma_load_4d_mc(mp, bar_full + stage * 8, sA + stage * STAGE_BYTES + r0 * 128, kb * 64, r0 & 63,
                             2 * t + (r0 >> 6), b, CMASK);
            }
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    } else if (lane == 1) {
      // ---------------------------------------------------------------------------------------- MMA issuer (lane 1)
      const uint32_t idesc_c = fused_idesc<T>(NCOL), idesc_r = fused_idesc<T>(NROW);
      mbar_wait(bar_w, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int t = 0; t < npairs; ++t) {
        const int buf = t & 1;
        mbar_wait(bar_tempty + buf * 8, (uint32_t)(((t >> 1) & 1) ^ 1));
        tc_fence_after();
        for (int half = 0; half < 2; ++half) {
          const uint32_t tmem_d = tmem_base + (uint32_t)(buf * NW + (half ? NCOL : 0));
          for (int kb = 0; kb < kblocks; ++kb) {
            mbar_wait(bar_full + stage * 8, phase);
            tc_fence_after();
            const int ksteps = min(4, (We - kb * 64 + 15) / 16);
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t ad = fused_desc_sw128(sA + stage * STAGE_BYTES + k * 32);
              const uint64_t bd = fused_desc_sw128((half ? sWr + kb * NROW * 128 : sWc + kb * NCOL * 128) + k * 32);
              if (!(dbg & 2)) tc_mma(tmem_d, ad, bd, half ? idesc_r : idesc_c, (uint32_t)((kb | k) != 0));
            }
            if (CL == 1) tc_commit(bar_empty + stage * 8);
            else tc_commit_mc(bar_empty + stage * 8, CMASK);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        tc_commit(bar_tfull + buf * 8);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------------------------------ attention warpgroups
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_ATTN));
    const int wg = (warp - 4) >> 2;              // 0..3
    const int wq = warp & 3;                     // TMEM lane quadrant of this warp
    const int wl = warp & 3;                     // warp index inside the warpgroup -> 16 query rows
    const int hl = wg & 1, dir = wg >> 1;
    const int h = hp * HPC + hl;
    const int g = lane >> 2, q = lane & 3;
    const int m0 = wl * 16;
    const uint32_t myT = sT + (uint32_t)wg * TILES_PER_WG * TILE_B;      // tiles: [jj][Q,K,V]
    const int bar_id = 1 + wg;

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

    // accumulator columns of this (head, direction) and the matching LN-fold vector entries
    //   inward : Q = col-tile cols [hl*16), K = row-tile cols [hl*16), V = row-tile cols [32 + hl*16)
    //   outward: Q = col-tile cols [32 + hl*16), K = [64 + hl*16), V = [96 + hl*16)   (all from the column tile)
    const int cQ = (dir ? HPC * HD : 0) + hl * HD;
    const int cK = dir ? (2 * HPC * HD + hl * HD) : (NCOL + hl * HD);
    const int cV = dir ? (3 * HPC * HD + hl * HD) : (NCOL + HPC * HD + hl * HD);
    const int co = dir * H * HD + h * HD;
    // the TMEM lane of this thread: r = wq*32 + lane.  Both tiles are ordered (jj, row): column tile rows are
    // (jj, i) = (r >> 6, r & 63), row tile rows are (jj, k) = (r >> 6, r & 63)
    const int r = wq * 32 + lane;
    const int ci = r & 63, cjj = r >> 6, rjj = r >> 6, rk = r & 63;

    for (int t = 0; t < npairs; ++t) {
      const int buf = t & 1;
      const int j0 = 2 * t;
      // LayerNorm statistics of the two edge rows this thread converts (column-tile row, row-tile row)
      float c_rstd = 0.f, c_nmr = 0.f, r_rstd = 0.f, r_nmr = 0.f;
      if (ci < N && j0 + cjj < N) {
        const int64_t gr = ((int64_t)(b * N + ci)) * N + j0 + cjj;
        c_rstd = row_rstd[gr];
        c_nmr = -row_mean[gr] * c_rstd;
      }
      if (!dir && rk < N && j0 + rjj < N) {
        const int64_t gr = ((int64_t)(b * N + j0 + rjj)) * N + rk;
        r_rstd = row_rstd[gr];
        r_nmr = -row_mean[gr] * r_rstd;
      }
      // the Va tiles of the previous junction pair must have left shared memory before their rows are overwritten
      if (wl == 0 && lane == 0) tma_store_wait_read();
      mbar_wait(bar_tfull + buf * 8, (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (dbg & 1) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_tempty + buf * 8);
        continue;
      }
      // ---- TMEM -> LN fold -> 16-bit operand tiles
      const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(buf * NW);
#pragma unroll
      for (int m = 0; m < 3; ++m) {
        const int c = m == 0 ? cQ : (m == 1 ? cK : cV);
        uint32_t v[16];
        fused_ld16(tbase + c, v);
        tc_ld_wait();
        const bool from_row = (m > 0) && !dir;
        const float rs = from_row ? r_rstd : c_rstd, nm = from_row ? r_nmr : c_nmr;
        const int jj = from_row ? rjj : cjj, trow = from_row ? rk : ci;
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 cs = *reinterpret_cast<const float4 *>(vec_cs + c + 4 * i);
          const float4 bs = *reinterpret_cast<const float4 *>(vec_bs + c + 4 * i);
          const float a0 = fmaf(__uint_as_float(v[4 * i + 0]), rs, fmaf(nm, cs.x, bs.x));
          const float a1 = fmaf(__uint_as_float(v[4 * i + 1]), rs, fmaf(nm, cs.y, bs.y));
          const float a2 = fmaf(__uint_as_float(v[4 * i + 2]), rs, fmaf(nm, cs.z, bs.z));
          const float a3 = fmaf(__uint_as_float(v[4 * i + 3]), rs, fmaf(nm, cs.w, bs.w));
          w[2 * i] = Mma<T>::pack(a0, a1);
          w[2 * i + 1] = Mma<T>::pack(a2, a3);
        }
        const uint32_t dst = myT + (uint32_t)(jj * 3 + m) * TILE_B;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + tile_off(trow, 0)), "r"(w[0]), "r"(w[1]),
                     "r"(w[2]), "r"(w[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst + tile_off(trow, 1)), "r"(w[4]), "r"(w[5]),
                     "r"(w[6]), "r"(w[7])
                     : "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_tempty + buf * 8);      // this warp no longer needs the accumulators
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");

      // ---- attention for the two junctions of the pair (same arithmetic as tri_attn_fwd_tma)
#pragma unroll 1
      for (int jj = 0; jj < 2; ++jj) {
        const int j = j0 + jj;
        if (j >= N || (dbg & 4)) break;
        const uint32_t sQ = myT + (uint32_t)(jj * 3) * TILE_B, sK = sQ + TILE_B, sV = sK + TILE_B;
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
        for (int tt = 0; tt < 4; ++tt) {
          uint32_t vb[4];
          load_b_kn(vb, sV, tt * 16, lane);
          Mma<T>::run(o[0], pa[tt], vb[0], vb[1]);
          Mma<T>::run(o[1], pa[tt], vb[2], vb[3]);
        }
        o[0][0] *= inv0; o[0][1] *= inv0; o[1][0] *= inv0; o[1][1] *= inv0;
        o[0][2] *= inv1; o[0][3] *= inv1; o[1][2] *= inv1; o[1][3] *= inv1;
        // Va rows of this warp overwrite the Q rows it alone has read (rows m0 .. m0+15 of the Q tile)
        __syncwarp();
        store_c_tile<T>(sQ, m0, lane, o, 1.f);
        const int i0 = m0 + g, i1 = m0 + g + 8;
        if (q == 0) {          // one float per row: log2-domain log-sum-exp  (P = exp2(x - lse2))
          float *stp = stats + (((int64_t)(b * 2 + dir) * H + h) * N + j) * N;
          if (i0 < N) stp[i0] = mx0 + __log2f(l0);
          if (i1 < N) stp[i1] = mx1 + __log2f(l1);
        }
      }
      fence_proxy_async();
      asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
      if (wl == 0 && lane == 0) {
        tma_store_4d(&mVA, myT, co, j0, 0, b);
        if (j0 + 1 < N) tma_store_4d(&mVA, myT + 3 * TILE_B, co, j0 + 1, 0, b);
        tma_store_commit();
      }
    }
    if (wl == 0 && lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();                    // nobody leaves while a peer may still multicast into it
  if (warp == 0) {
    tc_fence_after();
    tc_dealloc(tmem_base, fused::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------ host
static int fused_map(CUtensorMap *map, const void *base, int rank, const cuuint64_t *gdim, const cuuint64_t *gstride,
                     const cuuint32_t *box, CUtensorMapSwizzle sw, int dtype) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("triplet_attn_fused: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const CUtensorMapDataType dt = dtype == TGT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, (cuuint32_t)rank, const_cast<void *>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("triplet_attn_fused: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

bool triplet_attn_fused_supported(const tgt_triplet_attn_desc &D, int We) {
  if (D.dtype != TGT_BF16 && D.dtype != TGT_F16) return false;
  if (D.d != HD || D.N > TN || D.N < 1 || (D.H % fused::HPC)) return false;
  if (We % 8 || We < 16 || We > 64 * fused::KB_MAX) return false;
  return encode_tiled_fn() != nullptr;
}

template <typename T, int CL>
static int fused_impl(const tgt_triplet_attn_desc &D, int We, const void *x, int64_t ldx, const float *mean,
                      const float *rstd, const void *wf, const float *wcolsum, const float *wbias, void *va,
                      float *stats, const float *ws_e, const __half *ws_g, cudaStream_t st) {
  using namespace fused;
  CUtensorMap mXcol, mXrow, mWc, mWr, mVA;
  {
    // x[b, i, j, :]: the row tile walks the last edge index (stride ldx), the column tile the first one (stride N*ldx);
    // both maps list the walked index first so that either tile lands as [2 junctions][64 rows][64 channels]
    const cuuint64_t gdim[4] = {(cuuint64_t)We, (cuuint64_t)D.N, (cuuint64_t)D.N, (cuuint64_t)D.B};
    const cuuint64_t gstr_r[3] = {(cuuint64_t)ldx * 2, (cuuint64_t)D.N * ldx * 2, (cuuint64_t)D.N * D.N * ldx * 2};
    const cuuint64_t gstr_c[3] = {(cuuint64_t)D.N * ldx * 2, (cuuint64_t)ldx * 2, (cuuint64_t)D.N * D.N * ldx * 2};
    // CL == 1: one box = the whole tile [2][64 rows]; CL > 1: one box = this CTA's slice of 128 / CL rows
    const cuuint32_t bx[4] = {64u, CL == 1 ? 64u : (cuuint32_t)std::min(64, 128 / CL), CL == 1 ? 2u : (CL == 2 ? 1u : 1u), 1u};
    if (int e = fused_map(&mXcol, x, 4, gdim, gstr_c, bx, CU_TENSOR_MAP_SWIZZLE_128B, D.dtype)) return e;
    if (int e = fused_map(&mXrow, x, 4, gdim, gstr_r, bx, CU_TENSOR_MAP_SWIZZLE_128B, D.dtype)) return e;
  }
  {
    const cuuint64_t gdim[2] = {(cuuint64_t)We, (cuuint64_t)(D.H / HPC) * NW};
    const cuuint64_t gstr[1] = {(cuuint64_t)We * 2};
    const cuuint32_t bc[2] = {64u, (cuuint32_t)NCOL}, br[2] = {64u, (cuuint32_t)NROW};
    if (int e = fused_map(&mWc, wf, 2, gdim, gstr, bc, CU_TENSOR_MAP_SWIZZLE_128B, D.dtype)) return e;
    if (int e = fused_map(&mWr, wf, 2, gdim, gstr, br, CU_TENSOR_MAP_SWIZZLE_128B, D.dtype)) return e;
  }
  {
    const int Cv = 2 * D.H * HD;
    const cuuint64_t gdim[4] = {(cuuint64_t)Cv, (cuuint64_t)D.N, (cuuint64_t)D.N, (cuuint64_t)D.B};
    const cuuint64_t gstr[3] = {(cuuint64_t)Cv * 2, (cuuint64_t)D.N * Cv * 2, (cuuint64_t)D.N * D.N * Cv * 2};
    const cuuint32_t bx[4] = {(cuuint32_t)HD, 1u, (cuuint32_t)TN, 1u};
    if (int e = fused_map(&mVA, va, 4, gdim, gstr, bx, CU_TENSOR_MAP_SWIZZLE_32B, D.dtype)) return e;
  }
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(tri_fused_fwd<T, CL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
  });
  static const int dbg = [] { const char *v = getenv("TGT_FUSED_DEBUG"); return v ? atoi(v) : 0; }();   // ablations
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(D.H / HPC, D.B);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  KernelTimerScope ts("tri_fused_fwd", st);
  cudaError_t err = cudaLaunchKernelEx(&cfg, tri_fused_fwd<T, CL>, D, We, mXcol, mXrow, mWc, mWr, mVA, mean, rstd, wcolsum,
                                       wbias, ws_e, ws_g, stats, dbg);
  if (err != cudaSuccess) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    return fail("tri_fused_fwd: %s", cudaGetErrorString(err));
  }
  return check_launch("tri_fused_fwd");
}

// cluster size: the 8 head-pair CTAs of a graph share their activation tiles by TMA multicast (TGT_FUSED_CLUSTER=1/2/4/8)
static int fused_cluster(int ctas_per_graph) {
  static const int want = [] { const char *v = getenv("TGT_FUSED_CLUSTER"); return v ? atoi(v) : 1; }();
  int cl = want;
  while (cl > 1 && (ctas_per_graph % cl)) cl >>= 1;
  return cl < 1 ? 1 : cl;
}

int triplet_attn_fused_launch(const tgt_triplet_attn_desc &D, int We, const void *x, int64_t ldx, const float *mean,
                              const float *rstd, const void *wf, const float *wcolsum, const float *wbias, void *va,
                              float *stats, const float *ws_e, const __half *ws_g, cudaStream_t st) {
#define TGT_FUSED_GO(TT, C) return fused_impl<TT, C>(D, We, x, ldx, mean, rstd, wf, wcolsum, wbias, va, stats, ws_e, ws_g, st)
  const int cl = fused_cluster(D.H / fused::HPC);
  if (D.dtype == TGT_BF16) {
    if (cl == 8) TGT_FUSED_GO(__nv_bfloat16, 8);
    if (cl == 4) TGT_FUSED_GO(__nv_bfloat16, 4);
    if (cl == 2) TGT_FUSED_GO(__nv_bfloat16, 2);
    TGT_FUSED_GO(__nv_bfloat16, 1);
  }
  if (cl == 8) TGT_FUSED_GO(__half, 8);
  if (cl == 4) TGT_FUSED_GO(__half, 4);
  if (cl == 2) TGT_FUSED_GO(__half, 2);
  TGT_FUSED_GO(__half, 1);
#undef TGT_FUSED_GO
}

}  // namespace tgt
