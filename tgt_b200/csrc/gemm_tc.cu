// Persistent, warp-specialised tcgen05 GEMM for the edge-channel projections of the TGT hot path:
//
//     D[M, N] = epilogue( A[M, K] * B[N, K]^T )          A, B, D 16-bit (bf16 / fp16), fp32 accumulation in TMEM
//
// M = B*N*N edge rows (~1e6), K <= 512, N <= a few thousand: the weight matrix is tiny and the activation matrix is
// huge, so the kernel is WEIGHT-STATIONARY: N is cut into column slices of width bn <= 256; one CTA owns a slice,
// loads its [bn, K] weight panel into shared memory ONCE (TMA, 128B swizzle) and then streams 128-row activation
// tiles through a TMA ring.  One elected thread issues tcgen05.mma (M=128, N=bn, K=16) into one of two TMEM
// accumulators while four epilogue warps drain the other one (tcgen05.ld 32x32b), apply the fused epilogue and store
// with 16-byte vectors.  L2->SMEM traffic is 64 KB per 128x256x256 tile (32 B/clk/SM at full tensor rate), which is
// what lets small-K GEMMs run at tensor-core speed without 2-CTA multicast.
//
// Fused epilogues (flags):
//   LN    : the GEMM runs on the RAW rows x and  LN(x) W^T  is recovered as  rstd_r * (acc - mean_r * colsum_c) + b'_c
//           with B = W o gamma, colsum_c = sum_k B[c,k], b' = b + W beta   (replaces nn.LayerNorm -> nn.Linear,
//           lib/tgt/layers/triplet.py:207-211, layers.py:49-52,112-113,156-157)
//   BIAS  : + bias_c
//   GELU  : exact-erf GELU followed by counter-hash dropout (layers.py:157-158); optionally also stores the
//           pre-activation U (saved for backward)
//   RES   : D = res + scale[row / rows_per_scale] * v   (DropPath + residual add, layers.py:163-177, 269-290)
#include "ptx.cuh"

namespace tgt {

// K-major operand tile with 128-byte rows, 128B swizzle: 8-row groups are 1024 B apart (SBO), LBO unused.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor: fp32 accumulate, A/B K-major, M = 128
template <typename T> __device__ __forceinline__ uint32_t umma_idesc(int n) {
  const uint32_t fmt = (sizeof(T) == 2 && DT<T>::code == TGT_BF16) ? 1u : 0u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------- kernel
enum : int { EPI_LN = 1, EPI_BIAS = 2, EPI_GELU = 4, EPI_RES = 8, EPI_STORE_U = 16, EPI_RES_F32 = 32, EPI_ROWSCALE = 64,
              EPI_GELU_BWD = 128, EPI_STATS = 256, EPI_LN_BWD = 512, EPI_STORE_GP = 1024, EPI_MULRES = 2048 };

struct GemmParams {
  int64_t M;
  int N, K, bn, n_slices, ctas_per_slice, kblocks, stages, tmem_cols;
  void *D;
  int64_t ldd;
  void *U;
  int64_t ldu;
  const float *row_mean, *row_rstd, *col_sum, *bias, *row_scale;
  const void *res;
  int64_t ldres;
  int rows_per_scale;
  float p_drop;
  unsigned long long seed;
  const unsigned long long *seed_ptr;   // device-resident seed (CUDA-graph replays), overrides `seed` when non-null
  int flags;
  float *stat_mean, *stat_rstd;      // EPI_STATS: LayerNorm statistics of the OUTPUT rows (needs one column slice)
  float stat_eps;
  int debug;                         // TGT_GEMM_DEBUG (profiling only): 1 = skip the bulk stores, 2 = skip the MMAs, 4 = skip the accumulator reads, 8 = skip the activation loads
  const void *res2;                  // EPI_LN_BWD: residual gradient added to dx (16-bit, pitch ldres2); `res` is x
  int64_t ldres2;
};

// warp 0: TMA producer, warp 1: MMA issuer, warps 2..: epilogue, WQ per TMEM lane quadrant (they interleave 32-column
// groups).  WQ = 2 for the epilogues with coalesced-domain work, up to 4 for the TMA-store epilogues, whose per-group
// chain (tcgen05.ld -> FMA -> st.shared -> bulk store) is latency- and not issue-bound.
constexpr int A_STAGE_BYTES = 128 * 128;
// Epilogues whose work is all in the TMEM domain (lane = row) hand their 32x32 output tiles to the TMA store engine: no
// LDS / STG / address arithmetic in the epilogue warps, two staging tiles per warp so a store drains while the next fills.
__host__ __device__ constexpr bool epi_tma_store(int flags) { return (flags & (8 | 128 | 2048 | 256 | 4 | 512)) == 0; }
__host__ __device__ constexpr int epi_stage_bytes(int flags, int wq) {
  return 4 * wq * (((flags & 16) || epi_tma_store(flags)) ? 4096 : 2048);
}

// GELU(u) = u * Phi(u) with Phi from the Abramowitz-Stegun 7.1.26 erfc approximation (|error| < 2e-7 in Phi): two MUFU
// ops and ~10 FMA-pipe instructions per element -- the epilogue is issue-bound, CUDA's erff costs 3x as much.
__device__ __forceinline__ float gelu_fast(float u) {
  const float ax = fabsf(u) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  const float h = 0.5f * poly * t * e;          // 0.5 * erfc(|u| / sqrt 2)
  return u * (u < 0.f ? h : 1.f - h);
}
// d/du GELU(u) = Phi(u) + u * phi(u); the Gaussian exp(-u^2/2) is the one the erfc approximation already needs
__device__ __forceinline__ float gelu_grad_fast(float u) {
  const float ax = fabsf(u) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  const float h = 0.5f * poly * t * e;
  return fmaf(u * 0.39894228040143268f, e, u < 0.f ? h : 1.f - h);
}

// GELU(u) and d/du GELU(u) from one evaluation of the erfc approximation (the forward epilogue with EPI_STORE_GP stores
// the derivative * dropout mask instead of the pre-activation, so the backward epilogue is a plain multiply)
__device__ __forceinline__ void gelu_both_fast(float u, float &act, float &grad) {
  const float ax = fabsf(u) * 0.70710678118654752f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  const float h = 0.5f * poly * t * e;
  const float cdf = u < 0.f ? h : 1.f - h;
  act = u * cdf;
  grad = fmaf(u * 0.39894228040143268f, e, cdf);
}

template <typename T> __device__ __forceinline__ uint32_t pack2(float a, float b);
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__half>(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&h);
}
template <typename T> __device__ __forceinline__ void unpack2(uint32_t w, float &a, float &b);
template <> __device__ __forceinline__ void unpack2<__nv_bfloat16>(uint32_t w, float &a, float &b) {
  a = __uint_as_float(w << 16);
  b = __uint_as_float(w & 0xFFFF0000u);
}
template <> __device__ __forceinline__ void unpack2<__half>(uint32_t w, float &a, float &b) {
  float2 f = __half22float2(*reinterpret_cast<__half2 *>(&w));
  a = f.x;
  b = f.y;
}

// 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 w) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 w;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w) : "r"(addr) : "memory");
  return w;
}
// staging tile of one epilogue warp: 32 rows x 4 pieces of 16 B (32 output columns), xor-swizzled so that both the
// row-per-lane writes and the 8-rows-x-64-B coalesced reads are bank-conflict free
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, int row, int piece) {
  return base + row * 64 + ((piece ^ ((row >> 1) & 3)) << 4);
}

template <typename T, int FLAGS, int WQ>
__global__ void __launch_bounds__(64 + WQ * 128, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmD, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t *smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // warp index through a shuffle: provably warp-uniform, so tile / column-group bookkeeping and the TMA operands stay in
  // uniform registers
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;

  const int slice = blockIdx.x % p.n_slices;
  const int q = blockIdx.x / p.n_slices;
  const int64_t num_m_tiles = (p.M + 127) / 128;
  const int bn = p.bn;

  // shared memory carve-up (all operand regions 1024-byte aligned)
  const uint32_t sB = smem_base;                                           // kblocks x [bn rows x 128 B]
  const uint32_t sA = sB + (uint32_t)p.kblocks * bn * 128;                 // stages x 16 KB
  const uint32_t sStage = sA + (uint32_t)p.stages * A_STAGE_BYTES;         // 8 warps x (D 2 KB [+ U 2 KB])
  static_assert(WQ == 2 || epi_tma_store(FLAGS), "the pair exchanges of the STATS / LN_BWD epilogues assume two warps per quadrant");
  constexpr int EPI_WARPS = 4 * WQ;
  const uint32_t sVec = sStage + epi_stage_bytes(FLAGS, WQ);                         // colsum[256], bias[256] fp32
  const uint32_t sPart = sVec + 2048;                                      // EPI_STATS: [2 buffers][8 warps][32 rows] float2
  const uint32_t sBar = sPart + 4096;                                      // mbarriers
  float2 *part = reinterpret_cast<float2 *>(smem_gen + (sPart - smem_base));
  float *vec_colsum = reinterpret_cast<float *>(smem_gen + (sVec - smem_base));
  float *vec_bias = vec_colsum + 256;
  const uint32_t bar_full = sBar, bar_empty = sBar + 8 * 8, bar_bfull = sBar + 16 * 8;
  const uint32_t bar_tfull = sBar + 17 * 8, bar_tempty = sBar + 19 * 8;
  const uint32_t tmem_slot = sBar + 21 * 8;
  volatile uint32_t *tmem_slot_gen = reinterpret_cast<volatile uint32_t *>(smem_gen + (tmem_slot - smem_base));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(bar_full + i * 8, 1);
      mbar_init(bar_empty + i * 8, 1);
    }
    mbar_init(bar_bfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_tfull + i * 8, 1);
      mbar_init(bar_tempty + i * 8, EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tc_alloc(tmem_slot, (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  if (q < p.ctas_per_slice) {
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer
      // The whole warp runs the loop (uniform control flow, addresses in uniform registers) and one elected lane issues:
      // a `lane == 0` branch makes every operand a per-thread value that the compiler has to move to the uniform
      // datapath with a vote loop around each TMA / MMA / commit instruction.
      if (elect_one()) {
        mbar_expect_tx(bar_bfull, (uint32_t)p.kblocks * bn * 128);
        for (int kb = 0; kb < p.kblocks; ++kb) tma_load_2d(&tmB, bar_bfull, sB + kb * bn * 128, kb * 64, slice * bn);
      }
      __syncwarp();
      int stage = 0;
      uint32_t phase = 0;
      const bool no_loads = (p.debug & 8) != 0;
      for (int64_t mt = q; mt < num_m_tiles; mt += p.ctas_per_slice) {
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(bar_empty + stage * 8, phase ^ 1);
          if (elect_one()) {
            if (no_loads) {
              mbar_arrive(bar_full + stage * 8);
            } else {
              mbar_expect_tx(bar_full + stage * 8, A_STAGE_BYTES);
              tma_load_2d(&tmA, bar_full + stage * 8, sA + stage * A_STAGE_BYTES, kb * 64, (int)(mt * 128));
            }
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer (one elected lane, see above)
      // The issuing thread is the critical resource of small-K GEMMs: a 64-wide K block is 4 instructions of 120 tensor
      // cycles each, so everything it does between two blocks (barrier poll, descriptors, commit) must stay well under
      // 480 cycles.  Descriptors are built once per block and advanced by 2 (= 32 bytes >> 4) per K step.
      const uint32_t idesc = umma_idesc<T>(bn);
      const bool do_mma = !(p.debug & 2);
      const int last_ksteps = (p.K - (p.kblocks - 1) * 64 + 15) / 16;      // K steps of the last (possibly partial) block
      mbar_wait(bar_bfull, 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int64_t mt = q; mt < num_m_tiles; mt += p.ctas_per_slice) {
        mbar_wait(bar_tempty + acc * 8, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * bn);
        for (int kb = 0; kb < p.kblocks; ++kb) {
          mbar_wait(bar_full + stage * 8, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t ad = umma_desc_sw128(sA + stage * A_STAGE_BYTES), bd = umma_desc_sw128(sB + kb * bn * 128);
            const int ksteps = kb + 1 < p.kblocks ? 4 : last_ksteps;
            if (do_mma) {
              tc_mma(tmem_d, ad, bd, idesc, (uint32_t)(kb != 0));
              if (ksteps > 1) tc_mma(tmem_d, ad + 2, bd + 2, idesc, 1u);
              if (ksteps > 2) tc_mma(tmem_d, ad + 4, bd + 4, idesc, 1u);
              if (ksteps > 3) tc_mma(tmem_d, ad + 6, bd + 6, idesc, 1u);
            }
            tc_commit(bar_empty + stage * 8);
            if (kb + 1 == p.kblocks) tc_commit(bar_tfull + acc * 8);
          }
          __syncwarp();
          if (++stage == p.stages) {
            stage = 0;
            phase ^= 1;
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    } else {
      // ------------------------------------------------------------------ epilogue warps (TMEM lane quadrant = warp % 4)
      const int quad = warp & 3;
      const int half = (warp - 2) >> 2;  // which 32-column groups this warp drains: half, half + 2, ...
      const int et = threadIdx.x - 64;   // 0..255
      const int col0 = slice * bn;
      if (et < 256) {
        const int gc = col0 + et;
        vec_colsum[et] = ((FLAGS & (EPI_LN | EPI_LN_BWD)) && et < bn && gc < p.N) ? p.col_sum[gc] : 0.f;
        vec_bias[et] = ((FLAGS & EPI_BIAS) && et < bn && gc < p.N) ? p.bias[gc] : 0.f;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(WQ * 128) : "memory");
      const DropCfg dc = make_drop_cfg(p.p_drop, p.seed_ptr ? *p.seed_ptr : p.seed);
      constexpr bool TMA_ST = epi_tma_store(FLAGS);
      constexpr uint32_t STAGE_PER_WARP = ((FLAGS & EPI_STORE_U) || TMA_ST) ? 4096u : 2048u;
      const uint32_t stD0 = sStage + (uint32_t)(warp - 2) * STAGE_PER_WARP, stU = stD0 + 2048;
      uint32_t st_buf = 0;               // TMA_ST: which of the two staging tiles the next group fills
      if (TMA_ST && lane == 0) tma_prefetch_desc(&tmD);
      const int crow = lane >> 2, cpiece = lane & 3;     // coalesced phase: 8 rows x 4 pieces per instruction
      constexpr int RW = (FLAGS & EPI_RES_F32) ? 2 : 1;
      constexpr bool LNB = (FLAGS & EPI_LN_BWD) != 0;
      T *Dp = reinterpret_cast<T *>(p.D);
      int acc = 0;
      uint32_t acc_phase = 0;
      // per-row epilogue inputs (LN statistics, DropPath scale) are fetched one tile ahead
      float nx_rstd = 1.f, nx_mean = 0.f, nx_rs = 1.f;
      auto row_inputs = [&](int64_t mt_) {
        const int64_t r_ = mt_ * 128 + quad * 32 + lane;
        nx_rstd = 1.f, nx_mean = 0.f, nx_rs = 1.f;
        if (mt_ < num_m_tiles && r_ < p.M) {
          if (FLAGS & (EPI_LN | EPI_LN_BWD)) {
            nx_rstd = p.row_rstd[r_];
            nx_mean = p.row_mean[r_];
          }
          if ((FLAGS & (EPI_RES | EPI_ROWSCALE)) && p.row_scale) nx_rs = p.row_scale[r_ / p.rows_per_scale];
        }
      };
      row_inputs(q);
      for (int64_t mt = q; mt < num_m_tiles; mt += p.ctas_per_slice) {
        const int64_t row0 = mt * 128 + quad * 32;
        const int64_t row = row0 + lane;
        const float rstd = nx_rstd, nmr = -nx_mean * nx_rstd;
        const float rs_lane = nx_rs;         // DropPath scale of this lane's row (TMEM domain)
        row_inputs(mt + p.ctas_per_slice);
        // residual values of the coalesced phase are fetched one 32-column group ahead
        uint4 rcur[4][RW], rnext[4][RW];
        // second operand of the coalesced phase: the residual (RES), the pre-activation (GELU_BWD) or, with LN_BWD, the
        // residual gradient res2 (there `res` is x and is consumed in the TMEM domain)
        const void *cres = LNB ? p.res2 : p.res;
        const int64_t ldc = LNB ? p.ldres2 : p.ldres;
        auto res_load = [&](int c, uint4(&rb)[4][RW]) {
          const int gc = col0 + c + cpiece * 8;
          if (c + cpiece * 8 < bn && gc < p.N) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int64_t grow = row0 + it * 8 + crow;
              if (grow < p.M) {
                if (FLAGS & EPI_RES_F32) {
                  const float *rp = reinterpret_cast<const float *>(cres) + grow * ldc + gc;
                  rb[it][0] = *reinterpret_cast<const uint4 *>(rp);
                  rb[it][RW - 1] = *reinterpret_cast<const uint4 *>(rp + 4);
                } else {
                  rb[it][0] = *reinterpret_cast<const uint4 *>(reinterpret_cast<const T *>(cres) + grow * ldc + gc);
                }
              }
            }
          }
        };
        if (FLAGS & (EPI_RES | EPI_GELU_BWD | EPI_MULRES)) res_load(half * 32, rcur);
        float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};   // EPI_STATS: sum / sum of squares per row
        // LN_BWD: this lane's row of x for the (up to 4) column groups of this warp, fetched before the accumulator wait
        uint32_t xh[LNB ? 4 : 1][16];
        if (LNB) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = half * 32 + k * 64;
#pragma unroll
            for (int j = 0; j < 16; ++j) xh[LNB ? k : 0][j] = 0u;
            if (c < bn && row < p.M) {
              const uint4 *xp = reinterpret_cast<const uint4 *>(reinterpret_cast<const T *>(p.res) + row * p.ldres + col0 + c);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 t = xp[j];
                xh[LNB ? k : 0][4 * j + 0] = t.x; xh[LNB ? k : 0][4 * j + 1] = t.y;
                xh[LNB ? k : 0][4 * j + 2] = t.z; xh[LNB ? k : 0][4 * j + 3] = t.w;
              }
            }
          }
        }
        mbar_wait(bar_tfull + acc * 8, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * bn);
        uint32_t v[32], v2[32];
        float lnb_s1 = 0.f, lnb_s2 = 0.f;     // LN_BWD: mean_c(g), mean_c(g * xhat) of this lane's row, g = dy * gamma
        if (LNB) {
          // pass 1 over the accumulators: row sums; pass 2 (the main loop below) re-reads them from tensor memory
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int c = half * 32 + k * 64;
            if (c < bn) {
              tc_ld32(taddr + c, v);
              tc_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                float x0, x1;
                unpack2<T>(xh[LNB ? k : 0][i], x0, x1);
                const float g0 = __uint_as_float(v[2 * i]) * vec_colsum[c + 2 * i];
                const float g1 = __uint_as_float(v[2 * i + 1]) * vec_colsum[c + 2 * i + 1];
                lnb_s1 += g0 + g1;
                lnb_s2 = fmaf(g0, fmaf(x0, rstd, nmr), fmaf(g1, fmaf(x1, rstd, nmr), lnb_s2));
              }
            }
          }
          float2 *mine = part + (acc * EPI_WARPS + (warp - 2)) * 32;
          mine[lane] = make_float2(lnb_s1, lnb_s2);
          asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
          const float2 o = part[(acc * EPI_WARPS + ((warp - 2) ^ 4)) * 32 + lane];
          const float inv_n = 1.f / (float)p.N;
          lnb_s1 = (lnb_s1 + o.x) * inv_n;
          lnb_s2 = (lnb_s2 + o.y) * inv_n;
        }
        if (half * 32 < bn && !(p.debug & 4)) tc_ld32(taddr + half * 32, v);
        constexpr int CSTRIDE = WQ * 32;                     // column distance between two groups of one warp
        // one 32-column group: accumulators in vc; with the TMA-store epilogues the next group's accumulators are
        // requested into vn as soon as vc has landed, so the tcgen05.ld latency hides behind this group's arithmetic
        auto group = [&](uint32_t(&v)[32], uint32_t(&vn)[32], const int c, const int kq) {
          // per-column epilogue vectors of this group (shared-memory broadcasts), issued before the TMEM wait
          float4 csv[8], bsv[8];
          if (FLAGS & EPI_LN) {
#pragma unroll
            for (int k = 0; k < 8; ++k) csv[k] = *reinterpret_cast<const float4 *>(vec_colsum + c + k * 4);
          }
          if (FLAGS & EPI_BIAS) {
#pragma unroll
            for (int k = 0; k < 8; ++k) bsv[k] = *reinterpret_cast<const float4 *>(vec_bias + c + k * 4);
          }
          if ((FLAGS & (EPI_RES | EPI_GELU_BWD | EPI_MULRES)) && c + 64 < bn) res_load(c + 64, rnext);
          const bool tma_group = TMA_ST && c + 32 <= bn;      // a slice's ragged last group takes the vector-store path
          const uint32_t stD = stD0 + (TMA_ST ? st_buf * 2048u : 0u);
          if (TMA_ST) {
            // the bulk store committed two groups ago has finished reading this staging tile
            if (elect_one()) tma_store_wait_read1();
            __syncwarp();
            st_buf ^= 1u;
          }
          tc_ld_wait();
          if (TMA_ST && c + CSTRIDE < bn && !(p.debug & 4)) tc_ld32(taddr + c + CSTRIDE, vn);
          // ---- TMEM domain (lane = row): LN fold, bias, activation; 16-bit results go to the staging tile
#pragma unroll
          for (int pc = 0; pc < 4; ++pc) {
            float f[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v[pc * 8 + i]);
            if (LNB) {
              // dx = rstd * (g - mean(g) - xhat * mean(g xhat)),  g = dy * gamma,  xhat = (x - mean) * rstd
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                float x0, x1;
                unpack2<T>(xh[LNB ? kq : 0][pc * 4 + j], x0, x1);
                const float g0 = f[2 * j] * vec_colsum[c + pc * 8 + 2 * j];
                const float g1 = f[2 * j + 1] * vec_colsum[c + pc * 8 + 2 * j + 1];
                f[2 * j] = rstd * (g0 - lnb_s1 - fmaf(x0, rstd, nmr) * lnb_s2);
                f[2 * j + 1] = rstd * (g1 - lnb_s1 - fmaf(x1, rstd, nmr) * lnb_s2);
              }
            }
            // packed fp32 pairs (fma.rn.f32x2: two columns per issue slot, bit-identical to the scalar FMAs)
            // (measured, scripts/bench_gemm_flavors.py: the issue-bound GELU epilogues are 12 % faster with scalar FMAs --
            //  their values are consumed one by one -- everything else gains from the packed form)
            constexpr bool PACKED_LN = !(FLAGS & EPI_GELU);
            if ((FLAGS & EPI_LN) && !PACKED_LN) {
              const float4 cq[2] = {csv[2 * pc], csv[2 * pc + 1]};
              const float cs[8] = {cq[0].x, cq[0].y, cq[0].z, cq[0].w, cq[1].x, cq[1].y, cq[1].z, cq[1].w};
              if (FLAGS & EPI_BIAS) {
                const float4 bq[2] = {bsv[2 * pc], bsv[2 * pc + 1]};
                const float bs[8] = {bq[0].x, bq[0].y, bq[0].z, bq[0].w, bq[1].x, bq[1].y, bq[1].z, bq[1].w};
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], rstd, fmaf(nmr, cs[i], bs[i]));
              } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = fmaf(f[i], rstd, nmr * cs[i]);
              }
            } else if (FLAGS & EPI_LN) {
              const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(nmr, nmr);
              const float4 cq[2] = {csv[2 * pc], csv[2 * pc + 1]};
              const float2 cs2[4] = {make_float2(cq[0].x, cq[0].y), make_float2(cq[0].z, cq[0].w), make_float2(cq[1].x, cq[1].y),
                                     make_float2(cq[1].z, cq[1].w)};
              float2 t2[4];
              if (FLAGS & EPI_BIAS) {
                const float4 bq[2] = {bsv[2 * pc], bsv[2 * pc + 1]};
                const float2 bs2[4] = {make_float2(bq[0].x, bq[0].y), make_float2(bq[0].z, bq[0].w), make_float2(bq[1].x, bq[1].y),
                                       make_float2(bq[1].z, bq[1].w)};
#pragma unroll
                for (int j = 0; j < 4; ++j) t2[j] = __ffma2_rn(nmr2, cs2[j], bs2[j]);
              } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) t2[j] = __fmul2_rn(nmr2, cs2[j]);
              }
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 r2 = __ffma2_rn(make_float2(f[2 * j], f[2 * j + 1]), rstd2, t2[j]);
                f[2 * j] = r2.x;
                f[2 * j + 1] = r2.y;
              }
            } else if (FLAGS & EPI_BIAS) {
              const float4 bq[2] = {bsv[2 * pc], bsv[2 * pc + 1]};
              const float bs[8] = {bq[0].x, bq[0].y, bq[0].z, bq[0].w, bq[1].x, bq[1].y, bq[1].z, bq[1].w};
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] += bs[i];
            }
            if (FLAGS & (EPI_RES | EPI_ROWSCALE)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) f[i] *= rs_lane;
            }
            if (FLAGS & EPI_GELU) {
              uint4 wu;
              wu.x = pack2<T>(f[0], f[1]);
              wu.y = pack2<T>(f[2], f[3]);
              wu.z = pack2<T>(f[4], f[5]);
              wu.w = pack2<T>(f[6], f[7]);
              if (!(FLAGS & EPI_STORE_GP)) st_shared_v4(stage_addr(stU, lane, pc), wu);
              const uint32_t wr[4] = {wu.x, wu.y, wu.z, wu.w};
              const uint64_t pair0 = (uint64_t)(row * p.ldu + col0 + c + pc * 8) >> 1;
              uint32_t gpw[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                // the activation is computed from the ROUNDED pre-activation: backward recomputes from the stored U
                float u0, u1, m0 = 1.f, m1 = 1.f;
                unpack2<T>(wr[j], u0, u1);
                if (p.p_drop > 0.f) drop_mask_pair(dc, pair0 + j, m0, m1);
                if (FLAGS & EPI_STORE_GP) {
                  float a0, a1, g0, g1;
                  gelu_both_fast(u0, a0, g0);
                  gelu_both_fast(u1, a1, g1);
                  f[2 * j] = a0 * m0;
                  f[2 * j + 1] = a1 * m1;
                  gpw[j] = pack2<T>(g0 * m0, g1 * m1);
                } else {
                  f[2 * j] = gelu_fast(u0) * m0;
                  f[2 * j + 1] = gelu_fast(u1) * m1;
                }
              }
              if (FLAGS & EPI_STORE_GP) st_shared_v4(stage_addr(stU, lane, pc), make_uint4(gpw[0], gpw[1], gpw[2], gpw[3]));
            }
            uint4 w;
            w.x = pack2<T>(f[0], f[1]);
            w.y = pack2<T>(f[2], f[3]);
            w.z = pack2<T>(f[4], f[5]);
            w.w = pack2<T>(f[6], f[7]);
            st_shared_v4(stage_addr(stD, lane, pc), w);
          }
          if (!TMA_ST && c + 64 < bn) tc_ld32(taddr + c + 64, v);      // next group's accumulators fly during the store phase
          if (TMA_ST) fence_proxy_async();
          __syncwarp();
          if (TMA_ST) {
            if (elect_one()) {
              if (tma_group && !(p.debug & 1)) tma_store_2d(&tmD, stD, col0 + c, (int)row0);
              tma_store_commit();                           // (an empty group on the ragged path keeps the 2-deep accounting)
            }
            if (tma_group) return;
          }
          // ---- coalesced domain: each instruction moves 8 rows x 64 contiguous bytes
          const int gc = col0 + c + cpiece * 8;
          uint4 wst[4], ust[4];
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            wst[it] = ld_shared_v4(stage_addr(stD, it * 8 + crow, cpiece));
            if (FLAGS & EPI_STORE_U) ust[it] = ld_shared_v4(stage_addr(stU, it * 8 + crow, cpiece));
          }
          if (c + cpiece * 8 < bn && gc < p.N) {
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int r = it * 8 + crow;
              const int64_t grow = row0 + r;
              if (grow < p.M) {
                uint4 w = wst[it];
                if (FLAGS & EPI_GELU_BWD) {
                  // D = value * GELU'(u) * dropout mask, u = the saved pre-activation (p.res), same hash as forward
                  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
                  const uint32_t uw[4] = {rcur[it][0].x, rcur[it][0].y, rcur[it][0].z, rcur[it][0].w};
                  const uint64_t pair0 = (uint64_t)(grow * p.ldres + gc) >> 1;
                  float f[8];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float u0, u1, m0 = 1.f, m1 = 1.f;
                    unpack2<T>(ww[j], f[2 * j], f[2 * j + 1]);
                    unpack2<T>(uw[j], u0, u1);
                    if (p.p_drop > 0.f) drop_mask_pair(dc, pair0 + j, m0, m1);
                    f[2 * j] *= gelu_grad_fast(u0) * m0;
                    f[2 * j + 1] *= gelu_grad_fast(u1) * m1;
                  }
                  w.x = pack2<T>(f[0], f[1]);
                  w.y = pack2<T>(f[2], f[3]);
                  w.z = pack2<T>(f[4], f[5]);
                  w.w = pack2<T>(f[6], f[7]);
                }
                if (FLAGS & EPI_MULRES) {
                  // D = value * res: backward of GELU + dropout when the forward stored GELU'(u) * mask (EPI_STORE_GP)
                  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
                  const uint32_t gw[4] = {rcur[it][0].x, rcur[it][0].y, rcur[it][0].z, rcur[it][0].w};
                  float f[8];
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float g0, g1;
                    unpack2<T>(ww[j], f[2 * j], f[2 * j + 1]);
                    unpack2<T>(gw[j], g0, g1);
                    f[2 * j] *= g0;
                    f[2 * j + 1] *= g1;
                  }
                  w.x = pack2<T>(f[0], f[1]);
                  w.y = pack2<T>(f[2], f[3]);
                  w.z = pack2<T>(f[4], f[5]);
                  w.w = pack2<T>(f[6], f[7]);
                }
                if (FLAGS & EPI_RES) {
                  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
                  float f[8];
#pragma unroll
                  for (int j = 0; j < 4; ++j) unpack2<T>(ww[j], f[2 * j], f[2 * j + 1]);
                  if (FLAGS & EPI_RES_F32) {
                    const uint4 q0 = rcur[it][0], q1 = rcur[it][RW - 1];
                    const uint32_t rr[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] += __uint_as_float(rr[i]);
                  } else {
                    const uint32_t rw[4] = {rcur[it][0].x, rcur[it][0].y, rcur[it][0].z, rcur[it][0].w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                      float a, b;
                      unpack2<T>(rw[j], a, b);
                      f[2 * j] += a;
                      f[2 * j + 1] += b;
                    }
                  }
                  w.x = pack2<T>(f[0], f[1]);
                  w.y = pack2<T>(f[2], f[3]);
                  w.z = pack2<T>(f[4], f[5]);
                  w.w = pack2<T>(f[6], f[7]);
                }
                *reinterpret_cast<uint4 *>(Dp + grow * p.ldd + gc) = w;
                if (FLAGS & EPI_STATS) {          // statistics of the values as stored (16-bit rounded)
                  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    float a, b;
                    unpack2<T>(ww[j], a, b);
                    st1[it] += a + b;
                    st2[it] = fmaf(a, a, fmaf(b, b, st2[it]));
                  }
                }
                if (FLAGS & EPI_STORE_U) *reinterpret_cast<uint4 *>(reinterpret_cast<T *>(p.U) + grow * p.ldu + gc) = ust[it];
              }
            }
          }
          __syncwarp();
          if (FLAGS & (EPI_RES | EPI_GELU_BWD | EPI_MULRES)) {
#pragma unroll
            for (int it = 0; it < 4; ++it)
#pragma unroll
              for (int k = 0; k < RW; ++k) rcur[it][k] = rnext[it][k];
          }
        };
        if constexpr (TMA_ST) {
          for (int c = half * 32; c < bn; c += 2 * CSTRIDE) {
            group(v, v2, c, 0);
            if (c + CSTRIDE < bn) group(v2, v, c + CSTRIDE, 0);
          }
        } else {
#pragma unroll
          for (int kq = 0; kq < (LNB ? 4 : 1); ++kq) {
#pragma unroll((FLAGS & EPI_GELU) ? 1 : 2)      // two groups in flight hide the residual loads; the GELU bodies are too big
            for (int c = half * 32 + kq * 64; c < bn; c += (LNB ? (1 << 20) : 64)) group(v, v, c, kq);
          }
        }
        tc_fence_before();
        if (lane == 0) mbar_arrive(bar_tempty + acc * 8);
        if (FLAGS & EPI_STATS) {
          // row sums: 4 lanes (pieces) per row, then the two warps of this TMEM quadrant (even / odd column groups)
#pragma unroll
          for (int it = 0; it < 4; ++it) {
            st1[it] += __shfl_xor_sync(0xffffffffu, st1[it], 1);
            st1[it] += __shfl_xor_sync(0xffffffffu, st1[it], 2);
            st2[it] += __shfl_xor_sync(0xffffffffu, st2[it], 1);
            st2[it] += __shfl_xor_sync(0xffffffffu, st2[it], 2);
          }
          float2 *mine = part + (acc * EPI_WARPS + (warp - 2)) * 32;
          if (cpiece == 0) {
#pragma unroll
            for (int it = 0; it < 4; ++it) mine[it * 8 + crow] = make_float2(st1[it], st2[it]);
          }
          asm volatile("bar.sync %0, 64;" ::"r"(2 + quad) : "memory");
          if (half == 0 && cpiece == 0) {
            const float2 *other = part + (acc * EPI_WARPS + (warp - 2) + 4) * 32;
            const float inv_n = 1.f / (float)p.N;
#pragma unroll
            for (int it = 0; it < 4; ++it) {
              const int64_t grow = row0 + it * 8 + crow;
              if (grow < p.M) {
                const float2 o = other[it * 8 + crow];
                if (p.n_slices > 1) {
                  // the row spans several column slices (CTAs): emit this slice's (sum, sum of squares); stat_mean is a
                  // [n_slices][M] float2 buffer here and stats_finalize_kernel turns the partials into mean / rstd
                  reinterpret_cast<float2 *>(p.stat_mean)[(int64_t)slice * p.M + grow] = make_float2(st1[it] + o.x, st2[it] + o.y);
                } else {
                  const float mu = (st1[it] + o.x) * inv_n;
                  const float var = fmaxf((st2[it] + o.y) * inv_n - mu * mu, 0.f);
                  p.stat_mean[grow] = mu;
                  p.stat_rstd[grow] = rsqrtf(var + p.stat_eps);
                }
              }
            }
          }
        }
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
      if (TMA_ST) {
        __syncwarp();
        if (elect_one()) tma_store_wait_all();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tc_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// (sum, sum of squares) partials of the column slices of a multi-slice TGT_EPI_STATS GEMM -> mean / rstd
__global__ void __launch_bounds__(256) stats_finalize_kernel(const float2 *__restrict__ part, int n_slices, int64_t M, float inv_n,
                                                             float eps, float *__restrict__ mean, float *__restrict__ rstd) {
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < M; r += (int64_t)gridDim.x * blockDim.x) {
    float s1 = 0.f, s2 = 0.f;
    for (int s = 0; s < n_slices; ++s) {
      const float2 v = part[(int64_t)s * M + r];
      s1 += v.x;
      s2 += v.y;
    }
    const float mu = s1 * inv_n;
    mean[r] = mu;
    rstd[r] = rsqrtf(fmaxf(s2 * inv_n - mu * mu, 0.f) + eps);
  }
}

// per-row mean / rstd of x:[rows, W]: one warp per ROWS consecutive rows, ITS 16-byte vectors per lane and row
// (all loads of an iteration are issued before the first reduction, so each warp keeps ROWS*ITS*512 B in flight)
template <typename T, int ROWS, int ITS>
__global__ void __launch_bounds__(256) row_stats_kernel(const T *__restrict__ x, float *__restrict__ mean,
                                                        float *__restrict__ rstd, int64_t rows, int W, int64_t ldx,
                                                        float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  constexpr int NV = 16 / sizeof(T);
  for (int64_t r0 = warp0 * ROWS; r0 < rows; r0 += nwarps * ROWS) {
    float v[ROWS][ITS][NV];
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
#pragma unroll
      for (int it = 0; it < ITS; ++it) {
        const int c = (it * 32 + lane) * NV;
        if (r0 + rr < rows && c < W) {
          load_vec<T, NV>(x + (r0 + rr) * ldx + c, v[rr][it]);
        } else {
#pragma unroll
          for (int i = 0; i < NV; ++i) v[rr][it][i] = 0.f;
        }
      }
    }
#pragma unroll
    for (int rr = 0; rr < ROWS; ++rr) {
      float s = 0.f;
#pragma unroll
      for (int it = 0; it < ITS; ++it)
#pragma unroll
        for (int i = 0; i < NV; ++i) s += v[rr][it][i];
      s = warp_sum(s);
      const float mu = s / (float)W;
      float qd = 0.f;
#pragma unroll
      for (int it = 0; it < ITS; ++it) {
        const int c = (it * 32 + lane) * NV;
        if (c < W) {
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const float d = v[rr][it][i] - mu;
            qd += d * d;
          }
        }
      }
      qd = warp_sum(qd);
      if (lane == 0 && r0 + rr < rows) {
        mean[r0 + rr] = mu;
        rstd[r0 + rr] = rsqrtf(qd / (float)W + eps);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------- host side
// 2-D K-major operand map: rows x K elements of 2 bytes, row pitch ld elements; box = 64 columns x box_rows, 128B swizzle
static int make_operand_map(CUtensorMap *map, const void *base, int64_t rows, int K, int64_t ld, int box_rows,
                            int dtype) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld * 2) % 16)
    return fail("gemm_tc: operand base / pitch must be 16-byte aligned");
  const cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64u, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapDataType dt = dtype == TGT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("gemm_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// output map of the TMA-store epilogues: 32 columns x 32 rows per box, 64B swizzle (= stage_addr's xor pattern)
static int make_store_map(CUtensorMap *map, const void *base, int64_t rows, int N, int64_t ld, int dtype) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail("gemm_tc: cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {32u, 32u};
  const cuuint32_t estr[2] = {1u, 1u};
  const CUtensorMapDataType dt = dtype == TGT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUresult r = enc(map, dt, 2, const_cast<void *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("gemm_tc: cuTensorMapEncodeTiled (store map) failed (%d)", (int)r);
  return 0;
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

// epilogue warps per TMEM lane quadrant.  Measured (scripts/bench_gemm_flavors.py, r2): 3 or 4 warps per quadrant lose --
// their staging tiles cost activation-ring stages (232 KB is shared with the weight panel) -- so every flavour runs 2.
static constexpr int epi_wq(int) { return 2; }

static int gemm_smem_budget(int flags) { return 232448 - 1024 - 2048 - 4096 - 256 - epi_stage_bytes(flags, epi_wq(flags)); }
// Widest column slice whose weight panel leaves room for >= 3 activation stages.  Measured (r2, K = 512, N = 256): trading
// slice width for ring depth loses -- 2 slices x 4 stages 0.43 ms, 3 slices x 6 stages 0.60 ms, 4 slices x 8 stages 0.77 ms
// -- because a narrower UMMA re-reads the activation tile from shared memory once per slice width.
static int gemm_bn_max(int kblocks, int flags) {
  const int budget = gemm_smem_budget(flags);
  const int min_stages = 3;
  int bn_max = 256;
  while (bn_max > 16 && kblocks * bn_max * 128 + min_stages * A_STAGE_BYTES > budget) bn_max -= 16;
  return bn_max;
}

template <typename T, int FLAGS, int WQ>
static int launch_gemm_wq(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &md, const GemmParams &p, size_t smem,
                          cudaStream_t st) {
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm_tc_kernel<T, FLAGS, WQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  });
  gemm_tc_kernel<T, FLAGS, WQ><<<p.n_slices * p.ctas_per_slice, 64 + WQ * 128, smem, st>>>(ma, mb, md, p);
  return check_launch("gemm_tc");
}

template <typename T, int FLAGS>
static int launch_gemm(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &md, const GemmParams &p, size_t smem,
                       cudaStream_t st) {
  return launch_gemm_wq<T, FLAGS, 2>(ma, mb, md, p, smem, st);
}

template <typename T>
static int dispatch_gemm(const CUtensorMap &ma, const CUtensorMap &mb, const CUtensorMap &md, const GemmParams &p, size_t smem,
                         cudaStream_t st) {
  switch (p.flags) {
#define TGT_GEMM_CASE(F) \
  case (F): return launch_gemm<T, (F)>(ma, mb, md, p, smem, st);
    TGT_GEMM_CASE(0)
    TGT_GEMM_CASE(EPI_BIAS)
    TGT_GEMM_CASE(EPI_LN | EPI_BIAS)
    TGT_GEMM_CASE(EPI_LN | EPI_BIAS | EPI_GELU | EPI_STORE_U)
    TGT_GEMM_CASE(EPI_BIAS | EPI_GELU | EPI_STORE_U)
    TGT_GEMM_CASE(EPI_BIAS | EPI_RES)
    TGT_GEMM_CASE(EPI_BIAS | EPI_RES | EPI_RES_F32)
    TGT_GEMM_CASE(EPI_RES)
    TGT_GEMM_CASE(EPI_BIAS | EPI_RES | EPI_STATS)
    TGT_GEMM_CASE(EPI_BIAS | EPI_RES | EPI_RES_F32 | EPI_STATS)
    TGT_GEMM_CASE(EPI_ROWSCALE)
    TGT_GEMM_CASE(EPI_GELU_BWD)
    TGT_GEMM_CASE(EPI_GELU_BWD | EPI_ROWSCALE)
    TGT_GEMM_CASE(EPI_LN | EPI_BIAS | EPI_GELU | EPI_STORE_U | EPI_STORE_GP)
    TGT_GEMM_CASE(EPI_BIAS | EPI_GELU | EPI_STORE_U | EPI_STORE_GP)
    TGT_GEMM_CASE(EPI_MULRES)
    TGT_GEMM_CASE(EPI_MULRES | EPI_ROWSCALE)
    TGT_GEMM_CASE(EPI_LN_BWD)
    TGT_GEMM_CASE(EPI_LN_BWD | EPI_RES)
#undef TGT_GEMM_CASE
    default: return fail("gemm_tc: unsupported epilogue flag combination %d", p.flags);
  }
}

}  // namespace tgt

using namespace tgt;

extern "C" int tgt_gemm_tc(const tgt_gemm_desc *g, const void *A, const void *B, void *D, void *stream) {
  if (!g || !A || !B || !D) return fail("gemm_tc: null argument");
  if (g->dtype != TGT_BF16 && g->dtype != TGT_F16) return fail("gemm_tc: operands must be bf16 or fp16");
  if (g->M <= 0 || g->N <= 0 || g->K <= 0) return fail("gemm_tc: bad shape M=%lld N=%d K=%d", (long long)g->M, g->N, g->K);
  if (g->M >= (1ll << 31) - 128) return fail("gemm_tc: M too large for a 32-bit TMA coordinate");
  if (g->K % 8 || g->N % 8 || g->ldd % 8) return fail("gemm_tc: K, N and ldd must be multiples of 8");
  if (g->K > 512) return fail("gemm_tc: K=%d > 512 unsupported (weight panel must fit in shared memory)", g->K);
  if ((reinterpret_cast<uintptr_t>(D) & 15)) return fail("gemm_tc: D must be 16-byte aligned");
  int flags = g->flags;
  if ((flags & EPI_LN) && (!g->row_mean || !g->row_rstd || !g->col_sum)) return fail("gemm_tc: LN epilogue needs row stats and column sums");
  if ((flags & EPI_BIAS) && !g->bias) return fail("gemm_tc: BIAS epilogue without bias");
  if ((flags & EPI_RES) && (!g->res || g->ldres % 8 || (reinterpret_cast<uintptr_t>(g->res) & 15)))
    return fail("gemm_tc: RES epilogue needs a 16-byte aligned residual with ldres %% 8 == 0");
  if ((flags & EPI_GELU_BWD) && ((flags & EPI_RES) || !g->res || g->ldres % 8 || (reinterpret_cast<uintptr_t>(g->res) & 15) ||
                                 g->res_dtype != g->dtype))
    return fail("gemm_tc: GELU_BWD epilogue needs the 16-bit pre-activation in `res` (ldres %% 8 == 0) and excludes RES");
  if ((flags & EPI_MULRES) && ((flags & (EPI_RES | EPI_GELU_BWD)) || !g->res || g->ldres % 8 ||
                               (reinterpret_cast<uintptr_t>(g->res) & 15) || g->res_dtype != g->dtype))
    return fail("gemm_tc: MULRES epilogue needs a 16-bit multiplier in `res` (ldres %% 8 == 0) and excludes RES / GELU_BWD");
  if ((flags & EPI_STORE_GP) && !(flags & EPI_GELU)) return fail("gemm_tc: STORE_GP modifies the GELU epilogue");
  if ((flags & EPI_ROWSCALE) && !g->row_scale) return fail("gemm_tc: ROWSCALE epilogue without row_scale");
  if ((flags & (EPI_RES | EPI_ROWSCALE)) && g->row_scale && g->rows_per_scale <= 0) return fail("gemm_tc: rows_per_scale must be > 0");
  if (flags & EPI_STORE_U) {
    if (!(flags & EPI_GELU) || !g->U || g->ldu % 8 || (reinterpret_cast<uintptr_t>(g->U) & 15))
      return fail("gemm_tc: STORE_U needs GELU and a 16-byte aligned U with ldu %% 8 == 0");
  }
  if ((flags & EPI_GELU) && !(flags & EPI_STORE_U)) return fail("gemm_tc: GELU epilogue is only built with STORE_U");
  if ((flags & (EPI_GELU | EPI_GELU_BWD)) && (g->p_drop < 0.f || g->p_drop >= 1.f)) return fail("gemm_tc: p_drop=%f out of range", g->p_drop);

  GemmParams p{};
  p.M = g->M;
  p.N = g->N;
  p.K = g->K;
  p.kblocks = (g->K + 63) / 64;
  const int smem_budget = gemm_smem_budget(flags);
  const int bn_max = gemm_bn_max(p.kblocks, flags);
  p.n_slices = (g->N + bn_max - 1) / bn_max;
  p.bn = (((g->N + p.n_slices - 1) / p.n_slices) + 15) / 16 * 16;
  const int64_t m_tiles = (g->M + 127) / 128;
  const int sms = num_sms();
  if (p.n_slices > sms) return fail("gemm_tc: N=%d needs more column slices than SMs", g->N);
  p.ctas_per_slice = (int)std::min<int64_t>(sms / p.n_slices, m_tiles);
  p.stages = std::min(8, (smem_budget - p.kblocks * p.bn * 128) / A_STAGE_BYTES);
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.bn + (p.bn % 32)) p.tmem_cols *= 2;     // the epilogue reads 32-column groups
  p.D = D;
  p.ldd = g->ldd;
  p.U = g->U;
  p.ldu = g->ldu;
  p.row_mean = g->row_mean;
  p.row_rstd = g->row_rstd;
  p.col_sum = g->col_sum;
  p.bias = g->bias;
  p.row_scale = g->row_scale;
  p.res = g->res;
  p.ldres = g->ldres;
  p.rows_per_scale = g->rows_per_scale;
  p.p_drop = g->p_drop;
  p.seed = g->seed;
  p.seed_ptr = reinterpret_cast<const unsigned long long *>(g->seed_ptr);
  if (const char *dbg = getenv("TGT_GEMM_DEBUG")) p.debug = atoi(dbg);
  if ((flags & EPI_RES) && !(flags & EPI_LN_BWD) && g->res_dtype == TGT_F32) flags |= EPI_RES_F32;
  else if ((flags & EPI_RES) && g->res_dtype != g->dtype) return fail("gemm_tc: residual dtype must be fp32 or the operand dtype");
  p.flags = flags;
  if (flags & EPI_LN_BWD) {
    if (p.n_slices != 1 || p.bn != g->N || g->N % 64 || !g->res || g->res_dtype != g->dtype || !g->row_mean ||
        !g->row_rstd || !g->col_sum || g->ldres % 8 || (reinterpret_cast<uintptr_t>(g->res) & 15))
      return fail("gemm_tc: LN_BWD epilogue needs N %% 64 == 0, N <= 256, x (16-bit) in `res`, row stats and gamma in col_sum");
    if ((flags & EPI_RES) && (!g->res2 || g->ldres2 % 8 || (reinterpret_cast<uintptr_t>(g->res2) & 15)))
      return fail("gemm_tc: LN_BWD + RES needs the residual gradient in res2");
    p.res2 = g->res2;
    p.ldres2 = g->ldres2;
  }
  if (flags & EPI_STATS) {
    if (!g->stat_mean || !g->stat_rstd) return fail("gemm_tc: STATS epilogue needs the two output vectors");
    if (p.n_slices != 1 && !g->stat_partial)
      return fail("gemm_tc: STATS epilogue over %d column slices needs the stat_partial workspace (%d x M float2)", p.n_slices,
                  p.n_slices);
    p.stat_mean = p.n_slices != 1 ? reinterpret_cast<float *>(g->stat_partial) : g->stat_mean;
    p.stat_rstd = g->stat_rstd;
    p.stat_eps = g->stat_eps;
  }

  CUtensorMap ma, mb, md;
  if (int e = make_operand_map(&ma, A, g->M, g->K, g->lda, 128, g->dtype)) return e;
  if (int e = make_operand_map(&mb, B, g->N, g->K, g->ldb, p.bn, g->dtype)) return e;
  if (epi_tma_store(flags)) {
    if (int e = make_store_map(&md, D, g->M, g->N, g->ldd, g->dtype)) return e;
  } else {
    md = ma;
  }
  const size_t smem = (size_t)p.kblocks * p.bn * 128 + (size_t)p.stages * A_STAGE_BYTES + epi_stage_bytes(flags, epi_wq(flags)) + 2048 + 4096 + 256 + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  const int rc = g->dtype == TGT_BF16 ? dispatch_gemm<__nv_bfloat16>(ma, mb, md, p, smem, st) : dispatch_gemm<__half>(ma, mb, md, p, smem, st);
  if (rc || !(flags & EPI_STATS) || p.n_slices == 1) return rc;
  const int blocks = (int)std::min<int64_t>((g->M + 255) / 256, (int64_t)num_sms() * 8);
  stats_finalize_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float2 *>(g->stat_partial), p.n_slices, g->M,
                                                1.f / (float)g->N, g->stat_eps, g->stat_mean, g->stat_rstd);
  return check_launch("stats_finalize");
}

extern "C" int tgt_gemm_tc_slices(int N, int K, int flags) {
  // number of column slices tgt_gemm_tc cuts N into (callers size the stat_partial workspace with it)
  const int bn_max = gemm_bn_max((K + 63) / 64, flags);
  return (N + bn_max - 1) / bn_max;
}

extern "C" int tgt_row_stats(const void *x, float *mean, float *rstd, int64_t rows, int W, int64_t ldx, float eps,
                             int dtype, void *stream) {
  if (rows <= 0) return 0;
  if (W % 8 || W > 1024 || W <= 0) return fail("row_stats: W=%d must be a multiple of 8 and <= 1024", W);
  if (dtype != TGT_BF16 && dtype != TGT_F16) return fail("row_stats: 16-bit input only");
  if (ldx % 8 || (reinterpret_cast<uintptr_t>(x) & 15)) return fail("row_stats: x must be 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  const int rows_per_warp = W <= 256 ? 4 : 1;
  const int64_t blocks = std::min<int64_t>((rows + 8 * rows_per_warp - 1) / (8 * rows_per_warp), (int64_t)num_sms() * 8);
  TGT_DISPATCH_DTYPE(dtype, T, {
    if constexpr (sizeof(T) == 2) {
      const T *xp = reinterpret_cast<const T *>(x);
      if (W <= 256)
        row_stats_kernel<T, 4, 1><<<(unsigned)blocks, 256, 0, st>>>(xp, mean, rstd, rows, W, ldx, eps);
      else
        row_stats_kernel<T, 1, 4><<<(unsigned)blocks, 256, 0, st>>>(xp, mean, rstd, rows, W, ldx, eps);
    }
  });
  return check_launch("row_stats");
}
