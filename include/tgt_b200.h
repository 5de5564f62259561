/*
 * tgt_b200 C ABI  --  hand-written sm_100a kernels for the TGT hot path.
 *
 * The reference (shamim-hussain/tgt) has no FFI / plugin layer: its boundary is the
 * Python nn.Module surface of lib.tgt (lib/tgt/__init__.py:1, lib/tgt/layers/__init__.py:1).
 * This header is the C-ABI *under* that surface: each entry point replaces the ATen/cuBLAS
 * launches the cited reference lines produce.  tgt_b200/_C.py binds it with ctypes and
 * tgt_b200/ops.py wraps it in torch.autograd.Function objects.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (PyTorch caching allocator);
 *    the library never allocates, frees or retains memory, and never synchronises.
 *  - every call launches asynchronously on `stream` (a cudaStream_t passed as void*).
 *  - return value: 0 = ok; non-zero = error, message via tgt_last_error() (thread local).
 *    Unsupported shapes / dtypes are errors -- there is no fallback path.
 *  - dtype codes: 0 = float32, 1 = bfloat16, 2 = float16.  All accumulation is fp32.
 *  - R = B*N*N edge rows.  "head-major" means channel c = h*d + dd inside a block of H*d
 *    channels (the Python layer permutes the reference's d-major weights once per call).
 */
#ifndef TGT_B200_H
#define TGT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TGT_F32  0
#define TGT_BF16 1
#define TGT_F16  2

/* ---- library ------------------------------------------------------------------------- */
int         tgt_version(void);
const char *tgt_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t    tgt_launch_count(void);
/* 0 = default: the fastest measured kernel per op (triplet attention: tcgen05 / TMEM forward + software-pipelined mma.sync
 * backward); 1 = force the generic SIMT kernels; 3 = like 0 but the triplet-attention forward runs the fused projection +
 * attention kernel (tgt_triplet_attn_fused_fwd); 4 = triplet-attention core on tcgen05 / TMEM forward AND backward
 * (csrc/triplet_tc.cu); 5 (and 2, kept as an alias of the removed cp.async family) = the TMA-staged mma.sync core forward
 * and backward (csrc/triplet_tma.cu).  The tests cross-check the families. */
void        tgt_set_kernel_policy(int policy);

/* optional device-side timing of the MAIN kernel of each call (prep / post helpers excluded): enable,
 * run, then read "name launches total_ms" lines (the read synchronises on the recorded CUDA events;
 * disabling frees them).  Diagnostics only -- this is the one place the library creates CUDA objects. */
void        tgt_kernel_timer_enable(int on);
int         tgt_kernel_timer_read(char *buf, size_t n);

/* ---- LayerNorm over the channel dim of edge rows ---------------------------------------
 * replaces nn.LayerNorm calls at lib/tgt/layers/triplet.py:47,207; layers.py:49,112,156.
 * x:[rows,W] (x_dtype)  y:[rows,ldy] (y_dtype)  gamma,beta:[W] f32  mean,rstd:[rows] f32.
 * ldy >= W: when ldy > W the extra columns are written as [1, 0, 0, ...] ("augmented" LN output:
 * a weight-gradient GEMM dY^T * [LN(x) | 1] then yields the bias gradient in the same pass).   */
int tgt_layernorm_fwd(const void *x, const float *gamma, const float *beta, void *y,
                      float *mean, float *rstd, int64_t rows, int W, int64_t ldy, float eps,
                      int x_dtype, int y_dtype, void *stream);
/* dgamma,dbeta:[W] f32 must be ZERO on entry (accumulated with atomics).
 * dy has y_dtype, dx has x_dtype.  If `dres` is non-null it is added to dx (fused residual
 * gradient: dx = LN'(dy) + dres), dres has x_dtype.                                        */
int tgt_layernorm_bwd(const void *dy, const void *x, const float *gamma, const float *mean,
                      const float *rstd, const void *dres, void *dx, float *dgamma, float *dbeta,
                      int64_t rows, int W, int x_dtype, int y_dtype, void *stream);

/* Same, with by-products of the pass (W = 256 and 16-bit dy only, else an error):
 *  y (or NULL)          : y = LN(x) as [rows, ldy] in y_dtype (ldy = W or W + 8 with the augmentation columns of
 *                         tgt_layernorm_fwd) -- the weight-gradient GEMM of the Linear behind the LayerNorm then needs
 *                         no recompute pass; needs beta;
 *  dres_colsum (or NULL): [W] f32, ZERO on entry: += sum_r w_r * dres[r,:], w_r = res_row_scale[r / rows_per_scale]
 *                         (1 when res_row_scale is NULL) -- dres is the upstream gradient of the module this LayerNorm
 *                         opens, so this is the (DropPath-weighted) bias gradient of the module's output projection.  */
int tgt_layernorm_bwd_y(const void *dy, const void *x, const float *gamma, const float *beta,
                        const float *mean, const float *rstd, const void *dres, void *dx, float *dgamma,
                        float *dbeta, void *y, int64_t ldy, int64_t rows, int W, int x_dtype, int y_dtype,
                        const float *res_row_scale, int64_t rows_per_scale, float *dres_colsum, void *stream);

/* ---- triplet attention core -------------------------------------------------------------
 * replaces lib/tgt/layers/triplet.py:213-227 and 232-246 (both einsums, bias add, masked
 * softmax, sigmoid gate) for TripletAttention, TripletAttentionUngated and AxialAttention.
 * The O(N^3 H) tensor is never written to memory.
 *
 * proj : [R, ld]  row (b,i,j) holds the projections of LN(e)[b,i,j,:]; column blocks
 *        (head-major, H*d wide) at off_q/off_k/off_v[dir], and H-wide bias / gate blocks at
 *        off_e/off_g[dir] (dir 0 = inward, 1 = outward; -1 = block absent -> no bias / gate)
 * mask : [B,N,N] f32 additive (0 or finfo.min; -inf allowed)
 * va   : [R, 2*H*d] out, channel = dir*H*d + h*d + dd
 * stats: [B,2,H,N(j),N(i),2] f32 out: row max m and 1/l of the masked softmax            */
typedef struct {
  int32_t B, N, H, d;
  int64_t ld;
  int32_t off_q[2], off_k[2], off_v[2], off_e[2], off_g[2];
  float   scale;          /* d^-0.5 (triplet.py:190) */
  int32_t dtype;          /* dtype of proj / va / dva / dproj */
} tgt_triplet_attn_desc;

/* workspace: caller-provided scratch of at least tgt_triplet_attn_workspace_bytes(desc, backward)
 * bytes (0 for shapes only the generic kernels support); holds the [B,2,H,64,64] bias / gate
 * tiles (and, backward, their gradients).  Contents are undefined after the call.
 * stats written by the forward are private to the kernel family that wrote them: run the
 * backward under the same kernel policy as the forward.                                     */
size_t tgt_triplet_attn_workspace_bytes(const tgt_triplet_attn_desc *desc, int backward);
int tgt_triplet_attn_fwd(const tgt_triplet_attn_desc *desc, const void *proj, const float *mask,
                         void *va, float *stats, void *workspace, size_t workspace_bytes,
                         void *stream);
/* Test hook for the bf16 parity protocol ("fp32 kernel outputs on 16-bit-rounded inputs within
 * 1e-3 of the fp64 oracle"): the tcgen05 core with its Va written as fp32 [R, 2*H*d] (same channel order) instead
 * of being rounded to 16 bit.  Errors unless the tcgen05 kernel supports desc (d = 16, N <= 64, even H, 16-bit proj). */
int tgt_triplet_attn_fwd_f32out(const tgt_triplet_attn_desc *desc, const void *proj, const float *mask,
                                float *va_f32, float *stats, void *workspace, size_t workspace_bytes,
                                void *stream);
/* dproj: [R, ld] out, same column layout as proj; every column named in desc is written.   */
int tgt_triplet_attn_bwd(const tgt_triplet_attn_desc *desc, const void *proj, const float *mask,
                         const void *va, const void *dva, const float *stats, void *dproj,
                         void *workspace, size_t workspace_bytes, void *stream);

/* Same, reusing the bias / gate tiles of the forward: `fwd_workspace` is the workspace buffer the forward
 * call (tgt_triplet_attn_fwd or tgt_triplet_attn_fused_fwd, same desc geometry, same kernel policy) was
 * given, kept alive and unmodified by the caller; NULL recomputes the tiles from proj.                 */
int tgt_triplet_attn_bwd_tiles(const tgt_triplet_attn_desc *desc, const void *proj, const float *mask,
                               const void *va, const void *dva, const float *stats, void *dproj,
                               void *workspace, size_t workspace_bytes, const void *fwd_workspace,
                               void *stream);

/* ---- fused triplet attention forward: LayerNorm + lin_QKV_in/out + attention in one kernel ----------
 * replaces lib/tgt/layers/triplet.py:207-246 for head dim 16, edge width We <= 256 (TGT-At): the
 * [R, 6*We] q/k/v projection is produced on tcgen05 tensor cores tile by tile in tensor memory and
 * consumed in place -- it never reaches HBM.  Returns 1 from ..._supported when desc / We qualify.
 * x      : [B,N,N,We] raw edge rows (16-bit, row pitch ldx) -- NOT layer-normed
 * row_mean,row_rstd : [R] LayerNorm statistics of x (tgt_row_stats or TGT_EPI_STATS)
 * wf     : [(H/2)*192, We] LN-folded weights W*gamma, 16-bit, rows grouped per pair of heads (h0,h1) as
 *          Qin h0,h1 | Qout h0,h1 | Kout h0,h1 | Vout h0,h1 | Kin h0,h1 | Vin h0,h1 (16 rows each)
 * wcolsum, wbias : [(H/2)*192] f32, per wf row: sum_k wf[r,k] and b + W beta
 * proj_eg: [R, desc->ld] the E|G columns (desc->off_e / off_g index THIS matrix; off_q/k/v are ignored)
 * va, stats, workspace: as tgt_triplet_attn_fwd (the backward is tgt_triplet_attn_bwd on a recomputed
 * projection; stats are interchangeable with the TMA / cp.async tensor-core kernels).                  */
/* Same as ..._bwd_tiles, and -- when dbias is non-NULL -- also accumulates the column sums of dproj into dbias
 * [desc->ld] f32 (ZERO on entry): the bias gradient of the projection (lin_QKV_in/out, lin_EG_in/out) as a by-product
 * of the backward kernel, so the weight-gradient GEMM needs no augmented [LN(x) | 1] operand.  Only the pipelined
 * TMA kernel provides it: ask tgt_triplet_attn_bwd_bias_supported first (1 = yes).                               */
int tgt_triplet_attn_bwd_bias_supported(const tgt_triplet_attn_desc *desc);
int tgt_triplet_attn_bwd_bias(const tgt_triplet_attn_desc *desc, const void *proj, const float *mask,
                              const void *va, const void *dva, const float *stats, void *dproj,
                              void *workspace, size_t workspace_bytes, const void *fwd_workspace,
                              float *dbias, void *stream);

int tgt_triplet_attn_fused_supported(const tgt_triplet_attn_desc *desc, int We);
int tgt_triplet_attn_fused_fwd(const tgt_triplet_attn_desc *desc, const void *x, int64_t ldx, int We,
                               const float *row_mean, const float *row_rstd, const void *wf,
                               const float *wcolsum, const float *wbias, const void *proj_eg,
                               const float *mask, void *va, float *stats, void *workspace,
                               size_t workspace_bytes, void *stream);

/* ---- triplet aggregate core ---------------------------------------------------------------
 * replaces lib/tgt/layers/triplet.py:56-68 (TripletAggregate) and 106-120 (…Ungated).
 * proj column blocks: off_v[dir] (H*d, head-major), off_e[dir], off_g[dir] (H; off_g = -1 ->
 * ungated).  mask_dir[dir] != 0 -> add the mask in that direction (the gated reference masks
 * only the inward direction, triplet.py:56-57 vs 63-64).
 * aw : [B,2,H,N,N] f32 workspace/out, aw[b,dir,h,i,k] = attention weight A (query i, key k) */
typedef struct {
  int32_t B, N, H, d;
  int64_t ld;
  int32_t off_v[2], off_e[2], off_g[2];
  int32_t mask_dir[2];
  int32_t dtype;
} tgt_triplet_aggr_desc;

int tgt_triplet_aggr_fwd(const tgt_triplet_aggr_desc *desc, const void *proj, const float *mask,
                         void *va, float *aw, void *stream);
/* daw: [B,2,H,N,N] f32 scratch (written by the call).                                        */
int tgt_triplet_aggr_bwd(const tgt_triplet_aggr_desc *desc, const void *proj, const float *mask,
                         const void *dva, const float *aw, float *daw, void *dproj, void *stream);

/* ---- triangular update core ("next" row 8f-2) ----------------------------------------------
 * replaces lib/tgt/layers/triplet.py:154-170 (TriangularUpdate: mask add, four siglin gates, two einsums).
 * proj : [R, ld]; lin_V output (V_in_g | V_in_l | V_out_g | V_out_l, H columns each, triplet.py:154) at
 *        column off_v, lin_E output (E_in_g | E_in_l | E_out_g | E_out_l, triplet.py:155) at column off_e.
 * mask : [B,N,N] f32 additive, added to the four gates (triplet.py:157-160)
 * va   : [R, 2H] out = [Va_in | Va_out] (triplet.py:166-169).  N <= 64.
 * bwd  : dproj [R, ld] receives the gradient of all eight blocks (other columns are not touched). */
typedef struct {
  int32_t B, N, H;
  int64_t ld;
  int32_t off_v, off_e;
  int32_t dtype;
} tgt_triangular_desc;

int tgt_triangular_fwd(const tgt_triangular_desc *desc, const void *proj, const float *mask, void *va,
                       void *stream);
int tgt_triangular_bwd(const tgt_triangular_desc *desc, const void *proj, const float *mask,
                       const void *dva, void *dproj, void *stream);
/* y[r,c] = sigmoid(x[r,c]) * x[r,W+c], x:[rows,2W], y:[rows,W] (triplet.py:129-131, 172-175); W % (16/sizeof) == 0 */
int tgt_siglin_fwd(const void *x, void *y, int64_t rows, int W, int dtype, void *stream);
int tgt_siglin_bwd(const void *x, const void *dy, void *dx, int64_t rows, int W, int dtype, void *stream);

/* ---- EGT node/edge attention core ---------------------------------------------------------
 * replaces lib/tgt/layers/layers.py:62-78 (EGT_Attention) and 121-125 (EdgeUpdate).
 * Reference-native d-major channel layout: channel c = dd*H + h.
 * qkv  : [B*N, ld_qkv]; Q at column 0, K at column H*d, V at 2*H*d (absent when attend==0)
 * eg   : [R, ld_eg];    E at column 0, G at column H (absent when attend==0)
 * mask : [B,N,N] f32 ; src_mask: [B,N] f32 additive column mask or NULL (layers.py:55-59)
 * hhat : [R,H] out  (pre-mask logits, layers.py:69)
 * vatt : [B*N, H*d] out (after the degree scaler when scale_degree)  -- attend only
 * stats: [B*N, H, 3] f32 out: m, 1/l, sum of gates                   -- attend only         */
typedef struct {
  int32_t B, N, H, d;
  int64_t ld_qkv, ld_eg;
  float   scale;
  int32_t attend;         /* 1 = EGT_Attention, 0 = EdgeUpdate (logits only) */
  int32_t scale_degree;
  int32_t dtype;
} tgt_egt_desc;

int tgt_egt_attn_fwd(const tgt_egt_desc *desc, const void *qkv, const void *eg, const float *mask,
                     const float *src_mask, void *hhat, void *vatt, float *stats, void *stream);
/* vatt: the forward output (attend only; delta = dvatt . vatt).  dhhat may be NULL
 * (edge_update=False).  dqkv:[B*N, ld_qkv], deg:[R, ld_eg] out.  workspace: at least
 * tgt_egt_attn_workspace_bytes(desc) bytes of caller-provided scratch ([R,H] attention weights). */
size_t tgt_egt_attn_workspace_bytes(const tgt_egt_desc *desc);
int tgt_egt_attn_bwd(const tgt_egt_desc *desc, const void *qkv, const void *eg, const float *mask,
                     const float *src_mask, const float *stats, const void *vatt, const void *dhhat,
                     const void *dvatt, void *dqkv, void *deg, void *workspace, size_t workspace_bytes,
                     void *stream);

/* ---- tcgen05 / TMA projection GEMM with fused epilogues ----------------------------------------
 * D[M,N] = epilogue(A[M,K] * B[N,K]^T); A (pitch lda), B (pitch ldb, the nn.Linear weight layout)
 * and D (pitch ldd) are 16-bit (dtype = TGT_BF16 / TGT_F16), accumulation is fp32 in tensor memory.
 * K <= 512 and K, N, pitches multiples of 8; all base pointers 16-byte aligned.
 * replaces the nn.Linear calls on edge rows: lib/tgt/layers/triplet.py:210-211,229-230,249 (and
 * 50-51,72), layers.py:52,82,113,126,157,159 -- and, through the epilogue flags, the ops around them:
 *   TGT_EPI_LN    nn.LayerNorm in front of the Linear (triplet.py:207, layers.py:49,112,156): the GEMM
 *                 runs on the raw rows with B = W*gamma; out = rstd_r*(acc - mean_r*col_sum_c) (+ bias')
 *                 where col_sum_c = sum_k B[c,k] and bias' = bias + W beta are prepared by the caller
 *   TGT_EPI_BIAS  + bias_c (fp32 vector)
 *   TGT_EPI_GELU  exact GELU then dropout(p_drop, seed) with the same counter hash as
 *                 tgt_gelu_dropout_fwd (layers.py:157-158); requires TGT_EPI_STORE_U: the 16-bit
 *                 pre-activation is also written to U (pitch ldu, must be N for the hash to match)
 *   TGT_EPI_RES   D = res + row_scale[row / rows_per_scale] * value  (DropPath + residual add,
 *                 layers.py:163-177, 269-290); res is 16-bit (dtype) or fp32 (res_dtype), row_scale
 *                 may be NULL (= 1)                                                                  */
#define TGT_EPI_LN      1
#define TGT_EPI_BIAS    2
#define TGT_EPI_GELU    4
#define TGT_EPI_RES     8
#define TGT_EPI_STORE_U 16
#define TGT_EPI_ROWSCALE 64   /* value *= row_scale[row / rows_per_scale] (DropPath in backward), no residual      */
#define TGT_EPI_STATS    256  /* also write LayerNorm statistics (mean, 1/sqrt(var + stat_eps)) of the OUTPUT rows, for the
                               * next LayerNorm-folded GEMM; needs N <= 256                                            */
#define TGT_EPI_LN_BWD   512  /* LayerNorm backward as the epilogue of the data-gradient GEMM dy = A B^T (N = LN width <= 256,
                               * N % 64 == 0): D = rstd*(g - mean(g) - xhat*mean(g*xhat)) [+ res2], g = dy*gamma;
                               * x = `res` (16-bit), gamma = col_sum, row_mean / row_rstd = forward statistics;
                               * with TGT_EPI_RES the residual gradient res2 (16-bit, pitch ldres2) is added          */
#define TGT_EPI_STORE_GP 1024 /* with GELU | STORE_U: U receives GELU'(u) * dropout mask instead of the pre-activation u, so
                               * that the backward epilogue is a plain multiply (TGT_EPI_MULRES)                        */
#define TGT_EPI_MULRES   2048 /* D = value * res (16-bit, pitch ldres): backward of GELU + dropout for a STORE_GP forward;
                               * combines with ROWSCALE, excludes RES / GELU_BWD                                        */
#define TGT_EPI_GELU_BWD 128  /* D = value * GELU'(u) * dropout mask(p_drop, seed); u = `res` (16-bit, pitch ldres) */
typedef struct {
  int64_t M;
  int32_t N, K;
  int64_t lda, ldb, ldd;
  int32_t dtype, flags;
  const float *row_mean, *row_rstd, *col_sum, *bias, *row_scale;
  const void *res;
  int64_t ldres;
  int32_t res_dtype, rows_per_scale;
  void *U;
  int64_t ldu;
  float p_drop;
  uint64_t seed;
  float *stat_mean, *stat_rstd;   /* [M] out, TGT_EPI_STATS only */
  float stat_eps;
  const void *res2;               /* TGT_EPI_LN_BWD | TGT_EPI_RES */
  int64_t ldres2;
  void *stat_partial;             /* TGT_EPI_STATS with more than one column slice (tgt_gemm_tc_slices): workspace of
                                   * slices * M * 8 bytes for the per-slice (sum, sum of squares); a tiny second kernel
                                   * then writes stat_mean / stat_rstd                                              */
  const uint64_t *seed_ptr;       /* TGT_EPI_GELU / _GELU_BWD: if non-NULL the dropout seed is READ FROM DEVICE MEMORY when
                                   * the kernel runs (and `seed` is ignored), so that a captured CUDA graph draws a fresh
                                   * mask on every replay (Monte-Carlo-dropout inference, tgt_training.py:42)          */
} tgt_gemm_desc;
int tgt_gemm_tc(const tgt_gemm_desc *desc, const void *A, const void *B, void *D, void *stream);
/* number of column slices tgt_gemm_tc cuts an [M,N] output into for this K and epilogue */
int tgt_gemm_tc_slices(int N, int K, int flags);

/* per-row LayerNorm statistics of x:[rows,W] (16-bit, pitch ldx): mean, rstd = 1/sqrt(var + eps)   */
int tgt_row_stats(const void *x, float *mean, float *rstd, int64_t rows, int W, int64_t ldx, float eps,
                  int dtype, void *stream);

/* ---- FFN activation: y = dropout(gelu(u)) (exact erf GELU) -----------------------------------
 * replaces F.gelu + nn.Dropout at lib/tgt/layers/layers.py:157-158.  The keep mask is a
 * counter-based hash of (seed, element index): nothing is stored, backward regenerates it. */
int tgt_gelu_dropout_fwd(const void *u, void *y, int64_t n, float p_drop, uint64_t seed,
                         int dtype, void *stream);
int tgt_gelu_dropout_bwd(const void *u, const void *dy, void *du, int64_t n, float p_drop,
                         uint64_t seed, int dtype, void *stream);
/* forward only, seed read from device memory at kernel start: a captured CUDA graph resamples
 * the mask on every replay once the 8-byte seed is refreshed (inference, MC dropout).         */
int tgt_gelu_dropout_fwd_dseed(const void *u, void *y, int64_t n, float p_drop,
                               const uint64_t *seed_ptr, int dtype, void *stream);

/* ---- residual: out = res + scale[b] * x  (per-sample DropPath + in-place add) --------------
 * replaces DropPath + add_ at lib/tgt/layers/layers.py:163-177, 269-290.
 * x,out: [B, inner] (dtype) ; res: [B, inner] (res_dtype) or NULL (=0) ; scale: [B] f32 or NULL (=1). */
int tgt_scaled_residual(const void *x, const void *res, const float *scale, void *out,
                        int64_t B, int64_t inner, int dtype, int res_dtype, void *stream);

/* ---- cross-entropy over distance bins ("next" row 8f-3: the loss end of the distance head) -------
 * replaces F.cross_entropy(dist_logits, dist_targ, reduction='none') of DiscreteDistLoss
 * (lib/training_schemes/pcqm/commons.py:26-48); the masked mean stays in the caller ([R] vectors).
 * logits:[rows, nbins] (dtype, row pitch ld), target:[rows] int64 bin index; nbins % 8 == 0, <= 1024.
 * fwd: xent[r] = lse_r - logits[r, target_r], lse[r] = log sum_c exp(logits[r, c])   (fp32)
 * bwd: dlogits[r, c] = (exp(logits[r,c] - lse_r) - [c == target_r]) * gx[r]   (dense [rows, nbins], dtype);
 *      rows with gx[r] == 0 (masked pairs) are written as zeros without being read.              */
int tgt_xent_rows_fwd(const void *logits, int64_t ld, const int64_t *target, float *xent, float *lse,
                      int64_t rows, int nbins, int dtype, void *stream);
int tgt_xent_rows_bwd(const void *logits, int64_t ld, const int64_t *target, const float *lse,
                      const float *gx, void *dlogits, int64_t rows, int nbins, int dtype, void *stream);

/* ---- Gaussian basis of the 3-D distance embedding ("next" row 8f-3) ---------------------------
 * replaces the element-wise chain of lib/models/pcqm/layers.py:25-48 (GaussianLayer):
 * out[r,k] = exp(-0.5 ((x_r - mu_k)/sd_k)^2) / (sqrt(2*3.14159) sd_k),  x:[rows] f32, mu,sd:[K] f32,
 * K <= 128, K % 4 == 0.  Backward recomputes the basis from x; dmu, dsd:[K] must be ZERO on entry. */
int tgt_gaussian_basis_fwd(const float *x, const float *mu, const float *sd, void *out, int64_t rows,
                           int K, int out_dtype, void *stream);
int tgt_gaussian_basis_bwd(const float *x, const float *mu, const float *sd, const void *dout, float *dx,
                           float *dmu, float *dsd, int64_t rows, int K, int dout_dtype, void *stream);

/* ---- distance-bin decoding of the two-stage inference path ("next" row 8f-3) -------------------
 * replaces lib/training_schemes/pcqm/dist_pred/scheme.py:186-194 (softmax over bins, p + p^T over the atom pair,
 * argmax) and commons.py:72-82 (BinsProcessor.bins2dist: (bin + 0.5) * bin_size, d + d^T, zero diagonal) in one
 * pass over the logits.  logits:[B,N,N,num_bins] (dtype) ; bins:[B,N,N] int16 out or NULL ; dist:[B,N,N] f32 out
 * or NULL ; num_bins <= 512 ; ties go to the lowest bin (torch.argmax).                                          */
int tgt_bins_decode(const void *logits, int16_t *bins, float *dist, int B, int N, int num_bins,
                    float bin_size, int shift_half, int zero_diag, int dtype, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TGT_B200_H */
