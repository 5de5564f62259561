// Distance-bin decoding of the two-stage inference path ("next" row 8f-3), one pass over the logits:
//   reference dist_pred/scheme.py:186-194  (softmax over bins, p + p^T over the atom pair, argmax)
//   reference commons.py:72-82             (bins2dist: (bin + 0.5) * bin_size, d + d^T, zero diagonal)
// The reference materialises the fp32 probabilities ([B,N,N,bins], 2.4 GB at B = 512, N = 48, 256 bins), transposes and
// adds them, runs argmax, ships the bins to the host and re-reads them through parquet files; here one warp owns an
// unordered pair {i, j}: it reads the two logit rows once, keeps both softmaxes in registers, and writes the bin and the
// decoded distance of (i, j) and (j, i).  HBM traffic = the logits once + 6 bytes per pair.
#include "common.cuh"

namespace tgt {

constexpr int BINS_MAX_PER_LANE = 16;      // bins <= 512

template <typename T>
__global__ void __launch_bounds__(256)
bins_decode_kernel(const T *__restrict__ logits, int16_t *__restrict__ bins, float *__restrict__ dist, int B, int N,
                   int nb, float bin_size, int shift_half, int zero_diag) {
  const int lane = threadIdx.x & 31;
  const int64_t pairs_per_graph = (int64_t)N * (N + 1) / 2;
  const int64_t total = (int64_t)B * pairs_per_graph;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int per = (nb + 31) / 32;
  for (int64_t w = warp0; w < total; w += nwarps) {
    const int b = (int)(w / pairs_per_graph);
    int64_t r = w - (int64_t)b * pairs_per_graph;
    // unrank the pair: row i holds N - i pairs (j = i .. N-1)
    int i = 0;
    while (r >= N - i) {
      r -= N - i;
      ++i;
    }
    const int j = i + (int)r;
    const T *pa = logits + (((int64_t)b * N + i) * N + j) * nb;
    const T *pb = logits + (((int64_t)b * N + j) * N + i) * nb;
    float xa[BINS_MAX_PER_LANE], xb[BINS_MAX_PER_LANE];
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int k = 0; k < BINS_MAX_PER_LANE; ++k) {
      const int c = k * 32 + lane;
      xa[k] = xb[k] = -INFINITY;
      if (k < per && c < nb) {
        xa[k] = to_f(pa[c]);
        xb[k] = to_f(pb[c]);
      }
      ma = fmaxf(ma, xa[k]);
      mb = fmaxf(mb, xb[k]);
    }
    ma = warp_max(ma);
    mb = warp_max(mb);
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int k = 0; k < BINS_MAX_PER_LANE; ++k) {
      if (k < per) {
        xa[k] = expf(xa[k] - ma);
        xb[k] = expf(xb[k] - mb);
        sa += xa[k];
        sb += xb[k];
      }
    }
    sa = 1.f / warp_sum(sa);
    sb = 1.f / warp_sum(sb);
    float best = -1.f;
    int arg = 0;
#pragma unroll
    for (int k = 0; k < BINS_MAX_PER_LANE; ++k) {
      const int c = k * 32 + lane;
      if (k < per && c < nb) {
        const float p = xa[k] * sa + xb[k] * sb;
        if (p > best) {                    // strict: the first (lowest) bin wins ties, like torch.argmax
          best = p;
          arg = c;
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) {
        best = ob;
        arg = oa;
      }
    }
    if (lane == 0) {
      const float half_d = ((float)arg + (shift_half ? 0.5f : 0.f)) * bin_size;
      const float d = (zero_diag && i == j) ? 0.f : half_d + half_d;      // dist + dist^T of a symmetric bin matrix
      const int64_t o1 = ((int64_t)b * N + i) * N + j, o2 = ((int64_t)b * N + j) * N + i;
      if (bins) {
        bins[o1] = (int16_t)arg;
        bins[o2] = (int16_t)arg;
      }
      if (dist) {
        dist[o1] = d;
        dist[o2] = d;
      }
    }
  }
}

}  // namespace tgt

using namespace tgt;

extern "C" int tgt_bins_decode(const void *logits, int16_t *bins, float *dist, int B, int N, int num_bins,
                               float bin_size, int shift_half, int zero_diag, int dtype, void *stream) {
  if (B <= 0 || N <= 0) return 0;
  if (!logits || (!bins && !dist)) return fail("bins_decode: null argument");
  if (num_bins <= 0 || num_bins > 32 * BINS_MAX_PER_LANE) return fail("bins_decode: num_bins=%d unsupported (1..512)", num_bins);
  const int64_t warps = (int64_t)B * N * (N + 1) / 2;
  int64_t g = (warps + 7) / 8;
  if (g > 148 * 16) g = 148 * 16;
  TGT_DISPATCH_DTYPE(dtype, T, {
    bins_decode_kernel<T><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const T *)logits, bins, dist, B, N, num_bins,
                                                                   bin_size, shift_half, zero_diag);
  });
  return check_launch("bins_decode");
}
