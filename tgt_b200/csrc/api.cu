// C-ABI entry points that dispatch between the tensor-core and the generic kernels, plus library state.
#include "common.cuh"
#include <mutex>
#include <vector>
#include <string>
#include <map>
#include <algorithm>
#include <string.h>

namespace tgt {
thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_policy{0};

// ---- optional device-side kernel timer (see common.cuh)
static std::atomic<int> g_timer_on{0};
struct TimerRec {
  const char *name;
  cudaEvent_t a, b;
};
static std::mutex g_timer_mu;
static std::vector<TimerRec> g_timer_recs;
static thread_local cudaEvent_t g_timer_open_b = nullptr;

void kernel_timer_begin(const char *name, cudaStream_t st) {
  g_timer_open_b = nullptr;
  if (!g_timer_on.load(std::memory_order_relaxed)) return;
  TimerRec r{name, nullptr, nullptr};
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  g_timer_open_b = r.b;
  std::lock_guard<std::mutex> lk(g_timer_mu);
  g_timer_recs.push_back(r);
}
void kernel_timer_end(cudaStream_t st) {
  if (g_timer_open_b) cudaEventRecord(g_timer_open_b, st);
  g_timer_open_b = nullptr;
}

// triplet_simt.cu
int triplet_attn_fwd_simt(const tgt_triplet_attn_desc &, const void *, const float *, void *, float *, cudaStream_t);
int triplet_attn_bwd_simt(const tgt_triplet_attn_desc &, const void *, const float *, const void *, const void *,
                          const float *, void *, cudaStream_t);
int triplet_aggr_fwd_simt(const tgt_triplet_aggr_desc &, const void *, const float *, void *, float *, cudaStream_t);
int triplet_aggr_weights_simt(const tgt_triplet_aggr_desc &, const void *, const float *, float *, cudaStream_t);
int triplet_aggr_bwd_weights_simt(const tgt_triplet_aggr_desc &, const void *, const float *, const float *, void *,
                                  cudaStream_t);
bool triplet_aggr_tma_supported(const tgt_triplet_aggr_desc &);
int triplet_aggr_fwd_tma_launch(const tgt_triplet_aggr_desc &, const void *, void *, const float *, cudaStream_t);
int triplet_aggr_bwd_tma_launch(const tgt_triplet_aggr_desc &, const void *, const void *, const float *, float *, void *,
                                cudaStream_t);
int triplet_aggr_bwd_simt(const tgt_triplet_aggr_desc &, const void *, const float *, const void *, const float *,
                          float *, void *, cudaStream_t);
// triplet_mma.cu
bool triplet_attn_mma_supported(const tgt_triplet_attn_desc &);
size_t triplet_attn_mma_workspace(const tgt_triplet_attn_desc &, int backward);
int triplet_attn_fwd_mma(const tgt_triplet_attn_desc &, const void *, const float *, void *, float *, void *, size_t,
                         cudaStream_t);
int triplet_attn_bwd_mma(const tgt_triplet_attn_desc &, const void *, const float *, const void *, const void *,
                         const float *, void *, void *, size_t, const void *, float *, cudaStream_t);
bool triplet_attn_bwd_bias_available(const tgt_triplet_attn_desc &);
int triplet_attn_fwd_tc_f32out(const tgt_triplet_attn_desc &, const void *, const float *, float *, float *, void *, size_t,
                               cudaStream_t);
bool triplet_attn_fused_supported(const tgt_triplet_attn_desc &D, int We);
int triplet_attn_fused_fwd(const tgt_triplet_attn_desc &D, int We, const void *x, int64_t ldx, const float *mean,
                           const float *rstd, const void *wf, const float *wcolsum, const float *wbias,
                           const void *proj_eg, const float *mask, void *va, float *stats, void *ws, size_t ws_bytes,
                           cudaStream_t st);
}  // namespace tgt

using namespace tgt;

extern "C" int tgt_version(void) { return 100; }
extern "C" const char *tgt_last_error(void) { return g_err; }
extern "C" uint64_t tgt_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" void tgt_set_kernel_policy(int policy) { g_policy.store(policy); }

extern "C" void tgt_kernel_timer_enable(int on) {
  g_timer_on.store(on ? 1 : 0);
  if (on) return;
  std::lock_guard<std::mutex> lk(g_timer_mu);
  for (auto &r : g_timer_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  g_timer_recs.clear();
}

// writes "name launches total_ms\n" lines into buf (NUL-terminated, truncated to n); returns the number of lines
extern "C" int tgt_kernel_timer_read(char *buf, size_t n) {
  std::map<std::string, std::pair<int, double>> agg;
  {
    std::lock_guard<std::mutex> lk(g_timer_mu);
    for (auto &r : g_timer_recs) {
      float ms = 0.f;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) {
        auto &e = agg[r.name];
        e.first += 1;
        e.second += ms;
      }
    }
  }
  std::string out;
  for (auto &kv : agg) out += kv.first + " " + std::to_string(kv.second.first) + " " + std::to_string(kv.second.second) + "\n";
  if (buf && n) {
    const size_t m = std::min(n - 1, out.size());
    memcpy(buf, out.data(), m);
    buf[m] = 0;
  }
  return (int)agg.size();
}

static int attn_check(const tgt_triplet_attn_desc *D) {
  if (!D) return fail("triplet_attn: null descriptor");
  if (D->B <= 0 || D->N <= 0 || D->H <= 0 || D->d <= 0)
    return fail("triplet_attn: bad shape B=%d N=%d H=%d d=%d", D->B, D->N, D->H, D->d);
  if (D->N > 64) return fail("triplet_attn: N=%d > 64 unsupported", D->N);
  if (D->d > 64) return fail("triplet_attn: head dim %d > 64 unsupported", D->d);
  if (D->B > 65535 || D->H > 65535) return fail("triplet_attn: B or H > 65535 unsupported");
  for (int dir = 0; dir < 2; ++dir) {
    if (D->off_q[dir] < 0 || D->off_k[dir] < 0 || D->off_v[dir] < 0)
      return fail("triplet_attn: q/k/v column blocks are mandatory");
    if ((D->off_g[dir] >= 0) && (D->off_e[dir] < 0)) return fail("triplet_attn: gate without bias unsupported");
  }
  return 0;
}

extern "C" size_t tgt_triplet_attn_workspace_bytes(const tgt_triplet_attn_desc *D, int backward) {
  if (!D || attn_check(D)) return 0;
  if (g_policy.load() != 1 && triplet_attn_mma_supported(*D)) return triplet_attn_mma_workspace(*D, backward);
  return 0;
}

extern "C" int tgt_triplet_attn_fwd(const tgt_triplet_attn_desc *D, const void *proj, const float *mask, void *va,
                                    float *stats, void *ws, size_t ws_bytes, void *stream) {
  if (int e = attn_check(D)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_policy.load() != 1 && triplet_attn_mma_supported(*D))
    return triplet_attn_fwd_mma(*D, proj, mask, va, stats, ws, ws_bytes, st);
  return triplet_attn_fwd_simt(*D, proj, mask, va, stats, st);
}

extern "C" int tgt_triplet_attn_fwd_f32out(const tgt_triplet_attn_desc *D, const void *proj, const float *mask,
                                           float *va_f32, float *stats, void *ws, size_t ws_bytes, void *stream) {
  if (int e = attn_check(D)) return e;
  if (!va_f32) return fail("triplet_attn_fwd_f32out: null output");
  return triplet_attn_fwd_tc_f32out(*D, proj, mask, va_f32, stats, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int tgt_triplet_attn_bwd_bias_supported(const tgt_triplet_attn_desc *D) {
  if (!D || attn_check(D)) return 0;
  return (g_policy.load() != 1 && triplet_attn_bwd_bias_available(*D)) ? 1 : 0;
}

extern "C" int tgt_triplet_attn_bwd_bias(const tgt_triplet_attn_desc *D, const void *proj, const float *mask,
                                         const void *va, const void *dva, const float *stats, void *dproj, void *ws,
                                         size_t ws_bytes, const void *fwd_workspace, float *dbias, void *stream) {
  if (int e = attn_check(D)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_policy.load() != 1 && triplet_attn_mma_supported(*D))
    return triplet_attn_bwd_mma(*D, proj, mask, va, dva, stats, dproj, ws, ws_bytes, fwd_workspace, dbias, st);
  if (dbias) return fail("triplet_attn_bwd_bias: not available with the generic kernels (check ..._bias_supported)");
  return triplet_attn_bwd_simt(*D, proj, mask, va, dva, stats, dproj, st);
}

extern "C" int tgt_triplet_attn_bwd_tiles(const tgt_triplet_attn_desc *D, const void *proj, const float *mask,
                                          const void *va, const void *dva, const float *stats, void *dproj, void *ws,
                                          size_t ws_bytes, const void *fwd_workspace, void *stream) {
  return tgt_triplet_attn_bwd_bias(D, proj, mask, va, dva, stats, dproj, ws, ws_bytes, fwd_workspace, nullptr, stream);
}

extern "C" int tgt_triplet_attn_bwd(const tgt_triplet_attn_desc *D, const void *proj, const float *mask,
                                    const void *va, const void *dva, const float *stats, void *dproj, void *ws,
                                    size_t ws_bytes, void *stream) {
  return tgt_triplet_attn_bwd_tiles(D, proj, mask, va, dva, stats, dproj, ws, ws_bytes, nullptr, stream);
}

extern "C" int tgt_triplet_attn_fused_supported(const tgt_triplet_attn_desc *D, int We) {
  if (!D || D->B <= 0 || D->N <= 0 || D->H <= 0) return 0;
  // opt-in (policy 3): measured on B200 the fused kernel is bound by L2->SM operand traffic (8 CTAs per graph each
  // stream the graph's edge rows twice) and does not yet beat the un-fused TMA path -- see DESIGN.md section 4.3
  return (g_policy.load() == 3 && triplet_attn_fused_supported(*D, We)) ? 1 : 0;
}

extern "C" int tgt_triplet_attn_fused_fwd(const tgt_triplet_attn_desc *D, const void *x, int64_t ldx, int We,
                                          const float *row_mean, const float *row_rstd, const void *wf,
                                          const float *wcolsum, const float *wbias, const void *proj_eg,
                                          const float *mask, void *va, float *stats, void *ws, size_t ws_bytes,
                                          void *stream) {
  if (int e = attn_check(D)) return e;
  if (!x || !row_mean || !row_rstd || !wf || !wcolsum || !wbias || !proj_eg || !mask || !va || !stats)
    return fail("triplet_attn_fused_fwd: null argument");
  return triplet_attn_fused_fwd(*D, We, x, ldx, row_mean, row_rstd, wf, wcolsum, wbias, proj_eg, mask, va, stats, ws,
                                ws_bytes, (cudaStream_t)stream);
}

static int aggr_check(const tgt_triplet_aggr_desc *D) {
  if (!D) return fail("triplet_aggr: null descriptor");
  if (D->B <= 0 || D->N <= 0 || D->H <= 0 || D->d <= 0)
    return fail("triplet_aggr: bad shape B=%d N=%d H=%d d=%d", D->B, D->N, D->H, D->d);
  if (D->N > 64) return fail("triplet_aggr: N=%d > 64 unsupported", D->N);
  if (D->d > 64) return fail("triplet_aggr: head dim %d > 64 unsupported", D->d);
  if (D->B > 65535) return fail("triplet_aggr: B > 65535 unsupported");
  for (int dir = 0; dir < 2; ++dir)
    if (D->off_v[dir] < 0 || D->off_e[dir] < 0) return fail("triplet_aggr: v/e column blocks are mandatory");
  return 0;
}

// 16-bit, head dim 16: softmax/gate weights (O(N^2 H), SIMT) + tensor-core / TMA apply kernels (O(N^3), triplet_tma.cu);
// kernel policy 1 and every other shape: the generic SIMT kernels
extern "C" int tgt_triplet_aggr_fwd(const tgt_triplet_aggr_desc *D, const void *proj, const float *mask, void *va,
                                    float *aw, void *stream) {
  if (int e = aggr_check(D)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_policy.load() != 1 && triplet_aggr_tma_supported(*D)) {
    if (int e = triplet_aggr_weights_simt(*D, proj, mask, aw, st)) return e;
    return triplet_aggr_fwd_tma_launch(*D, proj, va, aw, st);
  }
  return triplet_aggr_fwd_simt(*D, proj, mask, va, aw, st);
}

extern "C" int tgt_triplet_aggr_bwd(const tgt_triplet_aggr_desc *D, const void *proj, const float *mask,
                                    const void *dva, const float *aw, float *daw, void *dproj, void *stream) {
  if (int e = aggr_check(D)) return e;
  cudaStream_t st = (cudaStream_t)stream;
  if (g_policy.load() != 1 && triplet_aggr_tma_supported(*D)) {
    if (int e = triplet_aggr_bwd_tma_launch(*D, proj, dva, aw, daw, dproj, st)) return e;
    return triplet_aggr_bwd_weights_simt(*D, proj, mask, daw, dproj, st);
  }
  return triplet_aggr_bwd_simt(*D, proj, mask, dva, aw, daw, dproj, st);
}
