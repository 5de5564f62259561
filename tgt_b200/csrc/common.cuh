// Shared device/host helpers for the tgt_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/tgt_b200.h"

namespace tgt {

// ---------------------------------------------------------------- host-side error plumbing
extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_policy;

inline int fail(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return 1;
}

inline int check_launch(const char *what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail("%s: %s", what, cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------- optional per-kernel device timing (bench / profiling)
// Off by default.  When enabled through tgt_kernel_timer_enable(1) the launchers bracket their MAIN kernel with CUDA
// events on the launching stream (prep / post helper kernels excluded); tgt_kernel_timer_read() synchronises on the
// recorded events and reports launches + total milliseconds per kernel name.
void kernel_timer_begin(const char *name, cudaStream_t st);
void kernel_timer_end(cudaStream_t st);
struct KernelTimerScope {
  cudaStream_t st;
  KernelTimerScope(const char *name, cudaStream_t s) : st(s) { kernel_timer_begin(name, s); }
  ~KernelTimerScope() { kernel_timer_end(st); }
};

#define TGT_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) return ::tgt::fail("%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// ---------------------------------------------------------------- dtype helpers
template <typename T> struct DT;
template <> struct DT<float> { static constexpr int code = TGT_F32; };
template <> struct DT<__nv_bfloat16> { static constexpr int code = TGT_BF16; };
template <> struct DT<__half> { static constexpr int code = TGT_F16; };

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f<__half>(float v) { return __float2half_rn(v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }
// accurate sigmoid for the fp32 parity path
__device__ __forceinline__ float sigmoid_acc(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---------------------------------------------------------------- counter-based dropout (shared by the elementwise
// gelu_dropout kernels and the fused GEMM epilogue: nothing is stored, backward regenerates the mask)
// One 32-bit hash serves TWO neighbouring elements (16 bits each): element idx is dropped iff its 16-bit lane of
// hash(idx >> 1) is below round(p * 65536); the survivors are scaled by 65536 / (65536 - threshold).
struct DropCfg {
  uint32_t seed_mix, thresh16;
  float keep_scale;
};
__host__ __device__ inline DropCfg make_drop_cfg(float p_drop, unsigned long long seed) {
  DropCfg c;
  c.seed_mix = (uint32_t)seed * 0x9E3779B9u + (uint32_t)(seed >> 32) * 0x85EBCA6Bu + 0x27D4EB2Fu;
  float t = p_drop * 65536.f + 0.5f;
  c.thresh16 = p_drop > 0.f ? (uint32_t)(t < 1.f ? 1.f : (t > 65535.f ? 65535.f : t)) : 0u;
  c.keep_scale = 65536.f / (float)(65536u - c.thresh16);
  return c;
}
__device__ __forceinline__ uint32_t drop_hash_pair(const DropCfg &c, uint64_t pair_idx) {
  uint32_t x = ((uint32_t)pair_idx ^ c.seed_mix) + (uint32_t)(pair_idx >> 32) * 0xC2B2AE35u;
  x ^= x >> 16;
  x *= 0x7FEB352Du;
  x ^= x >> 15;
  x *= 0x846CA68Bu;
  x ^= x >> 16;
  return x;
}
// multipliers (0 or keep_scale) of elements 2*pair_idx and 2*pair_idx + 1
__device__ __forceinline__ void drop_mask_pair(const DropCfg &c, uint64_t pair_idx, float &m0, float &m1) {
  const uint32_t h = drop_hash_pair(c, pair_idx);
  m0 = (h & 0xFFFFu) < c.thresh16 ? 0.f : c.keep_scale;
  m1 = (h >> 16) < c.thresh16 ? 0.f : c.keep_scale;
}
__device__ __forceinline__ float drop_mask_one(const DropCfg &c, uint64_t idx) {
  const uint32_t h = drop_hash_pair(c, idx >> 1);
  return ((idx & 1) ? (h >> 16) : (h & 0xFFFFu)) < c.thresh16 ? 0.f : c.keep_scale;
}

// 16-byte vector of T
template <typename T> struct Vec16 { static constexpr int n = 16 / sizeof(T); };

template <typename T, int NV>
__device__ __forceinline__ void load_vec(const T *p, float (&out)[NV]) {
  static_assert(NV * sizeof(T) == 16, "16B vectors only");
  uint4 raw = *reinterpret_cast<const uint4 *>(p);
  const T *e = reinterpret_cast<const T *>(&raw);
#pragma unroll
  for (int i = 0; i < NV; ++i) out[i] = to_f(e[i]);
}
template <typename T, int NV>
__device__ __forceinline__ void store_vec(T *p, const float (&in)[NV]) {
  static_assert(NV * sizeof(T) == 16, "16B vectors only");
  uint4 raw;
  T *e = reinterpret_cast<T *>(&raw);
#pragma unroll
  for (int i = 0; i < NV; ++i) e[i] = from_f<T>(in[i]);
  *reinterpret_cast<uint4 *>(p) = raw;
}

// dispatch on dtype code
#define TGT_DISPATCH_DTYPE(code, T, ...)                                   \
  switch (code) {                                                          \
    case TGT_F32: { using T = float; __VA_ARGS__; } break;                 \
    case TGT_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break;        \
    case TGT_F16: { using T = __half; __VA_ARGS__; } break;                \
    default: return ::tgt::fail("unsupported dtype code %d", (int)(code)); \
  }

}  // namespace tgt
