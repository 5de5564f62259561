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
