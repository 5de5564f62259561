// HBM write-bandwidth probe (not part of the library): write-only streams with 16-byte vector stores and with bulk
// shared->global copies, plus a read-only stream, each timed with CUDA events.  Build + run: see scripts/probes/run_write_bw.sh
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(256) k_store(uint4 *p, size_t n) {
  const uint4 v = make_uint4(1, 2, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void __launch_bounds__(256) k_store_cs(uint4 *p, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    asm volatile("st.global.cs.v4.b32 [%0], {%1, %1, %1, %1};" ::"l"(p + i), "r"(7) : "memory");
}
__global__ void __launch_bounds__(256) k_load(const uint4 *p, size_t n, uint32_t *out) {
  uint32_t acc = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    acc ^= v.x ^ v.y ^ v.z ^ v.w;
  }
  if (acc == 0x12345678u) *out = acc;
}
// bulk stores: each CTA fills a 32 KB shared tile once and streams it out with cp.async.bulk (one thread issues)
__global__ void __launch_bounds__(128) k_bulk(uint8_t *p, size_t bytes) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < 32768 / 16; i += blockDim.x) reinterpret_cast<uint4 *>(sm)[i] = make_uint4(i, 1, 2, 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    for (size_t off = (size_t)blockIdx.x * 32768; off + 32768 <= bytes; off += (size_t)gridDim.x * 32768) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(p + off), "r"(s), "r"(32768) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

template <typename F> static float timeit(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i) f();
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i) f();
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 10;
}

int main() {
  const size_t bytes = (size_t)3355443200ull;   // = R * 1600 * 2 at config 3
  uint8_t *p;
  uint32_t *o;
  cudaMalloc(&p, bytes);
  cudaMalloc(&o, 4);
  const size_t n = bytes / 16;
  for (int mult : {2, 4, 8, 16}) {
    const int grid = 148 * mult;
    const float s = timeit([&] { k_store<<<grid, 256>>>((uint4 *)p, n); });
    const float c = timeit([&] { k_store_cs<<<grid, 256>>>((uint4 *)p, n); });
    const float l = timeit([&] { k_load<<<grid, 256>>>((const uint4 *)p, n, o); });
    printf("{\"grid\": %d, \"store_gbs\": %.0f, \"store_cs_gbs\": %.0f, \"load_gbs\": %.0f}\n", grid, bytes / s / 1e6, bytes / c / 1e6,
           bytes / l / 1e6);
  }
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
  for (int mult : {1, 2, 4}) {
    const float t = timeit([&] { k_bulk<<<148 * mult, 128, 32768>>>(p, bytes); });
    printf("{\"grid\": %d, \"bulk_store_gbs\": %.0f}\n", 148 * mult, bytes / t / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
