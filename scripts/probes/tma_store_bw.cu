// TMA tensor-store bandwidth by box shape (not part of the library): 148 x k persistent CTAs, 8 warps each, every warp
// streams 32-row boxes of `cols` 16-bit columns from a shared-memory tile into a [R, 1600] bf16 matrix, walking the matrix
// the way the GEMM epilogue does (CTA = column slice x row tiles).  Prints GB/s for 32-, 64- and 128-column boxes.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void __launch_bounds__(256) k_store(const __grid_constant__ CUtensorMap map, int cols, int n_slices, int slice_cols,
                                               int64_t rows, int ctas_per_slice) {
  extern __shared__ __align__(1024) uint8_t sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % n_slices, q = blockIdx.x / n_slices;
  const uint32_t tile = (uint32_t)__cvta_generic_to_shared(sm) + warp * 8192;
  for (int i = lane; i < 8192 / 16; i += 32) reinterpret_cast<uint4 *>(sm + warp * 8192)[i] = make_uint4(i, warp, 2, 3);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    const int quad = warp & 3, half = warp >> 2;
    for (int64_t mt = q; mt * 128 < rows; mt += ctas_per_slice) {
      for (int c = half * cols; c + cols <= slice_cols; c += 2 * cols) {
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(&map), "r"(tile),
                     "r"(slice * slice_cols + c), "r"((int)(mt * 128 + quad * 32))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
      }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

int main() {
  const int64_t R = 256 * 64 * 64;
  const int N = 1600;
  void *p;
  cudaMalloc(&p, (size_t)R * N * 2);
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void **)&enc, cudaEnableDefault, &qr);
  cudaFuncSetAttribute(k_store, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int cols : {32, 64, 128}) {
    CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)R};
    const cuuint64_t gstr[1] = {(cuuint64_t)N * 2};
    const cuuint32_t box[2] = {(cuuint32_t)cols, 32u}, es[2] = {1u, 1u};
    const CUtensorMapSwizzle sw = cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : (cols == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE);
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    // slices of 256 columns (6 full slices = 1536 columns are written; enough for a bandwidth figure)
    const int n_slices = 6, slice_cols = 256, cps = 148 / n_slices;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int i = 0; i < 2; ++i) k_store<<<n_slices * cps, 256, 65536>>>(map, cols, n_slices, slice_cols, R, cps);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) k_store<<<n_slices * cps, 256, 65536>>>(map, cols, n_slices, slice_cols, R, cps);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    ms /= 5;
    const double bytes = (double)R * n_slices * slice_cols * 2;
    printf("{\"box_cols\": %d, \"ms\": %.3f, \"gbs\": %.0f}\n", cols, ms, bytes / ms / 1e6);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
