// Helpers shared by the tensor-core triplet attention kernels (triplet_mma.cu: cp.async staging, triplet_tma.cu: TMA
// staging): tile swizzles, ldmatrix / stmatrix fragment movers, mma.sync wrappers, the fully-masked-row rule.
#pragma once
#include "ptx.cuh"

namespace tgt {

constexpr int TN = 64;          // padded tile size (max N)
constexpr int HD = 16;          // head dim
constexpr float LOG2E = 1.4426950408889634f;
constexpr float NEG_BIG = -3.0e38f;

// ------------------------------------------------------------------------------------------------ helpers
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

template <typename T> struct Mma;
template <> struct Mma<__nv_bfloat16> {
  static __device__ __forceinline__ void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
  }
};
template <> struct Mma<__half> {
  static __device__ __forceinline__ void run(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  static __device__ __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&v);
  }
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;\n" : "=f"(y) : "f"(x));
  return y;
}

// 64 x 16 (32-byte rows) operand tile in shared memory: byte offset of 16-byte chunk c (0/1) of row r.
// chunk index is xor-ed with bit 2 of the row so that the 8 rows of an ldmatrix hit 8 distinct 16B bank groups.
__device__ __forceinline__ uint32_t tile_off(int r, int c) { return (uint32_t)(r * 32 + ((c ^ ((r >> 2) & 1)) << 4)); }
// 64 x 64 (128-byte rows) exchange tile: chunk (0..7) xor (row & 7)  (the classic 128B swizzle)
__device__ __forceinline__ uint32_t xch_off(int r, int c) { return (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4)); }

// A fragment (16 rows m0.., 16 k) from a [row][16] tile              (Q, dO)
__device__ __forceinline__ void load_a_rows(uint32_t (&a)[4], uint32_t base, int m0, int lane) {
  const int r = m0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  ldsm_x4(a, base + tile_off(r, lane >> 4));
}
// B fragments for two n-tiles (n = rows n0..n0+15 of the tile, k = the 16 columns), tile stored [n][k]   (K for S, V for dA)
__device__ __forceinline__ void load_b_nk(uint32_t (&b)[4], uint32_t base, int n0, int lane) {
  const int r = n0 + (lane & 7) + (lane >> 4) * 8;
  ldsm_x4(b, base + tile_off(r, (lane >> 3) & 1));
}
// B fragments for two n-tiles (n = the 16 columns), k = rows k0..k0+15, tile stored [k][n]  (V for PV, K for dQ, Q/dO for dK/dV)
__device__ __forceinline__ void load_b_kn(uint32_t (&b)[4], uint32_t base, int k0, int lane) {
  const int r = k0 + (lane & 7) + ((lane >> 3) & 1) * 8;
  ldsm_x4_t(b, base + tile_off(r, lane >> 4));
}
// A fragment = X^T block: m = columns m0..m0+15 of X, k = rows k0..k0+15 of X, X is the 64x64 exchange tile
__device__ __forceinline__ void load_a_xt(uint32_t (&a)[4], uint32_t base, int m0, int k0, int lane) {
  const int r = k0 + (lane & 7) + (lane >> 4) * 8;
  const int c = (m0 >> 3) + ((lane >> 3) & 1);
  ldsm_x4_t(a, base + xch_off(r, c));
}

// Rows whose N keys are ALL masked (padding atoms as queries) do not depend on j.  The reference gives them a
// uniform softmax (the -3.4e38 mask absorbs the logits).  Make that exact and lse-representable: zero the row's
// logit scale and bias, so x = 0, P = 1/N, lse2 = log2(N).  Returns the per-row logit scale.
__device__ __forceinline__ void fix_fully_masked_rows(float (&eb)[8][4], float c1, float &c1r0, float &c1r1) {
  int fm0 = 1, fm1 = 1;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    fm0 &= (eb[nt][0] <= -1e37f) & (eb[nt][1] <= -1e37f);
    fm1 &= (eb[nt][2] <= -1e37f) & (eb[nt][3] <= -1e37f);
  }
  fm0 &= __shfl_xor_sync(0xffffffffu, fm0, 1);
  fm0 &= __shfl_xor_sync(0xffffffffu, fm0, 2);
  fm1 &= __shfl_xor_sync(0xffffffffu, fm1, 1);
  fm1 &= __shfl_xor_sync(0xffffffffu, fm1, 2);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    if (fm0) {
      eb[nt][0] = eb[nt][0] == -INFINITY ? -INFINITY : 0.f;
      eb[nt][1] = eb[nt][1] == -INFINITY ? -INFINITY : 0.f;
    }
    if (fm1) {
      eb[nt][2] = eb[nt][2] == -INFINITY ? -INFINITY : 0.f;
      eb[nt][3] = eb[nt][3] == -INFINITY ? -INFINITY : 0.f;
    }
  }
  c1r0 = fm0 ? 0.f : c1;
  c1r1 = fm1 ? 0.f : c1;
}

struct RowMap {
  int64_t bN;      // b * N
  int N, j, dir;
  __device__ __forceinline__ int64_t qrow(int i) const { return (bN + i) * N + j; }
  __device__ __forceinline__ int64_t krow(int k) const { return dir == 0 ? (bN + j) * N + k : (bN + k) * N + j; }
};


// ---- stmatrix: four 8x8 16-bit matrices straight from mma accumulator-layout registers
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1,%2,%3,%4};\n" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
               : "memory");
}
// a warp's 16 x 16 fp32 accumulator block (two n-tiles) -> rows m0..m0+15 of a [64][16] operand tile, scaled by sc
template <typename T>
__device__ __forceinline__ void store_c_tile(uint32_t base, int m0, int lane, const float (&c)[2][4], float sc) {
  const int m = lane >> 3;
  const uint32_t addr = base + tile_off(m0 + (m & 1) * 8 + (lane & 7), m >> 1);
  stsm_x4(addr, Mma<T>::pack(c[0][0] * sc, c[0][1] * sc), Mma<T>::pack(c[0][2] * sc, c[0][3] * sc),
          Mma<T>::pack(c[1][0] * sc, c[1][1] * sc), Mma<T>::pack(c[1][2] * sc, c[1][3] * sc));
}
// packed fragments of n-tiles nt, nt+1 (rows g / g+8 of the warp's 16-row block) -> the 64 x 64 exchange tile
__device__ __forceinline__ void store_xch_pair(uint32_t base, int m0, int lane, int nt, uint32_t lo0, uint32_t hi0,
                                               uint32_t lo1, uint32_t hi1) {
  const int m = lane >> 3;
  const uint32_t addr = base + xch_off(m0 + (m & 1) * 8 + (lane & 7), nt + (m >> 1));
  stsm_x4(addr, lo0, hi0, lo1, hi1);
}

}  // namespace tgt
