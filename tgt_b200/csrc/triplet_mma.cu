// bf16/fp16 tensor-core triplet attention core (placeholder until the MMA kernels land).
#include "common.cuh"
namespace tgt {
bool triplet_attn_mma_supported(const tgt_triplet_attn_desc &) { return false; }
int triplet_attn_fwd_mma(const tgt_triplet_attn_desc &, const void *, const float *, void *, float *, cudaStream_t) {
  return fail("triplet_attn_fwd_mma: not built");
}
int triplet_attn_bwd_mma(const tgt_triplet_attn_desc &, const void *, const float *, const void *, const void *,
                         const float *, void *, cudaStream_t) {
  return fail("triplet_attn_bwd_mma: not built");
}
}  // namespace tgt
