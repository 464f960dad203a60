// attn_mma_dh16.cu - instantiations of the tensor-core attention core for head dim 16 (see attn_mma.cuh).
#include "attn_mma.cuh"

namespace edgl {
int launch_attention_mma_dh16(const AttnArgs& a, cudaStream_t st) { return launch_attention_mma_dh<16>(a, st); }
}  // namespace edgl
