// attn_mma_dh8.cu - instantiations of the tensor-core attention core for head dim 8 (see attn_mma.cuh).
#include "attn_mma.cuh"

namespace edgl {
int launch_attention_mma_dh8(const AttnArgs& a, cudaStream_t st) { return launch_attention_mma_dh<8>(a, st); }
}  // namespace edgl
