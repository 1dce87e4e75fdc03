// placeholder until the tcgen05 flash kernel lands: bf16 goes through the CUDA-core kernel
#include "wb_internal.h"
namespace wb {
void encoder_attention_tc(const void* qkv, void* out, int B, int S, int H, cudaStream_t stream) {
    encoder_attention_simt(qkv, out, BF16, B, S, H, stream);
}
}  // namespace wb
