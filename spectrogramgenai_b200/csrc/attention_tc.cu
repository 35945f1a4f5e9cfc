// K4 (tensor-core engine): flash-style self-attention on tcgen05 -- placeholder until the kernel lands.
#include "tc_common.cuh"

namespace sg {
int attention_tc(const void* qkv, void* out, int rows, int L, int C, int heads, int act_dtype, cudaStream_t stream) {
  (void)qkv; (void)out; (void)rows; (void)L; (void)C; (void)heads; (void)act_dtype; (void)stream;
  set_error("sg_attention: the tcgen05 attention kernel is not built into this library yet");
  return SG_ERR_ARG;
}
}  // namespace sg
