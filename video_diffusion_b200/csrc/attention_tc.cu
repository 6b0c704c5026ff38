// Spatial attention on tensor cores (bf16 in, fp32 softmax/accumulate).
#include "common.cuh"

namespace vdm {

int attn_spatial_tc(const void* qkv, int n_img, int L, int heads, int hd, void* out_a, int out_dtype,
                    cudaStream_t stream) {
  set_error("attn_spatial: bf16 tensor-core kernel not available in this build");
  return -1;
}

}  // namespace vdm
