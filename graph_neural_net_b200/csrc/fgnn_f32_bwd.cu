// fp32 backward operators (MlpBlock_Real backward with recomputation).  Placeholder until the
// training path lands; returns FGNN_ERR_UNSUPPORTED loudly.
#include "fgnn_f32.cuh"
namespace fgnn { namespace f32 {
int mlp_bwd(const fgnn_mlp_params&, const fgnn_mlp_grads&, const float*, const float*, const float*, float*,
            int, int, const int32_t*, void*, size_t, cudaStream_t) {
  return fail(FGNN_ERR_UNSUPPORTED, "fgnn_mlp_bwd_f32 not implemented yet");
}
}}
