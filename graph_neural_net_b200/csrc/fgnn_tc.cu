#include "fgnn_tc.cuh"
namespace fgnn { namespace tc {
size_t embed_workspace_bytes(const fgnn_embed_params&, int, int) { return 0; }
int embed_fwd(const fgnn_embed_params&, int, const float*, float*, int, int, const int32_t*, const int32_t*, void*, size_t, cudaStream_t) { return fail(FGNN_ERR_UNSUPPORTED, "tc path not built yet"); }
size_t debug_matmul_workspace_bytes(int, int, int) { return 0; }
int debug_matmul(int, const float*, const float*, float*, int, int, int, const int32_t*, void*, size_t, cudaStream_t) { return fail(FGNN_ERR_UNSUPPORTED, "tc path not built yet"); }
}}
