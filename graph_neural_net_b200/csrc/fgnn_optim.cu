// Fused multi-tensor Adam on ONE flat fp32 buffer: the reference's
// torch.optim.Adam(lr) of configure_optimizers (models/trainers.py:92-104) for every parameter in a single launch,
// reading the data-parallel step's scalars from device memory so that no host round trip sits between the gradient
// all-reduce and the update:
//   grad_div  (device, may be NULL): gradients are divided by *grad_div -- the GLOBAL number of rows of the loss
//             (toolbox/losses.py:32-34 divides by sum_b n_b after the sum over ranks);
//   skip_flag (device, may be NULL): *skip_flag > 0 -> the step is a no-op (AMP semantics: some rank's 16-bit
//             gradients overflowed for the current loss scale).
#include "fgnn_common.cuh"

namespace fgnn {
namespace {

__global__ void __launch_bounds__(256)
adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long n,
                 float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
                 const float* __restrict__ grad_div, const float* __restrict__ skip_flag) {
  if (skip_flag && skip_flag[0] > 0.f) return;
  const float inv = grad_div ? 1.f / fmaxf(grad_div[0], 1.f) : 1.f;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    float gi = g[i] * inv;
    const float pi = p[i];
    if (weight_decay != 0.f) gi = fmaf(weight_decay, pi, gi);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);          // torch.optim.Adam: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);     // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;                  // (exp_avg_sq.sqrt() / sqrt(bias_correction2)).add_(eps)
    p[i] = pi - (lr / bc1) * (mi / denom);                           // param.addcdiv_(exp_avg, denom, value=-lr / bias_correction1)
  }
}

}  // namespace
}  // namespace fgnn

extern "C" int fgnn_adam_step_f32(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2,
                                  float eps, float weight_decay, int32_t step, const float* grad_div, const float* skip_flag,
                                  void* stream) {
  using namespace fgnn;
  FGNN_CHECK_ARG(p && g && m && v, "null pointer");
  FGNN_CHECK_ARG(n >= 0 && step >= 1, "bad n=%lld step=%d", (long long)n, step);
  if (n == 0) return FGNN_OK;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  adam_step_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, (long)n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt,
                                                         grad_div, skip_flag);
  FGNN_LAUNCHED();
  return FGNN_OK;
}
