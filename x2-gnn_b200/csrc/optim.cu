// optim.cu -- the parameter-update tail of a training step (trainer.py:43-48 + train_ema.py:45-47) over FLAT fp32
// buffers, in two launches: global-norm gradient clipping (torch.nn.utils.clip_grad_norm_), the Adam step
// (torch.optim.Adam defaults: no weight decay, no amsgrad) and the exponential moving average of the parameters
// (AveragedModel with avg_fn = decay * ema + (1 - decay) * p).  The reference runs these as ~300 small launches
// (per-tensor norms, a stack + norm, a clamp, per-tensor multiplies, the Adam foreach groups, one lerp per
// parameter); here every element is read and written once.
//   k_optim_norm : fixed slices of the gradient -> one fp64 partial sum of squares per block; bumps the step count
//   k_optim_step : every block adds the partials in the same fixed order (deterministic, no atomics), forms the
//                  clip coefficient, then updates its elements: g' = g * grad_scale * coef ; m, v, p, ema.
// With data parallelism the flat gradient is all-reduced (sum) before the call and grad_scale = 1 / world.
#include "common.cuh"

namespace x2 {
namespace {

constexpr int kOptThreads = 256;
constexpr int kOptMaxBlocks = 2 * kNumSM;

__global__ void __launch_bounds__(kOptThreads)
k_optim_norm(const float* __restrict__ g, int64_t n, double* __restrict__ partial, float* __restrict__ step) {
  __shared__ double red[kOptThreads / 32];
  double s = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kOptThreads) {
    const double v = (double)g[i];
    s += v * v;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kOptThreads / 32; ++w) t += red[w];
    partial[blockIdx.x] = t;
    if (blockIdx.x == 0) step[0] += 1.0f;        // the update count t of Adam's bias correction
  }
}

struct OptHyper {
  float grad_scale, max_norm, lr, beta1, beta2, eps, ema_decay;
};

__global__ void __launch_bounds__(kOptThreads)
k_optim_step(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
             float* __restrict__ ema, int64_t n, OptHyper h, const double* __restrict__ partial, int nparts,
             const float* __restrict__ step, float* __restrict__ norm_out) {
  __shared__ float s_coef, s_bc1, s_bc2s;
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < nparts; ++i) t += partial[i];           // same order in every block
    const float norm = (float)sqrt(t) * fabsf(h.grad_scale);     // norm of the scaled gradient
    float coef = h.max_norm > 0.f ? h.max_norm / (norm + 1e-6f) : 1.0f;   // clip_grad_norm_: clamp(max / (norm + 1e-6), max = 1)
    coef = fminf(coef, 1.0f);
    s_coef = coef * h.grad_scale;
    const float tt = step[0];
    s_bc1 = 1.0f - powf(h.beta1, tt);
    s_bc2s = sqrtf(1.0f - powf(h.beta2, tt));
    if (blockIdx.x == 0 && norm_out) norm_out[0] = norm;
  }
  __syncthreads();
  const float coef = s_coef, step_size = h.lr / s_bc1, bc2s = s_bc2s;
  const float b1 = h.beta1, b2 = h.beta2, eps = h.eps, w = 1.0f - h.ema_decay;
  for (int64_t i = (int64_t)blockIdx.x * kOptThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kOptThreads) {
    const float gi = g[i] * coef;
    const float mi = m[i] + (gi - m[i]) * (1.0f - b1);          // lerp, as torch's fused Adam
    const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
    const float denom = sqrtf(vi) / bc2s + eps;
    const float pi = p[i] - step_size * (mi / denom);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
    if (ema) ema[i] = ema[i] + (pi - ema[i]) * w;               // decay * ema + (1 - decay) * p
  }
}

}  // namespace
}  // namespace x2

using namespace x2;

extern "C" {

size_t x2_optim_workspace_bytes(int64_t n) {
  (void)n;
  return (size_t)kOptMaxBlocks * sizeof(double) + 256;
}

int x2_optim_tail(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema, int64_t n,
                  float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float ema_decay,
                  float* step, float* norm_out, void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && step, "x2_optim_tail: null pointer");
  X2_CHECK_ARG(n >= 0 && lr >= 0.f && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && eps > 0.f &&
                   ema_decay >= 0.f && ema_decay <= 1.f, "x2_optim_tail: bad hyper-parameters");
  if (ws_bytes < x2_optim_workspace_bytes(n)) { set_error("x2_optim_tail: workspace too small"); return X2_EWORKSPACE; }
  if (n == 0) return X2_OK;
  cudaStream_t st = (cudaStream_t)stream;
  double* partial = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  const int64_t want = cdiv(n, (int64_t)kOptThreads * 8);
  const int blocks = (int)(want < 1 ? 1 : (want > kOptMaxBlocks ? kOptMaxBlocks : want));
  k_optim_norm<<<blocks, kOptThreads, 0, st>>>(grad, n, partial, step);
  X2_LAUNCH_OK();
  const OptHyper h{grad_scale, max_norm, lr, beta1, beta2, eps, ema_decay};
  k_optim_step<<<blocks, kOptThreads, 0, st>>>(param, grad, exp_avg, exp_avg_sq, ema, n, h, partial, blocks, step, norm_out);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // extern "C"
