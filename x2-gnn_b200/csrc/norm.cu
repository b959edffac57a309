// norm.cu -- graph-wise LayerNorm of the post-conv block.
//
// Reference: model.py:24,46 `LayerNorm(in_channels, eps=1e-8, affine=False)(x=out, batch=data.batch)` =
// torch_geometric.nn.LayerNorm 2.1.0 with a batch vector: statistics over ALL rows and
// channels of a molecule, mean first, variance of the centred values second; executed there as two
// scatter-adds, two gathers and ~8 elementwise passes over [E, D] (and twice that in the backward).
//
// Here: the rows of a molecule are contiguous (PyG collates graph after graph), so molecule g is one
// contiguous range of floats, rowptr[g]*D .. rowptr[g+1]*D.  One CTA per molecule streams its range with
// 128-bit loads: pass 1 sum -> mean, pass 2 centred sum of squares -> rstd, pass 3 writes y; passes 2 and 3
// re-read the range from L1/L2 (a QM9 molecule is <= 360 KB).  Block reductions are fixed-order trees in
// fp64, so the result does not depend on scheduling.  HBM traffic: x once in, y once out.
// Backward (y, dy -> dx): dx = rstd * (dy - mean(dy) - y * mean(dy*y)), same structure, two passes.
#include "common.cuh"

namespace x2 {
namespace {

constexpr int kLnThreads = 512;

// fixed-order block reduction of up to two doubles per thread; every thread returns the totals
__device__ __forceinline__ void block_sum2(double& a, double& b, double* sh /* [2 * 32] */) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();                       // sh may still be read by the previous call
  if (l == 0) { sh[w] = a; sh[32 + w] = b; }
  __syncthreads();
  a = (l < nw) ? sh[l] : 0.0;
  b = (l < nw) ? sh[32 + l] : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
}

__global__ void __launch_bounds__(kLnThreads)
k_graph_ln_fwd(const float* __restrict__ x, const int32_t* __restrict__ rowptr, int D, float eps,
               float* __restrict__ y, float* __restrict__ stats) {
  __shared__ double sh[64];
  const int g = blockIdx.x;
  const int64_t lo = (int64_t)rowptr[g] * D, hi = (int64_t)rowptr[g + 1] * D;
  const int64_t n = hi - lo;
  if (n <= 0) {
    if (threadIdx.x == 0) { stats[2 * g] = 0.f; stats[2 * g + 1] = 0.f; }
    return;
  }
  const float4* x4 = reinterpret_cast<const float4*>(x + lo);     // D % 4 == 0 and x is 16-byte aligned
  float4* y4 = reinterpret_cast<float4*>(y + lo);
  const int64_t n4 = n >> 2;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int64_t i = threadIdx.x; i < n4; i += kLnThreads) {
    const float4 v = __ldg(x4 + i);
    s0 += v.x; s1 += v.y; s2 += v.z; s3 += v.w;
  }
  double a = ((double)s0 + (double)s1) + ((double)s2 + (double)s3), b = 0.0;
  block_sum2(a, b, sh);
  const float mean = (float)(a / (double)n);
  s0 = s1 = s2 = s3 = 0.f;
  for (int64_t i = threadIdx.x; i < n4; i += kLnThreads) {
    const float4 v = __ldg(x4 + i);
    const float c0 = v.x - mean, c1 = v.y - mean, c2 = v.z - mean, c3 = v.w - mean;
    s0 = fmaf(c0, c0, s0); s1 = fmaf(c1, c1, s1); s2 = fmaf(c2, c2, s2); s3 = fmaf(c3, c3, s3);
  }
  a = ((double)s0 + (double)s1) + ((double)s2 + (double)s3);
  b = 0.0;
  block_sum2(a, b, sh);
  const float var = (float)(a / (double)n);
  const float rstd = 1.0f / sqrtf(var + eps);
  if (threadIdx.x == 0) { stats[2 * g] = mean; stats[2 * g + 1] = rstd; }
  for (int64_t i = threadIdx.x; i < n4; i += kLnThreads) {
    const float4 v = __ldg(x4 + i);
    float4 o;
    o.x = (v.x - mean) * rstd; o.y = (v.y - mean) * rstd; o.z = (v.z - mean) * rstd; o.w = (v.w - mean) * rstd;
    y4[i] = o;
  }
}

__global__ void __launch_bounds__(kLnThreads)
k_graph_ln_bwd(const float* __restrict__ y, const float* __restrict__ gy, const int32_t* __restrict__ rowptr,
               int D, const float* __restrict__ stats, float* __restrict__ gx) {
  __shared__ double sh[64];
  const int g = blockIdx.x;
  const int64_t lo = (int64_t)rowptr[g] * D, hi = (int64_t)rowptr[g + 1] * D;
  const int64_t n = hi - lo;
  if (n <= 0) return;
  const float4* y4 = reinterpret_cast<const float4*>(y + lo);
  const float4* g4 = reinterpret_cast<const float4*>(gy + lo);
  float4* o4 = reinterpret_cast<float4*>(gx + lo);
  const int64_t n4 = n >> 2;
  float s0 = 0.f, s1 = 0.f, t0 = 0.f, t1 = 0.f;
  for (int64_t i = threadIdx.x; i < n4; i += kLnThreads) {
    const float4 v = __ldg(y4 + i), d = __ldg(g4 + i);
    s0 += d.x + d.y; s1 += d.z + d.w;
    t0 = fmaf(d.x, v.x, t0); t1 = fmaf(d.y, v.y, t1); t0 = fmaf(d.z, v.z, t0); t1 = fmaf(d.w, v.w, t1);
  }
  double a = (double)s0 + (double)s1, b = (double)t0 + (double)t1;
  block_sum2(a, b, sh);
  const float mg = (float)(a / (double)n), mgy = (float)(b / (double)n);
  const float rstd = stats[2 * g + 1];
  for (int64_t i = threadIdx.x; i < n4; i += kLnThreads) {
    const float4 v = __ldg(y4 + i), d = __ldg(g4 + i);
    float4 o;
    o.x = rstd * (d.x - mg - v.x * mgy); o.y = rstd * (d.y - mg - v.y * mgy);
    o.z = rstd * (d.z - mg - v.z * mgy); o.w = rstd * (d.w - mg - v.w * mgy);
    o4[i] = o;
  }
}

}  // namespace
}  // namespace x2

using namespace x2;

extern "C" {

int x2_graph_layernorm_fwd(const float* x, const int32_t* rowptr, int64_t B, int32_t D, float eps, float* y,
                           float* stats, void* stream) {
  X2_CHECK_ARG(B >= 0 && D >= 4 && D % 4 == 0 && eps >= 0.f,
               "x2_graph_layernorm_fwd: need B >= 0, D a positive multiple of 4, eps >= 0");
  X2_CHECK_ARG(B < (int64_t)1 << 31, "x2_graph_layernorm_fwd: too many graphs");
  if (B == 0) return X2_OK;
  X2_CHECK_ARG(x && rowptr && y && stats, "x2_graph_layernorm_fwd: null pointer");
  X2_CHECK_ARG((((uintptr_t)x | (uintptr_t)y) & 15) == 0, "x2_graph_layernorm_fwd: x / y must be 16-byte aligned");
  k_graph_ln_fwd<<<(unsigned)B, kLnThreads, 0, (cudaStream_t)stream>>>(x, rowptr, D, eps, y, stats);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_graph_layernorm_bwd(const float* y, const float* grad_y, const int32_t* rowptr, int64_t B, int32_t D,
                           const float* stats, float* grad_x, void* stream) {
  X2_CHECK_ARG(B >= 0 && D >= 4 && D % 4 == 0, "x2_graph_layernorm_bwd: need B >= 0, D a positive multiple of 4");
  X2_CHECK_ARG(B < (int64_t)1 << 31, "x2_graph_layernorm_bwd: too many graphs");
  if (B == 0) return X2_OK;
  X2_CHECK_ARG(y && grad_y && rowptr && stats && grad_x, "x2_graph_layernorm_bwd: null pointer");
  X2_CHECK_ARG((((uintptr_t)y | (uintptr_t)grad_y | (uintptr_t)grad_x) & 15) == 0,
               "x2_graph_layernorm_bwd: buffers must be 16-byte aligned");
  k_graph_ln_bwd<<<(unsigned)B, kLnThreads, 0, (cudaStream_t)stream>>>(y, grad_y, rowptr, D, stats, grad_x);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // extern "C"
