// bw_probe.cu -- development micro-benchmark (not part of the library): read bandwidth of ONE CTA per
// SM as a function of warps per CTA and independent 128-bit loads in flight per thread, for plain
// LDG and for cp.async.bulk (UBLKCP) into shared memory.  Decides how the GEMM producers must fetch.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int U>
__global__ void k_ldg(const float4* __restrict__ src, size_t n4, float* out) {
  // each CTA streams its own contiguous range; each thread keeps U loads in flight
  const size_t per_cta = n4 / gridDim.x;
  const float4* p = src + (size_t)blockIdx.x * per_cta;
  float acc = 0.f;
  for (size_t i = threadIdx.x; i + (size_t)(U - 1) * blockDim.x < per_cta; i += (size_t)U * blockDim.x) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = __ldg(p + i + (size_t)u * blockDim.x);
#pragma unroll
    for (int u = 0; u < U; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// one thread issues bulk copies of `chunk` bytes into a ring of `stages` buffers; all threads then
// touch the buffer (one LDS each) and release it.
__global__ void k_bulk(const uint8_t* __restrict__ src, size_t bytes, int chunk, int stages, float* out) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  uint8_t* buf = sm + 256;
  const size_t per_cta = bytes / gridDim.x / chunk * chunk;
  const uint8_t* p = src + (size_t)blockIdx.x * per_cta;
  const int n = (int)(per_cta / chunk);
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[i])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[i])), "r"(blockDim.x - 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  float acc = 0.f;
  if (threadIdx.x < 32) {
    if (threadIdx.x == 0) {
      for (int c = 0; c < n; ++c) {
        const int st = c % stages, ph = (c / stages) & 1;
        asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(s32(&empty[st])), "r"(ph ^ 1) : "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(chunk) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(s32(buf + (size_t)st * chunk)), "l"(p + (size_t)c * chunk), "r"(chunk), "r"(s32(&full[st])) : "memory");
      }
    }
  } else {
    for (int c = 0; c < n; ++c) {
      const int st = c % stages, ph = (c / stages) & 1;
      asm volatile("{.reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(s32(&full[st])), "r"(ph) : "memory");
      acc += reinterpret_cast<const float*>(buf + (size_t)st * chunk)[threadIdx.x];
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
    }
  }
  if (acc == 123.456f) out[0] = acc;
}

template <typename F>
float time_ms(F f) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a);
  for (int i = 0; i < 5; ++i) f();
  cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  return ms / 5;
}

int main() {
  const size_t bytes = (size_t)2 << 30;   // 2 GiB, far larger than L2
  uint8_t* d; float* out;
  cudaMalloc(&d, bytes); cudaMalloc(&out, 4);
  cudaMemset(d, 1, bytes);
  const size_t n4 = bytes / 16;
  printf("LDG.128: 148 CTAs (1 per SM)\n");
  for (int warps : {4, 8, 16, 32}) {
    float ms;
    ms = time_ms([&] { k_ldg<2><<<148, warps * 32>>>((const float4*)d, n4, out); });  printf("  warps=%2d U=2  (%3d KB in flight/SM): %6.0f GB/s\n", warps, warps * 32 * 2 * 16 / 1024, bytes / ms / 1e6);
    ms = time_ms([&] { k_ldg<4><<<148, warps * 32>>>((const float4*)d, n4, out); });  printf("  warps=%2d U=4  (%3d KB in flight/SM): %6.0f GB/s\n", warps, warps * 32 * 4 * 16 / 1024, bytes / ms / 1e6);
    ms = time_ms([&] { k_ldg<8><<<148, warps * 32>>>((const float4*)d, n4, out); });  printf("  warps=%2d U=8  (%3d KB in flight/SM): %6.0f GB/s\n", warps, warps * 32 * 8 * 16 / 1024, bytes / ms / 1e6);
    ms = time_ms([&] { k_ldg<16><<<148, warps * 32>>>((const float4*)d, n4, out); }); printf("  warps=%2d U=16 (%3d KB in flight/SM): %6.0f GB/s\n", warps, warps * 32 * 16 * 16 / 1024, bytes / ms / 1e6);
  }
  printf("LDG.128 with a large dynamic smem allocation (L1 shrinks): 148 CTAs, 16 warps, U=4\n");
  for (int kb : {0, 64, 128, 160, 192, 212, 227}) {
    cudaFuncSetAttribute(k_ldg<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    float ms = time_ms([&] { k_ldg<4><<<148, 512, (size_t)kb * 1024>>>((const float4*)d, n4, out); });
    printf("  smem=%3d KB: %6.0f GB/s\n", kb, bytes / ms / 1e6);
  }
  printf("LDG.128: 296 / 592 CTAs of 8 warps, U=8\n");
  for (int g : {296, 592}) {
    float ms = time_ms([&] { k_ldg<8><<<g, 256>>>((const float4*)d, n4, out); });
    printf("  grid=%d: %6.0f GB/s\n", g, bytes / ms / 1e6);
  }
  printf("cp.async.bulk: 148 CTAs, 160 threads\n");
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int chunk : {4096, 16384, 32768}) for (int stages : {2, 4, 8}) {
    if ((size_t)chunk * stages > 190 * 1024) continue;
    float ms = time_ms([&] { k_bulk<<<148, 160, 256 + chunk * stages>>>(d, bytes, chunk, stages, out); });
    printf("  chunk=%5d B stages=%d (%3d KB in flight/SM): %6.0f GB/s\n", chunk, stages, chunk * stages / 1024, bytes / ms / 1e6);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return 0;
}
