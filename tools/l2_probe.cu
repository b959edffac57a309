// l2_probe.cu -- development micro-benchmark (not part of the library): how much of a freshly
// WRITTEN buffer is still in L2 when the next kernel reads it back?  Decides the triplet-chunk size
// for the chunked layer schedule (producer kernel -> consumer kernel per chunk).
//   mode 0: stream over a big arena, chunk c written then read (distinct addresses every chunk)
//   mode 1: the same scratch chunk is written then read every iteration
// Prints time of the write kernel, the read kernel and GB/s of each per chunk size.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void k_write(float4* __restrict__ dst, size_t n4, float v) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = make_float4(v, v, v, v);
}
__global__ void k_read(const float4* __restrict__ src, size_t n4, float* out) {
  float acc = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t st = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * st < n4; i += 4 * st) {
    float4 a = src[i], b = src[i + st], c = src[i + 2 * st], d = src[i + 3 * st];
    acc += a.x + b.y + c.z + d.w;
  }
  for (; i < n4; i += st) acc += src[i].x;
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const size_t arena = (size_t)4 << 30;
  float4* buf;
  float* out;
  cudaMalloc(&buf, arena);
  cudaMalloc(&out, 4);
  cudaMemset(buf, 0, arena);
  cudaEvent_t e[4];
  for (auto& x : e) cudaEventCreate(&x);
  const int grid = 148 * 4, block = 512;
  for (int mode = 0; mode < 2; ++mode)
    for (int mb : {8, 16, 24, 32, 48, 64, 80, 96, 128, 192, 256, 512}) {
      const size_t bytes = (size_t)mb << 20, n4 = bytes / 16;
      const int nchunk = (int)(arena / bytes) < 64 ? (int)(arena / bytes) : 64;
      float tw = 0, tr = 0;
      for (int rep = 0; rep < 2; ++rep) {
        tw = tr = 0;
        for (int c = 0; c < nchunk; ++c) {
          float4* p = buf + (mode == 0 ? (size_t)c * n4 : 0);
          cudaEventRecord(e[0]);
          k_write<<<grid, block>>>(p, n4, (float)c);
          cudaEventRecord(e[1]);
          k_read<<<grid, block>>>(p, n4, out);
          cudaEventRecord(e[2]);
          cudaEventSynchronize(e[2]);
          float a, b;
          cudaEventElapsedTime(&a, e[0], e[1]);
          cudaEventElapsedTime(&b, e[1], e[2]);
          tw += a;
          tr += b;
        }
      }
      printf("mode %d chunk %4d MB: write %7.1f GB/s  read-back %7.1f GB/s  (pair %.1f us)\n", mode, mb,
             bytes * nchunk / (tw * 1e6), bytes * nchunk / (tr * 1e6), (tw + tr) * 1e3 / nchunk);
    }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}
