// umma_ts_probe.cu -- development probe: tcgen05.mma kind::tf32 with the A operand in TENSOR MEMORY.
// A(m,k) is written with tcgen05.st (thread = lane/row m, 8 consecutive 32-bit columns = k 0..7);
// B is an identity-like selector in the known-good K-major SW128 smem layout, so D(m, n<8) = A(m, n).
// Also probes how the tensor core rounds fp32 operand bits to tf32 (A values with low mantissa bits).
#include <cstdio>
#include <vector>
#include "../x2-gnn_b200/csrc/tc_gemm.cuh"
using namespace x2::tc;
namespace x2 { void set_error(const char*, ...) {} void count_launch() {} }

__global__ void probe(int mode, float* out) {
  __shared__ __align__(1024) uint32_t sI[4096];    // selector: rows n (N=32), K-major SW128
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, lane = tid & 31, q = tid >> 5;
  for (int w = tid; w < 4096; w += blockDim.x) sI[w] = 0;
  __syncthreads();
  if (tid < 8) sI[(kmajor_off(tid, tid >> 2) + (tid & 3) * 4) >> 2] = __float_as_uint(1.0f);   // B(k=tid, n=tid) = 1
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 64);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  // A operand at columns [32, 40): thread (row m = q*32+lane) writes 8 values
  {
    uint32_t v[8];
    const int m = q * 32 + lane;
    for (int j = 0; j < 8; ++j) {
      float f;
      if (mode == 0) f = (float)(m * 8 + j);                       // layout check (exact in tf32 up to 1023)
      else {                                                       // rounding check: 1 + j * 2^-13 (+ 2^-11 steps)
        f = 1.0f + (float)(j + 1) * 0.0001220703125f;              // (j+1) * 2^-13 : below tf32 resolution 2^-10
      }
      v[j] = __float_as_uint(f);
    }
    const uint32_t taddr = tm + ((uint32_t)(q * 32) << 16) + 32;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint64_t dI = make_desc(smem_u32(sI), 16, 1024);
    const uint32_t idesc = make_idesc(32, 0, 0);
    const uint32_t a_tmem = tm + 32, d_tmem = tm;
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(dI),
        "r"(idesc), "r"(0)
        : "memory");
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 32; c0 += 16) {
    float v[16];
    tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) out[(q * 32 + lane) * 32 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tm, 64);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  std::vector<float> h(128 * 32);
  for (int mode = 0; mode < 2; ++mode) {
    cudaMemset(d, 0xff, 128 * 32 * 4);
    probe<<<1, 128>>>(mode, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h.data(), d, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    if (mode == 0) {
      int bad = 0;
      for (int m = 0; m < 128; ++m) for (int n = 0; n < 8; ++n) if (h[m * 32 + n] != (float)(m * 8 + n)) ++bad;
      printf("mode 0 (A in TMEM layout): mismatches = %d of 1024\n", bad);
      for (int m : {0, 1, 33, 127}) { printf("  D(m=%3d, n=0..9):", m); for (int n = 0; n < 10; ++n) printf(" %6.0f", h[m * 32 + n]); printf("\n"); }
    } else {
      printf("mode 1 (fp32 -> tf32 operand rounding), D - 1 in units of 2^-13 for inputs 1 + (j+1) 2^-13:\n  ");
      for (int n = 0; n < 8; ++n) printf(" in=%d->%g", n + 1, (h[n] - 1.0f) * 8192.0f);
      printf("\n  (truncation gives 0,0,0,0,0,0,0,8 ; round-to-nearest gives 0,0,0,8 or 0,0,0,0/8,8,8,8,8)\n");
    }
  }
  return 0;
}
