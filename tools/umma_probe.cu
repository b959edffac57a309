// umma_probe.cu -- development probe (not part of the library): issues ONE tcgen05.mma
// (kind::tf32, M=128, N=32, K=8) where one operand is an identity-like selector in the known-good
// K-major layout, so D reveals which shared-memory word the hardware reads for each logical element
// of the OTHER operand under a given (major, LBO, SBO) descriptor.  Used to pin down the MN-major
// SWIZZLE_128B descriptor semantics for csrc/tc_gemm.cuh (k_tc_wgrad).
#include <cstdio>
#include <vector>
#include "../x2-gnn_b200/csrc/tc_gemm.cuh"

using namespace x2::tc;
namespace x2 { void set_error(const char*, ...) {} void count_launch() {} }

// test_b = 0: probe operand A (D(m,n) = A(m,k=n), n<8).  test_b = 1: probe operand B (D(m,n) = B(k=m,n), m<8)
__global__ void probe(int test_b, int mn, uint32_t lbo, uint32_t sbo, uint32_t layout, int mode, float* out) {
  __shared__ __align__(1024) uint32_t sP[6144];    // probed operand region: 32 KB of tagged words
  __shared__ __align__(1024) uint32_t sI[4096];    // identity-like selector, K-major SW128, 128 rows x 32
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int w = tid; w < 6144; w += blockDim.x) {
    float v = mode == 0 ? (float)((w & 1023) + 1) : (float)((w >> 10) + 1);
    sP[w] = __float_as_uint(v);
  }
  for (int w = tid; w < 4096; w += blockDim.x) sI[w] = 0;
  __syncthreads();
  if (tid < 8) {  // selector(row r = tid, k = tid) = 1
    uint32_t off = kmajor_off(tid, tid >> 2) + (tid & 3) * 4;
    sI[off >> 2] = __float_as_uint(1.0f);
  }
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 32);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (tid == 0) {
    const uint64_t dP = make_desc(smem_u32(sP), lbo, sbo, layout);
    const uint64_t dI = make_desc(smem_u32(sI), 16, 1024);
    const uint32_t idesc = test_b ? make_idesc(32, 0, mn) : make_idesc(32, mn, 0);
    if (test_b) umma_tf32(tm, dI, dP, idesc, 0);
    else umma_tf32(tm, dP, dI, idesc, 0);
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  const int q = (tid >> 5) & 3, lane = tid & 31;
  for (int c0 = 0; c0 < 32; c0 += 16) {
    float v[16];
    tmem_ld16(tm + ((uint32_t)(q * 32) << 16) + c0, v);
    for (int j = 0; j < 16; ++j) out[(q * 32 + lane) * 32 + c0 + j] = v[j];
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tm, 32);
}

int main() {
  float* d;
  cudaMalloc(&d, 128 * 32 * 4);
  struct Cfg { int test_b, mn; uint32_t lbo, sbo, layout; const char* name; };
  std::vector<Cfg> cfgs = {
      {0, 1, 4096, 512, 1, "A MN-major BASE32B lbo=4096 sbo=512"},
      {0, 1, 512, 4096, 1, "A MN-major BASE32B lbo=512 sbo=4096"},
      {1, 1, 4096, 512, 1, "B MN-major BASE32B lbo=4096 sbo=512"},
      {0, 1, 4096, 1024, 2, "A MN-major SW128 lbo=4096 sbo=1024 (expected unsupported)"},
  };
  for (auto& c : cfgs) {
    std::vector<float> lo(128 * 32), hi(128 * 32);
    for (int mode = 0; mode < 2; ++mode) {
      cudaMemset(d, 0xff, 128 * 32 * 4);
      probe<<<1, 128>>>(c.test_b, c.mn, c.lbo, c.sbo, c.layout, mode, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(mode == 0 ? lo.data() : hi.data(), d, 128 * 32 * 4, cudaMemcpyDeviceToHost);
    }
    printf("== %s : byte offset of the word the hardware read\n", c.name);
    if (!c.test_b) {   // D(m, n<8) = A(m, k=n)
      for (int m : {0, 1, 2, 3, 4, 7, 8, 9, 16, 24, 31, 32, 33, 64, 96, 127}) {
        printf("  A(m=%3d, k=0..7):", m);
        for (int k = 0; k < 8; ++k) printf(" %6d", ((int)(hi[m * 32 + k] - 1) * 1024 + (int)(lo[m * 32 + k] - 1)) * 4);
        printf("\n");
      }
    } else {           // D(m<8, n) = B(k=m, n)
      for (int k = 0; k < 8; ++k) {
        printf("  B(k=%d, n=0,1,2,3,4,7,8,16,31):", k);
        for (int n : {0, 1, 2, 3, 4, 7, 8, 16, 31}) printf(" %6d", ((int)(hi[k * 32 + n] - 1) * 1024 + (int)(lo[k * 32 + n] - 1)) * 4);
        printf("\n");
      }
    }
  }
  return 0;
}
