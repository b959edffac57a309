// umma_rate.cu -- development probe: issue rate of tcgen05.mma (M = 128) on one SM as a function of
//   kind (tf32 / bf16), A operand source (tensor memory "TS" / shared memory "SS"), N, and the number of
//   accumulators the instruction stream rotates over (1 = every MMA accumulates into the same columns).
// One thread issues `nmma` MMAs back to back, commits, waits; cycles = clock64 around issue + completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate tools/umma_rate.cu && ./umma_rate
#include <cstdio>
#include <vector>
#include "../x2-gnn_b200/csrc/tc_gemm.cuh"
using namespace x2::tc;
namespace x2 { void set_error(const char*, ...) {} void count_launch() {} int ensure_dyn_smem(const void*, int) { return 0; } bool pdl_enabled() { return false; } }

// warp-collective issue (every lane executes, one elected lane issues): keeps descriptors in uniform registers,
// see umma_tf32_ts_w in tc_gemm.cuh -- under `if (tid == 0)` every MMA costs ~200 cycles of issue waterfall
template <int KIND, int TS>
__device__ __forceinline__ void mma_w(uint32_t leader, uint32_t d, uint32_t a_t, uint64_t a_d, uint64_t b_d,
                                      uint32_t idesc, uint32_t acc) {
  if constexpr (KIND == 0 && TS == 1)
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 q, %5, 0;\nsetp.ne.b32 p, %4, 0;\n@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc), "r"(leader) : "memory");
  if constexpr (KIND == 0 && TS == 0)
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 q, %5, 0;\nsetp.ne.b32 p, %4, 0;\n@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc), "r"(acc), "r"(leader) : "memory");
  if constexpr (KIND == 1 && TS == 1)
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 q, %5, 0;\nsetp.ne.b32 p, %4, 0;\n@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n"
                 ::"r"(d), "r"(a_t), "l"(b_d), "r"(idesc), "r"(acc), "r"(leader) : "memory");
  if constexpr (KIND == 1 && TS == 0)
    asm volatile("{\n.reg .pred p, q;\nsetp.ne.b32 q, %5, 0;\nsetp.ne.b32 p, %4, 0;\n@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
                 ::"r"(d), "l"(a_d), "l"(b_d), "r"(idesc), "r"(acc), "r"(leader) : "memory");
}

template <int KIND, int TS, int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate(int nmma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = smem;              // up to 256 rows x 128 B
  uint8_t* sA = smem + 32768;      // 128 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < (32768 + 16384) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3f800000u;
  if (tid == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  {  // define the A region (columns 448..511) so no NaN garbage is multiplied
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = 0x3f800000u;
    const uint32_t taddr = tm + ((uint32_t)((tid >> 5) * 32) << 16) + 448;
    for (int c = 0; c < 64; c += 8) tmem_st8(taddr + c, v);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = 0, t1 = 0, t2 = 0;
  if (tid < 32) {
    const uint32_t leader = elect_one();
    uint32_t idesc = make_idesc(N, 0, 0);        // a_format / b_format: TF32 = 2, BF16 = 1
    if (KIND == 1) idesc = (idesc & ~((7u << 7) | (7u << 10))) | (1u << 7) | (1u << 10);
    const uint64_t bd = make_desc(smem_u32(sB), 16, 1024);
    const uint64_t ad = make_desc(smem_u32(sA), 16, 1024);
    t0 = clock64();
    for (int i = 0; i < nmma; i += 4 * NACC) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
#pragma unroll
        for (int a = 0; a < NACC; ++a)
          mma_w<KIND, TS>(leader, tm + a * N, tm + 448 + ks * 8, ad + ks * 2, bd + ks * 2, idesc, 1);
    }
    t1 = clock64();
    umma_commit_w(leader, &bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (tid == 0) {
    t2 = clock64();
    out[blockIdx.x * 2] = t1 - t0;
    out[blockIdx.x * 2 + 1] = t2 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (tid < 32) tmem_dealloc(tm, 512);
}

template <int KIND, int TS, int N, int NACC>
void run(int grid, long long* d) {
  const int nmma = 3072;
  auto k = rate<KIND, TS, N, NACC>;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024);
  k<<<grid, 128, 52 * 1024>>>(nmma, d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s (kind %d ts %d N %d)\n", cudaGetErrorString(e), KIND, TS, N); exit(1); }
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  const double ci = (double)h[0] / nmma, cc = (double)h[1] / nmma;
  const double mac = 128.0 * N * (KIND == 0 ? 8 : 16) / cc;
  printf("%s %s %4d %d %4d | %8.1f %8.1f | %8.0f\n", KIND == 0 ? "tf32" : "bf16", TS ? "TMEM" : "smem", N, NACC, grid, ci, cc, mac);
}
template <int KIND, int TS>
void sweep(int grid, long long* d) {
  run<KIND, TS, 16, 1>(grid, d); run<KIND, TS, 16, 2>(grid, d);
  run<KIND, TS, 32, 1>(grid, d); run<KIND, TS, 32, 2>(grid, d); run<KIND, TS, 32, 3>(grid, d);
  run<KIND, TS, 64, 1>(grid, d); run<KIND, TS, 64, 2>(grid, d);
  run<KIND, TS, 128, 1>(grid, d); run<KIND, TS, 128, 2>(grid, d); run<KIND, TS, 128, 3>(grid, d);
  run<KIND, TS, 256, 1>(grid, d);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * 8);
  printf("kind A-src   N nacc grid | cycles/MMA (issue) cycles/MMA (complete) | MAC/cycle/SM\n");
  for (int grid : {1, 148}) {
    sweep<0, 1>(grid, d); sweep<0, 0>(grid, d); sweep<1, 1>(grid, d); sweep<1, 0>(grid, d);
  }
  return 0;
}
