// development probe: which setmaxnreg splits of an 896-thread CTA (72 registers at launch) run
#include <cstdio>
#include <cstdint>
#include <cstdlib>
template <int TR, int PR, int AT>
__global__ void __launch_bounds__(896, 1) k(const float* __restrict__ in, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5;
#ifdef WITH_STACK
  float loc[32];                       // dynamically indexed: lives in local memory
  for (int i = 0; i < 32; ++i) loc[i] = in[(threadIdx.x + i) & 1023];
  out[threadIdx.x + 896] = loc[(int)in[threadIdx.x] & 31];
#endif
#ifdef WITH_TMEM
  __shared__ uint32_t slot;
  if (warp == 12) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  __syncthreads();
#endif
  if (warp >= 12) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(AT));
    out[threadIdx.x] = in[threadIdx.x] * 2.f;
  } else if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(TR));
    out[threadIdx.x] = in[threadIdx.x];
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PR));
    out[threadIdx.x] = in[threadIdx.x] + 1.f;
  }
#ifdef WITH_TMEM
  __syncthreads();
  if (warp == 12) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
#endif
}
static int g_mode = 0;   // 1: 230 KB dynamic smem; 2: + programmatic stream serialization attribute
template <int TR, int PR, int AT>
void run() {
  float *a, *b;
  cudaMalloc(&a, 4096); cudaMalloc(&b, 4096);
  cudaMemset(a, 0, 4096);
  if (g_mode == 0) k<TR, PR, AT><<<2, 896>>>(a, b);
  else {
    cudaFuncSetAttribute(k<TR, PR, AT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(896); cfg.dynamicSmemBytes = 230000; cfg.stream = 0;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_mode == 2 ? 1 : 0;
    k<TR, PR, AT><<<148, 896, 230000>>>(a, b);       // a predecessor on the stream
    cudaLaunchKernelEx(&cfg, k<TR, PR, AT>, (const float*)a, b);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("transposers %d producers %d attention %d (sum %d): %s\n", TR, PR, AT, 128 * TR + 256 * PR + 512 * AT, cudaGetErrorString(e));
  if (e != cudaSuccess) exit(1);
}
int main(int argc, char** argv) {
  const int which = argc > 1 ? atoi(argv[1]) : 0;
  g_mode = argc > 2 ? atoi(argv[2]) : 0;
  if (which == 0) run<40, 40, 96>();
  if (which == 1) run<48, 40, 88>();
  if (which == 2) run<56, 40, 88>();
  if (which == 3) run<48, 40, 96>();
  if (which == 4) run<40, 40, 88>();
  if (which == 5) run<72, 72, 72>();
  fflush(stdout);
  return 0;
}
