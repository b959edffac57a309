// common.cuh -- shared host/device helpers for libx2gnn (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/x2gnn.h"

namespace x2 {

constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs

void set_error(const char* fmt, ...);

#define X2_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      x2::set_error(__VA_ARGS__);               \
      return X2_EINVAL;                         \
    }                                           \
  } while (0)

#define X2_CUDA_OK(expr)                                                          \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      x2::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return X2_ECUDA;                                                            \
    }                                                                             \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute of a kernel: ensure_dyn_smem sets it once
// per (kernel, current device) under a mutex (a process-global `static bool` would leave every device but
// the first without the opt-in size, and races between host threads).
int ensure_dyn_smem(const void* func, int bytes);
#define X2_DYN_SMEM(kernel, bytes) X2_CUDA_OK_RC(x2::ensure_dyn_smem(reinterpret_cast<const void*>(kernel), (int)(bytes)))
#define X2_CUDA_OK_RC(expr)               \
  do {                                    \
    int _rc = (expr);                     \
    if (_rc != X2_OK) return _rc;         \
  } while (0)

void count_launch();
#define X2_LAUNCH_OK()                 \
  do {                                 \
    x2::count_launch();                \
    X2_CUDA_OK(cudaGetLastError());    \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// The kernels of one conv step are launched back to back on one stream.  launch_k() sets the programmatic
// stream-serialisation attribute, and every kernel launched through it calls pdl_sync() before its first
// access to global memory: `griddepcontrol.wait` blocks until the preceding grid has completed and its
// writes are visible.  The preceding grid never triggers early (implicit trigger when its CTAs exit), so the
// effect is that the launch processing of a kernel overlaps the tail of its predecessor instead of
// following its completion.  Measured on the layer step (same box, 30 steps): 1.524 / 1.537 ms without
// the attribute, 1.484 / 1.486 ms with it.  An explicit `griddepcontrol.launch_dependents` at the top of
// every kernel was SLOWER than no PDL at all (1.536 vs 1.504 ms): early-resident CTAs of the next
// persistent kernel sit on SMs that the predecessor's tail could still use.  Kernels launched the ordinary
// way are unaffected (the wait is a no-op for them, and an ordinary launch waits for full completion).
bool pdl_enabled();           // X2GNN_PDL=0 turns the attribute off (A/B runs)
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                            Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);   // errors surface in X2_LAUNCH_OK()
}
#endif

// CUDA-event phase brackets (no-ops unless x2_timing_enable(1)).
void phase_begin(cudaStream_t st);
void phase_end(int phase, cudaStream_t st);

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), off(0) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};
// Same arithmetic without a buffer, for *_workspace_bytes.
struct ArenaSize {
  size_t off = 0;
  template <typename T>
  void take(size_t n) {
    off = align_up(off, 256);
    off += n * sizeof(T);
  }
  size_t bytes() const { return align_up(off, 256) + 256; }
};

// ---- exclusive scan of int32 (out has n+1 entries; out[n] = total) -------------------------
size_t scan_workspace_bytes(int64_t n);
int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes,
                       cudaStream_t stream);

}  // namespace x2
