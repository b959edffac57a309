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

void count_launch();
#define X2_LAUNCH_OK()                 \
  do {                                 \
    x2::count_launch();                \
    X2_CUDA_OK(cudaGetLastError());    \
  } while (0)

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
