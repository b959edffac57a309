// api.cu -- version / error / device checks and the int32 exclusive scan shared by the
// integer kernels.
#include "common.cuh"

#include <atomic>
#include <mutex>
#include <vector>

namespace x2 {

// ------------------------------------------------------------------ instrumentation
static std::atomic<long long> g_launches{0};
bool pdl_enabled() {
  static const int on = [] { const char* v = getenv("X2GNN_PDL"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}

int ensure_dyn_smem(const void* func, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<const void*, int>> done;      // (kernel, device) pairs already set
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("cudaGetDevice -> %s", cudaGetErrorString(e)); return X2_ECUDA; }
  std::lock_guard<std::mutex> lock(mu);
  for (const auto& d : done)
    if (d.first == func && d.second == dev) return X2_OK;
  e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) { set_error("cudaFuncSetAttribute(%d bytes) -> %s", bytes, cudaGetErrorString(e)); return X2_ECUDA; }
  done.emplace_back(func, dev);
  return X2_OK;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

static std::atomic<int> g_timing{0};
static std::mutex g_tmutex;
struct Span { cudaEvent_t a, b; int phase; };
static std::vector<Span> g_spans;          // recorded spans awaiting read()
static std::vector<cudaEvent_t> g_pool;    // recycled events
static thread_local cudaEvent_t t_open = nullptr;
constexpr size_t kMaxSpans = 1 << 16;

static cudaEvent_t take_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return e;
}

void phase_begin(cudaStream_t st) {
  if (!g_timing.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_tmutex);
  if (g_spans.size() >= kMaxSpans) return;
  if (t_open == nullptr) t_open = take_event();
  if (t_open) cudaEventRecord(t_open, st);
}

// Closes the span opened by the previous phase_begin/phase_end on this thread and opens the next.
void phase_end(int phase, cudaStream_t st) {
  if (!g_timing.load(std::memory_order_relaxed) || t_open == nullptr) return;
  std::lock_guard<std::mutex> lk(g_tmutex);
  cudaEvent_t e = take_event();
  if (!e) return;
  cudaEventRecord(e, st);
  g_spans.push_back(Span{t_open, e, phase});
  t_open = take_event();
  if (t_open) cudaEventRecord(t_open, st);
}

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// ------------------------------------------------------------------ scan
// Three-phase scan: per-block sums -> single-block scan of the sums -> per-block rescan.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ int32_t warp_incl_scan(int32_t v) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  return v;
}

// Exclusive scan across the block of one value per thread; returns the exclusive prefix and
// (optionally) the block total.  blockDim.x must be a multiple of 32, <= 1024.
__device__ __forceinline__ int32_t block_excl_scan(int32_t v, int32_t* total) {
  __shared__ int32_t wsum[33];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int32_t inc = warp_incl_scan(v);
  if (lane == 31) wsum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const int32_t w = lane < nw ? wsum[lane] : 0;
    const int32_t winc = warp_incl_scan(w);
    wsum[lane] = winc - w;           // exclusive offset of warp `lane`
    if (lane == 31) wsum[32] = winc;  // block total
  }
  __syncthreads();
  const int32_t excl = wsum[wid] + inc - v;
  if (total) *total = wsum[32];
  __syncthreads();  // wsum is reused by the next call
  return excl;
}

__global__ void k_scan_block_sums(const int32_t* __restrict__ in, int32_t* __restrict__ bsum,
                                  int64_t n) {
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + (int64_t)i * kScanThreads + threadIdx.x;
    if (idx < n) s += in[idx];
  }
  // block reduce
  __shared__ int32_t red[kScanThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int32_t t = 0;
    for (int i = 0; i < kScanThreads / 32; ++i) t += red[i];
    bsum[blockIdx.x] = t;
  }
}

// Single block: exclusive scan of bsum[nb] in place; bsum[nb] = total.
__global__ void k_scan_sums(int32_t* bsum, int64_t nb) {
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += blockDim.x) {
    int64_t idx = base + threadIdx.x;
    int32_t v = idx < nb ? bsum[idx] : 0;
    int32_t tot;
    int32_t ex = block_excl_scan(v, &tot);
    int32_t carry = carry_s;
    if (idx < nb) bsum[idx] = ex + carry;
    __syncthreads();
    if (threadIdx.x == 0) carry_s = carry + tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) bsum[nb] = carry_s;
}

__global__ void k_scan_final(const int32_t* __restrict__ in, int32_t* __restrict__ out,
                             const int32_t* __restrict__ bsum, int64_t n, int64_t nb) {
  // thread t owns kScanItems CONSECUTIVE items so the in-thread prefix is sequential
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  int32_t s = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    v[i] = idx < n ? in[idx] : 0;
    s += v[i];
  }
  int32_t ex = block_excl_scan(s, nullptr) + bsum[blockIdx.x];
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    int64_t idx = base + i;
    if (idx < n) out[idx] = ex;
    ex += v[i];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = bsum[nb];
}

size_t scan_workspace_bytes(int64_t n) {
  int64_t nb = cdiv(n > 0 ? n : 1, kScanTile);
  return align_up((size_t)(nb + 1) * sizeof(int32_t), 256) + 256;
}

int exclusive_scan_i32(const int32_t* in, int32_t* out, int64_t n, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
  if (n < 0) { set_error("scan: n < 0"); return X2_EINVAL; }
  if (ws_bytes < scan_workspace_bytes(n)) { set_error("scan: workspace too small"); return X2_EWORKSPACE; }
  int32_t* bsum = static_cast<int32_t*>(ws);
  int64_t nb = cdiv(n > 0 ? n : 1, kScanTile);
  if (n == 0) {
    X2_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int32_t), stream));
    return X2_OK;
  }
  k_scan_block_sums<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, bsum, n);
  X2_LAUNCH_OK();
  k_scan_sums<<<1, 1024, 0, stream>>>(bsum, nb);
  X2_LAUNCH_OK();
  k_scan_final<<<(unsigned)nb, kScanThreads, 0, stream>>>(in, out, bsum, n, nb);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // namespace x2

extern "C" {

int x2_version(void) { return 103; }  // 0.1.2: x2_graph_layernorm_*

int64_t x2_launch_count(void) { return (int64_t)x2::g_launches.load(); }

int x2_timing_enable(int on) {
  x2::g_timing.store(on ? 1 : 0);
  return X2_OK;
}

int x2_timing_read(double* ms, int64_t* calls, int n) {
  std::lock_guard<std::mutex> lk(x2::g_tmutex);
  for (auto& s : x2::g_spans) {
    float t = 0.f;
    X2_CUDA_OK(cudaEventSynchronize(s.b));
    X2_CUDA_OK(cudaEventElapsedTime(&t, s.a, s.b));
    if (s.phase >= 0 && s.phase < n) {
      if (ms) ms[s.phase] += (double)t;
      if (calls) calls[s.phase] += 1;
    }
    x2::g_pool.push_back(s.a);
    x2::g_pool.push_back(s.b);
  }
  x2::g_spans.clear();
  return X2_OK;
}

const char* x2_timing_phase_name(int phase) {
  static const char* names[X2_NUM_PHASES] = {"node_proj", "trow_proj", "attn_fwd", "attn_bwd_tgt",
                                             "attn_bwd_src", "trow_dgrad", "trow_wgrad", "node_bwd"};
  return (phase >= 0 && phase < X2_NUM_PHASES) ? names[phase] : "?";
}

const char* x2_last_error(void) { return x2::g_err; }

int x2_device_check(int device) {
  int count = 0;
  X2_CUDA_OK(cudaGetDeviceCount(&count));
  X2_CHECK_ARG(device >= 0 && device < count, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp p;
  X2_CUDA_OK(cudaGetDeviceProperties(&p, device));
  if (p.major != 10) {
    x2::set_error("device %d is sm_%d%d; libx2gnn is built for sm_100a (B200) only", device, p.major, p.minor);
    return X2_EDEVICE;
  }
  return X2_OK;
}

size_t x2_scan_workspace_bytes(int64_t n) { return x2::scan_workspace_bytes(n); }

}  // extern "C"
