// collate.cu -- device-side batch collation.  The reference collates on the host: a PyG
// DataLoader concatenates the per-molecule records of qm9_allprop.py:18 (x, atom_pos, edge_index, edge_attr[E,338],
// edge_num) in Python, adds the per-graph atom offset to edge_index and builds `batch`, then copies ~58 MB per
// batch of 128 to the device.  Here the dataset lives on the device in CSR form (atoms and bonds of molecule m are
// rows atom_ptr[m] .. atom_ptr[m+1] / edge_ptr[m] .. edge_ptr[m+1] of flat arrays) and a batch is a gather:
//   x2_collate_sizes : counts of the selected molecules -> exclusive offsets aoff[B+1], eoff[B+1]
//   x2_collate_fill  : one CTA row per selected molecule copies its atoms and streams its bond rows (128-bit)
// Bit-exact against the host collation (tests/test_gpu_graph.py).
#include "common.cuh"

namespace x2 {
namespace {

__global__ void k_collate_counts(const int64_t* __restrict__ ids, int64_t B, const int64_t* __restrict__ atom_ptr,
                                 const int64_t* __restrict__ edge_ptr, int64_t M, int32_t* __restrict__ acnt,
                                 int32_t* __restrict__ ecnt, int32_t* __restrict__ flags) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t m = ids[b];
  if (m < 0 || m >= M) {
    atomicAdd(&flags[0], 1);
    acnt[b] = ecnt[b] = 0;
    return;
  }
  acnt[b] = (int32_t)(atom_ptr[m + 1] - atom_ptr[m]);
  ecnt[b] = (int32_t)(edge_ptr[m + 1] - edge_ptr[m]);
}

constexpr int kColThreads = 256;
__global__ void __launch_bounds__(kColThreads)
k_collate_fill(const int64_t* __restrict__ ids, int64_t M, const int64_t* __restrict__ atom_ptr,
               const int64_t* __restrict__ edge_ptr, const int64_t* __restrict__ z_all,
               const float* __restrict__ pos_all, const float* __restrict__ feat_all, int F,
               const int64_t* __restrict__ ei_all, int64_t Etot_all, const int32_t* __restrict__ aoff,
               const int32_t* __restrict__ eoff, int64_t* __restrict__ z, float* __restrict__ pos,
               int64_t* __restrict__ batch, int64_t* __restrict__ edge_num, float* __restrict__ feat,
               int64_t* __restrict__ edge_index, int64_t E_total) {
  const int b = blockIdx.x;
  const int64_t m = ids[b];
  if (m < 0 || m >= M) return;
  const int64_t a0 = atom_ptr[m], e0 = edge_ptr[m];
  const int na = aoff[b + 1] - aoff[b], ne = eoff[b + 1] - eoff[b];
  const int64_t ao = aoff[b], eo = eoff[b];
  const int tid = blockIdx.y * kColThreads + threadIdx.x, nthr = gridDim.y * kColThreads;
  if (blockIdx.y == 0 && threadIdx.x == 0) edge_num[b] = ne;
  for (int i = tid; i < na; i += nthr) {
    z[ao + i] = z_all[a0 + i];
    batch[ao + i] = b;
    pos[(ao + i) * 3 + 0] = pos_all[(a0 + i) * 3 + 0];
    pos[(ao + i) * 3 + 1] = pos_all[(a0 + i) * 3 + 1];
    pos[(ao + i) * 3 + 2] = pos_all[(a0 + i) * 3 + 2];
  }
  if (edge_index) {                                  // local atom ids + the graph's atom offset (PyG collate)
    for (int i = tid; i < ne; i += nthr) {
      edge_index[eo + i] = ei_all[e0 + i] + ao;
      edge_index[E_total + eo + i] = ei_all[Etot_all + e0 + i] + ao;
    }
  }
  // bond feature rows: contiguous in the store and in the batch
  const int64_t nfl = (int64_t)ne * F;
  const float* src = feat_all + e0 * F;
  float* dst = feat + eo * F;
  if ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) {
    const int64_t n4 = nfl >> 2;
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int64_t i = tid; i < n4; i += nthr) d4[i] = __ldg(s4 + i);
    for (int64_t i = (n4 << 2) + tid; i < nfl; i += nthr) dst[i] = src[i];
  } else {
    for (int64_t i = tid; i < nfl; i += nthr) dst[i] = src[i];
  }
}

}  // namespace
}  // namespace x2

using namespace x2;

extern "C" {

size_t x2_collate_workspace_bytes(int64_t B) {
  if (B < 0) B = 0;
  return 2 * align_up((size_t)(B + 1) * 4, 256) + scan_workspace_bytes(B) + 512;
}

int x2_collate_sizes(const int64_t* ids, int64_t B, const int64_t* atom_ptr, const int64_t* edge_ptr, int64_t M,
                     int32_t* aoff, int32_t* eoff, int32_t* flags, void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(ids && atom_ptr && edge_ptr && aoff && eoff && flags, "x2_collate_sizes: null pointer");
  X2_CHECK_ARG(B >= 0 && M >= 0 && B < 2147483647LL, "x2_collate_sizes: bad sizes");
  if (ws_bytes < x2_collate_workspace_bytes(B)) { set_error("x2_collate_sizes: workspace too small"); return X2_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  char* w = static_cast<char*>(ws);
  const size_t seg = align_up((size_t)(B + 1) * 4, 256);
  int32_t* acnt = reinterpret_cast<int32_t*>(w);
  int32_t* ecnt = reinterpret_cast<int32_t*>(w + seg);
  void* sws = w + 2 * seg;
  const size_t sbytes = ws_bytes - 2 * seg;
  X2_CUDA_OK(cudaMemsetAsync(flags, 0, 4 * sizeof(int32_t), st));
  if (B > 0) {
    k_collate_counts<<<(unsigned)cdiv(B, 256), 256, 0, st>>>(ids, B, atom_ptr, edge_ptr, M, acnt, ecnt, flags);
    X2_LAUNCH_OK();
  }
  int rc = exclusive_scan_i32(acnt, aoff, B, sws, sbytes, st);
  if (rc) return rc;
  return exclusive_scan_i32(ecnt, eoff, B, sws, sbytes, st);
}

int x2_collate_fill(const int64_t* ids, int64_t B, int64_t M, const int64_t* atom_ptr, const int64_t* edge_ptr,
                    const int64_t* z_all, const float* pos_all, const float* feat_all, int32_t F,
                    const int64_t* ei_all, int64_t Etot_all, const int32_t* aoff, const int32_t* eoff, int64_t* z,
                    float* pos, int64_t* batch, int64_t* edge_num, float* feat, int64_t* edge_index, int64_t E_total,
                    void* stream) {
  X2_CHECK_ARG(ids && atom_ptr && edge_ptr && z_all && pos_all && feat_all && aoff && eoff && z && pos && batch &&
                   edge_num && feat, "x2_collate_fill: null pointer");
  X2_CHECK_ARG(B >= 0 && F >= 1 && (edge_index == nullptr || ei_all != nullptr), "x2_collate_fill: bad arguments");
  if (B == 0) return X2_OK;
  k_collate_fill<<<dim3((unsigned)B, 8), kColThreads, 0, (cudaStream_t)stream>>>(
      ids, M, atom_ptr, edge_ptr, z_all, pos_all, feat_all, F, ei_all, Etot_all, aoff, eoff, z, pos, batch, edge_num,
      feat, edge_index, E_total);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // extern "C"
