// graph.cu -- integer kernels: radius graph (atom_graph.py:32-45), triplet enumeration
// (edge_graph.py:12-30, ordering per SURVEY.md App. E) and the CSR metadata of the line graph
// that the segmented attention kernels consume.  All results are deterministic; integer atomics
// are only used for histograms / slot claims whose outcome is re-ordered by a per-segment sort.
#include "common.cuh"

namespace x2 {

// ------------------------------------------------------------------------------------------
// Per-segment ascending sort of (key[, payload]) in global memory, one thread per segment.
// Insertion sort for short segments (molecular graphs: <= ~100), heapsort otherwise so that a
// pathological hub cannot go quadratic.
template <bool HAS_PAYLOAD>
__device__ void seg_sort(int32_t* key, int32_t* pay, int n) {
  if (n <= 48) {
    for (int i = 1; i < n; ++i) {
      int32_t k = key[i], p = HAS_PAYLOAD ? pay[i] : 0;
      int j = i - 1;
      while (j >= 0 && (key[j] > k || (HAS_PAYLOAD && key[j] == k && pay[j] > p))) {
        key[j + 1] = key[j];
        if (HAS_PAYLOAD) pay[j + 1] = pay[j];
        --j;
      }
      key[j + 1] = k;
      if (HAS_PAYLOAD) pay[j + 1] = p;
    }
    return;
  }
  auto less = [&](int a, int b) {
    return key[a] < key[b] || (HAS_PAYLOAD && key[a] == key[b] && pay[a] < pay[b]);
  };
  auto swp = [&](int a, int b) {
    int32_t t = key[a]; key[a] = key[b]; key[b] = t;
    if (HAS_PAYLOAD) { t = pay[a]; pay[a] = pay[b]; pay[b] = t; }
  };
  auto sift = [&](int root, int end) {
    while (2 * root + 1 < end) {
      int c = 2 * root + 1;
      if (c + 1 < end && less(c, c + 1)) ++c;
      if (less(root, c)) { swp(root, c); root = c; } else return;
    }
  };
  for (int s = n / 2 - 1; s >= 0; --s) sift(s, n);
  for (int e = n - 1; e > 0; --e) { swp(0, e); sift(0, e); }
}

// ------------------------------------------------------------------------------------------
// Radius graph
__device__ __forceinline__ float gram_dist(float xi, float yi, float zi, float hi, float xj,
                                           float yj, float zj, float hj) {
  // atom_graph.py:33-35: relu((H + H^T - 2 G) ** 0.5).  sqrt of a (rounding-)negative radicand
  // is NaN and torch.relu keeps NaN; NaN < cutoff is false, so such pairs never become bonds.
  const float g = fmaf(zi, zj, fmaf(yi, yj, xi * xj));
  const float r = sqrtf((hi + hj) - 2.0f * g);
  return (r > 0.0f || r != r) ? r : 0.0f;
}

__global__ void k_dij(const float* __restrict__ pos, int64_t n, float* __restrict__ dij) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const int64_t i = idx / n, j = idx - i * n;
  const float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
  const float xj = pos[3 * j], yj = pos[3 * j + 1], zj = pos[3 * j + 2];
  const float hi = fmaf(zi, zi, fmaf(yi, yi, xi * xi));
  const float hj = fmaf(zj, zj, fmaf(yj, yj, xj * xj));
  dij[idx] = gram_dist(xi, yi, zi, hi, xj, yj, zj, hj);
}

__device__ __forceinline__ bool is_bond(float d, float cutoff) { return (d < cutoff) && (d != 0.0f); }

// One warp per row of the dense matrix.
__global__ void k_bonds_count(const float* __restrict__ dij, int64_t n, float cutoff,
                              int32_t* __restrict__ cnt) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int c = 0;
  for (int64_t j = lane; j < n; j += 32) c += is_bond(dij[row * n + j], cutoff) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) cnt[row] = c;
}

__global__ void k_bonds_fill(const float* __restrict__ dij, int64_t n, float cutoff,
                             const int32_t* __restrict__ rowptr, int64_t* __restrict__ ei,
                             int64_t E) {
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= n) return;
  int64_t out = rowptr[row];
  for (int64_t j0 = 0; j0 < n; j0 += 32) {
    const int64_t j = j0 + lane;
    const bool p = (j < n) && is_bond(dij[row * n + j], cutoff);
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (p) {
      const int64_t o = out + __popc(m & ((1u << lane) - 1u));
      ei[o] = row;
      ei[E + o] = j;
    }
    out += __popc(m);
  }
}

// Batched, straight from positions: one warp per atom, all atoms of the same graph.
template <bool FILL>
__global__ void k_radius_graph(const float* __restrict__ pos, const int64_t* __restrict__ batch,
                               const int64_t* __restrict__ ptr, int64_t n, float cutoff,
                               int32_t* __restrict__ cnt, const int32_t* __restrict__ rowptr,
                               int64_t* __restrict__ ei, int64_t E) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const int64_t g = batch[i];
  const int64_t lo = ptr[g], hi = ptr[g + 1];
  const float xi = pos[3 * i], yi = pos[3 * i + 1], zi = pos[3 * i + 2];
  const float hi_ = fmaf(zi, zi, fmaf(yi, yi, xi * xi));
  int c = 0;
  int64_t out = FILL ? rowptr[i] : 0;
  for (int64_t j0 = lo; j0 < hi; j0 += 32) {
    const int64_t j = j0 + lane;
    bool p = false;
    if (j < hi) {
      const float xj = pos[3 * j], yj = pos[3 * j + 1], zj = pos[3 * j + 2];
      const float hj = fmaf(zj, zj, fmaf(yj, yj, xj * xj));
      p = is_bond(gram_dist(xi, yi, zi, hi_, xj, yj, zj, hj), cutoff);
    }
    const unsigned m = __ballot_sync(0xffffffffu, p);
    if (FILL) {
      if (p) {
        const int64_t o = out + __popc(m & ((1u << lane) - 1u));
        ei[o] = i;
        ei[E + o] = j;
      }
      out += __popc(m);
    } else {
      c += __popc(m);
    }
  }
  if (!FILL && lane == 0) cnt[i] = c;
}

// ------------------------------------------------------------------------------------------
// Triplets
struct TripWs {
  int32_t *adeg, *astart, *cursor, *acol, *aeid, *skip, *cnt, *flags;
  void* scan;
  size_t scan_bytes;
};

static size_t trip_layout(int64_t E, int64_t N, void* ws, TripWs* w) {
  Arena a(ws, (size_t)-1);
  w->adeg = a.take<int32_t>(N + 1);
  w->astart = a.take<int32_t>(N + 2);
  w->cursor = a.take<int32_t>(N + 1);
  w->acol = a.take<int32_t>(E + 1);
  w->aeid = a.take<int32_t>(E + 1);
  w->skip = a.take<int32_t>(E + 1);
  w->cnt = a.take<int32_t>(E + 1);
  w->flags = a.take<int32_t>(8);
  w->scan_bytes = scan_workspace_bytes(E > N ? E : N);
  w->scan = a.take<char>(w->scan_bytes);
  return align_up(a.off, 256) + 256;
}

// flags[0] = not-sorted count (0 => lexicographically sorted, strictly), flags[1] = out of range
__global__ void k_trip_hist(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                            int32_t* __restrict__ adeg, int32_t* __restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = ei[e], d = ei[E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) {
    atomicAdd(&flags[1], 1);
    return;
  }
  atomicAdd(&adeg[s], 1);
  if (e > 0) {
    const int64_t ps = ei[e - 1], pd = ei[E + e - 1];
    if (ps > s || (ps == s && pd >= d)) atomicAdd(&flags[0], 1);
  }
}

__global__ void k_trip_scatter(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                               const int32_t* __restrict__ astart, int32_t* __restrict__ cursor,
                               int32_t* __restrict__ acol, int32_t* __restrict__ aeid,
                               const int32_t* __restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t s = ei[e], d = ei[E + e];
  if (s < 0 || s >= N || d < 0 || d >= N) return;
  int32_t p;
  if (flags[0] == 0 && flags[1] == 0) p = (int32_t)e;  // sorted input: CSR order == edge order
  else p = astart[s] + atomicAdd(&cursor[s], 1);
  acol[p] = (int32_t)d;
  aeid[p] = (int32_t)e;
}

__global__ void k_trip_sort_rows(int64_t N, const int32_t* __restrict__ astart,
                                 int32_t* __restrict__ acol, int32_t* __restrict__ aeid,
                                 const int32_t* __restrict__ flags) {
  if (flags[0] == 0 && flags[1] == 0) return;
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= N) return;
  const int32_t b = astart[a];
  seg_sort<true>(acol + b, aeid + b, astart[a + 1] - b);
}

// cnt[e] = outdeg(j) - [ (j -> i) exists ];  skip[e] = position of i inside row j or -1.
__global__ void k_trip_count(const int64_t* __restrict__ ei, int64_t E, int64_t N,
                             const int32_t* __restrict__ astart, const int32_t* __restrict__ acol,
                             int32_t* __restrict__ skip, int32_t* __restrict__ cnt) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int64_t i = ei[e], j = ei[E + e];
  if (i < 0 || i >= N || j < 0 || j >= N) { cnt[e] = 0; skip[e] = -1; return; }
  const int32_t b = astart[j], len = astart[j + 1] - b;
  int lo = 0, hi = len;  // lower_bound(i)
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (acol[b + mid] < (int32_t)i) lo = mid + 1; else hi = mid;
  }
  const bool found = lo < len && acol[b + lo] == (int32_t)i;
  skip[e] = found ? lo : -1;
  cnt[e] = len - (found ? 1 : 0);
}

// One warp per target bond e.
__global__ void k_trip_fill(const int64_t* __restrict__ ei, int64_t E, int64_t T,
                            const int32_t* __restrict__ rowptr, const int32_t* __restrict__ astart,
                            const int32_t* __restrict__ acol, const int32_t* __restrict__ aeid,
                            const int32_t* __restrict__ skip, int64_t* __restrict__ tri,
                            int64_t* __restrict__ ej, int64_t* __restrict__ ei_, int64_t* __restrict__ ek) {
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (e >= E) return;
  const int32_t beg = rowptr[e], n = rowptr[e + 1] - beg;
  if (n <= 0) return;
  const int64_t i = ei[e], j = ei[E + e];
  const int32_t b = astart[j], sk = skip[e];
  for (int s = lane; s < n; s += 32) {
    const int pos = b + s + ((sk >= 0 && s >= sk) ? 1 : 0);
    const int64_t t = (int64_t)beg + s;
    tri[t] = aeid[pos];
    tri[T + t] = e;
    ej[t] = j;
    ei_[t] = i;
    ek[t] = acol[pos];
  }
}

// ------------------------------------------------------------------------------------------
// Line-graph CSR metadata
__global__ void k_meta_hist(const int64_t* __restrict__ ei, int64_t T, int64_t E,
                            int32_t* __restrict__ src, int32_t* __restrict__ tgt,
                            int32_t* __restrict__ cnt_t, int32_t* __restrict__ cnt_s,
                            int32_t* __restrict__ flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int64_t s = ei[t], d = ei[T + t];
  const bool bad = (s < 0 || s >= E || d < 0 || d >= E);
  src[t] = bad ? 0 : (int32_t)s;
  tgt[t] = bad ? 0 : (int32_t)d;
  if (bad) { atomicAdd(&flags[1], 1); return; }
  atomicAdd(&cnt_t[d], 1);
  atomicAdd(&cnt_s[s], 1);
  if (t > 0 && ei[T + t - 1] > d) atomicAdd(&flags[0], 1);
}

// key = tgt (BY_TGT) or src; identity fast path when already target-sorted.
__global__ void k_meta_scatter(const int64_t* __restrict__ ei, int64_t T, int64_t E,
                               const int32_t* __restrict__ rp_t, const int32_t* __restrict__ rp_s,
                               int32_t* __restrict__ cur_t, int32_t* __restrict__ cur_s,
                               int32_t* __restrict__ ord_t, int32_t* __restrict__ ord_s,
                               const int32_t* __restrict__ flags) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const int64_t s = ei[t], d = ei[T + t];
  const bool bad = (s < 0 || s >= E || d < 0 || d >= E);
  const bool sorted = flags[0] == 0 && flags[1] == 0;
  if (sorted) ord_t[t] = (int32_t)t;
  if (bad) { if (!sorted) { /* keep arrays defined */ } return; }
  if (!sorted) ord_t[rp_t[d] + atomicAdd(&cur_t[d], 1)] = (int32_t)t;
  ord_s[rp_s[s] + atomicAdd(&cur_s[s], 1)] = (int32_t)t;
}

__global__ void k_meta_sort(int64_t E, const int32_t* __restrict__ rp, int32_t* __restrict__ ord,
                            const int32_t* __restrict__ flags, int skip_if_sorted) {
  if (skip_if_sorted && flags[0] == 0 && flags[1] == 0) return;
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= E) return;
  const int32_t b = rp[a];
  seg_sort<false>(ord + b, nullptr, rp[a + 1] - b);
}

// longest target segment (integer atomicMax: order-independent)
__global__ void k_seg_max(const int32_t* __restrict__ rowptr, int64_t E, int32_t* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int v = e < E ? rowptr[e + 1] - rowptr[e] : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  if ((threadIdx.x & 31) == 0 && v > 0) atomicMax(out, v);
}
__global__ void k_meta_finalize(int32_t* flags_out, const int32_t* flags_in) {
  flags_out[0] = (flags_in[0] == 0 && flags_in[1] == 0) ? 1 : 0;
  flags_out[1] = flags_in[1];
  flags_out[2] = flags_in[2];
  flags_out[3] = 0;
}

// Work items of the fused tile kernels (csrc/tile_attn.cuh): every target segment cut, from its own start, into
// runs of X2_ITEM_ROWS rows.  items[i] = (first triplet, (target << 4) | (rows - 1)), in row order; the entry
// after the last item is the sentinel (T, 0).
__global__ void k_item_count(const int32_t* __restrict__ rowptr, int64_t E, int32_t* __restrict__ cnt) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) cnt[e] = (rowptr[e + 1] - rowptr[e] + X2_ITEM_ROWS - 1) / X2_ITEM_ROWS;
}
__global__ void k_item_fill(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ itemptr, int64_t E,
                            int32_t* __restrict__ items) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e > E) return;
  if (e == E) {                                   // sentinel after the last item
    items[2 * itemptr[E]] = rowptr[E];
    items[2 * itemptr[E] + 1] = 0;
    return;
  }
  const int beg = rowptr[e], end = rowptr[e + 1];
  int i = itemptr[e];
  for (int t = beg; t < end; t += X2_ITEM_ROWS, ++i) {
    items[2 * i] = t;
    items[2 * i + 1] = (int32_t)((e << 4) | (min(X2_ITEM_ROWS, end - t) - 1));
  }
}

__global__ void k_flags_finalize(int32_t* flags_out, const int32_t* flags_in) {
  // user-facing: flags_out[0] = 1 if sorted, flags_out[1] = #out-of-range
  flags_out[0] = (flags_in[0] == 0 && flags_in[1] == 0) ? 1 : 0;
  flags_out[1] = flags_in[1];
}


// ------------------------------------------------------------------------------------------
// Closed blocks of the line graph (csrc/blk_attn.cuh).  A BLOCK is a contiguous range of line-nodes
// [sptr[b], sptr[b+1]) such that every target segment draws all of its sources from ONE block; the
// targets are grouped by that block.  Then every triplet has its source and its target in the same block:
// a CTA that owns a block sees all contributions to dQ of its targets AND to dK / dV of its sources, and
// the K / V (and per-source basis) rows of a block are shared by all of its targets.  For the line graph of
// a sorted atom graph (edge_graph.py:12-30) a block is "all bonds leaving atom j" and its targets are the
// bonds entering j (structural facts) -- but nothing here assumes that: f and f + 1 are put
// in one block when they are adjacent sources of some target segment, and k_blk_tgt then VERIFIES the
// closure, so an arbitrary edge_index either yields valid blocks or ok = 0 (callers keep the generic path).
__global__ void k_blk_link(const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, int64_t T,
                           int32_t* __restrict__ link) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p + 1 >= T) return;
  const int32_t s0 = src[p];
  if (tgt[p] == tgt[p + 1] && src[p + 1] == s0 + 1) link[s0] = 1;     // benign race: every writer stores 1
}
__global__ void k_blk_bnd(const int32_t* __restrict__ link, int64_t E, int32_t* __restrict__ bnd) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f < E) bnd[f] = (f == 0 || link[f - 1] == 0) ? 1 : 0;
}
__global__ void k_blk_ids(const int32_t* __restrict__ bnd, const int32_t* __restrict__ ex, int64_t E,
                          int32_t* __restrict__ blk_of, int32_t* __restrict__ sptr) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= E) return;
  const int32_t b = ex[f] + bnd[f] - 1;
  blk_of[f] = b;
  if (bnd[f]) sptr[b] = (int32_t)f;
  if (f == E - 1) sptr[ex[E]] = (int32_t)E;
}
__global__ void k_blk_tgt(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                          const int32_t* __restrict__ blk_of, int64_t E, int32_t* __restrict__ gt,
                          int32_t* __restrict__ cnt_g, int32_t* __restrict__ trip, int32_t* __restrict__ fl) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int beg = rowptr[e], end = rowptr[e + 1];
  int32_t b = blk_of[e];                         // a target without triplets: any block will do
  if (beg < end) {
    b = blk_of[src[beg]];
    bool bad = false;
    for (int i = beg + 1; i < end; ++i) bad |= blk_of[src[i]] != b;
    if (bad) atomicAdd(&fl[0], 1);
    atomicAdd(&trip[b], end - beg);
  }
  gt[e] = b;
  atomicAdd(&cnt_g[b], 1);
}
__global__ void k_blk_scatter(const int32_t* __restrict__ gt, const int32_t* __restrict__ tptr,
                              int32_t* __restrict__ cur, int64_t E, int32_t* __restrict__ tord) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const int32_t b = gt[e];
  tord[tptr[b] + atomicAdd(&cur[b], 1)] = (int32_t)e;
}
__global__ void k_blk_sort(const int32_t* __restrict__ tptr, int32_t* __restrict__ tord,
                           const int32_t* __restrict__ trip, const int32_t* __restrict__ sptr,
                           const int32_t* __restrict__ nblk, int32_t* __restrict__ tpos, int32_t* __restrict__ fl) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= *nblk) return;
  const int32_t t0 = tptr[b], n = tptr[b + 1] - t0;
  seg_sort<false>(tord + t0, nullptr, n);      // ascending target ids: results do not depend on atomic order
  for (int i = 0; i < n; ++i) tpos[tord[t0 + i]] = i;         // position of a target inside its block
  atomicMax(&fl[2], trip[b]);
  atomicMax(&fl[3], sptr[b + 1] - sptr[b]);
  atomicMax(&fl[4], n);
}
__global__ void k_blk_finalize(int32_t* flags_out, const int32_t* fl, const int32_t* nblk) {
  flags_out[0] = fl[0] == 0 ? 1 : 0;
  flags_out[1] = *nblk;
  flags_out[2] = fl[2];
  flags_out[3] = fl[3];
  flags_out[4] = fl[4];
  flags_out[5] = 0;
}

}  // namespace x2

using namespace x2;

extern "C" {

int x2_dij(const float* pos, int64_t n, float* dij, void* stream) {
  X2_CHECK_ARG(n >= 0 && n < 46340, "x2_dij: n=%lld out of range", (long long)n);
  if (n == 0) return X2_OK;
  const int64_t tot = n * n;
  k_dij<<<(unsigned)cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(pos, n, dij);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_bonds_count(const float* dij, int64_t n, float cutoff, int32_t* rowptr, void* ws,
                   size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(n >= 0, "x2_bonds_count: n < 0");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t need = align_up((size_t)(n + 1) * 4, 256) + scan_workspace_bytes(n);
  if (ws_bytes < need) { set_error("x2_bonds_count: workspace %zu < %zu", ws_bytes, need); return X2_EWORKSPACE; }
  int32_t* cnt = static_cast<int32_t*>(ws);
  void* sws = static_cast<char*>(ws) + align_up((size_t)(n + 1) * 4, 256);
  if (n > 0) {
    k_bonds_count<<<(unsigned)cdiv(n * 32, 256), 256, 0, st>>>(dij, n, cutoff, cnt);
    X2_LAUNCH_OK();
  }
  return exclusive_scan_i32(cnt, rowptr, n, sws, ws_bytes - ((char*)sws - (char*)ws), st);
}

int x2_bonds_fill(const float* dij, int64_t n, float cutoff, const int32_t* rowptr,
                  int64_t* edge_index, int64_t E, void* stream) {
  if (n == 0 || E == 0) return X2_OK;
  k_bonds_fill<<<(unsigned)cdiv(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(dij, n, cutoff, rowptr,
                                                                            edge_index, E);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_radius_graph_count(const float* pos, const int64_t* batch, const int64_t* ptr, int64_t n,
                          float cutoff, int32_t* rowptr, void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(n >= 0, "x2_radius_graph_count: n < 0");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t need = align_up((size_t)(n + 1) * 4, 256) + scan_workspace_bytes(n);
  if (ws_bytes < need) { set_error("x2_radius_graph_count: workspace %zu < %zu", ws_bytes, need); return X2_EWORKSPACE; }
  int32_t* cnt = static_cast<int32_t*>(ws);
  void* sws = static_cast<char*>(ws) + align_up((size_t)(n + 1) * 4, 256);
  if (n > 0) {
    k_radius_graph<false><<<(unsigned)cdiv(n * 32, 256), 256, 0, st>>>(pos, batch, ptr, n, cutoff, cnt,
                                                                     nullptr, nullptr, 0);
    X2_LAUNCH_OK();
  }
  return exclusive_scan_i32(cnt, rowptr, n, sws, ws_bytes - ((char*)sws - (char*)ws), st);
}

int x2_radius_graph_fill(const float* pos, const int64_t* batch, const int64_t* ptr, int64_t n,
                         float cutoff, const int32_t* rowptr, int64_t* edge_index, int64_t E,
                         void* stream) {
  if (n == 0 || E == 0) return X2_OK;
  k_radius_graph<true><<<(unsigned)cdiv(n * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      pos, batch, ptr, n, cutoff, nullptr, rowptr, edge_index, E);
  X2_LAUNCH_OK();
  return X2_OK;
}

size_t x2_triplets_workspace_bytes(int64_t E, int64_t N) {
  TripWs w;
  return trip_layout(E < 0 ? 0 : E, N < 0 ? 0 : N, nullptr, &w);
}

int x2_triplets_count(const int64_t* edge_index, int64_t E, int64_t N, int32_t* rowptr,
                      int32_t* flags, void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(E >= 0 && N >= 0 && E < 2147483647LL && N < 2147483647LL, "x2_triplets_count: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  TripWs w;
  const size_t need = trip_layout(E, N, ws, &w);
  if (ws_bytes < need) { set_error("x2_triplets_count: workspace %zu < %zu", ws_bytes, need); return X2_EWORKSPACE; }
  X2_CUDA_OK(cudaMemsetAsync(w.adeg, 0, (size_t)(N + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(w.cursor, 0, (size_t)(N + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(w.flags, 0, 8 * 4, st));
  const unsigned gE = (unsigned)cdiv(E > 0 ? E : 1, 256), gN = (unsigned)cdiv(N > 0 ? N : 1, 256);
  if (E > 0) { k_trip_hist<<<gE, 256, 0, st>>>(edge_index, E, N, w.adeg, w.flags); X2_LAUNCH_OK(); }
  int rc = exclusive_scan_i32(w.adeg, w.astart, N, w.scan, w.scan_bytes, st);
  if (rc) return rc;
  if (E > 0) {
    k_trip_scatter<<<gE, 256, 0, st>>>(edge_index, E, N, w.astart, w.cursor, w.acol, w.aeid, w.flags);
    X2_LAUNCH_OK();
    if (N > 0) { k_trip_sort_rows<<<gN, 256, 0, st>>>(N, w.astart, w.acol, w.aeid, w.flags); X2_LAUNCH_OK(); }
    k_trip_count<<<gE, 256, 0, st>>>(edge_index, E, N, w.astart, w.acol, w.skip, w.cnt);
    X2_LAUNCH_OK();
  }
  rc = exclusive_scan_i32(w.cnt, rowptr, E, w.scan, w.scan_bytes, st);
  if (rc) return rc;
  k_flags_finalize<<<1, 1, 0, st>>>(flags, w.flags);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_triplets_fill(const int64_t* edge_index, int64_t E, int64_t N, const int32_t* rowptr,
                     int64_t T, int64_t* triplets_index, int64_t* edge_j, int64_t* edge_i,
                     int64_t* edge_k, const void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(E >= 0 && N >= 0 && T >= 0, "x2_triplets_fill: bad sizes");
  if (E == 0 || T == 0) return X2_OK;
  TripWs w;
  const size_t need = trip_layout(E, N, const_cast<void*>(ws), &w);
  if (ws_bytes < need) { set_error("x2_triplets_fill: workspace %zu < %zu", ws_bytes, need); return X2_EWORKSPACE; }
  k_trip_fill<<<(unsigned)cdiv(E * 32, 256), 256, 0, (cudaStream_t)stream>>>(
      edge_index, E, T, rowptr, w.astart, w.acol, w.aeid, w.skip, triplets_index, edge_j, edge_i, edge_k);
  X2_LAUNCH_OK();
  return X2_OK;
}

size_t x2_meta_workspace_bytes(int64_t T, int64_t E) {
  ArenaSize a;
  if (E < 0) E = 0;
  if (T < 0) T = 0;
  a.take<int32_t>(E + 1);  // cnt_t
  a.take<int32_t>(E + 1);  // cnt_s
  a.take<int32_t>(E + 1);  // cur_t
  a.take<int32_t>(E + 1);  // cur_s
  a.take<int32_t>(8);      // flags
  a.take<char>(scan_workspace_bytes(E));
  return a.bytes();
}

int x2_meta_build(const int64_t* edge_index, int64_t T, int64_t E, int32_t* src, int32_t* tgt,
                  int32_t* rowptr_tgt, int32_t* order_tgt, int32_t* rowptr_src,
                  int32_t* order_src, int32_t* flags, void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(E >= 0 && T >= 0 && E < 2147483647LL && T < 2147483647LL, "x2_meta_build: bad sizes");
  cudaStream_t st = (cudaStream_t)stream;
  if (ws_bytes < x2_meta_workspace_bytes(T, E)) { set_error("x2_meta_build: workspace too small"); return X2_EWORKSPACE; }
  Arena a(ws, ws_bytes);
  int32_t* cnt_t = a.take<int32_t>(E + 1);
  int32_t* cnt_s = a.take<int32_t>(E + 1);
  int32_t* cur_t = a.take<int32_t>(E + 1);
  int32_t* cur_s = a.take<int32_t>(E + 1);
  int32_t* fl = a.take<int32_t>(8);
  const size_t sbytes = scan_workspace_bytes(E);
  void* sws = a.take<char>(sbytes);
  X2_CUDA_OK(cudaMemsetAsync(cnt_t, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(cnt_s, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(cur_t, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(cur_s, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(fl, 0, 8 * 4, st));
  const unsigned gT = (unsigned)cdiv(T > 0 ? T : 1, 256), gE = (unsigned)cdiv(E > 0 ? E : 1, 256);
  if (T > 0) { k_meta_hist<<<gT, 256, 0, st>>>(edge_index, T, E, src, tgt, cnt_t, cnt_s, fl); X2_LAUNCH_OK(); }
  int rc = exclusive_scan_i32(cnt_t, rowptr_tgt, E, sws, sbytes, st);
  if (rc) return rc;
  rc = exclusive_scan_i32(cnt_s, rowptr_src, E, sws, sbytes, st);
  if (rc) return rc;
  if (T > 0) {
    k_meta_scatter<<<gT, 256, 0, st>>>(edge_index, T, E, rowptr_tgt, rowptr_src, cur_t, cur_s, order_tgt,
                                      order_src, fl);
    X2_LAUNCH_OK();
    if (E > 0) {
      k_meta_sort<<<gE, 256, 0, st>>>(E, rowptr_tgt, order_tgt, fl, 1);
      X2_LAUNCH_OK();
      k_meta_sort<<<gE, 256, 0, st>>>(E, rowptr_src, order_src, fl, 0);
      X2_LAUNCH_OK();
    }
  }
  if (E > 0) { k_seg_max<<<gE, 256, 0, st>>>(rowptr_tgt, E, fl + 2); X2_LAUNCH_OK(); }
  k_meta_finalize<<<1, 1, 0, st>>>(flags, fl);
  X2_LAUNCH_OK();
  return X2_OK;
}

int64_t x2_items_bound(int64_t T, int64_t E) { return T / X2_ITEM_ROWS + E + 2; }

size_t x2_items_workspace_bytes(int64_t E) {
  return align_up((size_t)(E + 1) * 4, 256) + scan_workspace_bytes(E) + 512;
}

int x2_items_build(const int32_t* rowptr_tgt, int64_t E, int64_t T, int32_t* itemptr, int32_t* items, void* ws,
                   size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(rowptr_tgt && itemptr && items, "x2_items_build: null pointer");
  X2_CHECK_ARG(E > 0 && E < (1LL << 27) && T >= 0, "x2_items_build: needs 0 < E < 2^27 (got %lld)", (long long)E);
  if (ws_bytes < x2_items_workspace_bytes(E)) { set_error("x2_items_build: workspace too small"); return X2_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  int32_t* cnt = a.take<int32_t>(E + 1);
  const size_t sbytes = scan_workspace_bytes(E);
  void* sws = a.take<char>(sbytes);
  k_item_count<<<(unsigned)cdiv(E, 256), 256, 0, st>>>(rowptr_tgt, E, cnt);
  X2_LAUNCH_OK();
  int rc = exclusive_scan_i32(cnt, itemptr, E, sws, sbytes, st);
  if (rc) return rc;
  k_item_fill<<<(unsigned)cdiv(E + 1, 256), 256, 0, st>>>(rowptr_tgt, itemptr, E, items);
  X2_LAUNCH_OK();
  return X2_OK;
}

size_t x2_blocks_workspace_bytes(int64_t T, int64_t E) {
  ArenaSize a;
  if (E < 0) E = 0;
  (void)T;
  for (int i = 0; i < 7; ++i) a.take<int32_t>(E + 1);   // link, bnd, ex, blk_of, gt, cnt/cur, trip
  a.take<int32_t>(E + 1);
  a.take<int32_t>(8);
  a.take<char>(scan_workspace_bytes(E));
  return a.bytes();
}

int x2_blocks_build(const int32_t* src, const int32_t* tgt, const int32_t* rowptr_tgt, int64_t T, int64_t E,
                    int32_t* blk_sptr, int32_t* blk_tptr, int32_t* blk_tord, int32_t* blk_tpos, int32_t* flags, void* ws,
                    size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(src && tgt && rowptr_tgt && blk_sptr && blk_tptr && blk_tord && blk_tpos && flags, "x2_blocks_build: null pointer");
  X2_CHECK_ARG(E > 0 && T >= 0 && E < 2147483647LL && T < 2147483647LL, "x2_blocks_build: bad sizes");
  if (ws_bytes < x2_blocks_workspace_bytes(T, E)) { set_error("x2_blocks_build: workspace too small"); return X2_EWORKSPACE; }
  cudaStream_t st = (cudaStream_t)stream;
  Arena a(ws, ws_bytes);
  int32_t* link = a.take<int32_t>(E + 1);
  int32_t* bnd = a.take<int32_t>(E + 1);
  int32_t* ex = a.take<int32_t>(E + 1);
  int32_t* blk_of = a.take<int32_t>(E + 1);
  int32_t* gt = a.take<int32_t>(E + 1);
  int32_t* cnt = a.take<int32_t>(E + 1);
  int32_t* cur = a.take<int32_t>(E + 1);
  int32_t* trip = a.take<int32_t>(E + 1);
  int32_t* fl = a.take<int32_t>(8);
  const size_t sbytes = scan_workspace_bytes(E);
  void* sws = a.take<char>(sbytes);
  X2_CUDA_OK(cudaMemsetAsync(link, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(cnt, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(cur, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(trip, 0, (size_t)(E + 1) * 4, st));
  X2_CUDA_OK(cudaMemsetAsync(fl, 0, 8 * 4, st));
  const unsigned gT = (unsigned)cdiv(T > 0 ? T : 1, 256), gE = (unsigned)cdiv(E, 256);
  if (T > 1) { k_blk_link<<<gT, 256, 0, st>>>(src, tgt, T, link); X2_LAUNCH_OK(); }
  k_blk_bnd<<<gE, 256, 0, st>>>(link, E, bnd);
  X2_LAUNCH_OK();
  int rc = exclusive_scan_i32(bnd, ex, E, sws, sbytes, st);
  if (rc) return rc;
  k_blk_ids<<<gE, 256, 0, st>>>(bnd, ex, E, blk_of, blk_sptr);
  X2_LAUNCH_OK();
  k_blk_tgt<<<gE, 256, 0, st>>>(rowptr_tgt, src, blk_of, E, gt, cnt, trip, fl);
  X2_LAUNCH_OK();
  rc = exclusive_scan_i32(cnt, blk_tptr, E, sws, sbytes, st);
  if (rc) return rc;
  k_blk_scatter<<<gE, 256, 0, st>>>(gt, blk_tptr, cur, E, blk_tord);
  X2_LAUNCH_OK();
  k_blk_sort<<<gE, 256, 0, st>>>(blk_tptr, blk_tord, trip, blk_sptr, ex + E, blk_tpos, fl);
  X2_LAUNCH_OK();
  k_blk_finalize<<<1, 1, 0, st>>>(flags, fl, ex + E);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // extern "C"
