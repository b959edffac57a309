// conv.cu -- SBFTransformerConv forward / backward (sbftransformer_conv.py:93-162 + the PyG
// propagate / softmax / sum-aggregate it drives), X2_MODE_FP32 path.
//
// Forward:   rbf filter -> one batched node GEMM (Q | K | V | skip) -> T-row projections
//            (lin_edge, lin_sbf) -> ONE segmented attention kernel, warp per target line-node,
//            that gathers K/V rows, forms logits, runs an online segment softmax, modulates the
//            values by the sbf gate and aggregates -- deterministic and atomic-free (replaces 3
//            index_selects, ~8 scatter-softmax launches, 3 elementwise passes and an atomicAdd
//            scatter in the reference).
// Backward:  pass 1 (warp per target) recomputes the attention weights from the saved
//            log-sum-exp and emits dQ, d(lin_edge out), d(lin_sbf out); pass 2 (warp per SOURCE
//            line-node, using the source-sorted permutation) accumulates dK/dV without atomics;
//            the Linear gradients are dgrad / split-K wgrad GEMMs with fixed-order reductions.
#include "common.cuh"
#include "gemm_simt.cuh"
#include "tc_gemm.cuh"

namespace x2 {

// ------------------------------------------------------------------ small vector helpers
template <int VEC>
__device__ __forceinline__ void ldv(const float* __restrict__ p, float (&r)[VEC]) {
  if constexpr (VEC == 8) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w; r[4] = b.x; r[5] = b.y; r[6] = b.z; r[7] = b.w;
  } else if constexpr (VEC == 4) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    r[0] = a.x; r[1] = a.y; r[2] = a.z; r[3] = a.w;
  } else if constexpr (VEC == 2) {
    const float2 a = *reinterpret_cast<const float2*>(p);
    r[0] = a.x; r[1] = a.y;
  } else {
    r[0] = p[0];
  }
}
template <int VEC>
__device__ __forceinline__ void stv(float* __restrict__ p, const float (&r)[VEC]) {
  if constexpr (VEC == 8) {
    *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(r[4], r[5], r[6], r[7]);
  } else if constexpr (VEC == 4) {
    *reinterpret_cast<float4*>(p) = make_float4(r[0], r[1], r[2], r[3]);
  } else if constexpr (VEC == 2) {
    *reinterpret_cast<float2*>(p) = make_float2(r[0], r[1]);
  } else {
    p[0] = r[0];
  }
}

// sum over the `lph` (power of two) adjacent lanes that share a head
__device__ __forceinline__ float head_sum(float v, int lph) {
  for (int o = 1; o < lph; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPH>
__device__ __forceinline__ float head_sum_t(float v, int lph) {
  if constexpr (LPH > 0) {
#pragma unroll
    for (int o = 1; o < LPH; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  } else {
    return head_sum(v, lph);
  }
}

}  // namespace x2
#include "tile_attn.cuh"     // fused T-scale forward (uses head_sum_t)
namespace x2 {

// Attention-dropout keep factor for (triplet, head): 0 or 1/(1-p); counter-based so forward and
// backward regenerate the same mask.
__device__ __forceinline__ float keep_scale(uint64_t seed, int64_t t, int h, int H, float p) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(t * H + h + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = (float)(z >> 40) * (1.0f / 16777216.0f);
  return u < p ? 0.f : 1.0f / (1.0f - p);
}

// ------------------------------------------------------------------ rbf filter
// xs[e,:] = x[e,:] * (rbf[e,:] . W_r^T)   (sbftransformer_conv.py:99-100); optionally also F.
// One warp per row e: lane r holds rbf[e,r] (broadcast by shuffle), every lane produces 4 consecutive
// channels per step from a shared-memory copy of W_r; x / xs / F move as 128-bit accesses.
constexpr int kFilterRows = 8;   // rows (warps) per block
__global__ void __launch_bounds__(kFilterRows * 32)
k_rbf_filter(const float* __restrict__ x, const float* __restrict__ rbf, const float* __restrict__ w_rbf,
             int64_t E, int D, int R, float* __restrict__ xs, float* __restrict__ F) {
  extern __shared__ __align__(16) float s_w[];     // W_r transposed to [R][D]: lanes read consecutive words
  pdl_sync();
  for (int i = threadIdx.x; i < D * R; i += blockDim.x) s_w[(i % R) * D + i / R] = w_rbf[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  // grid-stride over rows: the W_r copy above is paid once per resident block, not once per 8 rows
  for (int64_t e = (int64_t)blockIdx.x * kFilterRows + (threadIdx.x >> 5); e < E;
       e += (int64_t)gridDim.x * kFilterRows) {
  float rv[2];                                     // R <= 64: lane holds rbf[e, lane] and rbf[e, lane + 32]
  rv[0] = lane < R ? rbf[e * R + lane] : 0.f;
  rv[1] = lane + 32 < R ? rbf[e * R + lane + 32] : 0.f;
  for (int d0 = lane * 4; d0 < D; d0 += 128) {
    float f[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < R; ++r) {
      const float b = __shfl_sync(0xffffffffu, rv[r >> 5], r & 31);
      const float4 wv = *reinterpret_cast<const float4*>(s_w + r * D + d0);
      f[0] = fmaf(b, wv.x, f[0]); f[1] = fmaf(b, wv.y, f[1]);
      f[2] = fmaf(b, wv.z, f[2]); f[3] = fmaf(b, wv.w, f[3]);
    }
    const float4 xv = *reinterpret_cast<const float4*>(x + e * D + d0);
    *reinterpret_cast<float4*>(xs + e * D + d0) = make_float4(xv.x * f[0], xv.y * f[1], xv.z * f[2], xv.w * f[3]);
    if (F) *reinterpret_cast<float4*>(F + e * D + d0) = make_float4(f[0], f[1], f[2], f[3]);
  }
  }
}
static inline unsigned filter_grid(int64_t E) {
  const int64_t need = cdiv(E, kFilterRows);
  const int64_t cap = (int64_t)kNumSM * 8;
  return (unsigned)(need < cap ? (need > 0 ? need : 1) : cap);
}

// dx += dxs * F ; dF = dxs * x (written over dxs)       (App. A last line)
__global__ void k_filter_bwd(const float* __restrict__ x, const float* __restrict__ F,
                             float* __restrict__ dxs, float* __restrict__ dx, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = dxs[i];
  dx[i] += g * F[i];
  dxs[i] = g * x[i];
}
// Same, with the sums the grouped dgrad launch leaves open folded in (fixed order):
// dxs = dxs + dxs2 ; dx = (dx + dx2) + dxs * F ; dF = dxs * x.  n4 = number of float4 (dx2 may be NULL).
__global__ void k_filter_bwd_sum(const float4* __restrict__ x, const float4* __restrict__ F,
                                 float4* __restrict__ dxs, const float4* __restrict__ dxs2,
                                 float4* __restrict__ dx, const float4* __restrict__ dx2, int64_t n4) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = dxs[i], b = dxs2[i], f = F[i], xv = x[i];
  float4 o = dx[i];
  if (dx2) { const float4 o2 = dx2[i]; o.x += o2.x; o.y += o2.y; o.z += o2.z; o.w += o2.w; }
  const float4 g = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  dx[i] = make_float4(o.x + g.x * f.x, o.y + g.y * f.y, o.z + g.z * f.z, o.w + g.w * f.w);
  dxs[i] = make_float4(g.x * xv.x, g.y * xv.y, g.z * xv.z, g.w * xv.w);
}

// The whole tail of the backward for D = 128 in one pass over the rows (replaces k_filter_bwd_sum, the
// F recompute, the d rbf GEMM and the dW_r weight-gradient GEMM + reduction):
//   F = rbf W_r^T (recomputed) ; g = dxs + dxs2 ; dx = (dx + dx2) + g * F ; dF = g * x ;
//   drbf[e, :] = dF W_r ; dW_r[d, r] = sum_e dF[e, d] rbf[e, r].
// Warp per row (grid-stride, fixed row -> warp assignment), lane owns 4 channels; the dW_r terms are
// accumulated in registers, combined over the block's warps in fixed order and written as one partial
// tile per block (summed afterwards by k_splitk_reduce): deterministic, fp32-exact.
template <int RMAX>
__global__ void __launch_bounds__(kFilterRows * 32)
k_filter_bwd_full(const float* __restrict__ x, const float* __restrict__ rbf, const float* __restrict__ w_rbf,
                  const float* __restrict__ dxs, const float* __restrict__ dxs2, float* __restrict__ dx,
                  const float* __restrict__ dx2, int64_t E, int R, float* __restrict__ drbf,
                  float* __restrict__ partial) {
  constexpr int D = 128;
  extern __shared__ __align__(16) float s_w[];     // [R][D] transposed W_r, then [warps][RMAX][D] for the reduction
  pdl_sync();
  for (int i = threadIdx.x; i < D * R; i += blockDim.x) s_w[(i % R) * D + i / R] = w_rbf[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d0 = lane * 4;
  float4 wv[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; ++r)
    wv[r] = r < R ? *reinterpret_cast<const float4*>(s_w + r * D + d0) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 acc[RMAX];
#pragma unroll
  for (int r = 0; r < RMAX; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int64_t e = (int64_t)blockIdx.x * kFilterRows + warp; e < E; e += (int64_t)gridDim.x * kFilterRows) {
    const float rv = lane < R ? rbf[e * R + lane] : 0.f;       // RMAX <= 32: lane r holds rbf[e, r]
    const float4 a = *reinterpret_cast<const float4*>(dxs + e * D + d0);
    const float4 b = *reinterpret_cast<const float4*>(dxs2 + e * D + d0);
    const float4 xv = *reinterpret_cast<const float4*>(x + e * D + d0);
    float4 o = *reinterpret_cast<const float4*>(dx + e * D + d0);
    if (dx2) {
      const float4 o2 = *reinterpret_cast<const float4*>(dx2 + e * D + d0);
      o.x += o2.x; o.y += o2.y; o.z += o2.z; o.w += o2.w;
    }
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      const float br = __shfl_sync(0xffffffffu, rv, r);
      f.x = fmaf(br, wv[r].x, f.x); f.y = fmaf(br, wv[r].y, f.y);
      f.z = fmaf(br, wv[r].z, f.z); f.w = fmaf(br, wv[r].w, f.w);
    }
    const float4 g = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    *reinterpret_cast<float4*>(dx + e * D + d0) =
        make_float4(o.x + g.x * f.x, o.y + g.y * f.y, o.z + g.z * f.z, o.w + g.w * f.w);
    const float4 dF = make_float4(g.x * xv.x, g.y * xv.y, g.z * xv.z, g.w * xv.w);
    float mine = 0.f;                                           // lane r ends up with drbf[e, r]
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
      const float br = __shfl_sync(0xffffffffu, rv, r);
      acc[r].x = fmaf(dF.x, br, acc[r].x); acc[r].y = fmaf(dF.y, br, acc[r].y);
      acc[r].z = fmaf(dF.z, br, acc[r].z); acc[r].w = fmaf(dF.w, br, acc[r].w);
      float p = dF.x * wv[r].x + dF.y * wv[r].y + dF.z * wv[r].z + dF.w * wv[r].w;
#pragma unroll
      for (int o_ = 16; o_ > 0; o_ >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o_);
      if (lane == r) mine = p;
    }
    if (lane < R) drbf[e * R + lane] = mine;
  }
  // block partial of dW_r: warps summed in fixed order
  __syncthreads();                                   // everyone is done with the W_r copy
  float* red = s_w;                                  // [warps][RMAX][D]
#pragma unroll
  for (int r = 0; r < RMAX; ++r) *reinterpret_cast<float4*>(red + (warp * RMAX + r) * D + d0) = acc[r];
  __syncthreads();
  for (int i = threadIdx.x; i < D * R; i += blockDim.x) {
    const int d = i / R, r = i - d * R;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < kFilterRows; ++w) t += red[(w * RMAX + r) * D + d];
    partial[(int64_t)blockIdx.x * D * R + i] = t;    // [block][D][R], the layout of dW_r
  }
}

// ------------------------------------------------------------------ segmented attention fwd
// lin_edge term of a triplet: none (edge_dim=None), one row per triplet (reference layout), or one row
// per target segment (segment-constant edge features, SURVEY.md §8f row 1: the row is loaded once).
enum { kEaNone = 0, kEaTriplet = 1, kEaSegment = 2 };
constexpr int kRing = 2;      // forward: register ring of triplet rows, kRing - 1 triplets of loads in flight ahead
constexpr int kRingBwd = 1;   // by-target backward: more rows in registers cost more occupancy than they hide latency

// GENERAL = attention dropout and/or the alpha output requested (rare paths; the plain instantiation
// has no branches in the triplet loop, so the loads of the next triplets are issued ahead).
template <int VEC, int EA, bool GENERAL, int LPH>
__global__ void __launch_bounds__(128, 8)
k_attn_fwd(const float* __restrict__ qkvs, int ldq, const float* __restrict__ ea,
           const int32_t* __restrict__ ea_index,
           const float* __restrict__ sg, const int32_t* __restrict__ src,
           const int32_t* __restrict__ rowptr, const int32_t* __restrict__ order, int64_t E, int H,
           int C, float scale, int fuse_skip, float dropout_p, uint64_t seed,
           float* __restrict__ attn, float* __restrict__ out, float* __restrict__ lse,
           float* __restrict__ alpha) {
  pdl_sync();
  constexpr int D = 32 * VEC;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;
  const int lane = threadIdx.x & 31;
  const int ch = lane * VEC;
  const int head = ch / C;
  const int lph = LPH > 0 ? LPH : C / VEC;      // lanes per head: compile-time for the common shapes
  const bool leader = (ch % C) == 0;

  float q[VEC];
  ldv<VEC>(qkvs + e * ldq + ch, q);
  const int beg = rowptr[e], end = rowptr[e + 1];

  float m = -INFINITY, z = 0.f;
  float acc[VEC], a_seg[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = a_seg[i] = 0.f;
  if constexpr (EA == kEaSegment) ldv<VEC>(ea + (int64_t)ea_index[e] * D + ch, a_seg);

  struct Row {
    float k[VEC], v[VEC], g[VEC], a[EA == kEaTriplet ? VEC : 1];
    int t;
  };
  for (int base = beg; base < end; base += 32) {
    const int my = base + lane;
    int t_l = 0, s_l = 0;
    if (my < end) {
      t_l = order ? order[my] : my;            // order == NULL: the triplet list is target-sorted
      s_l = src[t_l];
    }
    const int cnt = min(32, end - base);
    auto fetch = [&](Row& r, int i) {          // i is clamped: re-reading the last triplet is harmless
      i = min(i, cnt - 1);
      r.t = __shfl_sync(0xffffffffu, t_l, i);
      const int s = __shfl_sync(0xffffffffu, s_l, i);
      ldv<VEC>(qkvs + (int64_t)s * ldq + D + ch, r.k);
      ldv<VEC>(qkvs + (int64_t)s * ldq + 2 * D + ch, r.v);
      ldv<VEC>(sg + (int64_t)r.t * D + ch, r.g);
      if constexpr (EA == kEaTriplet) ldv<VEC>(ea + (int64_t)r.t * D + ch, r.a);
    };
    auto consume = [&](const Row& r) {
      float a_[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) a_[j] = EA == kEaTriplet ? r.a[EA == kEaTriplet ? j : 0] : a_seg[j];
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < VEC; ++j) dot = fmaf(q[j], r.k[j] + a_[j], dot);
      const float a = head_sum_t<LPH>(dot, lph) * scale;    // :150
      if constexpr (GENERAL) {
        if (alpha && leader) alpha[(int64_t)r.t * H + head] = a;  // raw logit, normalised below
      }
      const float mn = fmaxf(m, a);
      const float corr = expf(m - mn);                    // exp(-inf) = 0 on the first triplet
      const float p = expf(a - mn);
      z = z * corr + p;
      float pk = p;
      if constexpr (GENERAL) {
        if (dropout_p > 0.f) pk *= keep_scale(seed, r.t, head, H, dropout_p);
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] = acc[j] * corr + pk * (r.v[j] + a_[j]) * r.g[j];  // :155-160
      m = mn;
    };
    Row ring[kRing];
#pragma unroll
    for (int u = 0; u < kRing - 1; ++u) fetch(ring[u], u);
    for (int i = 0; i < cnt; i += kRing) {
#pragma unroll
      for (int u = 0; u < kRing; ++u) {
        if (i + u < cnt) {
          fetch(ring[(u + kRing - 1) % kRing], i + u + kRing - 1);
          consume(ring[u]);
        }
      }
    }
  }
  const float inv = 1.0f / (z + 1e-16f);                    // PyG softmax: out / (sum + 1e-16)
  float o[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) o[i] = acc[i] * inv;
  stv<VEC>(attn + e * D + ch, o);
  if (fuse_skip) {
    float sk[VEC];
    ldv<VEC>(qkvs + e * ldq + 3 * D + ch, sk);
#pragma unroll
    for (int i = 0; i < VEC; ++i) o[i] += sk[i];            // :127
  }
  stv<VEC>(out + e * D + ch, o);
  const float l = (end > beg) ? m + logf(z) : 0.f;
  if (leader) lse[e * H + head] = l;
  if constexpr (GENERAL) {
    if (alpha) {
      __syncwarp();
      for (int idx = beg; idx < end; ++idx) {
        const int t = order ? order[idx] : idx;
        if (leader) {
          const float a = alpha[(int64_t)t * H + head];
          alpha[(int64_t)t * H + head] = expf(a - m) * inv;
        }
      }
    }
  }
}

// ------------------------------------------------------------------ forward, bulk-copy staged variant
// Same arithmetic as k_attn_fwd (plain instantiation), for target-sorted triplet lists: the Sg (and EA)
// rows of a segment are then CONTIGUOUS in memory, so lane 0 of the warp streams them into a per-warp
// shared-memory ring with cp.async.bulk (one instruction per kStRows rows, completion on an mbarrier)
// while the lanes gather K / V with ordinary loads.  The streamed operand no longer occupies the
// warp's load slots or registers.
constexpr int kStRows = 4;          // rows per bulk copy (divides 32)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc::smem_u32(bar)), "r"(bytes) : "memory");
}

template <int VEC, int EA, int LPH, int DEPTH>
__global__ void __launch_bounds__(128, 7)
k_attn_fwd_stage(const float* __restrict__ qkvs, int ldq, const float* __restrict__ ea,
                 const int32_t* __restrict__ ea_index, const float* __restrict__ sg,
                 const int32_t* __restrict__ src, const int32_t* __restrict__ rowptr, int64_t E, int H, int C,
                 float scale, int fuse_skip, float* __restrict__ attn, float* __restrict__ out,
                 float* __restrict__ lse) {
  pdl_sync();
  constexpr int D = 32 * VEC;
  constexpr int NARR = EA == kEaTriplet ? 2 : 1;
  constexpr int STAGE = NARR * kStRows * D;            // floats per stage
  extern __shared__ __align__(128) float st_smem[];    // [4 warps][DEPTH][NARR][kStRows][D] | barriers
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* wbuf = st_smem + (size_t)warp * DEPTH * STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(st_smem + (size_t)4 * DEPTH * STAGE) + warp * DEPTH;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < DEPTH; ++i) tc::mbar_init(&bars[i], 1);
    tc::fence_barrier_init();
  }
  __syncwarp();
  const int ch = lane * VEC;
  const int head = ch / C;
  const int lph = LPH > 0 ? LPH : C / VEC;
  const bool leader = (ch % C) == 0;
  const int beg = rowptr[e], end = rowptr[e + 1];
  const int n = end - beg;
  const int nchunks = (n + kStRows - 1) / kStRows;
  auto issue = [&](int c) {                            // lane 0: rows [c kStRows, ...) of the segment
    const int stg = c % DEPTH;
    const int rows = min(kStRows, n - c * kStRows);
    const uint32_t bytes = (uint32_t)rows * D * 4;
    mbar_expect_tx(&bars[stg], bytes * NARR);
    const int64_t t0 = (int64_t)beg + c * kStRows;
    bulk_g2s(wbuf + stg * STAGE, sg + t0 * D, bytes, &bars[stg]);
    if constexpr (EA == kEaTriplet) bulk_g2s(wbuf + stg * STAGE + kStRows * D, ea + t0 * D, bytes, &bars[stg]);
  };
  if (lane == 0) {
#pragma unroll
    for (int c = 0; c < DEPTH; ++c)
      if (c < nchunks) issue(c);
  }
  float q[VEC];
  ldv<VEC>(qkvs + e * ldq + ch, q);
  float m = -INFINITY, z = 0.f;
  float acc[VEC], a_[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) acc[i] = a_[i] = 0.f;
  if constexpr (EA == kEaSegment) ldv<VEC>(ea + (int64_t)ea_index[e] * D + ch, a_);

  int s_l = 0;
  for (int c = 0; c < nchunks; ++c) {
    const int r0 = c * kStRows;
    if ((r0 & 31) == 0) s_l = (beg + r0 + lane < end) ? src[beg + r0 + lane] : 0;
    const int stg = c % DEPTH;
    const int rows = min(kStRows, n - r0);
    // gather K / V of the chunk's rows first (independent loads), then wait for the streamed rows
    float k[kStRows][VEC], v[kStRows][VEC];
#pragma unroll
    for (int i = 0; i < kStRows; ++i) {
      const int s = __shfl_sync(0xffffffffu, s_l, (r0 + min(i, rows - 1)) & 31);
      ldv<VEC>(qkvs + (int64_t)s * ldq + D + ch, k[i]);
      ldv<VEC>(qkvs + (int64_t)s * ldq + 2 * D + ch, v[i]);
    }
    tc::mbar_wait(&bars[stg], (uint32_t)(c / DEPTH) & 1u);
    const float* rowbuf = wbuf + stg * STAGE;
#pragma unroll
    for (int i = 0; i < kStRows; ++i) {
      float g[VEC];
      ldv<VEC>(rowbuf + i * D + ch, g);
      if constexpr (EA == kEaTriplet) ldv<VEC>(rowbuf + kStRows * D + i * D + ch, a_);
      float dot = 0.f;
#pragma unroll
      for (int j = 0; j < VEC; ++j) dot = fmaf(q[j], k[i][j] + a_[j], dot);
      const float a = head_sum_t<LPH>(dot, lph) * scale;
      if (i < rows) {
        const float mn = fmaxf(m, a);
        const float corr = expf(m - mn);
        const float p = expf(a - mn);
        z = z * corr + p;
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = acc[j] * corr + p * (v[i][j] + a_[j]) * g[j];
        m = mn;
      }
    }
    __syncwarp();                                      // every lane has read the stage
    if (lane == 0 && c + DEPTH < nchunks) issue(c + DEPTH);
  }
  const float inv = 1.0f / (z + 1e-16f);
  float o[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) o[i] = acc[i] * inv;
  stv<VEC>(attn + e * D + ch, o);
  if (fuse_skip) {
    float sk[VEC];
    ldv<VEC>(qkvs + e * ldq + 3 * D + ch, sk);
#pragma unroll
    for (int i = 0; i < VEC; ++i) o[i] += sk[i];
  }
  stv<VEC>(out + e * D + ch, o);
  const float l = (end > beg) ? m + logf(z) : 0.f;
  if (leader) lse[e * H + head] = l;
}

// ------------------------------------------------------------------ backward pass 1 (by target)
// EA == kEaSegment: d(lin_edge out) is summed over the segment in registers and written as ONE row per
// target, dea[e, :] -- the per-triplet [T, D] stream disappears.  DROP: attention dropout active.
template <int VEC, int EA, bool DROP, int LPH>
__global__ void __launch_bounds__(128, 8)
k_attn_bwd_tgt(const float* __restrict__ qkvs, int ldq, const float* __restrict__ ea,
               const int32_t* __restrict__ ea_index,
               const float* __restrict__ sg, const float* __restrict__ attn,
               const float* __restrict__ lse, const float* __restrict__ gout,
               const int32_t* __restrict__ src, const int32_t* __restrict__ rowptr,
               const int32_t* __restrict__ order, int64_t E, int H, int C,
               float scale, float dropout_p, uint64_t seed, float* __restrict__ dqkv, int ldg,
               float* __restrict__ dea, float* __restrict__ dsg, float* __restrict__ al,
               float* __restrict__ da_out) {
  pdl_sync();
  constexpr int D = 32 * VEC;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;
  const int lane = threadIdx.x & 31;
  const int ch = lane * VEC;
  const int head = ch / C;
  const int lph = LPH > 0 ? LPH : C / VEC;
  const bool leader = (ch % C) == 0;

  float q[VEC], g[VEC], o[VEC], dq[VEC];
  ldv<VEC>(qkvs + e * ldq + ch, q);
  ldv<VEC>(gout + e * D + ch, g);
  ldv<VEC>(attn + e * D + ch, o);
  float r = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    r = fmaf(g[i], o[i], r);
    dq[i] = 0.f;
  }
  r = head_sum_t<LPH>(r, lph);                // r_eh = sum_t alpha dalpha = <G, O>  (App. A)
  const float l = lse[e * H + head];
  const int beg = rowptr[e], end = rowptr[e + 1];
  float a_seg[VEC], dea_acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) a_seg[j] = dea_acc[j] = 0.f;
  if constexpr (EA == kEaSegment) ldv<VEC>(ea + (int64_t)ea_index[e] * D + ch, a_seg);

  struct Row {
    float k[VEC], v[VEC], g[VEC], a[EA == kEaTriplet ? VEC : 1];
    int t;
  };
  for (int base = beg; base < end; base += 32) {
    const int my = base + lane;
    int t_l = 0, s_l = 0;
    if (my < end) {
      t_l = order ? order[my] : my;            // order == NULL: the triplet list is target-sorted
      s_l = src[t_l];
    }
    const int cnt = min(32, end - base);
    auto fetch = [&](Row& rw, int i) {         // i is clamped: re-reading the last triplet is harmless
      i = min(i, cnt - 1);
      rw.t = __shfl_sync(0xffffffffu, t_l, i);
      const int s = __shfl_sync(0xffffffffu, s_l, i);
      ldv<VEC>(qkvs + (int64_t)s * ldq + D + ch, rw.k);
      ldv<VEC>(qkvs + (int64_t)s * ldq + 2 * D + ch, rw.v);
      ldv<VEC>(sg + (int64_t)rw.t * D + ch, rw.g);
      if constexpr (EA == kEaTriplet) ldv<VEC>(ea + (int64_t)rw.t * D + ch, rw.a);
    };
    auto consume = [&](const Row& rw) {
      float k[VEC], v[VEC];
      float dot = 0.f, dal = 0.f;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float a_ = EA == kEaTriplet ? rw.a[EA == kEaTriplet ? j : 0] : a_seg[j];
        k[j] = rw.k[j] + a_;                   // kk
        v[j] = rw.v[j] + a_;                   // vv
        dot = fmaf(q[j], k[j], dot);
        dal = fmaf(g[j] * v[j], rw.g[j], dal);
      }
      // both reductions share the shuffle steps
      if constexpr (LPH > 0) {
#pragma unroll
        for (int off = 1; off < LPH; off <<= 1) {
          dot += __shfl_xor_sync(0xffffffffu, dot, off);
          dal += __shfl_xor_sync(0xffffffffu, dal, off);
        }
      } else {
        for (int off = 1; off < lph; off <<= 1) {
          dot += __shfl_xor_sync(0xffffffffu, dot, off);
          dal += __shfl_xor_sync(0xffffffffu, dal, off);
        }
      }
      const float alpha = expf(dot * scale - l);
      float keep = 1.f;
      if constexpr (DROP) keep = keep_scale(seed, rw.t, head, H, dropout_p);
      const float alpha_d = alpha * keep;      // weight actually applied to the value
      const float da = alpha * (dal * keep - r);
      const float sda = scale * da;
      float o_ea[VEC], o_sg[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        dq[j] = fmaf(sda, k[j], dq[j]);
        const float dkk = sda * q[j];
        const float dvv = g[j] * rw.g[j] * alpha_d;
        o_ea[j] = dkk + dvv;
        o_sg[j] = g[j] * v[j] * alpha_d;
      }
      if constexpr (EA == kEaSegment) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) dea_acc[j] += o_ea[j];
      } else if constexpr (EA == kEaTriplet) {
        stv<VEC>(dea + (int64_t)rw.t * D + ch, o_ea);
      }
      stv<VEC>(dsg + (int64_t)rw.t * D + ch, o_sg);
      if (leader) {
        if (al) al[(int64_t)rw.t * H + head] = alpha_d;      // (NULL: the by-source pass reads dEA instead)
        da_out[(int64_t)rw.t * H + head] = da;
      }
    };
    Row ring[kRingBwd];
#pragma unroll
    for (int u = 0; u < kRingBwd - 1; ++u) fetch(ring[u], u);
    for (int i = 0; i < cnt; i += kRingBwd) {
#pragma unroll
      for (int u = 0; u < kRingBwd; ++u) {
        if (i + u < cnt) {
          fetch(ring[(u + kRingBwd - 1) % kRingBwd], i + u + kRingBwd - 1);
          consume(ring[u]);
        }
      }
    }
  }
  stv<VEC>(dqkv + e * ldg + ch, dq);
  if constexpr (EA == kEaSegment) stv<VEC>(dea + e * D + ch, dea_acc);
}

// out[m, :] = sum of in[order[i], :] over i in [rowptr[m], rowptr[m+1]) in that (ascending) order: the
// per-target d(lin_edge out) rows summed per edge_attr table row.  Warp per output row, deterministic.
template <int VEC>
__global__ void __launch_bounds__(128)
k_rows_segsum(const float* __restrict__ in, const int32_t* __restrict__ rowptr,
              const int32_t* __restrict__ order, int64_t M, float* __restrict__ out) {
  pdl_sync();
  constexpr int D = 32 * VEC;
  const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (m >= M) return;
  const int ch = (threadIdx.x & 31) * VEC;
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  const int beg = rowptr[m], end = rowptr[m + 1];
#pragma unroll 4
  for (int i = beg; i < end; ++i) {
    float v[VEC];
    ldv<VEC>(in + (int64_t)order[i] * D + ch, v);
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] += v[j];
  }
  stv<VEC>(out + m * D + ch, acc);
}

// ------------------------------------------------------------------ backward pass 2 (by source)
template <int VEC>
__global__ void __launch_bounds__(128)
k_attn_bwd_src(const float* __restrict__ qkvs, int ldq, const float* __restrict__ sg,
               const float* __restrict__ gout, const float* __restrict__ al,
               const float* __restrict__ da_in, const int32_t* __restrict__ tgt,
               const int32_t* __restrict__ rowptr_src, const int32_t* __restrict__ order_src,
               int64_t E, int H, int C, float scale, float* __restrict__ dqkv, int ldg) {
  pdl_sync();
  constexpr int D = 32 * VEC;
  const int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (f >= E) return;
  const int lane = threadIdx.x & 31;
  const int ch = lane * VEC;
  const int head = ch / C;
  float dk[VEC], dv[VEC];
#pragma unroll
  for (int i = 0; i < VEC; ++i) dk[i] = dv[i] = 0.f;
  const int beg = rowptr_src[f], end = rowptr_src[f + 1];
  for (int base = beg; base < end; base += 32) {
    const int my = base + lane;
    int t_l = 0, e_l = 0;
    if (my < end) {
      t_l = order_src[my];
      e_l = tgt[t_l];
    }
    const int cnt = min(32, end - base);
#pragma unroll 4
    for (int i = 0; i < cnt; ++i) {
      const int t = __shfl_sync(0xffffffffu, t_l, i);
      const int e = __shfl_sync(0xffffffffu, e_l, i);
      float q[VEC], g[VEC], sgv[VEC];
      ldv<VEC>(qkvs + (int64_t)e * ldq + ch, q);
      ldv<VEC>(gout + (int64_t)e * D + ch, g);
      ldv<VEC>(sg + (int64_t)t * D + ch, sgv);
      const float a = al[(int64_t)t * H + head];
      const float sda = scale * da_in[(int64_t)t * H + head];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        dv[j] = fmaf(g[j] * sgv[j], a, dv[j]);   // dV[s] += G . Sg . alpha
        dk[j] = fmaf(sda, q[j], dk[j]);          // dK[s] += sigma da Q[e]
      }
    }
  }
  stv<VEC>(dqkv + f * ldg + D + ch, dk);
  stv<VEC>(dqkv + f * ldg + 2 * D + ch, dv);
}

// The same pass when the by-target pass has written dEA_t = dkk_t + dvv_t per triplet (edge_attr [T, A]):
// dK[s] = sum_t sigma da_t Q[e] as above, and dV[s] = sum_t dvv_t = sum_t (dEA_t - dkk_t) -- one streamed row (dEA) and
// one gathered row (Q) per triplet instead of one streamed (Sg) and two gathered (Q, G) plus alpha: the pass is bound
// by L1 / L2 requests (1.66 kB per triplet, ~10 TB/s with the gathers), not by DRAM.  The subtraction costs at most
// eps (|dEA_t| + |dkk_t|) per term.
template <int VEC>
__global__ void __launch_bounds__(128)
k_attn_bwd_src_dea(const float* __restrict__ qkvs, int ldq, const float* __restrict__ dea,
                   const float* __restrict__ da_in, const int32_t* __restrict__ tgt,
                   const int32_t* __restrict__ rowptr_src, const int32_t* __restrict__ order_src,
                   int64_t E, int H, int C, float scale, float* __restrict__ dqkv, int ldg) {
  pdl_sync();
  constexpr int D = 32 * VEC;
  const int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (f >= E) return;
  const int lane = threadIdx.x & 31;
  const int ch = lane * VEC;
  const int head = ch / C;
  float dk[VEC], ds[VEC];           // ds = sum_t dEA_t
#pragma unroll
  for (int i = 0; i < VEC; ++i) dk[i] = ds[i] = 0.f;
  const int beg = rowptr_src[f], end = rowptr_src[f + 1];
  for (int base = beg; base < end; base += 32) {
    const int my = base + lane;
    int t_l = 0, e_l = 0;
    if (my < end) {
      t_l = order_src[my];
      e_l = tgt[t_l];
    }
    const int cnt = min(32, end - base);
#pragma unroll 8
    for (int i = 0; i < cnt; ++i) {
      const int t = __shfl_sync(0xffffffffu, t_l, i);
      const int e = __shfl_sync(0xffffffffu, e_l, i);
      float q[VEC], dv_[VEC];
      ldv<VEC>(qkvs + (int64_t)e * ldq + ch, q);
      ldv<VEC>(dea + (int64_t)t * D + ch, dv_);
      const float sda = scale * da_in[(int64_t)t * H + head];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float kk = sda * q[j];             // dkk_t, formed exactly as the by-target pass formed it
        dk[j] += kk;
        ds[j] += dv_[j] - kk;                    // dvv_t
      }
    }
  }
  stv<VEC>(dqkv + f * ldg + D + ch, dk);
  stv<VEC>(dqkv + f * ldg + 2 * D + ch, ds);
}

}  // namespace x2
#include "blk_attn.cuh"      // block-centric kernels (one CTA per closed block of the line graph)
namespace x2 {

// ------------------------------------------------------------------ Linear dispatch (SIMT / tensor core)
static inline int lin_mode(int mode) {
  return (mode == X2_MODE_TF32X3_FUSED || mode == X2_MODE_TF32) ? X2_MODE_TF32X3 : mode;
}

struct Lin {
  int mode;          // X2_MODE_FP32: SIMT fp32;  X2_MODE_TF32X3: tcgen05 (3xTF32, or one pass if `single`)
  void* img;         // weight-image scratch (tensor-core mode)
  float* wg;         // wgrad partial-tile scratch
  cudaStream_t st;
  int single;        // X2_MODE_TF32: one tf32 pass per product
};

constexpr int kTcBlock = 128;   // the tcgen05 kernels take N <= 128 and a K block whose image fits smem

// C[M,N] (+)= A[M,K] . B(K,N) + bias with B(k,n) = W[k*sbk + n*sbn], blocked over N and K
static int tc_linear(const Lin& L, const float* A, int64_t lda, int64_t M, int K, const float* W, int64_t sbk,
                     int64_t sbn, int N, const float* bias, float* C, int64_t ldc, int beta) {
  for (int n0 = 0; n0 < N; n0 += kTcBlock) {
    const int nb = N - n0 < kTcBlock ? N - n0 : kTcBlock;
    for (int k0 = 0; k0 < K; k0 += kTcBlock) {
      const int kb = K - k0 < kTcBlock ? K - k0 : kTcBlock;
      int rc = tc::tc_gemm(A + k0, lda, M, kb, W + (int64_t)k0 * sbk + (int64_t)n0 * sbn, sbk, sbn, nb,
                           (bias && k0 == 0) ? bias + n0 : nullptr, C + n0, ldc, (beta || k0 > 0) ? 1 : 0,
                           L.img, L.st, L.single);
      if (rc != X2_OK) return rc;
    }
  }
  return X2_OK;
}

// y[M,N] = x[M,K] W[N,K]^T + bias
static int lin_fwd(const Lin& L, const float* x, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                   float* y, int64_t ldy, int64_t M, int N, int K) {
  if (L.mode == X2_MODE_TF32X3) return tc_linear(L, x, ldx, M, K, W, 1, ldw, N, bias, y, ldy, 0);
  return gemm_nt(x, ldx, W, ldw, bias, y, ldy, M, N, K, L.st);
}
// dx[M,N] (+)= dy[M,K] W[K,N]
static int lin_dgrad(const Lin& L, const float* dy, int64_t lddy, const float* W, int64_t ldw, float* dx,
                     int64_t lddx, int64_t M, int N, int K, int beta) {
  if (L.mode == X2_MODE_TF32X3) return tc_linear(L, dy, lddy, M, K, W, ldw, 1, N, nullptr, dx, lddx, beta);
  return gemm_nn(dy, lddy, W, ldw, dx, lddx, M, N, K, beta, L.st);
}
// dW[Md,N] = dy[rows,Md]^T x[rows,N] ; db[Md] = colsum(dy)
static int lin_wgrad(const Lin& L, const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dW,
                     int64_t lddw, float* db, int64_t rows, int Md, int N) {
  if (L.mode == X2_MODE_TF32X3 && Md % kTcBlock == 0) {
    for (int m0 = 0; m0 < Md; m0 += kTcBlock)
      for (int n0 = 0; n0 < N; n0 += kTcBlock) {
        const int nb = N - n0 < kTcBlock ? N - n0 : kTcBlock;
        int rc = tc::tc_wgrad(dy + m0, lddy, x + n0, ldx, rows, nb, dW + (int64_t)m0 * lddw + n0, lddw,
                              (db && n0 == 0) ? db + m0 : nullptr, L.wg, L.st, L.single);
        if (rc != X2_OK) return rc;
      }
    return X2_OK;
  }
  return gemm_wgrad(dy, lddy, x, ldx, dW, lddw, db, rows, Md, N, L.wg, L.st);
}

// ------------------------------------------------------------------ host side
static int check_desc(const x2_conv_desc* d) {
  X2_CHECK_ARG(d != nullptr, "conv: null descriptor");
  X2_CHECK_ARG(d->E >= 0 && d->T >= 0 && d->E < 2147483647LL && d->T < 2147483647LL, "conv: bad E/T");
  X2_CHECK_ARG(d->mode == X2_MODE_FP32 || d->mode == X2_MODE_TF32X3 || d->mode == X2_MODE_TF32X3_FUSED ||
                   d->mode == X2_MODE_TF32, "conv: unknown mode %d", d->mode);
  X2_CHECK_ARG(d->mode == X2_MODE_FP32 || d->D % 128 == 0,
               "conv: X2_MODE_TF32X3 needs heads*out_channels to be a multiple of 128 (got %d)", d->D);
  X2_CHECK_ARG(d->D == d->H * d->C && d->H >= 1 && d->C >= 1, "conv: D=%d != H*C=%d*%d", d->D, d->H, d->C);
  X2_CHECK_ARG(d->D == 32 || d->D == 64 || d->D == 128 || d->D == 256,
               "conv: heads*out_channels must be 32, 64, 128 or 256 (got %d)", d->D);
  const int vec = d->D / 32;
  X2_CHECK_ARG(d->C % vec == 0 && ((d->C / vec) & (d->C / vec - 1)) == 0 && d->C / vec <= 32,
               "conv: out_channels=%d unsupported for D=%d (need C %% (D/32) == 0 and C/(D/32) a power of two)",
               d->C, d->D);
  X2_CHECK_ARG(d->S >= 1 && d->R >= 1 && d->R <= 64 && d->A >= 0, "conv: bad S/R/A (need 1 <= R <= 64)");
  X2_CHECK_ARG(d->dropout_p >= 0.f && d->dropout_p < 1.f, "conv: dropout must be in [0,1)");
  X2_CHECK_ARG((d->A > 0) == (d->w_edge != nullptr), "conv: w_edge must be given iff A > 0");
  if (d->ea_index) {
    X2_CHECK_ARG(d->A > 0, "conv: ea_index (segment-constant edge_attr) needs edge_dim > 0");
    X2_CHECK_ARG(d->ea_rows >= 1 && d->ea_rows < 2147483647LL, "conv: bad ea_rows");
    X2_CHECK_ARG(d->ea_rowptr && d->ea_order, "conv: ea_index needs the ea_rowptr / ea_order grouping");
  }
  return X2_OK;
}

struct FwdWs { float* xs; void* img; float* part; float* wsT; };
static size_t fwd_layout(const x2_conv_desc* d, void* ws, FwdWs* w) {
  Arena a(ws, (size_t)-1);
  w->xs = a.take<float>((size_t)d->E * d->D + 4);
  w->img = a.take<char>(tc::bimage_bytes(kTcBlock, kTcBlock) + 256);
  w->part = a.take<float>(d->items ? (size_t)d->items_bound * tc::kTaPart : 4);   // partial states of the fused forward
  w->wsT = a.take<float>((size_t)d->S * d->D + 4);                                 // lin_sbf.weight^T (factorised sbf)
  return align_up(a.off, 256) + 256;
}

// ---- block-centric kernels (csrc/blk_attn.cuh): when they apply
static bool env_on(const char* name) {       // default on; NAME=0 turns the path off (A/B runs)
  const char* v = getenv(name);
  return !(v && v[0] == '0');
}
constexpr int kDwsParts = 6 * kNumSM;
constexpr size_t kBlkSmemMax = 220 * 1024;    // shared memory of a block-centric CTA
// rows per source of the staged table: the config's 7 orders exactly, anything else padded to 8 with zero rows
static int blk_lt(const x2_conv_desc* d) { return d->sbf_L == 7 ? 7 : kBlkLMax; }
static size_t blk_table_bytes(const x2_conv_desc* d) { return (size_t)d->blk_max_src * blk_lt(d) * 128 * sizeof(float); }
// factorised lin_sbf: closed blocks of a target-sorted line graph + the factors of sbf, L <= 8 orders of <= 8 radial
// functions, D = 128, no dropout, and W_s^T + one block's table + the target offsets fit shared memory
static bool sgf_usable(const x2_conv_desc* d) {
  static const bool on = env_on("X2GNN_SGF") && env_on("X2GNN_BLOCK");
  if (!(on && d->nblk > 0 && d->nblk < 2147483647LL && d->blk_sptr && d->blk_tptr && d->blk_tord && d->blk_tpos &&
        d->tgt_sorted && d->T > 0 && d->D == 128 && d->dropout_p == 0.f))
    return false;
  if (!(d->sbf_tab && d->angles && d->sbf_L >= 1 && d->sbf_L <= kBlkLMax && d->sbf_R >= 1 && d->sbf_R <= kBlkRMax &&
        d->sbf_L * d->sbf_R == d->S && d->blk_max_src > 0 && d->blk_max_tgt > 0 && d->blk_max_trip > 0))
    return false;
  const size_t need = (size_t)d->S * 128 * sizeof(float) + blk_table_bytes(d) + ((size_t)d->blk_max_tgt + 8) * 4;
  return need <= kBlkSmemMax && (reinterpret_cast<uintptr_t>(d->b_sbf) & 15) == 0;
}

struct BwdWs {
  float *dea, *dea_tab, *dsg, *al, *da, *dqkv, *xs, *F, *dxs, *dxs2, *dx2, *wg, *wsT, *dP;
  void* img;
  size_t wg_floats;
};
static size_t bwd_layout(const x2_conv_desc* d, void* ws, BwdWs* w) {
  Arena a(ws, (size_t)-1);
  const size_t ED = (size_t)d->E * d->D, TD = (size_t)d->T * d->D, TH = (size_t)d->T * d->H;
  // segment-constant edge_attr: one d(lin_edge out) row per target + one per table row
  w->dea = d->A > 0 ? a.take<float>((d->ea_index ? ED : TD) + 4) : nullptr;
  w->dea_tab = d->ea_index ? a.take<float>((size_t)d->ea_rows * d->D + 4) : nullptr;
  const bool sgf = sgf_usable(d);
  w->dsg = a.take<float>(sgf ? 4 : TD + 4);
  w->wsT = a.take<float>((size_t)d->S * d->D + 4);
  w->dP = a.take<float>(sgf ? (size_t)d->E * d->sbf_L * d->D + 4 : 4);      // d(per-source basis table)
  w->al = a.take<float>(TH + 4);
  w->da = a.take<float>(TH + 4);
  w->dqkv = a.take<float>(3 * ED + 4);
  w->xs = a.take<float>(ED + 4);
  w->F = a.take<float>(ED + 4);
  w->dxs = a.take<float>(ED + 4);
  w->dxs2 = a.take<float>(ED + 4);
  w->dx2 = a.take<float>(ED + 4);
  size_t wg = wgrad_workspace_floats(d->T, d->D, d->A > 0 ? d->A : 1);
  size_t t2 = wgrad_workspace_floats(d->T, d->D, d->S);
  if (t2 > wg) wg = t2;
  t2 = wgrad_workspace_floats(d->E, d->D, d->D);
  if (t2 > wg) wg = t2;
  t2 = wgrad_workspace_floats(d->E, d->D, d->R);
  if (t2 > wg) wg = t2;
  t2 = tc::tc_wgrad_workspace_floats(d->T > d->E ? d->T : d->E, kTcBlock);
  if (t2 > wg) wg = t2;
  t2 = (size_t)kDwsParts * ((size_t)d->D * d->S + d->D) + 64;
  if (t2 > wg) wg = t2;
  w->wg_floats = wg;
  w->wg = a.take<float>(wg);
  w->img = a.take<char>(tc::bimage_bytes(kTcBlock, kTcBlock) + 256);
  return align_up(a.off, 256) + 256;
}

static bool staged_fwd_enabled() {
  static const int on = [] { const char* v = getenv("X2GNN_STAGED"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}

template <int VEC, int EA, bool GENERAL>
static void launch_attn_fwd_inst(const x2_conv_desc* d, const x2_conv_saved* s, float* out, float* alpha,
                                 cudaStream_t st) {
  const float scale = 1.0f / sqrtf((float)d->C);
  const int32_t* order = d->tgt_sorted ? nullptr : d->order_tgt;     // sorted: order_tgt is the identity
  const unsigned grid = (unsigned)cdiv(d->E * 32, 128);
  // measured (QM9 batch 128): the staged kernel wins when only Sg is streamed (0.134 -> 0.126 ms) and loses
  // slightly when EA is streamed too (two copies per stage, 6 instead of 8 blocks per SM)
  if constexpr (!GENERAL && VEC == 4 && EA != kEaTriplet) {
    if (d->tgt_sorted && d->C == 2 * VEC && staged_fwd_enabled() &&
        ((reinterpret_cast<uintptr_t>(s->sg) | reinterpret_cast<uintptr_t>(s->ea)) & 15) == 0) {
      constexpr int DEPTH = EA == kEaTriplet ? 2 : 3;
      constexpr int NARR = EA == kEaTriplet ? 2 : 1;
      const size_t smem = (size_t)4 * DEPTH * NARR * kStRows * 32 * VEC * sizeof(float) + 4 * DEPTH * sizeof(uint64_t);
      launch_k(k_attn_fwd_stage<VEC, EA, 2, DEPTH>, dim3(grid), dim3(128), smem, st, 
          s->qkvs, 4 * d->D, s->ea, d->ea_index, s->sg, d->src, d->rowptr_tgt, d->E, d->H, d->C, scale,
          d->fuse_skip, s->attn, out, s->lse);
      return;
    }
  }
  if (d->C == 2 * VEC)     // config.json: C = 8, 4 channels per lane => 2 lanes per head
    launch_k(k_attn_fwd<VEC, EA, GENERAL, 2>, dim3(grid), dim3(128), 0, st, 
        s->qkvs, 4 * d->D, s->ea, d->ea_index, s->sg, d->src, d->rowptr_tgt, order, d->E,
        d->H, d->C, scale, d->fuse_skip, d->dropout_p, d->seed, s->attn, out, s->lse, alpha);
  else
    launch_k(k_attn_fwd<VEC, EA, GENERAL, 0>, dim3(grid), dim3(128), 0, st, 
        s->qkvs, 4 * d->D, s->ea, d->ea_index, s->sg, d->src, d->rowptr_tgt, order, d->E,
        d->H, d->C, scale, d->fuse_skip, d->dropout_p, d->seed, s->attn, out, s->lse, alpha);
}
template <int VEC>
static int launch_attn_fwd(const x2_conv_desc* d, const x2_conv_saved* s, float* out, float* alpha,
                           cudaStream_t st) {
  const bool general = alpha != nullptr || d->dropout_p > 0.f;
  const int ea = d->A == 0 ? kEaNone : (d->ea_index ? kEaSegment : kEaTriplet);
  switch (ea * 2 + (general ? 1 : 0)) {
    case 0: launch_attn_fwd_inst<VEC, kEaNone, false>(d, s, out, alpha, st); break;
    case 1: launch_attn_fwd_inst<VEC, kEaNone, true>(d, s, out, alpha, st); break;
    case 2: launch_attn_fwd_inst<VEC, kEaTriplet, false>(d, s, out, alpha, st); break;
    case 3: launch_attn_fwd_inst<VEC, kEaTriplet, true>(d, s, out, alpha, st); break;
    case 4: launch_attn_fwd_inst<VEC, kEaSegment, false>(d, s, out, alpha, st); break;
    default: launch_attn_fwd_inst<VEC, kEaSegment, true>(d, s, out, alpha, st); break;
  }
  X2_LAUNCH_OK();
  return X2_OK;
}

static unsigned long long* g_tile_trace = nullptr;     // development only (x2_debug_tile_trace)
static bool tile_fwd_enabled() {      // X2GNN_FUSED=1: take the fused tile forward in X2_MODE_TF32X3 as well
  static const int on = [] { const char* v = getenv("X2GNN_FUSED"); return (v && v[0] == '1') ? 1 : 0; }();
  return on != 0;
}
// the fused forward needs: tensor-core mode, the tiling of a target-sorted list, D = 128, edge_attr [T,128]
// (or the table form / no lin_edge), S even <= 64, no dropout, no alpha output, 16-byte aligned rows
static bool tile_fwd_usable(const x2_conv_desc* d, const float* alpha) {
  if (!(d->mode == X2_MODE_TF32X3_FUSED || (d->mode == X2_MODE_TF32X3 && tile_fwd_enabled()))) return false;
  if (!d->items || !d->itemptr || d->items_bound <= 0 || !d->tgt_sorted || d->T <= 0 || alpha || d->dropout_p > 0.f) return false;
  if (!tc::tile_fwd_supported(d->D, d->H, d->C, d->A, d->S, d->ea_index != nullptr)) return false;
  if ((reinterpret_cast<uintptr_t>(d->sbf) & 7) != 0) return false;
  if (d->A > 0 && !d->ea_index && (reinterpret_cast<uintptr_t>(d->edge_attr) & 15) != 0) return false;
  return true;
}
template <int LPH>
static int launch_tile_fwd_lph(const tc::TaParams& p, int ea_mode, int grid, cudaStream_t st) {
  auto go = [&](auto kern) -> int {
    X2_DYN_SMEM(kern, tc::kTaSmem);
    launch_k(kern, dim3(grid), dim3(tc::kTaThreads), tc::kTaSmem, st, p);
    X2_LAUNCH_OK();
    return X2_OK;
  };
  switch (ea_mode) {
    case tc::kTaEaTriplet: return go(tc::k_tile_fwd<LPH, tc::kTaEaTriplet>);
    case tc::kTaEaSegment: return go(tc::k_tile_fwd<LPH, tc::kTaEaSegment>);
    default: return go(tc::k_tile_fwd<LPH, tc::kTaEaNone>);
  }
}
static int launch_tile_fwd(const x2_conv_desc* d, const x2_conv_saved* s, float* out, float* part, cudaStream_t st) {
  tc::TaParams p{};
  const int ea_mode = d->A == 0 ? tc::kTaEaNone : (d->ea_index ? tc::kTaEaSegment : tc::kTaEaTriplet);
  p.ea = ea_mode == tc::kTaEaSegment ? s->ea : d->edge_attr;
  p.ea_index = d->ea_index;
  p.sbf = d->sbf; p.S = d->S;
  p.w_edge = d->w_edge; p.w_sbf = d->w_sbf; p.b_sbf = d->b_sbf;
  p.qkvs = s->qkvs; p.ldq = 4 * d->D;
  p.src = d->src;
  p.items = d->items; p.itemptr = d->itemptr;
  p.E = d->E; p.T = d->T;
  p.H = d->H; p.C = d->C; p.scale = 1.0f / sqrtf((float)d->C);
  p.part = part;
  p.ea_out = ea_mode == tc::kTaEaTriplet ? s->ea : nullptr;      // still consumed by the backward kernels
  p.sg_out = s->sg;
  p.trace = g_tile_trace;
  { static const int dbg = [] { const char* v = getenv("X2GNN_TA_DBG"); return v ? atoi(v) : 0; }(); p.dbg = dbg; }
  // persistent: one CTA per SM (fewer when the list is short: a unit is ~100 rows)
  const int64_t units = d->T / (X2_UNIT_ITEMS * X2_ITEM_ROWS) + 1;
  const int grid = (int)(units < kNumSM ? units : kNumSM);
  int rc;
  switch (d->C / 4) {
    case 1: rc = launch_tile_fwd_lph<1>(p, ea_mode, grid, st); break;
    case 2: rc = launch_tile_fwd_lph<2>(p, ea_mode, grid, st); break;
    case 4: rc = launch_tile_fwd_lph<4>(p, ea_mode, grid, st); break;
    default: rc = launch_tile_fwd_lph<0>(p, ea_mode, grid, st); break;
  }
  if (rc != X2_OK) return rc;
  launch_k(tc::k_item_merge, dim3((unsigned)cdiv(d->E * 32, 128)), dim3(128), 0, st, (const float*)part, d->itemptr,
           (const float*)s->qkvs, 4 * d->D, d->E, d->H, d->C, d->fuse_skip, out, s->attn, s->lse);
  X2_LAUNCH_OK();
  return X2_OK;
}

// by-source pass from dEA (k_attn_bwd_src_dea) whenever edge_attr is [T, A]; X2GNN_SRC_DEA=0: the generic pass (A/B runs)
static bool src_from_dea() {
  static const bool on = env_on("X2GNN_SRC_DEA");
  return on;
}

template <int VEC, int EA, bool DROP>
static void launch_attn_bwd_tgt_inst(const x2_conv_desc* d, const x2_conv_saved* s, const float* gout,
                                     const BwdWs& w, cudaStream_t st) {
  const float scale = 1.0f / sqrtf((float)d->C);
  const unsigned grid = (unsigned)cdiv(d->E * 32, 128);
  const int32_t* order = d->tgt_sorted ? nullptr : d->order_tgt;     // sorted: order_tgt is the identity
  float* al = (EA == kEaTriplet && src_from_dea()) ? nullptr : w.al;  // alpha is only kept for the generic by-source pass
  if (d->C == 2 * VEC)
    launch_k(k_attn_bwd_tgt<VEC, EA, DROP, 2>, dim3(grid), dim3(128), 0, st, 
        s->qkvs, 4 * d->D, s->ea, d->ea_index, s->sg, s->attn, s->lse, gout, d->src, d->rowptr_tgt, order, d->E,
        d->H, d->C, scale, d->dropout_p, d->seed, w.dqkv, 3 * d->D, w.dea, w.dsg, al, w.da);
  else
    launch_k(k_attn_bwd_tgt<VEC, EA, DROP, 0>, dim3(grid), dim3(128), 0, st, 
        s->qkvs, 4 * d->D, s->ea, d->ea_index, s->sg, s->attn, s->lse, gout, d->src, d->rowptr_tgt, order, d->E,
        d->H, d->C, scale, d->dropout_p, d->seed, w.dqkv, 3 * d->D, w.dea, w.dsg, al, w.da);
}

template <int VEC>
static int launch_attn_bwd(const x2_conv_desc* d, const x2_conv_saved* s, const float* gout,
                           const BwdWs& w, cudaStream_t st) {
  const float scale = 1.0f / sqrtf((float)d->C);
  const unsigned grid = (unsigned)cdiv(d->E * 32, 128);
  phase_begin(st);
  const bool drop = d->dropout_p > 0.f;
  const int ea = d->A == 0 ? kEaNone : (d->ea_index ? kEaSegment : kEaTriplet);
  switch (ea * 2 + (drop ? 1 : 0)) {
    case 0: launch_attn_bwd_tgt_inst<VEC, kEaNone, false>(d, s, gout, w, st); break;
    case 1: launch_attn_bwd_tgt_inst<VEC, kEaNone, true>(d, s, gout, w, st); break;
    case 2: launch_attn_bwd_tgt_inst<VEC, kEaTriplet, false>(d, s, gout, w, st); break;
    case 3: launch_attn_bwd_tgt_inst<VEC, kEaTriplet, true>(d, s, gout, w, st); break;
    case 4: launch_attn_bwd_tgt_inst<VEC, kEaSegment, false>(d, s, gout, w, st); break;
    default: launch_attn_bwd_tgt_inst<VEC, kEaSegment, true>(d, s, gout, w, st); break;
  }
  X2_LAUNCH_OK();
  if (d->ea_index) {       // per-target rows -> per-table-row sums (fixed order)
    launch_k(k_rows_segsum<VEC>, dim3((unsigned)cdiv(d->ea_rows * 32, 128)), dim3(128), 0, st, w.dea, d->ea_rowptr, d->ea_order,
                                                                             d->ea_rows, w.dea_tab);
    X2_LAUNCH_OK();
  }
  phase_end(X2_PHASE_ATTN_BWD_TGT, st);
  if (ea == kEaTriplet && src_from_dea())
    launch_k(k_attn_bwd_src_dea<VEC>, dim3(grid), dim3(128), 0, st, s->qkvs, 4 * d->D, w.dea, w.da, d->tgt,
             d->rowptr_src, d->order_src, d->E, d->H, d->C, scale, w.dqkv, 3 * d->D);
  else
    launch_k(k_attn_bwd_src<VEC>, dim3(grid), dim3(128), 0, st, s->qkvs, 4 * d->D, s->sg, gout, w.al, w.da, d->tgt,
                                            d->rowptr_src, d->order_src, d->E, d->H, d->C, scale, w.dqkv,
                                            3 * d->D);
  X2_LAUNCH_OK();
  phase_end(X2_PHASE_ATTN_BWD_SRC, st);
  return X2_OK;
}

// ---- block-centric launches (factorised sbf)
static void blk_fill(const x2_conv_desc* d, const x2_conv_saved* s, BlkParams& p, const float* wsT) {
  p.qkvs = s->qkvs; p.ldq = 4 * d->D;
  p.ea = s->ea; p.ea_index = d->ea_index;
  p.stab = d->sbf_tab; p.angles = d->angles; p.wsT = wsT; p.b_sbf = d->b_sbf;
  p.L = d->sbf_L; p.Rr = d->sbf_R; p.S = d->S;
  p.src = d->src; p.tgt = d->tgt; p.rowptr_tgt = d->rowptr_tgt; p.rowptr_src = d->rowptr_src; p.order_src = d->order_src;
  p.blk_sptr = d->blk_sptr; p.blk_tptr = d->blk_tptr; p.blk_tord = d->blk_tord; p.blk_tpos = d->blk_tpos;
  p.nblk = (int)d->nblk; p.maxS = d->blk_max_src; p.maxTrip = d->blk_max_trip; p.maxTgt = d->blk_max_tgt;
  p.E = d->E; p.H = d->H; p.C = d->C; p.scale = 1.0f / sqrtf((float)d->C);
  p.fuse_skip = d->fuse_skip;
}
template <int LPH, int LT>
static int launch_blk_fwd_t(const x2_conv_desc* d, const BlkParams& p, cudaStream_t st) {
  const int ea = d->A == 0 ? kEaNone : (d->ea_index ? kEaSegment : kEaTriplet);
  // two halves (two blocks in flight per SM) when two staged tables fit next to W_s^T
  const size_t wbytes = (size_t)d->S * 128 * sizeof(float);
  const int nh = wbytes + 2 * blk_table_bytes(d) <= kBlkSmemMax ? 2 : 1;
  const size_t smem = wbytes + nh * blk_table_bytes(d);
  const int64_t want = cdiv(d->nblk, nh);
  const unsigned grid = (unsigned)(want < kNumSM ? want : kNumSM);
  auto go = [&](auto kern) -> int {
    X2_DYN_SMEM(kern, kBlkSmemMax);      // once per (kernel, device): the largest layout the path takes
    launch_k(kern, dim3(grid), dim3(nh * kBlkHalf), smem, st, p);
    X2_LAUNCH_OK();
    return X2_OK;
  };
  switch (ea) {
    case kEaNone: return go(k_blk_fwd<kEaNone, LPH, LT>);
    case kEaTriplet: return go(k_blk_fwd<kEaTriplet, LPH, LT>);
    default: return go(k_blk_fwd<kEaSegment, LPH, LT>);
  }
}
template <int LPH, int LT, bool ALDA>
static int launch_blk_bwd_t(const x2_conv_desc* d, const BlkParams& p, size_t smem, cudaStream_t st) {
  const int ea = d->A == 0 ? kEaNone : (d->ea_index ? kEaSegment : kEaTriplet);
  const unsigned grid = (unsigned)(d->nblk < kNumSM ? d->nblk : kNumSM);
  auto go = [&](auto kern) -> int {
    X2_DYN_SMEM(kern, kBlkSmemMax);
    launch_k(kern, dim3(grid), dim3(kBlkHalf), smem, st, p);
    X2_LAUNCH_OK();
    return X2_OK;
  };
  switch (ea) {
    case kEaNone: return go(k_blk_bwd<kEaNone, LPH, LT, ALDA>);
    case kEaTriplet: return go(k_blk_bwd<kEaTriplet, LPH, LT, ALDA>);
    default: return go(k_blk_bwd<kEaSegment, LPH, LT, ALDA>);
  }
}
static int launch_blk_fwd(const x2_conv_desc* d, const BlkParams& p, cudaStream_t st) {
  const bool l2 = d->C == 8;
  if (blk_lt(d) == 7) return l2 ? launch_blk_fwd_t<2, 7>(d, p, st) : launch_blk_fwd_t<0, 7>(d, p, st);
  return l2 ? launch_blk_fwd_t<2, 8>(d, p, st) : launch_blk_fwd_t<0, 8>(d, p, st);
}
static int launch_blk_bwd(const x2_conv_desc* d, const BlkParams& p, cudaStream_t st) {
  const bool l2 = d->C == 8;
  // alpha / d(logit) of a block's triplets stay in shared memory between the passes when they fit
  const size_t base = (size_t)d->S * 128 * sizeof(float) + blk_table_bytes(d) + (((size_t)d->blk_max_tgt + 1 + 3) & ~(size_t)3) * 4;
  const size_t ad = (size_t)d->blk_max_trip * 2 * d->H * sizeof(float);
  const bool alda = base + ad <= kBlkSmemMax;
  const size_t smem = base + (alda ? ad : 0);
  if (blk_lt(d) == 7) {
    if (alda) return l2 ? launch_blk_bwd_t<2, 7, true>(d, p, smem, st) : launch_blk_bwd_t<0, 7, true>(d, p, smem, st);
    return l2 ? launch_blk_bwd_t<2, 7, false>(d, p, smem, st) : launch_blk_bwd_t<0, 7, false>(d, p, smem, st);
  }
  if (alda) return l2 ? launch_blk_bwd_t<2, 8, true>(d, p, smem, st) : launch_blk_bwd_t<0, 8, true>(d, p, smem, st);
  return l2 ? launch_blk_bwd_t<2, 8, false>(d, p, smem, st) : launch_blk_bwd_t<0, 8, false>(d, p, smem, st);
}

}  // namespace x2

using namespace x2;

#define X2_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc != X2_OK) return _rc; \
  } while (0)

extern "C" {

// development hook (not in x2gnn.h): device buffer of 16 x 256 uint64 that CTA 0 of the fused tile kernel fills
// with clock64() stamps of its roles (tools/tile_probe.py); NULL turns it off
void x2_debug_tile_trace(void* dev_buf) { x2::g_tile_trace = static_cast<unsigned long long*>(dev_buf); }

// ---- tensor-core GEMM building blocks (also exercised directly by tests/test_gpu_tc_gemm.py)
size_t x2_tc_gemm_workspace_bytes(int32_t K, int32_t N) { return tc::bimage_bytes(K, N) + 256; }

int x2_tc_gemm(const float* A, int64_t lda, int64_t M, int32_t K, const float* W, int64_t sbk, int64_t sbn,
               int32_t N, const float* bias, float* C, int64_t ldc, int32_t beta, void* ws, size_t ws_bytes,
               void* stream) {
  X2_CHECK_ARG(A && W && C && M >= 0, "x2_tc_gemm: null pointer");
  X2_CHECK_ARG(tc::g1_supported(K, N), "x2_tc_gemm: unsupported K=%d N=%d (need N <= 128 and the weight image to fit smem)", K, N);
  if (ws_bytes < x2_tc_gemm_workspace_bytes(K, N)) { set_error("x2_tc_gemm: workspace too small"); return X2_EWORKSPACE; }
  return tc::tc_gemm(A, lda, M, K, W, sbk, sbn, N, bias, C, ldc, beta, ws, (cudaStream_t)stream);
}

size_t x2_tc_wgrad_workspace_bytes(int64_t rows, int32_t N) {
  return tc::tc_wgrad_workspace_floats(rows, N) * sizeof(float) + 256;
}

int x2_tc_wgrad(const float* Y, int64_t ldy, const float* X, int64_t ldx, int64_t rows, int32_t N, float* dW,
                int64_t lddw, float* db, void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(Y && X && dW && rows >= 0, "x2_tc_wgrad: null pointer");
  X2_CHECK_ARG(N >= 1 && N <= 128, "x2_tc_wgrad: need 1 <= N <= 128 (got %d)", N);
  if (ws_bytes < x2_tc_wgrad_workspace_bytes(rows, N)) { set_error("x2_tc_wgrad: workspace too small"); return X2_EWORKSPACE; }
  return tc::tc_wgrad(Y, ldy, X, ldx, rows, N, dW, lddw, db, static_cast<float*>(ws), (cudaStream_t)stream);
}

// njobs weight gradients over the same `rows` and N, kMaxProbG2 per launch (each launch: one k_tc_wgrad with the CTAs
// dealt over its problems + one fixed-order reduction).  The same workspace serves every launch (stream order).
int x2_tc_wgrad_batch(const x2_wgrad_job* jobs, int32_t njobs, int64_t rows, int32_t N, void* ws, size_t ws_bytes,
                      void* stream) {
  X2_CHECK_ARG(jobs && njobs >= 1 && rows >= 0, "x2_tc_wgrad_batch: bad arguments");
  X2_CHECK_ARG(N >= 1 && N <= 128, "x2_tc_wgrad_batch: need 1 <= N <= 128 (got %d)", N);
  if (ws_bytes < x2_tc_wgrad_workspace_bytes(rows, N)) { set_error("x2_tc_wgrad_batch: workspace too small"); return X2_EWORKSPACE; }
  for (int32_t i = 0; i < njobs; i += tc::kMaxProbG2) {
    tc::G2Job g[tc::kMaxProbG2];
    const int n = njobs - i < tc::kMaxProbG2 ? njobs - i : tc::kMaxProbG2;
    for (int j = 0; j < n; ++j) {
      const x2_wgrad_job& q = jobs[i + j];
      X2_CHECK_ARG(q.Y && q.X && q.dW, "x2_tc_wgrad_batch: null pointer in job %d", i + j);
      g[j] = tc::G2Job{q.Y, q.ldy, q.X, q.ldx, q.dW, q.lddw, q.db};
    }
    X2_TRY(tc::tc_wgrad_batch(g, n, rows, N, static_cast<float*>(ws), (cudaStream_t)stream));
  }
  return X2_OK;
}

int x2_sbfconv_plan(const x2_conv_desc* d) {
  if (!d) return 0;
  return sgf_usable(d) ? (X2_PLAN_BLOCKS | X2_PLAN_FACTORISED_SBF) : 0;
}

size_t x2_sbfconv_fwd_workspace_bytes(const x2_conv_desc* d) {
  if (!d) return 0;
  FwdWs w;
  return fwd_layout(d, nullptr, &w);
}

size_t x2_sbfconv_bwd_workspace_bytes(const x2_conv_desc* d) {
  if (!d) return 0;
  BwdWs w;
  return bwd_layout(d, nullptr, &w);
}

int x2_sbfconv_fwd(const x2_conv_desc* d, const x2_conv_saved* s, float* out, float* alpha,
                   void* ws, size_t ws_bytes, void* stream) {
  X2_TRY(check_desc(d));
  X2_CHECK_ARG(s && s->qkvs && s->attn && s->lse && out, "conv fwd: null output buffer");
  const bool sgf = sgf_usable(d) && !alpha;
  X2_CHECK_ARG(sgf || (s->sg && (d->sbf || d->T == 0)), "conv fwd: sbf / saved.sg required (no usable factorised sbf)");
  X2_CHECK_ARG(d->A == 0 || s->ea, "conv fwd: saved.ea required when A > 0");
  X2_CHECK_ARG(!d->fuse_skip || d->w_skip, "conv fwd: fuse_skip needs w_skip");
  cudaStream_t st = (cudaStream_t)stream;
  FwdWs w;
  const size_t need = fwd_layout(d, ws, &w);
  if (ws_bytes < need) { set_error("conv fwd: workspace %zu < %zu", ws_bytes, need); return X2_EWORKSPACE; }
  const int64_t E = d->E, T = d->T;
  const int D = d->D;
  if (E == 0) return X2_OK;

  phase_begin(st);
  if (s->xs) w.xs = s->xs;          // kept for the backward (saved.xs) instead of living in the workspace
  // (1) x_src = x * lin_rbf(rbf)                                             :99-100
  if ((size_t)D * d->R * sizeof(float) > 48 * 1024) X2_DYN_SMEM(k_rbf_filter, 64 * 1024);     // D = 256, R > 48
  launch_k(k_rbf_filter, dim3(filter_grid(E)), dim3(kFilterRows * 32), (size_t)D * d->R * sizeof(float), st, 
      d->x, d->rbf, d->w_rbf, E, D, d->R, w.xs, nullptr);
  X2_LAUNCH_OK();
  const Lin L{lin_mode(d->mode), w.img, nullptr, st, d->mode == X2_MODE_TF32};
  // (2) Q | K | V | skip                                                     :105-107, :121
  if (L.mode == X2_MODE_TF32X3 && D == kTcBlock) {   // the four projections as problems of ONE launch
    tc::G1Prob pr[4] = {
        {d->x, D, d->w_q, d->b_q, s->qkvs, 4 * D, 0},
        {w.xs, D, d->w_k, d->b_k, s->qkvs + D, 4 * D, 0},
        {w.xs, D, d->w_v, d->b_v, s->qkvs + 2 * D, 4 * D, 0},
        {d->x, D, d->w_skip, d->b_skip, s->qkvs + 3 * D, 4 * D, 0}};
    X2_TRY(tc::tc_gemm_batch(pr, d->fuse_skip ? 4 : 3, d->fuse_skip ? 4 : 3, E, D, 1, D, D, st, L.single));   // one group each
  } else if (L.mode == X2_MODE_TF32X3) {
    X2_TRY(lin_fwd(L, d->x, D, d->w_q, D, d->b_q, s->qkvs, 4 * D, E, D, D));
    X2_TRY(lin_fwd(L, w.xs, D, d->w_k, D, d->b_k, s->qkvs + D, 4 * D, E, D, D));
    X2_TRY(lin_fwd(L, w.xs, D, d->w_v, D, d->b_v, s->qkvs + 2 * D, 4 * D, E, D, D));
    if (d->fuse_skip) X2_TRY(lin_fwd(L, d->x, D, d->w_skip, D, d->b_skip, s->qkvs + 3 * D, 4 * D, E, D, D));
  } else {   // one batched SIMT launch
    GemmBatch p{};
    p.A[0] = d->x;  p.B[0] = d->w_q; p.bias[0] = d->b_q; p.C[0] = s->qkvs;
    p.A[1] = w.xs;  p.B[1] = d->w_k; p.bias[1] = d->b_k; p.C[1] = s->qkvs + D;
    p.A[2] = w.xs;  p.B[2] = d->w_v; p.bias[2] = d->b_v; p.C[2] = s->qkvs + 2 * D;
    int nb = 3;
    if (d->fuse_skip) {
      p.A[3] = d->x; p.B[3] = d->w_skip; p.bias[3] = d->b_skip; p.C[3] = s->qkvs + 3 * D;
      nb = 4;
    }
    X2_TRY((launch_gemm<true, true>(p, nb, E, D, D, D, D, 4 * D, 0, st)));
  }
  phase_end(X2_PHASE_NODE_PROJ, st);
  // (3+4, factorised sbf) lin_edge on its rows, then ONE block-centric attention kernel that forms lin_sbf(sbf_t)
  // from theta_t and a per-source table in shared memory (csrc/blk_attn.cuh): no [T, S] read, no Sg tensor
  if (sgf) {
    if (d->A > 0) {
      const int64_t n_ea = d->ea_index ? d->ea_rows : T;
      X2_TRY(lin_fwd(L, d->edge_attr, d->A, d->w_edge, d->A, nullptr, s->ea, D, n_ea, D, d->A));
    }
    launch_k(k_wsbf_transpose, dim3((unsigned)cdiv((int64_t)D * d->S, 256)), dim3(256), 0, st, d->w_sbf, d->S, w.wsT);
    X2_LAUNCH_OK();
    phase_end(X2_PHASE_TROW_PROJ, st);
    BlkParams bp{};
    blk_fill(d, s, bp, w.wsT);
    bp.attn = s->attn; bp.out = out; bp.lse = s->lse;
    X2_TRY(launch_blk_fwd(d, bp, st));
    phase_end(X2_PHASE_ATTN_FWD, st);
    return X2_OK;
  }
  // (3+4 fused) lin_edge + lin_sbf + segmented attention as ONE tcgen05 kernel over segment-aligned tiles
  // (csrc/tile_attn.cuh): edge_attr / sbf are streamed once, EA / Sg go from tensor memory through a
  // shared-memory slot ring straight into the attention warps.  Opt-in (X2_MODE_TF32X3_FUSED / X2GNN_FUSED=1).
  if (tile_fwd_usable(d, alpha)) {
    if (d->ea_index) X2_TRY(lin_fwd(L, d->edge_attr, d->A, d->w_edge, d->A, nullptr, s->ea, D, d->ea_rows, D, d->A));
    phase_end(X2_PHASE_TROW_PROJ, st);
    X2_TRY(launch_tile_fwd(d, s, out, w.part, st));
    phase_end(X2_PHASE_ATTN_FWD, st);
    return X2_OK;
  }
  // (3) T-row projections                                                    :144, :148
  // segment-constant edge_attr: lin_edge runs on the table rows only
  if (d->ea_index) X2_TRY(lin_fwd(L, d->edge_attr, d->A, d->w_edge, d->A, nullptr, s->ea, D, d->ea_rows, D, d->A));
  if (T > 0) {
    if (d->A > 0 && !d->ea_index) X2_TRY(lin_fwd(L, d->edge_attr, d->A, d->w_edge, d->A, nullptr, s->ea, D, T, D, d->A));
    X2_TRY(lin_fwd(L, d->sbf, d->S, d->w_sbf, d->S, d->b_sbf, s->sg, D, T, D, d->S));
  }
  phase_end(X2_PHASE_TROW_PROJ, st);
  // (4) fused gather + logits + segment softmax + gate + aggregate (+ skip)  :150-160, aggregate, :127
  int rc;
  switch (D / 32) {
    case 1: rc = launch_attn_fwd<1>(d, s, out, alpha, st); break;
    case 2: rc = launch_attn_fwd<2>(d, s, out, alpha, st); break;
    case 4: rc = launch_attn_fwd<4>(d, s, out, alpha, st); break;
    default: rc = launch_attn_fwd<8>(d, s, out, alpha, st); break;
  }
  phase_end(X2_PHASE_ATTN_FWD, st);
  return rc;
}

int x2_sbfconv_bwd(const x2_conv_desc* d, const x2_conv_saved* s, const float* grad_out,
                   const x2_conv_grads* g, void* ws, size_t ws_bytes, void* stream) {
  X2_TRY(check_desc(d));
  X2_CHECK_ARG(s && s->qkvs && s->attn && s->lse && grad_out && g, "conv bwd: null buffer");
  const bool sgf = sgf_usable(d);
  X2_CHECK_ARG(sgf || (s->sg && (d->sbf || d->T == 0)), "conv bwd: sbf / saved.sg required (no usable factorised sbf)");
  X2_CHECK_ARG(!(sgf && g->dsbf), "conv bwd: d sbf is not available with the factorised sbf (pass the dense tensor only)");
  X2_CHECK_ARG(g->dx && g->drbf && g->dw_rbf && g->dw_q && g->db_q && g->dw_k && g->db_k && g->dw_v &&
                   g->db_v && g->dw_sbf && g->db_sbf, "conv bwd: null gradient buffer");
  X2_CHECK_ARG(d->A == 0 || (s->ea && g->dw_edge), "conv bwd: lin_edge buffers required when A > 0");
  X2_CHECK_ARG(!d->fuse_skip || (d->w_skip && g->dw_skip), "conv bwd: fuse_skip needs w_skip/dw_skip");
  cudaStream_t st = (cudaStream_t)stream;
  BwdWs w;
  const size_t need = bwd_layout(d, ws, &w);
  if (ws_bytes < need) { set_error("conv bwd: workspace %zu < %zu", ws_bytes, need); return X2_EWORKSPACE; }
  const int64_t E = d->E, T = d->T;
  const int D = d->D, A = d->A, S = d->S, R = d->R;
  if (E == 0) {
    // no rows: every parameter gradient is zero
    X2_CUDA_OK(cudaMemsetAsync(g->dw_rbf, 0, sizeof(float) * D * R, st));
    float* ws_[] = {g->dw_q, g->dw_k, g->dw_v, d->fuse_skip ? g->dw_skip : nullptr};
    for (float* p : ws_) if (p) X2_CUDA_OK(cudaMemsetAsync(p, 0, sizeof(float) * D * D, st));
    float* bs_[] = {g->db_q, g->db_k, g->db_v, g->db_sbf, d->fuse_skip ? g->db_skip : nullptr};
    for (float* p : bs_) if (p) X2_CUDA_OK(cudaMemsetAsync(p, 0, sizeof(float) * D, st));
    if (A > 0) X2_CUDA_OK(cudaMemsetAsync(g->dw_edge, 0, sizeof(float) * D * A, st));
    X2_CUDA_OK(cudaMemsetAsync(g->dw_sbf, 0, sizeof(float) * D * S, st));
    if (d->ea_index && g->dedge_attr) X2_CUDA_OK(cudaMemsetAsync(g->dedge_attr, 0, sizeof(float) * d->ea_rows * A, st));
    return X2_OK;
  }

  // (1,2) attention backward: by target, then by source
  if (sgf) {
    // factorised sbf: both passes in ONE kernel, one CTA per closed block of the line graph (csrc/blk_attn.cuh)
    phase_begin(st);
    launch_k(k_wsbf_transpose, dim3((unsigned)cdiv((int64_t)D * S, 256)), dim3(256), 0, st, d->w_sbf, S, w.wsT);
    X2_LAUNCH_OK();
    BlkParams bp{};
    blk_fill(d, s, bp, w.wsT);
    bp.gout = grad_out; bp.attn_in = s->attn; bp.lse_in = s->lse;
    bp.dqkv = w.dqkv; bp.ldg = 3 * D;
    bp.dea = w.dea; bp.al = w.al; bp.da = w.da; bp.dP = w.dP;
    X2_TRY(launch_blk_bwd(d, bp, st));
    if (d->ea_index) {       // per-target rows -> per-table-row sums (fixed order)
      launch_k(k_rows_segsum<4>, dim3((unsigned)cdiv(d->ea_rows * 32, 128)), dim3(128), 0, st, w.dea, d->ea_rowptr, d->ea_order,
               d->ea_rows, w.dea_tab);
      X2_LAUNCH_OK();
    }
    phase_end(X2_PHASE_ATTN_BWD_TGT, st);
    phase_end(X2_PHASE_ATTN_BWD_SRC, st);
  } else {
  switch (D / 32) {
    case 1: X2_TRY(launch_attn_bwd<1>(d, s, grad_out, w, st)); break;
    case 2: X2_TRY(launch_attn_bwd<2>(d, s, grad_out, w, st)); break;
    case 4: X2_TRY(launch_attn_bwd<4>(d, s, grad_out, w, st)); break;
    default: X2_TRY(launch_attn_bwd<8>(d, s, grad_out, w, st)); break;
  }
  }
  const float* dq = w.dqkv;
  const float* dk = w.dqkv + D;
  const float* dv = w.dqkv + 2 * D;

  const Lin L{lin_mode(d->mode), w.img, w.wg, st, d->mode == X2_MODE_TF32};
  // (3) T-row input gradients
  const float* dea_rows = d->ea_index ? w.dea_tab : w.dea;          // rows lin_edge was applied to
  const int64_t n_ea = d->ea_index ? d->ea_rows : T;
  if (A > 0 && g->dedge_attr && n_ea > 0) X2_TRY(lin_dgrad(L, dea_rows, D, d->w_edge, A, g->dedge_attr, A, n_ea, A, D, 0));
  if (g->dsbf && T > 0 && !sgf) X2_TRY(lin_dgrad(L, w.dsg, D, d->w_sbf, S, g->dsbf, S, T, S, D, 0));
  phase_end(X2_PHASE_TROW_DGRAD, st);
  // (4) T-row weight gradients (split-K over triplets, fixed-order reduction)
  if (A > 0) X2_TRY(lin_wgrad(L, dea_rows, D, d->edge_attr, A, g->dw_edge, A, nullptr, n_ea, D, A));
  if (sgf) {
    // dW_s / db_s from the per-source d(table) rows the by-source pass left (E-scale, fixed-order reduction)
    const int64_t chunks = cdiv(E, kDwsChunk);
    const int parts = (int)(chunks < kDwsParts ? chunks : kDwsParts);
    float* colsum = w.wg + (size_t)parts * D * S;
    if (d->sbf_L == 7 && d->sbf_R == 6)
      launch_k(k_dws_partial<7, 6>, dim3(parts), dim3(kDwsThreads), 0, st, (const float*)w.dP, d->sbf_tab, E, d->sbf_L, d->sbf_R, w.wg, colsum);
    else
      launch_k(k_dws_partial<0, 0>, dim3(parts), dim3(kDwsThreads), 0, st, (const float*)w.dP, d->sbf_tab, E, d->sbf_L, d->sbf_R, w.wg, colsum);
    X2_LAUNCH_OK();
    launch_k(k_splitk_reduce, dim3(splitk_reduce_blocks(D, S, true)), dim3(256), 0, st, (const float*)w.wg, (const float*)colsum, parts,
             (int64_t)D, S, g->dw_sbf, (int64_t)S, g->db_sbf);
    X2_LAUNCH_OK();
  } else {
    X2_TRY(lin_wgrad(L, w.dsg, D, d->sbf, S, g->dw_sbf, S, g->db_sbf, T, D, S));
  }

  phase_end(X2_PHASE_TROW_WGRAD, st);
  // (5) recompute the filtered sources
  const bool batched = L.mode == X2_MODE_TF32X3 && D == kTcBlock &&
                       ((reinterpret_cast<uintptr_t>(d->x) | reinterpret_cast<uintptr_t>(g->dx)) & 15) == 0;
  const bool fused_tail = batched && R <= 16;        // k_filter_bwd_full recomputes F itself
  if (s->xs && fused_tail) {
    w.xs = s->xs;                                    // the forward kept x * lin_rbf(rbf): nothing to recompute
  } else {
    if ((size_t)D * R * sizeof(float) > 48 * 1024) X2_DYN_SMEM(k_rbf_filter, 64 * 1024);
    launch_k(k_rbf_filter, dim3(filter_grid(E)), dim3(kFilterRows * 32), (size_t)D * R * sizeof(float), st, 
        d->x, d->rbf, d->w_rbf, E, D, R, w.xs, fused_tail ? nullptr : w.F);
    X2_LAUNCH_OK();
  }
  // (6) node-level weight gradients
  if (batched) {          // the four weight gradients as problems of ONE launch (+ one reduction)
    const tc::G2Job jobs[4] = {
        {dq, 3 * D, d->x, D, g->dw_q, D, g->db_q},
        {dk, 3 * D, w.xs, D, g->dw_k, D, g->db_k},
        {dv, 3 * D, w.xs, D, g->dw_v, D, g->db_v},
        {grad_out, D, d->x, D, g->dw_skip, D, g->db_skip}};
    X2_TRY(tc::tc_wgrad_batch(jobs, d->fuse_skip ? 4 : 3, E, D, L.wg, st, L.single));
  } else {
  X2_TRY(lin_wgrad(L, dq, 3 * D, d->x, D, g->dw_q, D, g->db_q, E, D, D));
  X2_TRY(lin_wgrad(L, dk, 3 * D, w.xs, D, g->dw_k, D, g->db_k, E, D, D));
  X2_TRY(lin_wgrad(L, dv, 3 * D, w.xs, D, g->dw_v, D, g->db_v, E, D, D));
  if (d->fuse_skip) X2_TRY(lin_wgrad(L, grad_out, D, d->x, D, g->dw_skip, D, g->db_skip, E, D, D));
  }
  if (batched) {
    // (7) dK W_k, dV W_v and (8) dQ W_q (, G W_o) as independent problems of ONE launch, one group
    // each; k_filter_bwd_sum adds the pairs
    tc::G1Prob pr[4] = {
        {dk, 3 * D, d->w_k, nullptr, w.dxs, D, 0},
        {dv, 3 * D, d->w_v, nullptr, w.dxs2, D, 0},
        {dq, 3 * D, d->w_q, nullptr, g->dx, D, 0},
        {grad_out, D, d->w_skip, nullptr, w.dx2, D, 0}};
    const int np = d->fuse_skip ? 4 : 3;
    X2_TRY(tc::tc_gemm_batch(pr, np, np, E, D, D, 1, D, st, L.single));
    if (fused_tail) {
      // (9 + 10) the whole tail in one pass: dx, d rbf and the per-block partials of dW_r
      const int RM = R <= 8 ? 8 : 16;
      const int64_t fneed = cdiv(E, kFilterRows);
      const unsigned fgrid = (unsigned)(fneed < kNumSM * 4 ? fneed : kNumSM * 4);
      const size_t fsmem = (size_t)kFilterRows * RM * D * sizeof(float);
      X2_DYN_SMEM(k_filter_bwd_full<16>, 65536);
      if (RM == 8)
        launch_k(k_filter_bwd_full<8>, dim3(fgrid), dim3(kFilterRows * 32), fsmem, st, d->x, d->rbf, d->w_rbf, w.dxs, w.dxs2, g->dx,
            d->fuse_skip ? w.dx2 : nullptr, E, R, g->drbf, w.wg);
      else
        launch_k(k_filter_bwd_full<16>, dim3(fgrid), dim3(kFilterRows * 32), fsmem, st, d->x, d->rbf, d->w_rbf, w.dxs, w.dxs2, g->dx,
            d->fuse_skip ? w.dx2 : nullptr, E, R, g->drbf, w.wg);
      X2_LAUNCH_OK();
      launch_k(k_splitk_reduce, dim3(splitk_reduce_blocks(D, R, false)), dim3(256), 0, st, w.wg, nullptr, (int)fgrid, D, R, g->dw_rbf, R,
                                                                         nullptr);
      X2_LAUNCH_OK();
      phase_end(X2_PHASE_NODE_BWD, st);
      return X2_OK;
    }
    k_filter_bwd_sum<<<(unsigned)cdiv(E * D / 4, 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(d->x), reinterpret_cast<const float4*>(w.F),
        reinterpret_cast<float4*>(w.dxs), reinterpret_cast<const float4*>(w.dxs2),
        reinterpret_cast<float4*>(g->dx), d->fuse_skip ? reinterpret_cast<const float4*>(w.dx2) : nullptr,
        E * D / 4);
    X2_LAUNCH_OK();
  } else {
  // (7) dxs = dK W_k + dV W_v
  X2_TRY(lin_dgrad(L, dk, 3 * D, d->w_k, D, w.dxs, D, E, D, D, 0));
  X2_TRY(lin_dgrad(L, dv, 3 * D, d->w_v, D, w.dxs, D, E, D, D, 1));
  // (8) dx = dQ W_q (+ G W_o)
  X2_TRY(lin_dgrad(L, dq, 3 * D, d->w_q, D, g->dx, D, E, D, D, 0));
  if (d->fuse_skip) X2_TRY(lin_dgrad(L, grad_out, D, d->w_skip, D, g->dx, D, E, D, D, 1));
  // (9) dx += dxs * F ; dF = dxs * x
  k_filter_bwd<<<(unsigned)cdiv(E * D, 256), 256, 0, st>>>(d->x, w.F, w.dxs, g->dx, E * D);
  X2_LAUNCH_OK();
  }
  // (10) d rbf = dF W_r ; dW_r = dF^T rbf
  X2_TRY(lin_dgrad(L, w.dxs, D, d->w_rbf, R, g->drbf, R, E, R, D, 0));
  X2_TRY(lin_wgrad(L, w.dxs, D, d->rbf, R, g->dw_rbf, R, nullptr, E, D, R));
  phase_end(X2_PHASE_NODE_BWD, st);
  return X2_OK;
}

}  // extern "C"
