// blk_attn.cuh -- block-centric attention kernels of the SBFTransformerConv layer (sbftransformer_conv.py:138-162
// + PyG softmax / sum aggregation and their backward), included by conv.cu.
//
// A BLOCK (x2_blocks_build, graph.cu) is a contiguous range of source line-nodes together with the targets
// whose segments draw from it; for the line graph of a molecule: the bonds leaving atom j and the bonds
// entering it ("structural facts").  Every triplet has its source and its target in one
// block, so ONE CTA per block
//   * reuses the block's K | V rows (and per-source basis rows) across all of its targets through L1 /
//     shared memory instead of gathering them per triplet from L2,
//   * runs the by-target and the by-source pass of the backward back to back (k_blk_bwd): dQ of its targets
//     and dK / dV of its sources are complete inside the CTA -- no second launch, alpha / d(logit) and the
//     second read of the Sg rows come back from L2 while they are still resident; no atomics, and the
//     accumulation order is the one of k_attn_bwd_tgt / k_attn_bwd_src (bitwise identical results).
//
// SGF = factorised lin_sbf.  F_B_2D's output is sbf[t, l R + n] = tab[src(t), l R + n] *
// Y_l0(theta_t) (angular_basis_layer.py:80-93), so
//   lin_sbf(sbf_t)[c] = b_s[c] + sum_l Y_l0(theta_t) P[src(t), l, c],   P[f, l, c] = sum_n W_s[c, l R + n] tab[f, l R + n].
// The CTA builds P for the sources of its block in shared memory (nS x L x 128 floats: 3.5 KB per source at
// L = 7) and the attention loops form Sg_t from theta_t with L FMAs per channel: the [T, S] sbf stream, the
// T-row lin_sbf GEMM, the saved Sg [T, 128] tensor (written once, read three times) and d(Sg) [T, 128] never
// exist.  Backward: dP[f, l, :] = sum_{t : src(t) = f} Y_l0(theta_t) dSg_t is accumulated by the by-source
// pass in registers; dW_s[c, l R + n] = sum_f dP[f, l, c] tab[f, l R + n] and db_s are E-scale reductions
// (k_dws_partial + k_splitk_reduce, fixed order).
#pragma once

namespace x2 {

constexpr int kBlkThreads = 512;
constexpr int kBlkWarps = kBlkThreads / 32;
constexpr int kBlkLMax = 8;            // spherical orders (num_spherical) the factorised path takes
constexpr int kBlkRMax = 8;            // radial functions per order

struct BlkParams {
  const float* qkvs; int ldq;          // [E, 4 D]  Q | K | V | skip
  const float* ea;                     // EA rows: [T, 128] (kEaTriplet) | lin_edge(table) [M, 128] (kEaSegment)
  const int32_t* ea_index;             // [E] (kEaSegment)
  const float* sg;                     // dense lin_sbf(sbf) [T, 128] (SGF = false)
  const float* stab;                   // [E, S] per-bond radial table (SGF)
  const float* angles;                 // [T]
  const float* wsT;                    // [S, 128] lin_sbf.weight transposed (SGF)
  const float* b_sbf;                  // [128]
  int L, Rr, S;
  const int32_t *src, *tgt, *rowptr_tgt, *rowptr_src, *order_src;
  const int32_t *blk_sptr, *blk_tptr, *blk_tord;
  int64_t E;
  int H, C;
  float scale;
  int fuse_skip;
  float *attn, *out, *lse;             // forward outputs
  const float* gout;                   // backward
  const float* attn_in;
  const float* lse_in;
  float* dqkv; int ldg;
  float *dea, *dsg, *al, *da, *dP;
};

__device__ __forceinline__ float4 ldg_stream4(const float* p) {      // read once: do not allocate in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ex2f(float x) {        // 2^x, MUFU.EX2 (exp2(-inf) = 0)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 ldg_keep4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void stg_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

// Y_l0(theta), l < LT, with the arithmetic of k_sbf_fwd (basis.cu): fp32 Legendre recurrence, normalisation
// constants sqrt((2l+1)/(4 pi)) rounded from fp64.  (Orders l >= L meet zero rows of the staged table.)
template <int LT>
__device__ __forceinline__ void blk_ylm(float theta, const float* __restrict__ ynorm, float (&y)[LT]) {
  const float c = cosf(theta);
  float p0 = 1.f, p1 = c;
  y[0] = ynorm[0];
  if constexpr (LT > 1) y[1] = ynorm[1] * c;
#pragma unroll
  for (int j = 2; j < LT; ++j) {
    const float pj = ((float)(2 * j - 1) * c * p1 - (float)(j - 1) * p0) / (float)j;
    p0 = p1;
    p1 = pj;
    y[j] = ynorm[j] * pj;
  }
}

// P rows of the block's sources into shared memory: sP[(f - s0) LT + l][128]; LT >= L, rows l >= L are zero
template <int LT>
__device__ __forceinline__ void blk_stage_P(const BlkParams& p, int s0, int nS, float* sP) {
  const int L = p.L, Rr = p.Rr, S = p.S;
  for (int idx = threadIdx.x; idx < nS * LT * 32; idx += kBlkThreads) {
    const int c4 = idx & 31, fl = idx >> 5;          // fl = (f - s0) * LT + l
    const int f = fl / LT, l = fl - f * LT;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < L) {
      const float* trow = p.stab + (int64_t)(s0 + f) * S + l * Rr;
      const float* wrow = p.wsT + (int64_t)l * Rr * 128 + c4 * 4;
      for (int n = 0; n < Rr; ++n) {
        const float tv = __ldg(trow + n);
        const float4 w = ldg_keep4(wrow + n * 128);
        acc.x = fmaf(tv, w.x, acc.x); acc.y = fmaf(tv, w.y, acc.y);
        acc.z = fmaf(tv, w.z, acc.z); acc.w = fmaf(tv, w.w, acc.w);
      }
    }
    *reinterpret_cast<float4*>(sP + (size_t)fl * 128 + c4 * 4) = acc;
  }
}

// lin_sbf.weight [128, S] -> [S, 128] (once per call; the staging loop then reads consecutive words)
__global__ void k_wsbf_transpose(const float* __restrict__ w, int S, float* __restrict__ wT) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * S) {
    const int c = i / S, k = i - c * S;
    wT[k * 128 + c] = w[i];
  }
}

// lin_sbf(sbf_t) for the lane's 4 channels from the staged table row of the triplet's source
template <int LT>
__device__ __forceinline__ float4 blk_sg(const float* __restrict__ prow, const float (&y)[LT], float4 g) {
#pragma unroll
  for (int l = 0; l < LT; ++l) {
    const float4 pv = *reinterpret_cast<const float4*>(prow + l * 128);
    g.x = fmaf(y[l], pv.x, g.x); g.y = fmaf(y[l], pv.y, g.y);
    g.z = fmaf(y[l], pv.z, g.z); g.w = fmaf(y[l], pv.w, g.w);
  }
  return g;
}

// ------------------------------------------------------------------ forward
// LT: rows per source of the staged table (the number of spherical orders, or 8 with zero rows); 1 when !SGF
template <int EA, bool SGF, int LPH, int LT>
__global__ void __launch_bounds__(kBlkThreads, 2) k_blk_fwd(const BlkParams p) {
  extern __shared__ __align__(16) float blk_smem[];
  __shared__ float ynorm[kBlkLMax];
  constexpr int D = 128;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kBlkLMax) ynorm[threadIdx.x] = (float)sqrt((double)(2 * threadIdx.x + 1) * 0.07957747154594767);
  pdl_sync();
  const int s0 = p.blk_sptr[b];
  const int tp0 = p.blk_tptr[b], tp1 = p.blk_tptr[b + 1];
  if constexpr (SGF) blk_stage_P<LT>(p, s0, p.blk_sptr[b + 1] - s0, blk_smem);
  __syncthreads();
  const int ch = lane * 4;
  const int head = ch / p.C;
  const int lph = LPH > 0 ? LPH : p.C / 4;
  const bool leader = (ch % p.C) == 0;
  const float scale2 = p.scale * 1.4426950408889634f;        // logits in the log2 domain: ex2 softmax
  const int ldq = p.ldq;
  const float* __restrict__ kvbase = p.qkvs + D + ch;
  const float* __restrict__ sPl = blk_smem + ch;
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (SGF) bs = ldg_keep4(p.b_sbf + ch);

  for (int ti = tp0 + warp; ti < tp1; ti += kBlkWarps) {
    const int e = p.blk_tord[ti];
    float4 q = ldg_keep4(p.qkvs + (int64_t)e * ldq + ch);
    q.x *= scale2; q.y *= scale2; q.z *= scale2; q.w *= scale2;
    const int beg = p.rowptr_tgt[e], end = p.rowptr_tgt[e + 1];
    float m = -INFINITY, z = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 aseg = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (EA == kEaSegment) aseg = ldg_keep4(p.ea + (int64_t)__ldg(p.ea_index + e) * D + ch);
    for (int base = beg; base < end; base += 32) {
      const int my = base + lane;
      int s_l = s0;
      float y_l[LT];
      if (my < end) s_l = __ldg(p.src + my);
      if constexpr (SGF) blk_ylm<LT>(my < end ? __ldg(p.angles + my) : 0.f, ynorm, y_l);
      const int cnt = min(32, end - base);
      // one row of loads ahead of the arithmetic
      float4 k_n, v_n, a_n = aseg, g_n = bs;
      int s_n = __shfl_sync(0xffffffffu, s_l, 0);
      auto fetch = [&](int i) {
        const float* row = kvbase + (int64_t)s_n * ldq;
        k_n = ldg_keep4(row);
        v_n = ldg_keep4(row + D);
        if constexpr (EA == kEaTriplet) a_n = ldg_stream4(p.ea + (int64_t)(base + i) * D + ch);
        if constexpr (!SGF) g_n = ldg_stream4(p.sg + (int64_t)(base + i) * D + ch);
      };
      fetch(0);
      for (int i = 0; i < cnt; ++i) {
        const float4 k = k_n, v = v_n, a = a_n;
        float4 g = g_n;
        if constexpr (SGF) {
          float y[LT];
#pragma unroll
          for (int l = 0; l < LT; ++l) y[l] = __shfl_sync(0xffffffffu, y_l[l], i);
          g = blk_sg<LT>(sPl + (s_n - s0) * (LT * 128), y, bs);
        }
        const int inext = min(i + 1, cnt - 1);
        s_n = __shfl_sync(0xffffffffu, s_l, inext);
        fetch(inext);
        float dot = q.x * (k.x + a.x);
        dot = fmaf(q.y, k.y + a.y, dot);
        dot = fmaf(q.z, k.z + a.z, dot);
        dot = fmaf(q.w, k.w + a.w, dot);
        const float lg = head_sum_t<LPH>(dot, lph);                     // :150, in log2 units
        const float mn = fmaxf(m, lg);
        const float corr = ex2f(m - mn);
        const float pe = ex2f(lg - mn);
        z = fmaf(z, corr, pe);
        acc.x = fmaf(acc.x, corr, pe * (v.x + a.x) * g.x);              // :155-160
        acc.y = fmaf(acc.y, corr, pe * (v.y + a.y) * g.y);
        acc.z = fmaf(acc.z, corr, pe * (v.z + a.z) * g.z);
        acc.w = fmaf(acc.w, corr, pe * (v.w + a.w) * g.w);
        m = mn;
      }
    }
    const float inv = 1.0f / (z + 1e-16f);                              // PyG softmax: out / (sum + 1e-16)
    float4 o = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    *reinterpret_cast<float4*>(p.attn + (int64_t)e * D + ch) = o;
    if (p.fuse_skip) {
      const float4 sk = ldg_keep4(p.qkvs + (int64_t)e * ldq + 3 * D + ch);
      o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;               // :127
    }
    *reinterpret_cast<float4*>(p.out + (int64_t)e * D + ch) = o;
    // natural-log log-sum-exp, as the generic kernels save it
    if (leader) p.lse[(int64_t)e * p.H + head] = (end > beg) ? (m + log2f(z)) * 0.6931471805599453f : 0.f;
  }
}

// ------------------------------------------------------------------ backward: by target, then by source, one CTA
template <int EA, bool SGF, int LPH, int LT>
__global__ void __launch_bounds__(kBlkThreads, SGF ? 1 : 2) k_blk_bwd(const BlkParams p) {
  extern __shared__ __align__(16) float blk_smem[];
  __shared__ float ynorm[kBlkLMax];
  constexpr int D = 128;
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kBlkLMax) ynorm[threadIdx.x] = (float)sqrt((double)(2 * threadIdx.x + 1) * 0.07957747154594767);
  pdl_sync();
  const int s0 = p.blk_sptr[b], s1 = p.blk_sptr[b + 1];
  const int tp0 = p.blk_tptr[b], tp1 = p.blk_tptr[b + 1];
  if constexpr (SGF) blk_stage_P<LT>(p, s0, s1 - s0, blk_smem);
  __syncthreads();
  const int ch = lane * 4;
  const int head = ch / p.C;
  const int lph = LPH > 0 ? LPH : p.C / 4;
  const bool leader = (ch % p.C) == 0;
  const int H = p.H;
  const int ldq = p.ldq;
  const float scale = p.scale;
  const float scale2 = scale * 1.4426950408889634f;
  const float* __restrict__ kvbase = p.qkvs + D + ch;
  const float* __restrict__ sPl = blk_smem + ch;
  float4 bs = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (SGF) bs = ldg_keep4(p.b_sbf + ch);

  // ---- pass A: warp per target (the arithmetic and the order of k_attn_bwd_tgt)
  for (int ti = tp0 + warp; ti < tp1; ti += kBlkWarps) {
    const int e = p.blk_tord[ti];
    const float4 q = ldg_keep4(p.qkvs + (int64_t)e * ldq + ch);
    const float4 g = ldg_keep4(p.gout + (int64_t)e * D + ch);
    const float4 o = ldg_keep4(p.attn_in + (int64_t)e * D + ch);
    float r = g.x * o.x;
    r = fmaf(g.y, o.y, r); r = fmaf(g.z, o.z, r); r = fmaf(g.w, o.w, r);
    r = head_sum_t<LPH>(r, lph);                   // r_eh = sum_t alpha dalpha = <G, O>  (App. A)
    const float lse2 = __ldg(p.lse_in + (int64_t)e * H + head) * 1.4426950408889634f;
    const int beg = p.rowptr_tgt[e], end = p.rowptr_tgt[e + 1];
    float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), dea_acc = dq;
    float4 aseg = dq;
    if constexpr (EA == kEaSegment) aseg = ldg_keep4(p.ea + (int64_t)__ldg(p.ea_index + e) * D + ch);
    for (int base = beg; base < end; base += 32) {
      const int my = base + lane;
      int s_l = s0;
      float y_l[LT];
      if (my < end) s_l = __ldg(p.src + my);
      if constexpr (SGF) blk_ylm<LT>(my < end ? __ldg(p.angles + my) : 0.f, ynorm, y_l);
      const int cnt = min(32, end - base);
      float4 k_n, v_n, a_n = aseg, g_n = bs;
      int s_n = __shfl_sync(0xffffffffu, s_l, 0);
      auto fetch = [&](int i) {
        const float* row = kvbase + (int64_t)s_n * ldq;
        k_n = ldg_keep4(row);
        v_n = ldg_keep4(row + D);
        if constexpr (EA == kEaTriplet) a_n = ldg_stream4(p.ea + (int64_t)(base + i) * D + ch);
        if constexpr (!SGF) g_n = ldg_keep4(p.sg + (int64_t)(base + i) * D + ch);   // read again by pass B
      };
      fetch(0);
      for (int i = 0; i < cnt; ++i) {
        const int64_t t = base + i;
        float4 kk = k_n, vv = v_n;
        const float4 a = a_n;
        float4 sgv = g_n;
        if constexpr (SGF) {
          float y[LT];
#pragma unroll
          for (int l = 0; l < LT; ++l) y[l] = __shfl_sync(0xffffffffu, y_l[l], i);
          sgv = blk_sg<LT>(sPl + (s_n - s0) * (LT * 128), y, bs);
        }
        const int inext = min(i + 1, cnt - 1);
        s_n = __shfl_sync(0xffffffffu, s_l, inext);
        fetch(inext);
        kk.x += a.x; kk.y += a.y; kk.z += a.z; kk.w += a.w;
        vv.x += a.x; vv.y += a.y; vv.z += a.z; vv.w += a.w;
        float dot = q.x * kk.x, dal = g.x * vv.x * sgv.x;
        dot = fmaf(q.y, kk.y, dot); dal = fmaf(g.y * vv.y, sgv.y, dal);
        dot = fmaf(q.z, kk.z, dot); dal = fmaf(g.z * vv.z, sgv.z, dal);
        dot = fmaf(q.w, kk.w, dot); dal = fmaf(g.w * vv.w, sgv.w, dal);
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
        const float alpha = ex2f(fmaf(dot, scale2, -lse2));
        const float da = alpha * (dal - r);
        const float sda = scale * da;
        dq.x = fmaf(sda, kk.x, dq.x); dq.y = fmaf(sda, kk.y, dq.y);
        dq.z = fmaf(sda, kk.z, dq.z); dq.w = fmaf(sda, kk.w, dq.w);
        if constexpr (EA != kEaNone) {
          const float4 oe = make_float4(fmaf(sda, q.x, g.x * sgv.x * alpha), fmaf(sda, q.y, g.y * sgv.y * alpha),
                                        fmaf(sda, q.z, g.z * sgv.z * alpha), fmaf(sda, q.w, g.w * sgv.w * alpha));
          if constexpr (EA == kEaSegment) {
            dea_acc.x += oe.x; dea_acc.y += oe.y; dea_acc.z += oe.z; dea_acc.w += oe.w;
          } else {
            stg_stream4(p.dea + t * D + ch, oe);
          }
        }
        if constexpr (!SGF)
          stg_stream4(p.dsg + t * D + ch, make_float4(g.x * vv.x * alpha, g.y * vv.y * alpha, g.z * vv.z * alpha,
                                                      g.w * vv.w * alpha));
        if (leader) {
          p.al[t * H + head] = alpha;
          p.da[t * H + head] = da;
        }
      }
    }
    *reinterpret_cast<float4*>(p.dqkv + (int64_t)e * p.ldg + ch) = dq;
    if constexpr (EA == kEaSegment) *reinterpret_cast<float4*>(p.dea + (int64_t)e * D + ch) = dea_acc;
  }
  __syncthreads();       // alpha / d(logit) of every triplet of the block are written (and visible to the CTA)

  // ---- pass B: warp per source (the arithmetic and the order of k_attn_bwd_src)
  for (int f = s0 + warp; f < s1; f += kBlkWarps) {
    float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = dk;
    float4 Pl[LT], dP[LT];
    float4 vf = dk;
    if constexpr (SGF) {
      const float* pr = sPl + (f - s0) * (LT * 128);
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        Pl[l] = *reinterpret_cast<const float4*>(pr + l * 128);
        dP[l] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      vf = ldg_keep4(p.qkvs + (int64_t)f * ldq + 2 * D + ch);
    }
    const int beg = p.rowptr_src[f], end = p.rowptr_src[f + 1];
    for (int base = beg; base < end; base += 32) {
      const int my = base + lane;
      int t_l = 0, e_l = 0;
      float y_l[LT];
      if (my < end) {
        t_l = __ldg(p.order_src + my);
        e_l = __ldg(p.tgt + t_l);
      }
      if constexpr (SGF) blk_ylm<LT>(my < end ? __ldg(p.angles + t_l) : 0.f, ynorm, y_l);
      const int cnt = min(32, end - base);
#pragma unroll 2
      for (int i = 0; i < cnt; ++i) {
        const int64_t t = __shfl_sync(0xffffffffu, t_l, i);
        const int e = __shfl_sync(0xffffffffu, e_l, i);
        const float4 q = ldg_keep4(p.qkvs + (int64_t)e * ldq + ch);
        const float4 g = ldg_keep4(p.gout + (int64_t)e * D + ch);
        const float a = __ldcg(p.al + t * H + head);          // written by pass A of this CTA: L2, not the read-only path
        const float sda = scale * __ldcg(p.da + t * H + head);
        float4 sgv;
        if constexpr (SGF) {
          sgv = bs;
          float yv[LT];
#pragma unroll
          for (int l = 0; l < LT; ++l) {
            yv[l] = __shfl_sync(0xffffffffu, y_l[l], i);
            sgv.x = fmaf(yv[l], Pl[l].x, sgv.x); sgv.y = fmaf(yv[l], Pl[l].y, sgv.y);
            sgv.z = fmaf(yv[l], Pl[l].z, sgv.z); sgv.w = fmaf(yv[l], Pl[l].w, sgv.w);
          }
          float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (EA == kEaTriplet) ev = ldg_keep4(p.ea + t * D + ch);
          if constexpr (EA == kEaSegment) ev = ldg_keep4(p.ea + (int64_t)__ldg(p.ea_index + e) * D + ch);
          const float4 ds = make_float4(g.x * (vf.x + ev.x) * a, g.y * (vf.y + ev.y) * a, g.z * (vf.z + ev.z) * a,
                                        g.w * (vf.w + ev.w) * a);                      // d(Sg_t)
#pragma unroll
          for (int l = 0; l < LT; ++l) {
            dP[l].x = fmaf(yv[l], ds.x, dP[l].x); dP[l].y = fmaf(yv[l], ds.y, dP[l].y);
            dP[l].z = fmaf(yv[l], ds.z, dP[l].z); dP[l].w = fmaf(yv[l], ds.w, dP[l].w);
          }
        } else {
          sgv = ldg_keep4(p.sg + t * D + ch);
        }
        dv.x = fmaf(g.x * sgv.x, a, dv.x); dv.y = fmaf(g.y * sgv.y, a, dv.y);      // dV[s] += G . Sg . alpha
        dv.z = fmaf(g.z * sgv.z, a, dv.z); dv.w = fmaf(g.w * sgv.w, a, dv.w);
        dk.x = fmaf(sda, q.x, dk.x); dk.y = fmaf(sda, q.y, dk.y);                  // dK[s] += sigma da Q[e]
        dk.z = fmaf(sda, q.z, dk.z); dk.w = fmaf(sda, q.w, dk.w);
      }
    }
    *reinterpret_cast<float4*>(p.dqkv + (int64_t)f * p.ldg + D + ch) = dk;
    *reinterpret_cast<float4*>(p.dqkv + (int64_t)f * p.ldg + 2 * D + ch) = dv;
    if constexpr (SGF) {
      const int L = p.L;
#pragma unroll
      for (int l = 0; l < LT; ++l)
        if (l < L) stg_stream4(p.dP + ((int64_t)f * L + l) * D + ch, dP[l]);
    }
  }
}

// dW_s[c, l R + n] = sum_f dP[f, l, c] tab[f, l R + n] ;  db_s[c] = sum_t d(Sg_t)[c] = sum_f dP[f, 0, c] / Y_00.
// CTA z handles the sources z, z + grid, ... (fixed assignment) and writes one partial tile [128][S] (+ [128]);
// k_splitk_reduce sums the tiles in fixed order.
constexpr int kDwsThreads = 256;
__global__ void __launch_bounds__(kDwsThreads)
k_dws_partial(const float* __restrict__ dP, const float* __restrict__ stab, int64_t E, int L, int Rr,
              float* __restrict__ partial, float* __restrict__ colsum) {
  pdl_sync();
  const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int S = L * Rr;
  float acc[kBlkLMax / 2][kBlkRMax];
#pragma unroll
  for (int i = 0; i < kBlkLMax / 2; ++i)
#pragma unroll
    for (int n = 0; n < kBlkRMax; ++n) acc[i][n] = 0.f;
  float accb = 0.f;
  for (int64_t f = blockIdx.x; f < E; f += gridDim.x) {
    const float* prow = dP + f * L * 128 + c;
    const float* trow = stab + f * S;
#pragma unroll
    for (int i = 0; i < kBlkLMax / 2; ++i) {
      const int l = 2 * i + half;
      if (l < L) {
        const float v = prow[l * 128];
        if (l == 0) accb += v;
#pragma unroll
        for (int n = 0; n < kBlkRMax; ++n)
          if (n < Rr) acc[i][n] = fmaf(v, __ldg(trow + l * Rr + n), acc[i][n]);
      }
    }
  }
  float* out = partial + (int64_t)blockIdx.x * 128 * S + (int64_t)c * S;
#pragma unroll
  for (int i = 0; i < kBlkLMax / 2; ++i) {
    const int l = 2 * i + half;
    if (l < L) {
#pragma unroll
      for (int n = 0; n < kBlkRMax; ++n)
        if (n < Rr) out[l * Rr + n] = acc[i][n];
    }
  }
  if (half == 0) colsum[(int64_t)blockIdx.x * 128 + c] = accb * (1.0f / (float)sqrt(0.07957747154594767));
}

}  // namespace x2
