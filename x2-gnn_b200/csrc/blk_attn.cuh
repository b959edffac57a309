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
//     and dK / dV of its sources are complete inside the CTA -- no second launch, alpha / d(logit) of the
//     block's triplets stay in shared memory between the passes; no atomics, fixed accumulation order.
// (With a DENSE lin_sbf output the same one-kernel backward was measured slower than the two generic kernels --
// 0.58 vs 0.47 ms on the bench batch, profiles/r2_notes.md -- so these kernels serve the factorised form only.)
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

constexpr int kBlkHalf = 512;          // threads that work on one block (16 warps)
constexpr int kBlkWarps = kBlkHalf / 32;
constexpr int kBlkLMax = 8;            // spherical orders (num_spherical) the factorised path takes
constexpr int kBlkRMax = 8;            // radial functions per order

struct BlkParams {
  const float* qkvs; int ldq;          // [E, 4 D]  Q | K | V | skip
  const float* ea;                     // EA rows: [T, 128] (kEaTriplet) | lin_edge(table) [M, 128] (kEaSegment)
  const int32_t* ea_index;             // [E] (kEaSegment)
  const float* stab;                   // [E, S] per-bond radial table
  const float* angles;                 // [T]
  const float* wsT;                    // [S, 128] lin_sbf.weight transposed
  const float* b_sbf;                  // [128]
  int L, Rr, S;
  const int32_t *src, *tgt, *rowptr_tgt, *rowptr_src, *order_src;
  const int32_t *blk_sptr, *blk_tptr, *blk_tord, *blk_tpos;
  int nblk, maxS, maxTrip, maxTgt;
  int64_t E;
  int H, C;
  float scale;
  int fuse_skip;
  float *attn, *out, *lse;             // forward outputs
  const float* gout;                   // backward
  const float* attn_in;
  const float* lse_in;
  float* dqkv; int ldg;
  float *dea, *al, *da, *dP;
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
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
// barrier of one 512-thread half of the CTA (ids 1, 2; id 0 is __syncthreads)
__device__ __forceinline__ void half_sync(int half) { asm volatile("bar.sync %0, %1;" ::"r"(half + 1), "r"(kBlkHalf) : "memory"); }

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

// P rows of the block's sources into shared memory: sP[(f - s0) LT + l][128] (LT >= L, rows l >= L are zero), from
// the shared-memory copy of W_s^T.  A warp per (source, order): lane n holds tab[f, l R + n], every lane owns 4
// channels.  `wid` / `nw`: the warp's index among the warps that stage.
template <int LT>
__device__ __forceinline__ void blk_stage_P(const BlkParams& p, const float* __restrict__ sW, int s0, int nS, float* sP,
                                            int wid, int nw, int lane) {
  const int L = p.L, Rr = p.Rr, S = p.S;
  for (int fl = wid; fl < nS * LT; fl += nw) {
    const int f = fl / LT, l = fl - f * LT;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (l < L) {
      const float tv_l = __ldg(p.stab + (int64_t)(s0 + f) * S + l * Rr + (lane < Rr ? lane : 0));
      const float* wrow = sW + (size_t)l * Rr * 128 + lane * 4;
#pragma unroll
      for (int n = 0; n < kBlkRMax; ++n) {
        const float tv = __shfl_sync(0xffffffffu, tv_l, n);
        if (n < Rr) {
          const float4 w = *reinterpret_cast<const float4*>(wrow + n * 128);
          acc.x = fmaf(tv, w.x, acc.x); acc.y = fmaf(tv, w.y, acc.y);
          acc.z = fmaf(tv, w.z, acc.z); acc.w = fmaf(tv, w.w, acc.w);
        }
      }
    }
    *reinterpret_cast<float4*>(sP + (size_t)fl * 128 + lane * 4) = acc;
  }
}

// lin_sbf.weight [128, S] -> [S, 128] (once per call; the kernels then copy it to shared memory with 128-bit loads)
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
// Persistent: CTA = nh halves of 512 threads (nh = blockDim.x / 512 = 2 when two staged tables fit shared memory);
// a half walks the blocks h, h + stride, ... on its own (named barriers).  Shared memory: W_s^T [S][128] | one table
// [maxS][LT][128] per half.  LT: rows per source of the staged table (7, or 8 with zero rows).
template <int EA, int LPH, int LT>
__global__ void __launch_bounds__(2 * kBlkHalf, 1) k_blk_fwd(const BlkParams p) {
  extern __shared__ __align__(16) float blk_smem[];
  __shared__ float ynorm[kBlkLMax];
  constexpr int D = 128;
  const int half = threadIdx.x / kBlkHalf, nh = blockDim.x / kBlkHalf;
  const int tid = threadIdx.x - half * kBlkHalf;
  const int warp = tid >> 5, lane = tid & 31;
  float* sW = blk_smem;
  float* sP = blk_smem + (size_t)p.S * 128 + (size_t)half * p.maxS * LT * 128;
  if (threadIdx.x < kBlkLMax) ynorm[threadIdx.x] = (float)sqrt((double)(2 * threadIdx.x + 1) * 0.07957747154594767);
  pdl_sync();
  for (int i = threadIdx.x; i < p.S * 32; i += blockDim.x)
    reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(p.wsT) + i);
  __syncthreads();
  const int ch = lane * 4;
  const int head = ch / p.C;
  const int lph = LPH > 0 ? LPH : p.C / 4;
  const bool leader = (ch % p.C) == 0;
  const float scale2 = p.scale * 1.4426950408889634f;        // logits in the log2 domain: ex2 softmax
  const int ldq = p.ldq;
  const float* __restrict__ kvbase = p.qkvs + D + ch;
  const float* __restrict__ sPl = sP + ch;
  const float4 bs = ldg_keep4(p.b_sbf + ch);

  for (int b = blockIdx.x * nh + half; b < p.nblk; b += gridDim.x * nh) {
    const int s0 = p.blk_sptr[b], nS = p.blk_sptr[b + 1] - s0;
    const int tp0 = p.blk_tptr[b], tp1 = p.blk_tptr[b + 1];
    half_sync(half);                         // the previous block's readers of the table are done
    for (int i = tid; i < nS * 8; i += kBlkHalf)      // K | V rows of the block: 8 lines each, towards L1
      prefetch_l1(p.qkvs + (int64_t)(s0 + (i >> 3)) * ldq + D + (i & 7) * 32);
    blk_stage_P<LT>(p, sW, s0, nS, sP, warp, kBlkWarps, lane);
    half_sync(half);
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
        blk_ylm<LT>(my < end ? __ldg(p.angles + my) : 0.f, ynorm, y_l);
        const int cnt = min(32, end - base);
        // one row of loads ahead of the arithmetic
        float4 k_n, v_n, a_n = aseg;
        int s_n = __shfl_sync(0xffffffffu, s_l, 0);
        auto fetch = [&](int i) {
          const float* row = kvbase + (int64_t)s_n * ldq;
          k_n = ldg_keep4(row);
          v_n = ldg_keep4(row + D);
          if constexpr (EA == kEaTriplet) a_n = ldg_stream4(p.ea + (int64_t)(base + i) * D + ch);
        };
        fetch(0);
        for (int i = 0; i < cnt; ++i) {
          const float4 k = k_n, v = v_n, a = a_n;
          float y[LT];
#pragma unroll
          for (int l = 0; l < LT; ++l) y[l] = __shfl_sync(0xffffffffu, y_l[l], i);
          const float4 g = blk_sg<LT>(sPl + (s_n - s0) * (LT * 128), y, bs);
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
      // log-sum-exp in LOG2 units (the generic kernels save the natural log): k_blk_bwd, the only reader on this
      // path, forms alpha = 2^(logit2 - lse2) without another rounding of a large number
      if (leader) p.lse[(int64_t)e * p.H + head] = (end > beg) ? m + log2f(z) : 0.f;
    }
  }
}

// ------------------------------------------------------------------ backward: by target, then by source, one CTA
// Persistent, 512 threads.  Shared memory: W_s^T | table [maxS][LT][128] | tbase [maxTgt + 1] | (ALDA) alpha / d(logit)
// of the block's triplets [maxTrip][2 H] -- pass A leaves them there for pass B; when they do not fit they go
// through the al / da scratch in global memory (L2).
template <int EA, int LPH, int LT, bool ALDA>
__global__ void __launch_bounds__(kBlkHalf, 1) k_blk_bwd(const BlkParams p) {
  extern __shared__ __align__(16) float blk_smem[];
  __shared__ float ynorm[kBlkLMax];
  constexpr int D = 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sW = blk_smem;
  float* sP = sW + (size_t)p.S * 128;
  int* tbase = reinterpret_cast<int*>(sP + (size_t)p.maxS * LT * 128);
  float* sAD = reinterpret_cast<float*>(tbase + ((p.maxTgt + 1 + 3) & ~3));
  if (threadIdx.x < kBlkLMax) ynorm[threadIdx.x] = (float)sqrt((double)(2 * threadIdx.x + 1) * 0.07957747154594767);
  pdl_sync();
  for (int i = threadIdx.x; i < p.S * 32; i += kBlkHalf)
    reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(p.wsT) + i);
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
  const float* __restrict__ sPl = sP + ch;
  const float4 bs = ldg_keep4(p.b_sbf + ch);

  for (int b = blockIdx.x; b < p.nblk; b += gridDim.x) {
    const int s0 = p.blk_sptr[b], s1 = p.blk_sptr[b + 1];
    const int tp0 = p.blk_tptr[b], tp1 = p.blk_tptr[b + 1];
    __syncthreads();                         // the previous block's pass B is done with the table and alpha / d(logit)
    for (int i = threadIdx.x; i < (s1 - s0) * 8; i += kBlkHalf)
      prefetch_l1(p.qkvs + (int64_t)(s0 + (i >> 3)) * ldq + D + (i & 7) * 32);
    for (int i = threadIdx.x; i < (tp1 - tp0) * 8; i += kBlkHalf) {      // Q and G rows of the block's targets
      const int e = __ldg(p.blk_tord + tp0 + (i >> 3));
      const int part = i & 7;
      prefetch_l1(part < 4 ? p.qkvs + (int64_t)e * ldq + part * 32 : p.gout + (int64_t)e * D + (part - 4) * 32);
    }
    if (warp == kBlkWarps - 1) {             // tbase[i] = first local triplet index of the block's i-th target
      int run = 0;
      for (int i0 = 0; i0 < tp1 - tp0; i0 += 32) {
        int len = 0;
        if (i0 + lane < tp1 - tp0) {
          const int e = __ldg(p.blk_tord + tp0 + i0 + lane);
          len = __ldg(p.rowptr_tgt + e + 1) - __ldg(p.rowptr_tgt + e);
        }
        int inc = len;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += v;
        }
        if (i0 + lane < tp1 - tp0) tbase[i0 + lane] = run + inc - len;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
    }
    blk_stage_P<LT>(p, sW, s0, s1 - s0, sP, warp, kBlkWarps, lane);
    __syncthreads();

    // ---- pass A: warp per target (the arithmetic of k_attn_bwd_tgt)
    for (int ti = tp0 + warp; ti < tp1; ti += kBlkWarps) {
      const int e = p.blk_tord[ti];
      const float4 q = ldg_keep4(p.qkvs + (int64_t)e * ldq + ch);
      const float4 g = ldg_keep4(p.gout + (int64_t)e * D + ch);
      const float4 o = ldg_keep4(p.attn_in + (int64_t)e * D + ch);
      float r = g.x * o.x;
      r = fmaf(g.y, o.y, r); r = fmaf(g.z, o.z, r); r = fmaf(g.w, o.w, r);
      r = head_sum_t<LPH>(r, lph);                   // r_eh = sum_t alpha dalpha = <G, O>  (App. A)
      const float lse2 = __ldg(p.lse_in + (int64_t)e * H + head);          // log2 units (k_blk_fwd)
      const int beg = p.rowptr_tgt[e], end = p.rowptr_tgt[e + 1];
      float* adrow = sAD + (size_t)tbase[ti - tp0] * (2 * H) + head;
      float4 dq = make_float4(0.f, 0.f, 0.f, 0.f), dea_acc = dq;
      float4 aseg = dq;
      if constexpr (EA == kEaSegment) aseg = ldg_keep4(p.ea + (int64_t)__ldg(p.ea_index + e) * D + ch);
      for (int base = beg; base < end; base += 32) {
        const int my = base + lane;
        int s_l = s0;
        float y_l[LT];
        if (my < end) s_l = __ldg(p.src + my);
        blk_ylm<LT>(my < end ? __ldg(p.angles + my) : 0.f, ynorm, y_l);
        const int cnt = min(32, end - base);
        float4 k_n, v_n, a_n = aseg;
        int s_n = __shfl_sync(0xffffffffu, s_l, 0);
        auto fetch = [&](int i) {
          const float* row = kvbase + (int64_t)s_n * ldq;
          k_n = ldg_keep4(row);
          v_n = ldg_keep4(row + D);
          if constexpr (EA == kEaTriplet) a_n = ldg_stream4(p.ea + (int64_t)(base + i) * D + ch);
        };
        fetch(0);
        for (int i = 0; i < cnt; ++i) {
          const int64_t t = base + i;
          float4 kk = k_n, vv = v_n;
          const float4 a = a_n;
          float y[LT];
#pragma unroll
          for (int l = 0; l < LT; ++l) y[l] = __shfl_sync(0xffffffffu, y_l[l], i);
          const float4 sgv = blk_sg<LT>(sPl + (s_n - s0) * (LT * 128), y, bs);
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
          if (leader) {
            if constexpr (ALDA) {
              adrow[0] = alpha;
              adrow[H] = da;
            } else {
              p.al[t * H + head] = alpha;
              p.da[t * H + head] = da;
            }
          }
          adrow += 2 * H;
        }
      }
      *reinterpret_cast<float4*>(p.dqkv + (int64_t)e * p.ldg + ch) = dq;
      if constexpr (EA == kEaSegment) *reinterpret_cast<float4*>(p.dea + (int64_t)e * D + ch) = dea_acc;
    }
    __syncthreads();       // alpha / d(logit) of every triplet of the block are written (and visible to the CTA)

    // ---- pass B: warp per source (the arithmetic of k_attn_bwd_src + the d(table) rows)
    for (int f = s0 + warp; f < s1; f += kBlkWarps) {
      float4 dk = make_float4(0.f, 0.f, 0.f, 0.f), dv = dk;
      float4 Pl[LT], dP[LT];
      const float* pr = sPl + (f - s0) * (LT * 128);
#pragma unroll
      for (int l = 0; l < LT; ++l) {
        Pl[l] = *reinterpret_cast<const float4*>(pr + l * 128);
        dP[l] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const float4 vf = ldg_keep4(p.qkvs + (int64_t)f * ldq + 2 * D + ch);
      const int beg = p.rowptr_src[f], end = p.rowptr_src[f + 1];
      for (int base = beg; base < end; base += 32) {
        const int my = base + lane;
        int t_l = 0, e_l = 0, loc_l = 0;
        float y_l[LT];
        if (my < end) {
          t_l = __ldg(p.order_src + my);
          e_l = __ldg(p.tgt + t_l);
          if constexpr (ALDA) loc_l = tbase[__ldg(p.blk_tpos + e_l)] + t_l - __ldg(p.rowptr_tgt + e_l);
        }
        blk_ylm<LT>(my < end ? __ldg(p.angles + t_l) : 0.f, ynorm, y_l);
        const int cnt = min(32, end - base);
        float4 q_n, g_n;
        auto fetch = [&](int i) {
          const int e = __shfl_sync(0xffffffffu, e_l, i);
          q_n = ldg_keep4(p.qkvs + (int64_t)e * ldq + ch);
          g_n = ldg_keep4(p.gout + (int64_t)e * D + ch);
        };
        fetch(0);
        for (int i = 0; i < cnt; ++i) {
          const float4 q = q_n, g = g_n;
          float a, sda;
          if constexpr (ALDA) {
            const float* ad = sAD + (size_t)__shfl_sync(0xffffffffu, loc_l, i) * (2 * H) + head;
            a = ad[0];
            sda = scale * ad[H];
          } else {
            const int64_t t = __shfl_sync(0xffffffffu, t_l, i);
            a = __ldcg(p.al + t * H + head);          // written by pass A of this CTA: L2, not the read-only path
            sda = scale * __ldcg(p.da + t * H + head);
          }
          float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (EA == kEaTriplet) ev = ldg_keep4(p.ea + (int64_t)__shfl_sync(0xffffffffu, t_l, i) * D + ch);
          if constexpr (EA == kEaSegment)
            ev = ldg_keep4(p.ea + (int64_t)__ldg(p.ea_index + __shfl_sync(0xffffffffu, e_l, i)) * D + ch);
          fetch(min(i + 1, cnt - 1));
          float4 sgv = bs;
          float yv[LT];
#pragma unroll
          for (int l = 0; l < LT; ++l) {
            yv[l] = __shfl_sync(0xffffffffu, y_l[l], i);
            sgv.x = fmaf(yv[l], Pl[l].x, sgv.x); sgv.y = fmaf(yv[l], Pl[l].y, sgv.y);
            sgv.z = fmaf(yv[l], Pl[l].z, sgv.z); sgv.w = fmaf(yv[l], Pl[l].w, sgv.w);
          }
          const float4 ds = make_float4(g.x * (vf.x + ev.x) * a, g.y * (vf.y + ev.y) * a, g.z * (vf.z + ev.z) * a,
                                        g.w * (vf.w + ev.w) * a);                      // d(Sg_t)
#pragma unroll
          for (int l = 0; l < LT; ++l) {
            dP[l].x = fmaf(yv[l], ds.x, dP[l].x); dP[l].y = fmaf(yv[l], ds.y, dP[l].y);
            dP[l].z = fmaf(yv[l], ds.z, dP[l].z); dP[l].w = fmaf(yv[l], ds.w, dP[l].w);
          }
          dv.x = fmaf(g.x * sgv.x, a, dv.x); dv.y = fmaf(g.y * sgv.y, a, dv.y);      // dV[s] += G . Sg . alpha
          dv.z = fmaf(g.z * sgv.z, a, dv.z); dv.w = fmaf(g.w * sgv.w, a, dv.w);
          dk.x = fmaf(sda, q.x, dk.x); dk.y = fmaf(sda, q.y, dk.y);                  // dK[s] += sigma da Q[e]
          dk.z = fmaf(sda, q.z, dk.z); dk.w = fmaf(sda, q.w, dk.w);
        }
      }
      *reinterpret_cast<float4*>(p.dqkv + (int64_t)f * p.ldg + D + ch) = dk;
      *reinterpret_cast<float4*>(p.dqkv + (int64_t)f * p.ldg + 2 * D + ch) = dv;
      const int L = p.L;
#pragma unroll
      for (int l = 0; l < LT; ++l)
        if (l < L) stg_stream4(p.dP + ((int64_t)f * L + l) * D + ch, dP[l]);
    }
  }
}

// dW_s[c, l R + n] = sum_f dP[f, l, c] tab[f, l R + n] ;  db_s[c] = sum_t d(Sg_t)[c] = sum_f dP[f, 0, c] / Y_00.
// CTA z handles the 16-source chunks z, z + grid, ... (fixed assignment): the chunk's tab rows go to shared
// memory, thread (c, half) accumulates the orders l = half, half + 2, ...; one partial tile [128][S] (+ [128]) per
// CTA, summed by k_splitk_reduce in fixed order.  LC / RC: compile-time L / R (0 = run-time, <= 8).
constexpr int kDwsThreads = 256;
constexpr int kDwsChunk = 16;
template <int LC, int RC>
__global__ void __launch_bounds__(kDwsThreads)
k_dws_partial(const float* __restrict__ dP, const float* __restrict__ stab, int64_t E, int Lr, int Rrr,
              float* __restrict__ partial, float* __restrict__ colsum) {
  constexpr int LM = LC > 0 ? LC : kBlkLMax, RM = RC > 0 ? RC : kBlkRMax;
  constexpr int NL = (LM + 1) / 2;
  __shared__ float stile[kDwsChunk * LM * RM];
  pdl_sync();
  const int L = LC > 0 ? LC : Lr, Rr = RC > 0 ? RC : Rrr;
  const int c = threadIdx.x & 127, half = threadIdx.x >> 7;
  const int S = L * Rr;
  float acc[NL][RM];
#pragma unroll
  for (int i = 0; i < NL; ++i)
#pragma unroll
    for (int n = 0; n < RM; ++n) acc[i][n] = 0.f;
  float accb = 0.f;
  for (int64_t f0 = (int64_t)blockIdx.x * kDwsChunk; f0 < E; f0 += (int64_t)gridDim.x * kDwsChunk) {
    const int nf = (int)min((int64_t)kDwsChunk, E - f0);
    // the chunk's d(table) values first: NL x 16 independent loads in flight per thread
    float v[kDwsChunk][NL];
#pragma unroll
    for (int j = 0; j < kDwsChunk; ++j)
#pragma unroll
      for (int i = 0; i < NL; ++i) {
        const int l = 2 * i + half;
        v[j][i] = (j < nf && l < L) ? __ldg(dP + ((f0 + j) * L + l) * 128 + c) : 0.f;
      }
    __syncthreads();
    for (int i = threadIdx.x; i < nf * S; i += kDwsThreads) stile[i] = __ldg(stab + f0 * S + i);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kDwsChunk; ++j) {
      if (j < nf) {
        const float* trow = stile + j * S;
#pragma unroll
        for (int i = 0; i < NL; ++i) {
          const int l = 2 * i + half;
          if (l < L) {
            if (l == 0) accb += v[j][i];
#pragma unroll
            for (int n = 0; n < RM; ++n)
              if (n < Rr) acc[i][n] = fmaf(v[j][i], trow[l * Rr + n], acc[i][n]);
          }
        }
      }
    }
  }
  float* out = partial + (int64_t)blockIdx.x * 128 * S + (int64_t)c * S;
#pragma unroll
  for (int i = 0; i < NL; ++i) {
    const int l = 2 * i + half;
    if (l < L) {
#pragma unroll
      for (int n = 0; n < RM; ++n)
        if (n < Rr) out[l * Rr + n] = acc[i][n];
    }
  }
  if (half == 0) colsum[(int64_t)blockIdx.x * 128 + c] = accb * (1.0f / (float)sqrt(0.07957747154594767));
}

}  // namespace x2
