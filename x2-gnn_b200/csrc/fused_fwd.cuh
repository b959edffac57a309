// fused_fwd.cuh -- fused forward of the T-scale part of SBFTransformerConv (sbftransformer_conv.py:
// 138-162 + PyG softmax / aggregate): ONE persistent kernel that
//   * streams edge_attr[T,A] and sbf[T,S] once from HBM,
//   * projects them on the tensor cores (tcgen05, 3xTF32): EA = lin_edge(edge_attr), Sg = lin_sbf(sbf),
//     with lin_edge's weights resident in TENSOR MEMORY (A operand) and lin_sbf's in shared memory,
//   * and runs the segmented attention directly on the TMEM accumulators: gathers K/V/Q rows, logits,
//     online segment softmax, sbf gate, aggregation, skip add -- deterministic and atomic-free.
// EA / Sg never go to HBM for the forward result (they are optionally written for the backward pass).
//
// Work decomposition: the target-sorted triplet list is cut into contiguous UNITS of ~512 triplets at
// segment boundaries, dealt round-robin to the 2 x gridDim.x (CTA, stream) slots.  A stream has its own
// accumulator columns in TMEM (64-triplet tiles: EA^T and Sg^T, 64 columns each) and its own 4 epilogue
// warps, so the sequential online-softmax state simply carries from tile to tile inside a unit, while the
// single MMA thread and the producer warps alternate between the CTA's two streams.  Small round-robin
// units keep all SMs inside the same few molecules at any time: a source row is re-gathered across its
// whole molecule, and with one long contiguous slice per SM the concurrent working set was the entire
// Q|K|V table (ncu: L2 hit rate 13 %, 2x the DRAM reads).
//
// Accumulators are transposed (lane = channel d, column = triplet): an epilogue warp owns 32 channels
// (= 32 / C heads), a head's C lanes reduce the logit with xor-shuffles -- the same arithmetic, in the
// same order along a segment, as k_attn_fwd.
#pragma once
#include "tc_gemm.cuh"

namespace x2 {
namespace tc {

constexpr int kFT = 64;                         // triplets per tile (UMMA N)
constexpr int kFChunkBytes = kFT * kChunkK * 4; // 8 KB: one 64 x 32 fp32 operand chunk
constexpr int kFStages = 6;                     // ring of [hi 8K | lo 8K] chunks (one tile of lookahead)
constexpr int kWsKC = 2;                        // lin_sbf K padded to 64 (S <= 64)
// Thread layout of the fused kernel (roles on warpgroup boundaries):
//   warps 0..6   producers (7 warps stream ~3.5 TB/s, what this kernel needs)
//   warp  7      MMA issuer + TMEM allocation
//   warps 8..15  epilogue  (stream 0: warps 8-11, stream 1: warps 12-15; TMEM lane quarter = warp & 3)
// 512 threads => 128 registers per thread: the epilogue keeps two blocks of gathered K/V/Q values in
// registers (with 800 threads / 80 registers they spilled to local memory; ncu showed STL traffic and
// long-scoreboard stalls; ptxas does not raise the cap after setmaxnreg.inc).
constexpr int kFProdWarps = 7;
constexpr int kFEpiWarp0 = 8;
constexpr int kFMmaWarp = 7;
constexpr int kFThreads = 16 * 32;
constexpr int kFProdThreads = kFProdWarps * 32;
constexpr int kFB = 8;                          // triplets per epilogue block

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

struct F1Params {
  const float* ea;  int64_t ld_ea; int A, KC_A;   // edge_attr [T, A], A <= 128
  const float* sbf; int64_t ld_s;  int S, KC_S;   // sbf [T, S], S <= 64
  const float* w_edge;                            // [128, A]
  const uint32_t* ws_img;                         // lin_sbf weights, K-major SW128 image: hi | lo, 2 x 2 x 16 KB
  const float* b_sbf;                             // [128]
  const float* qkvs; int ldq;                     // [E, 4*128]  Q | K | V | skip
  const int32_t* src; const int32_t* tgt; const int32_t* rowptr;      // [T], [T], [E+1]
  const int32_t* unit_e;                          // [nunits + 1] first target of each unit
  int nunits;
  int64_t E, T;
  int H, C; float scale; int fuse_skip;
  float *out, *attn, *lse;                        // [E,128], [E,128], [E,H]
  float *ea_out, *sg_out;                         // [T,128] each or NULL (saved for backward)
};

// lin_sbf weights -> K-major SW128 image (rows = output channel d, 2 chunks of 32 k), hi then lo
__global__ void k_make_ws_image(const float* __restrict__ W, int S, uint32_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // kc * 128 * 32 + d * 32 + kk
  if (idx >= kWsKC * 128 * kChunkK) return;
  const int kc = idx / (128 * kChunkK), rem = idx - kc * 128 * kChunkK;
  const int d = rem / kChunkK, kk = rem - d * kChunkK;
  const int k = kc * kChunkK + kk;
  const float w = k < S ? W[(int64_t)d * S + k] : 0.f;
  const uint32_t off = (uint32_t)kc * kChunkBytes + kmajor_off(d, kk >> 2) + (kk & 3) * 4;
  img[off >> 2] = hi_bits(w);
  img[(kWsKC * kChunkBytes + off) >> 2] = __float_as_uint(lo_part(w));
}

constexpr int kFUnit = 512;                     // nominal triplets per unit

// unit u starts at the first target whose first triplet is >= u * kFUnit
__global__ void k_unit_bounds(const int32_t* __restrict__ rowptr, int64_t E, int nunits,
                              int32_t* __restrict__ unit_e) {
  const int u = blockIdx.x * blockDim.x + threadIdx.x;
  if (u > nunits) return;
  if (u == nunits) { unit_e[u] = (int32_t)E; return; }
  const int64_t want = (int64_t)u * kFUnit;
  int64_t lo = 0, hi = E;                       // first e with rowptr[e] >= want
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (rowptr[mid] < want) lo = mid + 1; else hi = mid;
  }
  unit_e[u] = (int32_t)(u == 0 ? 0 : lo);
}

template <int C>
__global__ void __launch_bounds__(kFThreads, 1) k_fused_fwd(const F1Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sW = smem;                                               // Ws image: hi 32 KB | lo 32 KB
  uint8_t* sR = smem + 2 * kWsKC * kChunkBytes;                     // ring: kFStages x [hi 8K | lo 8K]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sR + (size_t)kFStages * 2 * kFChunkBytes);
  uint64_t* full = bars;                        // [kFStages]
  uint64_t* empty = bars + kFStages;            // [kFStages]
  uint64_t* tfull = bars + 2 * kFStages;        // [2] per stream
  uint64_t* tempty = bars + 2 * kFStages + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kFStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int D = 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kFStages; ++i) {
      mbar_init(&full[i], kFProdThreads);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == kFMmaWarp) tmem_alloc(tmem_slot, 512);
  {  // lin_sbf weight image -> smem
    const uint4* src = reinterpret_cast<const uint4*>(p.ws_img);
    const uint32_t dst = smem_u32(sW);
    for (int i = threadIdx.x; i < 2 * kWsKC * kChunkBytes / 16; i += kFThreads) sts128(dst + i * 16, __ldg(src + i));
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // lin_edge weights -> tensor memory (lane = channel d, column = k): hi at 256.., lo at 384..
  if (warp >= kFEpiWarp0 && warp < kFEpiWarp0 + 4) {
    const int q = warp & 3, d = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int k0 = 0; k0 < p.KC_A * kChunkK; k0 += 32) {
      float w[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) w[j] = (k0 + j) < p.A ? __ldg(p.w_edge + (int64_t)d * p.A + k0 + j) : 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          hi[j] = hi_bits(w[g * 8 + j]);
          lo[j] = __float_as_uint(lo_part(w[g * 8 + j]));
        }
        tmem_st8(trow + kTmemWHi + k0 + g * 8, hi);
        tmem_st8(trow + kTmemWLo + k0 + g * 8, lo);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // ---- rounds: in round r, stream `set` of this CTA processes unit r * nslots + 2 * blockIdx.x + set
  const int nslots = 2 * gridDim.x;
  const int KC_A = p.KC_A, KC_S = p.KC_S;
  const int32_t* __restrict__ unit_e = p.unit_e;
  const int32_t* __restrict__ rowptr = p.rowptr;
  auto unit_range = [&](int u, int64_t& tb, int64_t& te) {      // triplet range of unit u (empty if u >= nunits)
    if (u < p.nunits) { tb = rowptr[unit_e[u]]; te = rowptr[unit_e[u + 1]]; }
    else { tb = 0; te = 0; }
  };

  if (warp == kFMmaWarp) {
    // =============================== MMA issuer ===============================
    // whole warp, uniform control flow; one elected lane issues (see umma_tf32_ts_w in tc_gemm.cuh)
    {
      const uint32_t leader = elect_one();
      const uint32_t idesc = make_idesc(kFT, 0, 0);          // M = 128 channels, N = 64 triplets
      const uint32_t sR_u = smem_u32(sR), sW_u = smem_u32(sW);
      uint32_t st = 0, ph = 0;
      uint32_t cnt[2] = {0, 0};                              // tiles issued per stream (TMEM handshake parity)
      for (int u0 = 2 * blockIdx.x; u0 < p.nunits; u0 += nslots) {
        int64_t tb[2], te[2];
        unit_range(u0, tb[0], te[0]);
        unit_range(u0 + 1, tb[1], te[1]);
        const int64_t nt0 = (te[0] - tb[0] + kFT - 1) / kFT, nt1 = (te[1] - tb[1] + kFT - 1) / kFT;
        const int64_t nmax = nt0 > nt1 ? nt0 : nt1;
        for (int64_t i = 0; i < nmax; ++i) {
          for (int set = 0; set < 2; ++set) {
            if (i >= (set ? nt1 : nt0)) continue;
            mbar_wait(&tempty[set], (cnt[set] & 1) ^ 1);
            ++cnt[set];
            tc_fence_after();
            const uint32_t t_ea = tmem_base + set * 128, t_sg = t_ea + kFT;
            for (int kc = 0; kc < KC_A; ++kc) {               // EA^T += W_e . ea^T   (A operand in TMEM)
              mbar_wait(&full[st], ph);
              tc_fence_after();
              const int ksteps = (min(kChunkK, p.A - kc * kChunkK) + 7) >> 3;
              uint64_t dh = make_desc(sR_u + st * 2 * kFChunkBytes, 16, 1024);
              uint64_t dl = make_desc(sR_u + st * 2 * kFChunkBytes + kFChunkBytes, 16, 1024);
              uint32_t w_hi = tmem_base + kTmemWHi + kc * kChunkK, w_lo = tmem_base + kTmemWLo + kc * kChunkK;
              for (int ks = 0; ks < ksteps; ++ks, dh += 2, dl += 2, w_hi += 8, w_lo += 8) {
                umma_tf32_ts_w(leader, t_ea, w_hi, dh, idesc, (kc | ks) != 0);
                umma_tf32_ts_w(leader, t_ea, w_lo, dh, idesc, 1);
                umma_tf32_ts_w(leader, t_ea, w_hi, dl, idesc, 1);
              }
              umma_commit_w(leader, &empty[st]);
              if (++st == kFStages) { st = 0; ph ^= 1; }
            }
            for (int kc = 0; kc < KC_S; ++kc) {               // Sg^T += W_s . sbf^T   (A operand in smem)
              mbar_wait(&full[st], ph);
              tc_fence_after();
              const int ksteps = (min(kChunkK, p.S - kc * kChunkK) + 7) >> 3;
              uint64_t dh = make_desc(sR_u + st * 2 * kFChunkBytes, 16, 1024);
              uint64_t dl = make_desc(sR_u + st * 2 * kFChunkBytes + kFChunkBytes, 16, 1024);
              uint64_t wh = make_desc(sW_u + kc * kChunkBytes, 16, 1024);
              uint64_t wl = make_desc(sW_u + kWsKC * kChunkBytes + kc * kChunkBytes, 16, 1024);
              for (int ks = 0; ks < ksteps; ++ks, dh += 2, dl += 2, wh += 2, wl += 2) {
                umma_tf32_w(leader, t_sg, wh, dh, idesc, (kc | ks) != 0);
                umma_tf32_w(leader, t_sg, wl, dh, idesc, 1);
                umma_tf32_w(leader, t_sg, wh, dl, idesc, 1);
              }
              umma_commit_w(leader, &empty[st]);
              if (++st == kFStages) { st = 0; ph ^= 1; }
            }
            umma_commit_w(leader, &tfull[set]);
          }
        }
      }
    }
  } else if (warp < kFProdWarps) {
    // =============================== producers ===============================
    // 224 threads; a 64 x 32 chunk is 512 float4: thread pt moves elements f = pt + 224 i (i < 3, f < 512),
    // row f >> 3, 16-byte column f & 7.  A whole tile (KC_A + KC_S chunks) of loads is issued, then its
    // chunks are split and stored; the 6-stage ring lets the MMA / epilogue work one tile behind.
    const int pt = threadIdx.x;
    const bool vecA = ((p.ld_ea & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.ea) & 15) == 0);
    const uint32_t sR_u = smem_u32(sR);
    constexpr int MAXC = 4 + kWsKC;
    constexpr int NP = 3;
    const float* __restrict__ ea_p = p.ea;
    const float* __restrict__ sbf_p = p.sbf;
    int rr[NP], cc[NP];
    uint32_t soff[NP];
    bool act[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      const int f = pt + kFProdThreads * i;
      act[i] = f < kFT * 8;
      rr[i] = (f >> 3) & (kFT - 1);
      cc[i] = f & 7;
      soff[i] = kmajor_off(rr[i], cc[i]);
    }
    uint32_t st = 0, ph = 0;
    for (int u0 = 2 * blockIdx.x; u0 < p.nunits; u0 += nslots) {
      int64_t tb[2], te[2];
      unit_range(u0, tb[0], te[0]);
      unit_range(u0 + 1, tb[1], te[1]);
      const int64_t nt0 = (te[0] - tb[0] + kFT - 1) / kFT, nt1 = (te[1] - tb[1] + kFT - 1) / kFT;
      const int64_t nmax = nt0 > nt1 ? nt0 : nt1;
      for (int64_t i = 0; i < nmax; ++i) {
        for (int set = 0; set < 2; ++set) {
          if (i >= (set ? nt1 : nt0)) continue;
          const int64_t t0 = tb[set] + i * kFT;
          const int64_t tend = te[set];
          float4 v[MAXC][NP];
#pragma unroll
          for (int u = 0; u < NP; ++u) {
            const int64_t t = t0 + rr[u];
            const bool ok = act[u] && t < tend;
#pragma unroll
            for (int kc = 0; kc < 4; ++kc) {
              v[kc][u] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (kc < KC_A && ok) {
                const int k = kc * kChunkK + cc[u] * 4;
                const float* src = ea_p + t * p.ld_ea + k;
                if (vecA && k + 3 < p.A) v[kc][u] = __ldcs(reinterpret_cast<const float4*>(src));
                else {
                  if (k < p.A) v[kc][u].x = __ldcs(src);
                  if (k + 1 < p.A) v[kc][u].y = __ldcs(src + 1);
                  if (k + 2 < p.A) v[kc][u].z = __ldcs(src + 2);
                  if (k + 3 < p.A) v[kc][u].w = __ldcs(src + 3);
                }
              }
            }
#pragma unroll
            for (int kc = 0; kc < kWsKC; ++kc) {
              float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
              if (kc < KC_S && ok) {
                const int k = kc * kChunkK + cc[u] * 4;
                const float* src = sbf_p + t * p.ld_s + k;
                if (k < p.S) x.x = __ldcs(src);
                if (k + 1 < p.S) x.y = __ldcs(src + 1);
                if (k + 2 < p.S) x.z = __ldcs(src + 2);
                if (k + 3 < p.S) x.w = __ldcs(src + 3);
              }
              v[4 + kc][u] = x;
            }
          }
#pragma unroll
          for (int kc = 0; kc < MAXC; ++kc) {
            if ((kc < 4 && kc < KC_A) || (kc >= 4 && kc - 4 < KC_S)) {
              mbar_wait(&empty[st], ph ^ 1);
              const uint32_t base = sR_u + st * 2 * kFChunkBytes;
#pragma unroll
              for (int u = 0; u < NP; ++u) {
                if (act[u]) {
                  uint4 hi, lo;
                  split4(v[kc][u], hi, lo);
                  sts128(base + soff[u], hi);
                  sts128(base + kFChunkBytes + soff[u], lo);
                }
              }
              fence_proxy_async();
              mbar_arrive(&full[st]);
              if (++st == kFStages) { st = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else {
    // =============================== epilogue: segmented attention on the accumulators ===============
    // lane = channel d; the C lanes of a head reduce the logit with xor-shuffles.  Logits are kept in
    // the log2 domain (scale * log2 e folded in) so the softmax uses ex2 directly.
    const int ew = warp - kFEpiWarp0;                // 0..7
    const int set = ew >> 2;                         // stream handled by this warp
    const int q = warp & 3;                          // TMEM lane quarter
    const int d = q * 32 + lane;                     // channel of this thread
    const int head = d / C;
    const bool leader = (d % C) == 0;
    const float bs = __ldg(p.b_sbf + d);
    const float scale2 = p.scale * 1.4426950408889634f;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + set * 128;
    const float* __restrict__ qkvs = p.qkvs;
    const int32_t* __restrict__ srcp = p.src;
    const int32_t* __restrict__ tgtp = p.tgt;
    float* __restrict__ outp = p.out;
    float* __restrict__ attnp = p.attn;
    float* __restrict__ lsep = p.lse;
    float* __restrict__ eao = p.ea_out;
    float* __restrict__ sgo = p.sg_out;
    const int ldq = p.ldq;
    const bool save = eao != nullptr;
    uint32_t tcnt = 0;                                // tiles consumed by this stream (TMEM handshake parity)

    for (int u = 2 * blockIdx.x + set; u < p.nunits; u += nslots) {
      int32_t cur_e = unit_e[u];
      const int32_t e_hi = unit_e[u + 1];
      const int64_t t_beg = rowptr[cur_e], t_end = rowptr[e_hi];
      float m = -INFINITY, z = 0.f, acc = 0.f;       // m in log2 units
      auto flush = [&]() {                            // finish target cur_e and move to the next one
        const float inv = 1.0f / (z + 1e-16f);
        float o = acc * inv;
        attnp[(int64_t)cur_e * D + d] = o;
        if (p.fuse_skip) o += __ldg(qkvs + (int64_t)cur_e * ldq + 3 * D + d);
        outp[(int64_t)cur_e * D + d] = o;
        if (leader) lsep[(int64_t)cur_e * p.H + head] = z > 0.f ? (m + log2f(z)) * 0.6931471805599453f : 0.f;
        ++cur_e;
        m = -INFINITY; z = 0.f; acc = 0.f;
      };

      const int64_t nblk = (t_end - t_beg + kFB - 1) / kFB;
      float ck[kFB], cv[kFB], cq[kFB];               // gathered values of the current block
      float nk[kFB], nv[kFB], nq[kFB];               // ... and of the next one (in flight)
      int ce_l = 0, ne_l = 0;                        // lane j holds the target of triplet j of the block

#define X2_GATHER(G, K_, V_, Q_, E_L)                                                   \
      {                                                                                 \
        const int64_t tb_ = t_beg + (G) * kFB;                                          \
        int s_l_ = 0;                                                                   \
        E_L = 0;                                                                        \
        if (lane < kFB && tb_ + lane < t_end) {                                         \
          s_l_ = __ldg(srcp + tb_ + lane);                                              \
          E_L = __ldg(tgtp + tb_ + lane);                                               \
        }                                                                               \
        _Pragma("unroll") for (int j = 0; j < kFB; ++j) {                               \
          const int sj_ = __shfl_sync(0xffffffffu, s_l_, j);                            \
          const int ej_ = __shfl_sync(0xffffffffu, E_L, j);                             \
          K_[j] = __ldg(qkvs + (int64_t)sj_ * ldq + D + d);                             \
          V_[j] = __ldg(qkvs + (int64_t)sj_ * ldq + 2 * D + d);                         \
          Q_[j] = __ldg(qkvs + (int64_t)ej_ * ldq + d);                                 \
        }                                                                               \
      }

      if (nblk > 0) X2_GATHER(0, ck, cv, cq, ce_l)
      for (int64_t g = 0; g < nblk; ++g) {
        if (g + 1 < nblk) X2_GATHER(g + 1, nk, nv, nq, ne_l)
        const int64_t tb = t_beg + g * kFB;
        const int cb = (int)((g * kFB) & (kFT - 1));  // column of this block inside its tile
        if (cb == 0) {                                // first block of a tile: wait for its accumulators
          mbar_wait(&tfull[set], tcnt & 1);
          ++tcnt;
          tc_fence_after();
        }
        float ea[kFB], sg[kFB], a[kFB];
        tmem_ld8(trow + cb, ea);
        tmem_ld8(trow + kFT + cb, sg);
        // phase A: independent per triplet -- logits
#pragma unroll
        for (int j = 0; j < kFB; ++j) {
          sg[j] += bs;
          float dot = cq[j] * (ck[j] + ea[j]);
#pragma unroll
          for (int off = 1; off < C; off <<= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
          a[j] = dot * scale2;
          cv[j] = (cv[j] + ea[j]) * sg[j];            // gated value
        }
        if (save) {
#pragma unroll
          for (int j = 0; j < kFB; ++j) {
            if (tb + j < t_end) {
              __stcs(eao + (tb + j) * D + d, ea[j]);  // one-touch data: evict-first
              __stcs(sgo + (tb + j) * D + d, sg[j]);
            }
          }
        }
        // phase B: sequential online softmax along the segment
#pragma unroll
        for (int j = 0; j < kFB; ++j) {
          if (tb + j < t_end) {
            const int ej = __shfl_sync(0xffffffffu, ce_l, j);
            while (cur_e < ej) flush();               // also steps over empty segments
            const float mn = fmaxf(m, a[j]);
            const float corr = exp2f(m - mn);
            const float pr = exp2f(a[j] - mn);
            z = fmaf(z, corr, pr);
            acc = fmaf(acc, corr, pr * cv[j]);
            m = mn;
          }
        }
        if (cb == kFT - kFB || g == nblk - 1) {       // last block of the tile: release the accumulators
          tc_fence_before();
          mbar_arrive(&tempty[set]);
        }
#pragma unroll
        for (int j = 0; j < kFB; ++j) { ck[j] = nk[j]; cv[j] = nv[j]; cq[j] = nq[j]; }
        ce_l = ne_l;
      }
#undef X2_GATHER
      while (cur_e < e_hi) flush();                   // last segment + trailing empty targets
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kFMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

static inline bool fused_fwd_supported(int D, int H, int C, int A, int S) {
  return D == 128 && A >= 1 && A <= 128 && S >= 1 && S <= kWsKC * kChunkK && (C == 4 || C == 8 || C == 16 || C == 32) &&
         H * C == D;
}
static inline size_t fused_fwd_workspace_bytes(int64_t T) {
  const size_t nunits = (size_t)((T > 0 ? T : 0) / kFUnit + 2);
  return align_up((size_t)2 * kWsKC * kChunkBytes, 256) + align_up((nunits + 1) * 4, 256) + 256;
}

// ws: [Ws image | stream_e]
static int fused_fwd(F1Params p, const float* w_sbf, void* ws, cudaStream_t st) {
  uint32_t* img = static_cast<uint32_t*>(ws);
  int32_t* unit_e = reinterpret_cast<int32_t*>(static_cast<char*>(ws) + align_up((size_t)2 * kWsKC * kChunkBytes, 256));
  const int nunits = (int)((p.T + kFUnit - 1) / kFUnit);
  const int grid = nunits < 2 * kNumSM ? (nunits + 1) / 2 : kNumSM;
  k_make_ws_image<<<(kWsKC * 128 * kChunkK + 255) / 256, 256, 0, st>>>(w_sbf, p.S, img);
  X2_LAUNCH_OK();
  k_unit_bounds<<<(nunits + 1 + 255) / 256, 256, 0, st>>>(p.rowptr, p.E, nunits, unit_e);
  X2_LAUNCH_OK();
  p.ws_img = img;
  p.unit_e = unit_e;
  p.nunits = nunits;
  p.KC_A = (p.A + kChunkK - 1) / kChunkK;
  p.KC_S = (p.S + kChunkK - 1) / kChunkK;
  const size_t smem = 1024 + (size_t)2 * kWsKC * kChunkBytes + (size_t)kFStages * 2 * kFChunkBytes + 256;
  static bool attr_set = false;
  if (!attr_set) {
    X2_CUDA_OK(cudaFuncSetAttribute(k_fused_fwd<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    X2_CUDA_OK(cudaFuncSetAttribute(k_fused_fwd<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    X2_CUDA_OK(cudaFuncSetAttribute(k_fused_fwd<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    X2_CUDA_OK(cudaFuncSetAttribute(k_fused_fwd<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_set = true;
  }
  switch (p.C) {
    case 4: k_fused_fwd<4><<<grid, kFThreads, smem, st>>>(p); break;
    case 8: k_fused_fwd<8><<<grid, kFThreads, smem, st>>>(p); break;
    case 16: k_fused_fwd<16><<<grid, kFThreads, smem, st>>>(p); break;
    default: k_fused_fwd<32><<<grid, kFThreads, smem, st>>>(p); break;
  }
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // namespace tc
}  // namespace x2
