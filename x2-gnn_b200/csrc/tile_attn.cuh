// tile_attn.cuh -- the T-scale part of the SBFTransformerConv forward (sbftransformer_conv.py:138-162 +
// PyG softmax / sum aggregation) as ONE persistent tcgen05 kernel on the default [T, A] path:
//
//   edge_attr[T,128], sbf[T,S]  --cp.async-->  swizzled smem ring (3xTF32 hi / lo operands)
//        --tcgen05.mma (weights resident in TENSOR MEMORY)-->  EA^T, Sg^T chunks in TMEM (32 triplets each)
//        --4 transposer warps (tcgen05.ld, conflict-free st.shared)-->  [triplet][channel] tile in smem
//        --15 consumer warps-->  gathers of K / V / Q rows, logits, online segment softmax, sbf gate,
//                                aggregation, skip add  -->  out[E,128], lse[E,H]
//
// EA = lin_edge(edge_attr) and Sg = lin_sbf(sbf) are never read back from HBM by the forward (the unfused pair
// k_tc_gemm x 2 + k_attn_fwd wrote 1 024 B and re-read 1 024 B per triplet); they are optionally stored for the
// backward (ea_out / sg_out).
//
// Work decomposition.  The target-sorted triplet list is cut into TILES of whole segments with at most kTaRows
// rows (x2_tiles_build: window cut, see graph.cu), dealt round-robin to the CTAs (one per SM), so that all SMs
// work inside the same few molecules at any time (the K / V gathers then hit L2).  A tile is streamed in
// 32-row chunks through a 2-stage ring; the accumulators of a chunk (EA 32 + Sg 32 columns) are double-buffered
// in TMEM.  The consumers do not own segments but ITEMS: every segment is cut, relative to its own start, into
// runs of at most 8 rows (x2_tiles_build lists them in row order; a tile has ~7 segments of a QM9 molecule or
// ~2 of a dense 500-atom ball -- far too few for 15 warps, but ~15-20 items).  A warp walks its items
// (item i of the tile -> warp i mod 15), keeps an online-softmax state (m, z, acc) and drops one PARTIAL state
// per item into the tile row the item starts at; after a consumer-wide barrier, warp-per-target merges the
// partial states of each segment in row order (flash-decoding style), normalises, adds the skip projection
// and writes out / attn / lse.  The grouping of a segment's rows depends on the segment alone -- not on where
// it lies in a tile or in the batch -- so results are deterministic AND bitwise independent of the rest of the
// batch.  An item starts as soon as the chunk that holds its last row has been transposed, so the attention
// arithmetic of a tile overlaps the streaming of its later chunks, and the front end runs two chunks into
// the next tile while the last items and the merge finish.
//
// Budget per CTA: ring 2 x 48 KB + EA / Sg tile 2 x 60 KB = 216 KB of shared memory; tensor memory: accumulators
// 128 + W_sbf hi / lo 128 + W_edge hi / lo 256 = 512 columns; 896 threads at 72 registers.
#pragma once
#include "tc_gemm.cuh"

namespace x2 {
namespace tc {

constexpr int kTaRows = X2_TILE_ROWS;                         // rows of the smem tile (max triplets per tile)
constexpr int kTaChunk = 32;                          // triplets per chunk (UMMA N)
constexpr int kTaStages = 2;
constexpr int kTaXHalf = kTaChunk * 128 * 4;          // 16 KB: [32 rows][128 k] fp32, 4 K-blocks of 4 KB
constexpr int kTaSHalf = kTaChunk * 64 * 4;           // 8 KB: sbf, K padded to 64 (2 K-blocks)
constexpr int kTaStage = 2 * kTaXHalf + 2 * kTaSHalf; // X hi | X lo | S hi | S lo
constexpr int kTaTileBytes = kTaRows * 512;
constexpr int kTaTrWarp0 = 1;                         // warp 0: MMA; warps 1..4: transposers
constexpr int kTaProdWarp0 = 5, kTaProdWarps = 8;     // warps 5..12
constexpr int kTaProdThreads = kTaProdWarps * 32;
constexpr int kTaConsWarp0 = 13, kTaConsWarps = 15;   // warps 13..27: one per 8-row block of a full tile
constexpr int kTaThreads = (kTaConsWarp0 + kTaConsWarps) * 32;   // 896 => 72 registers per thread
constexpr int kTaBlk = X2_TILE_ITEM_ROWS;             // rows per consumer item
constexpr int kTaRing = 3;                            // consumer register ring of gathered K / V rows
// tensor-memory columns
constexpr int kTaWsHi = 128, kTaWsLo = 192, kTaWeHi = 256, kTaWeLo = 384;
constexpr size_t kTaSmem = 1024 + (size_t)kTaStages * kTaStage + 2 * (size_t)kTaTileBytes + 256;

enum { kTaEaTriplet = 1, kTaEaSegment = 2, kTaEaNone = 0 };

struct TaParams {
  const float* ea;         // edge_attr [T, 128] (kTaEaTriplet) | lin_edge(table) [M, 128] (kTaEaSegment)
  const int32_t* ea_index; // [E] table row of every target (kTaEaSegment)
  const float* sbf; int S; // [T, S], S even, <= 64
  const float* w_edge;     // [128, 128]
  const float* w_sbf;      // [128, S]
  const float* b_sbf;      // [128]
  const float* qkvs; int ldq;                        // [E, 4*128]  Q | K | V | skip
  const int32_t* src; const int32_t* tgt; const int32_t* rowptr;
  const int32_t* tile;     // [ntiles + 1][4]: first target, first triplet, first item of every tile (x2_tiles_build)
  const int32_t* items;    // [nitems][2]: first triplet, (target << 4) | (rows - 1) of every item
  int ntiles;
  int H, C; float scale; int fuse_skip;
  float *out, *attn, *lse;                           // [E,128], [E,128], [E,H]
  float *ea_out, *sg_out;                            // [T,128] each or NULL (saved for the backward)
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Tile cursor of one role: tiles blockIdx.x, + gridDim.x, ...; the next tile's bounds are loaded one tile ahead.
struct TaTile {
  int e0, t0, i0, e1, t1, i1;
};

template <int LPH, int EA>
__global__ void __launch_bounds__(kTaThreads, 1) k_tile_fwd(const TaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sRing = smem;
  uint8_t* sEa = smem + (size_t)kTaStages * kTaStage;       // [kTaRows][128] fp32
  uint8_t* sSg = sEa + kTaTileBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSg + kTaTileBytes);
  uint64_t* full = bars;              // [2]  producers -> MMA
  uint64_t* empty = bars + 2;         // [2]  MMA -> producers
  uint64_t* accfull = bars + 4;       // [2]  MMA -> transposers
  uint64_t* accempty = bars + 6;      // [2]  transposers -> MMA
  uint64_t* chunkfull = bars + 8;     // [4]  transposers -> consumers (per chunk of the tile)
  uint64_t* tileempty = bars + 12;    // [1]  consumers -> transposers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int D = 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], kTaProdThreads);
      mbar_init(&empty[i], 1);
      mbar_init(&accfull[i], 1);
      mbar_init(&accempty[i], 128);
    }
    for (int i = 0; i < 4; ++i) mbar_init(&chunkfull[i], 128);
    mbar_init(tileempty, kTaConsWarps);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  {  // the padding of the sbf operand blocks (k >= S) is never written by the copies: zero it once
    const uint32_t ring_u = smem_u32(sRing);
    for (int i = threadIdx.x; i < kTaStages * 2 * kTaSHalf / 16; i += kTaThreads) {
      const int stg = i / (2 * kTaSHalf / 16), w = i - stg * (2 * kTaSHalf / 16);
      sts128(ring_u + stg * kTaStage + 2 * kTaXHalf + w * 16, make_uint4(0u, 0u, 0u, 0u));
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();        // global memory from here on

  const int32_t* __restrict__ tinfo = p.tile;
  auto load_tile = [&](int k, TaTile& t) {
    const int4 a = __ldg(reinterpret_cast<const int4*>(tinfo) + k);
    const int4 b = __ldg(reinterpret_cast<const int4*>(tinfo) + k + 1);
    t.e0 = a.x; t.t0 = a.y; t.i0 = a.z; t.e1 = b.x; t.t1 = b.y; t.i1 = b.z;
  };
  const int KS_S0 = (min(p.S, 32) + 7) >> 3;             // k-steps of the two sbf K-blocks
  const int KS_S1 = p.S > 32 ? (p.S - 32 + 7) >> 3 : 0;

  if (warp == 0) {
    // =============================== MMA issuer (whole warp, one elected lane issues) ===============
    const uint32_t leader = elect_one();
    const uint32_t idesc = make_idesc(kTaChunk, 0, 0);     // M = 128 channels, N = 32 triplets
    const uint32_t ring_u = smem_u32(sRing);
    asm volatile("bar.sync 2, 160;" ::: "memory");         // weights are in tensor memory
    tc_fence_after();
    uint32_t st = 0, ph = 0, cc = 0;
    TaTile cur, nxt;
    if ((int)blockIdx.x < p.ntiles) load_tile(blockIdx.x, cur);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
      if (tile + (int)gridDim.x < p.ntiles) load_tile(tile + gridDim.x, nxt);
      const int nch = (cur.t1 - cur.t0 + kTaChunk - 1) / kTaChunk;
      for (int c = 0; c < nch; ++c, ++cc) {
        const uint32_t buf = cc & 1;
        mbar_wait(&accempty[buf], ((cc >> 1) & 1) ^ 1);
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t base = ring_u + st * kTaStage;
        const uint32_t t_ea = tmem_base + buf * 64, t_sg = t_ea + 32;
        if constexpr (EA == kTaEaTriplet) {
#pragma unroll 1
          for (int kc = 0; kc < 4; ++kc) {                 // EA^T = W_e . X^T
            uint64_t dh = make_desc(base + kc * 4096, 16, 1024);
            uint64_t dl = make_desc(base + kTaXHalf + kc * 4096, 16, 1024);
            uint32_t w_hi = tmem_base + kTaWeHi + kc * 32, w_lo = tmem_base + kTaWeLo + kc * 32;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks, dh += 2, dl += 2, w_hi += 8, w_lo += 8) {
              umma_tf32_ts_w(leader, t_ea, w_hi, dh, idesc, (kc | ks) != 0);
              umma_tf32_ts_w(leader, t_ea, w_lo, dh, idesc, 1);
              umma_tf32_ts_w(leader, t_ea, w_hi, dl, idesc, 1);
            }
          }
        }
#pragma unroll 1
        for (int kc = 0; kc < 2; ++kc) {                   // Sg^T = W_s . sbf^T
          const int ksteps = kc == 0 ? KS_S0 : KS_S1;
          uint64_t dh = make_desc(base + 2 * kTaXHalf + kc * 4096, 16, 1024);
          uint64_t dl = make_desc(base + 2 * kTaXHalf + kTaSHalf + kc * 4096, 16, 1024);
          uint32_t w_hi = tmem_base + kTaWsHi + kc * 32, w_lo = tmem_base + kTaWsLo + kc * 32;
          for (int ks = 0; ks < ksteps; ++ks, dh += 2, dl += 2, w_hi += 8, w_lo += 8) {
            umma_tf32_ts_w(leader, t_sg, w_hi, dh, idesc, (kc | ks) != 0);
            umma_tf32_ts_w(leader, t_sg, w_lo, dh, idesc, 1);
            umma_tf32_ts_w(leader, t_sg, w_hi, dl, idesc, 1);
          }
        }
        umma_commit_w(leader, &empty[st]);
        umma_commit_w(leader, &accfull[buf]);
        if (++st == kTaStages) { st = 0; ph ^= 1; }
      }
      cur = nxt;
    }
    __syncwarp();
  } else if (warp < kTaProdWarp0) {
    // =============================== transposers ==================================================
    const int q = warp & 3;                                // TMEM lane quarter this warp may access
    const int ch = q * 32 + lane;                          // channel of this thread
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    // ---- weights -> tensor memory (lane = output channel, column = k), hi / lo split in software
    auto load_w = [&](const float* __restrict__ W, int K, int KP, uint32_t chi, uint32_t clo) {
      const float* __restrict__ wrow = W + (int64_t)ch * K;
      const bool vec = (K & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
      for (int k0 = 0; k0 < KP; k0 += 32) {
        float w[32];
        if (vec) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 v = (k0 + j < K) ? __ldg(reinterpret_cast<const float4*>(wrow + k0 + j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) w[j] = (k0 + j < K) ? __ldg(wrow + k0 + j) : 0.f;
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            hi[j] = hi_bits(w[g * 8 + j]);
            lo[j] = __float_as_uint(lo_part(w[g * 8 + j]));
          }
          tmem_st8(trow + chi + k0 + g * 8, hi);
          tmem_st8(trow + clo + k0 + g * 8, lo);
        }
      }
    };
    if constexpr (EA == kTaEaTriplet) load_w(p.w_edge, 128, 128, kTaWeHi, kTaWeLo);
    load_w(p.w_sbf, p.S, 64, kTaWsHi, kTaWsLo);
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    asm volatile("bar.arrive 2, 160;" ::: "memory");
    const float bs = __ldg(p.b_sbf + ch);
    const uint32_t ea_u = smem_u32(sEa) + ch * 4, sg_u = smem_u32(sSg) + ch * 4;
    float* __restrict__ eao = p.ea_out;
    float* __restrict__ sgo = p.sg_out;
    uint32_t cc = 0, tcount = 0;
    TaTile cur, nxt;
    if ((int)blockIdx.x < p.ntiles) load_tile(blockIdx.x, cur);
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
      if (tile + (int)gridDim.x < p.ntiles) load_tile(tile + gridDim.x, nxt);
      const int rows = cur.t1 - cur.t0;
      const int nch = (rows + kTaChunk - 1) / kTaChunk;
      mbar_wait(tileempty, (tcount & 1) ^ 1);              // consumers are done with the previous tile
      for (int c = 0; c < nch; ++c, ++cc) {
        const uint32_t buf = cc & 1;
        mbar_wait(&accfull[buf], (cc >> 1) & 1);
        tc_fence_after();
        const int valid = min(kTaChunk, rows - c * kTaChunk);
        const uint32_t taddr = trow + buf * 64;
        float v[32];
        if constexpr (EA == kTaEaTriplet) {
          tmem_ld32(taddr, v);
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < valid) sts32f(ea_u + (uint32_t)(c * kTaChunk + j) * 512, v[j]);
          if (eao) {
            float* dst = eao + ((int64_t)cur.t0 + c * kTaChunk) * D + ch;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < valid) dst[(int64_t)j * D] = v[j];
          }
        }
        tmem_ld32(taddr + 32, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          v[j] += bs;
          if (j < valid) sts32f(sg_u + (uint32_t)(c * kTaChunk + j) * 512, v[j]);
        }
        if (sgo) {
          float* dst = sgo + ((int64_t)cur.t0 + c * kTaChunk) * D + ch;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < valid) dst[(int64_t)j * D] = v[j];
        }
        tc_fence_before();
        mbar_arrive(&accempty[buf]);
        mbar_arrive(&chunkfull[c]);
      }
      for (int c = nch; c < 4; ++c) mbar_arrive(&chunkfull[c]);   // keep the phases of all four in step
      cur = nxt;
    }
  } else if (warp < kTaConsWarp0) {
    // =============================== producers ====================================================
    // X chunk = 32 rows x 32 16-byte pieces: thread (r0 = pt / 32, c16 = pt % 32) moves column piece c16 of
    // rows r0 + 8 i.  sbf chunk = 32 rows x S/2 8-byte pieces: thread moves pieces pt + 256 j.
    const int pt = threadIdx.x - kTaProdWarp0 * 32;
    const int r0 = pt >> 5, c16 = pt & 31;
    const uint32_t xoff = (uint32_t)((c16 >> 3) * 4096 + r0 * 128 + (((c16 & 7) ^ r0) << 4));
    const int S = p.S, S2 = S >> 1;
    uint32_t soff[3];
    int srow[3], scol[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int f = pt + kTaProdThreads * j;
      const int row = f / S2, col = (f - row * S2) * 2;
      srow[j] = f < kTaChunk * S2 ? row : 1 << 20;         // inactive pieces never pass the row test
      scol[j] = col;
      soff[j] = (uint32_t)(2 * kTaXHalf + (col >> 5) * 4096) + kmajor_off(row & 31, (col & 31) >> 2) + (uint32_t)(col & 3) * 4;
    }
    const uint32_t ring_u = smem_u32(sRing);
    const float* __restrict__ xg = p.ea;
    const float* __restrict__ sg_ = p.sbf;
    uint32_t ist = 0, iph = 0, cst = 0;
    int it_tile = blockIdx.x, it_c = 0;
    TaTile icur;
    if (it_tile < p.ntiles) load_tile(it_tile, icur);
    auto issue = [&]() -> bool {                            // copies of the next chunk; false when none is left
      while (it_tile < p.ntiles && it_c * kTaChunk >= icur.t1 - icur.t0) {
        it_tile += gridDim.x;
        it_c = 0;
        if (it_tile < p.ntiles) load_tile(it_tile, icur);
      }
      if (it_tile >= p.ntiles) return false;
      const int valid = min(kTaChunk, icur.t1 - icur.t0 - it_c * kTaChunk);
      const int64_t tb = (int64_t)icur.t0 + it_c * kTaChunk;
      ++it_c;
      mbar_wait(&empty[ist], iph ^ 1);                     // the MMAs that read this stage have retired
      const uint32_t base = ring_u + ist * kTaStage;
      if constexpr (EA == kTaEaTriplet) {
        const float* src = xg + (tb + r0) * D + c16 * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (r0 + 8 * i < valid) cp_async<16>(base + xoff + i * 1024, src + (int64_t)i * 8 * D);
      }
#pragma unroll
      for (int j = 0; j < 3; ++j)
        if (srow[j] < valid) cp_async<8>(base + soff[j], sg_ + (tb + srow[j]) * S + scol[j]);
      if (++ist == kTaStages) { ist = 0; iph ^= 1; }
      return true;
    };
    auto consume = [&]() {                                  // own pieces have landed: derive the lo halves
      const uint32_t base = ring_u + cst * kTaStage;
      if constexpr (EA == kTaEaTriplet) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = lds128f(base + xoff + i * 1024);
          sts128(base + kTaXHalf + xoff + i * 1024,
                 make_uint4(lo_of_raw(v.x), lo_of_raw(v.y), lo_of_raw(v.z), lo_of_raw(v.w)));
        }
      }
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        if (srow[j] < kTaChunk) {
          const float2 v = lds64f(base + soff[j]);
          sts64(base + kTaSHalf + soff[j], lo_of_raw(v.x), lo_of_raw(v.y));
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[cst]);
      if (++cst == kTaStages) cst = 0;
    };
    int issued = 0;
    bool more = true;
    if (issue()) ++issued; else more = false;
    cp_async_commit();
    for (int it = 0; it < issued; ++it) {
      if (more) { if (issue()) ++issued; else more = false; }
      cp_async_commit();
      cp_async_wait<1>();                                   // the group of chunk `it` is complete
      consume();
    }
  } else {
    // =============================== consumers ====================================================
    const int cw = warp - kTaConsWarp0;
    const int ch = lane * 4;
    const int head = ch / p.C;
    const int lph = LPH > 0 ? LPH : p.C / 4;
    const bool leader = (ch % p.C) == 0;
    const float scale2 = p.scale * 1.4426950408889634f;    // logits in the log2 domain: ex2 softmax
    const uint32_t ea_u = smem_u32(sEa) + ch * 4, sg_u = smem_u32(sSg) + ch * 4;
    const uint32_t mz_u = smem_u32(sSg) + head * 4;        // partial (m, z) of a run: words [head], [32 + head]
    const float* __restrict__ qkvs = p.qkvs;
    const int32_t* __restrict__ rowptr = p.rowptr;
    const int ldq = p.ldq;
    uint32_t tcount = 0;
    TaTile cur, nxt;
    // the record and the source ids of a warp's FIRST item of a tile are loaded during the previous tile: the
    // chain tile bounds -> item record -> source ids -> K / V rows is four dependent global loads
    int2 rec_n = make_int2(0, 0);
    int s_n = 0;
    if ((int)blockIdx.x < p.ntiles) {
      load_tile(blockIdx.x, cur);
      if (cur.i0 + cw < cur.i1) {
        rec_n = __ldg(reinterpret_cast<const int2*>(p.items) + cur.i0 + cw);
        if (lane <= (rec_n.y & 15)) s_n = __ldg(p.src + rec_n.x + lane);
      }
    }
    for (int tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x, ++tcount) {
      const bool has_next = tile + (int)gridDim.x < p.ntiles;
      if (has_next) load_tile(tile + gridDim.x, nxt);
      // merge-phase inputs of this warp's first target, loaded ahead
      const int e_mine = cur.e0 + cw;
      int pb = 0, pe = 0;
      float4 sk = make_float4(0.f, 0.f, 0.f, 0.f);
      if (e_mine < cur.e1) {
        pb = __ldg(rowptr + e_mine) - cur.t0;
        pe = __ldg(rowptr + e_mine + 1) - cur.t0;
        if (p.fuse_skip) sk = __ldg(reinterpret_cast<const float4*>(qkvs + (int64_t)e_mine * ldq + 3 * D + ch));
      }
      // ---- phase 1: items (runs of <= 8 rows of one segment), one partial softmax state each
      for (int it = cur.i0 + cw; it < cur.i1; it += kTaConsWarps) {
        int2 rec = rec_n;
        int s_l = s_n;
        if (it != cur.i0 + cw) {
          rec = __ldg(reinterpret_cast<const int2*>(p.items) + it);
          s_l = (lane <= (rec.y & 15)) ? __ldg(p.src + rec.x + lane) : 0;
        }
        const int rb = rec.x - cur.t0;                      // first row of the item inside the tile
        const int e = rec.y >> 4, cnt = (rec.y & 15) + 1;
        struct Row { float4 k, v; };
        Row ring[kTaRing];
        auto fetch = [&](Row& r, int j) {
          const int s = __shfl_sync(0xffffffffu, s_l, min(j, cnt - 1));
          const float* kp = qkvs + (int64_t)s * ldq + D + ch;
          r.k = __ldg(reinterpret_cast<const float4*>(kp));
          r.v = __ldg(reinterpret_cast<const float4*>(kp + D));
        };
#pragma unroll
        for (int u = 0; u < kTaRing - 1; ++u) fetch(ring[u], u);
        float4 q = __ldg(reinterpret_cast<const float4*>(qkvs + (int64_t)e * ldq + ch));
        q.x *= scale2; q.y *= scale2; q.z *= scale2; q.w *= scale2;
        float4 aseg = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (EA == kTaEaSegment)
          aseg = __ldg(reinterpret_cast<const float4*>(p.ea + (int64_t)__ldg(p.ea_index + e) * D + ch));
        float m = -INFINITY, z = 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        mbar_wait(&chunkfull[(rb + cnt - 1) / kTaChunk], tcount & 1);   // the chunk holding the last row is in the tile
#pragma unroll
        for (int j = 0; j < kTaBlk; ++j) {
          if (j < cnt) {
            fetch(ring[(j + kTaRing - 1) % kTaRing], j + kTaRing - 1);
            const Row& r = ring[j % kTaRing];
            const uint32_t ro = (uint32_t)(rb + j) * 512;
            float4 a;
            if constexpr (EA == kTaEaTriplet) a = lds128f(ea_u + ro);
            else a = aseg;
            const float4 g = lds128f(sg_u + ro);
            float dot = q.x * (r.k.x + a.x);
            dot = fmaf(q.y, r.k.y + a.y, dot);
            dot = fmaf(q.z, r.k.z + a.z, dot);
            dot = fmaf(q.w, r.k.w + a.w, dot);
            const float l2 = head_sum_t<LPH>(dot, lph);     // :150, in log2 units (scale folded into q)
            const float mn = fmaxf(m, l2);
            const float corr = ex2(m - mn);                 // ex2(-inf) = 0 on the first row
            const float pr = ex2(l2 - mn);
            z = fmaf(z, corr, pr);
            acc.x = fmaf(acc.x, corr, pr * (r.v.x + a.x) * g.x);          // :155-160
            acc.y = fmaf(acc.y, corr, pr * (r.v.y + a.y) * g.y);
            acc.z = fmaf(acc.z, corr, pr * (r.v.z + a.z) * g.z);
            acc.w = fmaf(acc.w, corr, pr * (r.v.w + a.w) * g.w);
            m = mn;
          }
        }
        __syncwarp();                                       // every lane has read the rows of the item
        const uint32_t ro = (uint32_t)rb * 512;
        sts128f(ea_u + ro, acc.x, acc.y, acc.z, acc.w);
        if (leader) {
          sts32f(mz_u + ro, m);
          sts32f(mz_u + ro + 128, z);
        }
      }
      const bool item_n = has_next && nxt.i0 + cw < nxt.i1;
      if (item_n) rec_n = __ldg(reinterpret_cast<const int2*>(p.items) + nxt.i0 + cw);
      asm volatile("bar.sync 1, %0;" ::"n"(kTaConsWarps * 32) : "memory");
      // ---- phase 2: warp per target, partial states merged in row order
      for (int e = e_mine; e < cur.e1; e += kTaConsWarps) {
        if (e != e_mine) {
          pb = __ldg(rowptr + e) - cur.t0;
          pe = __ldg(rowptr + e + 1) - cur.t0;
          if (p.fuse_skip) sk = __ldg(reinterpret_cast<const float4*>(qkvs + (int64_t)e * ldq + 3 * D + ch));
        }
        float M = -INFINITY, Z = 0.f;
        float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = pb; r < pe; r += kTaBlk) {
          const uint32_t ro = (uint32_t)r * 512;
          const float4 pa = lds128f(ea_u + ro);
          const float pm = lds32f(mz_u + ro), pz = lds32f(mz_u + ro + 128);
          const float Mn = fmaxf(M, pm);
          const float c0 = ex2(M - Mn), c1 = ex2(pm - Mn);
          Z = fmaf(Z, c0, pz * c1);
          A.x = fmaf(A.x, c0, pa.x * c1);
          A.y = fmaf(A.y, c0, pa.y * c1);
          A.z = fmaf(A.z, c0, pa.z * c1);
          A.w = fmaf(A.w, c0, pa.w * c1);
          M = Mn;
        }
        const float inv = 1.0f / (Z + 1e-16f);              // PyG softmax: out / (sum + 1e-16)
        float4 o = make_float4(A.x * inv, A.y * inv, A.z * inv, A.w * inv);
        *reinterpret_cast<float4*>(p.attn + (int64_t)e * D + ch) = o;
        if (p.fuse_skip) {
          o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;            // :127
          *reinterpret_cast<float4*>(p.out + (int64_t)e * D + ch) = o;
        }
        if (leader) p.lse[(int64_t)e * p.H + head] = pe > pb ? (M + log2f(Z)) * 0.6931471805599453f : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(tileempty);
      s_n = (item_n && lane <= (rec_n.y & 15)) ? __ldg(p.src + rec_n.x + lane) : 0;
      cur = nxt;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

static inline bool tile_fwd_supported(int D, int H, int C, int A, int S, bool ea_segment) {
  if (D != 128 || H * C != D || H > 32 || (C & 3) != 0) return false;
  const int lph = C / 4;
  if ((lph & (lph - 1)) != 0) return false;
  if (S < 2 || S > 64 || (S & 1) != 0) return false;
  if (A == 0 || ea_segment) return true;
  return A == 128;
}

}  // namespace tc
}  // namespace x2
