// tile_attn.cuh -- the T-scale part of the SBFTransformerConv forward (sbftransformer_conv.py:138-162 +
// PyG softmax / sum aggregation) as ONE persistent tcgen05 kernel on the default [T, A] path, plus a small
// merge kernel:
//
//   edge_attr[T,128], sbf[T,S]  --cp.async-->  swizzled smem ring (3xTF32 hi / lo operands, 32-row chunks)
//        --tcgen05.mma (weights resident in TENSOR MEMORY)-->  EA^T, Sg^T chunks in TMEM
//        --4 transposer warps (tcgen05.ld, conflict-free st.shared)-->  ring of [16 triplets][channel] slots
//        --15 attention warps-->  gathers of K / V / Q rows, logits, softmax pieces, sbf gate, aggregation
//        -->  one partial softmax state per ITEM (<= 8 rows of one segment)   --k_item_merge-->  out, lse
//
// EA = lin_edge(edge_attr) and Sg = lin_sbf(sbf) are never read back from HBM by the forward (the unfused pair
// k_tc_gemm x 2 + k_attn_fwd wrote 1 024 B and re-read 1 024 B per triplet); they are optionally stored for the
// backward (ea_out / sg_out), by the attention warps, as whole 512-byte rows.
//
// Work decomposition (x2_items_build, graph.cu).  Every target segment is cut, relative to its own start, into
// ITEMS of at most 8 rows; the item list is in row order, so 15 consecutive items -- a UNIT -- cover a
// contiguous row range of ~100 triplets.  Units are dealt round-robin to the CTAs (one per SM): all SMs then
// work inside the same few molecules at any time and the K / V gathers hit L2.  Inside a CTA the rows of its
// units flow as 32-row chunks through: a 2-stage operand ring -> 2 accumulator buffers in TMEM -> a ring of 8
// shared-memory slots of 16 rows.  Attention warp w owns item w of every unit: it requests the K rows of its item, waits
// until the chunk holding the item's last row is in a slot, forms the logits, requests the V rows, accumulates,
// and writes ONE partial state (m, z, acc[128]) per item to global scratch; a slot is recycled when the warps
// that own rows in it have released it (transposer warp 0 arrives for the others).  Nothing in the kernel waits for a whole unit or a whole segment: segments of any
// length are fine, and the roles only meet through the rings (the first version processed segment-aligned
// tiles in lock step -- fill the tile, consume it, merge, release -- and spent half of every tile period with
// either the front end or the attention warps idle; clock64 traces in profiles/r2_notes.md).
// k_item_merge (warp per target) then combines the partial states of a segment in row order, normalises, adds
// the skip projection and writes out / attn / lse.  The grouping of a segment's rows depends on the segment
// alone, not on its position in the batch: results are deterministic and bitwise independent of the rest of
// the batch.
//
// Budget per CTA: operand ring 2 x 48 KB + slot ring 8 x 16 KB = 224 KB of shared memory; tensor memory:
// accumulators 2 x 64 + W_sbf hi / lo 128 + W_edge hi / lo 256 = 512 columns; 896 threads at 72 registers.
#pragma once
#include "tc_gemm.cuh"

namespace x2 {
namespace tc {

constexpr int kTaChunk = 32;                          // triplets per chunk (UMMA N; 16 for a short last chunk of a unit).
                                                      // The issue of one tcgen05.mma costs the MMA warp ~19 cycles of
                                                      // uniform-datapath instructions, the tensor core 9 / 16 cycles at
                                                      // N = 16 / 32 (tools/umma_rate.cu): N = 16 chunks were issue-bound
constexpr int kTaStages = 2;                          // operand ring
constexpr int kTaXHalf = kTaChunk * 128 * 4;          // 16 KB: [32 rows][128 k] fp32 = 4 K-blocks of 4 KB
constexpr int kTaKBlk = kTaChunk * 128;               // bytes of one K-block: 32 rows x 128 B
constexpr int kTaSHalf = kTaChunk * 64 * 4;           // 8 KB: sbf, K padded to 64 (2 K-blocks)
constexpr int kTaStage = 2 * kTaXHalf + 2 * kTaSHalf; // X hi | X lo | S hi | S lo = 48 KB
constexpr int kTaSlotRows = 16;                       // rows of a slot (half a chunk)
constexpr int kTaSlots = 8;                           // slot ring: [16 rows][128] EA + [16 rows][128] Sg = 16 KB each; must
                                                      // hold a whole unit (<= 120 rows): see the parity note at the waits
constexpr int kTaSlotBytes = kTaSlotRows * 1024;
constexpr int kTaAcc = 2;                             // accumulator buffers in TMEM (64 columns each: EA 32 | Sg 32)
// Warp roles, aligned to warpgroups (4 warps).  Moving registers from the producers / transposers to the
// attention warps with setmaxnreg (40 / 48 / 88 of the 72 per thread at launch; tools/setmaxnreg_probe.cu shows
// which splits the pool allows) compiles and lets the attention warps keep 8 K rows in flight across the wait
// without spilling, but the kernel then dies with "illegal instruction" as soon as the producer warps run their
// cp.async copies after a setmaxnreg.dec -- even a dec to the unchanged 72 (bisected with the ablation bits on
// the hardware, profiles/r2_notes.md).  So every warp keeps the 72 registers of the launch.
constexpr int kTaProdWarp0 = 4, kTaProdWarps = 8;     // warps 0..3: transposers; warps 4..11: producers
constexpr int kTaProdThreads = kTaProdWarps * 32;
constexpr int kTaMmaWarp = 12;                        // warp 12: MMA issuer
constexpr int kTaConsWarp0 = 13, kTaConsWarps = X2_UNIT_ITEMS;   // warps 13..27: one per item of a unit
constexpr int kTaThreads = (kTaConsWarp0 + kTaConsWarps) * 32;   // 896 threads
constexpr int kTaBlk = X2_ITEM_ROWS;                  // rows per item
constexpr int kTaPart = 192;                          // floats of a partial state: acc[128] | m[32] | z[32]
// tensor-memory columns
constexpr int kTaWsHi = 128, kTaWsLo = 192, kTaWeHi = 256, kTaWeLo = 384;
constexpr size_t kTaSmem = 1024 + (size_t)kTaStages * kTaStage + (size_t)kTaSlots * kTaSlotBytes + 512;

enum { kTaEaTriplet = 1, kTaEaSegment = 2, kTaEaNone = 0 };

struct TaParams {
  const float* ea;         // edge_attr [T, 128] (kTaEaTriplet) | lin_edge(table) [M, 128] (kTaEaSegment)
  const int32_t* ea_index; // [E] table row of every target (kTaEaSegment)
  const float* sbf; int S; // [T, S], S even, <= 64
  const float* w_edge;     // [128, 128]
  const float* w_sbf;      // [128, S]
  const float* b_sbf;      // [128]
  const float* qkvs; int ldq;                        // [E, 4*128]  Q | K | V | skip
  const int32_t* src;      // [T]
  const int32_t* items;    // [nitems + 1][2]: first triplet, (target << 4) | (rows - 1); sentinel (T, 0)
  const int32_t* itemptr;  // [E + 1]: items of target e are itemptr[e] .. itemptr[e + 1]
  int64_t E, T;
  int H, C; float scale;
  float* part;             // [nitems][kTaPart] partial softmax states (scratch)
  float *ea_out, *sg_out;  // [T,128] each or NULL (saved for the backward)
  unsigned long long* trace;   // development: clock64() timeline of CTA 0 (tools/tile_probe.py), NULL otherwise
  int dbg;                 // ablation bits (X2GNN_TA_DBG, development only): 1 no item arithmetic, 2 no EA / Sg
                           // stores to HBM, 4 no MMAs, 8 no copies / lo pass, 16 no accumulator read-out
};

__device__ __forceinline__ void tmem_ld16u(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void* ptr, uint32_t bytes) {      // 16-byte aligned, multiple of 16
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}
// Gather load as a volatile asm statement: volatile asms keep their program order, so the K rows are requested
// before the wait, the V row of a triplet only after its K row has been consumed (the compiler otherwise hoists
// the V loads next to the K loads: 64 live registers, spilled -- and a spill store waits for its load).
__device__ __forceinline__ float4 ldg128_v(const float* ptr, uint64_t pol) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(ptr), "l"(pol));
  return v;
}
// L2 eviction policies: the Q | K | V table (E x 2 KB, gathered ~19 times per row) should stay in L2, the
// edge_attr / sbf stream (read once) and the EA / Sg saves (written once) should not displace it
__device__ __forceinline__ uint64_t l2_policy_keep() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_stream() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
template <int BYTES>
__device__ __forceinline__ void cp_async_pol(uint32_t dst, const void* src, uint64_t pol) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
  else
    asm volatile("cp.async.ca.shared.global.L2::cache_hint [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
// The same, ordered after `dep` by a data dependence (the load is predicated on dep == dep, false only for a
// NaN logit, and then keeps the register's old value): the assembler cannot hoist the V loads above the logits.
__device__ __forceinline__ void ldg128_after(float4& v, const float* ptr, float dep) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.eq.f32 p, %5, %5;\n@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n}\n"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w) : "l"(ptr), "f"(dep));
}
// release / acquire on a shared-memory word (slot fill counters: monotonic, so a waiter can never mistake an old
// phase for the one it waits for, as it can with a parity wait on an mbarrier it is more than one phase behind)
__device__ __forceinline__ void st_release_cta(uint32_t* p, uint32_t v) {
  asm volatile("st.release.cta.shared::cta.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ uint4 ld_acquire_cta_v4(const uint32_t* p) {
  uint4 v;
  asm volatile("ld.acquire.cta.shared::cta.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void mbar_arrive_n(uint64_t* bar, uint32_t n) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// development trace: slot s, index i < 256 of CTA 0
#define TA_TR(slot, idx)                                                                       \
  do {                                                                                         \
    if (p.trace && blockIdx.x == 0 && lane == 0 && (idx) < 256)                                \
      p.trace[(slot) * 256 + (idx)] = (unsigned long long)clock64();                           \
  } while (0)

template <int LPH, int EA>
__global__ void __launch_bounds__(kTaThreads, 1) k_tile_fwd(const TaParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sRing = smem;
  uint8_t* sSlot = smem + (size_t)kTaStages * kTaStage;     // kTaSlots x ([16][128] EA | [16][128] Sg)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sSlot + (size_t)kTaSlots * kTaSlotBytes);
  uint64_t* full = bars;                          // [4]  producers -> MMA
  uint64_t* empty = bars + kTaStages;             // [4]  MMA -> producers
  uint64_t* accfull = empty + kTaStages;          // [4]  MMA -> transposers
  uint64_t* accempty = accfull + kTaAcc;          // [4]  transposers -> MMA
  uint64_t* slotempty = accempty + kTaAcc;        // [8]  attention warps (+ transposer warp 0 for the non-owners) -> transposers
  uint32_t* done = reinterpret_cast<uint32_t*>(slotempty + kTaSlots);   // [4] slots filled so far, per transposer warp
  uint32_t* tmem_slot = done + 4;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int D = 128;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kTaStages; ++i) {
      mbar_init(&full[i], kTaProdThreads);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < kTaAcc; ++i) {
      mbar_init(&accfull[i], 1);
      mbar_init(&accempty[i], 128);
    }
    for (int i = 0; i < kTaSlots; ++i) mbar_init(&slotempty[i], kTaConsWarps);
    for (int i = 0; i < 4; ++i) done[i] = 0;
    fence_barrier_init();
  }
  if (warp == kTaMmaWarp) tmem_alloc(tmem_slot, 512);
  {  // the padding of the sbf operand blocks (k >= S) is never written by the copies: zero it once
    const uint32_t ring_u = smem_u32(sRing);
    constexpr int per = 2 * kTaSHalf / 16;
    for (int i = threadIdx.x; i < kTaStages * per; i += kTaThreads) {
      const int stg = i / per, w = i - stg * per;
      sts128(ring_u + stg * kTaStage + 2 * kTaXHalf + w * 16, make_uint4(0u, 0u, 0u, 0u));
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();        // global memory from here on

  // ---- units: 15 consecutive items; every role walks the same sequence (loads are unconditional with clamped
  // indices: conditionally loaded values are not provably uniform for the compiler, and the MMA warp then paid
  // an R2UR.BROADCAST per operand of every tcgen05.mma)
  const int2* __restrict__ items = reinterpret_cast<const int2*>(p.items);
  const int nitems = __ldg(p.itemptr + p.E);
  const int nunits = (p.dbg & 64) ? 0 : (nitems + X2_UNIT_ITEMS - 1) / X2_UNIT_ITEMS;   // dbg 64: roles start and stop
  auto unit_rows = [&](int u, int& t0, int& t1) {          // rows [t0, t1) of unit u (clamped: u may be past the end)
    t0 = __ldg(&items[min(u * X2_UNIT_ITEMS, nitems)].x);
    t1 = __ldg(&items[min((u + 1) * X2_UNIT_ITEMS, nitems)].x);
  };
  const int KS_S0 = (min(p.S, 32) + 7) >> 3;               // k-steps of the two sbf K-blocks
  const int KS_S1 = p.S > 32 ? (p.S - 32 + 7) >> 3 : 0;

  if (warp >= kTaMmaWarp) {
  if (warp == kTaMmaWarp) {
    // =============================== MMA issuer (whole warp, one elected lane issues) ===============
    const uint32_t leader = elect_one();
    const uint32_t idesc32 = make_idesc(kTaChunk, 0, 0);   // M = 128 channels, N = 32 triplets
    const uint32_t idesc16 = make_idesc(16, 0, 0);         // last chunk of a unit with <= 16 rows: half the MMA time
    const uint32_t ring_u = smem_u32(sRing);
    const uint32_t tmem_u = __reduce_max_sync(0xffffffffu, tmem_base);   // provably uniform copy (see the note on idesc)
    asm volatile("bar.sync 2, 160;" ::: "memory");         // weights are in tensor memory
    tc_fence_after();
    uint32_t st = 0, ph = 0, cc = 0;
    int t0, t1, n0, n1;
    unit_rows(blockIdx.x, t0, t1);
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      unit_rows(u + gridDim.x, n0, n1);
      const int nch = (t1 - t0 + kTaChunk - 1) / kTaChunk;
      for (int c = 0; c < nch; ++c, ++cc) {
        const uint32_t buf = cc % kTaAcc;
        mbar_wait_w(&accempty[buf], ((cc / kTaAcc) & 1) ^ 1);
        TA_TR(0, cc);
        mbar_wait_w(&full[st], ph);
        tc_fence_after();
        TA_TR(1, cc);
        const uint32_t base = ring_u + st * kTaStage;
        // (redux: a value the assembler knows to be uniform -- operands of the MMA that are derived from loaded
        // data otherwise go through an R2UR per instruction)
        const uint32_t idesc = __reduce_max_sync(0xffffffffu, t1 - t0 - c * kTaChunk) > 16 ? idesc32 : idesc16;
        const uint32_t t_ea = tmem_u + buf * 64, t_sg = t_ea + 32;
        if (!(p.dbg & 4)) {
          if constexpr (EA == kTaEaTriplet) {
#pragma unroll 1
            for (int kc = 0; kc < 4; ++kc) {               // EA^T = W_e . X^T
              uint64_t dh = make_desc(base + kc * kTaKBlk, 16, 1024);
              uint64_t dl = make_desc(base + kTaXHalf + kc * kTaKBlk, 16, 1024);
              uint32_t w_hi = tmem_u + kTaWeHi + kc * 32, w_lo = tmem_u + kTaWeLo + kc * 32;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks, dh += 2, dl += 2, w_hi += 8, w_lo += 8) {
                umma_tf32_ts_w(leader, t_ea, w_hi, dh, idesc, (kc | ks) != 0);
                umma_tf32_ts_w(leader, t_ea, w_lo, dh, idesc, 1);
                umma_tf32_ts_w(leader, t_ea, w_hi, dl, idesc, 1);
              }
            }
          }
#pragma unroll 1
          for (int kc = 0; kc < 2; ++kc) {                 // Sg^T = W_s . sbf^T
            const int ksteps = kc == 0 ? KS_S0 : KS_S1;
            uint64_t dh = make_desc(base + 2 * kTaXHalf + kc * kTaKBlk, 16, 1024);
            uint64_t dl = make_desc(base + 2 * kTaXHalf + kTaSHalf + kc * kTaKBlk, 16, 1024);
            uint32_t w_hi = tmem_u + kTaWsHi + kc * 32, w_lo = tmem_u + kTaWsLo + kc * 32;
            for (int ks = 0; ks < ksteps; ++ks, dh += 2, dl += 2, w_hi += 8, w_lo += 8) {
              umma_tf32_ts_w(leader, t_sg, w_hi, dh, idesc, (kc | ks) != 0);
              umma_tf32_ts_w(leader, t_sg, w_lo, dh, idesc, 1);
              umma_tf32_ts_w(leader, t_sg, w_hi, dl, idesc, 1);
            }
          }
        }
        umma_commit_w(leader, &empty[st]);
        umma_commit_w(leader, &accfull[buf]);
        if (++st == kTaStages) { st = 0; ph ^= 1; }
      }
      t0 = n0; t1 = n1;
    }
    __syncwarp();
  } else {
    // =============================== attention warps ==============================================
    const int cw = warp - kTaConsWarp0;
    const int ch = lane * 4;
    const int head = ch / p.C;
    const int lph = LPH > 0 ? LPH : p.C / 4;
    const bool leader = (ch % p.C) == 0;
    const float scale2 = p.scale * 1.4426950408889634f;    // logits in the log2 domain: ex2 softmax
    const uint64_t pol_keep = l2_policy_keep();
    const uint32_t slot_u = smem_u32(sSlot) + ch * 4;
    const float* __restrict__ qkvs = p.qkvs;
    const int ldq = p.ldq;
    float* __restrict__ eao = (p.dbg & 2) ? nullptr : p.ea_out;
    float* __restrict__ sgo = (p.dbg & 2) ? nullptr : p.sg_out;
    uint32_t sc = 0;                                        // slots (16-row half chunks) of the CTA before the current unit
    uint32_t ucount = 0;
    int t0, t1, n0, n1;
    unit_rows(blockIdx.x, t0, t1);
    // the record and the source ids of the warp's item of the NEXT unit are loaded during the current one: the
    // chain unit bounds -> item record -> source ids -> K / V rows is four dependent global loads
    int2 rec_n = __ldg(&items[min((int)blockIdx.x * X2_UNIT_ITEMS + cw, nitems)]);
    int s_n = __ldg(p.src + min((int64_t)rec_n.x + min(lane, rec_n.y & 15), p.T - 1));
    for (int u = blockIdx.x; u < nunits; u += gridDim.x, ++ucount) {
      unit_rows(u + gridDim.x, n0, n1);
      const int it = u * X2_UNIT_ITEMS + cw;
      const int2 rec = rec_n;
      const int s_l = s_n;
      rec_n = __ldg(&items[min((u + (int)gridDim.x) * X2_UNIT_ITEMS + cw, nitems)]);
      if (it < nitems && !(p.dbg & 1)) {
        const int rb = rec.x - t0;                          // first row of the item inside the unit
        const int e = rec.y >> 4, cnt = (rec.y & 15) + 1, last = cnt - 1;
        // All K rows of the item are requested at once (8 x 16 bytes per lane in flight), the V rows after the
        // logits.  Every loop runs over all 8 rows with the row index clamped to the item (rows past its end
        // repeat the last one with weight 0): predicated loads made the compiler merge each result through a
        // temporary, which serialised the gathers (one ~800-cycle load at a time; 4 100 cycles for the logits
        // of 8 rows in the clock64 trace).  No online rescaling: the logits are kept, the maximum taken once.
        float4 aseg = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (EA == kTaEaSegment)
          aseg = __ldg(reinterpret_cast<const float4*>(p.ea + (int64_t)__ldg(p.ea_index + e) * D + ch));
        // shared-memory address of row j of the item: slot of its half chunk, row inside the slot
        auto row_addr = [&](int j) -> uint32_t {
          const int r = rb + min(j, last);
          return slot_u + ((sc + (uint32_t)(r / kTaSlotRows)) % kTaSlots) * kTaSlotBytes + (uint32_t)(r % kTaSlotRows) * 512;
        };
        const uint32_t k_first = sc + (uint32_t)(rb / kTaSlotRows), k_last = sc + (uint32_t)((rb + last) / kTaSlotRows);
        if (cw == 0 || cw == kTaConsWarps - 1) TA_TR(cw == 0 ? 7 : 11, (int)ucount);
        for (;;) {                                          // all four transposer warps have filled slot use k_last
          const uint4 dn = ld_acquire_cta_v4(done);
          if (min(min(dn.x, dn.y), min(dn.z, dn.w)) > k_last) break;
        }
        if (cw == 0 || cw == kTaConsWarps - 1) TA_TR(cw == 0 ? 8 : 12, (int)ucount);
        // (requested after the wait: 32 registers of loads in flight across it do not fit the 72 of the warp --
        // the assembler spills exactly those values, and a spill store waits for its load)
        float4 kv[kTaBlk];
        const float* const kbase = qkvs + D + ch;
#pragma unroll
        for (int j = 0; j < kTaBlk; ++j) {
          const int s = __shfl_sync(0xffffffffu, s_l, min(j, last));
          kv[j] = ldg128_v(kbase + (int64_t)s * ldq, pol_keep);
        }
        float4 q = ldg128_v(qkvs + (int64_t)e * ldq + ch, pol_keep);
        q.x *= scale2; q.y *= scale2; q.z *= scale2; q.w *= scale2;
        float l2[kTaBlk];
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < kTaBlk; ++j) {
          float4 a;
          if constexpr (EA == kTaEaTriplet) a = lds128f(row_addr(j));
          else a = aseg;
          float dot = q.x * (kv[j].x + a.x);
          dot = fmaf(q.y, kv[j].y + a.y, dot);
          dot = fmaf(q.z, kv[j].z + a.z, dot);
          dot = fmaf(q.w, kv[j].w + a.w, dot);
          l2[j] = head_sum_t<LPH>(dot, lph);                // :150, in log2 units (scale folded into q)
          m = fmaxf(m, l2[j]);                              // a repeated row repeats a logit: the maximum is unchanged
        }
        // the K rows are dead: request the V rows into the same registers.  The fence keeps the assembler from
        // hoisting these loads above the logits (K and V together are 64 live registers, spilled -- and a spill
        // store waits for its load, which serialised the gathers again).  Ordering them through a data dependence
        // on the maximum instead (predicated load / address + isnan) cost more registers than it saved.
        __threadfence_block();
#pragma unroll
        for (int j = 0; j < kTaBlk; ++j) {
          const int s = __shfl_sync(0xffffffffu, s_l, min(j, last));
          kv[j] = ldg128_v(kbase + D + (int64_t)s * ldq, pol_keep);
        }
        if (cw == kTaConsWarps - 1) TA_TR(15, (int)ucount);
        float z = 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < kTaBlk; ++j) {
          const uint32_t ra = row_addr(j);
          float4 a;
          if constexpr (EA == kTaEaTriplet) a = lds128f(ra);
          else a = aseg;
          const float4 g = lds128f(ra + kTaSlotRows * 512);
          if (j < cnt) {
            if constexpr (EA == kTaEaTriplet) {
              if (eao) __stcs(reinterpret_cast<float4*>(eao + ((int64_t)rec.x + j) * D + ch), a);
            }
            if (sgo) __stcs(reinterpret_cast<float4*>(sgo + ((int64_t)rec.x + j) * D + ch), g);
          }
          const float pr = j < cnt ? ex2(l2[j] - m) : 0.f;
          z += pr;
          acc.x = fmaf(pr * (kv[j].x + a.x), g.x, acc.x);                 // :155-160
          acc.y = fmaf(pr * (kv[j].y + a.y), g.y, acc.y);
          acc.z = fmaf(pr * (kv[j].z + a.z), g.z, acc.z);
          acc.w = fmaf(pr * (kv[j].w + a.w), g.w, acc.w);
        }
        float* __restrict__ pp = p.part + (int64_t)it * kTaPart;
        *reinterpret_cast<float4*>(pp + ch) = acc;
        if (leader) {
          pp[128 + head] = m;
          pp[160 + head] = z;
        }
        __syncwarp();                                       // every lane has read the rows of the item
        if (lane == 0) {                                    // release the one or two slots that hold them
          mbar_arrive(&slotempty[k_first % kTaSlots]);
          if (k_last != k_first) mbar_arrive(&slotempty[k_last % kTaSlots]);
        }
      }
      if (cw == 0 || cw == kTaConsWarps - 1) TA_TR(cw == 0 ? 9 : 13, (int)ucount);
      s_n = __ldg(p.src + min((int64_t)rec_n.x + min(lane, rec_n.y & 15), p.T - 1));
      sc += (uint32_t)((t1 - t0 + kTaSlotRows - 1) / kTaSlotRows);
      t0 = n0; t1 = n1;
    }
  }
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
    const uint32_t slot_u = smem_u32(sSlot) + ch * 4;
    uint32_t cc = 0, sc = 0;                              // chunks / slots (half chunks) so far
    int t0, t1, n0, n1;
    unit_rows(blockIdx.x, t0, t1);
    // first rows of the unit's items (lane i: item i; lane 15: end of the unit), one unit ahead: warp 0 tells the
    // slot barrier how many attention warps have NO row in a slot, so that only the owners have to arrive
    auto item_start = [&](int u) { return __ldg(&items[min(u * X2_UNIT_ITEMS + min(lane, X2_UNIT_ITEMS), nitems)].x); };
    int xs_n = item_start(blockIdx.x);
    for (int u = blockIdx.x; u < nunits; u += gridDim.x) {
      unit_rows(u + gridDim.x, n0, n1);
      const int xs = xs_n;
      xs_n = item_start(u + gridDim.x);
      int f0 = 0, f1 = 0;
      if (q == 1) unit_rows(u + 2 * gridDim.x, f0, f1);     // used at the end of the iteration (L2 prefetch)
      const int xe = __shfl_down_sync(0xffffffffu, xs, 1);  // end of item `lane`
      const int nch = (t1 - t0 + kTaChunk - 1) / kTaChunk;
      for (int c = 0; c < nch; ++c, ++cc) {
        const uint32_t buf = cc % kTaAcc;
        mbar_wait(&accfull[buf], (cc / kTaAcc) & 1);
        tc_fence_after();
        if (q == 1) TA_TR(2, cc);
        const int halves = (t1 - t0 - c * kTaChunk) > kTaSlotRows ? 2 : 1;
        for (int h = 0; h < halves; ++h, ++sc) {              // one 16-row slot per half chunk
          const uint32_t slot = sc % kTaSlots;
          mbar_wait(&slotempty[slot], ((sc / kTaSlots) & 1) ^ 1);   // the previous use of the slot has been released
          if (q == 0) {                                       // non-owners of this slot: arrive on their behalf
            const int rlo = t0 + c * kTaChunk + h * kTaSlotRows, rhi = rlo + kTaSlotRows - 1;
            const unsigned own = __ballot_sync(0xffffffffu, lane < X2_UNIT_ITEMS && xs <= rhi && xe - 1 >= rlo && xe > xs);
            const int others = kTaConsWarps - __popc(own);
            if (lane == 0 && others > 0) mbar_arrive_n(&slotempty[slot], (uint32_t)others);
          }
          if (!(p.dbg & 16)) {
            const uint32_t dst = slot_u + slot * kTaSlotBytes;
            float v[16];
            if constexpr (EA == kTaEaTriplet) {
              tmem_ld16u(trow + buf * 64 + h * 16, v);
#pragma unroll
              for (int j = 0; j < 16; ++j) sts32f(dst + j * 512, v[j]);
            }
            tmem_ld16u(trow + buf * 64 + 32 + h * 16, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) sts32f(dst + kTaSlotRows * 512 + j * 512, v[j] + bs);
          }
          __syncwarp();
          if (lane == 0) st_release_cta(&done[q], sc + 1);    // this warp's 32 channels of slot use `sc` are written
        }
        tc_fence_before();
        mbar_arrive(&accempty[buf]);
        if (q == 1) TA_TR(3, cc);
      }
      if (q == 1 && lane == 0 && f1 > f0) {
        // pull the rows of the unit two rounds ahead into L2 (one thread, two instructions), so that the
        // producers' copies find their data in L2 rather than in HBM
        if constexpr (EA == kTaEaTriplet) l2_prefetch(p.ea + (int64_t)f0 * D, (uint32_t)(f1 - f0) * D * 4);
        const uintptr_t b0 = reinterpret_cast<uintptr_t>(p.sbf + (int64_t)f0 * p.S) & ~(uintptr_t)15;
        const uintptr_t b1 = reinterpret_cast<uintptr_t>(p.sbf + (int64_t)f1 * p.S) & ~(uintptr_t)15;
        if (b1 > b0) l2_prefetch(reinterpret_cast<const void*>(b0), (uint32_t)(b1 - b0));
      }
      t0 = n0; t1 = n1;
    }
  } else {
    // =============================== producers ====================================================
    // X chunk = 32 rows x 32 16-byte pieces: thread (r0 = pt / 32, c16 = pt % 32) moves column piece c16 of
    // rows r0 + 8 i.  sbf chunk = 32 rows x S/2 8-byte pieces: thread moves pieces pt + 256 j.
    const int pt = threadIdx.x - kTaProdWarp0 * 32;
    const int r0 = pt >> 5, c16 = pt & 31;
    const uint32_t xoff = (uint32_t)((c16 >> 3) * kTaKBlk + r0 * 128 + (((c16 & 7) ^ r0) << 4));
    const int S = p.S, S2 = S >> 1;
    constexpr int NX = kTaChunk / 8, NS = 3;
    uint32_t soff[NS];
    int srow[NS], scol[NS];
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      const int f = pt + kTaProdThreads * j;
      const int row = f / S2, col = (f - row * S2) * 2;
      srow[j] = f < kTaChunk * S2 ? row : 1 << 20;         // inactive pieces never pass the row test
      scol[j] = col;
      soff[j] = (uint32_t)(2 * kTaXHalf + (col >> 5) * kTaKBlk) + kmajor_off(row & (kTaChunk - 1), (col & 31) >> 2) + (uint32_t)(col & 3) * 4;
    }
    const uint32_t ring_u = smem_u32(sRing);
    const uint64_t pol_stream = l2_policy_stream();
    const float* __restrict__ xg = p.ea;
    const float* __restrict__ sg_ = p.sbf;
    uint32_t ist = 0, iph = 0, cst = 0;
    int ntr_issue = 0;
    int iu = blockIdx.x, ic = 0, it0, it1;
    unit_rows(iu, it0, it1);
    auto issue = [&]() -> bool {                            // copies of the next chunk; false when none is left
      while (iu < nunits && ic * kTaChunk >= it1 - it0) {
        iu += gridDim.x;
        ic = 0;
        unit_rows(iu, it0, it1);
      }
      if (iu >= nunits) return false;
      const int valid = min(kTaChunk, it1 - it0 - ic * kTaChunk);
      const int64_t tb = (int64_t)it0 + ic * kTaChunk;
      ++ic;
      mbar_wait(&empty[ist], iph ^ 1);                     // the MMAs that read this stage have retired
      if (warp == kTaProdWarp0) TA_TR(4, ntr_issue);
      ++ntr_issue;
      const uint32_t base = ring_u + ist * kTaStage;
      if (!(p.dbg & 8)) {
        if constexpr (EA == kTaEaTriplet) {
          const float* src = xg + (tb + r0) * D + c16 * 4;
#pragma unroll
          for (int i = 0; i < NX; ++i)
            if (r0 + 8 * i < valid) cp_async_pol<16>(base + xoff + i * 1024, src + (int64_t)i * 8 * D, pol_stream);
        }
#pragma unroll
        for (int j = 0; j < NS; ++j)
          if (srow[j] < valid) cp_async_pol<8>(base + soff[j], sg_ + (tb + srow[j]) * S + scol[j], pol_stream);
      }
      if (++ist == kTaStages) { ist = 0; iph ^= 1; }
      return true;
    };
    auto consume = [&]() {                                  // own pieces have landed: derive the lo halves
      const uint32_t base = ring_u + cst * kTaStage;
      if (!(p.dbg & 8)) {
        if constexpr (EA == kTaEaTriplet) {
#pragma unroll
          for (int i = 0; i < NX; ++i) {
            const float4 v = lds128f(base + xoff + i * 1024);
            sts128(base + kTaXHalf + xoff + i * 1024,
                   make_uint4(lo_of_raw(v.x), lo_of_raw(v.y), lo_of_raw(v.z), lo_of_raw(v.w)));
          }
        }
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          if (srow[j] < kTaChunk) {
            const float2 v = lds64f(base + soff[j]);
            sts64(base + kTaSHalf + soff[j], lo_of_raw(v.x), lo_of_raw(v.y));
          }
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[cst]);
      if (++cst == kTaStages) cst = 0;
    };
    // one chunk of copies in flight beside the one being split; the lo pass of chunk i runs before the wait for the
    // stage of chunk i + 1, i.e. concurrently with the MMAs of chunk i - 1
    int issued = 0;
    bool more = true;
#pragma unroll
    for (int u = 0; u < kTaStages - 1; ++u) {
      if (more) { if (issue()) ++issued; else more = false; }
      cp_async_commit();
    }
    for (int it = 0; it < issued; ++it) {
      cp_async_wait<kTaStages - 2>();                       // the group of chunk `it` is complete
      if (warp == kTaProdWarp0) TA_TR(5, it);
      consume();
      if (warp == kTaProdWarp0) TA_TR(6, it);
      if (more) { if (issue()) ++issued; else more = false; }
      cp_async_commit();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kTaMmaWarp) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// out[e] = sum over the items of segment e of their partial states, merged in row order (flash-decoding
// combine), normalised (PyG softmax: / (sum + 1e-16)), + skip projection; lse for the backward.  Warp per target.
__global__ void __launch_bounds__(128) k_item_merge(const float* __restrict__ part, const int32_t* __restrict__ itemptr,
                                                    const float* __restrict__ qkvs, int ldq, int64_t E, int H, int C,
                                                    int fuse_skip, float* __restrict__ out, float* __restrict__ attn,
                                                    float* __restrict__ lse) {
  pdl_sync();
  constexpr int D = 128;
  const int64_t e = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (e >= E) return;
  const int lane = threadIdx.x & 31, ch = lane * 4;
  const int head = ch / C;
  const bool leader = (ch % C) == 0;
  const int i0 = __ldg(itemptr + e), i1 = __ldg(itemptr + e + 1);
  float4 sk = make_float4(0.f, 0.f, 0.f, 0.f);
  if (fuse_skip) sk = __ldg(reinterpret_cast<const float4*>(qkvs + e * ldq + 3 * D + ch));
  float M = -INFINITY, Z = 0.f;
  float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i = i0; i < i1; ++i) {
    const float* __restrict__ pp = part + (int64_t)i * kTaPart;
    const float4 pa = __ldcs(reinterpret_cast<const float4*>(pp + ch));
    const float pm = __ldcs(pp + 128 + head), pz = __ldcs(pp + 160 + head);
    const float Mn = fmaxf(M, pm);
    const float c0 = ex2(M - Mn), c1 = ex2(pm - Mn);
    Z = fmaf(Z, c0, pz * c1);
    A.x = fmaf(A.x, c0, pa.x * c1);
    A.y = fmaf(A.y, c0, pa.y * c1);
    A.z = fmaf(A.z, c0, pa.z * c1);
    A.w = fmaf(A.w, c0, pa.w * c1);
    M = Mn;
  }
  const float inv = 1.0f / (Z + 1e-16f);                    // PyG softmax: out / (sum + 1e-16)
  float4 o = make_float4(A.x * inv, A.y * inv, A.z * inv, A.w * inv);
  *reinterpret_cast<float4*>(attn + e * D + ch) = o;
  if (fuse_skip) {
    o.x += sk.x; o.y += sk.y; o.z += sk.z; o.w += sk.w;     // :127
    *reinterpret_cast<float4*>(out + e * D + ch) = o;
  }
  if (leader) lse[e * H + head] = i1 > i0 ? (M + log2f(Z)) * 0.6931471805599453f : 0.f;
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
