// tc_gemm.cuh -- hand-written tcgen05 (5th-gen tensor core) GEMMs for the Linear layers of the
// conv path, in 3xTF32 split precision so the result keeps fp32 accuracy (1e-5 parity mode):
//
//     x = x_hi + x_lo   (x_hi = rna_tf32(x), x_lo = rna_tf32(x - x_hi))
//     A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo        (fp32 accumulation in TMEM)
//
// Two persistent, warp-specialised kernels (1 CTA / SM, 800 threads):
//   warp 0      : TMEM allocation + single-thread tcgen05.mma issue
//   warps 1..16 : producers.  Aligned operands (the conv path): every thread copies its 16-byte (8-byte
//                 for 168-byte sbf rows) pieces of raw fp32 with cp.async straight into the hi half of
//                 the UMMA canonical smem layout, several chunks ahead and without holding registers,
//                 then reads its own pieces back and derives the lo half (see "Split precision" below).
//                 Unaligned operands: LDG.128 -> hi/lo split in registers -> STS.128.
//   warps 17..24: epilogue -- tcgen05.ld of the fp32 accumulator (one output row per thread),
//                 transposed through a padded smem tile so that every global store instruction
//                 writes whole 128-byte lines, bias / accumulate fused
// synchronised with mbarriers (full/empty smem ring, full/empty double-buffered TMEM).
//
//   G1  C[M,N] (+)= X[M,K] . B(K,N) + bias.  The WEIGHTS are the MMA's A operand and live in TENSOR
//       MEMORY for the whole kernel (hi and lo halves, 2 x K columns, loaded once with tcgen05.st);
//       the streamed activations are the B operand (K-major smem ring).  The accumulator is therefore
//       C^T (lane = output channel, column = row of the tile): no smem for weights (160 KB ring instead
//       of a 128 KB weight image + 2 stages; keeping smem <= 192 KB also keeps a 60 KB L1, which LDG
//       streaming needs -- tools/bw_probe.cu) and the epilogue's stores are coalesced as they come
//       (32 lanes = 32 consecutive channels of one output row).
//   G2  P[cta][Dm,N] = sum_{rows of this CTA} Y[row,:]^T X[row,:]   both operands streamed and
//       MN-major; the accumulator stays in TMEM over the CTA's whole row range (split-K across
//       CTAs, partial tiles summed in fixed order by k_splitk_reduce => deterministic wgrad).
//
// Split precision: the tensor core TRUNCATES fp32 operand bits to tf32 (verified with
// tools/umma_ts_probe.cu).  Register-staged operands and the weights are split in software into
// hi = rn_tf32(x), lo = rn_tf32(x - hi).  cp.async-staged activations keep the RAW word as the hi operand
// (hardware truncation) and lo = rn_tf32(x - trunc_tf32(x)): the difference is exact, |x - hi - lo| <=
// 2^-21 |x| with either scheme.
#pragma once
#include "common.cuh"

namespace x2 {
namespace tc {

constexpr int kProducerWarps = 16;   // LDG bandwidth of one CTA/SM is capped by its warp count: 8 warps ->
                                     // 3.8 TB/s, 16 -> 6.0 TB/s whatever the loads in flight (tools/bw_probe.cu)
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpilogueWarps = 4;    // one per TMEM lane quarter.  21 warps in all: registers are allocated per
                                     // 4 warps, so 24 warps' worth = 80 registers per thread (25 warps gave 72),
                                     // which is what a 4th chunk of loads in flight per producer thread needs
constexpr int kEpiHalves = kEpilogueWarps / 4;
constexpr int kThreads = 32 + kProducerThreads + kEpilogueWarps * 32;   // MMA warp + producers + epilogue
constexpr int kEpilogueThreads = kEpilogueWarps * 32;
constexpr int kStageWords = 20;      // padded row pitch (words) of the [32][16] epilogue transpose tiles
constexpr int kTileM = 128;          // rows per tile (UMMA M)
constexpr int kChunkK = 32;          // fp32 per 128-byte swizzle row
constexpr int kChunkBytes = kTileM * kChunkK * 4;   // 16 KB: one 128 x 32 fp32 operand chunk
constexpr int kMaxSmem = 232448;     // 227 KB opt-in limit per CTA
constexpr int kTmemCols = 256;       // 2 accumulator buffers x 128 columns

// ------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// Warp-collective wait: every lane polls, the loop branch is taken on a VOTE result.  A vote result is uniform by
// construction, so the compiler keeps treating the code after the wait as convergent and may use the uniform
// datapath for it -- with the per-thread branch of mbar_wait() the MMA-issuing warp computed its descriptors in
// vector registers and paid an R2UR.BROADCAST waterfall (~8 instructions, ~38 cycles) per tcgen05.mma, which
// made the ISSUE loop, not the tensor core (16 cycles per M128 N32 K8 tf32 MMA, tools/umma_rate.cu), the
// period of the pipeline.
__device__ __forceinline__ void mbar_wait_w(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "WAIT_LOOP_W:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "vote.sync.all.pred q, p, 0xffffffff;\n"
      "@q bra WAIT_DONE_W;\n"
      "bra WAIT_LOOP_W;\n"
      "WAIT_DONE_W:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t taddr, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(taddr),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before)
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
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
// Round-to-nearest (ties away) fp32 -> tf32 as two integer ops.  cvt.rna.tf32.f32 compiles to four
// instructions per element (add, |x|<inf compare, select, mask); inputs here are finite activations,
// so the inf/nan guard is dropped: (bits + 2^12) & ~(2^13 - 1).
__device__ __forceinline__ uint32_t to_tf32(float x) {
  return (__float_as_uint(x) + 0x1000u) & 0xffffe000u;
}
__device__ __forceinline__ void sts128f(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128f(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// 3xTF32 split.  The tensor core truncates operand bits to tf32, so both parts are rounded to NEAREST
// tf32 here (two integer ops each): hi = rn(x), lo = rn(x - hi).  |x - hi - lo| <= 2^-23 |x| and unbiased;
// feeding the raw bits instead leaves a one-sided 2^-21 residual (measured: 5e-6 vs 1e-6 layer error).
__device__ __forceinline__ uint32_t hi_bits(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ float lo_part(float x) {
  const float l = x - __uint_as_float(hi_bits(x));
  return __uint_as_float(hi_bits(l));
}
__device__ __forceinline__ void split4(const float4& v, uint4& hi, uint4& lo) {
  hi.x = hi_bits(v.x); hi.y = hi_bits(v.y); hi.z = hi_bits(v.z); hi.w = hi_bits(v.w);
  lo.x = __float_as_uint(lo_part(v.x));
  lo.y = __float_as_uint(lo_part(v.y));
  lo.z = __float_as_uint(lo_part(v.z));
  lo.w = __float_as_uint(lo_part(v.w));
}
// D[tmem] (+)= A[tmem] . B[smem desc]  (A operand in tensor memory: lane = row, 32-bit column = k)
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective forms: EVERY lane of a converged warp executes the call with the same operands and one
// elected lane issues the instruction.  Keeping the MMA warp's loops warp-uniform lets the compiler hold
// descriptors and tensor-memory addresses in uniform registers; with the loops under `if (lane == 0)` it
// wrapped every UTCHMMA in a vote / elect / 4 x R2UR.BROADCAST waterfall (~14 dependent instructions per
// MMA): the single issuing thread then needed ~1 650 cycles per 32-row chunk whatever the number of
// MMAs, and that -- not HBM, not the tensor pipe -- was the period of the whole pipeline (ablation in
// profiles/r1_notes.md: MMA path alone 0.152 ms of gemm_e's 0.181 ms, one pass instead of three: same).
__device__ __forceinline__ uint32_t elect_one() {        // 1 in exactly one lane of the (converged) warp
  uint32_t is_leader;
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "elect.sync _|q, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, q;\n"
      "}\n" : "=r"(is_leader));
  return is_leader;
}
// `leader` comes from ONE elect_one() call per role, so the same thread issues the MMAs and the commits
// that track them.
__device__ __forceinline__ void umma_tf32_ts_w(uint32_t leader, uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_tf32_w(uint32_t leader, uint32_t taddr, uint64_t adesc, uint64_t bdesc,
                                            uint32_t idesc, uint32_t accumulate) {      // A and B from smem
  asm volatile(
      "{\n"
      ".reg .pred p, q;\n"
      "setp.ne.b32 q, %5, 0;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(taddr),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint32_t leader, uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred q;\n"
      "setp.ne.b32 q, %1, 0;\n"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(leader)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

// Ampere-style asynchronous copies global -> shared (LDGSTS): no register staging, arbitrary 8/16-byte
// destination, so a thread can drop raw fp32 straight into a swizzled UMMA operand layout and keep as
// many chunks in flight as the smem ring is deep.
template <int BYTES>
__device__ __forceinline__ void cp_async(uint32_t dst, const void* src) {
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// lo part when the hi part is the RAW fp32 word (the tensor core truncates it to tf32):
// lo = rn_tf32(x - trunc_tf32(x)); x - trunc(x) is exact (13 significant bits), |x - hi - lo| <= 2^-21 |x|
__device__ __forceinline__ uint32_t lo_of_raw(float x) {
  const float l = x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
  return hi_bits(l);
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts64(uint32_t addr, uint32_t a, uint32_t b) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}

// UMMA shared-memory descriptor, Blackwell version (cute SmemDescriptor layout):
// [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type
constexpr uint32_t kLayoutSW128 = 2;         // SWIZZLE_128B           (K-major operands)
constexpr uint32_t kLayoutSW128Base32 = 1;   // SWIZZLE_128B_BASE32B   (the only MN-major tf32 layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout = kLayoutSW128) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}
// UMMA instruction descriptor (cute InstrDescriptor): c=F32, a=b=TF32, M=128
__host__ __device__ constexpr uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

// Byte offset of element (row r, 16-byte column chunk c16 in [0,8)) inside a K-major SW128 operand
// chunk (rows x 32 fp32): 8-row groups of 1024 B, chunk index XOR-swizzled by the row.
__host__ __device__ __forceinline__ uint32_t kmajor_off(int r, int c16) {
  return (uint32_t)((r >> 3) * 1024 + (r & 7) * 128 + ((c16 ^ (r & 7)) << 4));
}

// ------------------------------------------------------------------ G1
constexpr int kTmemWHi = 256;        // TMEM columns [256, 384): W_hi ; [384, 512): W_lo ; [0, 256): 2 accumulators
constexpr int kTmemWLo = 384;

constexpr int kMaxProb = 4;
constexpr int kMaxProbG2 = 16;       // weight gradients of one k_tc_wgrad launch (x2_tc_wgrad_batch: deferred wgrads)
constexpr int kG1Flight = 4;         // chunks of global loads in flight per producer thread
struct G1Prob {
  const float* A;        // streamed activations X[M, K]
  int64_t lda;
  const float* W;        // weights: B(k, n) = W[k * sbk + n * sbn]
  const float* bias;
  float* C;
  int64_t ldc;
  int beta;
};
// Up to kMaxProb problems of identical shape run back to back inside one persistent launch (the Q / K /
// V / skip projections, or the dgrad pairs): the ~15 us start-up of a launch (TMEM allocation, pipeline
// fill, tail) is paid once; only the weights in tensor memory are swapped between problems.
// The problems form `ngroups` groups of `ppg` consecutive problems: CTA b works for group b % ngroups
// only (its tiles are b / ngroups, + gridDim / ngroups, ...) and runs that group's problems back to
// back.  Problems that write different outputs go to different groups, so a CTA loads its weights
// once and its tile pipeline never drains; problems accumulating into one output share a group.
struct G1Params {
  G1Prob prob[kMaxProb];
  int nprob;
  int ngroups, ppg;      // nprob = ngroups * ppg ; gridDim.x is a multiple of ngroups
  int64_t M;
  int K, KC;             // KC = ceil(K / 32) <= 4
  int64_t sbk, sbn;      // weight strides (shared by the problems)
  int N;                 // output channels <= 128
  int stages;
  int single;            // 1: one tf32 pass (hi . hi only): X2_MODE_TF32, ~5e-4 relative error
};

// PIECE = 0: producers stage through registers (LDG -> split -> STS; any alignment).
// PIECE = 16 / 8: producers copy raw fp32 pieces of that many bytes with cp.async straight into the hi
// half of the ring stage (kG1Ahead chunks ahead of their use, no registers held), then read their OWN
// pieces back, derive the lo half and publish the stage.  The hi operand is the raw word: the tensor core
// truncates it.  Needs lda * 4 and the base address to be multiples of PIECE and K * 4 a multiple of PIECE.
template <int PIECE>
__global__ void __launch_bounds__(kThreads, 1) k_tc_gemm(const G1Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  uint8_t* sA = smem;                                             // S x [hi 16K | lo 16K]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sA + (size_t)S * 2 * kChunkBytes);
  uint64_t* full = bars;            // [S]   producers -> MMA
  uint64_t* empty = bars + S;       // [S]   MMA -> producers
  uint64_t* tfull = bars + 2 * S;   // [2]   MMA -> epilogue
  uint64_t* tempty = bars + 2 * S + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], kProducerThreads);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], kEpilogueThreads);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_sync();        // everything above overlapped the previous launch's tail; global memory from here on

  // ---- weights -> tensor memory (first epilogue warp of each lane quarter): lane = output channel n.
  // Loads are issued 32 at a time; only the MMA warp waits for the weights (named barrier 2), the
  // producers start streaming immediately.
  auto load_weights = [&](const float* __restrict__ W) {
    const int q = warp & 3;
    const int n = q * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    const float* __restrict__ wrow = W + (int64_t)n * p.sbn;
    const bool nv = n < p.N;
    // forward layout (sbk == 1: a thread's row is contiguous): 128-bit loads, 8 instead of 32 per batch -- the
    // scalar version cost ~7 us of the ~19 us fixed time of a launch (tools/kbench.py with 128 rows)
    const bool vecw = p.sbk == 1 && (p.sbn & 3) == 0 && (p.K & 3) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0;
    for (int k0 = 0; k0 < p.KC * kChunkK; k0 += 32) {
      float w[32];
      if (vecw) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int k = k0 + j;
          const float4 v = (nv && k < p.K) ? __ldg(reinterpret_cast<const float4*>(wrow + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
          w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
        }
      } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int k = k0 + j;
        w[j] = (nv && k < p.K) ? __ldg(wrow + (int64_t)k * p.sbk) : 0.f;
      }
      }
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
    tc_fence_before();
    asm volatile("bar.arrive 2, 160;" ::: "memory");
  };

  const int64_t ntiles = (p.M + kTileM - 1) / kTileM;
  const int KC = p.KC;
  const int grp = (int)(blockIdx.x % (unsigned)p.ngroups);
  const int64_t cta = blockIdx.x / (unsigned)p.ngroups, ncta = gridDim.x / (unsigned)p.ngroups;
  const int pi_beg = grp * p.ppg, pi_end = pi_beg + p.ppg;

  if (warp == 0) {
    // =============================== MMA issuer ===============================
    const uint32_t idesc = make_idesc(kTileM, 0, 0);     // M = 128 channels, N = 128 rows of the tile
    const uint32_t sA_u = smem_u32(sA);
    const uint32_t leader = elect_one();
    uint32_t st = 0, ph = 0;   // smem ring position / phase
    uint32_t tcount = 0;       // tile counter (TMEM accumulator buffer)
    for (int pi = pi_beg; pi < pi_end; ++pi) {
    asm volatile("bar.sync 2, 160;" ::: "memory");        // this problem's weights are in tensor memory
    tc_fence_after();
    // the whole warp walks the loops (uniform control flow); one elected lane issues each MMA / commit
    for (int64_t tile = cta; tile < ntiles; tile += ncta, ++tcount) {
      const uint32_t buf = tcount & 1;
      mbar_wait(&tempty[buf], ((tcount >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + buf * 128;
      for (int kc = 0; kc < KC; ++kc) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const int kvalid = min(kChunkK, p.K - kc * kChunkK);
        const int ksteps = (kvalid + 7) >> 3;
        uint64_t dxh = make_desc(sA_u + st * 2 * kChunkBytes, 16, 1024);
        uint64_t dxl = make_desc(sA_u + st * 2 * kChunkBytes + kChunkBytes, 16, 1024);
        uint32_t w_hi = tmem_base + kTmemWHi + kc * kChunkK, w_lo = tmem_base + kTmemWLo + kc * kChunkK;
        if (p.single) {
          for (int ks = 0; ks < ksteps; ++ks, dxh += 2, w_hi += 8)
            umma_tf32_ts_w(leader, taddr, w_hi, dxh, idesc, (kc | ks) != 0);
        } else {
          for (int ks = 0; ks < ksteps; ++ks, dxh += 2, dxl += 2, w_hi += 8, w_lo += 8) {   // +32 bytes (>>4) along K
            umma_tf32_ts_w(leader, taddr, w_hi, dxh, idesc, (kc | ks) != 0);
            umma_tf32_ts_w(leader, taddr, w_lo, dxh, idesc, 1);
            umma_tf32_ts_w(leader, taddr, w_hi, dxl, idesc, 1);
          }
        }
        umma_commit_w(leader, &empty[st]);  // smem slot reusable once these MMAs retire
        if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
      }
      umma_commit_w(leader, &tfull[buf]);   // accumulator complete
    }
    __syncwarp();
    }
  } else if (warp <= kProducerWarps) {
    // =============================== producers ===============================
    if constexpr (PIECE != 0) {
      constexpr int PPR = 128 / PIECE;                       // pieces per 128-byte row segment (8 / 16)
      constexpr int NP = kTileM * PPR / kProducerThreads;    // pieces per thread per chunk (2 / 4)
      constexpr int RS = kProducerThreads / PPR;             // rows between a thread's pieces (64 / 32)
      constexpr int EPP = PIECE / 4;                         // fp32 per piece
      const int pt = threadIdx.x - 32;
      const int cp = pt % PPR, r0 = pt / PPR;
      const int kel = cp * EPP;                              // first k of the piece inside its chunk
      const int64_t my_tiles = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;
      const int64_t nchunk = my_tiles * KC;
      const uint32_t sA_u = smem_u32(sA);
      uint32_t soff[NP];
#pragma unroll
      for (int i = 0; i < NP; ++i) soff[i] = kmajor_off(r0 + RS * i, (cp * PIECE) >> 4) + (uint32_t)((cp * PIECE) & 15);
      uint32_t ist = 0, iph = 0;                             // issue cursor (ring stage, phase)
      uint32_t cst = 0;                                      // consume cursor
      const int64_t m_step = ncta * kTileM;
      for (int pi = pi_beg; pi < pi_end; ++pi) {
        const G1Prob& pr = p.prob[pi];
        const int64_t lda = pr.lda;
        const float* const a_thr = pr.A + (int64_t)r0 * lda + kel;
        const int64_t row_step = (int64_t)RS * lda;
        int64_t ld_m0 = cta * kTileM;
        int ld_kc = 0;
        auto issue = [&]() {
          const int64_t m0 = ld_m0;
          const int kc0 = ld_kc * kChunkK;
          if (++ld_kc == KC) { ld_kc = 0; ld_m0 += m_step; }
          mbar_wait(&empty[ist], iph ^ 1);                   // the MMAs that read this stage have retired
          const uint32_t base = sA_u + ist * 2 * kChunkBytes;
          const float* src = a_thr + m0 * lda + kc0;
          if (kc0 + kel + EPP <= p.K) {
            if (m0 + kTileM <= p.M) {
#pragma unroll
              for (int i = 0; i < NP; ++i) cp_async<PIECE>(base + soff[i], src + i * row_step);
            } else {                                         // rows past M: zeros (row-local, never stored)
#pragma unroll
              for (int i = 0; i < NP; ++i) {
                if (m0 + r0 + RS * i < p.M) cp_async<PIECE>(base + soff[i], src + i * row_step);
                else if constexpr (PIECE == 16) sts128(base + soff[i], make_uint4(0u, 0u, 0u, 0u));
                else sts64(base + soff[i], 0u, 0u);
              }
            }
          } else {                                           // k >= K inside the last k-step: zeros
#pragma unroll
            for (int i = 0; i < NP; ++i) {
              if constexpr (PIECE == 16) sts128(base + soff[i], make_uint4(0u, 0u, 0u, 0u));
              else sts64(base + soff[i], 0u, 0u);
            }
          }
          if (++ist == (uint32_t)S) { ist = 0; iph ^= 1; }
        };
        auto consume = [&]() {                               // the thread's own pieces of this stage have landed
          const uint32_t base = sA_u + cst * 2 * kChunkBytes;
#pragma unroll
          for (int i = 0; i < NP; ++i) {
            if (p.single) break;                               // one tf32 pass: no lo operand
            if constexpr (PIECE == 16) {
              const float4 v = lds128f(base + soff[i]);
              sts128(base + kChunkBytes + soff[i],
                     make_uint4(lo_of_raw(v.x), lo_of_raw(v.y), lo_of_raw(v.z), lo_of_raw(v.w)));
            } else {
              const float2 v = lds64f(base + soff[i]);
              sts64(base + kChunkBytes + soff[i], lo_of_raw(v.x), lo_of_raw(v.y));
            }
          }
          fence_proxy_async();
          mbar_arrive(&full[cst]);
          if (++cst == (uint32_t)S) cst = 0;
        };
        constexpr int AHEAD = 3;                             // chunks of copies in flight per thread (stages - 2)
#pragma unroll
        for (int u = 0; u < AHEAD; ++u) {
          if (u < nchunk) issue();
          cp_async_commit();
        }
        for (int64_t it = 0; it < nchunk; ++it) {
          if (it + AHEAD < nchunk) issue();
          cp_async_commit();
          cp_async_wait<AHEAD>();                            // the group of chunk `it` is complete
          consume();
        }
      }
    } else {
    // 512 threads; thread (c16, r0) moves the 16-byte column chunk c16 of rows r0 + 64 i.
    const int pt = threadIdx.x - 32;                 // 0..511
    const int c16 = pt & 7, r0 = pt >> 3;            // 8 threads cover one 128-byte row segment
    const int64_t my_tiles = cta < ntiles ? (ntiles - cta + ncta - 1) / ncta : 0;
    const int64_t nchunk = my_tiles * KC;
    const uint32_t sA_u = smem_u32(sA);
    constexpr int NV = kTileM * 8 / kProducerThreads;   // float4 per thread per chunk (2)
    constexpr int RS = kProducerThreads / 8;            // row stride between a thread's loads (64)
    uint32_t st = 0, ph = 0;
    for (int pi = pi_beg; pi < pi_end; ++pi) {
    const G1Prob& pr = p.prob[pi];
    const int64_t lda = pr.lda;
    const bool vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(pr.A) & 15) == 0);
    const float* const a_thr = pr.A + (int64_t)r0 * lda + c16 * 4;
    const int64_t row_step = (int64_t)RS * lda;
    uint32_t soff[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) soff[i] = kmajor_off(r0 + RS * i, c16);

    // load cursor (tile row offset, K chunk) and ring cursor advance incrementally: no div/mod
    int64_t ld_m0 = cta * kTileM;                     // restarts for every problem
    int ld_kc = 0;
    const int64_t m_step = ncta * kTileM;
    auto issue = [&](float4 (&v)[NV]) {
      const int64_t m0 = ld_m0;
      const int kc0 = ld_kc * kChunkK;
      if (++ld_kc == KC) { ld_kc = 0; ld_m0 += m_step; }
      const float* src = a_thr + m0 * lda + kc0;
      if (vec && m0 + kTileM <= p.M && kc0 + kChunkK <= p.K) {       // interior chunk: no guards
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(src + i * row_step));
      } else {
        const int kcol = kc0 + c16 * 4;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r0 + RS * i < p.M) {
            const float* q = src + i * row_step;
            if (vec && kcol + 3 < p.K) v[i] = __ldg(reinterpret_cast<const float4*>(q));
            else {
              if (kcol < p.K) v[i].x = __ldg(q);
              if (kcol + 1 < p.K) v[i].y = __ldg(q + 1);
              if (kcol + 2 < p.K) v[i].z = __ldg(q + 2);
              if (kcol + 3 < p.K) v[i].w = __ldg(q + 3);
            }
          }
        }
      }
    };
    auto commit = [&](const float4 (&v)[NV]) {
      mbar_wait(&empty[st], ph ^ 1);
      const uint32_t base = sA_u + st * 2 * kChunkBytes;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint4 hi, lo;
        split4(v[i], hi, lo);
        sts128(base + soff[i], hi);
        if (!p.single) sts128(base + kChunkBytes + soff[i], lo);
      }
      fence_proxy_async();
      mbar_arrive(&full[st]);
      if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
    };

    // kG1Flight chunks of loads in flight per thread (register ring, statically indexed)
    float4 v[kG1Flight][NV];
#pragma unroll
    for (int u = 0; u < kG1Flight - 1; ++u)
      if (u < nchunk) issue(v[u]);
    for (int64_t it = 0; it < nchunk; it += kG1Flight) {
#pragma unroll
      for (int u = 0; u < kG1Flight; ++u) {
        if (it + u < nchunk) {
          if (it + u + kG1Flight - 1 < nchunk) issue(v[(u + kG1Flight - 1) % kG1Flight]);
          commit(v[u]);
        }
      }
    }
    }
    }
  } else {
    // =============================== epilogue ===============================
    // The accumulator is C^T: lane = output channel n = 32 q + lane, column = row of the tile.  Warp
    // (q, half) stores the 16-column blocks c0 = 16 half + 32 j: for each column the 32 lanes write 32
    // consecutive floats of one output row (one 128-byte line per store instruction).
    const int ew = warp - (1 + kProducerWarps);      // 0..7
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int half = ew >> 2;
    const int n = q * 32 + lane;
    const bool nvalid = n < p.N;
    uint32_t tcount = 0;
    for (int pi = pi_beg; pi < pi_end; ++pi) {
    const G1Prob& pr = p.prob[pi];
    // every epilogue warp has waited for the previous problem's last accumulator, i.e. all MMAs that
    // read the old weights have retired: the loader warps may overwrite them
    if (ew < 4) load_weights(pr.W);
    const float bv = (pr.bias && nvalid) ? __ldg(pr.bias + n) : 0.f;
    for (int64_t tile = cta; tile < ntiles; tile += ncta, ++tcount) {
      const uint32_t buf = tcount & 1;
      mbar_wait(&tfull[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const int64_t row0 = tile * kTileM;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128;
      for (int c0 = half * 16; c0 < kTileM; c0 += 16 * kEpiHalves) {
        if (row0 + c0 >= p.M) break;                   // the rest of a partial last tile
        float v[16];
        tmem_ld16(taddr + c0, v);
        if (nvalid) {
          float* dst = pr.C + (row0 + c0) * pr.ldc + n;
          if (row0 + c0 + 16 <= p.M) {
            if (pr.beta) {
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] += dst[(int64_t)j * pr.ldc];
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) dst[(int64_t)j * pr.ldc] = v[j] + bv;
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (row0 + c0 + j < p.M) {
                float o = v[j] + bv;
                if (pr.beta) o += dst[(int64_t)j * pr.ldc];
                dst[(int64_t)j * pr.ldc] = o;
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
    }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ G2 (wgrad)
struct G2Prob {
  const float* Y;        // [rows, >=128] : the 128 columns starting at Y are the output rows (Dm)
  int64_t ldy;
  const float* X;        // [rows, N]
  int64_t ldx;
  float* partial;        // [ctas per problem][128][N]
  float* colsum;         // [ctas per problem][128] or NULL
};
// Up to kMaxProbG2 weight gradients of identical shape (the four node-level ones: dW_q, dW_k, dW_v,
// dW_skip; the deferred weight gradients of a training step, x2_tc_wgrad_batch) share one launch: CTA b
// works on problem b % nprob, row slab b / nprob.
struct G2Params {
  G2Prob prob[kMaxProbG2];
  int nprob;
  int N, N_pad;          // N_pad multiple of 32, <= 128
  int64_t rows, rows_per_cta;   // rows_per_cta multiple of 32
  int stages;
  int single;            // 1: one tf32 pass (hi . hi only): X2_MODE_TF32
};

// Byte offset of (k-row r in [0,32), 16-byte chunk c16 along MN) inside an MN-major operand chunk.
// tf32 MN-major operands must use SWIZZLE_128B_BASE32B (cutlass sm100_common.inl: "the only available
// smem layout"): MN blocks of 32 fp32 (LBO = 4096 B apart = 8 k-groups), k-groups of 4 rows (SBO =
// 512 B), 128 B per row, 32-byte chunks XOR-swizzled by (row & 3).  Verified word by word on
// hardware with tools/umma_probe.cu.
__device__ __forceinline__ uint32_t mnmajor_off(int r, int c16) {
  const int c32 = (c16 & 7) >> 1;
  return (uint32_t)((c16 >> 3) * 4096 + (r >> 2) * 512 + (r & 3) * 128 + ((c32 ^ (r & 3)) << 5) + ((c16 & 1) << 4));
}

constexpr int kG2YCol = 128;         // TMEM columns [128, 128 + 64 S): per stage Y^T hi (32 cols) | lo (32 cols)
// The tensor core adds into its fp32 accumulator with TRUNCATION: every tcgen05.mma loses ~half an ulp of the
// running sum, always towards zero.  Measured on dW = Y^T X with N(0,1) operands (tools/wgrad_err.py): the result
// shrinks by 6e-9 per accumulated row -- 3e-6 at 346 rows per CTA, 8.9e-6 at 1 350, 4.2e-5 at the bench batch's
// 5 485 (cuBLAS fp32: 2.5e-6) -- a bias, not noise, and above the 1e-5 parity bar.  So the accumulation is cut
// into PERIODS of kG2Period chunks (256 rows: 96 MMAs) that alternate between two accumulators (columns
// [0, 128) and [384, 512)); a finished period is added, round-to-nearest, to fp32 running sums held in the
// registers of the producer threads (thread (q, g, lane): channel 32 q + lane, columns 32 g .. 32 g + 31), which
// also write the CTA's partial tile at the end.  No extra barrier: a producer that has passed the `empty` wait
// for chunk i knows chunk i - S has retired, hence period p once i = (p + 1) F - 1 + S; and the MMA warp cannot
// reach period p + 2 (same accumulator) before every producer has published chunk (p + 2) F > i.  Needs F > S.
constexpr int kG2Period = 8;
constexpr int kG2PeriodMinRows = 768;  // CTAs with fewer rows keep the single accumulator (bias < 5e-6)
constexpr int kG2Acc1 = 384;         // second accumulator (4 stages: Y^T operands end at column 384)

__device__ __forceinline__ void tmem_st8f(uint32_t taddr, const float (&v)[8], bool lo) {
  uint32_t r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = lo ? __float_as_uint(lo_part(v[j])) : hi_bits(v[j]);
  tmem_st8(taddr, r);
}

// The dY^T operand (M = 128 output channels x K = rows) is the MMA's A operand and lives in TENSOR
// MEMORY: producer thread (q, g, lane) owns channel d = 32 q + lane (its TMEM lane) and the 8 rows
// 8g..8g+7 of every 32-row chunk, loads them coalesced across lanes and writes them with tcgen05.st.
// Only X goes through shared memory (MN-major B operand), halving the smem traffic of the first
// version, which was bound by shared-memory bandwidth (SS-mode MMAs re-read both operands 3 times).
// warp 0 = MMA issuer, then epilogue of lane quarter 0; warps 1..16 producers; warps 17..19 epilogue of
// quarters 1..3.  20 warps: the register file is allocated in groups of 4 warps, so 21 would cap a
// thread at 80 registers instead of 96 (three chunks of loads in flight need ~90).
constexpr int kG2Threads = 32 + kProducerThreads + 3 * 32;
// XMODE selects the producer implementation:
//   0  generic: loads staged through registers with per-element guards (any N, ldx, alignment)
//   1  cp.async, N == 128, 16-byte aligned X rows
//   2  cp.async, ldx == N <= 64 and even (e.g. sbf [T, 42]): 8-byte pieces
// ncu on the generic version: 298 warp-instructions per producer warp per 32-row chunk, 41 % of the stall
// samples on the first use of the loaded registers (3.9 TB/s); the cp.async paths hold no registers for
// data in flight and compute only the lo halves (the hi operands are the raw words).
// PERIODS: the accumulation periods of kG2Period (rows_per_cta above kG2PeriodMinRows); false = one accumulator for
// the CTA's whole range, read out by the epilogue warps (short ranges: the bias is below 3e-6 and the 32 running
// sums per producer thread would only cost registers).
template <int XMODE, bool PERIODS>
__global__ void __launch_bounds__(kG2Threads, 1) k_tc_wgrad(const G2Params p) {
  const G2Prob& pr = p.prob[blockIdx.x % (unsigned)p.nprob];
  const int64_t cta = blockIdx.x / (unsigned)p.nprob;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  const uint32_t xbytes = (uint32_t)(p.N_pad / 32) * 4096;      // one X half (hi or lo)
  const uint32_t stage_bytes = 2 * xbytes + (XMODE != 0 ? 16384u : 0u);   // X hi | X lo (| raw Y block)
  uint8_t* sS = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + (size_t)S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  float* cs_smem = reinterpret_cast<float*>(bars + 2 * S + 2);   // [4][128] column-sum staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], kProducerThreads);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base_v = *tmem_slot;
  const uint32_t tmem_base = tmem_base_v;
  pdl_sync();        // global memory from here on

  const int64_t rbeg = cta * p.rows_per_cta;
  const int64_t rend = min(p.rows, rbeg + p.rows_per_cta);
  const int64_t nchunks = rend > rbeg ? (rend - rbeg + kChunkK - 1) / kChunkK : 0;

  if (warp == 0) {
    if (nchunks > 0) {                                            // whole warp, uniform control flow (see umma_tf32_ts_w)
      const uint32_t idesc = make_idesc(p.N_pad, 0, 1);          // A from TMEM, B MN-major
      const uint32_t s_u = smem_u32(sS);
      const uint32_t leader = elect_one();
      // tensor-memory addresses derived from the value loaded from shared memory went through an R2UR.BROADCAST per
      // MMA operand (4 per tcgen05.mma in the SASS); a redux result is uniform for the assembler
      const uint32_t tmem_base = __reduce_max_sync(0xffffffffu, tmem_base_v);
      uint32_t st = 0, ph = 0;
      for (int64_t c = 0; c < nchunks; ++c) {
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t x_hi = s_u + st * stage_bytes, x_lo = x_hi + xbytes;
        uint32_t y_hi = tmem_base + kG2YCol + st * 64, y_lo = y_hi + 32;
        const int kvalid = (int)min((int64_t)kChunkK, rend - rbeg - c * kChunkK);
        const int ksteps = (kvalid + 7) >> 3;
        uint64_t dxh = make_desc(x_hi, 4096, 512, kLayoutSW128Base32);
        uint64_t dxl = make_desc(x_lo, 4096, 512, kLayoutSW128Base32);
        // one K=8 step = two 4-row k-groups = 1024 B (>>4 = 64)
        const uint32_t cp = PERIODS ? (uint32_t)(c % kG2Period) : (uint32_t)(c != 0);   // 0: first chunk of a period
        const uint32_t acc = tmem_base + ((PERIODS && ((c / kG2Period) & 1)) ? (uint32_t)kG2Acc1 : 0u);
        if (p.single) {
          for (int ks = 0; ks < ksteps; ++ks, dxh += 64, y_hi += 8)
            umma_tf32_ts_w(leader, acc, y_hi, dxh, idesc, (cp | (uint32_t)ks) != 0);
        } else {
          for (int ks = 0; ks < ksteps; ++ks, dxh += 64, dxl += 64, y_hi += 8, y_lo += 8) {
            umma_tf32_ts_w(leader, acc, y_hi, dxh, idesc, (cp | (uint32_t)ks) != 0);
            umma_tf32_ts_w(leader, acc, y_lo, dxh, idesc, 1);
            umma_tf32_ts_w(leader, acc, y_hi, dxl, idesc, 1);
          }
        }
        umma_commit_w(leader, &empty[st]);
        if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
      }
      umma_commit_w(leader, tfull);
    }
    __syncwarp();
  } else if (warp <= kProducerWarps) {
    const int pt = threadIdx.x - 32;                 // 0..511
    const int q = warp & 3;                          // TMEM lane quarter of this warp
    const int g = (warp - 1) >> 2;                   // 8-row group of the chunk (0..3)
    const int d = q * 32 + lane;                     // output channel owned by this thread
    const uint32_t s_u = smem_u32(sS);
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + kG2YCol + 8 * g;
    float cs = 0.f;
    // running sums of the finished accumulation periods (see kG2Period): channel d, columns 32 g .. 32 g + 31
    float racc[PERIODS ? 32 : 1];
#pragma unroll
    for (int j = 0; j < (PERIODS ? 32 : 1); ++j) racc[j] = 0.f;
    int64_t drained = 0;                             // periods added to racc so far
    const bool my_cols = 32 * g < p.N_pad;           // warp-uniform
    auto drain = [&](int64_t period) {               // whole warp (tcgen05.ld is .sync.aligned)
      if constexpr (PERIODS) if (my_cols) {
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + ((period & 1) ? (uint32_t)kG2Acc1 : 0u) + 32 * g;
        float v[16];
        tmem_ld16(ta, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) racc[j] += v[j];
        tmem_ld16(ta + 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) racc[16 + j] += v[j];
      }
    };
    // called after the `empty` wait that precedes writing chunk i: chunk i - S has retired
    auto drain_if_due = [&](int64_t i) {
      const int64_t r = i - S + 1;                   // chunks known to be retired
      if (PERIODS && r > 0 && r % kG2Period == 0 && r / kG2Period > drained) {
        drain(drained);
        ++drained;
      }
    };
    if constexpr (XMODE != 0) {
      // ---- cp.async producers: raw fp32 goes straight into the X hi operand (MN-major layout) and into a
      // per-stage raw Y block; the hi operands are the raw words (the tensor core truncates them), only the
      // lo halves are computed.  Warp (q, g) owns the 8-row x 32-channel Y block it later feeds to tensor
      // memory, so a __syncwarp() is the only synchronisation between the copies and their readers; X pieces
      // are read back by the thread that copied them.
      const bool want_cs = pr.colsum != nullptr;
      const int64_t ldy = pr.ldy;
      const uint32_t yraw_off = 2 * xbytes;          // stage: X hi | X lo | Y raw [32][128] fp32
      // Y pieces of this lane: piece pc = lane + 32 i -> row 8g + (pc >> 3), 16-byte column pc & 7 of the block
      uint32_t ydst[2];
      const float* ysrc[2];
      int yrow[2];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int pc = lane + 32 * i;
        yrow[i] = 8 * g + (pc >> 3);
        ydst[i] = yraw_off + (uint32_t)(yrow[i] * 128 + q * 32 + (pc & 7) * 4) * 4;
        ysrc[i] = pr.Y + (rbeg + yrow[i]) * ldy + q * 32 + (pc & 7) * 4;
      }
      const uint32_t yrd = yraw_off + (uint32_t)((8 * g) * 128 + d) * 4;     // word read for row 8g + j: + j * 512
      // X pieces of this thread
      constexpr int XP = 2;                          // pieces per thread per chunk
      uint32_t xdst[XP];
      const float* xsrc[XP];
      int xrow[XP];
      bool xact[XP];
#pragma unroll
      for (int i = 0; i < XP; ++i) {
        const int f = pt + kProducerThreads * i;
        if constexpr (XMODE == 1) {                  // N == 128: 16-byte pieces, row f >> 5, float4 column f & 31
          xrow[i] = f >> 5;
          xact[i] = true;
          xdst[i] = mnmajor_off(xrow[i], f & 31);
          xsrc[i] = pr.X + (rbeg + xrow[i]) * pr.ldx + (f & 31) * 4;
        } else {                                     // ldx == N even: 8-byte pieces, N / 2 per row
          const int hp = p.N >> 1;
          xrow[i] = f / hp;
          const int col = (f - xrow[i] * hp) * 2;
          xact[i] = f < kChunkK * hp;
          xdst[i] = mnmajor_off(xrow[i] & 31, col >> 2) + (uint32_t)(col & 3) * 4;
          xsrc[i] = pr.X + (rbeg + xrow[i]) * (int64_t)p.N + col;
        }
      }
      const int64_t ystep = (int64_t)kChunkK * ldy;
      const int64_t xstep = XMODE == 1 ? (int64_t)kChunkK * pr.ldx : (int64_t)kChunkK * p.N;
      uint32_t ist = 0, iph = 0, cst = 0;
      int64_t ic = 0;                                // next chunk to issue
      auto issue = [&]() {
        const int64_t row0 = rbeg + ic * kChunkK;
        mbar_wait(&empty[ist], iph ^ 1);             // MMAs of the chunk that used this stage have retired
        tc_fence_after();
        ++ic;
        const uint32_t sb = s_u + ist * stage_bytes;
        if (row0 + kChunkK <= rend) {
#pragma unroll
          for (int i = 0; i < 2; ++i) cp_async<16>(sb + ydst[i], ysrc[i]);
#pragma unroll
          for (int i = 0; i < XP; ++i)
            if (xact[i]) cp_async<XMODE == 1 ? 16 : 8>(sb + xdst[i], xsrc[i]);
        } else {                                     // last, partial chunk: rows past the end are zeros
#pragma unroll
          for (int i = 0; i < 2; ++i) {
            if (row0 + yrow[i] < rend) cp_async<16>(sb + ydst[i], ysrc[i]);
            else sts128(sb + ydst[i], make_uint4(0u, 0u, 0u, 0u));
          }
#pragma unroll
          for (int i = 0; i < XP; ++i) {
            if (!xact[i]) continue;
            if (row0 + xrow[i] < rend) cp_async<XMODE == 1 ? 16 : 8>(sb + xdst[i], xsrc[i]);
            else if constexpr (XMODE == 1) sts128(sb + xdst[i], make_uint4(0u, 0u, 0u, 0u));
            else sts64(sb + xdst[i], 0u, 0u);
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) ysrc[i] += ystep;
#pragma unroll
        for (int i = 0; i < XP; ++i) xsrc[i] += xstep;
        if (++ist == (uint32_t)S) { ist = 0; iph ^= 1; }
      };
      auto consume = [&]() {
        const uint32_t sb = s_u + cst * stage_bytes;
        __syncwarp();                                // the warp's Y block: every lane's copies have landed
        uint32_t yh[8], yl[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = lds32f(sb + yrd + j * 512);
          if (want_cs) cs += v;                      // fixed order per thread
          yh[j] = __float_as_uint(v);
          yl[j] = lo_of_raw(v);
        }
        tmem_st8(trow + cst * 64, yh);
        if (!p.single) tmem_st8(trow + cst * 64 + 32, yl);
#pragma unroll
        for (int i = 0; i < XP; ++i) {
          if (!xact[i] || p.single) continue;
          if constexpr (XMODE == 1) {
            const float4 v = lds128f(sb + xdst[i]);
            sts128(sb + xbytes + xdst[i], make_uint4(lo_of_raw(v.x), lo_of_raw(v.y), lo_of_raw(v.z), lo_of_raw(v.w)));
          } else {
            const float2 v = lds64f(sb + xdst[i]);
            sts64(sb + xbytes + xdst[i], lo_of_raw(v.x), lo_of_raw(v.y));
          }
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        tc_fence_before();
        fence_proxy_async();
        mbar_arrive(&full[cst]);
        if (++cst == (uint32_t)S) cst = 0;
      };
      constexpr int AHEAD = 2;                       // chunks of copies in flight per thread (stages - 2); depths
                                                     // 1..4 and 3..6 stages measured the same (on-chip bound)
#pragma unroll
      for (int u = 0; u < AHEAD; ++u) {
        if (u < nchunks) issue();
        cp_async_commit();
      }
      for (int64_t c = 0; c < nchunks; ++c) {
        if (c + AHEAD < nchunks) {
          issue();
          cp_async_commit();
          drain_if_due(ic - 1);                      // after the copies are in flight: the drain overlaps them
        } else {
          cp_async_commit();
        }
        cp_async_wait<AHEAD>();                      // this thread's copies of chunk c are complete
        consume();
      }
    } else {
    const bool vecX = ((pr.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(pr.X) & 15) == 0) && ((p.N & 3) == 0);
    const int xq = p.N_pad / 4;                      // float4 per X row (8..32)
    const int nx = kChunkK * xq;                     // float4 per X chunk (256..1024)
    const float* const y_thr = pr.Y + (int64_t)(8 * g) * pr.ldy + d;
    const int64_t ldy = pr.ldy;
    constexpr int NV = kChunkK * 32 / kProducerThreads;   // X float4 per thread per chunk (2)
    uint32_t xoff[NV];
    int xr[NV], xcol[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int f = pt + i * kProducerThreads;
      xr[i] = f / xq;
      const int xc = f - xr[i] * xq;
      xcol[i] = xc * 4;
      xoff[i] = mnmajor_off(xr[i] & 31, xc);
    }

    auto issue = [&](float (&vy)[8], float4 (&vx)[NV], int64_t c) {
      const int64_t row0 = rbeg + c * kChunkK;
      const float* ysrc = y_thr + row0 * ldy;
      if (row0 + kChunkK <= rend) {
#pragma unroll
        for (int j = 0; j < 8; ++j) vy[j] = __ldg(ysrc + (int64_t)j * ldy);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) vy[j] = (row0 + 8 * g + j < rend) ? __ldg(ysrc + (int64_t)j * ldy) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        vx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pt + i * kProducerThreads < nx) {
          const int64_t row = row0 + xr[i];
          if (row < rend) {
            const float* qx = pr.X + row * pr.ldx + xcol[i];
            if (vecX && xcol[i] + 3 < p.N) vx[i] = __ldg(reinterpret_cast<const float4*>(qx));
            else {
              if (xcol[i] < p.N) vx[i].x = __ldg(qx);
              if (xcol[i] + 1 < p.N) vx[i].y = __ldg(qx + 1);
              if (xcol[i] + 2 < p.N) vx[i].z = __ldg(qx + 2);
              if (xcol[i] + 3 < p.N) vx[i].w = __ldg(qx + 3);
            }
          }
        }
      }
    };
    uint32_t st = 0, ph = 0;
    int64_t cc = 0;                                  // chunk being written
    auto commit = [&](const float (&vy)[8], const float4 (&vx)[NV]) {
      mbar_wait(&empty[st], ph ^ 1);
      tc_fence_after();
      drain_if_due(cc);
      ++cc;
#pragma unroll
      for (int j = 0; j < 8; ++j) cs += vy[j];                   // fixed order per thread
      tmem_st8f(trow + st * 64, vy, false);
      tmem_st8f(trow + st * 64 + 32, vy, true);
      const uint32_t xb = s_u + st * stage_bytes;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (pt + i * kProducerThreads < nx) {
          uint4 hi, lo;
          split4(vx[i], hi, lo);
          sts128(xb + xoff[i], hi);
          sts128(xb + xbytes + xoff[i], lo);
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      fence_proxy_async();
      mbar_arrive(&full[st]);
      if (++st == (uint32_t)S) { st = 0; ph ^= 1; }
    };

    // three chunks of loads in flight per thread
    float ya[8], yb[8], yc[8];
    float4 xa[NV], xb_[NV], xc[NV];
    if (nchunks > 0) issue(ya, xa, 0);
    if (nchunks > 1) issue(yb, xb_, 1);
    for (int64_t c = 0; c < nchunks; c += 3) {
      if (c + 2 < nchunks) issue(yc, xc, c + 2);
      commit(ya, xa);
      if (c + 1 < nchunks) {
        if (c + 3 < nchunks) issue(ya, xa, c + 3);
        commit(yb, xb_);
      }
      if (c + 2 < nchunks) {
        if (c + 4 < nchunks) issue(yb, xb_, c + 4);
        commit(yc, xc);
      }
    }
    }
    // the periods still in tensor memory (at most two, in different accumulators), then this thread's slice of
    // the CTA's partial tile
    if constexpr (PERIODS) {
      if (nchunks > 0) {
        mbar_wait(tfull, 0);
        tc_fence_after();
        const int64_t nper = (nchunks + kG2Period - 1) / kG2Period;
        for (; drained < nper; ++drained) drain(drained);
      }
      if (my_cols) {
        float* dst = pr.partial + (cta * 128 + d) * p.N + 32 * g;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (32 * g + j < p.N) dst[j] = racc[j];
      }
    }
    // column sums of Y over this CTA's rows: combine the 4 row groups in fixed order
    if (pr.colsum) {
      cs_smem[g * 128 + d] = cs;
      asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");
      if (g == 0) pr.colsum[cta * 128 + d] = (cs_smem[d] + cs_smem[128 + d]) + (cs_smem[256 + d] + cs_smem[384 + d]);
    }
  }
  if constexpr (!PERIODS) {
  if (warp == 0 || warp > kProducerWarps) {
    // =============================== epilogue: accumulator -> this CTA's partial tile
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float* dst = pr.partial + (cta * 128 + m) * p.N;
    if (nchunks > 0) {
      mbar_wait(tfull, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
      for (int c0 = 0; c0 < p.N_pad; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c0 + j < p.N) dst[c0 + j] = v[j];
      }
    } else {
      for (int j = 0; j < p.N; ++j) dst[j] = 0.f;
    }
  }
  }   // (PERIODS: warps 17..19 have nothing left to do, the producers drain the accumulators)
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host launchers
static inline int ceil_to(int v, int m) { return (v + m - 1) / m * m; }

// kept for ABI compatibility of the workspace queries: the weights now live in tensor memory
static inline size_t bimage_bytes(int K, int N) { (void)K; (void)N; return 256; }

static inline bool g1_supported(int K, int N) { return K >= 1 && K <= 128 && N >= 1 && N <= 128; }

static inline bool cp_async_enabled() {         // X2GNN_CPASYNC=0 forces the register-staged producers (A/B runs)
  static const int on = [] { const char* v = getenv("X2GNN_CPASYNC"); return (v && v[0] == '0') ? 0 : 1; }();
  return on != 0;
}
// nprob problems C_i[M,N] (+)= A_i[M,K] . B_i(K,N) + bias_i, B_i(k,n) = W_i[k*sbk + n*sbn], in one launch.
// The problems are split into `ngroups` groups of nprob / ngroups consecutive problems (see G1Params).
static int tc_gemm_batch(const G1Prob* probs, int nprob, int ngroups, int64_t M, int K, int64_t sbk, int64_t sbn,
                         int N, cudaStream_t st, int single = 0) {
  if (M <= 0 || nprob <= 0) return X2_OK;
  if (!g1_supported(K, N) || nprob > kMaxProb || ngroups < 1 || nprob % ngroups != 0) {
    set_error("tc_gemm: unsupported K=%d N=%d nprob=%d ngroups=%d", K, N, nprob, ngroups);
    return X2_EINVAL;
  }
  const int stages = 5;                                   // 5 x 32 KB = 160 KB (6 and 7 stages measured the same)
  const size_t smem = 1024 + (size_t)stages * 2 * kChunkBytes + 256;
  X2_DYN_SMEM(k_tc_gemm<0>, kMaxSmem);
  X2_DYN_SMEM(k_tc_gemm<8>, kMaxSmem);
  X2_DYN_SMEM(k_tc_gemm<16>, kMaxSmem);
  // producer path: the widest cp.async piece every problem's rows are aligned to (0 = register staging)
  int piece = cp_async_enabled() ? 16 : 0;
  for (int i = 0; i < nprob && piece; ++i) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(probs[i].A);
    while (piece && ((a % piece) || ((probs[i].lda * 4) % piece) || ((K * 4) % piece))) piece = piece == 16 ? 8 : 0;
  }
  G1Params p{};
  for (int i = 0; i < nprob; ++i) p.prob[i] = probs[i];
  p.nprob = nprob; p.ngroups = ngroups; p.ppg = nprob / ngroups;
  p.M = M; p.K = K; p.KC = (K + kChunkK - 1) / kChunkK;
  p.sbk = sbk; p.sbn = sbn; p.N = N; p.stages = stages;
  p.single = single;
  const int64_t ntiles = cdiv(M, kTileM);
  const int64_t per_group = kNumSM / ngroups;
  const int grid = (int)(ntiles < per_group ? ntiles : per_group) * ngroups;
  if (piece == 16) launch_k(k_tc_gemm<16>, dim3(grid), dim3(kThreads), smem, st, p);
  else if (piece == 8) launch_k(k_tc_gemm<8>, dim3(grid), dim3(kThreads), smem, st, p);
  else launch_k(k_tc_gemm<0>, dim3(grid), dim3(kThreads), smem, st, p);
  X2_LAUNCH_OK();
  return X2_OK;
}

// C[M,N] (+)= A[M,K] . B(K,N) + bias, B(k,n) = W[k*sbk + n*sbn].
static int tc_gemm(const float* A, int64_t lda, int64_t M, int K, const float* W, int64_t sbk, int64_t sbn,
                   int N, const float* bias, float* C, int64_t ldc, int beta, void* img, cudaStream_t st,
                   int single = 0) {
  (void)img;
  G1Prob pr{A, lda, W, bias, C, ldc, beta};
  return tc_gemm_batch(&pr, 1, 1, M, K, sbk, sbn, N, st, single);
}

static inline int wgrad_ctas(int64_t rows, int nprob = 1) {
  int64_t c = cdiv(rows > 0 ? rows : 1, 256);     // >= 256 rows per CTA
  if (c > kNumSM / nprob) c = kNumSM / nprob;
  if (c < 1) c = 1;
  return (int)c;
}
// sized for the largest grid any batch may use (<= kNumSM CTAs in total)
static inline size_t tc_wgrad_workspace_floats(int64_t rows, int N) {
  (void)rows;
  return (size_t)kNumSM * (128 * (size_t)N + 128) + 64;
}

struct G2Job {           // one weight gradient: dW[128,N] = Y^T X ; db[128] = colsum(Y) (db may be NULL)
  const float* Y;
  int64_t ldy;
  const float* X;
  int64_t ldx;
  float* dW;
  int64_t lddw;
  float* db;
};

struct ReduceBatch {
  const float* partial[kMaxProbG2];
  const float* colsum[kMaxProbG2];
  float* out[kMaxProbG2];
  float* bias[kMaxProbG2];
  int64_t ldo[kMaxProbG2];
};

// Fixed-order reduction of the per-CTA partial tiles of up to kMaxProbG2 weight gradients
// (blockIdx.y = problem).  Same summation tree as k_splitk_reduce: warp g sums the splits g, g+8, ...
// and the eight warp sums are added in order.
__global__ void __launch_bounds__(256)
k_splitk_reduce_batch(const ReduceBatch b, int splits, int64_t M, int N) {
  __shared__ float red[8][33];
  pdl_sync();
  const int pi = blockIdx.y;
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t MN = M * N;
  const int64_t nblk_out = (MN + 31) / 32;
  const bool is_bias = (int64_t)blockIdx.x >= nblk_out;      // trailing blocks reduce the column sums
  const int64_t idx = is_bias ? ((int64_t)blockIdx.x - nblk_out) * 32 + lane : (int64_t)blockIdx.x * 32 + lane;
  const int64_t lim = is_bias ? M : MN;
  const float* src = is_bias ? b.colsum[pi] : b.partial[pi];
  const bool live = idx < lim && src != nullptr;
  float s = 0.f;
  if (live) {
    int z = g;                                     // loads eight at a time, additions in the original order
    for (; z + 56 < splits; z += 64) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(int64_t)(z + 8 * u) * lim + idx];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; z < splits; z += 8) s += src[(int64_t)z * lim + idx];
  }
  red[g][lane] = s;
  __syncthreads();
  if (g == 0 && live) {
    float t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][lane];
    if (is_bias) b.bias[pi][idx] = t;
    else {
      const int64_t m = idx / N;
      b.out[pi][m * b.ldo[pi] + (idx - m * N)] = t;
    }
  }
}

// nprob weight gradients over the same `rows` and the same N in one launch + one reduction launch.
// ws: tc_wgrad_workspace_floats
static int tc_wgrad_batch(const G2Job* jobs, int nprob, int64_t rows, int N, float* ws, cudaStream_t st,
                          int single = 0) {
  if (N < 1 || N > 128 || nprob < 1 || nprob > kMaxProbG2) { set_error("tc_wgrad: unsupported N=%d nprob=%d", N, nprob); return X2_EINVAL; }
  const int N_pad = ceil_to(N, 32);
  const int cpp = wgrad_ctas(rows, nprob);                              // CTAs per problem
  const int64_t rpc = cdiv(cdiv(rows > 0 ? rows : 1, cpp), kChunkK) * kChunkK;
  // producer path (see k_tc_wgrad): every job of the batch must qualify
  int xmode = cp_async_enabled() ? (N == 128 ? 1 : ((N <= 64 && (N & 1) == 0) ? 2 : 0)) : 0;
  for (int i = 0; i < nprob; ++i) {
    const uintptr_t xa = reinterpret_cast<uintptr_t>(jobs[i].X), ya = reinterpret_cast<uintptr_t>(jobs[i].Y);
    if ((ya & 15) || (jobs[i].ldy & 3)) xmode = 0;
    if (xmode == 1 && ((xa & 15) || (jobs[i].ldx & 3))) xmode = 0;
    if (xmode == 2 && ((xa & 7) || jobs[i].ldx != N)) xmode = 0;
  }
  // stage: X hi | X lo (| raw Y block, cp.async paths); Y^T hi / lo live in tensor memory
  const uint32_t stage_bytes = 2 * (uint32_t)(N_pad / 32) * 4096 + (xmode ? 16384u : 0u);
  const int stages = 4;                                                 // <= 192 KB smem, 4 x 64 TMEM columns
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 256 + 4 * 128 * sizeof(float);
  X2_DYN_SMEM((k_tc_wgrad<0, false>), kMaxSmem);
  X2_DYN_SMEM((k_tc_wgrad<1, false>), kMaxSmem);
  X2_DYN_SMEM((k_tc_wgrad<2, false>), kMaxSmem);
  X2_DYN_SMEM((k_tc_wgrad<0, true>), kMaxSmem);
  X2_DYN_SMEM((k_tc_wgrad<1, true>), kMaxSmem);
  X2_DYN_SMEM((k_tc_wgrad<2, true>), kMaxSmem);
  G2Params p{};
  ReduceBatch rb{};
  bool any_bias = false;
  const size_t per_prob = (size_t)cpp * (128 * (size_t)N + 128);
  for (int i = 0; i < nprob; ++i) {
    float* base = ws + (size_t)i * per_prob;
    p.prob[i] = G2Prob{jobs[i].Y, jobs[i].ldy, jobs[i].X, jobs[i].ldx, base,
                       jobs[i].db ? base + (size_t)cpp * 128 * N : nullptr};
    rb.partial[i] = p.prob[i].partial; rb.colsum[i] = p.prob[i].colsum;
    rb.out[i] = jobs[i].dW; rb.ldo[i] = jobs[i].lddw; rb.bias[i] = jobs[i].db;
    any_bias |= jobs[i].db != nullptr;
  }
  p.nprob = nprob; p.N = N; p.N_pad = N_pad; p.rows = rows; p.rows_per_cta = rpc; p.stages = stages;
  p.single = single;
  const bool periods = rpc > kG2PeriodMinRows;
  const dim3 grid(cpp * nprob), blk(kG2Threads);
  if (periods) {
    if (xmode == 1) launch_k(k_tc_wgrad<1, true>, grid, blk, smem, st, p);
    else if (xmode == 2) launch_k(k_tc_wgrad<2, true>, grid, blk, smem, st, p);
    else launch_k(k_tc_wgrad<0, true>, grid, blk, smem, st, p);
  } else {
    if (xmode == 1) launch_k(k_tc_wgrad<1, false>, grid, blk, smem, st, p);
    else if (xmode == 2) launch_k(k_tc_wgrad<2, false>, grid, blk, smem, st, p);
    else launch_k(k_tc_wgrad<0, false>, grid, blk, smem, st, p);
  }
  X2_LAUNCH_OK();
  const unsigned nblk = (unsigned)((128 * (int64_t)N + 31) / 32 + (any_bias ? 4 : 0));
  launch_k(k_splitk_reduce_batch, dim3(nblk, (unsigned)nprob), dim3(256), 0, st, rb, cpp, (int64_t)128, N);
  X2_LAUNCH_OK();
  return X2_OK;
}

// dW[128,N] = Y[rows,128]^T X[rows,N] ; db[128] = colsum(Y)  (db may be NULL).  ws: tc_wgrad_workspace_floats
static int tc_wgrad(const float* Y, int64_t ldy, const float* X, int64_t ldx, int64_t rows, int N, float* dW,
                    int64_t lddw, float* db, float* ws, cudaStream_t st, int single = 0) {
  const G2Job job{Y, ldy, X, ldx, dW, lddw, db};
  return tc_wgrad_batch(&job, 1, rows, N, ws, st, single);
}

}  // namespace tc
}  // namespace x2
