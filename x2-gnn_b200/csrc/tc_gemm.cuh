// tc_gemm.cuh -- hand-written tcgen05 (5th-gen tensor core) GEMMs for the Linear layers of the
// conv path, in 3xTF32 split precision so the result keeps fp32 accuracy (1e-5 parity mode):
//
//     x = x_hi + x_lo   (x_hi = rna_tf32(x), x_lo = rna_tf32(x - x_hi))
//     A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo        (fp32 accumulation in TMEM)
//
// Two persistent, warp-specialised kernels (1 CTA / SM, 416 threads):
//   warp 0      : TMEM allocation + single-thread tcgen05.mma issue
//   warps 1..8  : producers -- coalesced 128-bit global loads of fp32 activations, hi/lo split in
//                 registers, 128-bit stores into the UMMA canonical SWIZZLE_128B smem layout
//                 (the split has to touch every element anyway, so LDG->STS replaces TMA here).
//                 Two producer warps per SM sub-partition and two chunks of loads in flight per
//                 thread: with one warp per sub-partition the kernel was issue-latency bound
//                 (ncu: 13.7 cycles per issued instruction, profiles/r1_notes.md).
//   warps 9..12 : epilogue -- tcgen05.ld of the fp32 accumulator (one output row per thread),
//                 transposed through a padded smem tile so that every global store instruction
//                 writes whole 128-byte lines, bias / accumulate fused
// synchronised with mbarriers (full/empty smem ring, full/empty double-buffered TMEM).
//
//   G1  C[M,N] (+)= A[M,K] . B(K,N) + bias     A streamed (K-major), B = weights, pre-split into a
//       resident smem image (K-major).  Used for forward (y = x W^T) and dgrad (dx = dy W).
//   G2  P[cta][Dm,N] = sum_{rows of this CTA} Y[row,:]^T X[row,:]   both operands streamed and
//       MN-major; the accumulator stays in TMEM over the CTA's whole row range (split-K across
//       CTAs, partial tiles summed in fixed order by k_splitk_reduce => deterministic wgrad).
#pragma once
#include "common.cuh"

namespace x2 {
namespace tc {

constexpr int kThreads = 416;
constexpr int kProducerThreads = 256;
constexpr int kEpilogueThreads = 128;
constexpr int kStageWords = 36;      // padded row pitch (words) of the epilogue transpose tile
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
__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split4(const float4& v, uint4& hi, uint4& lo) {
  hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
  // lo is exact in fp32; the tensor core ignores its 13 low mantissa bits (~2^-22 of |x|)
  lo.x = __float_as_uint(v.x - __uint_as_float(hi.x));
  lo.y = __float_as_uint(v.y - __uint_as_float(hi.y));
  lo.z = __float_as_uint(v.z - __uint_as_float(hi.z));
  lo.w = __float_as_uint(v.w - __uint_as_float(hi.w));
}
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
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

// ------------------------------------------------------------------ weight image
// Builds the resident B operand of G1: B(k,n) = W[k*sbk + n*sbn] (zero padded to KC*32 x N_pad),
// split into hi / lo and laid out exactly as the CTA keeps it in smem:
//   img[half][kc][kmajor_off(n, kk/4)] , half 0 = hi, 1 = lo, each KC * N_pad * 128 bytes.
__global__ void k_make_bimage(const float* __restrict__ W, int64_t sbk, int64_t sbn, int K, int N,
                              int KC, int N_pad, uint32_t* __restrict__ img) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = KC * N_pad * kChunkK;
  if (idx >= total) return;
  const int kc = idx / (N_pad * kChunkK);
  const int rem = idx - kc * N_pad * kChunkK;
  const int n = rem / kChunkK, kk = rem - n * kChunkK;
  const int k = kc * kChunkK + kk;
  const float v = (k < K && n < N) ? W[(int64_t)k * sbk + (int64_t)n * sbn] : 0.f;
  const uint32_t hi = to_tf32(v);
  const uint32_t lo = to_tf32(v - __uint_as_float(hi));
  const uint32_t off = (uint32_t)kc * N_pad * 128 + kmajor_off(n, kk >> 2) + (kk & 3) * 4;
  img[off >> 2] = hi;
  img[((uint32_t)KC * N_pad * 128 + off) >> 2] = lo;
}

// ------------------------------------------------------------------ G1
struct G1Params {
  const float* A;
  int64_t lda, M;
  int K, KC;
  const uint32_t* bimg;
  int N, N_pad;
  const float* bias;
  float* C;
  int64_t ldc;
  int beta;
  int stages;
};

__global__ void __launch_bounds__(kThreads, 1) k_tc_gemm(const G1Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  const uint32_t bhalf = (uint32_t)p.KC * p.N_pad * 128;          // bytes of one B image half
  uint8_t* sB = smem;                                             // [hi | lo]
  uint8_t* sA = smem + 2 * bhalf;                                 // S x [hi 16K | lo 16K]
  float* sT = reinterpret_cast<float*>(sA + (size_t)S * 2 * kChunkBytes);   // 4 x [32][36] transpose tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sT + 4 * 32 * kStageWords);
  uint64_t* full = bars;            // [S]   producers -> MMA
  uint64_t* empty = bars + S;       // [S]   MMA -> producers
  uint64_t* tfull = bars + 2 * S;   // [2]   MMA -> epilogue
  uint64_t* tempty = bars + 2 * S + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // ---- one-time setup: barriers, TMEM, resident weight image
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
  if (warp == 0) tmem_alloc(tmem_slot, kTmemCols);
  {
    const uint4* src = reinterpret_cast<const uint4*>(p.bimg);
    uint4* dst = reinterpret_cast<uint4*>(sB);
    const int n16 = (int)(2 * bhalf / 16);
    for (int i = threadIdx.x; i < n16; i += kThreads) dst[i] = __ldg(src + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t ntiles = (p.M + kTileM - 1) / kTileM;
  const int KC = p.KC;

  if (warp == 0) {
    // =============================== MMA issuer ===============================
    if (lane == 0) {
      const uint32_t idesc = make_idesc(p.N_pad, 0, 0);
      const uint32_t sA_u = smem_u32(sA), sB_u = smem_u32(sB);
      uint32_t it = 0;       // global chunk counter (ring position)
      uint32_t tcount = 0;   // tile counter (TMEM buffer)
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        const uint32_t buf = tcount & 1;
        mbar_wait(&tempty[buf], ((tcount >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + buf * 128;
        for (int kc = 0; kc < KC; ++kc, ++it) {
          const uint32_t st = it % S, ph = (it / S) & 1;
          mbar_wait(&full[st], ph);
          tc_fence_after();
          const int kvalid = min(kChunkK, p.K - kc * kChunkK);
          const int ksteps = (kvalid + 7) >> 3;
          const uint64_t dah = make_desc(sA_u + st * 2 * kChunkBytes, 16, 1024);
          const uint64_t dal = make_desc(sA_u + st * 2 * kChunkBytes + kChunkBytes, 16, 1024);
          const uint64_t dbh = make_desc(sB_u + (uint32_t)kc * p.N_pad * 128, 16, 1024);
          const uint64_t dbl = make_desc(sB_u + (uint32_t)kc * p.N_pad * 128 + bhalf, 16, 1024);
          for (int ks = 0; ks < ksteps; ++ks) {
            const uint64_t adv = (uint64_t)(ks * 2);      // +32 bytes (>>4) along K inside the swizzle row
            umma_tf32(taddr, dah + adv, dbh + adv, idesc, (kc | ks) != 0);
            umma_tf32(taddr, dal + adv, dbh + adv, idesc, 1);
            umma_tf32(taddr, dah + adv, dbl + adv, idesc, 1);
          }
          umma_commit(&empty[st]);          // smem slot reusable once these MMAs retire
        }
        umma_commit(&tfull[buf]);           // accumulator complete
      }
    }
  } else if (warp <= 8) {
    // =============================== producers ===============================
    // 256 threads; thread (c16, r0) moves the 16-byte column chunk c16 of rows r0 + 32 i.  Two
    // chunks of global loads are in flight per thread (register double buffering).
    const int pt = threadIdx.x - 32;                 // 0..255
    const int c16 = pt & 7, r0 = pt >> 3;            // 8 threads cover one 128-byte row segment
    const bool vec = ((p.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.A) & 15) == 0);
    const int64_t my_tiles = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t nchunk = my_tiles * KC;
    const uint32_t sA_u = smem_u32(sA);
    const float* const a_thr = p.A + (int64_t)r0 * p.lda + c16 * 4;
    const int64_t row_step = 32 * p.lda;
    uint32_t soff[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) soff[i] = kmajor_off(r0 + 32 * i, c16);

    auto issue = [&](float4 (&v)[4], int64_t it) {
      const int64_t m0 = (blockIdx.x + (it / KC) * gridDim.x) * kTileM;
      const int kc0 = (int)(it % KC) * kChunkK;
      const float* src = a_thr + m0 * p.lda + kc0;
      if (vec && m0 + kTileM <= p.M && kc0 + kChunkK <= p.K) {       // interior chunk: no guards
#pragma unroll
        for (int i = 0; i < 4; ++i) v[i] = __ldg(reinterpret_cast<const float4*>(src + i * row_step));
      } else {
        const int kcol = kc0 + c16 * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + r0 + 32 * i < p.M) {
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
    auto commit = [&](const float4 (&v)[4], int64_t it) {
      const uint32_t st = (uint32_t)(it % S), ph = (uint32_t)((it / S) & 1);
      mbar_wait(&empty[st], ph ^ 1);
      const uint32_t base = sA_u + st * 2 * kChunkBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 hi, lo;
        split4(v[i], hi, lo);
        sts128(base + soff[i], hi);
        sts128(base + kChunkBytes + soff[i], lo);
      }
      fence_proxy_async();
      mbar_arrive(&full[st]);
    };

    float4 va[4], vb[4];
    if (nchunk > 0) issue(va, 0);
    for (int64_t it = 0; it < nchunk; it += 2) {
      if (it + 1 < nchunk) issue(vb, it + 1);
      commit(va, it);
      if (it + 1 < nchunk) {
        if (it + 2 < nchunk) issue(va, it + 2);
        commit(vb, it + 1);
      }
    }
  } else {
    // =============================== epilogue ===============================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    float* tile_s = sT + q * 32 * kStageWords;       // this warp's [32][36] transpose tile
    const bool vecC = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) && ((p.N & 3) == 0);
    const int rr = lane >> 3, cc = (lane & 7) * 4;   // read-back mapping: 4 rows x 8 float4 per pass
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
      const uint32_t buf = tcount & 1;
      mbar_wait(&tfull[buf], (tcount >> 1) & 1);
      tc_fence_after();
      const int64_t row0 = tile * kTileM + q * 32;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128;
      for (int c0 = 0; c0 < p.N_pad; c0 += 32) {
        // thread = accumulator row: 32 (or 16) columns -> padded smem tile
        float v[16];
        tmem_ld16(taddr + c0, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          *reinterpret_cast<float4*>(tile_s + lane * kStageWords + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (c0 + 16 < p.N_pad) {
          tmem_ld16(taddr + c0 + 16, v);
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(tile_s + lane * kStageWords + 16 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
        __syncwarp();
        if (vecC) {
          // each store instruction writes 4 rows x 128 contiguous bytes
          const int col = c0 + cc;
          if (col < p.N) {
            float4 bv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) bv = __ldg(reinterpret_cast<const float4*>(p.bias + col));
#pragma unroll
            for (int pass = 0; pass < 8; ++pass) {
              const int r = pass * 4 + rr;
              const int64_t row = row0 + r;
              if (row < p.M) {
                float4 o = *reinterpret_cast<const float4*>(tile_s + r * kStageWords + cc);
                o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
                float* dst = p.C + row * p.ldc + col;
                if (p.beta) {
                  const float4 old = *reinterpret_cast<const float4*>(dst);
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                *reinterpret_cast<float4*>(dst) = o;
              }
            }
          }
        } else {
          const int col = c0 + lane;
          if (col < p.N) {
            const float bv = p.bias ? __ldg(p.bias + col) : 0.f;
            for (int r = 0; r < 32; ++r) {
              const int64_t row = row0 + r;
              if (row < p.M) {
                float o = tile_s[r * kStageWords + lane] + bv;
                float* dst = p.C + row * p.ldc + col;
                if (p.beta) o += *dst;
                *dst = o;
              }
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      mbar_arrive(&tempty[buf]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ G2 (wgrad)
struct G2Params {
  const float* Y;        // [rows, >=128] : the 128 columns starting at Y are the output rows (Dm)
  int64_t ldy;
  const float* X;        // [rows, N]
  int64_t ldx;
  int N, N_pad;          // N_pad multiple of 32, <= 128
  int64_t rows, rows_per_cta;   // rows_per_cta multiple of 32
  float* partial;        // [grid][128][N]
  float* colsum;         // [grid][128] or NULL
  int stages;
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

__global__ void __launch_bounds__(kThreads, 1) k_tc_wgrad(const G2Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int S = p.stages;
  const uint32_t xbytes = (uint32_t)(p.N_pad / 32) * 4096;      // one X half (hi or lo)
  const uint32_t stage_bytes = 2 * kChunkBytes + 2 * xbytes;    // Y hi | Y lo | X hi | X lo
  uint8_t* sS = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sS + (size_t)S * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  uint64_t* tfull = bars + 2 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  float* cs_smem = reinterpret_cast<float*>(bars + 2 * S + 2);   // [8][128] column-sum staging

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full[i], kProducerThreads);
      mbar_init(&empty[i], 1);
    }
    mbar_init(tfull, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int64_t rbeg = (int64_t)blockIdx.x * p.rows_per_cta;
  const int64_t rend = min(p.rows, rbeg + p.rows_per_cta);
  const int64_t nchunks = rend > rbeg ? (rend - rbeg + kChunkK - 1) / kChunkK : 0;

  if (warp == 0) {
    if (lane == 0 && nchunks > 0) {
      const uint32_t idesc = make_idesc(p.N_pad, 1, 1);
      const uint32_t s_u = smem_u32(sS);
      for (int64_t c = 0; c < nchunks; ++c) {
        const uint32_t st = (uint32_t)(c % S), ph = (uint32_t)((c / S) & 1);
        mbar_wait(&full[st], ph);
        tc_fence_after();
        const uint32_t y_hi = s_u + st * stage_bytes, y_lo = y_hi + kChunkBytes;
        const uint32_t x_hi = y_lo + kChunkBytes, x_lo = x_hi + xbytes;
        const int kvalid = (int)min((int64_t)kChunkK, rend - rbeg - c * kChunkK);
        const int ksteps = (kvalid + 7) >> 3;
        const uint64_t dyh = make_desc(y_hi, 4096, 512, kLayoutSW128Base32);
        const uint64_t dyl = make_desc(y_lo, 4096, 512, kLayoutSW128Base32);
        const uint64_t dxh = make_desc(x_hi, 4096, 512, kLayoutSW128Base32);
        const uint64_t dxl = make_desc(x_lo, 4096, 512, kLayoutSW128Base32);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t adv = (uint64_t)(ks * 64);       // one K=8 step = two 4-row k-groups = 1024 B (>>4)
          umma_tf32(tmem_base, dyh + adv, dxh + adv, idesc, (c | ks) != 0);
          umma_tf32(tmem_base, dyl + adv, dxh + adv, idesc, 1);
          umma_tf32(tmem_base, dyh + adv, dxl + adv, idesc, 1);
        }
        umma_commit(&empty[st]);
      }
      umma_commit(tfull);
    }
  } else if (warp <= 8) {
    // 256 producer threads.  Y chunk = 32 rows x 32 float4: thread owns float4 column yc of rows
    // yr + 8 i.  X chunk = 32 rows x xq float4, flattened (up to 4 per thread).  Two chunks of loads
    // in flight per thread.
    const int pt = threadIdx.x - 32;
    const int yc = pt & 31, yr = pt >> 5;
    const bool vecY = ((p.ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.Y) & 15) == 0);
    const bool vecX = ((p.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.X) & 15) == 0) && ((p.N & 3) == 0);
    const int xq = p.N_pad / 4;                      // float4 per X row (8..32)
    const int nx = kChunkK * xq;                     // float4 per X chunk (256..1024)
    float cs0 = 0.f, cs1 = 0.f, cs2 = 0.f, cs3 = 0.f;
    const uint32_t s_u = smem_u32(sS);
    const float* const y_thr = p.Y + (int64_t)yr * p.ldy + yc * 4;
    const int64_t y_step = 8 * p.ldy;
    uint32_t yoff[4], xoff[4];
    int xr[4], xcol[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      yoff[i] = mnmajor_off(yr + 8 * i, yc);
      const int f = pt + i * kProducerThreads;
      xr[i] = f / xq;
      const int xc = f - xr[i] * xq;
      xcol[i] = xc * 4;
      xoff[i] = mnmajor_off(xr[i] & 31, xc);
    }

    auto issue = [&](float4 (&vy)[4], float4 (&vx)[4], int64_t c) {
      const int64_t row0 = rbeg + c * kChunkK;
      const bool interior = row0 + kChunkK <= rend;
      const float* ysrc = y_thr + row0 * p.ldy;
      if (interior && vecY) {
#pragma unroll
        for (int i = 0; i < 4; ++i) vy[i] = __ldg(reinterpret_cast<const float4*>(ysrc + i * y_step));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          vy[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row0 + yr + 8 * i < rend) {
            const float* q = ysrc + i * y_step;
            if (vecY) vy[i] = __ldg(reinterpret_cast<const float4*>(q));
            else { vy[i].x = __ldg(q); vy[i].y = __ldg(q + 1); vy[i].z = __ldg(q + 2); vy[i].w = __ldg(q + 3); }
          }
        }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        vx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (pt + i * kProducerThreads < nx) {
          const int64_t row = row0 + xr[i];
          if (row < rend) {
            const float* q = p.X + row * p.ldx + xcol[i];
            if (vecX && xcol[i] + 3 < p.N) vx[i] = __ldg(reinterpret_cast<const float4*>(q));
            else {
              if (xcol[i] < p.N) vx[i].x = __ldg(q);
              if (xcol[i] + 1 < p.N) vx[i].y = __ldg(q + 1);
              if (xcol[i] + 2 < p.N) vx[i].z = __ldg(q + 2);
              if (xcol[i] + 3 < p.N) vx[i].w = __ldg(q + 3);
            }
          }
        }
      }
    };
    auto commit = [&](const float4 (&vy)[4], const float4 (&vx)[4], int64_t c) {
      const uint32_t st = (uint32_t)(c % S), ph = (uint32_t)((c / S) & 1);
      mbar_wait(&empty[st], ph ^ 1);
      const uint32_t base = s_u + st * stage_bytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        cs0 += vy[i].x; cs1 += vy[i].y; cs2 += vy[i].z; cs3 += vy[i].w;   // fixed order per thread
        uint4 hi, lo;
        split4(vy[i], hi, lo);
        sts128(base + yoff[i], hi);
        sts128(base + kChunkBytes + yoff[i], lo);
      }
      const uint32_t xb = base + 2 * kChunkBytes;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (pt + i * kProducerThreads < nx) {
          uint4 hi, lo;
          split4(vx[i], hi, lo);
          sts128(xb + xoff[i], hi);
          sts128(xb + xbytes + xoff[i], lo);
        }
      }
      fence_proxy_async();
      mbar_arrive(&full[st]);
    };

    float4 ya[4], xa[4], yb[4], xb_[4];
    if (nchunks > 0) issue(ya, xa, 0);
    for (int64_t c = 0; c < nchunks; c += 2) {
      if (c + 1 < nchunks) issue(yb, xb_, c + 1);
      commit(ya, xa, c);
      if (c + 1 < nchunks) {
        if (c + 2 < nchunks) issue(ya, xa, c + 2);
        commit(yb, xb_, c + 1);
      }
    }
    // column sums of Y over this CTA's rows: combine the 8 row-phase threads in fixed order
    if (p.colsum) {
      cs_smem[yr * 128 + yc * 4 + 0] = cs0;
      cs_smem[yr * 128 + yc * 4 + 1] = cs1;
      cs_smem[yr * 128 + yc * 4 + 2] = cs2;
      cs_smem[yr * 128 + yc * 4 + 3] = cs3;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (pt < 128) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < 8; ++g) s += cs_smem[g * 128 + pt];
        p.colsum[(int64_t)blockIdx.x * 128 + pt] = s;
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    float* dst = p.partial + ((int64_t)blockIdx.x * 128 + m) * p.N;
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
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------ host launchers
static inline int ceil_to(int v, int m) { return (v + m - 1) / m * m; }

// bytes of workspace for one weight image (hi + lo)
static inline size_t bimage_bytes(int K, int N) {
  const int KC = (K + kChunkK - 1) / kChunkK, N_pad = ceil_to(N, 16);
  return align_up((size_t)2 * KC * N_pad * 128, 256);
}

constexpr size_t kEpiBytes = 4 * 32 * kStageWords * sizeof(float);   // epilogue transpose tiles

static inline bool g1_supported(int K, int N) {
  if (K < 1 || N < 1 || N > 128) return false;
  const int KC = (K + kChunkK - 1) / kChunkK, N_pad = ceil_to(N, 16);
  const size_t b = (size_t)2 * KC * N_pad * 128;
  return b + 2 * (size_t)2 * kChunkBytes + kEpiBytes + 2048 <= (size_t)kMaxSmem;
}

// C[M,N] (+)= A[M,K] . B(K,N) + bias, B(k,n) = W[k*sbk + n*sbn].  img: scratch for the weight image.
static int tc_gemm(const float* A, int64_t lda, int64_t M, int K, const float* W, int64_t sbk, int64_t sbn,
                   int N, const float* bias, float* C, int64_t ldc, int beta, void* img, cudaStream_t st) {
  if (M <= 0) return X2_OK;
  if (!g1_supported(K, N)) { set_error("tc_gemm: unsupported K=%d N=%d", K, N); return X2_EINVAL; }
  const int KC = (K + kChunkK - 1) / kChunkK, N_pad = ceil_to(N, 16);
  const int total = KC * N_pad * kChunkK;
  k_make_bimage<<<(total + 255) / 256, 256, 0, st>>>(W, sbk, sbn, K, N, KC, N_pad, static_cast<uint32_t*>(img));
  X2_LAUNCH_OK();
  const size_t bbytes = (size_t)2 * KC * N_pad * 128;
  int stages = (int)(((size_t)kMaxSmem - bbytes - kEpiBytes - 2048) / (2 * kChunkBytes));
  if (stages > 4) stages = 4;
  const size_t smem = 1024 + bbytes + (size_t)stages * 2 * kChunkBytes + kEpiBytes + 256;
  static bool attr_set = false;
  if (!attr_set) {
    X2_CUDA_OK(cudaFuncSetAttribute(k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_set = true;
  }
  G1Params p;
  p.A = A; p.lda = lda; p.M = M; p.K = K; p.KC = KC;
  p.bimg = static_cast<const uint32_t*>(img);
  p.N = N; p.N_pad = N_pad; p.bias = bias; p.C = C; p.ldc = ldc; p.beta = beta; p.stages = stages;
  const int64_t ntiles = cdiv(M, kTileM);
  const int grid = (int)(ntiles < kNumSM ? ntiles : kNumSM);
  k_tc_gemm<<<grid, kThreads, smem, st>>>(p);
  X2_LAUNCH_OK();
  return X2_OK;
}

static inline int wgrad_ctas(int64_t rows) {
  int64_t c = cdiv(rows > 0 ? rows : 1, 256);     // >= 256 rows per CTA
  if (c > kNumSM) c = kNumSM;
  if (c < 1) c = 1;
  return (int)c;
}
static inline size_t tc_wgrad_workspace_floats(int64_t rows, int N) {
  return (size_t)wgrad_ctas(rows) * (128 * (size_t)N + 128) + 64;
}

}  // namespace tc

__global__ void k_splitk_reduce(const float* __restrict__ partial, const float* __restrict__ colsum,
                                int splits, int64_t M, int N, float* __restrict__ out, int64_t ldo,
                                float* __restrict__ bias);
static inline unsigned splitk_reduce_blocks(int64_t M, int N, bool with_bias);

namespace tc {

// dW[128,N] = Y[rows,128]^T X[rows,N] ; db[128] = colsum(Y)  (db may be NULL).  ws: tc_wgrad_workspace_floats
static int tc_wgrad(const float* Y, int64_t ldy, const float* X, int64_t ldx, int64_t rows, int N, float* dW,
                    int64_t lddw, float* db, float* ws, cudaStream_t st) {
  if (N < 1 || N > 128) { set_error("tc_wgrad: unsupported N=%d", N); return X2_EINVAL; }
  const int N_pad = ceil_to(N, 32);
  const int grid = wgrad_ctas(rows);
  const int64_t rpc = cdiv(cdiv(rows > 0 ? rows : 1, grid), kChunkK) * kChunkK;
  const uint32_t stage_bytes = 2 * kChunkBytes + 2 * (uint32_t)(N_pad / 32) * 4096;
  int stages = (int)(((size_t)kMaxSmem - 8192 - 2048) / stage_bytes);
  if (stages > 4) stages = 4;
  const size_t smem = 1024 + (size_t)stages * stage_bytes + 256 + 8 * 128 * sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    X2_CUDA_OK(cudaFuncSetAttribute(k_tc_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    attr_set = true;
  }
  G2Params p;
  p.Y = Y; p.ldy = ldy; p.X = X; p.ldx = ldx; p.N = N; p.N_pad = N_pad; p.rows = rows; p.rows_per_cta = rpc;
  p.partial = ws; p.colsum = db ? ws + (size_t)grid * 128 * N : nullptr; p.stages = stages;
  k_tc_wgrad<<<grid, kThreads, smem, st>>>(p);
  X2_LAUNCH_OK();
  k_splitk_reduce<<<splitk_reduce_blocks(128, N, db != nullptr), 256, 0, st>>>(ws, p.colsum, grid, 128, N, dW, lddw, db);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // namespace tc
}  // namespace x2
