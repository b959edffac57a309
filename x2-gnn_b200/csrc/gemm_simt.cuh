// gemm_simt.cuh -- fp32 SIMT GEMM used by the X2_MODE_FP32 (1e-5 parity) path of the conv layer.
//
//   C[M,N] = (beta ? C : 0) + A(M,K) * B(K,N) (+ bias[n])
//
// Operand layouts are template flags so the same kernel serves the three shapes of a Linear:
//   forward   y = x W^T      A k-contiguous (x[M,K]),   B k-contiguous (W[N,K])
//   dgrad     dx = dy W      A k-contiguous (dy[M,N']), B n-contiguous (W[K,N])
//   wgrad     dW = dy^T x    A m-contiguous (dy[K,M]),  B n-contiguous (x[K,N]); reduction over
//                            rows is split across gridDim.z into partial tiles that a second
//                            kernel sums in a fixed order (deterministic, no atomics); the column
//                            sums of dy (bias gradient) fall out of the A tiles already in smem.
//
// 128 x BN x 16 tiles, 256 threads, 8 x (BN/16) register micro-tile, double-buffered smem with
// 128-bit global loads when the operand is 16-byte aligned.
#pragma once
#include "common.cuh"

namespace x2 {

constexpr int GM = 128;      // tile rows
constexpr int GK = 16;       // tile depth
constexpr int GPAD = 4;      // smem row padding (keeps float4 alignment, breaks store conflicts)
constexpr int GTHREADS = 256;
constexpr int kMaxBatch = 4;

struct GemmBatch {
  const float* A[kMaxBatch];
  const float* B[kMaxBatch];
  const float* bias[kMaxBatch];
  float* C[kMaxBatch];
};

// MODE: 0 = plain (blockIdx.z = batch index), 1 = split-K partial output (blockIdx.z = split)
template <int BN, bool A_KC, bool B_KC, int MODE>
__global__ void __launch_bounds__(GTHREADS)
k_gemm(GemmBatch p, int64_t M, int N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc, int beta,
       int64_t kchunk, float* __restrict__ colsum) {
  constexpr int TN = BN / 16;
  __shared__ __align__(16) float As[2][GK][GM + GPAD];
  __shared__ __align__(16) float Bs[2][GK][BN + GPAD];

  const int z = blockIdx.z;
  const float* __restrict__ A = MODE == 0 ? p.A[z] : p.A[0];
  const float* __restrict__ B = MODE == 0 ? p.B[z] : p.B[0];
  const int64_t m0 = (int64_t)blockIdx.y * GM;
  const int n0 = blockIdx.x * BN;
  const int64_t kbeg = MODE == 1 ? (int64_t)z * kchunk : 0;
  const int64_t kend = MODE == 1 ? min(K, kbeg + kchunk) : K;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const bool vecA = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool vecB = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
  float asum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) asum[i] = 0.f;

  // ---- staging registers: A tile = 128x16 = 2048 floats = 8/thread; B tile = BN x 16
  float ra[8];
  float rb[TN];

  auto load_a = [&](int64_t k0) {
    if (A_KC) {
      if (vecA) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = tid + i * GTHREADS;
          const int m = f >> 2, kq = (f & 3) * 4;
          const int64_t gm = m0 + m, gk = k0 + kq;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gm < M) {
            if (gk + 3 < kend) v = *reinterpret_cast<const float4*>(A + gm * lda + gk);
            else {
              if (gk < kend) v.x = A[gm * lda + gk];
              if (gk + 1 < kend) v.y = A[gm * lda + gk + 1];
              if (gk + 2 < kend) v.z = A[gm * lda + gk + 2];
            }
          }
          ra[i * 4 + 0] = v.x; ra[i * 4 + 1] = v.y; ra[i * 4 + 2] = v.z; ra[i * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = tid & 15, m = (tid >> 4) + i * 16;
          const int64_t gm = m0 + m, gk = k0 + k;
          ra[i] = (gm < M && gk < kend) ? A[gm * lda + gk] : 0.f;
        }
      }
    } else {  // m-contiguous: A(m,k) = A[k*lda + m]
      if (vecA) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = tid + i * GTHREADS;
          const int k = f >> 5, mq = (f & 31) * 4;
          const int64_t gm = m0 + mq, gk = k0 + k;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gk < kend) {
            if (gm + 3 < M) v = *reinterpret_cast<const float4*>(A + gk * lda + gm);
            else {
              if (gm < M) v.x = A[gk * lda + gm];
              if (gm + 1 < M) v.y = A[gk * lda + gm + 1];
              if (gm + 2 < M) v.z = A[gk * lda + gm + 2];
            }
          }
          ra[i * 4 + 0] = v.x; ra[i * 4 + 1] = v.y; ra[i * 4 + 2] = v.z; ra[i * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int m = tid & 127, k = (tid >> 7) + i * 2;
          const int64_t gm = m0 + m, gk = k0 + k;
          ra[i] = (gm < M && gk < kend) ? A[gk * lda + gm] : 0.f;
        }
      }
    }
  };
  auto store_a = [&](int buf) {
    if (A_KC) {
      if (vecA) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = tid + i * GTHREADS;
          const int m = f >> 2, kq = (f & 3) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) As[buf][kq + j][m] = ra[i * 4 + j];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][tid & 15][(tid >> 4) + i * 16] = ra[i];
      }
    } else {
      if (vecA) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int f = tid + i * GTHREADS;
          const int k = f >> 5, mq = (f & 31) * 4;
          *reinterpret_cast<float4*>(&As[buf][k][mq]) =
              make_float4(ra[i * 4], ra[i * 4 + 1], ra[i * 4 + 2], ra[i * 4 + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) As[buf][(tid >> 7) + i * 2][tid & 127] = ra[i];
      }
    }
  };

  // B tile: BN x 16 floats = TN per thread (scalar element mapping; BN*16/256 = TN)
  auto load_b = [&](int64_t k0) {
    if (B_KC) {  // B(k,n) = B[n*ldb + k]
      if (TN % 4 == 0 && vecB) {
#pragma unroll
        for (int i = 0; i < TN / 4; ++i) {
          const int f = tid + i * GTHREADS;
          const int n = f >> 2, kq = (f & 3) * 4;
          const int gn = n0 + n;
          const int64_t gk = k0 + kq;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gn < N) {
            if (gk + 3 < kend) v = *reinterpret_cast<const float4*>(B + (int64_t)gn * ldb + gk);
            else {
              if (gk < kend) v.x = B[(int64_t)gn * ldb + gk];
              if (gk + 1 < kend) v.y = B[(int64_t)gn * ldb + gk + 1];
              if (gk + 2 < kend) v.z = B[(int64_t)gn * ldb + gk + 2];
            }
          }
          rb[i * 4 + 0] = v.x; rb[i * 4 + 1] = v.y; rb[i * 4 + 2] = v.z; rb[i * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TN; ++i) {
          const int k = tid & 15, n = (tid >> 4) + i * 16;
          const int gn = n0 + n;
          const int64_t gk = k0 + k;
          rb[i] = (gn < N && gk < kend) ? B[(int64_t)gn * ldb + gk] : 0.f;
        }
      }
    } else {  // n-contiguous: B(k,n) = B[k*ldb + n]
      if (TN % 4 == 0 && vecB) {
#pragma unroll
        for (int i = 0; i < TN / 4; ++i) {
          const int f = tid + i * GTHREADS;
          const int k = f / (BN / 4), nq = (f % (BN / 4)) * 4;
          const int gn = n0 + nq;
          const int64_t gk = k0 + k;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (gk < kend) {
            if (gn + 3 < N) v = *reinterpret_cast<const float4*>(B + gk * ldb + gn);
            else {
              if (gn < N) v.x = B[gk * ldb + gn];
              if (gn + 1 < N) v.y = B[gk * ldb + gn + 1];
              if (gn + 2 < N) v.z = B[gk * ldb + gn + 2];
            }
          }
          rb[i * 4 + 0] = v.x; rb[i * 4 + 1] = v.y; rb[i * 4 + 2] = v.z; rb[i * 4 + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < TN; ++i) {
          const int f = tid + i * GTHREADS;
          const int k = f / BN, n = f % BN;
          const int gn = n0 + n;
          const int64_t gk = k0 + k;
          rb[i] = (gn < N && gk < kend) ? B[gk * ldb + gn] : 0.f;
        }
      }
    }
  };
  auto store_b = [&](int buf) {
    if (B_KC) {
      if (TN % 4 == 0 && vecB) {
#pragma unroll
        for (int i = 0; i < TN / 4; ++i) {
          const int f = tid + i * GTHREADS;
          const int n = f >> 2, kq = (f & 3) * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) Bs[buf][kq + j][n] = rb[i * 4 + j];
        }
      } else {
#pragma unroll
        for (int i = 0; i < TN; ++i) Bs[buf][tid & 15][(tid >> 4) + i * 16] = rb[i];
      }
    } else {
      if (TN % 4 == 0 && vecB) {
#pragma unroll
        for (int i = 0; i < TN / 4; ++i) {
          const int f = tid + i * GTHREADS;
          const int k = f / (BN / 4), nq = (f % (BN / 4)) * 4;
          *reinterpret_cast<float4*>(&Bs[buf][k][nq]) =
              make_float4(rb[i * 4], rb[i * 4 + 1], rb[i * 4 + 2], rb[i * 4 + 3]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < TN; ++i) {
          const int f = tid + i * GTHREADS;
          Bs[buf][f / BN][f % BN] = rb[i];
        }
      }
    }
  };

  const int64_t ntiles = kend > kbeg ? (kend - kbeg + GK - 1) / GK : 0;
  if (ntiles > 0) {
    load_a(kbeg);
    load_b(kbeg);
    store_a(0);
    store_b(0);
  }
  __syncthreads();
  for (int64_t it = 0; it < ntiles; ++it) {
    const int buf = (int)(it & 1);
    if (it + 1 < ntiles) {
      load_a(kbeg + (it + 1) * GK);
      load_b(kbeg + (it + 1) * GK);
    }
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float a[8], b[TN];
      *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      if constexpr (TN == 8) {
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
        *reinterpret_cast<float4*>(&b[4]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      } else if constexpr (TN == 4) {
        *reinterpret_cast<float4*>(&b[0]) = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      if (MODE == 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asum[i] += a[i];
      }
    }
    if (it + 1 < ntiles) {
      store_a(buf ^ 1);
      store_b(buf ^ 1);
    }
    __syncthreads();
  }

  // ---- epilogue
  float* __restrict__ C = MODE == 0 ? p.C[z] : p.C[0] + (int64_t)z * M * ldc;
  const float* __restrict__ bias = MODE == 0 ? p.bias[z] : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n;
      if constexpr (TN == 8) n = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      else n = tx * TN + j;
      const int gn = n0 + n;
      if (gn >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[gn];
      if (beta) v += C[gm * ldc + gn];
      C[gm * ldc + gn] = v;
    }
  }
  if (MODE == 1 && colsum != nullptr && blockIdx.x == 0 && tx == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      if (gm < M) colsum[(int64_t)z * M + gm] = asum[i];
    }
  }
}

// out[m*ldo + n] = sum_z partial[z][m][n] ; bias[m] = sum_z colsum[z][m].
// 256 threads = 8 warps x 32 outputs: warp g sums the splits z = g, g+8, ... in order, then the 8
// per-warp sums are combined in order g = 0..7 -- a fixed summation tree (deterministic) that is 8x
// shorter than one serial loop over all splits.
__global__ void __launch_bounds__(256)
k_splitk_reduce(const float* __restrict__ partial, const float* __restrict__ colsum, int splits, int64_t M,
                int N, float* __restrict__ out, int64_t ldo, float* __restrict__ bias) {
  __shared__ float red[8][33];
  pdl_sync();
  const int lane = threadIdx.x & 31, g = threadIdx.x >> 5;
  const int64_t MN = M * N;
  const int64_t nblk_out = (MN + 31) / 32;
  const bool is_bias = (int64_t)blockIdx.x >= nblk_out;      // trailing blocks reduce the column sums
  const int64_t idx = is_bias ? ((int64_t)blockIdx.x - nblk_out) * 32 + lane : (int64_t)blockIdx.x * 32 + lane;
  const int64_t lim = is_bias ? M : MN;
  const float* src = is_bias ? colsum : partial;
  float s = 0.f;
  if (idx < lim) {
    // loads issued eight at a time, additions in the original order (same bits): with hundreds of splits
    // (the 592 partial tiles of dW_r) the serial load -> add chain was 17 us for 1.8 MB
    int z = g;
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
  if (g == 0 && idx < lim) {
    float t = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][lane];
    if (is_bias) bias[idx] = t;
    else {
      const int64_t m = idx / N;
      out[m * ldo + (idx - m * N)] = t;
    }
  }
}

static inline unsigned splitk_reduce_blocks(int64_t M, int N, bool with_bias) {
  return (unsigned)(((M * N + 31) / 32) + (with_bias ? (M + 31) / 32 : 0));
}

// ------------------------------------------------------------------------- host launchers
template <bool A_KC, bool B_KC>
static int launch_gemm(const GemmBatch& p, int nbatch, int64_t M, int N, int64_t K, int64_t lda,
                       int64_t ldb, int64_t ldc, int beta, cudaStream_t st) {
  if (M <= 0 || N <= 0) return X2_OK;
  dim3 block(GTHREADS);
  if (N > 64) {
    dim3 grid((unsigned)cdiv(N, 128), (unsigned)cdiv(M, GM), nbatch);
    k_gemm<128, A_KC, B_KC, 0><<<grid, block, 0, st>>>(p, M, N, K, lda, ldb, ldc, beta, 0, nullptr);
  } else if (N > 32) {
    dim3 grid((unsigned)cdiv(N, 64), (unsigned)cdiv(M, GM), nbatch);
    k_gemm<64, A_KC, B_KC, 0><<<grid, block, 0, st>>>(p, M, N, K, lda, ldb, ldc, beta, 0, nullptr);
  } else {
    dim3 grid((unsigned)cdiv(N, 32), (unsigned)cdiv(M, GM), nbatch);
    k_gemm<32, A_KC, B_KC, 0><<<grid, block, 0, st>>>(p, M, N, K, lda, ldb, ldc, beta, 0, nullptr);
  }
  X2_LAUNCH_OK();
  return X2_OK;
}

// y[M,N] = x[M,K] W[N,K]^T + bias
static inline int gemm_nt(const float* x, int64_t ldx, const float* W, int64_t ldw, const float* bias,
                          float* y, int64_t ldy, int64_t M, int N, int64_t K, cudaStream_t st) {
  GemmBatch p{};
  p.A[0] = x; p.B[0] = W; p.bias[0] = bias; p.C[0] = y;
  return launch_gemm<true, true>(p, 1, M, N, K, ldx, ldw, ldy, 0, st);
}

// dx[M,N] (+)= dy[M,K] W[K,N]
static inline int gemm_nn(const float* dy, int64_t lddy, const float* W, int64_t ldw, float* dx,
                          int64_t lddx, int64_t M, int N, int64_t K, int beta, cudaStream_t st) {
  GemmBatch p{};
  p.A[0] = dy; p.B[0] = W; p.bias[0] = nullptr; p.C[0] = dx;
  return launch_gemm<true, false>(p, 1, M, N, K, lddy, ldw, lddx, beta, st);
}

static inline int wgrad_splits(int64_t rows, int64_t tiles) {
  int64_t s = (2 * (int64_t)kNumSM + tiles - 1) / tiles;      // ~2 CTAs per SM
  const int64_t maxs = cdiv(rows, 4 * GK);                     // >= 64 rows per split
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  if (s > 1024) s = 1024;
  return (int)s;
}

static inline size_t wgrad_workspace_floats(int64_t rows, int M, int N) {
  const int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  const int64_t tiles = cdiv(M, GM) * cdiv(N, bn);
  const int s = wgrad_splits(rows, tiles);
  return (size_t)s * ((size_t)M * N + (size_t)M) + 64;
}

// dW[M,N] = dy[rows,M]^T x[rows,N];  db[M] = colsum(dy)   (db may be NULL)
static inline int gemm_wgrad(const float* dy, int64_t lddy, const float* x, int64_t ldx, float* dW,
                             int64_t lddw, float* db, int64_t rows, int M, int N, float* ws,
                             cudaStream_t st) {
  if (M <= 0 || N <= 0) return X2_OK;
  const int bn = N > 64 ? 128 : (N > 32 ? 64 : 32);
  const int64_t tiles = cdiv(M, GM) * cdiv(N, bn);
  const int splits = wgrad_splits(rows, tiles);
  int64_t kchunk = cdiv(cdiv(rows > 0 ? rows : 1, splits), GK) * GK;
  float* partial = ws;
  float* colsum = ws + (size_t)splits * M * N;
  GemmBatch p{};
  p.A[0] = dy; p.B[0] = x; p.bias[0] = nullptr; p.C[0] = partial;
  dim3 block(GTHREADS);
  dim3 grid((unsigned)cdiv(N, bn), (unsigned)cdiv(M, GM), splits);
  float* cs = db ? colsum : nullptr;
  if (bn == 128) k_gemm<128, false, false, 1><<<grid, block, 0, st>>>(p, M, N, rows, lddy, ldx, N, 0, kchunk, cs);
  else if (bn == 64) k_gemm<64, false, false, 1><<<grid, block, 0, st>>>(p, M, N, rows, lddy, ldx, N, 0, kchunk, cs);
  else k_gemm<32, false, false, 1><<<grid, block, 0, st>>>(p, M, N, rows, lddy, ldx, N, 0, kchunk, cs);
  X2_LAUNCH_OK();
  const int64_t MN = (int64_t)M * N;
  (void)MN;
  k_splitk_reduce<<<splitk_reduce_blocks(M, N, db != nullptr), 256, 0, st>>>(partial, cs, splits, M, N, dW, lddw, db);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // namespace x2
