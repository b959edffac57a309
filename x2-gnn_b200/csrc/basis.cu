// basis.cu -- fused envelope / radial / spherical-Bessel x spherical-harmonic basis kernels.
// Replaces envelop.py:16-21, radial_basis_layer.py:36-40 and angular_basis_layer.py:80-93 (~150
// elementwise launches of sympy-lambdified closures in the reference) with one E-scale table
// kernel and one T-scale streaming kernel.
//
// Numerics: the reference evaluates closed-form j_l polynomials in 1/x in fp32, which cancel
// catastrophically for l >= 4.  Here everything that is E-scale (envelope,
// j_l) or O(L) per triplet (Legendre recurrence) runs in fp64 and is rounded to fp32 once, so
// results track the fp64 evaluation of the reference formulas to ~1 ulp.  The T-scale part is a
// pure HBM stream: read theta[T], idx[T] and an L2-resident table row, write sbf[T, L*R].
#include "common.cuh"

namespace x2 {

constexpr int kMaxL = 32;

__device__ __forceinline__ double envelope_d(double d, double inv_cutoff, int p, double a,
                                             double b, double c) {
  const double x = d * inv_cutoff;
  double xp = 1.0;  // x^(p-1)
  for (int i = 0; i < p - 1; ++i) xp *= x;
  return 1.0 / x + xp * (a + x * (b + x * c));
}

__global__ void k_envelope(const float* __restrict__ d, int64_t n, float inv_cutoff, int p, float a,
                           float b, float c, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  // the reference forms x = d * (1/cutoff) in fp32 (envelop.py:17); keep that product, then
  // evaluate the polynomial in fp64
  const float xs = d[i] * inv_cutoff;
  out[i] = (float)envelope_d((double)xs, 1.0, p, (double)a, (double)b, (double)c);
}

__global__ void k_radial_fwd(const float* __restrict__ d, const float* __restrict__ freq,
                             const float* __restrict__ env, int64_t n, int R, float inv_cutoff,
                             float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * R) return;
  const int64_t e = idx / R;
  const int r = (int)(idx - e * R);
  const float ds = d[e] * inv_cutoff;              // radial_basis_layer.py:37
  float v = sinf(freq[r] * ds);                    // :40
  if (env) v *= env[e];
  out[idx] = v;
}

// Stage 1 of the frequency gradient: each warp owns a fixed, contiguous row range and keeps one
// running sum per basis index (lane r, r+32, ...); also writes grad_d if requested.
__global__ void k_radial_bwd_partial(const float* __restrict__ d, const float* __restrict__ freq,
                                     const float* __restrict__ env, const float* __restrict__ go,
                                     int64_t n, int R, float inv_cutoff, int64_t rows_per_warp,
                                     int64_t nwarps, float* __restrict__ partial,
                                     float* __restrict__ grad_d) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= nwarps) return;
  const int64_t r0 = w * rows_per_warp;
  const int64_t r1 = min(n, r0 + rows_per_warp);
  float acc0 = 0.f, acc1 = 0.f;  // R <= 64
  const float f0 = lane < R ? freq[lane] : 0.f;
  const float f1 = lane + 32 < R ? freq[lane + 32] : 0.f;
  for (int64_t e = r0; e < r1; ++e) {
    const float ds = d[e] * inv_cutoff;
    const float ev = env ? env[e] : 1.f;
    float gd = 0.f;
    if (lane < R) {
      const float g = go[e * R + lane] * ev * cosf(f0 * ds);
      acc0 += g * ds;
      gd += g * f0;
    }
    if (lane + 32 < R) {
      const float g = go[e * R + lane + 32] * ev * cosf(f1 * ds);
      acc1 += g * ds;
      gd += g * f1;
    }
    if (grad_d) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) gd += __shfl_xor_sync(0xffffffffu, gd, o);
      if (lane == 0) grad_d[e] = gd * inv_cutoff;
    }
  }
  if (lane < R) partial[w * R + lane] = acc0;
  if (lane + 32 < R) partial[w * R + lane + 32] = acc1;
}

__global__ void k_radial_bwd_final(const float* __restrict__ partial, int64_t nwarps, int R,
                                   float* __restrict__ grad_freq) {
  const int r = threadIdx.x;
  if (r >= R) return;
  float s = 0.f;
  for (int64_t w = 0; w < nwarps; ++w) s += partial[w * R + r];  // fixed order => deterministic
  grad_freq[r] = s;
}

// spherical Bessel j_l(x), fp64.  Power series where upward recurrence is unstable (x < ~l).
__device__ double sph_jl(int l, double x) {
  const double thr = fmax(1.0, 0.75 * (double)l);
  if (x < thr) {
    // j_l(x) = x^l/(2l+1)!! * sum_k (-x^2/2)^k / (k! (2l+3)(2l+5)...(2l+2k+1))
    double pref = 1.0;
    for (int i = 1; i <= l; ++i) pref *= x / (double)(2 * i + 1);
    const double q = -0.5 * x * x;
    double term = 1.0, sum = 1.0;
    for (int k = 1; k < 80; ++k) {
      term *= q / ((double)k * (double)(2 * l + 2 * k + 1));
      sum += term;
      if (fabs(term) < 1e-18 * fabs(sum)) break;
    }
    return pref * sum;
  }
  double s, c;
  sincos(x, &s, &c);
  const double inv = 1.0 / x;
  double jm = s * inv;                    // j_0
  if (l == 0) return jm;
  double j = (s * inv - c) * inv;         // j_1
  for (int n = 1; n < l; ++n) {
    const double jn = (double)(2 * n + 1) * inv * j - jm;
    jm = j;
    j = jn;
  }
  return j;
}

__global__ void k_sbf_table(const float* __restrict__ d, int64_t n, int L, int R,
                            const float* __restrict__ zeros, const float* __restrict__ norm,
                            float cutoff, float env_cutoff, int p, float a, float b, float c,
                            float* __restrict__ table) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int S = L * R;
  if (idx >= n * S) return;
  const int64_t e = idx / S;
  const int col = (int)(idx - e * S);
  const int l = col / R;
  const double dd = (double)d[e];
  const double x = (double)zeros[col] * (dd / (double)cutoff);   // angular_basis_layer.py:81
  const double env = envelope_d(dd, 1.0 / (double)env_cutoff, p, (double)a, (double)b, (double)c);
  table[idx] = (float)(env * (double)norm[col] * sph_jl(l, x));
}

// Y_l0(theta) = sqrt((2l+1)/(4 pi)) P_l(cos theta), l < L, three-term Legendre recurrence
// (basis_func.py:94-96), fp64.
__device__ __forceinline__ void y_l0(double theta, int L, float* out, int stride) {
  const double c = cos(theta);
  double p0 = 1.0, p1 = c;
  const double inv4pi = 0.07957747154594767;
  out[0] = (float)sqrt(inv4pi);
  if (L > 1) out[stride] = (float)(sqrt(3.0 * inv4pi) * c);
  for (int j = 2; j < L; ++j) {
    const double pj = ((double)(2 * j - 1) * c * p1 - (double)(j - 1) * p0) / (double)j;
    p0 = p1;
    p1 = pj;
    out[j * stride] = (float)(sqrt((double)(2 * j + 1) * inv4pi) * pj);
  }
}

constexpr int kSbfTile = 256;  // triplets per block (= threads per block)

// sbf[t, l R + n] = table[idx[t], l R + n] * Y_l0(theta_t)       (angular_basis_layer.py:80-93)
// Phase 1: one thread per triplet evaluates Y_l0 in fp64 (normalisation constants from a per-block
// table, no per-triplet sqrt).  Phase 2: the block's [256, S] output tile is contiguous; consecutive
// threads produce consecutive 2-column units (the (triplet, column) cursor advances incrementally --
// no division in the loop): coalesced 64-bit table-row loads and output stores.
// Write-bandwidth bound: 4 (S+1) + 8 bytes per triplet.
__global__ void __launch_bounds__(kSbfTile)
k_sbf_fwd(const float* __restrict__ table, const float* __restrict__ angles,
          const int64_t* __restrict__ idx, int64_t T, int64_t E, int L, int R,
          float* __restrict__ out) {
  extern __shared__ __align__(16) float sbf_smem[];          // [kSbfTile][L] ys | [S] column -> l (as int)
  __shared__ int64_t rowoff[kSbfTile];
  __shared__ float ynorm[kMaxL];
  const int S = L * R;
  float* ys = sbf_smem;
  int* lcol = reinterpret_cast<int*>(sbf_smem + kSbfTile * L);
  const int64_t t0 = (int64_t)blockIdx.x * kSbfTile;
  const int nt = (int)min((int64_t)kSbfTile, T - t0);
  if (threadIdx.x < L) ynorm[threadIdx.x] = (float)sqrt((double)(2 * threadIdx.x + 1) * 0.07957747154594767);
  for (int c = threadIdx.x; c < S; c += kSbfTile) lcol[c] = c / R;
  __syncthreads();
  if (threadIdx.x < nt) {
    // the Legendre recurrence is stable on [-1, 1]: fp32 (error ~1e-7 per step) is enough here, unlike
    // the Bessel part, which lives in the fp64 table
    const float c = cosf(angles[t0 + threadIdx.x]);
    float p0 = 1.f, p1 = c;
    float* o = ys + threadIdx.x * L;
    o[0] = ynorm[0];
    if (L > 1) o[1] = ynorm[1] * c;
    for (int j = 2; j < L; ++j) {
      const float pj = ((float)(2 * j - 1) * c * p1 - (float)(j - 1) * p0) / (float)j;
      p0 = p1;
      p1 = pj;
      o[j] = ynorm[j] * pj;
    }
    const int64_t r = idx[t0 + threadIdx.x];
    rowoff[threadIdx.x] = ((r < 0 || r >= E) ? 0 : r) * S;  // bounds are validated by the caller
  }
  __syncthreads();
  float* o = out + t0 * S;
  if ((S & 1) == 0 && (R & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0 &&
      (reinterpret_cast<uintptr_t>(table) & 7) == 0) {
    // even S and R (config: 42, 6): units of 2 consecutive columns never straddle a triplet or an l
    // block; 64-bit table loads and output stores, cursor advanced without divisions
    const int S2 = S >> 1;
    const int total2 = nt * S2;
    const int dq = kSbfTile / S2, dr = kSbfTile - dq * S2;
    int u = threadIdx.x;
    int tb = u / S2, c2 = u - tb * S2;             // the only division: once per thread
    for (; u < total2; u += kSbfTile) {
      const float2 tv = __ldg(reinterpret_cast<const float2*>(table + rowoff[tb]) + c2);
      const float y = ys[tb * L + lcol[2 * c2]];
      reinterpret_cast<float2*>(o + (size_t)tb * S)[c2] = make_float2(tv.x * y, tv.y * y);
      tb += dq;
      c2 += dr;
      if (c2 >= S2) { c2 -= S2; ++tb; }
    }
    return;
  }
  const int total = nt * S;
  int i = threadIdx.x;
  int tb = i / S, col = i - tb * S;
  const int dq = kSbfTile / S, dr = kSbfTile - dq * S;
  for (; i < total; i += kSbfTile) {
    o[i] = __ldg(table + rowoff[tb] + col) * ys[tb * L + lcol[col]];
    tb += dq;
    col += dr;
    if (col >= S) { col -= S; ++tb; }
  }
}


__global__ void k_angular_fwd(const float* __restrict__ angles, int64_t T, int L,
                              float* __restrict__ out) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  y_l0((double)angles[t], L, out + t * L, 1);
}

// ------------------------------------------------------------------ gradients w.r.t. the geometry
// The reference's bases are plain torch expressions (sympy-lambdified closures), so autograd differentiates them
// w.r.t. distances and angles (force training differentiates the energy w.r.t. positions through them).  The
// kernels below are the analytic derivatives of the same formulas, E-scale parts in fp64 like the forward.

// u'(x) of envelop.py:16-21: -1/x^2 + (p-1) a x^(p-2) + x^(p-1) (p b + (p+1) c x)
__device__ __forceinline__ double envelope_dx(double x, int p, double a, double b, double c) {
  double xp = 1.0;  // x^(p-2)   (p >= 2; for p = 1 the a-term is constant and drops out)
  for (int i = 0; i < p - 2; ++i) xp *= x;
  const double xp1 = p >= 2 ? xp * x : 1.0;   // x^(p-1)
  const double t_a = p >= 2 ? (double)(p - 1) * a * xp : 0.0;
  return -1.0 / (x * x) + t_a + xp1 * ((double)p * b + (double)(p + 1) * c * x);
}

__global__ void k_envelope_bwd(const float* __restrict__ d, const float* __restrict__ go, int64_t n,
                               float inv_cutoff, int p, float a, float b, float c, float* __restrict__ gd) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float xs = d[i] * inv_cutoff;
  gd[i] = (float)((double)go[i] * envelope_dx((double)xs, p, (double)a, (double)b, (double)c) * (double)inv_cutoff);
}

// d/dtheta Y_l0(theta) = -sin(theta) sqrt((2l+1)/(4 pi)) P_l'(cos theta), P_l' = P_{l-2}' + (2l-1) P_{l-1}
// (no division by 1 - c^2: regular at the poles).  Lane l of the calling warp gets its own order.
__device__ __forceinline__ void y_l0_and_dtheta(double theta, int l, double* y, double* dy) {
  double s, c;
  sincos(theta, &s, &c);
  double p0 = 1.0, p1 = c, q0 = 0.0, q1 = 1.0;     // P_0, P_1, P_0', P_1'
  double P = l == 0 ? p0 : p1, Q = l == 0 ? q0 : q1;
  for (int j = 2; j <= l; ++j) {
    const double pj = ((double)(2 * j - 1) * c * p1 - (double)(j - 1) * p0) / (double)j;
    const double qj = q0 + (double)(2 * j - 1) * p1;
    p0 = p1; p1 = pj; q0 = q1; q1 = qj;
    P = pj; Q = qj;
  }
  const double nrm = sqrt((double)(2 * l + 1) * 0.07957747154594767);
  *y = nrm * P;
  *dy = -s * nrm * Q;
}

__global__ void k_angular_bwd(const float* __restrict__ angles, const float* __restrict__ go, int64_t T, int L,
                              float* __restrict__ gang) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const double th = (double)angles[t];
  double acc = 0.0;
  for (int l = 0; l < L; ++l) {
    double y, dy;
    y_l0_and_dtheta(th, l, &y, &dy);
    acc += (double)go[t * L + l] * dy;
  }
  gang[t] = (float)acc;
}

constexpr int kSbfBwdWarps = 8;       // warps per block of the two sbf gradient kernels
constexpr int kSbfBwdMaxS = 256;      // L R of the gradient kernels (8 column slots per lane)

// d sbf / d theta: one warp per triplet.  g_theta[t] = sum_col go[t, col] table[idx[t], col] Y_l'(theta_t)
__global__ void __launch_bounds__(kSbfBwdWarps * 32)
k_sbf_bwd_theta(const float* __restrict__ table, const float* __restrict__ angles, const int64_t* __restrict__ idx,
                const float* __restrict__ go, int64_t T, int64_t E, int L, int R, float* __restrict__ gang) {
  __shared__ float dys[kSbfBwdWarps][kMaxL];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * kSbfBwdWarps + w;
  if (t >= T) return;
  const int S = L * R;
  if (lane < L) {
    double y, dy;
    y_l0_and_dtheta((double)angles[t], lane, &y, &dy);
    dys[w][lane] = (float)dy;
  }
  __syncwarp();
  int64_t r = idx[t];
  r = (r < 0 || r >= E) ? 0 : r;
  const float* trow = table + r * S;
  const float* grow = go + t * S;
  float acc = 0.f;
  for (int col = lane; col < S; col += 32) acc += grow[col] * __ldg(trow + col) * dys[w][col / R];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) gang[t] = acc;
}

// d sbf / d d: one warp per bond e, its triplets (order[rowptr[e] .. rowptr[e+1])) in a fixed order =>
// deterministic, no atomics.  g_table[e, col] = sum_t go[t, col] Y_l(theta_t), then
// g_d[e] = sum_col g_table[e, col] d table[e, col] / d d, with
// d table / d d = N (u'(d / c_env) / c_env j_l(x) + u(d / c_env) j_l'(x) z / cutoff),  x = z d / cutoff,
// j_l' = j_{l-1} - (l + 1) / x j_l  (j_0' = -j_1), all fp64.
__global__ void __launch_bounds__(kSbfBwdWarps * 32)
k_sbf_bwd_d(const float* __restrict__ d, const float* __restrict__ angles, const int64_t* __restrict__ order,
            const int64_t* __restrict__ rowptr, const float* __restrict__ go, int64_t E, int L, int R,
            const float* __restrict__ zeros, const float* __restrict__ norm, float cutoff, float env_cutoff, int p,
            float a, float b, float c, float* __restrict__ gd) {
  __shared__ float ys[kSbfBwdWarps][kMaxL];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * kSbfBwdWarps + w;
  if (e >= E) return;
  const int S = L * R;
  float acc[kSbfBwdMaxS / 32];
#pragma unroll
  for (int j = 0; j < kSbfBwdMaxS / 32; ++j) acc[j] = 0.f;
  for (int64_t k = rowptr[e]; k < rowptr[e + 1]; ++k) {
    const int64_t t = order[k];
    __syncwarp();
    if (lane < L) {
      double y, dy;
      y_l0_and_dtheta((double)angles[t], lane, &y, &dy);
      ys[w][lane] = (float)y;
    }
    __syncwarp();
    const float* grow = go + t * S;
#pragma unroll
    for (int j = 0; j < kSbfBwdMaxS / 32; ++j) {
      const int col = lane + 32 * j;
      if (col < S) acc[j] += grow[col] * ys[w][col / R];
    }
  }
  const double dd = (double)d[e];
  const double xe = dd / (double)env_cutoff;
  const double env = envelope_d(dd, 1.0 / (double)env_cutoff, p, (double)a, (double)b, (double)c);
  const double denv = envelope_dx(xe, p, (double)a, (double)b, (double)c) / (double)env_cutoff;
  double g = 0.0;
#pragma unroll
  for (int j = 0; j < kSbfBwdMaxS / 32; ++j) {
    const int col = lane + 32 * j;
    if (col < S) {
      const int l = col / R;
      const double zc = (double)zeros[col] / (double)cutoff;
      const double x = zc * dd;
      const double jl = sph_jl(l, x);
      const double djl = l == 0 ? -sph_jl(1, x) : sph_jl(l - 1, x) - (double)(l + 1) / x * jl;
      g += (double)acc[j] * (double)norm[col] * (denv * jl + env * djl * zc);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
  if (lane == 0) gd[e] = (float)g;
}

// ------------------------------------------------------------------ geometry of the line graph (xgnn.py:46,60-66)
// bond lengths d_e = |pos[a0] - pos[a1]| and triplet angles theta_t = atan2(|ji x jk|, <ji, jk>) with
// ji = pos[i] - pos[j], jk = pos[k] - pos[j]: ~15 gather / elementwise / reduction launches of the caller as two
// kernels (no [T, 3] intermediates).  fp32, the reference's operation order; products and sums are not contracted
// into FMAs across the reference's kernel boundaries (a difference is exact, a product is rounded before it is added).
__global__ void k_bond_lengths(const float* __restrict__ pos, const int64_t* __restrict__ a0,
                               const int64_t* __restrict__ a1, int64_t E, float* __restrict__ d) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float* p = pos + 3 * a0[e];
  const float* q = pos + 3 * a1[e];
  const float x = p[0] - q[0], y = p[1] - q[1], z = p[2] - q[2];
  d[e] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
}

__global__ void k_triplet_angles(const float* __restrict__ pos, const int64_t* __restrict__ ai,
                                 const int64_t* __restrict__ aj, const int64_t* __restrict__ ak, int64_t T,
                                 float* __restrict__ ang) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const float* pj = pos + 3 * aj[t];
  const float* pi = pos + 3 * ai[t];
  const float* pk = pos + 3 * ak[t];
  const float jx = pj[0], jy = pj[1], jz = pj[2];
  const float ax = pi[0] - jx, ay = pi[1] - jy, az = pi[2] - jz;      // ji
  const float bx = pk[0] - jx, by = pk[1] - jy, bz = pk[2] - jz;      // jk
  const float c = __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
  const float cx = __fsub_rn(__fmul_rn(ay, bz), __fmul_rn(az, by));
  const float cy = __fsub_rn(__fmul_rn(az, bx), __fmul_rn(ax, bz));
  const float cz = __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
  const float s = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz)));
  ang[t] = atan2f(s, c);
}

}  // namespace x2

using namespace x2;

extern "C" {

int x2_envelope_fwd(const float* d, int64_t n, float inv_cutoff, int32_t p, float a, float b,
                    float c, float* out, void* stream) {
  X2_CHECK_ARG(n >= 0 && p >= 1, "x2_envelope_fwd: bad arguments");
  if (n == 0) return X2_OK;
  k_envelope<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(d, n, inv_cutoff, p, a, b, c, out);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_radial_fwd(const float* d, const float* freq, const float* env, int64_t n, int32_t R,
                  float inv_cutoff, float* out, void* stream) {
  X2_CHECK_ARG(n >= 0 && R >= 1, "x2_radial_fwd: bad arguments");
  if (n == 0) return X2_OK;
  k_radial_fwd<<<(unsigned)cdiv(n * R, 256), 256, 0, (cudaStream_t)stream>>>(d, freq, env, n, R, inv_cutoff, out);
  X2_LAUNCH_OK();
  return X2_OK;
}

static int64_t radial_bwd_warps(int64_t n) {
  int64_t w = cdiv(n, 64);                 // >= 64 rows per warp
  const int64_t cap = (int64_t)kNumSM * 32;  // 148 SMs x 32 warps
  if (w > cap) w = cap;
  if (w < 1) w = 1;
  return w;
}

size_t x2_radial_bwd_workspace_bytes(int64_t n, int32_t R) {
  return align_up((size_t)radial_bwd_warps(n) * (size_t)R * sizeof(float), 256) + 256;
}

int x2_radial_bwd(const float* d, const float* freq, const float* env, const float* grad_out,
                  int64_t n, int32_t R, float inv_cutoff, float* grad_freq, float* grad_d,
                  void* ws, size_t ws_bytes, void* stream) {
  X2_CHECK_ARG(n >= 0 && R >= 1 && R <= 64, "x2_radial_bwd: need 1 <= R <= 64 (got %d)", R);
  cudaStream_t st = (cudaStream_t)stream;
  if (ws_bytes < x2_radial_bwd_workspace_bytes(n, R)) { set_error("x2_radial_bwd: workspace too small"); return X2_EWORKSPACE; }
  if (n == 0) {
    X2_CUDA_OK(cudaMemsetAsync(grad_freq, 0, R * sizeof(float), st));
    return X2_OK;
  }
  const int64_t nw = radial_bwd_warps(n);
  const int64_t rpw = cdiv(n, nw);
  float* partial = static_cast<float*>(ws);
  k_radial_bwd_partial<<<(unsigned)cdiv(nw * 32, 128), 128, 0, st>>>(d, freq, env, grad_out, n, R, inv_cutoff,
                                                                   rpw, nw, partial, grad_d);
  X2_LAUNCH_OK();
  k_radial_bwd_final<<<1, 64, 0, st>>>(partial, nw, R, grad_freq);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_sbf_table(const float* d, int64_t n, int32_t L, int32_t R, const float* zeros,
                 const float* norm, float cutoff, float env_cutoff, int32_t p, float a, float b,
                 float c, float* table, void* stream) {
  X2_CHECK_ARG(n >= 0 && L >= 1 && L <= kMaxL && R >= 1 && R <= 64, "x2_sbf_table: need 1<=L<=%d, 1<=R<=64", kMaxL);
  if (n == 0) return X2_OK;
  k_sbf_table<<<(unsigned)cdiv(n * L * R, 128), 128, 0, (cudaStream_t)stream>>>(d, n, L, R, zeros, norm, cutoff,
                                                                            env_cutoff, p, a, b, c, table);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_sbf_fwd(const float* table, const float* angles, const int64_t* idx, int64_t T, int64_t E,
               int32_t L, int32_t R, float* out, void* stream) {
  X2_CHECK_ARG(T >= 0 && L >= 1 && L <= kMaxL && R >= 1 && R <= 64, "x2_sbf_fwd: need 1<=L<=%d, 1<=R<=64", kMaxL);
  if (T == 0) return X2_OK;
  const size_t smem = (size_t)(kSbfTile * L + L * R) * sizeof(float);
  k_sbf_fwd<<<(unsigned)cdiv(T, kSbfTile), kSbfTile, smem, (cudaStream_t)stream>>>(table, angles, idx, T, E, L, R, out);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_angular_fwd(const float* angles, int64_t T, int32_t L, float* out, void* stream) {
  X2_CHECK_ARG(T >= 0 && L >= 1 && L <= kMaxL, "x2_angular_fwd: need 1<=L<=%d", kMaxL);
  if (T == 0) return X2_OK;
  k_angular_fwd<<<(unsigned)cdiv(T, 256), 256, 0, (cudaStream_t)stream>>>(angles, T, L, out);
  X2_LAUNCH_OK();
  return X2_OK;
}

// ---- gradients w.r.t. distances / angles (autograd of the reference's torch expressions; force training)
int x2_envelope_bwd(const float* d, const float* grad_out, int64_t n, float inv_cutoff, int32_t p, float a, float b,
                    float c, float* grad_d, void* stream) {
  X2_CHECK_ARG(n >= 0 && p >= 1, "x2_envelope_bwd: bad arguments");
  if (n == 0) return X2_OK;
  k_envelope_bwd<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(d, grad_out, n, inv_cutoff, p, a, b, c, grad_d);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_angular_bwd(const float* angles, const float* grad_out, int64_t T, int32_t L, float* grad_angles, void* stream) {
  X2_CHECK_ARG(T >= 0 && L >= 1 && L <= kMaxL, "x2_angular_bwd: need 1<=L<=%d", kMaxL);
  if (T == 0) return X2_OK;
  k_angular_bwd<<<(unsigned)cdiv(T, 256), 256, 0, (cudaStream_t)stream>>>(angles, grad_out, T, L, grad_angles);
  X2_LAUNCH_OK();
  return X2_OK;
}

// grad_angles [T] and / or grad_d [E] of sbf = x2_sbf_fwd(x2_sbf_table(d), angles, idx) given grad_out [T, L R].
// order [T] / rowptr [E + 1]: the triplets grouped by idx (stable), needed for grad_d only.
int x2_sbf_bwd(const float* d, const float* table, const float* angles, const int64_t* idx, const int64_t* order,
               const int64_t* rowptr, const float* grad_out, int64_t T, int64_t E, int32_t L, int32_t R,
               const float* zeros, const float* norm, float cutoff, float env_cutoff, int32_t p, float a, float b,
               float c, float* grad_d, float* grad_angles, void* stream) {
  X2_CHECK_ARG(T >= 0 && E >= 0 && L >= 1 && L <= kMaxL && R >= 1 && R <= 64 && L * R <= kSbfBwdMaxS,
               "x2_sbf_bwd: need 1<=L<=%d, 1<=R<=64, L*R<=%d", kMaxL, kSbfBwdMaxS);
  X2_CHECK_ARG(!grad_d || (order && rowptr), "x2_sbf_bwd: grad_d needs the triplets grouped by idx (order, rowptr)");
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_angles && T > 0) {
    k_sbf_bwd_theta<<<(unsigned)cdiv(T, kSbfBwdWarps), kSbfBwdWarps * 32, 0, st>>>(table, angles, idx, grad_out, T, E,
                                                                                 L, R, grad_angles);
    X2_LAUNCH_OK();
  }
  if (grad_d && E > 0) {
    k_sbf_bwd_d<<<(unsigned)cdiv(E, kSbfBwdWarps), kSbfBwdWarps * 32, 0, st>>>(d, angles, order, rowptr, grad_out, E, L,
                                                                             R, zeros, norm, cutoff, env_cutoff, p, a,
                                                                             b, c, grad_d);
    X2_LAUNCH_OK();
  }
  return X2_OK;
}

// ---- geometry of the line graph (xgnn.py:46,60-66): indices are the caller's int64 tensors, bounds are the caller's
int x2_bond_lengths(const float* pos, const int64_t* a0, const int64_t* a1, int64_t E, float* d, void* stream) {
  X2_CHECK_ARG(E >= 0 && (E == 0 || (pos && a0 && a1 && d)), "x2_bond_lengths: bad arguments");
  if (E == 0) return X2_OK;
  k_bond_lengths<<<(unsigned)cdiv(E, 256), 256, 0, (cudaStream_t)stream>>>(pos, a0, a1, E, d);
  X2_LAUNCH_OK();
  return X2_OK;
}

int x2_triplet_angles(const float* pos, const int64_t* ai, const int64_t* aj, const int64_t* ak, int64_t T, float* ang,
                      void* stream) {
  X2_CHECK_ARG(T >= 0 && (T == 0 || (pos && ai && aj && ak && ang)), "x2_triplet_angles: bad arguments");
  if (T == 0) return X2_OK;
  k_triplet_angles<<<(unsigned)cdiv(T, 256), 256, 0, (cudaStream_t)stream>>>(pos, ai, aj, ak, T, ang);
  X2_LAUNCH_OK();
  return X2_OK;
}

}  // extern "C"
