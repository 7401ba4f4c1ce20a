// trsv.cuh -- HBM-bound side stages: blocked
// triangular solves with a few right-hand sides, using the diagonal-block
// inverses kept in W's diagonal blocks; 1-norms for the condition estimate;
// the p x p GLS (Gram, Cholesky, beta) on one CTA; deterministic reductions.
// Reference: solve_lower / solve_upper (src/lib/LinearAlgebra.cpp:695-701) at
// src/lib/KrigingImpl.cpp:102-103, 122 and src/lib/Kriging.cpp:294.
#pragma once
#include "common.cuh"

namespace lk {

constexpr int TRSV_MAX_RHS = 8;  // right-hand sides processed per sweep

// Forward sweep step j  (L z = b, in place in B (ldb = N, nrhs columns)):
//   for row blocks i > j:  b_i -= L[i, j] z_j ;  block i == j+1 then applies z_i = Dinv_i b_i.
// Launch with j = -1 (grid 1) to start: z_0 = Dinv_0 b_0.
// grid = (number of row blocks below j) ; 128 threads, thread r <-> row r of the block.
__global__ void __launch_bounds__(128)
trsv_fwd_step_kernel(const double* __restrict__ L, const double* __restrict__ W, long long ld, double* __restrict__ B,
                     long long ldb, int nrhs, int j) {
  __shared__ double zs[128 * TRSV_MAX_RHS];
  const int r = threadIdx.x;
  const int i = j + 1 + blockIdx.x;
  const long long ib = (long long)i * 128;
  double acc[TRSV_MAX_RHS];
#pragma unroll
  for (int q = 0; q < TRSV_MAX_RHS; ++q) acc[q] = (q < nrhs) ? B[q * ldb + ib + r] : 0.0;
  if (j >= 0) {
    const long long jb = (long long)j * 128;
    for (int q = 0; q < nrhs; ++q) zs[q * 128 + r] = B[q * ldb + jb + r];
    __syncthreads();
    const double* Lp = L + jb * ld + ib + r;
#pragma unroll 8
    for (int k = 0; k < 128; ++k) {
      const double l = Lp[(long long)k * ld];
#pragma unroll
      for (int q = 0; q < TRSV_MAX_RHS; ++q)
        if (q < nrhs) acc[q] -= l * zs[q * 128 + k];
    }
  }
  if (i == j + 1) {
    __syncthreads();
    for (int q = 0; q < nrhs; ++q) zs[q * 128 + r] = acc[q];
    __syncthreads();
    const double* Dp = W + ib * ld + ib + r;  // Dinv[r, k], k <= r
    double z[TRSV_MAX_RHS];
#pragma unroll
    for (int q = 0; q < TRSV_MAX_RHS; ++q) z[q] = 0.0;
    for (int k = 0; k <= r; ++k) {
      const double dv = Dp[(long long)k * ld];
#pragma unroll
      for (int q = 0; q < TRSV_MAX_RHS; ++q)
        if (q < nrhs) z[q] += dv * zs[q * 128 + k];
    }
#pragma unroll
    for (int q = 0; q < TRSV_MAX_RHS; ++q) acc[q] = z[q];
  }
  for (int q = 0; q < nrhs; ++q) B[q * ldb + ib + r] = acc[q];
}

// Backward sweep step j  (L^T x = e, in place):
//   for column blocks k < j:  e_k -= L[j, k]^T x_j ;  block k == j-1 then applies x_k = Dinv_k^T e_k.
// Launch with j = nblk (grid 1, blockIdx -> k = nblk-1) to start: x_last = Dinv_last^T e_last.
// 128 threads = 4 warps; each warp reduces columns with lanes striding the 128 rows (coalesced).
__global__ void __launch_bounds__(128)
trsv_bwd_step_kernel(const double* __restrict__ L, const double* __restrict__ W, long long ld, double* __restrict__ B,
                     long long ldb, int nrhs, int j, int nblk) {
  __shared__ double xs[128 * TRSV_MAX_RHS];
  __shared__ double es[128 * TRSV_MAX_RHS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k = (j >= nblk) ? (nblk - 1) : (j - 1 - (int)blockIdx.x);
  const long long kb = (long long)k * 128;
  for (int q = 0; q < nrhs; ++q) es[q * 128 + tid] = B[q * ldb + kb + tid];
  if (j < nblk) {
    const long long jb = (long long)j * 128;
    for (int q = 0; q < nrhs; ++q) xs[q * 128 + tid] = B[q * ldb + jb + tid];
    __syncthreads();
    for (int c = warp; c < 128; c += 4) {
      const double* Lc = L + (kb + c) * ld + jb;
      const double l0 = Lc[lane], l1 = Lc[lane + 32], l2 = Lc[lane + 64], l3 = Lc[lane + 96];
      for (int q = 0; q < nrhs; ++q) {
        const double* x = xs + q * 128;
        double s = l0 * x[lane] + l1 * x[lane + 32] + l2 * x[lane + 64] + l3 * x[lane + 96];
        s = warp_sum(s);
        if (lane == 0) es[q * 128 + c] -= s;
      }
    }
  }
  __syncthreads();
  if (k == j - 1 || j >= nblk) {
    // x_k = Dinv_k^T e_k :  x[c] = sum_{r >= c} Dinv[r, c] e[r]
    for (int c = warp; c < 128; c += 4) {
      const double* Dc = W + (kb + c) * ld + kb;
      double d[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int rr = lane + 32 * u;
        d[u] = (rr >= c) ? Dc[rr] : 0.0;
      }
      for (int q = 0; q < nrhs; ++q) {
        const double* e = es + q * 128;
        double s = d[0] * e[lane] + d[1] * e[lane + 32] + d[2] * e[lane + 64] + d[3] * e[lane + 96];
        s = warp_sum(s);
        if (lane == 0) xs[q * 128 + c] = s;
      }
    }
    __syncthreads();
    for (int q = 0; q < nrhs; ++q) B[q * ldb + kb + tid] = xs[q * 128 + tid];
  } else {
    for (int q = 0; q < nrhs; ++q) B[q * ldb + kb + tid] = es[q * 128 + tid];
  }
}

// max_j sum_{i >= j} |M[i, j]| over the lower triangle, two-stage deterministic:
// stage 1: one warp per column -> colsum[j] ; stage 2 on one CTA -> max.
__global__ void __launch_bounds__(256)
tri_colsum_abs_kernel(const double* __restrict__ M, long long ld, int n, double* __restrict__ colsum) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double* c = M + (long long)warp * ld;
  double s = 0.0;
  for (int i = warp + lane; i < n; i += 32) s += fabs(c[i]);
  s = warp_sum(s);
  if (lane == 0) colsum[warp] = s;
}

// NaN-propagating maximum (fmax drops NaN operands: a factor with NaN entries would otherwise report a finite
// norm and pass the isfinite guards of Engine::eval / factor_update).
__device__ __forceinline__ double nanmax(double a, double b) { return (a != a || b != b) ? NAN : fmax(a, b); }
__global__ void __launch_bounds__(256) vec_max_kernel(const double* __restrict__ v, int n, double* __restrict__ out) {
  __shared__ double sh[8];
  double m = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = nanmax(m, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = nanmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) m = nanmax(m, sh[w]);
    *out = m;
  }
}

// Gram partials of Z (n x q, ldz): for every pair (a <= b) partial[chunk][pair] = sum_{rows in chunk} Z[r,a] Z[r,b].
constexpr int GRAM_CHUNK = 2048;
__global__ void __launch_bounds__(256)
gram_partial_kernel(const double* __restrict__ Z, long long ldz, int n, int q, double* __restrict__ partial) {
  const int npairs = q * (q + 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * GRAM_CHUNK;
  const int r1 = min(n, r0 + GRAM_CHUNK);
  for (int pr = warp; pr < npairs; pr += 8) {
    // pair index -> (a, b), a <= b, enumerated column by column of the upper triangle
    int b = 0;
    while ((b + 1) * (b + 2) / 2 <= pr) ++b;
    const int a = pr - b * (b + 1) / 2;
    const double* za = Z + (long long)a * ldz;
    const double* zb = Z + (long long)b * ldz;
    double s = 0.0;
    for (int r = r0 + lane; r < r1; r += 32) s += za[r] * zb[r];
    s = warp_sum(s);
    if (lane == 0) partial[(long long)blockIdx.x * npairs + pr] = s;
  }
}

// One CTA: reduce the Gram partials of Z = [Fstar | ystar] (q = p + 1) in fixed order, then
//   Rstar = chol_upper(Fstar' Fstar)           (src/lib/KrigingImpl.cpp:107-108)
//   beta  = Rstar^-1 Rstar^-T Fstar' ystar      (src/lib/KrigingImpl.cpp:116-119)
// out: Rstar (p x p, column-major, strict lower zero), beta (p), gls_info (1 = not PD).
__global__ void __launch_bounds__(256)
gls_final_kernel(const double* __restrict__ partial, int nchunks, int p, double* __restrict__ Rstar,
                 double* __restrict__ beta, int* __restrict__ gls_info) {
  extern __shared__ double sm[];
  const int q = p + 1;
  const int npairs = q * (q + 1) / 2;
  double* G = sm;           // q x q (upper part used), column-major
  double* h = sm + q * q;   // p
  for (int pr = threadIdx.x; pr < npairs; pr += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(long long)c * npairs + pr];
    int b = 0;
    while ((b + 1) * (b + 2) / 2 <= pr) ++b;
    const int a = pr - b * (b + 1) / 2;
    G[b * q + a] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // dpotrf('U') semantics on the leading p x p block: G = U^T U
    int bad = 0;
    for (int j = 0; j < p; ++j) {
      double s = G[j * q + j];
      for (int k = 0; k < j; ++k) s -= G[j * q + k] * G[j * q + k];
      if (!(s > 0.0)) bad = 1;
      const double ujj = sqrt(s);
      G[j * q + j] = ujj;
      for (int c = j + 1; c < p; ++c) {
        double t = G[c * q + j];
        for (int k = 0; k < j; ++k) t -= G[j * q + k] * G[c * q + k];
        G[c * q + j] = t / ujj;
      }
    }
    *gls_info = bad;
    // h = Fstar' ystar = column p of the Gram matrix ; solve U^T t = h ; U beta = t
    for (int i = 0; i < p; ++i) {
      double s = G[p * q + i];
      for (int k = 0; k < i; ++k) s -= G[i * q + k] * h[k];
      h[i] = s / G[i * q + i];
    }
    for (int i = p - 1; i >= 0; --i) {
      double s = h[i];
      for (int k = i + 1; k < p; ++k) s -= G[k * q + i] * beta[k];
      beta[i] = s / G[i * q + i];
    }
    for (int c = 0; c < p; ++c)
      for (int r = 0; r < p; ++r) Rstar[c * p + r] = (r <= c) ? G[c * q + r] : 0.0;
  }
}

// resid = y - F beta  (n rows), written into column `col` of the rhs buffer.
__global__ void residual_kernel(const double* __restrict__ y, const double* __restrict__ F, long long ldf, int n, int p,
                                const double* __restrict__ beta, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int c = 0; c < p; ++c) s += F[(long long)c * ldf + i] * beta[c];
  out[i] = y[i] - s;
}

// out[0] = sum of partial[0..m) in fixed order (single thread: m is small).
__global__ void sum_partials_kernel(const double* __restrict__ partial, int m, int stride, int count,
                                    double* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  double s = 0.0;
  for (int c = 0; c < m; ++c) s += partial[(long long)c * stride + k];
  out[k] = s;
}

}  // namespace lk
