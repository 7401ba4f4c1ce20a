// cov.cuh -- pair kernels: everything that the
// reference does through the d x n^2 matrix dX and std::function callbacks is
// regenerated here from an X tile pair staged in shared memory.
//   K1  covariance build, lower-triangle tiles only
//       (cholCov build loop src/lib/LinearAlgebra.cpp:148-194 + Cov_* src/lib/Covariance.cpp:24-161)
//   K8  fused gradient reduction: R_ij * dln rho/dtheta_k regenerated per pair and contracted
//       with x_i x_j and with R^-1 (or R^-1 - U U^T for LMP) without materialising dR/dtheta_k
//       (compute_ll_grad_theta_vecs src/lib/KrigingImpl.cpp:855-885, DlnCovDtheta_* Covariance.cpp:38-179)
//   K9  theta-bounds pair reduction (Optim::theta_bounds src/lib/Optim.cpp:179-209)
// Tiles are 64 x 64 pairs, 256 threads, 4 x 4 pairs per thread with the row
// index contiguous (32-byte accesses per thread and column).
#pragma once
#include "common.cuh"

namespace lk {

constexpr int PT = 64;           // pair tile edge
constexpr int PAIR_THREADS = 256;
constexpr int LK_MAX_D = 64;
constexpr double LK_SQRT3 = 1.7320508075688772;
constexpr double LK_SQRT5 = 2.2360679774997898;

struct KernelParams {
  double inv_theta[LK_MAX_D];
};

// ln-free correlation factor pieces.  The pair kernels stage X pre-scaled by c_K / theta_k (c = 1, 1, sqrt3, sqrt5),
// so per pair and dimension they see  u = c_K (x_i - x_j) / theta_k  directly:
//   gauss:  rho = exp(-0.5 sum u^2)
//   exp:    rho = exp(-sum |u|)
//   m32:    rho = prod(1+s) exp(-sum s),            s = |u| = sqrt3 |dx|/theta   == exp(-sum(s - log1p(s)))
//   m52:    rho = prod(1+s+s^2/3) exp(-sum s),      s = |u| = sqrt5 |dx|/theta   == exp(-sum(s - log1p(s+s^2/3)))
// Every per-pair-per-dimension step is a handful of FP64 pipe operations (ncu of round 1's kernels: FP64 pipe 57 %
// busy, the rest issue slots -- the divisions of 1 + s + s^2/3 and of the log-derivatives were most of both).
template <int KERNEL>
__host__ __device__ constexpr double corr_scale() {
  return KERNEL == 2 ? LK_SQRT3 : (KERNEL == 3 ? LK_SQRT5 : 1.0);
}
// 1 / x for x in [1, 1e300): hardware seed (rcp.approx.ftz.f64, ~2^-23) + two Newton steps = full precision to the
// last ulp or two -- no special-case paths, 5 FP64 operations against the ~30 of an IEEE division.
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}
template <int KERNEL>
__device__ __forceinline__ void corr_accum(double u, double& esum, double& prod) {
  if (KERNEL == 0) {
    esum = fma(u, u, esum);
  } else if (KERNEL == 1) {
    esum += fabs(u);
  } else if (KERNEL == 2) {
    const double s = fabs(u);
    esum += s;
    prod *= (1.0 + s);
  } else {
    const double s = fabs(u);
    esum += s;
    prod *= fma(s, fma(s, 1.0 / 3.0, 1.0), 1.0);
  }
}
template <int KERNEL>
__device__ __forceinline__ double corr_finish(double esum, double prod) {
  if (KERNEL == 0) return exp(-0.5 * esum);
  if (KERNEL == 1) return exp(-esum);
  return prod * exp(-esum);
}
// theta_k * dln rho / dtheta_k as a function of the scaled u  (the 1/theta_k is applied by the host at the end):
//   gauss u^2 ; exp |u| ; m32 s^2 / (1 + s) ; m52 (1 + s) (s^2/3) / (1 + s + s^2/3) = (1 + s) s^2 / (3 (1 + s) + s^2)
template <int KERNEL>
__device__ __forceinline__ double dlnrho_times_theta(double u) {
  if (KERNEL == 0) return u * u;
  if (KERNEL == 1) return fabs(u);
  const double s = fabs(u);
  const double a = 1.0 + s, s2 = s * s;
  if (KERNEL == 2) return s2 * fast_rcp(a);
  return (a * s2) * fast_rcp(fma(3.0, a, s2));
}

// lower-triangle tile id -> (ti >= tj)
__device__ __forceinline__ void tri_tile(int id, int& ti, int& tj) {
  int t = (int)((sqrt(8.0 * (double)id + 1.0) - 1.0) * 0.5);
  while ((long long)t * (t + 1) / 2 > id) --t;
  while ((long long)(t + 1) * (t + 2) / 2 <= id) ++t;
  ti = t;
  tj = id - t * (t + 1) / 2;
}

// stage X rows [r0, r0+64) scaled by c_K / theta into smem as xs[k*64 + r]
template <int KERNEL>
__device__ __forceinline__ void stage_x(const double* __restrict__ X, int n, int d, const KernelParams& kp, int r0,
                                        double* xs) {
  for (int e = threadIdx.x; e < d * PT; e += PAIR_THREADS) {
    const int k = e / PT, r = e % PT;
    const int row = r0 + r;
    xs[e] = (row < n) ? X[(long long)k * n + row] * (kp.inv_theta[k] * corr_scale<KERNEL>()) : 0.0;
  }
}

// ---------------------------------------------------------------------------
// K1: A[i, j] = alpha * rho_ij (i != j), A[i, i] = 1 (+ noise_i * inv_sigma2) + diag_add ; padding = identity.
// Tiles with ti >= tj only; the diagonal tile is written in full (symmetric).
// Partial builds (block extension of a kept factor, LinearAlgebra::update_cholCov, LinearAlgebra.cpp:206-243):
// tile = tri_tile(blockIdx-strided id + id0) shifted by `shift` on both axes -- id0 = t0 (t0 + 1) / 2, shift = 0
// builds the row tiles >= t0 over all their columns; id0 = 0, shift = t0 builds the lower-right block only.
// Rows below diag_split receive diag_add_lo (the jitter the kept factor was accepted with) instead of diag_add.
// ---------------------------------------------------------------------------
template <int KERNEL>
__global__ void __launch_bounds__(PAIR_THREADS)
cov_build_kernel(const double* __restrict__ X, int n, int d, const __grid_constant__ KernelParams kp, double alpha,
                 const double* __restrict__ noise, double inv_sigma2, double diag_add, double* __restrict__ A,
                 long long ld, int ntiles, int id0, int shift, int diag_split, double diag_add_lo) {
  extern __shared__ double sm[];
  double* xi = sm;
  double* xj = sm + d * PT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int ti, tj;
    tri_tile(tile + id0, ti, tj);
    ti += shift;
    tj += shift;
    __syncthreads();
    stage_x<KERNEL>(X, n, d, kp, ti * PT, xi);
    stage_x<KERNEL>(X, n, d, kp, tj * PT, xj);
    __syncthreads();
    double es[4][4], pr[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        es[a][b] = 0.0;
        pr[a][b] = 1.0;
      }
    for (int k = 0; k < d; ++k) {
      double vi[4], vj[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) vi[a] = xi[k * PT + tx * 4 + a];
#pragma unroll
      for (int b = 0; b < 4; ++b) vj[b] = xj[k * PT + ty * 4 + b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) corr_accum<KERNEL>(vi[a] - vj[b], es[a][b], pr[a][b]);
    }
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = tj * PT + ty * 4 + b;
      double v[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int i = ti * PT + tx * 4 + a;
        double r;
        if (i >= n || j >= n) {
          r = (i == j) ? 1.0 : 0.0;
        } else if (i == j) {
          r = 1.0 + (noise ? noise[i] * inv_sigma2 : 0.0) + (i < diag_split ? diag_add_lo : diag_add);
        } else {
          r = alpha * corr_finish<KERNEL>(es[a][b], pr[a][b]);
        }
        v[a] = r;
      }
      double* dst = A + (long long)j * ld + ti * PT + tx * 4;
      *reinterpret_cast<double2*>(dst) = make_double2(v[0], v[1]);
      *reinterpret_cast<double2*>(dst + 2) = make_double2(v[2], v[3]);
    }
  }
}

// ---------------------------------------------------------------------------
// K8: per-CTA partial sums over strictly-lower pairs i > j:
//   S1[k] = sum c_ij R_ij f_k(u_ij,k)           S2[k] = sum Wt_ij R_ij f_k(u_ij,k),   k < d
//   S1[d] = sum c_ij R_ij                       S2[d] = sum Wt_ij R_ij               (f == 1)
// c_ij = (a_i b_j + b_i a_j) / 2  (a == b == x for LL / LMP: x_i x_j; a = v, b = x for the LOO gradient),
// with R_ij = alpha rho_ij, f_k = theta_k dln rho/dtheta_k, Wt = Rinv (LL) or Rinv - U U^T (LMP, pdim > 0).
// The host applies the factors 2 / -2 and 1/theta_k.  partial layout: [cta][2][d+1].
// ---------------------------------------------------------------------------
template <int KERNEL>
__global__ void __launch_bounds__(PAIR_THREADS)
grad_reduce_kernel(const double* __restrict__ X, int n, int d, const __grid_constant__ KernelParams kp, double alpha,
                   const double* __restrict__ V, long long ld, const double* __restrict__ avec,
                   const double* __restrict__ bvec, const double* __restrict__ U, long long ldu, int pdim, double* __restrict__ partial, int ntiles) {
  extern __shared__ double sm[];
  double* xi = sm;
  double* xj = xi + d * PT;
  double* xvi = xj + d * PT;       // 64  a_i
  double* xvj = xvi + PT;          // 64  a_j
  double* yvi = xvj + PT;          // 64  b_i
  double* yvj = yvi + PT;          // 64  b_j
  double* wacc = yvj + PT;         // [8 warps][2][d+1]
  double* ui = wacc + 8 * 2 * (d + 1);  // pdim*64
  double* uj = ui + pdim * PT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < 8 * 2 * (d + 1); e += PAIR_THREADS) wacc[e] = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int ti, tj;
    tri_tile(tile, ti, tj);
    __syncthreads();
    stage_x<KERNEL>(X, n, d, kp, ti * PT, xi);
    stage_x<KERNEL>(X, n, d, kp, tj * PT, xj);
    if (threadIdx.x < PT) {
      const int row = ti * PT + threadIdx.x;
      xvi[threadIdx.x] = (row < n) ? avec[row] : 0.0;
      yvi[threadIdx.x] = (row < n) ? bvec[row] : 0.0;
    } else if (threadIdx.x < 2 * PT) {
      const int row = tj * PT + threadIdx.x - PT;
      xvj[threadIdx.x - PT] = (row < n) ? avec[row] : 0.0;
      yvj[threadIdx.x - PT] = (row < n) ? bvec[row] : 0.0;
    }
    for (int e = threadIdx.x; e < pdim * PT; e += PAIR_THREADS) {
      const int q = e / PT, r = e % PT;
      ui[e] = (ti * PT + r < n) ? U[(long long)q * ldu + ti * PT + r] : 0.0;
      uj[e] = (tj * PT + r < n) ? U[(long long)q * ldu + tj * PT + r] : 0.0;
    }
    __syncthreads();
    double w1[4][4], w2[4][4];
    {
      double es[4][4], pr[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          es[a][b] = 0.0;
          pr[a][b] = 1.0;
        }
      for (int k = 0; k < d; ++k) {
        double vi[4], vj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) vi[a] = xi[k * PT + tx * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) vj[b] = xj[k * PT + ty * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) corr_accum<KERNEL>(vi[a] - vj[b], es[a][b], pr[a][b]);
      }
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int j = tj * PT + ty * 4 + b;
        const double* vcol = V + (long long)j * ld + ti * PT + tx * 4;
        const double2 v01 = *reinterpret_cast<const double2*>(vcol);
        const double2 v23 = *reinterpret_cast<const double2*>(vcol + 2);
        double vv[4] = {v01.x, v01.y, v23.x, v23.y};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
          const int i = ti * PT + tx * 4 + a;
          const bool valid = (i > j) && (i < n);
          const double r = valid ? alpha * corr_finish<KERNEL>(es[a][b], pr[a][b]) : 0.0;
          double wt = vv[a];
          for (int q = 0; q < pdim; ++q) wt -= ui[q * PT + tx * 4 + a] * uj[q * PT + ty * 4 + b];
          w1[a][b] = 0.5 * (xvi[tx * 4 + a] * yvj[ty * 4 + b] + yvi[tx * 4 + a] * xvj[ty * 4 + b]) * r;
          w2[a][b] = valid ? wt * r : 0.0;
        }
      }
    }
    for (int k = 0; k <= d; ++k) {
      double s1 = 0.0, s2 = 0.0;
      if (k < d) {
        double vi[4], vj[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) vi[a] = xi[k * PT + tx * 4 + a];
#pragma unroll
        for (int b = 0; b < 4; ++b) vj[b] = xj[k * PT + ty * 4 + b];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double f = dlnrho_times_theta<KERNEL>(vi[a] - vj[b]);
            s1 = fma(w1[a][b], f, s1);
            s2 = fma(w2[a][b], f, s2);
          }
      } else {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            s1 += w1[a][b];
            s2 += w2[a][b];
          }
      }
      s1 = warp_sum(s1);
      s2 = warp_sum(s2);
      if (lane == 0) {
        wacc[(warp * 2 + 0) * (d + 1) + k] += s1;
        wacc[(warp * 2 + 1) * (d + 1) + k] += s2;
      }
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < 2 * (d + 1); e += PAIR_THREADS) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += wacc[w * 2 * (d + 1) + e];
    partial[(long long)blockIdx.x * 2 * (d + 1) + e] = s;
  }
}

// diagonal sums: out partial [cta][4] = { sum x_i^2, sum V_ii, sum noise_i V_ii, sum noise_i x_i^2 }
__global__ void __launch_bounds__(256)
diag_sums_kernel(const double* __restrict__ V, long long ld, const double* __restrict__ xvec,
                 const double* __restrict__ noise, int n, double* __restrict__ partial) {
  __shared__ double sh[8][4];
  double s[4] = {0, 0, 0, 0};
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double x = xvec[i], v = V[(long long)i * ld + i], nz = noise ? noise[i] : 0.0;
    s[0] += x * x;
    s[1] += v;
    s[2] += nz * v;
    s[3] += nz * x * x;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    s[q] = warp_sum(s[q]);
    if (lane == 0) sh[warp][q] = s[q];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w][threadIdx.x];
    partial[blockIdx.x * 4 + threadIdx.x] = t;
  }
}

// ---------------------------------------------------------------------------
// K9: theta bounds.  Pass 1 (max |dx_k| per dim) is separable: max_k - min_k of X.
// Pass 2: per-CTA partial sums over unordered pairs i > j (the reference sums over
// ordered pairs: both w and w|dx| double, the ratio is unchanged):
//   w_ij = (y_i - y_j)^2 / ||x_i - x_j||^2  (NaN -> 0) ;  partial[cta][k] = sum w_ij |dx_ij,k| (k<d), [d] = sum w_ij
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(PAIR_THREADS)
theta_bounds_kernel(const double* __restrict__ X, const double* __restrict__ y, int n, int d,
                    double* __restrict__ partial, int ntiles) {
  extern __shared__ double sm[];
  double* xi = sm;
  double* xj = xi + d * PT;
  double* yi = xj + d * PT;
  double* yj = yi + PT;
  double* wacc = yj + PT;  // [8][d+1]
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < 8 * (d + 1); e += PAIR_THREADS) wacc[e] = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int ti, tj;
    tri_tile(tile, ti, tj);
    __syncthreads();
    for (int e = threadIdx.x; e < d * PT; e += PAIR_THREADS) {
      const int k = e / PT, r = e % PT;
      xi[e] = (ti * PT + r < n) ? X[(long long)k * n + ti * PT + r] : 0.0;
      xj[e] = (tj * PT + r < n) ? X[(long long)k * n + tj * PT + r] : 0.0;
    }
    if (threadIdx.x < PT) yi[threadIdx.x] = (ti * PT + threadIdx.x < n) ? y[ti * PT + threadIdx.x] : 0.0;
    else if (threadIdx.x < 2 * PT)
      yj[threadIdx.x - PT] = (tj * PT + threadIdx.x - PT < n) ? y[tj * PT + threadIdx.x - PT] : 0.0;
    __syncthreads();
    double w[4][4];
    {
      double d2[4][4];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) d2[a][b] = 0.0;
      for (int k = 0; k < d; ++k)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const double dx = xi[k * PT + tx * 4 + a] - xj[k * PT + ty * 4 + b];
            d2[a][b] = fma(dx, dx, d2[a][b]);
          }
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int i = ti * PT + tx * 4 + a, j = tj * PT + ty * 4 + b;
          const double dy = yi[tx * 4 + a] - yj[ty * 4 + b];
          double v = (dy * dy) / d2[a][b];
          if (v != v) v = 0.0;  // replace(nan, 0)
          w[a][b] = (i > j && i < n) ? v : 0.0;
        }
    }
    for (int k = 0; k <= d; ++k) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double f = (k < d) ? fabs(xi[k * PT + tx * 4 + a] - xj[k * PT + ty * 4 + b]) : 1.0;
          s = fma(w[a][b], f, s);
        }
      s = warp_sum(s);
      if (lane == 0) wacc[warp * (d + 1) + k] += s;
    }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < d + 1; e += PAIR_THREADS) {
    double s = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) s += wacc[w8 * (d + 1) + e];
    partial[(long long)blockIdx.x * (d + 1) + e] = s;
  }
}

// per-dimension min / max of X (one CTA per dimension)
__global__ void __launch_bounds__(256)
col_minmax_kernel(const double* __restrict__ X, int n, double* __restrict__ mn, double* __restrict__ mx) {
  __shared__ double s0[8], s1[8];
  const double* c = X + (long long)blockIdx.x * n;
  double lo = 1e300, hi = -1e300;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    lo = fmin(lo, c[i]);
    hi = fmax(hi, c[i]);
  }
  lo = -warp_max(-lo);
  hi = warp_max(hi);
  if ((threadIdx.x & 31) == 0) {
    s0[threadIdx.x >> 5] = lo;
    s1[threadIdx.x >> 5] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) {
      lo = fmin(lo, s0[w]);
      hi = fmax(hi, s1[w]);
    }
    mn[blockIdx.x] = lo;
    mx[blockIdx.x] = hi;
  }
}

}  // namespace lk
