// potrf_panel.cuh -- Cholesky panel kernel: POTF2 of one
// 128x128 diagonal block entirely in shared memory + registers, fused with
//   * sum_i log L_ii of the block (written per block, summed later in fixed order),
//   * the failure flag that replaces LAPACK's info / arma::chol's bool
//     (reference: arma::chol -> dpotrf, armadillo op_chol_meat.hpp:45-70),
//   * the explicit inverse of the diagonal block, written into the diagonal
//     block of the W buffer (strict upper zeroed).  The inverse turns the panel
//     TRSM, the triangular solves and TRTRI's leaves into DMMA GEMMs / GEMVs.
//
// Layout in smem: As[c*128 + r] = A[r, c] (column-major).  128 threads; thread
// r owns row r.  The block is processed in four 32-column sub-panels:
//   1. left-looking update of the sub-panel with the previous sub-panels
//      (row r in registers, multipliers broadcast from smem);
//   2. the warp owning the 32x32 diagonal sub-block factors it with
//      warp-shuffle broadcasts only (no block barrier inside the 32 columns);
//   3. the rows below do their 32-wide triangular solve in registers.
// The inverse is computed by forward substitution, thread c owning column c of
// L^-1, stored in the (unused) upper triangle of the same smem array.
#pragma once
#include "common.cuh"

namespace lk {

constexpr int POTF2_THREADS = 128;
constexpr int POTF2_SMEM_BYTES = 128 * 128 * 8 + 2 * 128 * 8 + 64;

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, double* __restrict__ W, long long ld, int jb, double* __restrict__ logdet_blocks,
                 int blk_index, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* As = sm;                  // 128*128
  double* diag = sm + 128 * 128;    // L_ii
  double* rdiag = diag + 128;       // 1 / L_ii
  const int r = threadIdx.x;
  const int w = r >> 5, lane = r & 31;
  double* Ablk = A + (long long)jb * ld + jb;

  for (int c = 0; c < 128; ++c) As[c * 128 + r] = Ablk[(long long)c * ld + r];
  __syncthreads();

  bool ok = true;
  for (int kb = 0; kb < 4; ++kb) {
    const int c0 = 32 * kb;
    double a[32];
    if (r >= c0) {
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = As[(c0 + j) * 128 + r];
      for (int k = 0; k < c0; ++k) {
        const double lr = As[k * 128 + r];
        const double2* row = reinterpret_cast<const double2*>(&As[k * 128 + c0]);
#pragma unroll
        for (int j2 = 0; j2 < 16; ++j2) {
          const double2 v = row[j2];
          a[2 * j2] -= lr * v.x;
          a[2 * j2 + 1] -= lr * v.y;
        }
      }
    }
    if (w == kb) {
      double logsum = 0.0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double piv = __shfl_sync(0xffffffffu, a[j], j);
        ok = ok && (piv > 0.0);
        const double dj = sqrt(piv);
        const double rd = 1.0 / dj;
        a[j] = (lane == j) ? dj : a[j] * rd;
        if (lane == j) logsum = log(dj);
#pragma unroll
        for (int j2 = j + 1; j2 < 32; ++j2) {
          const double l2 = __shfl_sync(0xffffffffu, a[j], j2);  // L[c0+j2, c0+j]
          a[j2] -= a[j] * l2;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) As[(c0 + j) * 128 + r] = (j <= lane) ? a[j] : 0.0;
      // own diagonal entry: a[lane] (static indexing via select chain)
      double dself = 0.0;
#pragma unroll
      for (int j = 0; j < 32; ++j) dself = (lane == j) ? a[j] : dself;
      diag[r] = dself;
      rdiag[r] = 1.0 / dself;
      const double ls = warp_sum(logsum);
      if (lane == 0) {
        // per 32-column partial; the 4 partials of the block are added in order below
        rdiag[128 + kb] = ls;  // scratch slots after rdiag (64 spare bytes)
      }
    }
    __syncthreads();
    if (r >= c0 + 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double s = a[j];
#pragma unroll
        for (int i = 0; i < j; ++i) s -= a[i] * As[(c0 + i) * 128 + c0 + j];
        a[j] = s * rdiag[c0 + j];
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) As[(c0 + j) * 128 + r] = a[j];
    }
    __syncthreads();
  }

  // failure flag (NaN-safe: piv > 0 is false for NaN)
  if (!__all_sync(0xffffffffu, ok) && lane == 0) atomicExch(info, 1);
  if (r == 0) logdet_blocks[blk_index] = ((rdiag[128] + rdiag[129]) + rdiag[130]) + rdiag[131];

  // write L back (lower part incl. diagonal; strict upper zero)
  for (int c = 0; c < 128; ++c) Ablk[(long long)c * ld + r] = (r >= c) ? As[c * 128 + r] : 0.0;
  __syncthreads();

  // ---- inverse: thread c owns column c of X = L^-1; X[k, c] kept at As[k*128 + c] (k >= c) ----
  {
    const int c = r;
    for (int rb = w; rb < 4; ++rb) {
      const int R0 = 32 * rb;
      double acc[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) acc[i] = (R0 + i == c) ? 1.0 : 0.0;
      for (int k = 32 * w; k < R0; ++k) {
        const double xk = (k >= c) ? As[k * 128 + c] : 0.0;
        const double2* col = reinterpret_cast<const double2*>(&As[k * 128 + R0]);  // L[R0+i, k]
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          const double2 v = col[i2];
          acc[2 * i2] -= v.x * xk;
          acc[2 * i2 + 1] -= v.y * xk;
        }
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double s = acc[i];
#pragma unroll
        for (int i2 = 0; i2 < i; ++i2) s -= As[(R0 + i2) * 128 + R0 + i] * acc[i2];  // L[R0+i, R0+i2] * x[i2]
        acc[i] = (R0 + i >= c) ? s * rdiag[R0 + i] : 0.0;
      }
      // NOTE: the strict-lower L entries As[k*128 + row] (row > k) read above are never
      // overwritten: X[row, c] goes to As[row*128 + c] with c <= row, i.e. the upper triangle.
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (R0 + i >= c) As[(R0 + i) * 128 + c] = acc[i];
    }
  }
  __syncthreads();
  double* Wblk = W + (long long)jb * ld + jb;
  for (int c = 0; c < 128; ++c) Wblk[(long long)c * ld + r] = (r >= c) ? As[r * 128 + c] : 0.0;
}

}  // namespace lk
