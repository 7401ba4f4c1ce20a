// potrf_panel.cuh -- Cholesky panel kernel: POTF2 of one 128x128 diagonal block
// entirely in shared memory, fused with
//   * sum_i log L_ii of the block (written per block, summed later in fixed order),
//   * the failure flag that replaces LAPACK's info / arma::chol's bool
//     (reference: arma::chol -> dpotrf, armadillo op_chol_meat.hpp:45-70),
//   * the explicit inverse of the diagonal block, written into the diagonal block of the W buffer
//     (strict upper zeroed).  The inverse turns the panel TRSM, the triangular sweeps and TRTRI's leaves
//     into DMMA GEMMs / tile products.
//
// This kernel sits on the critical path of the factorisation (one launch per 128 columns), so it is
// organised for latency, 512 threads:
//   for each 32-column sub-panel:
//     A. one warp factors the 32x32 diagonal sub-block with warp shuffles only (lane = row; the pivot's
//        1/sqrt comes from one rsqrt + two Newton steps instead of a sqrt and a divide) and inverts it
//        (lane = column, forward substitution against broadcast shared-memory reads);
//     B. all threads form the rows below as A21 * inv(L11)^T -- no dependency chain;
//     C. all threads apply the rank-32 update to the rest of the tile (4x4 register micro-tiles,
//        lower part only).
//   The off-diagonal 32x32 blocks of the inverse follow from X_ij = -X_ii sum_k L_ik X_kj, one block
//   sub-diagonal at a time, as small all-thread products.
// Shared memory: As[c*128 + r] = A[r, c] (column-major, 128 KB) + the 10 lower 32x32 blocks of the
// inverse in natural orientation (80 KB); T blocks of the inverse stage live in As's unused upper part.
#pragma once
#include "common.cuh"

namespace lk {

constexpr int POTF2_THREADS = 512;
constexpr int POTF2_SMEM_DOUBLES = 128 * 128 + 10 * 1024 + 128 + 8;
constexpr int POTF2_SMEM_BYTES = POTF2_SMEM_DOUBLES * 8;

__device__ __forceinline__ int potf2_blk(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, double* __restrict__ W, long long ld, int jb, double* __restrict__ logdet_blocks,
                 int blk_index, int* __restrict__ info) {
  extern __shared__ double sm[];
  double* As = sm;                     // 128*128
  double* Xb = sm + 128 * 128;         // 10 blocks of 32x32: Xb[blk][c*32 + r] = X[32i + r, 32j + c]
  double* rdiag = Xb + 10 * 1024;      // 1 / L_ii
  double* logp = rdiag + 128;          // 4 partial log sums
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  double* Ablk = A + (long long)jb * ld + jb;

  for (int idx = tid; idx < 128 * 64; idx += POTF2_THREADS) {
    const int c = idx >> 6, r2 = (idx & 63) * 2;
    *reinterpret_cast<double2*>(As + c * 128 + r2) = *reinterpret_cast<const double2*>(Ablk + (long long)c * ld + r2);
  }
  __syncthreads();

  bool ok = true;
  for (int kb = 0; kb < 4; ++kb) {
    const int c0 = 32 * kb;
    // ---------------- A: diagonal 32x32 sub-block, one warp ----------------
    if (warp == 0) {
      double a[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = As[(c0 + j) * 128 + c0 + lane];
      double logsum = 0.0, rd_self = 0.0;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double piv = __shfl_sync(0xffffffffu, a[j], j);
        ok = ok && (piv > 0.0);
        double rd = rsqrt(piv);
        double dj = piv * rd;
        dj = fma(fma(-dj, dj, piv) * 0.5, rd, dj);  // Newton step: dj = sqrt(piv) to working accuracy
        rd = fma(fma(-dj, rd, 1.0), rd, rd);        // rd = 1 / dj
        a[j] = (lane == j) ? dj : a[j] * rd;
        if (lane == j) {
          logsum = log(dj);
          rd_self = rd;
        }
#pragma unroll
        for (int j2 = j + 1; j2 < 32; ++j2) {
          const double l2 = __shfl_sync(0xffffffffu, a[j], j2);  // L[c0+j2, c0+j]
          a[j2] -= a[j] * l2;
        }
      }
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j <= lane) As[(c0 + j) * 128 + c0 + lane] = a[j];
      rdiag[c0 + lane] = rd_self;
      const double ls = warp_sum(logsum);
      if (lane == 0) logp[kb] = ls;
      __syncwarp();
      // inverse of the sub-block: lane = column c of X, x[i] = X[i, c]
      double x[32];
      double* Xd = Xb + potf2_blk(kb, kb) * 1024;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < i; ++k) s = fma(As[(c0 + k) * 128 + c0 + i], x[k], s);
        const double rdi = rdiag[c0 + i];
        x[i] = (i == lane) ? rdi : ((i > lane) ? -s * rdi : 0.0);
        Xd[lane * 32 + i] = x[i];
      }
    }
    __syncthreads();
    const int R0 = c0 + 32;
    const int m = 128 - R0;  // rows / columns left after this sub-panel
    if (m > 0) {
      // ---------------- B: L21 = A21 * inv(L11)^T ----------------
      {
        const int rl = tid & 127, part = tid >> 7;  // row, group of 8 output columns
        const bool act = rl < m;
        double in[32], out[8];
        if (act) {
#pragma unroll
          for (int k = 0; k < 32; ++k) in[k] = As[(c0 + k) * 128 + R0 + rl];
          const double* Xd = Xb + potf2_blk(kb, kb) * 1024;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              // L21[r, j] = sum_{k <= j} A21[r, k] X[j, k] ; X[j, k] = Xd[k*32 + j], zero above the diagonal
              s = fma(in[k], Xd[k * 32 + part * 8 + jj], s);
            }
            out[jj] = s;
          }
        }
        __syncthreads();
        if (act) {
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) As[(c0 + part * 8 + jj) * 128 + R0 + rl] = out[jj];
        }
      }
      __syncthreads();
      // ---------------- C: rank-32 update of the remaining lower part ----------------
      const int mq = m >> 2;
      for (int mt = tid; mt < mq * mq; mt += POTF2_THREADS) {
        const int tr = mt % mq, tc = mt / mq;
        if (tr < tc) continue;
        const int r0 = R0 + 4 * tr, q0 = R0 + 4 * tc;
        double acc[4][4];
#pragma unroll
        for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
          for (int b_ = 0; b_ < 4; ++b_) acc[a_][b_] = 0.0;
#pragma unroll 8
        for (int k = 0; k < 32; ++k) {
          const double* col = As + (c0 + k) * 128;
          const double2 r01 = *reinterpret_cast<const double2*>(col + r0);
          const double2 r23 = *reinterpret_cast<const double2*>(col + r0 + 2);
          const double2 c01 = *reinterpret_cast<const double2*>(col + q0);
          const double2 c23 = *reinterpret_cast<const double2*>(col + q0 + 2);
          const double rv[4] = {r01.x, r01.y, r23.x, r23.y};
          const double cv[4] = {c01.x, c01.y, c23.x, c23.y};
#pragma unroll
          for (int a_ = 0; a_ < 4; ++a_)
#pragma unroll
            for (int b_ = 0; b_ < 4; ++b_) acc[a_][b_] = fma(rv[a_], cv[b_], acc[a_][b_]);
        }
#pragma unroll
        for (int b_ = 0; b_ < 4; ++b_)
#pragma unroll
          for (int a_ = 0; a_ < 4; ++a_)
            if (r0 + a_ >= q0 + b_) As[(q0 + b_) * 128 + r0 + a_] -= acc[a_][b_];
      }
      __syncthreads();
    }
  }

  // failure flag (NaN-safe: piv > 0 is false for NaN) and the block's log-determinant part
  if (warp == 0 && lane == 0) {
    if (!ok) atomicExch(info, 1);
    logdet_blocks[blk_index] = ((logp[0] + logp[1]) + logp[2]) + logp[3];
  }

  // write L back (lower part incl. diagonal; strict upper zero)
  for (int idx = tid; idx < 128 * 64; idx += POTF2_THREADS) {
    const int c = idx >> 6, r2 = (idx & 63) * 2;
    double2 v = *reinterpret_cast<const double2*>(As + c * 128 + r2);
    if (r2 < c) v.x = 0.0;
    if (r2 + 1 < c) v.y = 0.0;
    *reinterpret_cast<double2*>(Ablk + (long long)c * ld + r2) = v;
  }

  // ---------------- off-diagonal blocks of the inverse, one block sub-diagonal at a time ----------------
  // T_ij = sum_{k=j}^{i-1} L_ik X_kj  is parked in As's upper block (j, i);  X_ij = -X_ii T_ij.
  for (int dd = 1; dd < 4; ++dd) {
    const int nblk = 4 - dd;
    for (int e = tid; e < nblk * 1024; e += POTF2_THREADS) {
      const int b = e >> 10, rr = e & 31, cc = (e >> 5) & 31;
      const int j = b, i = b + dd;
      double s = 0.0;
      for (int k = j; k < i; ++k) {
        const double* Lik = As + (32 * k) * 128 + 32 * i + rr;          // L[32i+rr, 32k+kk] = Lik[kk*128]
        const double* Xkj = Xb + potf2_blk(k, j) * 1024 + cc * 32;      // X[32k+kk, 32j+cc] = Xkj[kk]
#pragma unroll 8
        for (int kk = 0; kk < 32; ++kk) s = fma(Lik[kk * 128], Xkj[kk], s);
      }
      As[(32 * i + cc) * 128 + 32 * j + rr] = s;  // T[rr, cc] in the upper block (j, i)
    }
    __syncthreads();
    for (int e = tid; e < nblk * 1024; e += POTF2_THREADS) {
      const int b = e >> 10, rr = e & 31, cc = (e >> 5) & 31;
      const int j = b, i = b + dd;
      const double* Xii = Xb + potf2_blk(i, i) * 1024 + rr;              // X[32i+rr, 32i+kk] = Xii[kk*32]
      const double* T = As + (32 * i + cc) * 128 + 32 * j;               // T[kk, cc] = T[kk]
      double s = 0.0;
#pragma unroll 8
      for (int kk = 0; kk < 32; ++kk) s = fma(Xii[kk * 32], T[kk], s);
      Xb[potf2_blk(i, j) * 1024 + cc * 32 + rr] = -s;
    }
    __syncthreads();
  }

  // write the inverse into W's diagonal block (strict upper zero)
  double* Wblk = W + (long long)jb * ld + jb;
  for (int idx = tid; idx < 128 * 64; idx += POTF2_THREADS) {
    const int c = idx >> 6, r2 = (idx & 63) * 2;
    const int i = r2 >> 5, j = c >> 5;
    double2 v = make_double2(0.0, 0.0);
    if (i >= j) v = *reinterpret_cast<const double2*>(Xb + potf2_blk(i, j) * 1024 + (c & 31) * 32 + (r2 & 31));
    *reinterpret_cast<double2*>(Wblk + (long long)c * ld + r2) = v;
  }
}

}  // namespace lk
