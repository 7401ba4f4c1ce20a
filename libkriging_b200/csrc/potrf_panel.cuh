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
//   for each 32-column sub-panel (c0 = 0, 32, 64, 96):
//     AB. one thread per row r >= c0 keeps its 32 entries of the sub-panel in registers and runs the 32 column
//         steps left-looking: v = a[j] - sum_k a[k] L[c0+j, c0+k].  Every thread recomputes the pivot
//         p = A[jj] - sum_k L[c0+j, c0+k]^2 from the same shared-memory row, so ONE named barrier per column is
//         enough (the diagonal rows publish their new entry, everybody reads the finished row j at the next
//         step); 1/sqrt(p) is one rsqrt + two Newton steps instead of a sqrt and a divide.
//         Meanwhile an otherwise idle warp inverts the previous 32x32 diagonal sub-block (lane = column).
//     C.  all threads apply the rank-32 update to the rest of the tile (4x4 register micro-tiles, lower part).
//   The off-diagonal 32x32 blocks of the inverse follow from X_ij = -X_ii sum_k L_ik X_kj, one block
//   sub-diagonal at a time, as small all-thread products.
// Shared memory: As[c*128 + r] = A[r, c] (column-major, 128 KB) + the 10 lower 32x32 blocks of the inverse in
// natural orientation (80 KB) + two row-major copies of the current / previous diagonal sub-block (16 KB);
// T blocks of the inverse stage live in As's unused upper part.
#pragma once
#include "common.cuh"

namespace lk {

constexpr int POTF2_THREADS = 512;
constexpr int POTF2_SMEM_DOUBLES = 128 * 128 + 10 * 1024 + 2 * 1024 + 128 + 128 + 8;
constexpr int POTF2_SMEM_BYTES = POTF2_SMEM_DOUBLES * 8;

__device__ __forceinline__ int potf2_blk(int i, int j) { return i * (i + 1) / 2 + j; }  // i >= j

// -DLKGPU_POTF2_PROFILE: thread 0 stamps clock64() at the phase boundaries of the panel kernel (read back with
// lkgpu_debug_potf2_profile; tools/potf2_phases.py).  Not compiled into the product library.
#ifdef LKGPU_POTF2_PROFILE
__device__ long long g_potf2_prof[32];
#define POTF2_STAMP(k) do { if (threadIdx.x == 0) g_potf2_prof[k] = clock64(); } while (0)
#else
#define POTF2_STAMP(k) do { } while (0)
#endif

__device__ __forceinline__ void potf2_bar(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// inverse of the 32x32 lower-triangular diagonal sub-block kb, by one warp: lane = column c of X.
// Ld: row-major copy of the sub-block's strict lower part (Ld[i*32 + k] = L[c0+i, c0+k], k < i).
__device__ __forceinline__ void potf2_invert_diag(const double* __restrict__ Ld, const double* __restrict__ rdiag_c0,
                                                  double* __restrict__ Xd, int lane) {
  double x[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    // four partial sums: the dot product's dependent chain is i / 4 FMAs deep instead of i / 2
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k + 3 < i; k += 4) {
      const double2 l = *reinterpret_cast<const double2*>(Ld + i * 32 + k);
      const double2 m = *reinterpret_cast<const double2*>(Ld + i * 32 + k + 2);
      s0 = fma(l.x, x[k], s0);
      s1 = fma(l.y, x[k + 1], s1);
      s2 = fma(m.x, x[k + 2], s2);
      s3 = fma(m.y, x[k + 3], s3);
    }
#pragma unroll
    for (int k = i & ~3; k < i; ++k) s0 = fma(Ld[i * 32 + k], x[k], s0);
    s0 = (s0 + s1) + (s2 + s3);
    s1 = 0.0;
    const double rdi = rdiag_c0[i];
    x[i] = (i == lane) ? rdi : ((i > lane) ? -(s0 + s1) * rdi : 0.0);
    Xd[lane * 32 + i] = x[i];
  }
}

// acc[c] += sum_{kk < 32} left[kk * sl] * right[c * sr + kk],  c < 8.   `left` already points at this lane's row
// (rows are contiguous across lanes); `right` is warp-uniform: its reads are 16-byte broadcasts, two kk per load.
__device__ __forceinline__ void potf2_block_product(const double* __restrict__ left, int sl,
                                                    const double* __restrict__ right, int sr, double (&acc)[8]) {
#pragma unroll 4
  for (int kk = 0; kk < 32; kk += 2) {
    const double l0 = left[kk * sl], l1 = left[(kk + 1) * sl];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const double2 r = *reinterpret_cast<const double2*>(right + c * sr + kk);
      acc[c] = fma(l0, r.x, acc[c]);
      acc[c] = fma(l1, r.y, acc[c]);
    }
  }
}

__global__ void __launch_bounds__(POTF2_THREADS, 1)
potf2_inv_kernel(double* __restrict__ A, double* __restrict__ W, long long ld, int jb, double* __restrict__ logdet_blocks,
                 int blk_index, int* __restrict__ info, int check_abort) {
  extern __shared__ double sm[];
  double* As = sm;                     // 128*128
  double* Xb = sm + 128 * 128;         // 10 blocks of 32x32: Xb[blk][c*32 + r] = X[32i + r, 32j + c]
  double* Ldt = Xb + 10 * 1024;        // 2 x (32x32) row-major diagonal sub-blocks (double buffered)
  double* rdiag = Ldt + 2 * 1024;      // 1 / L_ii
  double* diag = rdiag + 128;          // L_ii
  double* logp = diag + 128;           // 4 partial log sums
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  double* Ablk = A + (long long)jb * ld + jb;
  // an earlier panel of this attempt already failed: the attempt is discarded, skip the rest of it
  if (check_abort) {
    __shared__ int s_abort;
    if (tid == 0) s_abort = *reinterpret_cast<const volatile int*>(info);
    __syncthreads();
    if (s_abort != 0) return;
  }

  POTF2_STAMP(0);
  {
    // 128 x 128 doubles = 16 double2 per thread: all 16 loads are in flight before the first store
    double2 v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + q * POTF2_THREADS;
      v[q] = *reinterpret_cast<const double2*>(Ablk + (long long)(idx >> 6) * ld + (idx & 63) * 2);
    }
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int idx = tid + q * POTF2_THREADS;
      *reinterpret_cast<double2*>(As + (idx >> 6) * 128 + (idx & 63) * 2) = v[q];
    }
  }
  __syncthreads();
  POTF2_STAMP(1);

  bool ok = true;
  for (int kb = 0; kb < 4; ++kb) {
    const int c0 = 32 * kb;
    const int nrow = 128 - c0;
    double* Ld = Ldt + (kb & 1) * 1024;
    if (tid < nrow) {
      // ---------------- AB: 32 column steps, one thread per row ----------------
      const int row = c0 + tid;
      double a[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) a[j] = As[(c0 + j) * 128 + row];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        potf2_bar(nrow);  // row j of Ld is complete (entries k < j)
        double v0 = a[j], v1 = 0.0;
        double p0 = As[(c0 + j) * 128 + c0 + j], p1 = 0.0;
#pragma unroll
        for (int k = 0; k + 1 < j; k += 2) {
          const double2 l = *reinterpret_cast<const double2*>(Ld + j * 32 + k);
          v0 = fma(-a[k], l.x, v0);
          v1 = fma(-a[k + 1], l.y, v1);
          p0 = fma(-l.x, l.x, p0);
          p1 = fma(-l.y, l.y, p1);
        }
        if (j & 1) {
          const double l = Ld[j * 32 + j - 1];
          v0 = fma(-a[j - 1], l, v0);
          p0 = fma(-l, l, p0);
        }
        const double piv = p0 + p1;
        ok = ok && (piv > 0.0);
        double rd = rsqrt(piv);
        double dj = piv * rd;
        dj = fma(fma(-dj, dj, piv) * 0.5, rd, dj);  // Newton step: dj = sqrt(piv) to working accuracy
        rd = fma(fma(-dj, rd, 1.0), rd, rd);        // rd = 1 / dj
        a[j] = (tid == j) ? dj : (v0 + v1) * rd;
        if (tid > j && tid < 32) Ld[tid * 32 + j] = a[j];
        if (tid == j) {
          diag[c0 + j] = dj;
          rdiag[c0 + j] = rd;
        }
      }
      potf2_bar(nrow);  // everybody is done reading the original diagonal entries from As
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (tid >= j) As[(c0 + j) * 128 + row] = a[j];
    } else if (warp == 4 && kb > 0) {
      // inverse of the previous diagonal sub-block, off the critical chain
      potf2_invert_diag(Ldt + ((kb - 1) & 1) * 1024, rdiag + c0 - 32, Xb + potf2_blk(kb - 1, kb - 1) * 1024, lane);
    }
    __syncthreads();
    POTF2_STAMP(2 + 2 * kb);
    const int R0 = c0 + 32;
    const int m = 128 - R0;  // rows / columns left after this sub-panel
    if (m > 0) {
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
    POTF2_STAMP(3 + 2 * kb);
  }

  // failure flag (NaN-safe: piv > 0 is false for NaN; every thread of rows 0..127 saw every pivot) and the block's
  // log-determinant part, summed in a fixed order
  if (tid < 128) {
    const double ls = warp_sum(log(diag[tid]));
    if (lane == 0) logp[warp] = ls;
  }
  if (tid == 0 && !ok) atomicExch(info, 1);  // thread 0 took part in all four sub-panels
  __syncthreads();
  if (tid == 0) logdet_blocks[blk_index] = ((logp[0] + logp[1]) + logp[2]) + logp[3];

  // write L back (lower part incl. diagonal; strict upper zero) -- warps 0..14; meanwhile warp 15 inverts the last
  // diagonal sub-block (the first three were inverted behind the column steps)
  if (warp == 15) {
    potf2_invert_diag(Ldt + 1024, rdiag + 96, Xb + potf2_blk(3, 3) * 1024, lane);
  } else {
    for (int idx = tid; idx < 128 * 64; idx += POTF2_THREADS - 32) {
      const int c = idx >> 6, r2 = (idx & 63) * 2;
      // pairs that lie wholly above the diagonal are not read: the upper blocks of As become the inverse stage's T
      // scratch as soon as the first warp gets there (racecheck: read here against the write of T below)
      double2 v = make_double2(0.0, 0.0);
      if (r2 + 1 >= c) v = *reinterpret_cast<const double2*>(As + c * 128 + r2);
      if (r2 < c) v.x = 0.0;
      *reinterpret_cast<double2*>(Ablk + (long long)c * ld + r2) = v;
    }
  }

  POTF2_STAMP(10);
  // ---------------- off-diagonal blocks of the inverse, one block sub-diagonal at a time ----------------
  // T_ij = sum_{k=j}^{i-1} L_ik X_kj  is parked in As's upper block (j, i);  X_ij = -X_ii T_ij.
  // Both are 32 x 32 x 32 block products run by potf2_block_product: one warp per (block, 8-column group),
  // lane = row, 8 accumulators -- the left operand is read once per k (conflict-free, rows contiguous), the right one
  // as warp-uniform 16-byte broadcasts.  (Round 1 computed one element per thread with two LDS per FMA: 8.4 MB of
  // shared-memory traffic per panel, 27 of the kernel's 71 us by the phase stamps.)
  for (int dd = 1; dd < 4; ++dd) {
    const int nblk = 4 - dd;
    for (int item = warp; item < nblk * 4; item += POTF2_THREADS / 32) {
      const int j = item >> 2, i = j + dd, c0 = (item & 3) * 8;
      double acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.0;
      for (int k = j; k < i; ++k)  // L[32i + lane, 32k + kk] = Lik[kk * 128] ; X[32k + kk, 32j + c] = Xkj[c * 32 + kk]
        potf2_block_product(As + (32 * k) * 128 + 32 * i + lane, 128, Xb + potf2_blk(k, j) * 1024 + c0 * 32, 32, acc);
#pragma unroll
      for (int c = 0; c < 8; ++c) As[(32 * i + c0 + c) * 128 + 32 * j + lane] = acc[c];  // T[lane, c0 + c]
    }
    __syncthreads();
    for (int item = warp; item < nblk * 4; item += POTF2_THREADS / 32) {
      const int j = item >> 2, i = j + dd, c0 = (item & 3) * 8;
      double acc[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[c] = 0.0;
      // X[32i + lane, 32i + kk] = Xii[kk * 32] ; T[kk, c] = T[c * 128 + kk]
      potf2_block_product(Xb + potf2_blk(i, i) * 1024 + lane, 32, As + (32 * i + c0) * 128 + 32 * j, 128, acc);
#pragma unroll
      for (int c = 0; c < 8; ++c) Xb[potf2_blk(i, j) * 1024 + (c0 + c) * 32 + lane] = -acc[c];
    }
    __syncthreads();
  }

  POTF2_STAMP(11);
  // write the inverse into W's diagonal block (strict upper zero)
  double* Wblk = W + (long long)jb * ld + jb;
  for (int idx = tid; idx < 128 * 64; idx += POTF2_THREADS) {
    const int c = idx >> 6, r2 = (idx & 63) * 2;
    const int i = r2 >> 5, j = c >> 5;
    double2 v = make_double2(0.0, 0.0);
    if (i >= j) v = *reinterpret_cast<const double2*>(Xb + potf2_blk(i, j) * 1024 + (c & 31) * 32 + (r2 & 31));
    *reinterpret_cast<double2*>(Wblk + (long long)c * ld + r2) = v;
  }
  POTF2_STAMP(12);
}

}  // namespace lk
