// variogram.cuh -- sigma2 bounds of NoiseModel::Heterogeneous on the device.
//
// Reference (src/lib/Kriging.cpp:1784-1797):
//   dX2 = sum(m_dX % m_dX, 0)            (n^2 ordered pairs, diagonal included)
//   sigma2_variogram = 0.5 * mean( dy2[ dX2 >= median(dX2) ] )
// The reference materialises dX (8 d n^2 B), dX % dX, dX2 and dy2 and runs std::nth_element over n^2 doubles.  Here
// nothing of size n^2 exists: the median is found by an MSD radix select over the IEEE bit patterns of the pair
// distances (non-negative doubles order like their 64-bit patterns), 11 bits per pass, each pass regenerating the
// distances from 64 x d tile pairs of X in shared memory and histogramming the digit of the keys that match the
// prefix selected so far; a last pass sums dy^2 over the pairs at or above the median.  Every off-diagonal distance
// appears twice in the reference's multiset (d_ij == d_ji bit for bit) and the n diagonal zeros once: the kernels
// walk the pairs i > j with weight 2 and the host adds the diagonal.
// The distance itself is accumulated the way Armadillo's arrayops::accumulate sums a column of dX % dX (two
// interleaved accumulators, products rounded before the adds, no FMA), so the selected median is the reference's.
#pragma once
#include "cov.cuh"

namespace lk {

constexpr int VG_DIGIT_BITS = 11;
constexpr int VG_BINS = 1 << VG_DIGIT_BITS;

// squared distance between row a of tile xi and row b of tile xj (tiles are [d][PT]), Armadillo summation order
__device__ __forceinline__ double vg_dist2(const double* xi, const double* xj, int d, int a, int b) {
  double acc1 = 0.0, acc2 = 0.0;
  int k = 0;
  for (; k + 1 < d; k += 2) {
    const double d0 = xi[k * PT + a] - xj[k * PT + b];
    const double d1 = xi[(k + 1) * PT + a] - xj[(k + 1) * PT + b];
    acc1 = __dadd_rn(acc1, __dmul_rn(d0, d0));
    acc2 = __dadd_rn(acc2, __dmul_rn(d1, d1));
  }
  if (k < d) {
    const double d0 = xi[k * PT + a] - xj[k * PT + b];
    acc1 = __dadd_rn(acc1, __dmul_rn(d0, d0));
  }
  return __dadd_rn(acc1, acc2);
}

__device__ __forceinline__ void vg_stage(const double* __restrict__ X, int n, int d, int ti, int tj, double* xi,
                                         double* xj) {
  for (int e = threadIdx.x; e < d * PT; e += PAIR_THREADS) {
    const int k = e / PT, r = e % PT;
    xi[e] = (ti * PT + r < n) ? X[(long long)k * n + ti * PT + r] : 0.0;
    xj[e] = (tj * PT + r < n) ? X[(long long)k * n + tj * PT + r] : 0.0;
  }
}

// One radix-select pass: hist[digit] += number of pairs i > j whose key matches `prefix` above bit `hi_shift`
// (hi_shift == 64: every key matches), digit = (key >> shift) & mask.  Counts are integers: the result does not
// depend on the order of the atomics.
__global__ void __launch_bounds__(PAIR_THREADS)
vario_hist_kernel(const double* __restrict__ X, int n, int d, unsigned long long prefix, int hi_shift, int shift,
                  unsigned int mask, unsigned long long* __restrict__ hist, int ntiles) {
  extern __shared__ double sm[];
  double* xi = sm;
  double* xj = xi + d * PT;
  __shared__ unsigned int shist[VG_BINS];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, lane = threadIdx.x & 31;
  for (int e = threadIdx.x; e < VG_BINS; e += PAIR_THREADS) shist[e] = 0u;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int ti, tj;
    tri_tile(tile, ti, tj);
    __syncthreads();
    vg_stage(X, n, d, ti, tj, xi, xj);
    __syncthreads();
#pragma unroll 1
    for (int a = 0; a < 4; ++a)
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        const int i = ti * PT + tx * 4 + a, j = tj * PT + ty * 4 + b;
        int bin = -1;
        if (i > j && i < n) {
          const unsigned long long key = (unsigned long long)__double_as_longlong(vg_dist2(xi, xj, d, tx * 4 + a, ty * 4 + b));
          if (hi_shift >= 64 || (key >> hi_shift) == prefix) bin = (int)((key >> shift) & mask);
        }
        // warp-aggregated shared-memory atomics: the leading digits of the distances fall into a handful of bins
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin >= 0 && lane == __ffs(peers) - 1) atomicAdd(&shist[bin], (unsigned)__popc(peers));
      }
  }
  __syncthreads();
  for (int e = threadIdx.x; e < VG_BINS; e += PAIR_THREADS)
    if (shist[e]) atomicAdd(&hist[e], (unsigned long long)shist[e]);
}

// partial[2 * cta] = sum of dy^2, partial[2 * cta + 1] = number of pairs, over the pairs i > j with dist2 >= med.
// Fixed tile -> CTA assignment and fixed reduction order: deterministic.
__global__ void __launch_bounds__(PAIR_THREADS)
vario_sum_kernel(const double* __restrict__ X, const double* __restrict__ y, int n, int d, double med,
                 double* __restrict__ partial, int ntiles) {
  extern __shared__ double sm[];
  double* xi = sm;
  double* xj = xi + d * PT;
  double* yi = xj + d * PT;
  double* yj = yi + PT;
  __shared__ double wsum[8], wcnt[8];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double s = 0.0, c = 0.0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int ti, tj;
    tri_tile(tile, ti, tj);
    __syncthreads();
    vg_stage(X, n, d, ti, tj, xi, xj);
    if (threadIdx.x < PT) yi[threadIdx.x] = (ti * PT + threadIdx.x < n) ? y[ti * PT + threadIdx.x] : 0.0;
    else if (threadIdx.x < 2 * PT)
      yj[threadIdx.x - PT] = (tj * PT + threadIdx.x - PT < n) ? y[tj * PT + threadIdx.x - PT] : 0.0;
    __syncthreads();
    double ts = 0.0, tc = 0.0;
#pragma unroll 1
    for (int a = 0; a < 4; ++a)
#pragma unroll 1
      for (int b = 0; b < 4; ++b) {
        const int i = ti * PT + tx * 4 + a, j = tj * PT + ty * 4 + b;
        if (i > j && i < n && vg_dist2(xi, xj, d, tx * 4 + a, ty * 4 + b) >= med) {
          const double dy = yi[tx * 4 + a] - yj[ty * 4 + b];
          ts += dy * dy;
          tc += 1.0;
        }
      }
    s += warp_sum(ts);
    c += warp_sum(tc);
  }
  if (lane == 0) {
    wsum[warp] = s;
    wcnt[warp] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double S = 0.0, Cn = 0.0;
    for (int w = 0; w < 8; ++w) {
      S += wsum[w];
      Cn += wcnt[w];
    }
    partial[2 * blockIdx.x] = S;
    partial[2 * blockIdx.x + 1] = Cn;
  }
}

}  // namespace lk
