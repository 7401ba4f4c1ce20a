// trsv_wave.cuh -- wavefront triangular sweeps (reference solve_lower / solve_upper,
// src/lib/LinearAlgebra.cpp:695-701, called at src/lib/KrigingImpl.cpp:102-103, 122 and src/lib/Kriging.cpp:294).
//
// One persistent kernel per sweep replaces the nb-step launch chain.  The sweep is HBM-bound (4 n^2 bytes of L
// per sweep, a few right-hand sides), and its only serial part is the chain z_0 -> z_1 -> ... of 128-row blocks:
//   forward  (L z = b):    z_i = Dinv_i (b_i - sum_{j<i} L[i,j] z_j)
//   backward (L^T x = e):  x_k = Dinv_k^T (e_k - sum_{j>k} L[j,k]^T x_j)
// Dinv_i = inverse of L's diagonal block, kept in W's diagonal blocks by the panel kernel (and still there after
// TRTRI: the diagonal blocks of L^-1 are the inverses of the diagonal blocks of L).
//
// CTAs take row blocks from a ticket counter in sweep order, so a CTA only ever waits on blocks owned by CTAs
// that are already resident (no deadlock for any grid size).  Each CTA streams ITS row (column) of 128 x 128
// tiles through a TMA ring as fast as HBM delivers them -- L is read-only, so the prefetch never waits -- and
// consumes tile j as soon as z_j is published.  Publication is by VALUE: the solution blocks go to a scratch vector
// Zpub that the host pre-fills with a sentinel bit pattern (a NaN payload no computation produces; computed NaNs
// are canonicalised before the store), and each consumer thread polls its own element with ld.relaxed.gpu until it
// is not the sentinel -- one L2 round trip from the producer's store to the consumer's register, no flag, no fence
// (every 8-byte element is published and consumed on its own).  The serial chain per block is then:
// see z_j -> one tile product -> one Dinv product -> store z_i.  (Round 1 published a flag per block with
// st.release after a __threadfence and re-read the block after the acquire: three dependent L2 round trips and a
// gpu-scope fence per step, ~4 us x nb steps per sweep -- the sweep was bound by that chain, not by HBM:
// profiles/r02_side_kernels.md.)
//
// Tile traffic: forward uses boxes {128 rows, 32 cols} (dense, thread <-> row: conflict-free LDS.64),
// backward uses boxes {16 rows, 128 cols} with SWIZZLE_128B (thread <-> column: conflict-free LDS.128).
#pragma once
#include "common.cuh"

#ifndef LKGPU_RING_RELEASE
#define LKGPU_RING_RELEASE 1
#endif

namespace lk {

constexpr int WAVE_COMPUTE_THREADS = 256;
constexpr int WAVE_THREADS = WAVE_COMPUTE_THREADS + 32;  // + 1 TMA producer warp
constexpr int WAVE_STAGES = 6;
constexpr int WAVE_STAGE_BYTES = 32768;  // fwd: one {128 x 32} box ; bwd: two {16 x 128} boxes
constexpr int WAVE_SUBTILES = 4;         // sub-tiles (stages) per 128 x 128 tile
constexpr int WAVE_MAX_RHS = 8;

__host__ __device__ constexpr int wave_smem_bytes(int nq) {
  return WAVE_STAGES * WAVE_STAGE_BYTES + 1024 /*align*/ + 3 * nq * 128 * 8 /*vec + 2 partials*/ + 256 /*barriers*/;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// strong (gpu-scope) accesses for the right-hand-side blocks that travel between CTAs inside one sweep
__device__ __forceinline__ double ld_relaxed_gpu_f64(const double* p) {
  double v;
  asm volatile("ld.relaxed.gpu.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_gpu_f64(double* p, double v) {
  asm volatile("st.relaxed.gpu.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ void bar_sync_compute() {
  asm volatile("bar.sync 1, %0;" ::"n"(WAVE_COMPUTE_THREADS) : "memory");
}

constexpr unsigned long long WAVE_SENTINEL = 0xFFF85EA71E5EA711ull;  // "not yet published"
__device__ __forceinline__ unsigned long long ld_relaxed_gpu_u64(const double* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// before each launch: ticket counter = 0, Zpub[0 .. count) = sentinel
__global__ void wave_reset_kernel(int* __restrict__ ctl, double* __restrict__ zpub, long long count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) ctl[0] = 0;
  if (i < count) zpub[i] = __longlong_as_double((long long)WAVE_SENTINEL);
}

// ctl[0] = ticket counter; Zpub = [nrhs][ldb] published solution blocks (see above).
template <bool BWD, int NQ>
__global__ void __launch_bounds__(WAVE_THREADS, 1)
trsv_wave_kernel(const CUtensorMap* tmapL, const CUtensorMap* tmapW,  // tensor maps in device memory
                 double* __restrict__ B, long long ldb, int nrhs, int nb, int* __restrict__ ctl,
                 double* __restrict__ Zpub, int sched_fence) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by pointer arithmetic on the shared array (not through an integer cast): the compiler keeps
  // the shared address space, so the tile / vector reads below are LDS, not generic loads.
  uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  double* vec = reinterpret_cast<double*>(ring + WAVE_STAGES * WAVE_STAGE_BYTES);  // [NQ][128] current z_j / t
  double* part = vec + NQ * 128;                                                   // [2][NQ][128] partials
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(part + 2 * NQ * 128);
  uint64_t* empty_bar = full_bar + WAVE_STAGES;
  __shared__ int s_ticket;

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const bool producer = warp == WAVE_COMPUTE_THREADS / 32;
  if (tid == 0) {
    for (int s = 0; s < WAVE_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], WAVE_COMPUTE_THREADS / 32);
    }
    fence_barrier_init();
  }
  __syncthreads();
  if (producer && lane == 0) {
    tma_prefetch_desc(tmapL);
    tma_prefetch_desc(tmapW);
  }
  int stage = 0;
  uint32_t phase = 0;

  while (true) {
    if (tid == 0) s_ticket = atomicAdd(ctl, 1);
    __syncthreads();
    const int ticket = s_ticket;
    __syncthreads();
    if (ticket >= nb) break;
    const int i = BWD ? (nb - 1 - ticket) : ticket;  // this CTA's row block
    const int ntiles = ticket + 1;                   // off-diagonal tiles in sweep order, then the Dinv tile
    const int ib = i * 128;

    if (producer) {
      if (lane == 0) {
        for (int tq = 0; tq < ntiles; ++tq) {
          const bool diag = (tq == ntiles - 1);
          const int j = BWD ? (nb - 1 - tq) : tq;  // for the last tile j == i
          const CUtensorMap* map = diag ? tmapW : tmapL;
          for (int s = 0; s < WAVE_SUBTILES; ++s) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            uint8_t* dst = ring + stage * WAVE_STAGE_BYTES;
            mbar_arrive_expect_tx(&full_bar[stage], WAVE_STAGE_BYTES);
            if (!BWD) {
              // tile rows = block i, cols = block j : box {128 rows, 32 cols} at (row ib, col 128 j + 32 s)
              tma_load_2d(dst, map, &full_bar[stage], ib, j * 128 + 32 * s);
            } else {
              // tile rows = block j, cols = block i : two boxes {16 rows, 128 cols} at rows 128 j + 32 s (+16)
              tma_load_2d(dst, map, &full_bar[stage], j * 128 + 32 * s, ib);
              tma_load_2d(dst + 16384, map, &full_bar[stage], j * 128 + 32 * s + 16, ib);
            }
            if (++stage == WAVE_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      __syncwarp();
      continue;  // producer warp goes back for the next ticket
    }

    // ============================ compute threads ============================
    const int r = tid & 127, h = tid >> 7;
    double acc[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] = (h == 0 && q < nrhs) ? ld_relaxed_gpu_f64(B + q * ldb + ib + r) : 0.0;

    for (int tq = 0; tq < ntiles; ++tq) {
      const bool diag = (tq == ntiles - 1);
      if (!diag) {
        const int j = BWD ? (nb - 1 - tq) : tq;
        double zj[NQ];
        if (tid < 128) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) {
            zj[q] = 0.0;
            if (q < nrhs) {
              unsigned long long bits;
              do {
                bits = ld_relaxed_gpu_u64(Zpub + q * ldb + j * 128 + tid);
              } while (bits == WAVE_SENTINEL);
              zj[q] = __longlong_as_double((long long)bits);
            }
          }
        }
        bar_sync_compute();  // everybody is done with the previous tile's vec
        if (tid < 128) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) vec[q * 128 + tid] = zj[q];
        }
        bar_sync_compute();
      } else {
        // t = b_i - sum(...) : combine the two half-sums, negate-free (acc already holds b - sum), then restart
        bar_sync_compute();
#pragma unroll
        for (int q = 0; q < NQ; ++q) part[(h * NQ + q) * 128 + r] = acc[q];
        bar_sync_compute();
        if (h == 0) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) vec[q * 128 + r] = part[q * 128 + r] + part[(NQ + q) * 128 + r];
        }
#pragma unroll
        for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
        bar_sync_compute();
      }
      const double sgn = diag ? 1.0 : -1.0;
      for (int s = 0; s < WAVE_SUBTILES; ++s) {
        mbar_wait(&full_bar[stage], phase);
        const uint8_t* tile = ring + stage * WAVE_STAGE_BYTES;
        if (!BWD) {
          // thread (r, h): row r, columns 16 h .. 16 h + 15 of this 32-column sub-tile
          const double* tp = reinterpret_cast<const double*>(tile) + (h * 16) * 128 + r;
          const double* vp = vec + 32 * s + 16 * h;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const double a = sgn * tp[c * 128];
#pragma unroll
            for (int q = 0; q < NQ; ++q) acc[q] = fma(a, vp[q * 128 + c], acc[q]);
          }
        } else {
          // thread (c = r, h): column c, rows 16 h .. 16 h + 15 of this 32-row sub-tile (box h, 8 chunks of 16 B)
          const uint8_t* line = tile + h * 16384 + r * 128;
          const double* vp = vec + 32 * s + 16 * h;
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const double2 a = *reinterpret_cast<const double2*>(line + ((ch ^ (r & 7)) << 4));
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
              acc[q] = fma(sgn * a.x, vp[q * 128 + 2 * ch], acc[q]);
              acc[q] = fma(sgn * a.y, vp[q * 128 + 2 * ch + 1], acc[q]);
            }
          }
        }
        // release of the ring slot: per-lane proxy fence, warp barrier, then lane 0's arrive (see gemm_dmma.cuh)
#if LKGPU_RING_RELEASE
        fence_proxy_async_smem();
        __syncwarp();
#else
        __syncwarp();
        if (sched_fence) fence_proxy_async();
#endif
        if (lane == 0) mbar_arrive(&empty_bar[stage]);
        if (++stage == WAVE_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    // z_i = sum of the two half-products of the Dinv tile
    bar_sync_compute();
#pragma unroll
    for (int q = 0; q < NQ; ++q) part[(h * NQ + q) * 128 + r] = acc[q];
    bar_sync_compute();
    if (h == 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (q < nrhs) {
          double z = part[q * 128 + r] + part[(NQ + q) * 128 + r];
          if (z != z) z = __longlong_as_double(0x7FF8000000000000ll);  // never the sentinel
          st_relaxed_gpu_f64(Zpub + q * ldb + ib + r, z);  // published: consumers poll this element
          B[q * ldb + ib + r] = z;                          // the result
        }
    }
    bar_sync_compute();  // part[] is reused by the next row block
  }
}

}  // namespace lk
