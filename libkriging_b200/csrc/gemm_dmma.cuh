// gemm_dmma.cuh -- the FP64 tensor-core tile engine behind Cholesky's trailing
// update (SYRK), the panel TRSM, TRTRI and LAUUM.
//
// One persistent, warp-specialised kernel:
//   * 1 producer warp: one elected lane issues TMA (cp.async.bulk.tensor.2d,
//     SWIZZLE_128B) loads of the two operand tiles of each k-stage into a
//     4-deep shared-memory ring guarded by full/empty mbarriers, running ahead
//     across tile boundaries;
//   * 4 consumer warps: each owns a 64(row) x 32(col) slab of the 64 x 128
//     output tile as 32 independent m8n8k4 FP64 accumulators (DMMA.8x8x4),
//     reads its fragments conflict-free from the swizzled tiles, and applies
//     the epilogue (C = acc | C = -acc | C -= acc) with 16-byte accesses.
// Two CTAs fit per SM (96 KB smem, <= 168 regs) so one CTA's epilogue overlaps
// the other's main loop.
//
// All matrices are column-major doubles with dimensions padded to 128.
//   C[c_row + m, c_col + n] (op)= sum_{k in [k_begin, k_end)} Ms(m, k) * Ns(n, k)
// "M-side" operand Ms supplies C's rows, "N-side" operand Ns supplies C's columns:
//   M-major (KMAJ = false): element (row = c_row + m, col = k)   of the operand buffer
//   K-major (KMAJ = true ): element (row = k,         col = c_row + m)
// NT (false,false): SYRK / TRSM-by-inverse;  NN (false,true): TRTRI;  TN (true,true): LAUUM.
//
// MMA mapping (transposed so that each lane's accumulator pair is two
// consecutive ROWS of one column -> one 16-byte global access):
//   mma A-fragment <- N-side tile, mma B-fragment <- M-side tile,
//   lane (g = lane/4, t = lane%4): acc[i][j] = C[c_row + 8j + 2t (+1)][c_col + 8(4w+i) + g].
#pragma once
#include "common.cuh"

namespace lk {

constexpr int TM = 64;       // C rows per tile
constexpr int TN = 128;      // C cols per tile
constexpr int TK = 16;       // k per pipeline stage (16 doubles = one 128-byte swizzle row)
constexpr int GSTAGES = 4;   // pipeline depth
constexpr int GEMM_CONSUMER_WARPS = 4;
constexpr int GEMM_THREADS = 32 * (GEMM_CONSUMER_WARPS + 1);
constexpr int NS_TILE_BYTES = TN * TK * 8;  // 16384
constexpr int MS_TILE_BYTES = TM * TK * 8;  // 8192
constexpr int STAGE_BYTES = NS_TILE_BYTES + MS_TILE_BYTES;
constexpr int GEMM_SMEM_BYTES = GSTAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum GemmEpilogue { EPI_SET = 0, EPI_SETNEG = 1, EPI_SUB = 2 };
enum GemmSched { SCHED_TABLE = 0, SCHED_RECT = 1, SCHED_TRAP = 2 };

struct TileDesc {
  int c_row, c_col, k_begin, k_end;
};

struct GemmArgs {
  double* C;          // output buffer (column-major)
  long long ldc;
  int sched;          // GemmSched
  int epilogue;       // GemmEpilogue
  int ntiles;
  // SCHED_TABLE
  const TileDesc* table;
  // SCHED_RECT: tiles (tm, tn), tm fastest; SCHED_TRAP: lower trapezoid tm >= 2*tn of a square region
  int row0, col0, mt, nt;
  int k_begin, k_end;
};

__device__ __forceinline__ TileDesc gemm_get_tile(const GemmArgs& a, int id) {
  TileDesc t;
  if (a.sched == SCHED_TABLE) {
    t = a.table[id];
  } else if (a.sched == SCHED_RECT) {
    int tm = id % a.mt, tn = id / a.mt;
    t.c_row = a.row0 + tm * TM;
    t.c_col = a.col0 + tn * TN;
    t.k_begin = a.k_begin;
    t.k_end = a.k_end;
  } else {
    // column tn holds row tiles 2*tn .. mt-1 ; offset(tn) = tn*mt - tn*(tn-1)
    int mt = a.mt;
    double disc = (double)(mt + 1) * (double)(mt + 1) - 4.0 * (double)id;
    int tn = (int)(((double)(mt + 1) - sqrt(disc > 0.0 ? disc : 0.0)) * 0.5);
    if (tn < 0) tn = 0;
    while (tn > 0 && tn * mt - tn * (tn - 1) > id) --tn;
    while ((tn + 1) * mt - (tn + 1) * tn <= id) ++tn;
    int tm = 2 * tn + (id - (tn * mt - tn * (tn - 1)));
    t.c_row = a.row0 + tm * TM;
    t.c_col = a.col0 + tn * TN;
    t.k_begin = a.k_begin;
    t.k_end = a.k_end;
  }
  return t;
}

// Byte offset of element (r, kk) inside a swizzled operand tile.
//  M-major tile: [r/16][kk][r%16] 128-byte rows (one TMA box {16 rows, 16 k} per r/16), chunk ^= kk&7
//  K-major tile: [r][kk]          128-byte rows (TMA boxes {16 k, 64 r}),              chunk ^= r&7
template <bool KMAJ>
__device__ __forceinline__ uint32_t tile_off(int r, int kk) {
  if (KMAJ) {
    return (uint32_t)(r * 128 + ((((kk >> 1) ^ (r & 7)) << 4) | ((kk & 1) << 3)));
  } else {
    int mi = r & 15;
    return (uint32_t)((((r >> 4) * 16 + kk) * 128) + ((((mi >> 1) ^ (kk & 7)) << 4) | ((mi & 1) << 3)));
  }
}

template <bool MS_KMAJ, bool NS_KMAJ>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_dmma_kernel(const __grid_constant__ CUtensorMap tmapM, const __grid_constant__ CUtensorMap tmapN,
                 const GemmArgs args) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms are 1024 bytes: align the ring.
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + GSTAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GSTAGES;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], GEMM_CONSUMER_WARPS);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == GEMM_CONSUMER_WARPS) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      tma_prefetch_desc(&tmapM);
      tma_prefetch_desc(&tmapN);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x) {
        TileDesc td = gemm_get_tile(args, tile);
        for (int k0 = td.k_begin; k0 < td.k_end; k0 += TK) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sN = ring + stage * STAGE_BYTES;
          uint8_t* sM = sN + NS_TILE_BYTES;
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (NS_KMAJ) {
#pragma unroll
            for (int b = 0; b < TN / 64; ++b)
              tma_load_2d(sN + b * 8192, &tmapN, &full_bar[stage], k0, td.c_col + 64 * b);
          } else {
#pragma unroll
            for (int b = 0; b < TN / 16; ++b)
              tma_load_2d(sN + b * 2048, &tmapN, &full_bar[stage], td.c_col + 16 * b, k0);
          }
          if (MS_KMAJ) {
            tma_load_2d(sM, &tmapM, &full_bar[stage], k0, td.c_row);
          } else {
#pragma unroll
            for (int b = 0; b < TM / 16; ++b)
              tma_load_2d(sM + b * 2048, &tmapM, &full_bar[stage], td.c_row + 16 * b, k0);
          }
          if (++stage == GSTAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    return;
  }

  // ===================== DMMA consumers =====================
  const int g = lane >> 2, t = lane & 3;
  const uint32_t ring_u32 = smem_u32(ring);
  int stage = 0;
  uint32_t phase = 0;

  // Per-lane fragment offsets inside a stage, for the 4 k-steps of a stage.
  // k index used by lane t in k-step s (a permutation of 0..15 chosen so that the
  // 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs):
  //   any M-major operand present: kk = 2t + (s&1) + 8(s>>1)
  //   both K-major:                kk = 2s + (t&1) + 8(t>>1)
  uint32_t offN[4][4];  // [kstep][i]
  uint32_t offM[4][8];  // [kstep][j]
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    const int kk = (MS_KMAJ && NS_KMAJ) ? (2 * s + (t & 1) + 8 * (t >> 1)) : (2 * t + (s & 1) + 8 * (s >> 1));
#pragma unroll
    for (int i = 0; i < 4; ++i) offN[s][i] = tile_off<NS_KMAJ>((warp * 4 + i) * 8 + g, kk);
#pragma unroll
    for (int j = 0; j < 8; ++j) offM[s][j] = NS_TILE_BYTES + tile_off<MS_KMAJ>(j * 8 + g, kk);
  }

  for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x) {
    TileDesc td = gemm_get_tile(args, tile);
    double acc[4][8][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int k0 = td.k_begin; k0 < td.k_end; k0 += TK) {
      mbar_wait(&full_bar[stage], phase);
      const uint32_t base = ring_u32 + stage * STAGE_BYTES;
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        double a[4], b[8];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = lds_f64(base + offN[s][i]);
#pragma unroll
        for (int j = 0; j < 8; ++j) b[j] = lds_f64(base + offM[s][j]);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == GSTAGES) {
        stage = 0;
        phase ^= 1;
      }
    }

    // ---- epilogue ----
    double* Cbase = args.C + (long long)(td.c_col + warp * 32 + g) * args.ldc + td.c_row + 2 * t;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double* Ccol = Cbase + (long long)(8 * i) * args.ldc;
      if (args.epilogue == EPI_SUB) {
        double2 old[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) old[j] = *reinterpret_cast<const double2*>(Ccol + 8 * j);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double2 v;
          v.x = old[j].x - acc[i][j][0];
          v.y = old[j].y - acc[i][j][1];
          *reinterpret_cast<double2*>(Ccol + 8 * j) = v;
        }
      } else {
        const double sgn = (args.epilogue == EPI_SETNEG) ? -1.0 : 1.0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          double2 v;
          v.x = sgn * acc[i][j][0];
          v.y = sgn * acc[i][j][1];
          *reinterpret_cast<double2*>(Ccol + 8 * j) = v;
        }
      }
    }
  }
}

// Plain CUDA-core restatement of the same tile contract, used ONLY by the GPU
// test-suite (LKGPU_DEBUG_SIMPLE_GEMM=1) to localise faults in the DMMA/TMA path.
template <bool MS_KMAJ, bool NS_KMAJ>
__global__ void gemm_simple_kernel(const double* __restrict__ Mbuf, long long ldm, const double* __restrict__ Nbuf,
                                   long long ldn, const GemmArgs args) {
  // blockDim.x must be 256: each thread owns 32 elements; sums are formed before any store so that the
  // in-place panel TRSM (C aliases the M-side operand) is safe, exactly as in the DMMA kernel.
  for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x) {
    TileDesc td = gemm_get_tile(args, tile);
    double sums[32];
    for (int q = 0; q < 32; ++q) {
      const int e = threadIdx.x + 256 * q;
      int m = e % TM, n = e / TM;
      double s = 0.0;
      for (int k = td.k_begin; k < td.k_end; ++k) {
        double a = MS_KMAJ ? Mbuf[(long long)(td.c_row + m) * ldm + k] : Mbuf[(long long)k * ldm + td.c_row + m];
        double b = NS_KMAJ ? Nbuf[(long long)(td.c_col + n) * ldn + k] : Nbuf[(long long)k * ldn + td.c_col + n];
        s += a * b;
      }
      sums[q] = s;
    }
    __syncthreads();
    for (int q = 0; q < 32; ++q) {
      const int e = threadIdx.x + 256 * q;
      int m = e % TM, n = e / TM;
      double* c = args.C + (long long)(td.c_col + n) * args.ldc + td.c_row + m;
      if (args.epilogue == EPI_SUB) *c -= sums[q];
      else if (args.epilogue == EPI_SETNEG) *c = -sums[q];
      else *c = sums[q];
    }
    __syncthreads();
  }
}

}  // namespace lk
