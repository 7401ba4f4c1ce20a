// gemm_dmma.cuh -- the FP64 tensor-core tile engine behind Cholesky's trailing
// update (SYRK), the panel TRSM, TRTRI and LAUUM.
//
// One persistent kernel, four warps per CTA, two CTAs per SM (96 KB smem each):
//   * warp 0 is the TMA producer, inline in its consumer loop: one elected lane issues the cp.async.bulk.tensor
//     (SWIZZLE_128B) loads of the two operand tiles of k-stage c+3 just before the CTA consumes k-stage c,
//     into a 4-deep shared-memory ring guarded by full/empty mbarriers, running ahead across tile boundaries.
//     (A producer-only fifth warp would put three warps on one SM sub-partition and cap every thread at
//     168 registers -- measured: spills in the main loop; with 2 warps per sub-partition the kernel uses 190 - 214.)
//   * all four warps are consumers: each owns a 64(row) x 32(col) slab of the 64 x 128 output tile as 32
//     independent m8n8k4 FP64 accumulators (DMMA.8x8x4), reads its fragments conflict-free from the swizzled
//     tiles with register + immediate addressing, software-pipelined one k-step ahead, and applies the
//     epilogue (C = acc | C = -acc | C -= acc) with 16-byte accesses.
// One CTA's epilogue overlaps the other CTA's main loop.
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
#include "tile_tables.hpp"

#ifndef LKGPU_RING_RELEASE
#define LKGPU_RING_RELEASE 1  // 1: real per-lane proxy fence before the ring-slot release; 0: round 1's never-taken fence
#endif

namespace lk {

// TM = 64 (C rows per tile), TN = 128 (C cols per tile) and TileDesc live in tile_tables.hpp (plain C++, shared with
// the CPU self-test of the table builders)
constexpr int TK = 16;       // k per pipeline stage (16 doubles = one 128-byte swizzle row)
constexpr int GSTAGES = 4;   // pipeline depth
constexpr int GEMM_CONSUMER_WARPS = 4;
constexpr int GEMM_THREADS = 32 * GEMM_CONSUMER_WARPS;  // the TMA producer is an elected lane of consumer warp 0
constexpr int NS_TILE_BYTES = TN * TK * 8;  // 16384
constexpr int MS_TILE_BYTES = TM * TK * 8;  // 8192
constexpr int STAGE_BYTES = NS_TILE_BYTES + MS_TILE_BYTES;
constexpr int GEMM_SMEM_BYTES = GSTAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

enum GemmEpilogue { EPI_SET = 0, EPI_SETNEG = 1, EPI_SUB = 2 };
enum GemmSched { SCHED_TABLE = 0, SCHED_RECT = 1, SCHED_TRAP = 2 };

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
  // optional: when *abort_flag != 0 at kernel start the launch is a no-op.  Set for the launches of a Cholesky
  // attempt: once a panel has found a non-positive pivot (info != 0) the attempt is discarded by safe_chol_lower's
  // ladder (LinearAlgebra.cpp:66-90), so the rest of its trailing updates is skipped.
  const int* abort_flag;
  // Always 0; only read by the LKGPU_RING_RELEASE=0 build (round 1's never-taken fence, kept for A/B timing).
  int sched_fence;
  // 1: M-major operands are described by 3D tensor maps (one box of 16 k x 64 / 128 rows per stage and operand)
  int mm3;
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

// ---- fragment addressing -------------------------------------------------------------------------------
// Byte offset of element (row, kk) of an operand tile = R[sel(s, idx)] + imm(s, idx), where row = 8 idx + g
// (+ 32 warp on the N side, folded into R by init's side_base), kk is lane t's k index in k-step s, R[] are
// per-lane registers and imm is a compile-time constant (it becomes the LDS immediate).
template <bool KMAJ, bool BOTHK>
struct FragAddr {
  uint32_t R[4];
  __device__ __forceinline__ void init(int g, int t, uint32_t side_base) {
    if (!KMAJ) {
      // M-major: off = ((row>>4)*16 + kk)*128 + ((((row&15)>>1) ^ (kk&7)) << 4 | (row&1) << 3), kk = 2t + (s&1) + 8(s>>1)
#pragma unroll
      for (int sel = 0; sel < 4; ++sel) {
        const int sp = sel >> 1, ip = sel & 1;
        const int kk7 = 2 * t + sp;
        R[sel] = side_base + (uint32_t)(kk7 * 128 + ((((ip * 4 + (g >> 1)) ^ kk7) << 4) | ((g & 1) << 3)));
      }
    } else if (!BOTHK) {
      // K-major beside an M-major operand: off = row*128 + (((kk>>1) ^ (row&7)) << 4 | (kk&1) << 3), same kk
#pragma unroll
      for (int sel = 0; sel < 4; ++sel) R[sel] = side_base + (uint32_t)(g * 128 + ((((t + 4 * (sel & 1)) ^ g) << 4)));
    } else {
      // both K-major: kk = 2s + (t&1) + 8(t>>1)
#pragma unroll
      for (int sel = 0; sel < 4; ++sel)
        R[sel] = side_base + (uint32_t)(g * 128 + ((((sel + 4 * (t >> 1)) ^ g) << 4) | ((t & 1) << 3)));
    }
  }
};
template <bool KMAJ, bool BOTHK, int S, int IDX>
struct FragSel {
  static constexpr int sel = !KMAJ ? ((S & 1) * 2 + (IDX & 1)) : (!BOTHK ? (S >> 1) : S);
  static constexpr int imm = !KMAJ ? ((S >> 1) * 1024 + (IDX >> 1) * 2048) : (!BOTHK ? (IDX * 1024 + (S & 1) * 8) : IDX * 1024);
};
template <int IMM>
__device__ __forceinline__ double lds_f64_imm(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM));
  return v;
}
template <bool KMAJ, bool BOTHK, int S, int IDX, int CNT>
struct FragLoader {
  static __device__ __forceinline__ void run(const uint32_t (&r)[4], double* out) {
    out[IDX] = lds_f64_imm<FragSel<KMAJ, BOTHK, S, IDX>::imm>(r[FragSel<KMAJ, BOTHK, S, IDX>::sel]);
    FragLoader<KMAJ, BOTHK, S, IDX + 1, CNT>::run(r, out);
  }
};
template <bool KMAJ, bool BOTHK, int S, int CNT>
struct FragLoader<KMAJ, BOTHK, S, CNT, CNT> {
  static __device__ __forceinline__ void run(const uint32_t (&)[4], double*) {}
};
template <bool KMAJ, bool BOTHK, int S, int CNT>
__device__ __forceinline__ void load_frags(const uint32_t (&r)[4], double* out) {
  FragLoader<KMAJ, BOTHK, S, 0, CNT>::run(r, out);
}

template <bool MS_KMAJ, bool NS_KMAJ>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_dmma_kernel(const CUtensorMap* tmapM, const CUtensorMap* tmapN, const GemmArgs args) {  // maps in device memory
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B atoms are 1024 bytes: align the ring.
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + GSTAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + GSTAGES;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler too
  const int lane = threadIdx.x & 31;
  // Abort flag: one read per CTA (so that the whole CTA takes the same branch even if the flag is raised right
  // now), issued before the barrier initialisation so that its latency hides behind it.
  __shared__ int s_abort;
  if (threadIdx.x == 0) {
    int aborted = 0;
    if (args.abort_flag != nullptr) aborted = *reinterpret_cast<const volatile int*>(args.abort_flag);
    for (int s = 0; s < GSTAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], GEMM_CONSUMER_WARPS);
    }
    fence_barrier_init();
    s_abort = aborted;
  }
  __syncthreads();
  if (s_abort != 0) return;

  // ===================== TMA producer: warp 0, inline in its consumer loop =====================
  // Four warps per CTA and two CTAs per SM = two warps per SM sub-partition, so each thread may hold the
  // 64 accumulators + double-buffered fragments without spilling (a fifth, producer-only warp would put three
  // warps on one sub-partition and cap everyone at 168 registers).  Warp 0 issues the loads of k-stage
  // c + GSTAGES - 1 just before the CTA consumes k-stage c, i.e. it refills the slot released one iteration ago.
  // All 32 lanes run the producer's bookkeeping on warp-uniform values and one elected lane issues the TMA
  // instructions: ptxas keeps the operands in uniform registers (round 1-2a issued from `threadIdx.x == 0`, which
  // made every UTMALDG a divergent-operand loop -- ELECT / 4 x R2UR / UTMALDG / BRA.U.ANY -- and with 12 boxes per
  // stage in the NT layout warp 0 spent 21 % of a stage in the producer, the other warps waiting on the full barrier
  // behind it: profiles/r02c_syrk_producer.md).  M-major operands come as ONE 3D box per operand (args.mm3, see
  // MatMaps) instead of 8 + 4 two-dimensional ones.
  const bool is_producer = warp == 0;
  int p_tile = blockIdx.x, p_k0 = 0, p_stage = 0;
  uint32_t p_phase = 0;
  bool p_live = false;
  TileDesc p_td = {0, 0, 0, 0};
  if (is_producer) {
    if (elect_one()) {
      tma_prefetch_desc(tmapM);
      tma_prefetch_desc(tmapN);
    }
    p_live = p_tile < args.ntiles;
    if (p_live) {
      p_td = gemm_get_tile(args, p_tile);
      p_k0 = p_td.k_begin;
    }
  }
  auto produce_one = [&]() {
    if (!p_live) return;
    mbar_wait(&empty_bar[p_stage], p_phase ^ 1);
    if (elect_one()) {
      uint8_t* sN = ring + p_stage * STAGE_BYTES;
      uint8_t* sM = sN + NS_TILE_BYTES;
      mbar_arrive_expect_tx(&full_bar[p_stage], STAGE_BYTES);
      if (NS_KMAJ) {
#pragma unroll
        for (int b = 0; b < TN / 64; ++b) tma_load_2d(sN + b * 8192, tmapN, &full_bar[p_stage], p_k0, p_td.c_col + 64 * b);
      } else if (args.mm3) {
        tma_load_3d(sN, tmapN, &full_bar[p_stage], 0, p_k0, p_td.c_col >> 4);
      } else {
#pragma unroll
        for (int b = 0; b < TN / 16; ++b) tma_load_2d(sN + b * 2048, tmapN, &full_bar[p_stage], p_td.c_col + 16 * b, p_k0);
      }
      if (MS_KMAJ) {
        tma_load_2d(sM, tmapM, &full_bar[p_stage], p_k0, p_td.c_row);
      } else if (args.mm3) {
        tma_load_3d(sM, tmapM, &full_bar[p_stage], 0, p_k0, p_td.c_row >> 4);
      } else {
#pragma unroll
        for (int b = 0; b < TM / 16; ++b) tma_load_2d(sM + b * 2048, tmapM, &full_bar[p_stage], p_td.c_row + 16 * b, p_k0);
      }
    }
    if (++p_stage == GSTAGES) {
      p_stage = 0;
      p_phase ^= 1;
    }
    p_k0 += TK;
    if (p_k0 >= p_td.k_end) {
      p_tile += gridDim.x;
      p_live = p_tile < args.ntiles;
      if (p_live) {
        p_td = gemm_get_tile(args, p_tile);
        p_k0 = p_td.k_begin;
      }
    }
  };
  if (is_producer) {
#pragma unroll 1
    for (int q = 0; q < GSTAGES - 1; ++q) produce_one();
  }
  __syncwarp();

  // ===================== DMMA consumers (all four warps) =====================
  const int g = lane >> 2, t = lane & 3;
  const uint32_t ring_u32 = smem_u32(ring);
  int stage = 0;
  uint32_t phase = 0;

  // Per-lane fragment addresses.  k index used by lane t in k-step s (a permutation of 0..15 chosen so that the
  // 16 lanes of a half-warp hit 16 distinct 8-byte bank pairs):
  //   any M-major operand present: kk = 2t + (s&1) + 8(s>>1)
  //   both K-major:                kk = 2s + (t&1) + 8(t>>1)
  // Every fragment address is  stage base + one of <= 4 per-lane registers + a compile-time immediate
  // (FragAddr above), so the main loop is LDS-with-immediate + DMMA only.
  constexpr bool BOTHK = MS_KMAJ && NS_KMAJ;
  FragAddr<NS_KMAJ, BOTHK> fN;
  FragAddr<MS_KMAJ, BOTHK> fM;
  fN.init(g, t, (uint32_t)(warp * 4096));
  fM.init(g, t, (uint32_t)NS_TILE_BYTES);

  for (int tile = blockIdx.x; tile < args.ntiles; tile += gridDim.x) {
    TileDesc td = gemm_get_tile(args, tile);
    double acc[4][8][2];
    double* Cbase = args.C + (long long)(td.c_col + warp * 32 + g) * args.ldc + td.c_row + 2 * t;
    if (args.epilogue == EPI_SUB) {
      // C -= M N^T: the accumulators START at -C (loads issued here, in the shadow of the first TMA round trip) and
      // the epilogue stores -acc: no read-modify-write latency at the end of the tile, where nothing hides it.
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const double* Ccol = Cbase + (long long)(8 * i) * args.ldc;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double2 c = *reinterpret_cast<const double2*>(Ccol + 8 * j);
          acc[i][j][0] = -c.x;
          acc[i][j][1] = -c.y;
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    }

    for (int k0 = td.k_begin; k0 < td.k_end; k0 += TK) {
      if (is_producer) produce_one();
      __syncwarp();
      mbar_wait(&full_bar[stage], phase);
      const uint32_t base = ring_u32 + stage * STAGE_BYTES;
      uint32_t rn[4], rm[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        rn[q] = base + fN.R[q];
        rm[q] = base + fM.R[q];
      }
      double a[2][4], b[2][8];
      load_frags<NS_KMAJ, BOTHK, 0, 4>(rn, a[0]);
      load_frags<MS_KMAJ, BOTHK, 0, 8>(rm, b[0]);
#pragma unroll
      for (int s = 0; s < 4; ++s) {
        // software pipeline: fragments of k-step s+1 are in flight while the 32 DMMAs of k-step s issue
        if (s == 0) {
          load_frags<NS_KMAJ, BOTHK, 1, 4>(rn, a[1]);
          load_frags<MS_KMAJ, BOTHK, 1, 8>(rm, b[1]);
        } else if (s == 1) {
          load_frags<NS_KMAJ, BOTHK, 2, 4>(rn, a[0]);
          load_frags<MS_KMAJ, BOTHK, 2, 8>(rm, b[0]);
        } else if (s == 2) {
          load_frags<NS_KMAJ, BOTHK, 3, 4>(rn, a[1]);
          load_frags<MS_KMAJ, BOTHK, 3, 8>(rm, b[1]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[s & 1][i], b[s & 1][j]);
      }
      // Release of the ring slot -- ordered by construction.  The slot is refilled by TMA (async proxy) while it was
      // read by LDS (generic proxy): every lane first executes fence.proxy.async.shared::cta, which orders ITS reads
      // of the slot before later async-proxy accesses (and, as a memory fence, cannot complete before those loads
      // have), the lanes then meet at bar.warp.sync, and only then lane 0 arrives (release) on the empty barrier that
      // the producer acquires before issuing the refill.  History: without any fence ptxas 12.9 placed the
      // SYNCS.ARRIVE two instructions after the last LDS *issue*, ahead of the DMMAs consuming those loads, and
      // under heavy shared-memory traffic the refill overtook loads still queued (profiles/r01c_ring_release.md;
      // found with tools/diag_foreign.py).  Round 1 pinned the arrive with a never-taken fence -- a scheduling
      // trick; LKGPU_RING_RELEASE=0 still builds that variant for A/B timing.  tests/test_abi.py checks the SASS.
#if LKGPU_RING_RELEASE
      fence_proxy_async_smem();
      __syncwarp();
#else
      __syncwarp();
      if (args.sched_fence) fence_proxy_async();
#endif
      if (lane == 0) mbar_arrive(&empty_bar[stage]);
      if (++stage == GSTAGES) {
        stage = 0;
        phase ^= 1;
      }
    }

    // ---- epilogue: C = acc (EPI_SET) | C = -acc (EPI_SETNEG, and EPI_SUB whose accumulators started at -C) ----
    const double sgn = (args.epilogue == EPI_SET) ? 1.0 : -1.0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      double* Ccol = Cbase + (long long)(8 * i) * args.ldc;
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

// Plain CUDA-core restatement of the same tile contract, used ONLY by the GPU
// test-suite (LKGPU_DEBUG_SIMPLE_GEMM=1) to localise faults in the DMMA/TMA path.
template <bool MS_KMAJ, bool NS_KMAJ>
__global__ void gemm_simple_kernel(const double* __restrict__ Mbuf, long long ldm, const double* __restrict__ Nbuf,
                                   long long ldn, const GemmArgs args) {
  // blockDim.x must be 256: each thread owns 32 elements; sums are formed before any store so that the
  // in-place panel TRSM (C aliases the M-side operand) is safe, exactly as in the DMMA kernel.
  if (args.abort_flag != nullptr && *reinterpret_cast<const volatile int*>(args.abort_flag) != 0) return;
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
