// engine.cu -- the lkgpu engine: device workspaces (a1: KModel), the evaluation
// pipeline (a2..a10 of SURVEY.md §8) and the C ABI of include/lkgpu.h.
//
// Device layout (all column-major fp64, N = n rounded up to 128, padding = identity):
//   A  N x N : R(theta) lower tiles  -> overwritten by L (lower), diagonal blocks fully stored
//   W  N x N : diagonal blocks = inverses of L's diagonal blocks (written by the panel kernel)
//              -> after TRTRI the whole lower triangle holds L^-1
//   V  N x N : scratch for TRTRI, then R^-1 = L^-T L^-1 (lower tiles) after LAUUM
//   Bv N x (p+1) : [F | y] -> [Fstar | ystar] ;  Ev, Xv : N vectors (Estar, x)
// No CPU fallback anywhere: every entry point needs a CUDA device.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <mutex>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lkgpu.h"
#include "common.cuh"
#include "cov.cuh"
#include "gemm_dmma.cuh"
#include "potrf_panel.cuh"
#include "trsv.cuh"
#include "trsv_wave.cuh"
#include "variogram.cuh"

using namespace lk;

namespace {

thread_local std::string g_last_error;

// Hardware work queues: the default of 8 connections maps the 2 streams x (up to 16) handles that run concurrently
// on one device onto 8 queues, with false dependencies between streams that share one (measured, 8 handles of
// n = 5000 in flight: 5.13 ms per evaluation with 8 connections, 4.70 ms with 32).  The variable is read when the CUDA
// context is created, so it is set when this library is loaded -- unless the user has set it.
__attribute__((constructor)) void lk_set_max_connections() { setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0); }
// Per-device gate (see Engine::SweepGate): mid-size handles overlap, large unflagged ones take the device in turn.
struct DeviceGate {
  std::mutex m;
  std::condition_variable cv;
  int shared_active = 0;       // evaluations of flagged handles in flight
  bool exclusive_active = false;
  int exclusive_waiting = 0;   // unflagged evaluations waiting for the device (flagged newcomers queue behind them)
};
DeviceGate g_gate[64];

struct LkError {
  std::string msg;
};

#define CUDA_CHECK(expr)                                                                          \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      throw LkError{std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                    ":" + std::to_string(__LINE__) + ")"};                                        \
    }                                                                                             \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    if (!p || qres != cudaDriverEntryPointSuccess) throw LkError{"cuTensorMapEncodeTiled not available"};
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 2D map over a column-major N x ncols fp64 matrix: dim0 = rows (contiguous), dim1 = columns.
CUtensorMap make_map(double* base, long long rows, long long cols, long long ld, int box_rows, int box_cols,
                     bool swizzle128 = true) {
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)rows, (cuuint64_t)cols};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {(cuuint32_t)box_rows, (cuuint32_t)box_cols};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE,
                               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw LkError{"cuTensorMapEncodeTiled failed with code " + std::to_string((int)r)};
  return m;
}

// 3D map of the same matrix for M-major operand tiles: dim0 = row inside a 16-row group (contiguous), dim1 = column
// (k), dim2 = group of 16 rows (stride 128 B).  A box {16, 16, G} lands in shared memory as [group][k][16 rows] =
// exactly G consecutive boxes {16 rows, 16 columns} of the 2D map, so one instruction per operand and stage replaces
// 8 (N side, G = 8) or 4 (M side, G = 4); the 128-byte swizzle is a function of the shared-memory address alone and
// stays what the fragment loads expect.  ok = false when the driver rejects the descriptor (the 2D path is used).
CUtensorMap make_map3(double* base, long long rows, long long cols, long long ld, int groups, bool* ok) {
  CUtensorMap m;
  memset(&m, 0, sizeof(m));
  cuuint64_t gdim[3] = {16, (cuuint64_t)cols, (cuuint64_t)(rows / 16)};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 8, 128};
  cuuint32_t box[3] = {16, 16, (cuuint32_t)groups};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, gdim, gstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  *ok = (r == CUDA_SUCCESS);
  return m;
}

// The tensor maps of one matrix live in DEVICE memory, one allocation per matrix that is written once and never
// modified; kernels receive pointers to them (two 8-byte kernel parameters instead of two 128-byte __grid_constant__
// descriptors, and the same descriptors serve every launch of the handle).
struct MatMaps {
  const CUtensorMap* mm = nullptr;  // M-major operand tiles: box {16 rows, 16 k-columns}
  const CUtensorMap* km = nullptr;  // K-major operand tiles: box {16 k-rows, 64 columns}
  const CUtensorMap* wf = nullptr;  // wavefront sweep, forward: box {128 rows, 32 columns}, dense
  const CUtensorMap* wb = nullptr;  // wavefront sweep, backward: box {16 rows, 128 columns}, SWIZZLE_128B
  const CUtensorMap* mm8 = nullptr;  // M-major operand tiles, 3D: box {16 rows, 16 k-columns, 8 row groups} (N side)
  const CUtensorMap* mm4 = nullptr;  // the same with 4 row groups (M side); both null when the driver rejects them
};

// FP64 peak probe kernels ---------------------------------------------------
__global__ void __launch_bounds__(256) probe_dmma_kernel(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) probe_dfma_kernel(double* out, int iters) {
  double c[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) c[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 32; ++i) c[i] = fma(c[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 32; ++i) s += c[i];
  if (s == 123.456) out[0] = s;
}
__global__ void __launch_bounds__(256) probe_mixed_kernel(double* out, int iters) {
  double c[8][2], f[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = i;
  double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dmma884(c[i][0], c[i][1], a, b);
      f[2 * i] = fma(f[2 * i], a, b);
      f[2 * i + 1] = fma(f[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  if (s == 123.456) out[0] = s;
}

// copy helpers ----------------------------------------------------------------
__global__ void pad_copy_kernel(const double* __restrict__ src, long long lds, int n, int cols, double* __restrict__ dst,
                                long long ldd, int N) {
  // dst (N x cols) = [src (n x cols); 0]
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)N * cols) return;
  const int c = (int)(idx / N), r = (int)(idx % N);
  dst[(long long)c * ldd + r] = (r < n) ? src[(long long)c * lds + r] : 0.0;
}

// Device-side fills and copies in the evaluation path are small kernels on the handle's own stream (they stay in
// stream order with the compute kernels and are counted in lkgpu_launch_count).
__global__ void zero_ints_kernel(int* p, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = 0;
}
__global__ void fill_zero_kernel(double* __restrict__ p, long long count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    p[i] = 0.0;
}
__global__ void copy_diag_blocks_kernel(double* __restrict__ dst, long long dst_ld, long long dst_blk,
                                       const double* __restrict__ src, long long src_ld, long long src_blk) {
  double* d_ = dst + blockIdx.x * dst_blk;
  const double* s_ = src + blockIdx.x * src_blk;
  for (int e = threadIdx.x; e < BLK * BLK; e += blockDim.x) {
    const int c = e / BLK, r = e % BLK;
    d_[c * dst_ld + r] = s_[c * src_ld + r];
  }
}
__global__ void copy_kernel(double* __restrict__ dst, const double* __restrict__ src, long long count) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// dscal slot map (device scalars copied back at the end of an evaluation)
enum {
  SC_SUMLOG = 0, SC_SSE = 1, SC_DIAG = 4 /*4 slots*/, SC_LOO = 8, SC_NORML = 9, SC_NORMW = 10,
  SC_GRAD = 16 /* 2*(LK_MAX_D+1) slots */, SC_BOUNDS = 160 /* 2*LK_MAX_D slots */, SC_COUNT = 320
};

struct Engine {
  int device = 0, n = 0, d = 0, p = 0, kernel = 0, noise_model = 0;
  int N = 0, nb = 0;
  long long ld = 0;
  int sm_count = 148;
  bool debug_simple = false;
  bool use_lookahead = true;
  bool use_step_trsv = false;
  // Evaluations of handles with n <= overlap_max_n overlap on the device by default (a mid-size factorisation cannot
  // fill 148 SMs); larger ones queue unless flagged by lkgpu_set_concurrent.  LKGPU_OVERLAP_MAX_N overrides (0: never).
  int overlap_max_n = 8192;
  int persistent_update_reserve = 0;
  bool l2_order = true;  // LKGPU_NO_L2_ORDER=1: tile tables sorted by k-length only (the r01c order)
  int wave_grid_cap = 0;  // LKGPU_WAVE_GRID=k: at most k CTAs per sweep (fault localisation)
  bool no_mm3 = false;  // LKGPU_NO_MM3=1: M-major operands as 2D boxes (12 TMA instructions per NT stage instead of 2)
  bool no_persistent = false;  // LKGPU_NO_PERSISTENT=1: one CTA per tile everywhere (fault localisation)
  bool ladder_shortcut = true;  // lkgpu_set_ladder_shortcut / LKGPU_FULL_LADDER=1 (see Engine::eval)
  bool use_abort = true;  // LKGPU_NO_ABORT=1: failed Cholesky attempts run to the end (fault localisation)
  int outer_panels = 0;  // Cholesky outer block = outer_panels * 128 columns; 0 = by size (LKGPU_OUTER_PANELS overrides)
  // numerics (LinearAlgebra statics of the reference)
  double num_nugget = 1e-10, min_rcond = 1e-18;
  int max_inc = 10;
  bool rcond_check = true;
  // model flags / fixed values used by the objective assembly (m_est_sigma2, m_sigma2, ... of the reference)
  bool est_sigma2 = true, est_nugget = true;
  double sigma2 = 1.0, nugget = 0.0, alpha0 = 1.0;
  std::vector<double> xmin, xmax;
  // m_est_beta == false: the trend coefficients the caller fixed (lkgpu_set_fixed_beta); they only enter the
  // committed z = ystar - Fstar beta used by predict (Kriging.cpp:2168-2172), never the objective (KrigingImpl.cpp:113-123)
  bool has_fixed_beta = false;
  std::vector<double> fixed_beta;

  double *dX = nullptr, *dy = nullptr, *dF = nullptr, *dnoise = nullptr;
  double *A = nullptr, *W = nullptr, *V = nullptr;
  double *Bv = nullptr, *Ev = nullptr, *Xv = nullptr, *Tv = nullptr;  // rhs / vectors (N rows)
  double *Uv = nullptr;                                               // N x p (LMP/LOO: U = Rinv F LX^-T)
  double *Zv = nullptr, *Qv = nullptr;                                // N x (p+1) [Rinv_X | yt_Rinv] ; N (Qo)
  double *dS2loo = nullptr, *dErr = nullptr, *dSqrtC = nullptr, *dEs = nullptr;  // LOO row vectors (N)
  double *Q1 = nullptr, *Q2 = nullptr;                                // LOO gradient only (lazily allocated N x N)
  double* dsmall = nullptr;                                           // (p+1)^2 + p small device matrix
  TileDesc* loo_table = nullptr;
  int loo_tiles = 0;
  double h_sum_log_diagLX = 0.0, h_S2 = 0.0;
  bool have_loo = false;
  double* logdet_blocks = nullptr;
  int* dinfo = nullptr;       // [0] chol info, [1] gls info
  int* wave_ctl = nullptr;    // wavefront sweeps: [0] ticket counter
  double* wave_z = nullptr;   // wavefront sweeps: published solution blocks, N x WAVE_MAX_RHS (trsv_wave.cuh)
  double* dscal = nullptr;    // small device scalars / results (64 doubles)
  double* dpartial = nullptr; // reduction partials
  size_t partial_doubles = 0;
  double* dcolsum = nullptr;  // N
  double *dRstar = nullptr, *dbeta = nullptr;
  TileDesc *trtri_tables = nullptr, *lauum_table = nullptr;
  int lauum_tiles = 0;
  // Cholesky trailing updates in L2 order (build_chol_plans): one look-ahead + one rest table per outer block
  TileDesc* chol_tables = nullptr;
  struct CholPlan {
    size_t off_la = 0, n_la = 0, off_rest = 0, n_rest = 0;
  };
  std::vector<CholPlan> chol_plans;  // indexed by J0 / chol_plan_OB
  int chol_plan_OB = 0;
  // LKGPU_LAUUM_ORDER: 0 super-tiles (order_for_l2), 1 rows (longest k first) in serpentine rounds, 2 rows, 3 bands of
  // LKGPU_LAUUM_BAND row tiles in serpentine rounds.  n = 20000, DRAM reads / L2 hit rate / time of the launch: 0: 121 GB,
  // 51 %, 77.07 ms; 1: 101 GB, 56 %, 76.88; 2: 148 GB, 45 %, 77.30; 3 with bands of 4 / 8 / 16 / 32: 90 / 67 / 57 / 75 GB,
  // 60 / 69 / 70 / 64 %, 76.60 / 76.61 / 76.58 / 76.67 ms (profiles/r02c_table_orders.log)
  int lauum_order = 3;
  int lauum_band = 16;
  int trtri_order = 1;  // LKGPU_TRTRI_ORDER: 0 longest k first, 1 + serpentine rounds
  int chol_band = 32;  // 64-row tiles per band of the rest update (LKGPU_CHOL_BAND; 0: closed-form column order)
  struct TrtriLevel {
    size_t off1, n1, off2, n2;
  };
  std::vector<TrtriLevel> trtri_levels;
  MatMaps mapA, mapW, mapV, mapQ1;
  cudaStream_t s_main = nullptr, s_upd = nullptr;
  std::vector<cudaEvent_t> ev_panel, ev_upd;
  cudaEvent_t ev_t[LKGPU_N_STAGES + 2];
  cudaEvent_t ev_misc = nullptr;
  long long launches = 0;
  double* hpin = nullptr;  // pinned host staging
  size_t hpin_doubles = 0;
  KernelParams kp;
  // state of the last evaluation
  bool have_model = false, have_W = false, have_V = false, have_x = false;
  double last_alpha = 1.0, last_inv_sigma2 = 0.0, last_diag_add = 0.0;
  int last_n_jitter = 0;
  long long evals_done = 0;  // evaluations completed on this handle since its data were set (ladder shortcut history)
  std::vector<double> last_theta;
  // factor kept across lkgpu_append_data (the reference's m_T while m_X has more rows than m_T)
  int keep_n = 0;
  std::vector<double> keep_theta;
  double keep_alpha = 1.0, keep_inv_sigma2 = 0.0, keep_diag_add = 0.0;
  bool last_was_update = false, last_mixed_jitter = false;
  // committed model: the reference's members m_T, m_M, m_z, m_circ, m_beta (commit at Kriging.cpp:2156-2173), which
  // live apart from the per-evaluation KModel workspaces.  One extra n x n buffer, allocated on the first commit.
  struct Committed {
    bool valid = false, have_x = false, mixed = false;
    double *L = nullptr, *D = nullptr, *Bv = nullptr, *Ev = nullptr, *Xv = nullptr, *Rstar = nullptr, *beta = nullptr,
           *logdet = nullptr;
    std::vector<double> theta;
    double alpha = 1.0, inv_sigma2 = 0.0, diag_add = 0.0;
    int n_jitter = 0;
  } cm;
  bool live_is_committed = false;

  ~Engine() { release(); }
  // everything whose size depends on n (lkgpu_append_data re-creates these for the extended data set)
  void release_sized() {
    cudaSetDevice(device);
    double* bufs[] = {dX, dy, dF, dnoise, A, W, V, Bv, Ev, Xv, Tv, Uv, Zv, Qv, dS2loo, dErr, dSqrtC, dEs, Q1, Q2, dsmall,
                      logdet_blocks, dscal, dpartial, dcolsum, dRstar, dbeta};
    for (double* b : bufs)
      if (b) cudaFree(b);
    if (dinfo) cudaFree(dinfo);
    if (wave_ctl) cudaFree(wave_ctl);
    wave_ctl = nullptr;
    if (wave_z) cudaFree(wave_z);
    wave_z = nullptr;
    if (ladder_D) cudaFree(ladder_D);
    if (ladder_logdet) cudaFree(ladder_logdet);
    ladder_D = ladder_logdet = nullptr;
    if (trtri_tables) cudaFree(trtri_tables);
    if (lauum_table) cudaFree(lauum_table);
    if (chol_tables) cudaFree(chol_tables);
    chol_tables = nullptr;
    chol_plans.clear();
    if (loo_table) cudaFree(loo_table);
    if (hpin) cudaFreeHost(hpin);
    for (CUtensorMap* m : map_allocs) cudaFree(m);
    map_allocs.clear();
    for (auto e : ev_panel) cudaEventDestroy(e);
    for (auto e : ev_upd) cudaEventDestroy(e);
    {
      double* cbufs[] = {cm.L, cm.D, cm.Bv, cm.Ev, cm.Xv, cm.Rstar, cm.beta, cm.logdet};
      for (double* b : cbufs)
        if (b) cudaFree(b);
      cm.L = cm.D = cm.Bv = cm.Ev = cm.Xv = cm.Rstar = cm.beta = cm.logdet = nullptr;
      cm.valid = false;
      live_is_committed = false;
    }
    dX = dy = dF = dnoise = A = W = V = Bv = Ev = Xv = Tv = Uv = logdet_blocks = dscal = dpartial = dcolsum = dRstar = dbeta = nullptr;
    Zv = Qv = dS2loo = dErr = dSqrtC = dEs = Q1 = Q2 = dsmall = nullptr;
    dinfo = nullptr;
    trtri_tables = lauum_table = loo_table = nullptr;
    loo_tiles = lauum_tiles = 0;
    hpin = nullptr;
    ev_panel.clear();
    ev_upd.clear();
  }
  void release() {
    release_sized();
    for (auto& e : ev_t)
      if (e) cudaEventDestroy(e);
    for (auto& e : ev_t) e = nullptr;
    if (ev_misc) cudaEventDestroy(ev_misc);
    ev_misc = nullptr;
    if (s_main) cudaStreamDestroy(s_main);
    if (s_upd) cudaStreamDestroy(s_upd);
    s_main = s_upd = nullptr;
  }

  template <typename T>
  T* dalloc(size_t count) {
    T* p_ = nullptr;
    CUDA_CHECK(cudaMalloc(&p_, std::max<size_t>(count, 1) * sizeof(T)));
    return p_;
  }

  void init(int device_, int n_, int d_, int p_, const double* X, const double* y, const double* F, const double* noise,
            int kernel_, int noise_model_) {
    device = device_;
    n = n_;
    d = d_;
    p = p_;
    kernel = kernel_;
    noise_model = noise_model_;
    if (n < 1 || d < 1 || d > LK_MAX_D) throw LkError{"lkgpu_create: need n >= 1 and 1 <= d <= 64"};
    if (p < 0 || p > 128) throw LkError{"lkgpu_create: need 0 <= p <= 128 trend columns"};  // p = 0: regmodel "none"
    if (kernel < 0 || kernel > 3) throw LkError{"lkgpu_create: unknown kernel id"};
    if (noise_model < 0 || noise_model > 2) throw LkError{"lkgpu_create: unknown noise model"};
    if (noise_model == LKGPU_NOISE_HETERO && !noise) throw LkError{"lkgpu_create: heterogeneous noise needs a noise vector"};
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
      throw LkError{"lkgpu: no CUDA device available (this engine has no CPU fallback)"};
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) throw LkError{"lkgpu: built for sm_100a only; device is sm_" + std::to_string(prop.major * 10 + prop.minor)};
    sm_count = prop.multiProcessorCount;
    const char* dbg = getenv("LKGPU_DEBUG_SIMPLE_GEMM");
    debug_simple = dbg && dbg[0] == '1';
    const char* nla = getenv("LKGPU_NO_LOOKAHEAD");
    use_lookahead = !(nla && nla[0] == '1');
    if (const char* op = getenv("LKGPU_OUTER_PANELS")) outer_panels = std::max(1, std::min(16, atoi(op)));
    if (const char* wg = getenv("LKGPU_WAVE_GRID")) wave_grid_cap = atoi(wg);
    if (const char* v = getenv("LKGPU_PERSISTENT_UPDATE")) persistent_update_reserve = std::max(0, atoi(v));
    if (const char* v = getenv("LKGPU_OVERLAP_MAX_N")) overlap_max_n = std::max(0, atoi(v));
    if (const char* nlo = getenv("LKGPU_NO_L2_ORDER")) l2_order = !(nlo[0] == '1');
    if (const char* v = getenv("LKGPU_CHOL_BAND")) chol_band = std::max(0, atoi(v));
    if (const char* v = getenv("LKGPU_LAUUM_ORDER")) lauum_order = atoi(v);
    if (const char* v = getenv("LKGPU_TRTRI_ORDER")) trtri_order = atoi(v);
    if (const char* v = getenv("LKGPU_LAUUM_BAND")) lauum_band = std::max(1, atoi(v));
    if (const char* v = getenv("LKGPU_NO_MM3")) no_mm3 = v[0] == '1';
    if (const char* v = getenv("LKGPU_TRACE_CHOL")) trace_chol = v[0] == '1';
    if (const char* v = getenv("LKGPU_UPDATE_AFTER_LOOKAHEAD")) update_after_lookahead = v[0] != '0' ? 1 : 0;
    const char* npe = getenv("LKGPU_NO_PERSISTENT");
    no_persistent = npe && npe[0] == '1';
    const char* nab = getenv("LKGPU_NO_ABORT");
    use_abort = !(nab && nab[0] == '1');
    if (const char* fl = getenv("LKGPU_FULL_LADDER")) ladder_shortcut = !(fl[0] == '1');
    const char* stv = getenv("LKGPU_STEP_TRSV");
    use_step_trsv = stv && stv[0] == '1';

    int lo_pri = 0, hi_pri = 0;
    CUDA_CHECK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
    if (const char* npr = getenv("LKGPU_NO_PRIORITY"); npr && npr[0] == '1') hi_pri = lo_pri;  // fault localisation
    CUDA_CHECK(cudaStreamCreateWithPriority(&s_main, cudaStreamNonBlocking, hi_pri));
    CUDA_CHECK(cudaStreamCreateWithPriority(&s_upd, cudaStreamNonBlocking, lo_pri));
    for (auto& ev : ev_t) CUDA_CHECK(cudaEventCreate(&ev));
    CUDA_CHECK(cudaEventCreateWithFlags(&ev_misc, cudaEventDisableTiming));
#define LK_WAVE_ATTR(BW, NQ_) \
  CUDA_CHECK(cudaFuncSetAttribute(trsv_wave_kernel<BW, NQ_>, cudaFuncAttributeMaxDynamicSharedMemorySize, wave_smem_bytes(NQ_)))
    LK_WAVE_ATTR(false, 1); LK_WAVE_ATTR(false, 2); LK_WAVE_ATTR(false, 4); LK_WAVE_ATTR(false, 8);
    LK_WAVE_ATTR(true, 1); LK_WAVE_ATTR(true, 2); LK_WAVE_ATTR(true, 4); LK_WAVE_ATTR(true, 8);
#undef LK_WAVE_ATTR
    CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(gemm_dmma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
    CUDA_CHECK(cudaFuncSetAttribute(potf2_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, POTF2_SMEM_BYTES));
    alloc_sized(noise != nullptr);
    set_data(X, y, F, noise);
    CUDA_CHECK(cudaStreamSynchronize(s_main));
  }

  std::vector<CUtensorMap*> map_allocs;  // device copies of the tensor maps (freed with the sized buffers)
  MatMaps maps_of(double* buf, bool with_sweep_maps) {
    CUtensorMap h[6];
    bool ok8 = false, ok4 = false;
    h[4] = make_map3(buf, N, N, ld, 8, &ok8);
    h[5] = make_map3(buf, N, N, ld, 4, &ok4);
    const bool use3 = ok8 && ok4 && !no_mm3;
    if (!(ok8 && ok4)) {
      static bool warned = false;
      if (!warned) fprintf(stderr, "[lkgpu] 3D tensor maps rejected by the driver; M-major operands use 2D boxes\n");
      warned = true;
    }
    h[0] = make_map(buf, N, N, ld, 16, 16);
    h[1] = make_map(buf, N, N, ld, 16, 64);
    if (with_sweep_maps) {
      h[2] = make_map(buf, N, N, ld, 128, 32, false);
      h[3] = make_map(buf, N, N, ld, 16, 128);
    } else {
      h[2] = h[0];
      h[3] = h[0];
    }
    CUtensorMap* dm = nullptr;
    CUDA_CHECK(cudaMalloc(&dm, 6 * sizeof(CUtensorMap)));
    map_allocs.push_back(dm);
    CUDA_CHECK(cudaMemcpy(dm, h, 6 * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    return MatMaps{dm + 0, dm + 1, dm + 2, dm + 3, use3 ? dm + 4 : nullptr, use3 ? dm + 5 : nullptr};
  }

  // device workspaces, tensor maps and tile plans for the current n (a1: KModel)
  void alloc_sized(bool with_noise) {
    N = ((n + BLK - 1) / BLK) * BLK;
    nb = N / BLK;
    ld = N;
    ev_panel.resize(nb);
    ev_upd.resize(nb);
    for (int i = 0; i < nb; ++i) {
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_panel[i], cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&ev_upd[i], cudaEventDisableTiming));
    }
    dX = dalloc<double>((size_t)n * d);
    dy = dalloc<double>(n);
    dF = dalloc<double>((size_t)n * p);
    if (with_noise) dnoise = dalloc<double>(n);
    A = dalloc<double>((size_t)N * N);
    W = dalloc<double>((size_t)N * N);
    V = dalloc<double>((size_t)N * N);
    Bv = dalloc<double>((size_t)N * (p + 1));
    Ev = dalloc<double>(N);
    Xv = dalloc<double>(N);
    Tv = dalloc<double>((size_t)N * 2);
    Uv = dalloc<double>((size_t)N * p);
    Zv = dalloc<double>((size_t)N * (p + 1));
    Qv = dalloc<double>(N);
    dS2loo = dalloc<double>(N);
    dErr = dalloc<double>(N);
    dSqrtC = dalloc<double>(N);
    dEs = dalloc<double>(N);
    dsmall = dalloc<double>((size_t)(p + 1) * (p + 1) + p + 8);
    logdet_blocks = dalloc<double>(nb);
    dinfo = dalloc<int>(4);
    dscal = dalloc<double>(SC_COUNT);
    CUDA_CHECK(cudaMemset(dscal, 0, SC_COUNT * 8));
    partial_doubles = (size_t)std::max({(size_t)2 * sm_count * 2 * (LK_MAX_D + 1) + 4 * 64,
                                        (size_t)((n + GRAM_CHUNK - 1) / GRAM_CHUNK) * ((size_t)(p + 1) * (p + 2)) + 4 * 64,
                                        (size_t)(N / 256 + 1) + 4 * 64, (size_t)4096});
    dpartial = dalloc<double>(partial_doubles);
    dcolsum = dalloc<double>(N);
    dRstar = dalloc<double>((size_t)p * p);
    dbeta = dalloc<double>(p);
    hpin_doubles = std::max<size_t>((size_t)N + 256, (size_t)SC_COUNT + 512 + (size_t)(p + 1) * (p + 1) + p);
    CUDA_CHECK(cudaMallocHost(&hpin, hpin_doubles * sizeof(double)));
    CUDA_CHECK(cudaMemsetAsync(W, 0, (size_t)N * N * 8, s_main));
    CUDA_CHECK(cudaMemsetAsync(V, 0, (size_t)N * N * 8, s_main));
    CUDA_CHECK(cudaMemsetAsync(A, 0, (size_t)N * N * 8, s_main));
    mapA = maps_of(A, true);
    mapW = maps_of(W, true);
    mapV = maps_of(V, true);
    wave_ctl = dalloc<int>(nb + 1);
    wave_z = dalloc<double>((size_t)N * WAVE_MAX_RHS);
    build_plans();
  }

  // ---- commit / restore (the reference moves km.L, km.Fstar, ... into m_T, m_M, ...: Kriging.cpp:2156-2173) ----
  void copy_diag_blocks(double* dst, long long dst_ld, long long dst_blk, const double* src, long long src_ld,
                        long long src_blk) {
    ++launches;
    copy_diag_blocks_kernel<<<nb, 256, 0, s_main>>>(dst, dst_ld, dst_blk, src, src_ld, src_blk);
    CUDA_CHECK(cudaGetLastError());
  }
  void commit() {
    CUDA_CHECK(cudaSetDevice(device));
    SweepGate gate(*this);
    if (!have_model) throw LkError{"lkgpu_commit_model: no evaluation has been run on this handle"};
    if (!cm.L) {
      cm.L = dalloc<double>((size_t)N * N);
      cm.D = dalloc<double>((size_t)nb * BLK * BLK);
      cm.Bv = dalloc<double>((size_t)N * (p + 1));
      cm.Ev = dalloc<double>(N);
      cm.Xv = dalloc<double>(N);
      cm.Rstar = dalloc<double>((size_t)p * p);
      cm.beta = dalloc<double>(p);
      cm.logdet = dalloc<double>(nb);
    }
    dev_copy(cm.L, A, (long long)N * N);
    copy_diag_blocks(cm.D, BLK, (long long)BLK * BLK, W, ld, (long long)BLK * ld + BLK);
    dev_copy(cm.Bv, Bv, (long long)N * (p + 1));
    dev_copy(cm.Ev, Ev, N);
    dev_copy(cm.Xv, Xv, N);
    dev_copy(cm.Rstar, dRstar, (long long)p * p);
    dev_copy(cm.beta, dbeta, p);
    dev_copy(cm.logdet, logdet_blocks, nb);
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    cm.theta = last_theta;
    cm.alpha = last_alpha;
    cm.inv_sigma2 = last_inv_sigma2;
    cm.diag_add = last_diag_add;
    cm.n_jitter = last_n_jitter;
    cm.mixed = last_mixed_jitter;
    cm.have_x = have_x;
    cm.valid = true;
    live_is_committed = true;
  }
  void restore() {
    CUDA_CHECK(cudaSetDevice(device));
    SweepGate gate(*this);
    if (!cm.valid) throw LkError{"lkgpu_restore_model: no committed model on this handle"};
    if (live_is_committed) return;
    dev_copy(A, cm.L, (long long)N * N);
    copy_diag_blocks(W, ld, (long long)BLK * ld + BLK, cm.D, BLK, (long long)BLK * BLK);
    dev_copy(Bv, cm.Bv, (long long)N * (p + 1));
    dev_copy(Ev, cm.Ev, N);
    dev_copy(Xv, cm.Xv, N);
    dev_copy(dRstar, cm.Rstar, (long long)p * p);
    dev_copy(dbeta, cm.beta, p);
    dev_copy(logdet_blocks, cm.logdet, nb);
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    last_theta = cm.theta;
    last_alpha = cm.alpha;
    last_inv_sigma2 = cm.inv_sigma2;
    last_diag_add = cm.diag_add;
    last_n_jitter = cm.n_jitter;
    last_mixed_jitter = cm.mixed;
    have_model = true;
    have_x = cm.have_x;
    have_W = have_V = have_loo = false;  // W holds the diagonal-block inverses only; L^-1 / R^-1 are re-derived on demand
    keep_n = 0;
    live_is_committed = true;
  }

  // ---- Kriging::update, data side (Kriging.cpp:2476-2491, KrigingImpl.cpp:576-610): append n_u observations.
  // The workspaces are re-created for n + n_u rows.  The factor of the last evaluation (the committed model: the
  // reference's m_T) is kept -- its whole 128-column panels; the rows of a partial last panel are re-derived -- so
  // that the next evaluation at the SAME (theta, extra) runs as a block extension (factor_update below) instead of
  // a factorisation from scratch: populate_Model's `update_eligible` (Kriging.cpp:170-188).
  void append(int n_u, const double* X_u, const double* y_u, const double* F_u, const double* noise_u) {
    CUDA_CHECK(cudaSetDevice(device));
    SweepGate gate(*this);
    if (n_u < 1) throw LkError{"lkgpu_append_data: need n_u >= 1"};
    if (noise_model == LKGPU_NOISE_HETERO && !noise_u) throw LkError{"lkgpu_append_data: heterogeneous noise needs noise_u"};
    if (cm.valid && !live_is_committed) restore();  // the factor that is extended is the committed one (m_T)
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    CUDA_CHECK(cudaStreamSynchronize(s_upd));
    const int n_old = n;
    const long long ld_old = ld;
    double *A_old = A, *W_old = W, *ldb_old = logdet_blocks, *X_old = dX, *y_old = dy, *F_old = dF, *nz_old = dnoise;
    A = W = logdet_blocks = dX = dy = dF = dnoise = nullptr;
    const bool keep = have_model;
    release_sized();
    auto free_old = [&]() {
      double* olds[] = {A_old, W_old, ldb_old, X_old, y_old, F_old, nz_old};
      for (double* b : olds)
        if (b) cudaFree(b);
    };
    try {
      n = n_old + n_u;
      alloc_sized(nz_old != nullptr);
      // data: column-major with the new column stride n
      CUDA_CHECK(cudaMemcpy2DAsync(dX, (size_t)n * 8, X_old, (size_t)n_old * 8, (size_t)n_old * 8, d, cudaMemcpyDeviceToDevice, s_main));
      CUDA_CHECK(cudaMemcpy2DAsync(dX + n_old, (size_t)n * 8, X_u, (size_t)n_u * 8, (size_t)n_u * 8, d, cudaMemcpyHostToDevice, s_main));
      if (p > 0) {
        CUDA_CHECK(cudaMemcpy2DAsync(dF, (size_t)n * 8, F_old, (size_t)n_old * 8, (size_t)n_old * 8, p, cudaMemcpyDeviceToDevice, s_main));
        CUDA_CHECK(cudaMemcpy2DAsync(dF + n_old, (size_t)n * 8, F_u, (size_t)n_u * 8, (size_t)n_u * 8, p, cudaMemcpyHostToDevice, s_main));
      }
      CUDA_CHECK(cudaMemcpyAsync(dy, y_old, (size_t)n_old * 8, cudaMemcpyDeviceToDevice, s_main));
      CUDA_CHECK(cudaMemcpyAsync(dy + n_old, y_u, (size_t)n_u * 8, cudaMemcpyHostToDevice, s_main));
      if (dnoise) {
        CUDA_CHECK(cudaMemcpyAsync(dnoise, nz_old, (size_t)n_old * 8, cudaMemcpyDeviceToDevice, s_main));
        if (noise_u) CUDA_CHECK(cudaMemcpyAsync(dnoise + n_old, noise_u, (size_t)n_u * 8, cudaMemcpyHostToDevice, s_main));
        else CUDA_CHECK(cudaMemsetAsync(dnoise + n_old, 0, (size_t)n_u * 8, s_main));
      }
      for (int k = 0; k < d; ++k) {
        const double* c = X_u + (size_t)k * n_u;
        xmin[k] = std::min(xmin[k], *std::min_element(c, c + n_u));
        xmax[k] = std::max(xmax[k], *std::max_element(c, c + n_u));
      }
      keep_n = 0;
      // (a factor whose kept and appended rows were accepted with different jitter is not kept a second time:
      //  the next evaluation factors from scratch, the reference's own fallback)
      if (keep && !last_mixed_jitter) {
        const int pb = n_old / BLK;
        const size_t c0 = (size_t)pb * BLK;
        if (c0 > 0) {
          CUDA_CHECK(cudaMemcpy2DAsync(A, (size_t)ld * 8, A_old, (size_t)ld_old * 8, c0 * 8, c0, cudaMemcpyDeviceToDevice, s_main));
          CUDA_CHECK(cudaMemcpy2DAsync(W, (size_t)ld * 8, W_old, (size_t)ld_old * 8, c0 * 8, c0, cudaMemcpyDeviceToDevice, s_main));
          CUDA_CHECK(cudaMemcpyAsync(logdet_blocks, ldb_old, (size_t)pb * 8, cudaMemcpyDeviceToDevice, s_main));
        }
        keep_n = n_old;
        keep_theta = last_theta;
        keep_alpha = last_alpha;
        keep_inv_sigma2 = last_inv_sigma2;
        keep_diag_add = last_diag_add;
      }
      CUDA_CHECK(cudaStreamSynchronize(s_main));
    } catch (...) {
      free_old();
      throw;
    }
    free_old();
    have_model = have_W = have_V = have_x = have_loo = false;
    evals_done = 0;
  }

  void set_data(const double* X, const double* y, const double* F, const double* noise) {
    CUDA_CHECK(cudaMemcpyAsync(dX, X, (size_t)n * d * 8, cudaMemcpyHostToDevice, s_main));
    CUDA_CHECK(cudaMemcpyAsync(dy, y, (size_t)n * 8, cudaMemcpyHostToDevice, s_main));
    if (p > 0) CUDA_CHECK(cudaMemcpyAsync(dF, F, (size_t)n * p * 8, cudaMemcpyHostToDevice, s_main));
    if (noise && dnoise) CUDA_CHECK(cudaMemcpyAsync(dnoise, noise, (size_t)n * 8, cudaMemcpyHostToDevice, s_main));
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    have_model = have_W = have_V = have_x = false;
    evals_done = 0;
    keep_n = 0;
    cm.valid = false;
    live_is_committed = false;
    xmin.assign(d, 0.0);
    xmax.assign(d, 0.0);
    for (int k = 0; k < d; ++k) {
      const double* c = X + (size_t)k * n;
      xmin[k] = *std::min_element(c, c + n);
      xmax[k] = *std::max_element(c, c + n);
    }
  }

  // Order of a tile table for the persistent grid (CTA b takes tiles b, b + G, b + 2G, ...: consecutive G tiles run
  // together).  Tiles are grouped into super-tiles of RB x CB output tiles; the tiles of a super-tile share RB
  // M-side and CB N-side operand strips and walk k at the same pace, so a k-window of those strips stays in the
  // 126 MB L2 instead of every tile streaming its own N-side strip from HBM.  Measured (LAUUM, n = 20000): DRAM reads
  // 150 -> 108 GB per launch at unchanged 77.5 ms (the kernel is DMMA-bound at 96 % pipe utilisation either way; the
  // CTAs drift apart, so the reuse is far from the 16 x a lock-step walk would give).  Super-tiles are issued
  // longest-k first (tail).  Pure reordering: every tile computes what it computed before, bit for bit.
  static void order_for_l2(std::vector<TileDesc>& t, int RB = 16, int CB = 16) {
    struct Group {
      long long key;
      int len;
      std::vector<TileDesc> tiles;
    };
    std::vector<Group> groups;
    std::vector<std::pair<long long, size_t>> index;  // key -> position in groups (kept sorted)
    for (const TileDesc& d_ : t) {
      const long long key = (long long)(d_.c_row / (RB * TM)) * 1000003LL + d_.c_col / (CB * TN);
      auto it = std::lower_bound(index.begin(), index.end(), std::make_pair(key, (size_t)0));
      if (it == index.end() || it->first != key) {
        it = index.insert(it, {key, groups.size()});
        groups.push_back({key, 0, {}});
      }
      Group& g = groups[it->second];
      g.len = std::max(g.len, d_.k_end - d_.k_begin);
      g.tiles.push_back(d_);
    }
    std::stable_sort(groups.begin(), groups.end(), [](const Group& x, const Group& y) {
      return x.len != y.len ? x.len > y.len : x.key < y.key;
    });
    t.clear();
    for (const Group& g : groups) t.insert(t.end(), g.tiles.begin(), g.tiles.end());
  }

  // rounds of the persistent grid (2 CTAs per SM): see tables::serpentine
  void serpentine(std::vector<TileDesc>& t) const { tables::serpentine(t, (size_t)2 * sm_count); }

  // ---- TRTRI (recursive, level-batched) and LAUUM tile tables ----
  void build_plans() {
    struct Node {
      int a, m, b, depth;
    };
    std::vector<Node> nodes;
    std::vector<Node> stack;
    int maxdepth = 0;
    {
      std::vector<Node> todo{{0, 0, nb, 0}};
      while (!todo.empty()) {
        Node nd = todo.back();
        todo.pop_back();
        if (nd.b - nd.a <= 1) continue;
        nd.m = nd.a + (nd.b - nd.a + 1) / 2;
        nodes.push_back(nd);
        maxdepth = std::max(maxdepth, nd.depth);
        todo.push_back({nd.a, 0, nd.m, nd.depth + 1});
        todo.push_back({nd.m, 0, nd.b, nd.depth + 1});
      }
    }
    std::vector<TileDesc> all;
    trtri_levels.clear();
    for (int depth = maxdepth; depth >= 0; --depth) {
      std::vector<TileDesc> t1, t2;
      for (const Node& nd : nodes) {
        if (nd.depth != depth) continue;
        for (int ct = nd.a; ct < nd.m; ++ct)
          for (int rt = 2 * nd.m; rt < 2 * nd.b; ++rt) {
            t1.push_back({rt * TM, ct * TN, ct * BLK, nd.m * BLK});
            t2.push_back({rt * TM, ct * TN, nd.m * BLK, rt * TM + TM});
          }
      }
      // (longest k first; the super-tile order of the LAUUM table was measured here too: DRAM reads 88 -> 71 GB
      //  per evaluation but 80.3 -> 82.3 ms, its coarser length order costs more in the tail than the traffic gains)
      auto by_len = [](const TileDesc& x, const TileDesc& y) { return (x.k_end - x.k_begin) > (y.k_end - y.k_begin); };
      std::stable_sort(t1.begin(), t1.end(), by_len);
      std::stable_sort(t2.begin(), t2.end(), by_len);
      // serpentine rounds: TRTRI 77.0 -> 75.6 ms at n = 20000, DRAM reads of the top level 28 + 49 -> 19 + 23 GB.  (Blocks of
      // 4 columns x 74 rows / 8 rows x 37 columns in serpentine rounds, the LAUUM recipe, were slower here: 76.7 ms --
      // the block edges do not line up with the rounds; profiles/r02c_table_orders.log)
      if (trtri_order == 1) {
        serpentine(t1);
        serpentine(t2);
      }
      TrtriLevel lv;
      lv.off1 = all.size();
      lv.n1 = t1.size();
      all.insert(all.end(), t1.begin(), t1.end());
      lv.off2 = all.size();
      lv.n2 = t2.size();
      all.insert(all.end(), t2.begin(), t2.end());
      trtri_levels.push_back(lv);
    }
    trtri_tables = dalloc<TileDesc>(all.size());
    if (!all.empty())
      CUDA_CHECK(cudaMemcpy(trtri_tables, all.data(), all.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
    std::vector<TileDesc> lt = tables::lower_tiles_by_row(nb, N, true);
    if (lauum_order == 0) {
      if (l2_order) order_for_l2(lt);
    } else if (lauum_order == 1) {
      serpentine(lt);  // (rows in ascending order = longest k first already)
    } else if (lauum_order == 3) {
      // bands of lauum_band row tiles (k ranges within lauum_band * 64 of each other), column by column inside a band:
      // a round of 2 * SMs consecutive tiles is a block of lauum_band rows x ~(2 * SMs / lauum_band) columns whose
      // tiles start together and walk k together -- every N-side strip is shared by lauum_band tiles, every M-side
      // strip by the round's columns; serpentine rounds keep the CTAs in step
      lt = tables::lower_tiles_in_bands(nb, N, lauum_band, true);
      serpentine(lt);
    }
    lauum_tiles = (int)lt.size();
    lauum_table = dalloc<TileDesc>(lt.size());
    CUDA_CHECK(cudaMemcpy(lauum_table, lt.data(), lt.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
    build_chol_plans();
  }

  // ---- Cholesky trailing updates: tile tables in L2 order ----
  // The closed-form trapezoid order (SCHED_TRAP) walks one 128-column strip of C at a time: its N-side operand strip
  // is shared by consecutive CTAs, but every column streams ALL M-side strips of the panel again, with a reuse
  // distance of the whole panel (123 MB for the first k = 768 update at n = 20000) -- ncu: L2 hit rate 64 % = exactly
  // the N-side share of the operand bytes, 6.5 GB of DRAM reads per launch, and 13 % of the warp time in the TRYWAIT
  // on the ring's full barrier (profiles/r02b_ncu_syrk.csv, gpurun_out/r02c3/prof_syrk.ncu-rep source page).
  // Here the rest update runs in horizontal bands of `chol_band` 64-row tiles, column by column inside a band: the
  // band's M-side strips (chol_band x 393 KB) stay in L2 for the whole band and every N-side strip is read once per
  // band by CTAs that run together; the look-ahead update (OB columns, all rows) runs row by row, its OB N-side
  // strips resident.  Pure reordering: every tile computes what it computed before, bit for bit.
  void build_chol_plans() {
    chol_plans.clear();
    chol_plan_OB = outer_panels > 0 ? outer_panels : (nb >= 96 ? 6 : 4);
    if (chol_band <= 0 || !l2_order) return;
    const int OB = chol_plan_OB;
    std::vector<TileDesc> all;
    for (int J0 = 0; J0 < nb; J0 += OB) {
      CholPlan pl;
      const int J1 = std::min(nb, J0 + OB);
      const int rem = nb - J1;
      const int c0 = J0 * BLK, c1 = J1 * BLK;
      if (rem > OB) {
        // look-ahead: region origin c1, 2 rem row tiles, OB column tiles, tm >= 2 tn; row-major
        pl.off_la = all.size();
        tables::chol_lookahead_tiles(all, c0, c1, rem, OB);
        pl.n_la = all.size() - pl.off_la;
        // rest: region origin c1 + OB * BLK, 2 (rem - OB) row tiles, rem - OB column tiles; bands of rows
        pl.off_rest = all.size();
        tables::chol_rest_tiles(all, c0, c1, c1 + OB * BLK, 2 * (rem - OB), rem - OB, chol_band);
        pl.n_rest = all.size() - pl.off_rest;
      }
      chol_plans.push_back(pl);
    }
    chol_tables = dalloc<TileDesc>(all.size());
    if (!all.empty())
      CUDA_CHECK(cudaMemcpy(chol_tables, all.data(), all.size() * sizeof(TileDesc), cudaMemcpyHostToDevice));
  }
  GemmArgs chol_table_args(size_t off, size_t count) {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.C = A;
    a.ldc = ld;
    a.sched = SCHED_TABLE;
    a.epilogue = EPI_SUB;
    a.table = chol_tables + off;
    a.ntiles = (int)count;
    return a;
  }

  // ---- GEMM launcher ----
  // layout: 0 = NT (both M-major), 1 = NN (M-side M-major, N-side K-major), 2 = TN (both K-major)
  void gemm(int layout, const MatMaps& mM, const double* Mbuf, const MatMaps& mN, const double* Nbuf, GemmArgs args,
            cudaStream_t st, bool persistent, int grid_cap = 0) {
    if (args.ntiles <= 0) return;
    args.sched_fence = 0;
    ++launches;
    if (debug_simple) {
      int grid = std::min(args.ntiles, 4 * sm_count);
      if (layout == 0) gemm_simple_kernel<false, false><<<grid, 256, 0, st>>>(Mbuf, ld, Nbuf, ld, args);
      else if (layout == 1) gemm_simple_kernel<false, true><<<grid, 256, 0, st>>>(Mbuf, ld, Nbuf, ld, args);
      else gemm_simple_kernel<true, true><<<grid, 256, 0, st>>>(Mbuf, ld, Nbuf, ld, args);
      CUDA_CHECK(cudaGetLastError());
      return;
    }
    int grid = (persistent && !no_persistent) ? std::min(args.ntiles, 2 * sm_count) : args.ntiles;
    if (grid_cap > 0) grid = std::min(grid, grid_cap);
    args.mm3 = (mM.mm4 != nullptr && mN.mm8 != nullptr) ? 1 : 0;
    if (layout == 0)
      gemm_dmma_kernel<false, false><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(args.mm3 ? mM.mm4 : mM.mm,
                                                                                  args.mm3 ? mN.mm8 : mN.mm, args);
    else if (layout == 1)
      gemm_dmma_kernel<false, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(args.mm3 ? mM.mm4 : mM.mm, mN.km, args);
    else
      gemm_dmma_kernel<true, true><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(mM.km, mN.km, args);
    CUDA_CHECK(cudaGetLastError());
  }

  GemmArgs rect_args(double* C, int epi, int row0, int col0, int mt, int nt, int k0, int k1) {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.C = C;
    a.ldc = ld;
    a.sched = SCHED_RECT;
    a.epilogue = epi;
    a.row0 = row0;
    a.col0 = col0;
    a.mt = mt;
    a.nt = nt;
    a.k_begin = k0;
    a.k_end = k1;
    a.ntiles = mt * nt;
    return a;
  }

  // ---- covariance build (a2) ----
  // Full lower triangle (t0 = 0), or -- for the block extension of a kept factor -- the 64-row tiles >= t0 only:
  // over all their columns (lower_right_only = false) or over the columns >= 64 t0 only (true).  Rows below
  // diag_split get diag_add_lo on the diagonal (the jitter of the kept factor), the others diag_add.
  void cov_build(double* dst, double alpha, double inv_sigma2, double diag_add, cudaStream_t st, int t0 = 0,
                 bool lower_right_only = false, int diag_split = 0, double diag_add_lo = 0.0) {
    const int t = N / PT;
    int ntiles, id0, shift;
    if (lower_right_only) {
      const int r = t - t0;
      ntiles = r * (r + 1) / 2;
      id0 = 0;
      shift = t0;
    } else {
      id0 = t0 * (t0 + 1) / 2;
      ntiles = t * (t + 1) / 2 - id0;
      shift = 0;
    }
    if (ntiles <= 0) return;
    const int grid = std::min(ntiles, 8 * sm_count);
    const size_t smem = (size_t)2 * d * PT * 8;
    const double* nz = (noise_model == LKGPU_NOISE_HETERO) ? dnoise : nullptr;
    ++launches;
#define LK_COV(K)                                                                                                     \
  cov_build_kernel<K><<<grid, PAIR_THREADS, smem, st>>>(dX, n, d, kp, alpha, nz, inv_sigma2, diag_add, dst, ld, ntiles, \
                                                        id0, shift, diag_split, diag_add_lo)
    switch (kernel) {
      case 0: LK_COV(0); break;
      case 1: LK_COV(1); break;
      case 2: LK_COV(2); break;
      default: LK_COV(3); break;
    }
#undef LK_COV
    CUDA_CHECK(cudaGetLastError());
  }

  // ---- two-level blocked right-looking Cholesky with look-ahead (a3) ----
  // Outer blocks of `outer_panels` 128-column panels.  Inside an outer block every panel is factored (POTF2 +
  // inverse), its column is solved (TRSM as a GEMM with the inverse) and only the REST OF THE OUTER BLOCK is
  // updated (k = 128).  The trailing matrix is then updated once per outer block with k = 128 * outer_panels:
  // 1/outer_panels of the C read-modify-write traffic and main loops long enough to fill the DMMA pipe.
  // Look-ahead: the next outer block's columns are updated first on the high-priority stream, which then factors
  // them while the low-priority stream updates the rest of the trailing matrix.
  GemmArgs trap_args(int row0, int mt, int ncols, int k0, int k1) {
    GemmArgs a = rect_args(A, EPI_SUB, row0, row0, mt, ncols, k0, k1);
    a.sched = SCHED_TRAP;
    a.ntiles = ncols * mt - ncols * (ncols - 1);
    return a;
  }
  // Every launch of an attempt carries dinfo as its abort flag: after the first non-positive pivot the remaining
  // panels and updates are no-ops (the attempt is discarded by the ladder either way).
  GemmArgs chol_gemm(GemmArgs a) {
    a.abort_flag = use_abort ? dinfo : nullptr;
    return a;
  }
  // Jstart > 0: the panels < Jstart already hold L (kept factor + row-block solve) and the trailing block
  // [Jstart.., Jstart..] holds the Schur complement: chol_block's last step (LinearAlgebra.cpp:286).
  // LKGPU_TRACE_CHOL=1 (diagnostics): timing events after every launch of the factorisation stream; the intervals
  // (completion to completion, so launch gaps and dependency waits are inside them) go to stderr after the stage.
  bool trace_chol = false;
  // 1: the rest update starts after the look-ahead update, 0: beside it, -1: by size (after it for nb < 96: n = 5000
  // chol 4.46 -> 4.31 ms; beside it above: n = 20000 83.9 against 84.9 ms).  LKGPU_UPDATE_AFTER_LOOKAHEAD overrides.
  int update_after_lookahead = -1;
  std::vector<std::pair<cudaEvent_t, std::string>> trace_marks;
  void trace_mark(const char* what, int j) {
    if (!trace_chol) return;
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    CUDA_CHECK(cudaEventRecord(e, s_main));
    trace_marks.push_back({e, std::string(what) + " " + std::to_string(j)});
  }
  void trace_dump() {
    if (!trace_chol || trace_marks.empty()) return;
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    for (size_t i = 1; i < trace_marks.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, trace_marks[i - 1].first, trace_marks[i].first);
      fprintf(stderr, "[trace] %-14s %9.2f us\n", trace_marks[i].second.c_str(), 1e3 * ms);
    }
    for (auto& m : trace_marks) cudaEventDestroy(m.first);
    trace_marks.clear();
  }
  void cholesky(int Jstart = 0) {
    dev_zero_ints(dinfo, 4);
    trace_mark("start", Jstart);
    // measured at n = 20000 (nb = 157): 3 panels 97.9 ms, 4: 95.3, 5: 94.3, 6: 93.9, 8: 93.8; mid-size matrices keep 4
    const int OB = outer_panels > 0 ? outer_panels : (nb >= 96 ? 6 : 4);
    const bool planned = Jstart == 0 && OB == chol_plan_OB && !chol_plans.empty() && persistent_update_reserve == 0;
    int last_upd = -1;
    for (int J0 = Jstart; J0 < nb; J0 += OB) {
      const int J1 = std::min(nb, J0 + OB);
      const int c0 = J0 * BLK, c1 = J1 * BLK;
      for (int j = J0; j < J1; ++j) {
        const int jb = j * BLK;
        ++launches;
        potf2_inv_kernel<<<1, POTF2_THREADS, POTF2_SMEM_BYTES, s_main>>>(A, W, ld, jb, logdet_blocks, j, dinfo, use_abort ? 1 : 0);
        CUDA_CHECK(cudaGetLastError());
        trace_mark("potf2", j);
        const int rem = nb - j - 1;
        if (rem == 0) break;
        // panel TRSM as GEMM with the inverted diagonal block: A[i, j] <- A[i, j] * Dinv_j^T  (in place)
        gemm(0, mapA, A, mapW, W, chol_gemm(rect_args(A, EPI_SET, jb + BLK, jb, 2 * rem, 1, jb, jb + BLK)), s_main, false);
        trace_mark("trsm", j);
        const int ncol = J1 - (j + 1);  // panels of this outer block still to be factored
        if (ncol > 0) gemm(0, mapA, A, mapA, A, chol_gemm(trap_args(jb + BLK, 2 * rem, ncol, jb, jb + BLK)), s_main, false);
        if (ncol > 0) trace_mark("inblock", j);
      }
      const int rem = nb - J1;  // panels after this outer block
      if (rem <= 0) break;
      if (use_lookahead && rem > OB) {
        CUDA_CHECK(cudaEventRecord(ev_panel[J0], s_main));
        if (last_upd >= 0) CUDA_CHECK(cudaStreamWaitEvent(s_main, ev_upd[last_upd], 0));
        // look-ahead: the next outer block's columns first, on the factorisation stream ...
        const CholPlan* pl = planned ? &chol_plans[J0 / OB] : nullptr;
        if (pl) gemm(0, mapA, A, mapA, A, chol_gemm(chol_table_args(pl->off_la, pl->n_la)), s_main, false);
        else gemm(0, mapA, A, mapA, A, chol_gemm(trap_args(c1, 2 * rem, OB, c0, c1)), s_main, false);
        trace_mark("lookahead", J0);
        // ... the rest of the trailing matrix on the low-priority stream.  Mid-size matrices start it when the look-ahead
        // update has finished, together with the next block's first panel kernel: that kernel needs a whole SM (231 KB of
        // shared memory), and if the rest update -- started beside the look-ahead update -- already fills the machine, it
        // waits for an SM to drain (LKGPU_TRACE_CHOL: 127 instead of 59 us at n = 5000, once per outer block).
        if (update_after_lookahead > 0 || (update_after_lookahead < 0 && nb < 96))
          CUDA_CHECK(cudaEventRecord(ev_panel[J0], s_main));
        CUDA_CHECK(cudaStreamWaitEvent(s_upd, ev_panel[J0], 0));
        // (experiment for round 2, off by default: LKGPU_PERSISTENT_UPDATE=r runs this update as a persistent grid on
        //  2 (SMs - r) CTAs -- the run-ahead producer then hides each tile's prologue -- leaving r SMs to the panel chain)
        if (persistent_update_reserve > 0)
          gemm(0, mapA, A, mapA, A, chol_gemm(trap_args(c1 + OB * BLK, 2 * (rem - OB), rem - OB, c0, c1)), s_upd, true,
               2 * std::max(1, sm_count - persistent_update_reserve));
        else if (pl)
          gemm(0, mapA, A, mapA, A, chol_gemm(chol_table_args(pl->off_rest, pl->n_rest)), s_upd, false);
        else
          gemm(0, mapA, A, mapA, A, chol_gemm(trap_args(c1 + OB * BLK, 2 * (rem - OB), rem - OB, c0, c1)), s_upd, false);
        CUDA_CHECK(cudaEventRecord(ev_upd[J0], s_upd));
        last_upd = J0;
      } else {
        if (last_upd >= 0) {
          CUDA_CHECK(cudaStreamWaitEvent(s_main, ev_upd[last_upd], 0));
          last_upd = -1;
        }
        gemm(0, mapA, A, mapA, A, chol_gemm(trap_args(c1, 2 * rem, rem, c0, c1)), s_main, false);
        trace_mark("update", J0);
      }
    }
    if (last_upd >= 0) CUDA_CHECK(cudaStreamWaitEvent(s_main, ev_upd[last_upd], 0));
    trace_dump();
  }

  // ---- block extension of the kept factor (f3) ----------------------------------------------------------------
  // LinearAlgebra::update_cholCov + chol_block (LinearAlgebra.cpp:206-299) for populate_Model's update_eligible
  // case: with o = kept rows, u = appended rows,
  //     L_oo = kept factor,   L_uo = C_uo L_oo^-T,   L_uu = safe_chol_lower(C_uu - L_uo L_uo^T),
  // the jitter ladder and the rcond test acting on the Schur complement only; if that ladder is exhausted the
  // reference falls back to a from-scratch safe_chol_lower(C) (:287-293) -- this returns false and eval() does that.
  // On the device the split is moved down to the last whole 128-panel boundary c0 <= o: rows [c0, o) of the kept
  // factor are re-derived (their diagonal carries the jitter the kept factor was accepted with), which leaves
  // L_oo unchanged up to rounding and lets every step run on whole tiles:
  //   1. covariance rows >= c0 (cov_build, partial)
  //   2. row-block solve against the kept panels, right-looking so that every step fills the machine even when few
  //      rows are appended:  A[c0.., k] <- A[c0.., k] Dinv_k^T ;  A[c0.., k+1..] -= A[c0.., k] A[k+1.., k]^T
  //   3. Schur complement (one SYRK, k = c0)      4. cholesky(Jstart = c0 / 128).
  bool factor_update(double alpha, double inv_sigma2, lkgpu_out* out, float& ms_cov, float& ms_chol, float& ms_rcond,
                     double& rc2_out, int& inc_out, double& diag_out) {
    const int no = keep_n;
    const int pb = no / BLK, c0 = pb * BLK, t0 = c0 / PT;
    const int mrows = 2 * (nb - pb);  // 64-row tiles below c0
    float t;
    CUDA_CHECK(cudaEventRecord(ev_t[1], s_main));
    cov_build(A, alpha, inv_sigma2, 0.0, s_main, t0, false, no, keep_diag_add);
    CUDA_CHECK(cudaEventRecord(ev_t[2], s_main));
    for (int k = 0; k < pb; ++k) {
      const int kb = k * BLK;
      gemm(0, mapA, A, mapW, W, rect_args(A, EPI_SET, c0, kb, mrows, 1, kb, kb + BLK), s_main, false);
      if (k + 1 < pb)
        gemm(0, mapA, A, mapA, A, rect_args(A, EPI_SUB, c0, kb + BLK, mrows, pb - k - 1, kb, kb + BLK), s_main, false);
    }
    CUDA_CHECK(cudaEventRecord(ev_t[3], s_main));
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    cudaEventElapsedTime(&t, ev_t[1], ev_t[2]); ms_cov += t;
    cudaEventElapsedTime(&t, ev_t[2], ev_t[3]); ms_chol += t;
    double diag_add = 0.0;
    int inc = 0;
    const int nu = n - no;
    while (true) {
      CUDA_CHECK(cudaEventRecord(ev_t[1], s_main));
      if (inc > 0) cov_build(A, alpha, inv_sigma2, diag_add, s_main, t0, true, no, keep_diag_add);
      CUDA_CHECK(cudaEventRecord(ev_t[2], s_main));
      if (pb > 0) gemm(0, mapA, A, mapA, A, trap_args(c0, mrows, nb - pb, 0, c0), s_main, false);
      cholesky(pb);
      CUDA_CHECK(cudaEventRecord(ev_t[3], s_main));
      // info + ||L_uu||_1
      const double* Luu = A + (long long)no * ld + no;
      launches += 2;
      tri_colsum_abs_kernel<<<(nu * 32 + 255) / 256, 256, 0, s_main>>>(Luu, ld, nu, dcolsum);
      vec_max_kernel<<<1, 256, 0, s_main>>>(dcolsum, nu, dscal + SC_NORML);
      CUDA_CHECK(cudaMemcpyAsync(hpin, dscal + SC_NORML, 8, cudaMemcpyDeviceToHost, s_main));
      CUDA_CHECK(cudaMemcpyAsync(hpin + 2, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s_main));
      CUDA_CHECK(cudaStreamSynchronize(s_main));
      cudaEventElapsedTime(&t, ev_t[1], ev_t[2]); ms_cov += t;
      cudaEventElapsedTime(&t, ev_t[2], ev_t[3]); ms_chol += t;
      const double normL = hpin[0];
      int info;
      memcpy(&info, hpin + 2, sizeof(int));
      const bool ok = (info == 0) && std::isfinite(normL);
      bool wrong_rcond = rcond_check;
      double rc2 = NAN;
      if (ok && rcond_check) {
        CUDA_CHECK(cudaEventRecord(ev_t[5], s_main));
        const double rc = rcond_estimate(normL, no);  // dtrcon on L_uu
        rc2 = rc * rc;
        wrong_rcond = rc2 < min_rcond;
        CUDA_CHECK(cudaEventRecord(ev_t[6], s_main));
        CUDA_CHECK(cudaStreamSynchronize(s_main));
        cudaEventElapsedTime(&t, ev_t[5], ev_t[6]); ms_rcond += t;
      }
      if (!ok || wrong_rcond) {
        if (inc > max_inc || num_nugget <= 0.0) return false;  // chol_block's catch: factor C from scratch
        diag_add += num_nugget * std::pow(10.0, inc);
        ++inc;
        continue;
      }
      rc2_out = rc2;
      inc_out = inc;
      diag_out = diag_add;
      return true;
    }
  }

  // ---- TRTRI: W <- L^-1 (lower), V used as scratch (a4) ----
  void trtri() {
    for (const TrtriLevel& lv : trtri_levels) {
      GemmArgs a;
      memset(&a, 0, sizeof(a));
      a.ldc = ld;
      a.sched = SCHED_TABLE;
      // X = L21 * W11 -> V
      a.C = V;
      a.epilogue = EPI_SET;
      a.table = trtri_tables + lv.off1;
      a.ntiles = (int)lv.n1;
      gemm(1, mapA, A, mapW, W, a, s_main, true);
      // W21 = -W22 * X
      a.C = W;
      a.epilogue = EPI_SETNEG;
      a.table = trtri_tables + lv.off2;
      a.ntiles = (int)lv.n2;
      gemm(1, mapW, W, mapV, V, a, s_main, true);
    }
    have_W = true;
  }

  // ---- LAUUM: V <- W^T W (lower tiles) = R^-1 (a4) ----
  void lauum() {
    GemmArgs a;
    memset(&a, 0, sizeof(a));
    a.C = V;
    a.ldc = ld;
    a.sched = SCHED_TABLE;
    a.epilogue = EPI_SET;
    a.table = lauum_table;
    a.ntiles = lauum_tiles;
    gemm(2, mapW, W, mapW, W, a, s_main, true);
    have_V = true;
  }

  // ---- triangular sweeps (a5) ----
  template <bool BWD>
  void solve_wave(double* B, int nrhs) {
    int grid = std::min(nb, sm_count);
    if (wave_grid_cap > 0) grid = std::min(grid, wave_grid_cap);
    for (int q0 = 0; q0 < nrhs; q0 += WAVE_MAX_RHS) {
      const int nq = std::min(WAVE_MAX_RHS, nrhs - q0);
      double* Bq = B + (long long)q0 * N;
      launches += 2;
      wave_reset_kernel<<<(unsigned)(((long long)N * nq + 255) / 256), 256, 0, s_main>>>(wave_ctl, wave_z, (long long)N * nq);
      const CUtensorMap* mL = BWD ? mapA.wb : mapA.wf;
      const CUtensorMap* mW = BWD ? mapW.wb : mapW.wf;
      if (nq == 1)
        trsv_wave_kernel<BWD, 1><<<grid, WAVE_THREADS, wave_smem_bytes(1), s_main>>>(mL, mW, Bq, N, nq, nb, wave_ctl, wave_z, 0);
      else if (nq == 2)
        trsv_wave_kernel<BWD, 2><<<grid, WAVE_THREADS, wave_smem_bytes(2), s_main>>>(mL, mW, Bq, N, nq, nb, wave_ctl, wave_z, 0);
      else if (nq <= 4)
        trsv_wave_kernel<BWD, 4><<<grid, WAVE_THREADS, wave_smem_bytes(4), s_main>>>(mL, mW, Bq, N, nq, nb, wave_ctl, wave_z, 0);
      else
        trsv_wave_kernel<BWD, 8><<<grid, WAVE_THREADS, wave_smem_bytes(8), s_main>>>(mL, mW, Bq, N, nq, nb, wave_ctl, wave_z, 0);
      CUDA_CHECK(cudaGetLastError());
    }
  }
  // ---- evaluations of different handles on one device ----
  // A factorisation of n <= ~8192 cannot fill 148 SMs (its panel chain is latency-bound), so the evaluations of
  // several such handles -- multistart rows (BASELINE cfg 5), NestedKriging sub-models -- overlap: every handle has
  // its own workspaces and streams and is driven by its own host thread.  Overlapping evaluations return the bits of
  // a lone handle (every kernel is deterministic and touches only its handle's buffers; validated on the device with
  // tools/diag_concurrent2.py and tools/diag_foreign.py, profiles/r02_relax_validation.md).  Policy of this gate:
  //  * handles with n <= overlap_max_n, or flagged by lkgpu_set_concurrent: shared -- they overlap with each other;
  //  * larger unflagged handles: exclusive -- each fills the GPU by itself, so they queue (their host work still
  //    overlaps) instead of thrashing each other's L2 working sets.
  bool concurrent_flag = false;  // lkgpu_set_concurrent
  int gate_depth = 0;  // the gate is re-entrant per handle (append -> restore)
  struct SweepGate {
    Engine& e;
    bool shared;
    bool outer;
    explicit SweepGate(Engine& e_) : e(e_), shared(e_.concurrent_flag || e_.n <= e_.overlap_max_n), outer(e_.gate_depth++ == 0) {
      if (!outer) return;
      DeviceGate& g = g_gate[e.device & 63];
      std::unique_lock<std::mutex> lk(g.m);
      if (shared) {
        g.cv.wait(lk, [&] { return !g.exclusive_active && g.exclusive_waiting == 0; });
        ++g.shared_active;
      } else {
        ++g.exclusive_waiting;
        g.cv.wait(lk, [&] { return !g.exclusive_active && g.shared_active == 0; });
        --g.exclusive_waiting;
        g.exclusive_active = true;
      }
    }
    ~SweepGate() {
      --e.gate_depth;
      if (!outer) return;
      DeviceGate& g = g_gate[e.device & 63];
      {
        std::lock_guard<std::mutex> lk(g.m);
        if (shared) --g.shared_active;
        else g.exclusive_active = false;
      }
      g.cv.notify_all();
    }
  };
  bool sweeps_by_launch_chain() const { return use_step_trsv; }
  void solve_fwd(double* B, int nrhs) {
    if (sweeps_by_launch_chain()) return solve_fwd_steps(B, nrhs);
    solve_wave<false>(B, nrhs);
  }
  void solve_bwd(double* B, int nrhs) {
    if (sweeps_by_launch_chain()) return solve_bwd_steps(B, nrhs);
    solve_wave<true>(B, nrhs);
  }
  // launch-chain variant (one kernel per 128-row step), kept for fault localisation: LKGPU_STEP_TRSV=1
  void solve_fwd_steps(double* B, int nrhs) {
    for (int q0 = 0; q0 < nrhs; q0 += TRSV_MAX_RHS) {
      const int nq = std::min(TRSV_MAX_RHS, nrhs - q0);
      double* Bq = B + (long long)q0 * N;
      for (int j = -1; j < nb - 1; ++j) {
        ++launches;
        trsv_fwd_step_kernel<<<nb - 1 - j < 1 ? 1 : (j < 0 ? 1 : nb - 1 - j), 128, 0, s_main>>>(A, W, ld, Bq, N, nq, j);
      }
      CUDA_CHECK(cudaGetLastError());
    }
  }
  void solve_bwd_steps(double* B, int nrhs) {
    for (int q0 = 0; q0 < nrhs; q0 += TRSV_MAX_RHS) {
      const int nq = std::min(TRSV_MAX_RHS, nrhs - q0);
      double* Bq = B + (long long)q0 * N;
      for (int j = nb; j >= 1; --j) {
        ++launches;
        trsv_bwd_step_kernel<<<(j >= nb) ? 1 : j, 128, 0, s_main>>>(A, W, ld, Bq, N, nq, j, nb);
      }
      CUDA_CHECK(cudaGetLastError());
    }
  }

  // dtrcon('1','L','N') restated: Higham/Hager estimator dlacn2 driven from the host,
  // triangular solves on the device (reference: arma::rcond -> dtrcon, auxlib_meat.hpp:6777-6800).
  // r0 > 0: the estimate is for the trailing block L[r0.., r0..] (chol_block's L_uu): its solves are sweeps of
  // the whole factor on right-hand sides that are zero above r0, read back below r0.
  double rcond_estimate(double anorm, int r0 = 0) {
    if (!(anorm > 0.0)) return 0.0;
    const int itmax = 5;
    const int n = this->n - r0;  // order of the block (shadows the member on purpose)
    std::vector<double> x(n), v(n);
    std::vector<int> isgn(n);
    auto dev_solve = [&](bool transpose) {
      // rows above r0 and the padded tail are zero
      for (int i = 0; i < N; ++i) hpin[i] = 0.0;
      memcpy(hpin + r0, x.data(), (size_t)n * 8);
      CUDA_CHECK(cudaMemcpyAsync(Tv, hpin, (size_t)N * 8, cudaMemcpyHostToDevice, s_main));
      if (transpose) solve_bwd(Tv, 1);
      else solve_fwd(Tv, 1);
      CUDA_CHECK(cudaMemcpyAsync(hpin, Tv, (size_t)N * 8, cudaMemcpyDeviceToHost, s_main));
      CUDA_CHECK(cudaStreamSynchronize(s_main));
      memcpy(x.data(), hpin + r0, (size_t)n * 8);
    };
    auto dasum = [&](const std::vector<double>& a) {
      double s = 0;
      for (double t : a) s += std::fabs(t);
      return s;
    };
    auto idamax = [&](const std::vector<double>& a) {
      int im = 0;
      double m = std::fabs(a[0]);
      for (int i = 1; i < n; ++i)
        if (std::fabs(a[i]) > m) {
          m = std::fabs(a[i]);
          im = i;
        }
      return im;
    };
    double est = 0.0;
    for (int i = 0; i < n; ++i) x[i] = 1.0 / n;
    dev_solve(false);  // kase = 1: x <- inv(A) x
    if (n == 1) {
      v = x;
      est = std::fabs(v[0]);
      return (1.0 / anorm) / est;
    }
    est = dasum(x);
    for (int i = 0; i < n; ++i) {
      x[i] = (x[i] >= 0.0) ? 1.0 : -1.0;
      isgn[i] = (int)x[i];
    }
    dev_solve(true);  // kase = 2: x <- inv(A)^T x
    int j = idamax(x);
    int iter = 2;
    bool alt = false;
    while (true) {
      std::fill(x.begin(), x.end(), 0.0);
      x[j] = 1.0;
      dev_solve(false);
      v = x;
      const double estold = est;
      est = dasum(v);
      bool same = true;
      for (int i = 0; i < n; ++i) {
        const int s = (x[i] >= 0.0) ? 1 : -1;
        if (s != isgn[i]) {
          same = false;
          break;
        }
      }
      if (same || est <= estold) {
        alt = true;
        break;
      }
      for (int i = 0; i < n; ++i) {
        x[i] = (x[i] >= 0.0) ? 1.0 : -1.0;
        isgn[i] = (int)x[i];
      }
      dev_solve(true);
      const int jlast = j;
      j = idamax(x);
      if (x[jlast] != std::fabs(x[j]) && iter < itmax) {
        ++iter;
        continue;
      }
      alt = true;
      break;
    }
    if (alt) {
      double altsgn = 1.0;
      for (int i = 0; i < n; ++i) {
        x[i] = altsgn * (1.0 + (double)i / (double)(n - 1));
        altsgn = -altsgn;
      }
      dev_solve(false);
      const double temp = 2.0 * (dasum(x) / (double)(3 * n));
      if (temp > est) est = temp;
    }
    if (est == 0.0) return 0.0;
    return (1.0 / anorm) / est;
  }

  void dev_zero(double* p_, long long count) {
    if (count <= 0) return;
    ++launches;
    fill_zero_kernel<<<(unsigned)std::min<long long>((count + 255) / 256, 8LL * sm_count), 256, 0, s_main>>>(p_, count);
    CUDA_CHECK(cudaGetLastError());
  }
  void dev_copy(double* dst, const double* src, long long count) {
    if (count <= 0) return;
    ++launches;
    copy_kernel<<<(unsigned)std::min<long long>((count + 255) / 256, 8LL * sm_count), 256, 0, s_main>>>(dst, src, count);
    CUDA_CHECK(cudaGetLastError());
  }
  void dev_zero_ints(int* p_, int count) {
    ++launches;
    zero_ints_kernel<<<(count + 255) / 256, 256, 0, s_main>>>(p_, count);
    CUDA_CHECK(cudaGetLastError());
  }

  // ---- ladder shortcut: set the accepted factor aside while a lower rung is tried (Engine::eval) ----
  // The factor lives in A (L), W's diagonal blocks (their inverses) and logdet_blocks.  A and V swap roles (V is
  // scratch until TRTRI), the two small pieces are copied.
  double *ladder_D = nullptr, *ladder_logdet = nullptr;
  void stash_factor() {
    if (!ladder_D) {
      ladder_D = dalloc<double>((size_t)nb * BLK * BLK);
      ladder_logdet = dalloc<double>(nb);
    }
    copy_diag_blocks(ladder_D, BLK, (long long)BLK * BLK, W, ld, (long long)BLK * ld + BLK);
    dev_copy(ladder_logdet, logdet_blocks, nb);
    std::swap(A, V);
    std::swap(mapA, mapV);
  }
  void unstash_factor() {
    std::swap(A, V);
    std::swap(mapA, mapV);
    copy_diag_blocks(W, ld, (long long)BLK * ld + BLK, ladder_D, BLK, (long long)BLK * BLK);
    dev_copy(logdet_blocks, ladder_logdet, nb);
  }

  void tic(int idx) { CUDA_CHECK(cudaEventRecord(ev_t[idx], s_main)); }

  // deterministic sum of squares of a vector (n rows) into dscal[slot]
  void sum_sq(const double* vec, int slot) {
    const int chunks = (n + GRAM_CHUNK - 1) / GRAM_CHUNK;
    launches += 2;
    gram_partial_kernel<<<chunks, 256, 0, s_main>>>(vec, N, n, 1, dpartial);
    sum_partials_kernel<<<1, 32, 0, s_main>>>(dpartial, chunks, 1, 1, dscal + slot);
    CUDA_CHECK(cudaGetLastError());
  }

  // pair reduction (K8).  Wmat: weight matrix (lower tiles); (avec, bvec): rank-1 term 1/2 (a_i b_j + b_i a_j)
  void grad_reduce(double alpha, int pdim, int out_slot, const double* Wmat, const double* avec, const double* bvec) {
    const int t = N / PT;
    const int ntiles = t * (t + 1) / 2;
    const int grid = std::min(ntiles, 2 * sm_count);
    const size_t smem = ((size_t)2 * d * PT + 4 * PT + 16 * (d + 1) + (size_t)2 * pdim * PT) * 8;
    if (smem > 200 * 1024) throw LkError{"lkgpu: d / p too large for the pair-reduction kernel's shared memory"};
    ++launches;
#define LAUNCH_GRAD(K)                                                                                              \
  do {                                                                                                              \
    if (smem > 48 * 1024)                                                                                           \
      CUDA_CHECK(cudaFuncSetAttribute(grad_reduce_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    grad_reduce_kernel<K><<<grid, PAIR_THREADS, smem, s_main>>>(dX, n, d, kp, alpha, Wmat, ld, avec, bvec, Uv, N, pdim, \
                                                                  dpartial, ntiles);                                 \
  } while (0)
    switch (kernel) {
      case 0: LAUNCH_GRAD(0); break;
      case 1: LAUNCH_GRAD(1); break;
      case 2: LAUNCH_GRAD(2); break;
      default: LAUNCH_GRAD(3); break;
    }
#undef LAUNCH_GRAD
    CUDA_CHECK(cudaGetLastError());
    ++launches;
    const int cnt = 2 * (d + 1);
    sum_partials_kernel<<<(cnt + 63) / 64, 64, 0, s_main>>>(dpartial, grid, cnt, cnt, dscal + out_slot);
    CUDA_CHECK(cudaGetLastError());
  }

  // =====================  one objective evaluation  =====================
  void eval(int objective, const double* theta, double extra, int want_grad, lkgpu_out* out) {
    CUDA_CHECK(cudaSetDevice(device));
    SweepGate gate(*this);
    if (objective != LKGPU_OBJ_LL && objective != LKGPU_OBJ_LOO && objective != LKGPU_OBJ_LMP)
      throw LkError{"lkgpu_eval: unknown objective"};
    if (objective == LKGPU_OBJ_LOO && noise_model != LKGPU_NOISE_NONE)
      throw LkError{"LOO objective not supported for Nugget/Heterogeneous noise modes"};  // Kriging.cpp:1472
    if (objective == LKGPU_OBJ_LMP && noise_model == LKGPU_NOISE_HETERO)
      throw LkError{"LMP objective not supported for Heterogeneous noise mode"};  // Kriging.cpp:1488
    for (int k = 0; k < d; ++k) {
      if (!(theta[k] > 0.0)) throw LkError{"lkgpu_eval: theta must be positive"};
      kp.inv_theta[k] = 1.0 / theta[k];
    }
    last_theta.assign(theta, theta + d);
    double alpha = 1.0, inv_sigma2 = 0.0;
    if (noise_model == LKGPU_NOISE_NUGGET) alpha = extra;
    if (noise_model == LKGPU_NOISE_HETERO) inv_sigma2 = 1.0 / extra;
    last_alpha = alpha;
    last_inv_sigma2 = inv_sigma2;
    have_model = have_W = have_V = have_x = have_loo = false;
    live_is_committed = false;
    memset(out->stage_ms, 0, sizeof(out->stage_ms));
    const bool need_inverse = want_grad || objective == LKGPU_OBJ_LOO;

    tic(0);
    double diag_add = 0.0;
    int inc = 0;
    double rc2 = 0.0;
    float ms_cov = 0, ms_chol = 0, ms_rcond = 0, ms_trtri = 0;
    int n_attempt_info = 0, n_attempt_rcond = 0;  // rejected rungs: failed factorisation / rcond below min_rcond
    int n_rungs_skipped = 0;                      // rungs the ladder shortcut did not have to factor
    // ---- populate_Model's update_eligible (Kriging.cpp:170-188): same theta / extra as the kept factor and more
    //      rows than it has -> block extension; anything else (or an exhausted ladder there) -> from scratch ----
    bool updated = false;
    last_was_update = last_mixed_jitter = false;
    if (keep_n > 0) {
      const bool eligible = keep_n < n && (int)keep_theta.size() == d && std::equal(theta, theta + d, keep_theta.begin()) &&
                            alpha == keep_alpha && inv_sigma2 == keep_inv_sigma2;
      if (eligible) updated = factor_update(alpha, inv_sigma2, out, ms_cov, ms_chol, ms_rcond, rc2, inc, diag_add);
      last_was_update = updated;
      last_mixed_jitter = updated && diag_add != keep_diag_add;
      keep_n = 0;  // the kept factor is consumed (or overwritten) by this evaluation
    }
    // ---- safe_chol_lower (LinearAlgebra.cpp:43-98): jitter ladder driven from the host ----
    // Rung r = r cumulative bumps on the diagonal: diag_add(r) = sum_{i<r} num_nugget 10^i, summed in the ladder's
    // own order so that a rung reached directly carries bit for bit the jitter it has when climbed to.
    auto ladder_diag = [&](int r) {
      double s_ = 0.0;
      for (int i = 0; i < r; ++i) s_ += num_nugget * std::pow(10.0, i);
      return s_;
    };
    // One attempt: R + diag_add(r) I -> L; accepted iff the factorisation succeeds and rcond_1(L)^2 >= min_rcond.
    // exact_first: gradient path on the first rung of a handle that has not been climbing -- L^-1 is needed anyway
    // and its exact 1-norm is a lower bound of what dtrcon estimates, so accepting on it is exactly the reference's
    // decision; otherwise (or when the exact value says "reject") the estimator decides, as in the reference.
    auto attempt = [&](int r, bool exact_first) -> bool {
      have_W = false;
      diag_add = ladder_diag(r);
      CUDA_CHECK(cudaEventRecord(ev_t[1], s_main));
      cov_build(A, alpha, inv_sigma2, diag_add, s_main);
      CUDA_CHECK(cudaEventRecord(ev_t[2], s_main));
      cholesky();
      CUDA_CHECK(cudaEventRecord(ev_t[3], s_main));
      // info + ||L||_1 : a failed factorisation goes straight to the next rung of the ladder
      launches += 2;
      tri_colsum_abs_kernel<<<(n * 32 + 255) / 256, 256, 0, s_main>>>(A, ld, n, dcolsum);
      vec_max_kernel<<<1, 256, 0, s_main>>>(dcolsum, n, dscal + SC_NORML);
      CUDA_CHECK(cudaMemcpyAsync(hpin, dscal + SC_NORML, 8, cudaMemcpyDeviceToHost, s_main));
      CUDA_CHECK(cudaMemcpyAsync(hpin + 2, dinfo, sizeof(int), cudaMemcpyDeviceToHost, s_main));
      CUDA_CHECK(cudaStreamSynchronize(s_main));
      float t;
      cudaEventElapsedTime(&t, ev_t[1], ev_t[2]); ms_cov += t;
      cudaEventElapsedTime(&t, ev_t[2], ev_t[3]); ms_chol += t;
      const double normL = hpin[0];
      int info;
      memcpy(&info, hpin + 2, sizeof(int));
      const bool ok = (info == 0) && std::isfinite(normL);
      bool wrong_rcond = rcond_check;
      if (ok && rcond_check) {
        double rc = -1.0;
        if (exact_first) {
          CUDA_CHECK(cudaEventRecord(ev_t[3], s_main));
          trtri();
          CUDA_CHECK(cudaEventRecord(ev_t[4], s_main));
          launches += 2;
          tri_colsum_abs_kernel<<<(n * 32 + 255) / 256, 256, 0, s_main>>>(W, ld, n, dcolsum);
          vec_max_kernel<<<1, 256, 0, s_main>>>(dcolsum, n, dscal + SC_NORMW);
          CUDA_CHECK(cudaMemcpyAsync(hpin, dscal + SC_NORMW, 8, cudaMemcpyDeviceToHost, s_main));
          CUDA_CHECK(cudaStreamSynchronize(s_main));
          cudaEventElapsedTime(&t, ev_t[3], ev_t[4]); ms_trtri += t;
          const double normW = hpin[0];
          if (std::isfinite(normW) && normW > 0.0) rc = (1.0 / normL) / normW;
        }
        CUDA_CHECK(cudaEventRecord(ev_t[5], s_main));
        if (!(rc * rc >= min_rcond) || rc < 0.0) rc = rcond_estimate(normL);  // dtrcon restated (Higham / Hager)
        rc2 = rc * rc;
        wrong_rcond = rc2 < min_rcond;
        CUDA_CHECK(cudaEventRecord(ev_t[6], s_main));
        CUDA_CHECK(cudaStreamSynchronize(s_main));
        cudaEventElapsedTime(&t, ev_t[5], ev_t[6]); ms_rcond += t;
      } else if (ok) {
        rc2 = NAN;
      }
      if (!ok || wrong_rcond) {
        if (!ok) ++n_attempt_info; else ++n_attempt_rcond;
        return false;
      }
      return true;
    };
    // what the reference does after a rejected rung r (LinearAlgebra.cpp:75-90)
    auto after_reject = [&](int r) {
      if (r > max_inc)
        throw LkError{"[ERROR] Exceed max numerical nugget (" + std::to_string(r) + " x 1e" +
                      std::to_string(std::log10(num_nugget)) + ") added to force chol matrix"};
      if (num_nugget <= 0.0)
        throw LkError{"[ERROR] Cannot add numerical nugget which is not strictly positive: " + std::to_string(num_nugget)};
    };
    if (!updated) {
      // Ladder shortcut (lkgpu_set_ladder_shortcut, on by default; LKGPU_FULL_LADDER=1 turns it off).
      // safe_chol_lower returns the LOWEST accepted rung by trying rung 0, 1, 2, ...: k + 1 full factorisations when
      // the answer is k.  With the shortcut the first probe is the rung k >= 2 the previous evaluation on this handle
      // accepted (an optimiser walking through the numerically singular region stays near it):
      //  * rung k rejected: the ladder climbs on from k + 1, one rung at a time;
      //  * rung k accepted: the factor is set aside (A and V swap roles; the inverted diagonal blocks and per-panel
      //    log-determinants are copied, 20 MB at n = 20000) and the rungs below are tried downwards, one at a time,
      //    until one is rejected; the lowest accepted factor is taken back;
      //  * rung k accepted with an rcond so far above the threshold that the un-jittered matrix is expected to pass
      //    (rcond_1(L)^2 of R + j I grows about one decade per rung once j dominates the smallest eigenvalue): rung 0
      //    is probed first -- if it is accepted the answer is 0, which is the plain ladder's answer UNCONDITIONALLY
      //    (it tries rung 0 first), and a point that has left the singular region costs 2 factorisations, not k + 1.
      // Usual cost: 2 factorisations (k accepted, k - 1 rejected) instead of k + 1.  The result is the plain ladder's
      // whenever acceptance is monotone in the jitter below rung k (see DESIGN.md for the one observed exception).
      // Without the shortcut, and with a hint below 2, the probes are 0, 1, 2, ...: the plain ladder itself.
      // (a handle without history -- its first evaluation, or the first after new data -- runs the plain ladder)
      const bool shortcut = ladder_shortcut && evals_done > 0;
      const int hint = shortcut ? last_n_jitter : 0;
      const int top = max_inc + 1;  // the last rung the reference tries (LinearAlgebra.cpp:75-90)
      int lo = -1, hi = -1, n_fact = 0;
      double rc_hi = NAN;
      bool acc_in_stash = false, probed_zero = false;
      int probe = hint >= 2 ? std::min(hint, top) : 0;
      while (true) {
        if (hi >= 0 && !acc_in_stash) {
          stash_factor();
          acc_in_stash = true;
        }
        rc2 = NAN;
        ++n_fact;
        if (attempt(probe, need_inverse && probe == 0 && last_n_jitter == 0)) {
          hi = probe;
          rc_hi = rc2;
          acc_in_stash = false;
        } else {
          if (hi < 0) after_reject(probe);  // throws where the reference's ladder ends
          lo = probe;
        }
        if (hi >= 0 && hi - lo == 1) break;
        if (hi < 0) {
          probe = lo + 1;  // climb, one rung at a time (after_reject has thrown where the reference's ladder ends)
        } else if (lo < 0 && hi > 1 && !probed_zero && std::isfinite(rc_hi) &&
                   rc_hi >= min_rcond * std::pow(10.0, hi)) {
          probe = 0;       // the un-jittered matrix is expected to pass: if it does, 0 is the plain ladder's answer
          probed_zero = true;
        } else {
          probe = hi - 1;  // walk down, one rung at a time
        }
      }
      if (acc_in_stash) {
        unstash_factor();
        rc2 = rc_hi;
        have_W = false;
      }
      inc = hi;
      diag_add = ladder_diag(hi);
      n_rungs_skipped = std::max(0, (hi + 1) - n_fact);
    }
    if (need_inverse && !have_W) {
      // accepted on a later rung: L^-1 is formed once, after the ladder
      CUDA_CHECK(cudaEventRecord(ev_t[3], s_main));
      trtri();
      CUDA_CHECK(cudaEventRecord(ev_t[4], s_main));
      CUDA_CHECK(cudaStreamSynchronize(s_main));
      float t;
      cudaEventElapsedTime(&t, ev_t[3], ev_t[4]); ms_trtri += t;
    }
    last_diag_add = diag_add;
    last_n_jitter = inc;
    out->n_jitter = inc;
    out->rcond = rc2;
    out->info = 0;
    out->stage_ms[LKGPU_ST_COV] = ms_cov;
    out->stage_ms[LKGPU_ST_CHOL] = ms_chol;
    out->stage_ms[LKGPU_ST_RCOND] = ms_rcond;
    out->stage_ms[LKGPU_ST_TRTRI] = ms_trtri;
    out->stage_ms[LKGPU_CT_REJECT_INFO] = n_attempt_info;
    out->stage_ms[LKGPU_CT_REJECT_RCOND] = n_attempt_rcond;
    out->stage_ms[LKGPU_CT_RUNGS_SKIPPED] = n_rungs_skipped;

    // ---- sum log diag L ----
    launches += 1;
    sum_partials_kernel<<<1, 32, 0, s_main>>>(logdet_blocks, nb, 1, 1, dscal + SC_SUMLOG);

    // ---- GLS: Fstar, ystar, Rstar, beta, Estar, SSE, x  (KrigingImpl.cpp:102-123; Kriging.cpp:294) ----
    CUDA_CHECK(cudaEventRecord(ev_t[5], s_main));
    {
      const long long tot = (long long)N * (p + 1);
      launches += 2;
      if (p > 0) pad_copy_kernel<<<(unsigned)(((long long)N * p + 255) / 256), 256, 0, s_main>>>(dF, n, n, p, Bv, N, N);
      pad_copy_kernel<<<(N + 255) / 256, 256, 0, s_main>>>(dy, n, n, 1, Bv + (long long)N * p, N, N);
      (void)tot;
      solve_fwd(Bv, p + 1);
      const int chunks = (n + GRAM_CHUNK - 1) / GRAM_CHUNK;
      launches += 3;
      gram_partial_kernel<<<chunks, 256, 0, s_main>>>(Bv, N, n, p + 1, dpartial);
      gls_final_kernel<<<1, 256, ((p + 1) * (p + 1) + p) * 8, s_main>>>(dpartial, chunks, p, dRstar, dbeta, dinfo + 1);
      dev_zero(Ev, N);
      residual_kernel<<<(n + 255) / 256, 256, 0, s_main>>>(dy, dF, n, n, p, dbeta, Ev);
      solve_fwd(Ev, 1);
      sum_sq(Ev, SC_SSE);
      dev_copy(Xv, Ev, N);
      solve_bwd(Xv, 1);
      have_x = true;
    }
    CUDA_CHECK(cudaEventRecord(ev_t[6], s_main));

    int pdim = 0;
    if ((objective == LKGPU_OBJ_LMP || objective == LKGPU_OBJ_LOO) && p == 0)
      throw LkError{"LOO / LMP objectives need at least one trend column (regmodel 'none' is supported for LL only)"};
    if (objective == LKGPU_OBJ_LMP || objective == LKGPU_OBJ_LOO) lmp_loo_prepare(objective, pdim);
    CUDA_CHECK(cudaEventRecord(ev_t[7], s_main));

    if (need_inverse) lauum();
    CUDA_CHECK(cudaEventRecord(ev_t[8], s_main));
    if (want_grad && objective != LKGPU_OBJ_LOO) {
      if (objective == LKGPU_OBJ_LMP) grad_reduce(alpha, pdim, SC_GRAD, V, Qv, Qv);
      else grad_reduce(alpha, 0, SC_GRAD, V, Xv, Xv);
      launches += 2;
      const int g2 = std::min(64, (n + 255) / 256);
      diag_sums_kernel<<<g2, 256, 0, s_main>>>(V, ld, Xv, dnoise, n, dpartial + partial_doubles - 4 * 64);
      sum_partials_kernel<<<1, 32, 0, s_main>>>(dpartial + partial_doubles - 4 * 64, g2, 4, 4, dscal + SC_DIAG);
    }
    if (objective == LKGPU_OBJ_LOO) loo_finish(want_grad);
    CUDA_CHECK(cudaEventRecord(ev_t[9], s_main));

    // ---- results ----
    const int nres = SC_COUNT;
    std::vector<double> hres(nres);
    CUDA_CHECK(cudaMemcpyAsync(hpin, dscal, nres * 8, cudaMemcpyDeviceToHost, s_main));
    CUDA_CHECK(cudaMemcpyAsync(hpin + nres, dbeta, p * 8, cudaMemcpyDeviceToHost, s_main));
    CUDA_CHECK(cudaMemcpyAsync(hpin + nres + 256, dinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, s_main));
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    memcpy(hres.data(), hpin, nres * 8);
    int infos[2];
    memcpy(infos, hpin + nres + 256, sizeof(infos));
    if (infos[1] != 0) throw LkError{"chol(): decomposition failed (F*' F* not positive definite)"};
    out->sum_log_diagL = hres[SC_SUMLOG];
    out->SSEstar = hres[SC_SSE];
    out->sum_log_diagLX = h_sum_log_diagLX;
    out->S2 = h_S2;
    out->sum_x2 = hres[SC_DIAG + 0];
    out->trace_Rinv = hres[SC_DIAG + 1];
    out->sum_noise_Rinv = hres[SC_DIAG + 2];
    out->sum_noise_x2 = hres[SC_DIAG + 3];
    out->loo = hres[SC_LOO] / (double)n;
    if (out->betahat) memcpy(out->betahat, hpin + nres, p * 8);
    if (want_grad) {
      // dscal[SC_GRAD ..): S1[0..d], S2[0..d]
      const double* S1 = hres.data() + SC_GRAD;
      const double* S2v = S1 + (d + 1);
      if (objective != LKGPU_OBJ_LOO) {
        for (int k = 0; k < d; ++k) {
          if (out->t1) out->t1[k] = 2.0 * S1[k] * kp.inv_theta[k];
          if (out->t2) out->t2[k] = -2.0 * S2v[k] * kp.inv_theta[k];
        }
        out->sum_offdiag_xRx = 2.0 * S1[d];
        out->sum_offdiag_RinvR = 2.0 * S2v[d];
      } else if (out->obj_grad) {
        // dloo/dtheta_k = (2/n) sum_{l != j} G_k,lj (M_lj - v_l x_j)
        for (int k = 0; k < d; ++k) out->obj_grad[k] = (2.0 / (double)n) * (2.0 * S2v[k] - 2.0 * S1[k]) * kp.inv_theta[k];
      }
    }
    float t;
    cudaEventElapsedTime(&t, ev_t[5], ev_t[6]); out->stage_ms[LKGPU_ST_SOLVES] = t;
    cudaEventElapsedTime(&t, ev_t[6], ev_t[7]); out->stage_ms[LKGPU_ST_EXTRA] = t;
    cudaEventElapsedTime(&t, ev_t[7], ev_t[8]); out->stage_ms[LKGPU_ST_LAUUM] = t;
    cudaEventElapsedTime(&t, ev_t[8], ev_t[9]); out->stage_ms[LKGPU_ST_GRAD] = t;
    cudaEventElapsedTime(&t, ev_t[0], ev_t[9]); out->stage_ms[LKGPU_ST_TOTAL] = t;
    have_model = true;
    ++evals_done;
  }

  // LMP / LOO common: U = Rinv F LX^-T (n x p), S2, sum log diag LX.   (filled in lmp_loo.cuh)
  void lmp_loo_prepare(int objective, int& pdim);
  void loo_finish(int want_grad);
  void predict(int m, const double* Xn, const double* Fn, const double* beta, double r_on_factor, double* mean_out,
               double* var_out);

  // ---- exports ----
  void export_mat(int which, double* dst) {
    CUDA_CHECK(cudaSetDevice(device));
    SweepGate gate(*this);
    if (!have_model) throw LkError{"lkgpu_export: no evaluation has been run on this handle"};
    auto copy_square = [&](const double* src) {
      CUDA_CHECK(cudaMemcpy2D(dst, (size_t)n * 8, src, (size_t)ld * 8, (size_t)n * 8, n, cudaMemcpyDeviceToHost));
    };
    switch (which) {
      case LKGPU_EXPORT_L:
        CUDA_CHECK(cudaStreamSynchronize(s_main));
        copy_square(A);
        for (long long c = 0; c < n; ++c)
          for (long long r = 0; r < c; ++r) dst[c * n + r] = 0.0;
        break;
      case LKGPU_EXPORT_LINV:
        if (!have_W) { trtri(); }
        CUDA_CHECK(cudaStreamSynchronize(s_main));
        copy_square(W);
        for (long long c = 0; c < n; ++c)
          for (long long r = 0; r < c; ++r) dst[c * n + r] = 0.0;
        break;
      case LKGPU_EXPORT_RINV:
        if (!have_W) trtri();
        if (!have_V) lauum();
        CUDA_CHECK(cudaStreamSynchronize(s_main));
        copy_square(V);
        for (long long c = 0; c < n; ++c)
          for (long long r = 0; r < c; ++r) dst[c * n + r] = dst[r * n + c];
        break;
      case LKGPU_EXPORT_R: {
        // the un-jittered matrix (quirk (i) of SURVEY.md §8c): rebuilt into a temporary
        double* tmp = dalloc<double>((size_t)N * N);
        cov_build(tmp, last_alpha, last_inv_sigma2, 0.0, s_main);
        CUDA_CHECK(cudaStreamSynchronize(s_main));
        copy_square(tmp);
        cudaFree(tmp);
        for (long long c = 0; c < n; ++c)
          for (long long r = 0; r < c; ++r) dst[c * n + r] = dst[r * n + c];
        break;
      }
      case LKGPU_EXPORT_FSTAR:
        if (p > 0) CUDA_CHECK(cudaMemcpy2D(dst, (size_t)n * 8, Bv, (size_t)N * 8, (size_t)n * 8, p, cudaMemcpyDeviceToHost));
        break;
      case LKGPU_EXPORT_YSTAR:
        CUDA_CHECK(cudaMemcpy(dst, Bv + (long long)N * p, (size_t)n * 8, cudaMemcpyDeviceToHost));
        break;
      case LKGPU_EXPORT_RSTAR:
        if (p > 0) CUDA_CHECK(cudaMemcpy(dst, dRstar, (size_t)p * p * 8, cudaMemcpyDeviceToHost));
        break;
      case LKGPU_EXPORT_ESTAR:
        CUDA_CHECK(cudaMemcpy(dst, Ev, (size_t)n * 8, cudaMemcpyDeviceToHost));
        break;
      case LKGPU_EXPORT_X:
        CUDA_CHECK(cudaMemcpy(dst, Xv, (size_t)n * 8, cudaMemcpyDeviceToHost));
        break;
      case LKGPU_EXPORT_Z:
        // m_z (Kriging.cpp:2168-2172): Estar when beta is estimated, ystar - Fstar beta when it is fixed
        if (!has_fixed_beta || p == 0) {
          CUDA_CHECK(cudaMemcpy(dst, Ev, (size_t)n * 8, cudaMemcpyDeviceToHost));
        } else {
          std::vector<double> fs((size_t)n * p);
          CUDA_CHECK(cudaMemcpy2D(fs.data(), (size_t)n * 8, Bv, (size_t)N * 8, (size_t)n * 8, p, cudaMemcpyDeviceToHost));
          CUDA_CHECK(cudaMemcpy(dst, Bv + (long long)N * p, (size_t)n * 8, cudaMemcpyDeviceToHost));
          for (int q = 0; q < p; ++q)
            for (int i = 0; i < n; ++i) dst[i] -= fs[(size_t)q * n + i] * fixed_beta[q];
        }
        break;
      case LKGPU_EXPORT_LOO_ERR:
      case LKGPU_EXPORT_LOO_S2:
        if (!have_loo) throw LkError{"lkgpu_export: the last evaluation was not a LOO evaluation"};
        CUDA_CHECK(cudaMemcpy(dst, which == LKGPU_EXPORT_LOO_ERR ? dErr : dS2loo, (size_t)n * 8, cudaMemcpyDeviceToHost));
        break;
      default:
        throw LkError{"lkgpu_export: unknown selector"};
    }
  }

  // ---- sigma2 bounds of NoiseModel::Heterogeneous (f2; Kriging.cpp:1784-1797): variogram.cuh ----
  double sigma2_variogram() {
    CUDA_CHECK(cudaSetDevice(device));
    SweepGate gate(*this);
    const int t = (n + PT - 1) / PT;
    const int ntiles = t * (t + 1) / 2;
    const int grid = std::min(ntiles, 2 * sm_count);
    const size_t smem = (size_t)(2 * d * PT + 2 * PT) * 8;
    if (smem > 48 * 1024) {
      CUDA_CHECK(cudaFuncSetAttribute(vario_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      CUDA_CHECK(cudaFuncSetAttribute(vario_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    unsigned long long* dhist = dalloc<unsigned long long>(VG_BINS);
    std::vector<unsigned long long> hist(VG_BINS);
    const unsigned long long total = (unsigned long long)n * (unsigned long long)n;
    // key of the element of 0-based rank r in the sorted multiset of all n^2 ordered-pair distances;
    // *below = number of elements strictly smaller than that key
    auto select = [&](unsigned long long r, unsigned long long* below) -> unsigned long long {
      unsigned long long prefix = 0, skipped = 0;
      int hi_shift = 64;
      while (hi_shift > 0) {
        const int width = std::min(VG_DIGIT_BITS, hi_shift);
        const int shift = hi_shift - width;
        CUDA_CHECK(cudaMemsetAsync(dhist, 0, VG_BINS * sizeof(unsigned long long), s_main));
        ++launches;
        vario_hist_kernel<<<grid, PAIR_THREADS, smem, s_main>>>(dX, n, d, prefix, hi_shift, shift, (1u << width) - 1u, dhist,
                                                              ntiles);
        CUDA_CHECK(cudaGetLastError());
        CUDA_CHECK(cudaMemcpyAsync(hist.data(), dhist, VG_BINS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s_main));
        CUDA_CHECK(cudaStreamSynchronize(s_main));
        for (auto& h : hist) h *= 2ull;               // (i, j) and (j, i)
        if (prefix == 0) hist[0] += (unsigned long long)n;  // the diagonal: n zeros
        int b = 0;
        while (b < (1 << width) - 1 && r >= hist[b]) {
          r -= hist[b];
          skipped += hist[b];
          ++b;
        }
        if (r >= hist[b]) throw LkError{"lkgpu: radix select lost its rank (internal error)"};
        prefix = (prefix << width) | (unsigned long long)b;
        hi_shift = shift;
      }
      if (below) *below = skipped;
      return prefix;
    };
    auto as_double = [](unsigned long long k) {
      double v;
      memcpy(&v, &k, 8);
      return v;
    };
    const unsigned long long half = total / 2;
    unsigned long long below = 0;
    const double val1 = as_double(select(half, &below));  // op_median::direct_median: *nth
    double med = val1;
    if (total % 2 == 0) {
      // val2 = max of the lower half = rank half - 1: the same key unless `half` is the first of its run
      const double val2 = (below < half) ? val1 : as_double(select(half - 1, nullptr));
      med = val1 + (val2 - val1) / 2.0;  // op_mean::robust_mean
    }
    launches += 2;
    vario_sum_kernel<<<grid, PAIR_THREADS, smem, s_main>>>(dX, dy, n, d, med, dpartial, ntiles);
    sum_partials_kernel<<<1, 32, 0, s_main>>>(dpartial, grid, 2, 2, dscal + SC_BOUNDS);
    CUDA_CHECK(cudaGetLastError());
    CUDA_CHECK(cudaMemcpyAsync(hpin, dscal + SC_BOUNDS, 16, cudaMemcpyDeviceToHost, s_main));
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    cudaFree(dhist);
    const double sum = 2.0 * hpin[0];
    const double cnt = 2.0 * hpin[1] + ((0.0 >= med) ? (double)n : 0.0);
    return 0.5 * sum / cnt;
  }

  // ---- theta bounds (a8) ----
  void theta_bounds(double lo_f, double up_f, int heuristic, double* lower, double* upper) {
    CUDA_CHECK(cudaSetDevice(device));
    SweepGate gate(*this);
    launches += 1;
    col_minmax_kernel<<<d, 256, 0, s_main>>>(dX, n, dscal + SC_BOUNDS, dscal + SC_BOUNDS + LK_MAX_D);
    CUDA_CHECK(cudaMemcpyAsync(hpin, dscal + SC_BOUNDS, 2 * LK_MAX_D * 8, cudaMemcpyDeviceToHost, s_main));
    CUDA_CHECK(cudaStreamSynchronize(s_main));
    std::vector<double> maxdX(d);
    for (int k = 0; k < d; ++k) {
      maxdX[k] = hpin[LK_MAX_D + k] - hpin[k];
      lower[k] = lo_f * maxdX[k];
      upper[k] = up_f * maxdX[k];
    }
    if (heuristic && n > 1) {
      const int t = (n + PT - 1) / PT;
      const int ntiles = t * (t + 1) / 2;
      const int grid = std::min(ntiles, 2 * sm_count);
      const size_t smem = ((size_t)2 * d * PT + 2 * PT + 8 * (d + 1)) * 8;
      if (smem > 48 * 1024)
        CUDA_CHECK(cudaFuncSetAttribute(theta_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      launches += 2;
      theta_bounds_kernel<<<grid, PAIR_THREADS, smem, s_main>>>(dX, dy, n, d, dpartial, ntiles);
      sum_partials_kernel<<<(d + 1 + 63) / 64, 64, 0, s_main>>>(dpartial, grid, d + 1, d + 1, dscal + SC_BOUNDS);
      CUDA_CHECK(cudaGetLastError());
      CUDA_CHECK(cudaMemcpyAsync(hpin, dscal + SC_BOUNDS, (d + 1) * 8, cudaMemcpyDeviceToHost, s_main));
      CUDA_CHECK(cudaStreamSynchronize(s_main));
      const double wsum = hpin[d];
      if (wsum > 0.0) {
        for (int k = 0; k < d; ++k) {
          const double steep = hpin[k] / wsum;
          lower[k] = std::max(lower[k], lo_f * steep);
          lower[k] = std::min(lower[k], upper[k]);
          upper[k] = std::max(lower[k], upper[k]);
        }
      }
    }
  }
};

#include "lmp_loo.cuh"
#include "objective.inl"

}  // namespace

// ============================== C ABI ==============================
#define LK_TRY try {
#define LK_CATCH                                   \
  }                                                \
  catch (const LkError& e) {                       \
    g_last_error = e.msg;                          \
    return -1;                                     \
  }                                                \
  catch (const std::exception& e) {                \
    g_last_error = e.what();                       \
    return -1;                                     \
  }                                                \
  catch (...) {                                    \
    g_last_error = "unknown error";                \
    return -1;                                     \
  }                                                \
  return 0;

extern "C" {

int lkgpu_abi_version(void) { return LKGPU_ABI_VERSION; }
const char* lkgpu_last_error(void) { return g_last_error.c_str(); }

int lkgpu_create(void** handle, int device, int n, int d, int p, const double* X, const double* y, const double* F,
                 const double* noise, int kernel, int noise_model) {
  LK_TRY
  if (!handle || !X || !y || (!F && p > 0)) throw LkError{"lkgpu_create: null argument"};
  *handle = nullptr;
  Engine* e = new Engine();
  try {
    e->init(device, n, d, p, X, y, F, noise, kernel, noise_model);
  } catch (...) {
    delete e;
    throw;
  }
  *handle = e;
  LK_CATCH
}

int lkgpu_set_numerics(void* handle, double num_nugget, int max_inc_choldiag, double min_rcond, int chol_rcond_check) {
  LK_TRY
  if (!handle) throw LkError{"null handle"};
  Engine* e = static_cast<Engine*>(handle);
  e->num_nugget = num_nugget;
  e->max_inc = max_inc_choldiag;
  e->min_rcond = min_rcond;
  e->rcond_check = chol_rcond_check != 0;
  LK_CATCH
}

int lkgpu_set_data(void* handle, const double* X, const double* y, const double* F, const double* noise) {
  LK_TRY
  if (!handle || !X || !y || (!F && static_cast<Engine*>(handle)->p > 0)) throw LkError{"lkgpu_set_data: null argument"};
  Engine* e = static_cast<Engine*>(handle);
  CUDA_CHECK(cudaSetDevice(e->device));
  e->set_data(X, y, F, noise);
  LK_CATCH
}

int lkgpu_append_data(void* handle, int n_u, const double* X_u, const double* y_u, const double* F_u,
                      const double* noise_u) {
  LK_TRY
  if (!handle || !X_u || !y_u || (!F_u && static_cast<Engine*>(handle)->p > 0))
    throw LkError{"lkgpu_append_data: null argument"};
  static_cast<Engine*>(handle)->append(n_u, X_u, y_u, F_u, noise_u);
  LK_CATCH
}

int lkgpu_commit_model(void* handle) {
  LK_TRY
  if (!handle) throw LkError{"null handle"};
  static_cast<Engine*>(handle)->commit();
  LK_CATCH
}

int lkgpu_restore_model(void* handle) {
  LK_TRY
  if (!handle) throw LkError{"null handle"};
  static_cast<Engine*>(handle)->restore();
  LK_CATCH
}

int lkgpu_set_concurrent(void* handle, int flag) {
  LK_TRY
  if (!handle) throw LkError{"null handle"};
  static_cast<Engine*>(handle)->concurrent_flag = flag != 0;
  LK_CATCH
}

int lkgpu_set_fixed_beta(void* handle, const double* beta) {
  LK_TRY
  if (!handle) throw LkError{"null handle"};
  Engine* e = static_cast<Engine*>(handle);
  e->has_fixed_beta = beta != nullptr;
  if (beta) e->fixed_beta.assign(beta, beta + e->p);
  else e->fixed_beta.clear();
  LK_CATCH
}

int lkgpu_set_ladder_shortcut(void* handle, int flag) {
  LK_TRY
  if (!handle) throw LkError{"null handle"};
  static_cast<Engine*>(handle)->ladder_shortcut = flag != 0;
  LK_CATCH
}

int lkgpu_last_eval_was_update(void* handle) {
  return handle ? (static_cast<Engine*>(handle)->last_was_update ? 1 : 0) : 0;
}

int lkgpu_theta_bounds(void* handle, double lower_factor, double upper_factor, int heuristic, double* lower,
                       double* upper) {
  LK_TRY
  if (!handle || !lower || !upper) throw LkError{"lkgpu_theta_bounds: null argument"};
  static_cast<Engine*>(handle)->theta_bounds(lower_factor, upper_factor, heuristic, lower, upper);
  LK_CATCH
}

int lkgpu_sigma2_variogram(void* handle, double* sigma2_variogram) {
  LK_TRY
  if (!handle || !sigma2_variogram) throw LkError{"lkgpu_sigma2_variogram: null argument"};
  *sigma2_variogram = static_cast<Engine*>(handle)->sigma2_variogram();
  LK_CATCH
}

int lkgpu_eval(void* handle, int objective, const double* theta, double extra, int want_grad, lkgpu_out* out) {
  LK_TRY
  if (!handle || !theta || !out) throw LkError{"lkgpu_eval: null argument"};
  static_cast<Engine*>(handle)->eval(objective, theta, extra, want_grad, out);
  LK_CATCH
}

int lkgpu_set_params(void* handle, int est_sigma2, double sigma2, int est_nugget, double nugget, double alpha) {
  LK_TRY
  if (!handle) throw LkError{"null handle"};
  Engine* e = static_cast<Engine*>(handle);
  e->est_sigma2 = est_sigma2 != 0;
  e->est_nugget = est_nugget != 0;
  e->sigma2 = sigma2;
  e->nugget = nugget;
  e->alpha0 = alpha;
  LK_CATCH
}

int lkgpu_objective_fun(void* handle, int objective, const double* gamma, int gamma_n, int return_grad,
                        double* value_out, double* grad_out, lkgpu_out* out) {
  LK_TRY
  if (!handle || !gamma || !value_out) throw LkError{"lkgpu_objective_fun: null argument"};
  if (return_grad && !grad_out) throw LkError{"lkgpu_objective_fun: return_grad set but grad_out is null"};
  Engine* e = static_cast<Engine*>(handle);
  *value_out = objective_fun(*e, objective, gamma, gamma_n, return_grad ? grad_out : nullptr, out);
  LK_CATCH
}

int lkgpu_export(void* handle, int which, double* dst) {
  LK_TRY
  if (!handle || !dst) throw LkError{"lkgpu_export: null argument"};
  static_cast<Engine*>(handle)->export_mat(which, dst);
  LK_CATCH
}

int lkgpu_predict(void* handle, int m, const double* Xn, const double* Fn, const double* beta, double r_on_factor,
                  double* mean_out, double* var_out) {
  LK_TRY
  if (!handle || !Xn || !mean_out || ((!Fn || !beta) && static_cast<Engine*>(handle)->p > 0))
    throw LkError{"lkgpu_predict: null argument"};
  static_cast<Engine*>(handle)->predict(m, Xn, Fn, beta, r_on_factor, mean_out, var_out);
  LK_CATCH
}

void lkgpu_destroy(void* handle) {
  if (handle) delete static_cast<Engine*>(handle);
}

void* lkgpu_get_stream(void* handle) { return handle ? (void*)static_cast<Engine*>(handle)->s_main : nullptr; }

long long lkgpu_launch_count(void* handle) { return handle ? static_cast<Engine*>(handle)->launches : 0; }

int lkgpu_mem_info(int device, unsigned long long* free_bytes, unsigned long long* total_bytes) {
  LK_TRY
  if (!free_bytes || !total_bytes) throw LkError{"null argument"};
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw LkError{"lkgpu: no CUDA device available"};
  CUDA_CHECK(cudaSetDevice(device));
  size_t f = 0, t = 0;
  CUDA_CHECK(cudaMemGetInfo(&f, &t));
  *free_bytes = f;
  *total_bytes = t;
  LK_CATCH
}

#ifdef LKGPU_POTF2_PROFILE
// debug builds only (tools/potf2_phases.py): clock64() stamps of the last panel kernel that ran
int lkgpu_debug_potf2_profile(long long* out32) {
  return cudaMemcpyFromSymbol(out32, lk::g_potf2_prof, 32 * sizeof(long long)) == cudaSuccess ? 0 : -1;
}
#endif

int lkgpu_probe_fp64_peak(int device, int mode, double* tflops) {
  LK_TRY
  if (!tflops) throw LkError{"null argument"};
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) throw LkError{"lkgpu: no CUDA device available"};
  CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  double* dout;
  CUDA_CHECK(cudaMalloc(&dout, 8));
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0));
  CUDA_CHECK(cudaEventCreate(&e1));
  const int grid = prop.multiProcessorCount * 2, iters = 4096;
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CUDA_CHECK(cudaEventRecord(e0));
    if (mode == 0) probe_dmma_kernel<<<grid, 256>>>(dout, iters);
    else if (mode == 1) probe_dfma_kernel<<<grid, 256>>>(dout, iters);
    else probe_mixed_kernel<<<grid, 256>>>(dout, iters);
    CUDA_CHECK(cudaEventRecord(e1));
    CUDA_CHECK(cudaEventSynchronize(e1));
    float ms;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    double flops;
    const double warps = (double)grid * 8;
    if (mode == 0) flops = warps * iters * 16.0 * 512.0;
    else if (mode == 1) flops = warps * 32 * iters * 32.0 * 2.0;
    else flops = warps * iters * (8.0 * 512.0 + 32 * 16.0 * 2.0);
    best = std::max(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaFree(dout);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops = best;
  LK_CATCH
}

}  // extern "C"
