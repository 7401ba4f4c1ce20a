/* lkgpu.h -- C ABI of the B200-native likelihood engine (liblkgpu.so).
 *
 * Drop-in boundary for the objective-evaluation hot path of libKriging's
 * Kriging::fit.  The reference has no plugin interface for this path; the seam
 * is the std::function `FitOfn` (src/lib/include/libKriging/Kriging.hpp:285)
 * called by lbfgsb::Optimizer::minimize (src/lib/Kriging.cpp:1980-1987) plus
 * the workspace struct KModel (src/lib/include/libKriging/KrigingImpl.hpp:26-38).
 * This ABI follows the conventions of the reference's only C ABI, the Julia
 * shim (bindings/Julia/jlibkriging/csrc/libkriging_c.h:8-10): opaque handles,
 * int return 0 / -1, thread-local last-error string, caller-allocated output
 * buffers, column-major doubles.
 *
 * Every function is host-callable from one thread per handle.  No torch types,
 * no C++ types.  There is NO CPU fallback: every entry point fails with -1 if
 * no CUDA device is usable.
 */
#ifndef LKGPU_H
#define LKGPU_H

#ifdef __cplusplus
extern "C" {
#endif

#define LKGPU_ABI_VERSION 1

/* covariance kernels: Covariance::resolve (src/lib/Covariance.cpp:219-231) */
enum { LKGPU_KERNEL_GAUSS = 0, LKGPU_KERNEL_EXP = 1, LKGPU_KERNEL_MATERN32 = 2, LKGPU_KERNEL_MATERN52 = 3 };
/* Kriging::NoiseModel (src/lib/include/libKriging/Kriging.hpp:45-49) */
enum { LKGPU_NOISE_NONE = 0, LKGPU_NOISE_NUGGET = 1, LKGPU_NOISE_HETERO = 2 };
/* objectives: _logLikelihood / _leaveOneOut / _logMargPost (src/lib/Kriging.cpp:214, 353, 488) */
enum { LKGPU_OBJ_LL = 0, LKGPU_OBJ_LOO = 1, LKGPU_OBJ_LMP = 2 };
/* lkgpu_export selectors: the KModel members fit() commits (src/lib/Kriging.cpp:2156-2173) */
enum {
  LKGPU_EXPORT_L = 0,      /* n*n, lower Cholesky factor, strict upper zeroed (m_T)          */
  LKGPU_EXPORT_R = 1,      /* n*n, un-jittered correlation matrix, both triangles (m_R)      */
  LKGPU_EXPORT_RINV = 2,   /* n*n, R^-1 of the (jittered) matrix, both triangles (m_Rinv)    */
  LKGPU_EXPORT_FSTAR = 3,  /* n*p (m_M)                                                      */
  LKGPU_EXPORT_RSTAR = 4,  /* p*p upper (m_circ)                                             */
  LKGPU_EXPORT_YSTAR = 5,  /* n                                                              */
  LKGPU_EXPORT_ESTAR = 6,  /* n  (m_z)                                                       */
  LKGPU_EXPORT_LINV = 7,   /* n*n, L^-1 lower (LOO's KModel::Linv)                           */
  LKGPU_EXPORT_X = 8,      /* n, x = L^-T Estar                                              */
  LKGPU_EXPORT_LOO_ERR = 9,  /* n, errorsLOO (yhat_loo = y - errorsLOO), after a LOO evaluation   */
  LKGPU_EXPORT_LOO_S2 = 10,  /* n, sigma2LOO = 1/diag(B) (unscaled), after a LOO evaluation       */
  LKGPU_EXPORT_Z = 11        /* n, m_z: Estar, or ystar - Fstar beta after lkgpu_set_fixed_beta     */
};

#define LKGPU_N_STAGES 12
/* stage_ms[] indices; names mirror the reference's Bench keys (src/lib/KrigingImpl.cpp:94-123,
 * src/lib/Kriging.cpp:294-302) */
enum {
  LKGPU_ST_COV = 0,     /* "R = _Cov(dX)"                       */
  LKGPU_ST_CHOL = 1,    /* "L = Chol(R)"                        */
  LKGPU_ST_RCOND = 2,   /* rcond_chol (LinearAlgebra.cpp:106)   */
  LKGPU_ST_SOLVES = 3,  /* F*, y*, beta, z, SSE, x              */
  LKGPU_ST_TRTRI = 4,   /* L^-1                                 */
  LKGPU_ST_LAUUM = 5,   /* "R^-1 = L^-T * L^-1"                 */
  LKGPU_ST_GRAD = 6,    /* "gradR computation"                  */
  LKGPU_ST_EXTRA = 7,   /* LOO / LMP specific work              */
  LKGPU_ST_TOTAL = 8,   /* whole evaluation, device time        */
  /* counters kept in the spare slots (not times): rungs of safe_chol_lower's ladder rejected ...            */
  LKGPU_CT_REJECT_INFO = 9,   /* ... because the factorisation failed (non-positive pivot; attempt aborted) */
  LKGPU_CT_REJECT_RCOND = 10, /* ... because rcond_1(L)^2 < min_rcond                                      */
  LKGPU_CT_RUNGS_SKIPPED = 11 /* rungs the ladder shortcut did not factor (lkgpu_set_ladder_shortcut)      */
};

typedef struct lkgpu_out {
  /* ---- scalars, always filled ---- */
  double sum_log_diagL; /* sum_i log L_ii                                   */
  double SSEstar;       /* ||L \ (y - F betahat)||^2                        */
  double rcond;         /* rcond_1(L)^2 as compared with min_rcond          */
  int n_jitter;         /* number of diagonal bumps applied (cumulative)    */
  int info;             /* 0 ok; >0 = Cholesky failed even at max jitter    */
  /* ---- LL extras (want_grad) ---- */
  double sum_offdiag_xRx;    /* sum_{i!=j} x_i x_j R_ij   (R includes alpha) */
  double sum_offdiag_RinvR;  /* sum_{i!=j} Rinv_ij R_ij                      */
  double sum_x2;             /* sum_i x_i^2                                  */
  double trace_Rinv;         /* sum_i Rinv_ii                                */
  double sum_noise_Rinv;     /* sum_i noise_i Rinv_ii   (hetero)             */
  double sum_noise_x2;       /* sum_i noise_i x_i^2     (hetero)             */
  /* ---- LMP ---- */
  double sum_log_diagLX;     /* sum log diag chol(F' R^-1 F)                 */
  double S2;                 /* y' R^-1 y - y' P y                           */
  /* ---- LOO ---- */
  double loo;                /* sum(errorsLOO^2) / n                         */
  /* ---- timing ---- */
  double stage_ms[LKGPU_N_STAGES];
  /* ---- caller-allocated arrays (may be NULL to skip) ---- */
  double* betahat;  /* [p] GLS beta                                            */
  double* t1;       /* [d] LL: 2 sum_{i>j} x_i x_j R_ij g_k ; LMP: same        */
  double* t2;       /* [d] LL: -2 sum_{i>j} Rinv_ij R_ij g_k ; LMP: with Rinv-P */
  double* obj_grad; /* [d] LOO: d loo / d theta_k                              */
} lkgpu_out;

/* Create an engine for one (process, start).  X: n*d column-major, already
 * normalised by the caller (fit_setup_impl, src/lib/KrigingImpl.cpp:780-802);
 * y: n; F: n*p column-major trend matrix (Trend::regressionModelMatrix);
 * noise: n or NULL.  Allocates all device workspaces (a1: KModel). */
int lkgpu_create(void** handle, int device, int n, int d, int p, const double* X, const double* y, const double* F,
                 const double* noise, int kernel, int noise_model);

/* LinearAlgebra::{num_nugget,max_inc_choldiag,min_rcond,chol_rcond_check}
 * (src/lib/LinearAlgebra.cpp:33,53,63,100) -- process globals in the reference,
 * per-handle here. */
int lkgpu_set_numerics(void* handle, double num_nugget, int max_inc_choldiag, double min_rcond, int chol_rcond_check);

/* Optim::theta_bounds + m_maxdX (src/lib/Optim.cpp:179-209, src/lib/KrigingImpl.cpp:807)
 * computed from X, y on the device without materialising dX.  lower/upper: [d]. */
int lkgpu_theta_bounds(void* handle, double lower_factor, double upper_factor, int heuristic, double* lower,
                       double* upper);

/* sigma2 bounds of NoiseModel::Heterogeneous (src/lib/Kriging.cpp:1784-1797):
 *   *sigma2_variogram = 0.5 * mean(dy2[dX2 >= median(dX2)]) over the n^2 ordered pairs (diagonal included),
 * by a radix select over the pair distances regenerated from X tiles -- neither dX nor dX2 nor dy2 exists.
 * The caller forms extra_lower = 0.1 (s - max noise), extra_upper = 10 (s - min noise) (:1798-1799). */
int lkgpu_sigma2_variogram(void* handle, double* sigma2_variogram);

/* One objective evaluation = populate_Model (src/lib/KrigingImpl.cpp:73-125)
 * + the objective-specific reductions.  theta: [d]; extra = alpha (nugget) or
 * sigma2 (hetero), ignored for none.  sigma2_for_lmp unused unless LMP. */
int lkgpu_eval(void* handle, int objective, const double* theta, double extra, int want_grad, lkgpu_out* out);

/* m_est_sigma2 / m_sigma2 / m_est_nugget / m_nugget / m_alpha of the reference's Kriging object
 * (src/lib/include/libKriging/Kriging.hpp; used by the LL / LMP formulas at src/lib/Kriging.cpp:247-289, 579-596).
 * Defaults: everything estimated, sigma2 = 1, nugget = 0, alpha = 1. */
int lkgpu_set_params(void* handle, int est_sigma2, double sigma2, int est_nugget, double nugget, double alpha);

/* m_est_beta == false (Parameters::is_beta_estim = false with beta given, src/lib/Kriging.cpp:1668-1676): the trend
 * coefficients are fixed by the caller.  The objectives keep using the GLS estimate (src/lib/KrigingImpl.cpp:113-123);
 * what changes is the committed z = ystar - M beta (src/lib/Kriging.cpp:2168-2172; KrigingImpl.cpp:611-617), which
 * lkgpu_predict and LKGPU_EXPORT_Z then use.  beta: [p] (in the model's normalised output scale), or NULL to go back
 * to the estimated beta. */
int lkgpu_set_fixed_beta(void* handle, const double* beta);

/* The reference's objective functions, value and analytic gradient in the reference's parametrisation:
 * Kriging::_logLikelihood / _leaveOneOut / _logMargPost (src/lib/Kriging.cpp:214, 353, 488); same calling
 * convention as the Julia shim's lk_kriging_log_likelihood_fun
 * (bindings/Julia/jlibkriging/csrc/libkriging_c.h:109-130).
 * gamma: [theta (d)] or [theta, alpha | sigma2] (d+1) for Nugget / Heterogeneous; grad_out: [gamma_n].
 * out (may be NULL) receives the raw reductions and stage timings of the evaluation. */
int lkgpu_objective_fun(void* handle, int objective, const double* gamma, int gamma_n, int return_grad,
                        double* value_out, double* grad_out, lkgpu_out* out);

/* Copy a member of the last evaluation's model to host memory (column-major). */
int lkgpu_export(void* handle, int which, double* dst);

/* predict mean / variance factor at m new points (src/lib/KrigingImpl.cpp:145-243),
 * using the model of the last evaluation.  Xn: m*d column-major (normalised),
 * Fn: m*p; beta: [p]; r_on_factor: alpha (nugget) else 1.
 * mean_out[m] = Fn beta + Rstar_on' z  (z = Estar, or ystar - Fstar beta after lkgpu_set_fixed_beta);  var_out[m] = 1 - colsum(Rstar_on^2) + rowsum(Ecirc^2)
 * (clamped at 0, not yet scaled by sigma2). */
int lkgpu_predict(void* handle, int m, const double* Xn, const double* Fn, const double* beta, double r_on_factor,
                  double* mean_out, double* var_out);

/* Replace X / y / F / noise of an existing handle (same n, d, p). */
int lkgpu_set_data(void* handle, const double* X, const double* y, const double* F, const double* noise);

/* The reference keeps the committed model (m_T, m_M, m_z, m_circ, m_beta; commit at src/lib/Kriging.cpp:2156-2173)
 * apart from the per-evaluation KModel workspaces, so objective calls at other points never disturb predict().
 * lkgpu_commit_model snapshots the model of the last evaluation into a device-side store (one extra n*n buffer,
 * allocated on the first commit); lkgpu_restore_model makes it the live model again (a device-to-device copy, no
 * re-evaluation; a no-op when nothing was evaluated since).  lkgpu_append_data extends the committed model. */
int lkgpu_commit_model(void* handle);
int lkgpu_restore_model(void* handle);

/* Kriging::update, data side (src/lib/Kriging.cpp:2476-2491; KrigingImpl::update_no_refit_impl,
 * src/lib/KrigingImpl.cpp:576-610): append n_u observations (X_u: n_u*d column-major, normalised with the model's
 * own centre / scale; y_u: n_u; F_u: n_u*p; noise_u: n_u or NULL) to the handle's data set.  The device workspaces
 * are re-created for n + n_u rows and the Cholesky factor of the last evaluation -- the caller makes that the
 * committed model, the reference's m_T -- is kept, so that the next lkgpu_eval / lkgpu_objective_fun at the SAME
 * theta (and alpha | sigma2) runs as the block extension of LinearAlgebra::update_cholCov / chol_block
 * (src/lib/LinearAlgebra.cpp:206-299) instead of a factorisation from scratch: populate_Model's `update_eligible`
 * (src/lib/Kriging.cpp:170-188), including chol_block's fall-back to a full safe_chol_lower when the ladder on the
 * Schur complement is exhausted.  An evaluation at any other point factors from scratch and drops the kept factor. */
int lkgpu_append_data(void* handle, int n_u, const double* X_u, const double* y_u, const double* F_u,
                      const double* noise_u);
/* safe_chol_lower's ladder (src/lib/LinearAlgebra.cpp:68-98) returns the LOWEST accepted rung by trying rung 0, 1,
 * 2, ...: k + 1 factorisations when the answer is k.  With the shortcut (default on; flag = 0 or the environment
 * variable LKGPU_FULL_LADDER=1 restore the plain ladder) an evaluation on a handle whose PREVIOUS evaluation was
 * accepted on rung k >= 2 probes rung k first: rejected, it climbs on from k + 1; accepted, the factor is set aside and
 * the rungs below are tried downwards until one is rejected (rung 0 first when the rcond of rung k says the
 * un-jittered matrix should pass: if rung 0 is accepted the answer is 0 unconditionally).  Usually 2 factorisations
 * instead of k + 1; same n_jitter, factor and value as the plain ladder WHENEVER acceptance is monotone in the jitter
 * below rung k.  It is not always: in the numerically singular regime a low rung can be "accepted" by rounding luck
 * while a higher one is rejected (DESIGN.md documents a case); a caller that needs the plain ladder's decision there
 * turns the shortcut off.  The first evaluation on a handle (or after new data) always runs the plain ladder.
 * stage_ms[LKGPU_CT_RUNGS_SKIPPED] counts the factorisations saved. */
int lkgpu_set_ladder_shortcut(void* handle, int flag);
/* 1 if the last evaluation on this handle ran as a block extension of a kept factor, else 0 */
int lkgpu_last_eval_was_update(void* handle);

void lkgpu_destroy(void* handle);
const char* lkgpu_last_error(void);
int lkgpu_abi_version(void);
/* number of kernels launched by this handle since creation (bench.py "gpu_launches") */
long long lkgpu_launch_count(void* handle);
/* the CUDA stream (cudaStream_t) every kernel of this handle is launched on; bench.py records its
 * CUDA events on it.  The main stream is non-blocking with respect to the legacy default stream. */
void* lkgpu_get_stream(void* handle);
/* FP64 DMMA peak probe (dependent-free mma.sync.m8n8k4.f64 chains on every SM):
 * returns TFLOP/s in *tflops.  Used by bench.py for the roofline denominator. */
int lkgpu_probe_fp64_peak(int device, int mode, double* tflops);

/* Evaluations of different handles on one device are exclusive by default: they queue, and only their host work
 * overlaps (results are those of a lone handle, bit for bit).  Handles flagged here overlap with each other (one per
 * multistart row in flight, BASELINE cfg 5; one per NestedKriging sub-model): the throughput mode for many mid-size
 * factorisations.  The default is conservative for historical reasons ("The ring release, and concurrent
 * handles"): the bug that made overlapping evaluations deviate is fixed -- 0 deviations in 2400 overlapping
 * evaluations since -- but the policy has not been re-validated for removal yet. */
int lkgpu_set_concurrent(void* handle, int flag);

/* Free / total bytes of device memory: the host sizes the number of concurrent handles (one per
 * multistart row in flight, BASELINE cfg 5) with it.  The reference preallocates one KModel per start
 * (src/lib/Kriging.cpp:1861-1874) without such a check. */
int lkgpu_mem_info(int device, unsigned long long* free_bytes, unsigned long long* total_bytes);

#ifdef __cplusplus
}
#endif
#endif /* LKGPU_H */
