"""oracle/kriging_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement (numpy + LAPACK through scipy) of libKriging's objective
evaluation hot path, function by function, citing the reference file:line each
one follows.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
leg may import this module; the product path (libkriging_b200) never does.

Parity status: PINNED.  tests/test_oracle_golden.py checks this module against
(i) the reference's own 11 golden vectors (tests/references/data1-scal-*,
data2-grad-{1..10}-* of the reference, converted into
tests/golden/reference_vectors.json by tests/golden/make_golden.py) at the
reference's own 1e-12 tolerance, and (ii) outputs of the reference itself run
in the build container (oracle/_ref/ref_driver, fixtures in
tests/golden/refgen_vectors.json) for the kernels / noise models / objectives
that have no golden file.

All matrices are dense float64; X is (n, d); theta is (d,).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
from scipy.linalg import lapack

KERNELS = ("gauss", "exp", "matern3_2", "matern5_2")
SQRT3 = math.sqrt(3.0)
SQRT5 = math.sqrt(5.0)


# --------------------------------------------------------------------------
# Numerics knobs (reference: src/lib/LinearAlgebra.cpp:33, 47, 53, 63, 100)
# --------------------------------------------------------------------------
@dataclass
class Numerics:
    num_nugget: float = 1e-10
    max_inc_choldiag: int = 10
    min_rcond: float = 1e-18
    chol_rcond_check: bool = True


# --------------------------------------------------------------------------
# Covariance kernels  (src/lib/Covariance.cpp:24-36, 65-75, 104-116, 149-161)
# --------------------------------------------------------------------------
def _pair_diffs(X: np.ndarray) -> np.ndarray:
    """dX[i, j, k] = X[i,k] - X[j,k]  (src/lib/LinearAlgebra.cpp:440-485)."""
    return X[:, None, :] - X[None, :, :]


def corr_from_dx(dx: np.ndarray, theta: np.ndarray, kernel: str) -> np.ndarray:
    """rho(dx; theta) for arrays dx[..., d] -- Cov_* of the reference."""
    u = dx / theta
    if kernel == "gauss":
        return np.exp(-0.5 * np.sum(u * u, axis=-1))
    if kernel == "exp":
        return np.exp(-np.sum(np.abs(u), axis=-1))
    if kernel == "matern3_2":
        s = SQRT3 * np.abs(u)
        return np.exp(-np.sum(s - np.log1p(s), axis=-1))
    if kernel == "matern5_2":
        s = SQRT5 * np.abs(u)
        return np.exp(-np.sum(s - np.log1p(s + (s * s) / 3.0), axis=-1))
    raise ValueError(f"Unsupported covariance kernel: {kernel}")


def dlncorr_dtheta(dx: np.ndarray, theta: np.ndarray, kernel: str) -> np.ndarray:
    """d ln rho / d theta_k, shape dx.shape -- DlnCovDtheta_* of the reference
    (src/lib/Covariance.cpp:38-50, 77-88, 118-130, 163-179)."""
    if kernel == "gauss":
        return (dx * dx) / (theta * theta * theta)
    if kernel == "exp":
        return np.abs(dx) / (theta * theta)
    if kernel == "matern3_2":
        s = SQRT3 * np.abs(dx / theta)
        return (s * s) / (1.0 + s) / theta
    if kernel == "matern5_2":
        s = SQRT5 * np.abs(dx / theta)
        a = 1.0 + s
        b = (s * s) / 3.0
        return (a * b) / (a + b) / theta
    raise ValueError(kernel)


def build_R(X, theta, kernel, alpha=1.0, diag=None, block=1024):
    """cholCov's build loop (src/lib/LinearAlgebra.cpp:148-194): off-diagonal
    alpha * rho_ij, diagonal 1 (or `diag`)."""
    n = X.shape[0]
    R = np.empty((n, n))
    for i0 in range(0, n, block):
        i1 = min(n, i0 + block)
        dx = X[i0:i1, None, :] - X[None, :, :]
        R[i0:i1, :] = corr_from_dx(dx, theta, kernel)
    R *= alpha
    if diag is None:
        np.fill_diagonal(R, 1.0)
    else:
        np.fill_diagonal(R, diag)
    return R


# --------------------------------------------------------------------------
# safe_chol_lower + rcond_chol  (src/lib/LinearAlgebra.cpp:43-98, 106-115)
# --------------------------------------------------------------------------
def rcond_chol(L: np.ndarray) -> float:
    """arma::rcond on a triangular Mat -> dtrcon('1','L','N'), squared."""
    rc, info = lapack.dtrcon(L, norm="1", uplo="L", diag="N")
    assert info == 0
    return float(rc) * float(rc)


def safe_chol_lower(R: np.ndarray, num: Numerics | None = None):
    """Returns (L, n_jitter, rcond2).  Jitter is cumulative:
    X.diag() += num_nugget * 10**inc, inc = 0, 1, ... (LinearAlgebra.cpp:80-90)."""
    num = num or Numerics()
    Xm = np.array(R, order="F", copy=True)
    inc = 0
    while True:
        L, info = lapack.dpotrf(Xm, lower=1, clean=1, overwrite_a=0)
        ok = info == 0
        wrong = num.chol_rcond_check
        rc2 = float("nan")
        if ok:
            rc2 = rcond_chol(L) if num.chol_rcond_check else float("nan")
            wrong = wrong and (rc2 < num.min_rcond)
        if (not ok) or wrong:
            if inc > num.max_inc_choldiag:
                raise RuntimeError("[ERROR] Exceed max numerical nugget")
            if num.num_nugget <= 0.0:
                raise RuntimeError("[ERROR] Cannot add numerical nugget which is not strictly positive")
            Xm[np.diag_indices_from(Xm)] += num.num_nugget * (10.0 ** inc)
            inc += 1
            continue
        return L, inc, rc2


def solve_lower(L, B):
    if np.size(B) == 0:  # regmodel "none": F has no column
        return np.array(B, dtype=float, copy=True)
    x, info = lapack.dtrtrs(L, B, lower=1, trans=0)
    assert info == 0
    return x


def chol_block(Cm: np.ndarray, Loo: np.ndarray, num: Numerics | None = None):
    """LinearAlgebra::chol_block (src/lib/LinearAlgebra.cpp:254-299): Cholesky root of C knowing the root Loo of its
    leading block.  Lou = Loo \\ Cou, Luu = safe_chol_lower(Cuu - Lou' Lou) -- the jitter ladder and the rcond test
    act on the Schur complement only; if that ladder is exhausted, a from-scratch safe_chol_lower(C) (:287-293).
    Returns (L, n_jitter, rcond2, used_block)."""
    n, no = Cm.shape[0], Loo.shape[0]
    Cou = Cm[:no, no:]
    Cuu = Cm[no:, no:]
    L = np.zeros((n, n))
    L[:no, :no] = Loo
    Lou = solve_lower(Loo, Cou)
    L[no:, :no] = Lou.T
    try:
        Luu, inc, rc2 = safe_chol_lower(Cuu - Lou.T @ Lou, num)
    except RuntimeError:
        Lf, inc, rc2 = safe_chol_lower(Cm, num)
        return np.tril(Lf), inc, rc2, False
    L[no:, no:] = Luu
    return np.tril(L), inc, rc2, True


def update_chol_cov(X, theta, kernel, alpha, diag, T_old, R_old, num: Numerics | None = None):
    """LinearAlgebra::update_cholCov (src/lib/LinearAlgebra.cpp:206-243): R = [R_old, new columns; ...] with the new
    off-diagonal entries alpha * rho, the diagonal 1 (or `diag`), then chol_block(R, T_old).
    Returns (R, L, n_jitter, rcond2, used_block)."""
    no = T_old.shape[0]
    R = build_R(X, theta, kernel, alpha, diag)
    R[:no, :no] = R_old
    if diag is None:
        np.fill_diagonal(R, 1.0)
    else:
        np.fill_diagonal(R, diag)
    L, inc, rc2, used = chol_block(R, T_old, num)
    return R, L, inc, rc2, used


def solve_upper_Lt(L, B):
    """solve(trimatu(L.t()), B) == L^T \\ B."""
    x, info = lapack.dtrtrs(L, B, lower=1, trans=1)
    assert info == 0
    return x


# --------------------------------------------------------------------------
# KModel / populate_Model  (src/lib/KrigingImpl.cpp:73-125; Kriging.cpp:166-190)
# --------------------------------------------------------------------------
@dataclass
class KModel:
    R: np.ndarray = None
    L: np.ndarray = None
    Rinv: np.ndarray = None
    Fstar: np.ndarray = None
    ystar: np.ndarray = None
    Rstar: np.ndarray = None
    betahat: np.ndarray = None
    Estar: np.ndarray = None
    SSEstar: float = 0.0
    n_jitter: int = 0
    rcond2: float = float("nan")
    used_block_update: bool = False


@dataclass
class KeptModel:
    """The committed model an update extends: m_T, m_R, m_theta and m_alpha | m_sigma2 of the reference, as
    populate_Model's update_eligible test reads them (src/lib/Kriging.cpp:170-188)."""
    T: np.ndarray
    R: np.ndarray
    theta: np.ndarray
    extra: float | None = None


@dataclass
class Problem:
    """Fixed inputs of a fit (members m_X, m_y, m_F, m_noise of the reference)."""
    X: np.ndarray
    y: np.ndarray
    F: np.ndarray
    kernel: str = "gauss"
    noise_model: str = "none"  # none | nugget | hetero
    noise: np.ndarray | None = None
    # estimation flags / fixed values (Kriging.cpp:247-289)
    est_sigma2: bool = True
    est_nugget: bool = True
    est_beta: bool = True
    sigma2: float = 1.0
    nugget: float = 0.0
    alpha: float = 1.0
    num: Numerics = field(default_factory=Numerics)
    kept: KeptModel | None = None  # set by an update: the model whose factor may be block-extended


def populate_model(pb: Problem, theta, extra=None) -> KModel:
    alpha, diag = 1.0, None
    ex = None
    if pb.noise_model == "nugget":
        alpha = ex = pb.alpha if extra is None else extra
    elif pb.noise_model == "hetero":
        s2 = ex = pb.sigma2 if extra is None else extra
        diag = 1.0 + pb.noise / s2
    m = KModel()
    theta = np.asarray(theta, float)
    n = pb.X.shape[0]
    k = pb.kept
    # update_eligible (src/lib/Kriging.cpp:170-188)
    eligible = (k is not None and k.theta.size == theta.size and not np.any(theta - k.theta) and n > k.T.shape[0]
                and (pb.noise_model == "none" or ex == k.extra))
    if eligible:
        m.R, m.L, m.n_jitter, m.rcond2, m.used_block_update = update_chol_cov(pb.X, theta, pb.kernel, alpha, diag,
                                                                              k.T, k.R, pb.num)
    else:
        m.R = build_R(pb.X, theta, pb.kernel, alpha, diag)
        m.L, m.n_jitter, m.rcond2 = safe_chol_lower(m.R, pb.num)
    # inv_sympd (LinearAlgebra.cpp:708-710)
    m.Rinv = solve_upper_Lt(m.L, solve_lower(m.L, np.eye(n)))
    m.Fstar = solve_lower(m.L, pb.F)
    m.ystar = solve_lower(m.L, pb.y)
    if pb.F.shape[1] == 0:  # regmodel "none" (Trend.cpp:39): no trend, Rstar and betahat are empty
        m.Rstar, m.betahat = np.zeros((0, 0)), np.zeros(0)
    else:
        G = m.Fstar.T @ m.Fstar
        Rs, info = lapack.dpotrf(G, lower=0, clean=1)
        if info != 0:
            raise RuntimeError("chol(F*'F*) failed")
        m.Rstar = Rs
        rhs = m.Fstar.T @ m.ystar
        t = lapack.dtrtrs(Rs, rhs, lower=0, trans=1)[0]
        m.betahat = lapack.dtrtrs(Rs, t, lower=0, trans=0)[0]
    resid = pb.y - pb.F @ m.betahat
    m.Estar = solve_lower(m.L, resid)
    m.SSEstar = float(m.Estar @ m.Estar)
    if not pb.est_beta:
        m.betahat = np.zeros(pb.F.shape[1])
    return m


def _pair_grad_sums(pb: Problem, m: KModel, theta, x, weight2, block=512):
    """compute_ll_grad_theta_vecs (KrigingImpl.cpp:855-885):
    t1_k = 2 sum_{i>j} x_i x_j R_ij g_k ; t2_k = -2 sum_{i>j} W2_ij R_ij g_k,
    W2 = Rinv for LL."""
    n, d = pb.X.shape
    t1 = np.zeros(d)
    t2 = np.zeros(d)
    for i0 in range(0, n, block):
        i1 = min(n, i0 + block)
        dx = pb.X[i0:i1, None, :] - pb.X[None, :, :]
        g = dlncorr_dtheta(dx, theta, pb.kernel)
        mask = (np.arange(i0, i1)[:, None] > np.arange(n)[None, :])
        Rg = (m.R[i0:i1, :] * mask)[:, :, None] * g
        t1 += 2.0 * np.einsum("i,ijk,j->k", x[i0:i1], Rg, x)
        t2 -= 2.0 * np.einsum("ij,ijk->k", weight2[i0:i1, :], Rg)
    return t1, t2


def log_likelihood(pb: Problem, gamma, want_grad=True, model_out: list | None = None):
    """Kriging::_logLikelihood (src/lib/Kriging.cpp:214-341).
    gamma = [theta] (none) | [theta, alpha] (nugget) | [theta, sigma2] (hetero)."""
    n, d = pb.X.shape
    gamma = np.asarray(gamma, float)
    theta = gamma[:d]
    if gamma.size > d:
        extra = float(gamma[d])
    else:
        extra = pb.alpha if pb.noise_model == "nugget" else pb.sigma2
    if pb.noise_model == "hetero" and not pb.est_sigma2:
        extra = pb.sigma2
    elif pb.noise_model == "nugget" and not pb.est_sigma2 and not pb.est_nugget:
        extra = pb.sigma2 / (pb.sigma2 + pb.nugget)
    m = populate_model(pb, theta, extra)
    if model_out is not None:
        model_out.append(m)
    sumlog = float(np.sum(np.log(np.diag(m.L))))
    if pb.noise_model == "nugget":
        a = extra
        s2, nug = pb.sigma2, pb.nugget
        if pb.est_sigma2:
            if pb.est_nugget:
                var = m.SSEstar / n
                s2, nug = a * var, (1.0 - a) * var
            else:
                s2 = pb.nugget * a / (1.0 - a)
        else:
            if pb.est_nugget:
                nug = pb.sigma2 * (1.0 - a) / a
            else:
                a = pb.sigma2 / (pb.sigma2 + pb.nugget)
        tv = s2 + nug
        ll = -0.5 * (n * math.log(2 * math.pi * tv) + 2 * sumlog + m.SSEstar / tv)
        s2g = tv
    elif pb.noise_model == "hetero":
        s2 = extra if pb.est_sigma2 else pb.sigma2
        ll = -0.5 * (n * math.log(2 * math.pi * s2) + 2 * sumlog + m.SSEstar / s2)
        s2g = s2
    else:
        if pb.est_sigma2:
            s2g = m.SSEstar / n
            ll = -0.5 * (n * math.log(2 * math.pi * s2g) + 2 * sumlog + n)
        else:
            s2g = pb.sigma2
            ll = -0.5 * (n * math.log(2 * math.pi * s2g) + 2 * sumlog + m.SSEstar / s2g)
    if not want_grad:
        return ll, None
    x = solve_upper_Lt(m.L, m.Estar)
    t1, t2 = _pair_grad_sums(pb, m, theta, x, m.Rinv)
    grad = np.zeros(gamma.size)
    grad[:d] = (t1 / s2g + t2) / 2.0
    if gamma.size > d:
        if pb.noise_model == "nugget":
            a = extra
            if pb.est_sigma2 and pb.est_nugget:
                dR = m.R / a
                np.fill_diagonal(dR, 0.0)
                term1 = -float(x @ dR @ x) / s2g
                term2 = float(np.sum(m.Rinv * dR))
                grad[d] = -0.5 * (term1 + term2)
            elif pb.est_sigma2 and not pb.est_nugget:
                dR = m.R / a
                np.fill_diagonal(dR, 1.0)
                term1 = -float(x @ dR @ x) / (s2g * s2g)
                term2 = float(np.sum((m.Rinv / s2g) * dR))
                grad[d] = -0.5 * (term1 + term2) * pb.nugget / (1.0 - a) / (1.0 - a)
            else:
                grad[d] = 0.0
        elif pb.noise_model == "hetero":
            if not pb.est_sigma2:
                grad[d] = 0.0
            else:
                s2 = extra
                s2sq = s2 * s2
                nR = float(pb.noise @ np.diag(m.Rinv))
                nx2 = float(pb.noise @ (x * x))
                grad[d] = -0.5 * (n / s2 - nR / s2sq + nx2 / (s2sq * s2) - m.SSEstar / s2sq)
    return ll, grad


def leave_one_out(pb: Problem, theta, want_grad=True, return_vec=False):
    """Kriging::_leaveOneOut (src/lib/Kriging.cpp:353-468); NoiseModel::None only."""
    n, d = pb.X.shape
    theta = np.asarray(theta, float)
    m = populate_model(pb, theta)
    Linv = solve_lower(m.L, np.eye(n))
    By = Linv.T @ m.Estar
    Q, _ = np.linalg.qr(m.Fstar)  # qr_econ (LinearAlgebra.cpp:716)
    A = Q.T @ Linv
    B = Linv.T @ Linv - A.T @ A
    s2loo = 1.0 / np.diag(B)
    err = s2loo * By
    loo = float(np.sum(err * err)) / n
    out_vec = (pb.y - err, np.sqrt(s2loo)) if return_vec else None
    if not want_grad:
        return (loo, None, out_vec) if return_vec else (loo, None)
    grad = np.zeros(d)
    dxfull = _pair_diffs(pb.X)
    g = dlncorr_dtheta(dxfull, theta, pb.kernel)
    for k in range(d):
        G = m.R * g[:, :, k]
        np.fill_diagonal(G, 0.0)
        # diagABA (LinearAlgebra.cpp:428-433)
        D = np.triu(2 * G)
        np.fill_diagonal(D, np.diag(G))
        diagdB = -np.sum((B @ D) * B, axis=1)
        ds2 = -s2loo * s2loo * diagdB
        derr = ds2 * By - s2loo * (B @ (G @ By))
        grad[k] = 2.0 * float(err @ derr) / n
    return (loo, grad, out_vec) if return_vec else (loo, grad)


def log_marg_post(pb: Problem, gamma, want_grad=True):
    """Kriging::_logMargPost (src/lib/Kriging.cpp:488-648) + compute_lmp_theta_ans
    (src/lib/KrigingImpl.cpp:887-921).  est_sigma2 branch only (the fixed-sigma2
    branch of the reference is a forward finite difference of this value)."""
    n, d = pb.X.shape
    p = pb.F.shape[1]
    gamma = np.asarray(gamma, float)
    theta = gamma[:d]
    alpha = float(gamma[d]) if pb.noise_model == "nugget" else pb.alpha
    m = populate_model(pb, theta, alpha if pb.noise_model == "nugget" else None)
    Rinv_X = solve_upper_Lt(m.L, m.Fstar)
    XtRX = pb.F.T @ Rinv_X
    LX, _, _ = safe_chol_lower(XtRX, pb.num)
    P = Rinv_X @ lapack.dtrtrs(LX, lapack.dtrtrs(LX, Rinv_X.T, lower=1, trans=0)[0], lower=1, trans=1)[0]
    yt_Rinv = solve_upper_Lt(m.L, m.ystar)
    S2 = float(yt_Rinv @ pb.y - pb.y @ P @ pb.y)
    if pb.noise_model == "nugget":
        if pb.est_sigma2 and pb.est_nugget:
            sigma2 = S2 / (n - p)
        elif pb.est_sigma2 or pb.est_nugget:
            sigma2 = pb.sigma2 / alpha
        else:
            sigma2 = pb.sigma2 + pb.nugget
    elif pb.est_sigma2:
        sigma2 = S2 / (n - p)
    else:
        sigma2 = pb.sigma2
    logS2 = math.log(sigma2 * (n - p))
    lml = -float(np.sum(np.log(np.diag(m.L)))) - float(np.sum(np.log(np.diag(LX)))) - (n - p) / 2.0 * logS2
    a = 0.2
    b = 1.0 / (n ** (1.0 / d)) * (a + d)
    CL = (pb.X.max(axis=0) - pb.X.min(axis=0)) / (n ** (1.0 / d))
    nugget_ratio = (1.0 - alpha) / alpha if pb.noise_model == "nugget" else 0.0
    t = float(np.sum(CL / theta)) + nugget_ratio
    lprior = -b * t + a * math.log(t)
    val = lml + lprior
    if not want_grad:
        return val, None
    grad = np.zeros(gamma.size)
    Qo = yt_Rinv - P @ pb.y
    dxfull = _pair_diffs(pb.X)
    g = dlncorr_dtheta(dxfull, theta, pb.kernel)

    def ans_for(G):
        Wb = solve_upper_Lt(m.L, solve_lower(m.L, G)).T - G @ P
        return -float(np.trace(Wb)) / 2.0 + float(pb.y @ Wb.T @ Qo) / (2.0 * sigma2)

    for k in range(d):
        G = m.R * g[:, :, k]
        np.fill_diagonal(G, 0.0)
        grad[k] = ans_for(G)
    grad[:d] -= (a * CL / t - b * CL) / (theta * theta)
    if pb.noise_model == "nugget":
        if pb.est_sigma2 or pb.est_nugget:
            G = m.R / alpha
            np.fill_diagonal(G, 0.0)
            grad[d] = ans_for(G) - (a / t - b) / (alpha ** 2.0)
        else:
            grad[d] = 0.0
    return val, grad


# --------------------------------------------------------------------------
# Trend basis (src/lib/Trend.cpp:34-93) and theta bounds (src/lib/Optim.cpp:179-209)
# --------------------------------------------------------------------------
def regression_matrix(regmodel: str, X: np.ndarray) -> np.ndarray:
    """Trend::regressionModelMatrix (src/lib/Trend.cpp:34-93), same column order."""
    n, d = X.shape
    if regmodel == "none":
        return np.ones((n, 0))
    cols = [np.ones(n)]
    if regmodel == "constant":
        pass
    elif regmodel == "linear":
        cols += [X[:, i] for i in range(d)]
    elif regmodel == "interactive":
        for i in range(d):
            cols.append(X[:, i])
            for j in range(i):
                cols.append(X[:, i] * X[:, j])
    elif regmodel == "quadratic":
        for i in range(d):
            cols.append(X[:, i])
            for j in range(i + 1):
                cols.append(X[:, i] * X[:, j])
    else:
        raise ValueError(regmodel)
    return np.column_stack(cols)


def theta_bounds(X, y, lower_factor=0.02, upper_factor=10.0, heuristic=True):
    """Optim::theta_bounds (src/lib/Optim.cpp:179-209) with m_maxdX
    (src/lib/KrigingImpl.cpp:807).  Sums run over ALL ordered pairs."""
    n, d = X.shape
    dX = _pair_diffs(X)  # [i, j, k]
    maxdX = np.max(np.abs(dX.reshape(-1, d)), axis=0)
    lower = lower_factor * maxdX
    upper = upper_factor * maxdX
    if heuristic:
        dy2 = (y[:, None] - y[None, :]) ** 2
        dX2 = np.sum(dX * dX, axis=2)
        with np.errstate(divide="ignore", invalid="ignore"):
            w = dy2 / dX2
        w = np.where(np.isnan(w), 0.0, w)  # replace(nan, 0) only (Optim.cpp:198)
        wsum = float(np.sum(w))
        if wsum > 0.0:
            w = w / wsum
            steep = np.einsum("ij,ijk->k", w, np.abs(dX))
            lower = np.maximum(lower, lower_factor * steep)
            lower = np.minimum(lower, upper)
            upper = np.maximum(lower, upper)
    return lower, upper


# --------------------------------------------------------------------------
# predict mean / stdev (src/lib/KrigingImpl.cpp:145-243), constant-sigma2 form
# --------------------------------------------------------------------------
def sigma2_variogram(X, y):
    """sigma2 bounds of NoiseModel::Heterogeneous (src/lib/Kriging.cpp:1784-1797): dX2 = sum(dX % dX, 0) over all n^2
    ordered pairs (Armadillo's arrayops::accumulate order: two interleaved accumulators), median as
    op_median::direct_median (upper middle element + robust_mean with the lower one), then
    0.5 * mean(dy2[dX2 >= median])."""
    X = np.asarray(X, float)
    y = np.asarray(y, float).ravel()
    n, d = X.shape
    dX2 = np.empty((n, n))
    for i0 in range(0, n, 1024):
        sq = (X[i0:i0 + 1024, None, :] - X[None, :, :]) ** 2
        acc1 = np.zeros(sq.shape[:2])
        acc2 = np.zeros(sq.shape[:2])
        for k in range(0, d - 1, 2):
            acc1 += sq[:, :, k]
            acc2 += sq[:, :, k + 1]
        if d % 2 == 1:
            acc1 += sq[:, :, d - 1]
        dX2[i0:i0 + 1024] = acc1 + acc2
    flat = np.sort(dX2.ravel())
    half = flat.size // 2
    val1 = flat[half]
    med = val1 + (flat[half - 1] - val1) / 2.0 if flat.size % 2 == 0 else val1
    dy2 = (y[:, None] - y[None, :]) ** 2
    return 0.5 * float(np.mean(dy2[dX2 >= med]))


def predict(pb: Problem, theta, sigma2, Xn, Fn, m: KModel | None = None, fixed_beta=None):
    """KrigingImpl::predict_impl (KrigingImpl.cpp:145-243).  fixed_beta: the caller's trend coefficients when
    m_est_beta is false -- then m_z = ystar - M beta (Kriging.cpp:2168-2172) instead of Estar."""
    if m is None:
        m = populate_model(pb, theta)
    dx = pb.X[:, None, :] - Xn[None, :, :]
    R_on = corr_from_dx(dx, np.asarray(theta, float), pb.kernel)
    if pb.noise_model == "nugget":
        R_on = R_on * pb.alpha
    # coincident points: exactly 1, no R_on_factor (KrigingImpl.cpp:199-202, dij.is_zero(eps))
    R_on = np.where(np.all(np.abs(dx) <= np.finfo(float).eps, axis=-1), 1.0, R_on)
    Rstar_on = solve_lower(m.L, R_on)
    if fixed_beta is None:
        z, beta = m.Estar, m.betahat
    else:
        beta = np.asarray(fixed_beta, float).ravel()
        z = m.ystar - m.Fstar @ beta
    mean = Fn @ beta + Rstar_on.T @ z
    if m.Rstar.size == 0:
        Ecirc = np.zeros((Xn.shape[0], 0))
    else:
        Ecirc = lapack.dtrtrs(m.Rstar, (Fn - Rstar_on.T @ m.Fstar).T, lower=0, trans=1)[0].T
    var = 1.0 - np.sum(Rstar_on * Rstar_on, axis=0) + np.sum(Ecirc * Ecirc, axis=1)
    var = np.maximum(var, 0.0)
    return mean, np.sqrt(var * sigma2)
