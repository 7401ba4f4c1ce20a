"""tests/golden/make_golden_cfg4_oracle.py -- fixture generator (BUILD container, CPU, ~15 minutes, ~30 GB).

BASELINE config 4 (NuggetKriging('matern3_2') LL + gradient, n = 40000, d = 8) cannot be run by the reference: its
m_dX alone is 8 d n^2 = 102 GB and 1.28e10 elements exceed Armadillo's 32-bit uword of the default build
(CMakeLists.txt:285, ARMA_32BIT_WORD).  This script evaluates the same formulas with numpy / LAPACK IN PLACE (one
n x n buffer: R -> dpotrf -> L -> dpotri -> R^-1, pair sums regenerated blockwise) -- a memory-lean restatement of
oracle/kriging_oracle.py:log_likelihood for the Nugget model with everything estimated
(Kriging.cpp:243-339: total variance SSE/n, theta gradient (t1/tv + t2)/2, alpha gradient :308-326).  Before the big
run it is checked against oracle/kriging_oracle.py (itself pinned on the reference) at n = 1500 to 1e-11.

The result goes to tests/golden/refgen_fullsize.json as case "cfg4-oracle" with source = "oracle" (NOT the reference).
Usage: python tests/golden/make_golden_cfg4_oracle.py [n]
"""
import json
import os
import sys
import time

import numpy as np
from scipy.linalg import lapack, solve_triangular

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import kriging_oracle as ko  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "refgen_fullsize.json")
SQRT3 = np.sqrt(3.0)


def synth(n, d, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((n, d))
    y = np.sin(3.0 * X[:, 0]) + np.sum(X * X, axis=1) + 0.05 * rng.standard_normal(n)
    return X, y


def m32_block(Xi, Xj, theta):
    """rho (matern3_2) and theta_k dln rho/dtheta_k for all pairs of the two row blocks."""
    s = SQRT3 * np.abs(Xi[:, None, :] - Xj[None, :, :]) / theta  # (bi, bj, d)
    rho = np.exp(-np.sum(s - np.log1p(s), axis=2))               # Covariance.cpp:104-116
    g = (s * s) / (1.0 + s) / theta                              # Covariance.cpp:118-130 : dln rho / dtheta_k
    return rho, g


def ll_grad_nugget_m32_lean(X, y, theta, alpha, blk=400):
    n, d = X.shape
    theta = np.asarray(theta, float)
    A = np.empty((n, n), order="F")
    for i0 in range(0, n, blk):  # lower part, row blocks (dpotrf('L') reads the lower triangle only)
        i1 = min(n, i0 + blk)
        s = SQRT3 * np.abs(X[i0:i1, None, :] - X[None, :i1, :]) / theta
        A[i0:i1, :i1] = alpha * np.exp(-np.sum(s - np.log1p(s), axis=2))
    A[np.arange(n), np.arange(n)] = 1.0
    c, info = lapack.dpotrf(A, lower=1, overwrite_a=1, clean=0)
    assert info == 0 and np.shares_memory(c, A)
    rc, info = lapack.dtrcon(A, norm="1", uplo="L", diag="N")
    assert info == 0 and rc * rc >= 1e-18, rc  # safe_chol_lower accepts without jitter
    sumlog = float(np.sum(np.log(np.diag(A))))
    F = np.ones((n, 1))
    Fstar = solve_triangular(A, F, lower=True, check_finite=False)
    ystar = solve_triangular(A, y, lower=True, check_finite=False)
    Rstar = np.linalg.cholesky(Fstar.T @ Fstar).T
    beta = np.linalg.solve(Rstar, np.linalg.solve(Rstar.T, Fstar.T @ ystar))
    Estar = solve_triangular(A, y - F @ beta, lower=True, check_finite=False)
    SSE = float(Estar @ Estar)
    x = solve_triangular(A, Estar, lower=True, trans="T", check_finite=False)
    tv = SSE / n
    ll = -0.5 * (n * np.log(2 * np.pi * tv) + 2 * sumlog + n)
    c, info = lapack.dpotri(A, lower=1, overwrite_c=1)  # A <- R^-1 (lower)
    assert info == 0 and np.shares_memory(c, A)
    t1 = np.zeros(d)
    t2 = np.zeros(d)
    xRx = 0.0
    RiR = 0.0
    for i0 in range(0, n, blk):
        i1 = min(n, i0 + blk)
        rho, g = m32_block(X[i0:i1], X[:i1], theta)
        R = alpha * rho
        R[np.arange(i1 - i0)[:, None] + i0 <= np.arange(i1)[None, :]] = 0.0  # strictly lower pairs i > j only
        w1 = (x[i0:i1, None] * x[None, :i1]) * R
        w2 = A[i0:i1, :i1] * R
        t1 += 2.0 * np.einsum("ij,ijk->k", w1, g)
        t2 += -2.0 * np.einsum("ij,ijk->k", w2, g)
        xRx += 2.0 * float(np.sum(w1))
        RiR += 2.0 * float(np.sum(w2))
    grad = np.empty(d + 1)
    grad[:d] = (t1 / tv + t2) / 2.0
    grad[d] = -0.5 * (-(xRx / alpha) / tv + RiR / alpha)
    return ll, grad, rc * rc


def main():
    # ---- self-check against the pinned oracle ----
    Xs, ys = synth(1500, 8, 7)
    pb = ko.Problem(X=Xs, y=ys, F=np.ones((1500, 1)), kernel="matern3_2", noise_model="nugget")
    gam = np.append(np.full(8, 0.6), 0.9)
    v0, g0 = ko.log_likelihood(pb, gam, True)
    v1, g1, _ = ll_grad_nugget_m32_lean(Xs, ys, gam[:8], 0.9)
    assert abs(v0 - v1) <= 1e-11 * abs(v0), (v0, v1)
    assert np.linalg.norm(g0 - g1) <= 1e-11 * np.linalg.norm(g0), (g0, g1)
    print("self-check vs oracle/kriging_oracle.py at n=1500: value relerr %.1e, gradient relerr %.1e" % (
        abs(v0 - v1) / abs(v0), np.linalg.norm(g0 - g1) / np.linalg.norm(g0)), flush=True)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40000
    d, theta, alpha = 8, 0.6, 0.9
    X, y = synth(n, d, 123)
    t0 = time.time()
    v, g, rc2 = ll_grad_nugget_m32_lean(X, y, np.full(d, theta), alpha)
    wall = time.time() - t0
    print(n, v, g, "rcond^2", rc2, "wall", wall, flush=True)
    doc = json.load(open(OUT))
    doc["cases"]["cfg4-oracle" if n == 40000 else f"cfg4-oracle-n{n}"] = dict(
        n=n, d=d, seed=123, kernel="matern3_2", noise_model="nugget", objective="LL", theta=theta, extra=alpha,
        value=v, grad=[float(t) for t in g], eval_s=wall, threads=len(os.sched_getaffinity(0)),
        y_sum=float(np.sum(y)), X_sum=float(np.sum(X)), rcond2=rc2,
        source="oracle (tests/golden/make_golden_cfg4_oracle.py: numpy / LAPACK in place; the reference cannot "
               "allocate this size)")
    json.dump(doc, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
