"""tests/golden/make_golden_cfg4_oracle.py -- fixture generator (BUILD container, CPU, ~15 minutes, ~30 GB).

BASELINE config 4 (NuggetKriging('matern3_2') LL + gradient, n = 40000, d = 8) cannot be run by the reference: its
m_dX alone is 8 d n^2 = 102 GB and 1.28e10 elements exceed Armadillo's 32-bit uword of the default build
(CMakeLists.txt:285, ARMA_32BIT_WORD).  This script evaluates the same formulas with numpy / LAPACK BLOCKWISE (blocked
Cholesky, blocked triangular inverse, R^-1 = L^-T L^-1 by blocks, pair sums regenerated blockwise; two n x n buffers) -- a memory-lean restatement of
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


def _blocks(n, nbk):
    e = list(range(0, n, nbk)) + [n]
    return [(e[i], e[i + 1]) for i in range(len(e) - 1)]


def chol_blocked(A, nbk):
    """In-place lower Cholesky of the lower triangle of A, right-looking on nbk x nbk blocks.  Every LAPACK call sees
    a contiguous block of at most nbk^2 elements (the LP64 LAPACK behind scipy faults on one 40000 x 40000 call);
    the O(n^3) work is numpy matmul on views."""
    B = _blocks(A.shape[0], nbk)
    for k, (k0, k1) in enumerate(B):
        Lkk, info = lapack.dpotrf(np.asfortranarray(A[k0:k1, k0:k1]), lower=1, clean=1)
        assert info == 0
        A[k0:k1, k0:k1] = Lkk
        for (i0, i1) in B[k + 1:]:
            # A_ik <- A_ik L_kk^-T   ==   (L_kk^-1 A_ik^T)^T
            A[i0:i1, k0:k1] = solve_triangular(Lkk, np.asfortranarray(A[i0:i1, k0:k1].T), lower=True, check_finite=False).T
        for bi, (i0, i1) in enumerate(B[k + 1:], k + 1):
            for (j0, j1) in B[k + 1:bi + 1]:
                A[i0:i1, j0:j1] -= A[i0:i1, k0:k1] @ A[j0:j1, k0:k1].T


def trsv_blocked(A, b, nbk, trans=False):
    """L x = b (trans=False) or L^T x = b (trans=True), L = lower triangle of A, blockwise."""
    B = _blocks(A.shape[0], nbk)
    x = np.array(b, dtype=float, copy=True)
    order = B[::-1] if trans else B
    for (k0, k1) in order:
        Lkk = np.asfortranarray(A[k0:k1, k0:k1])
        x[k0:k1] = solve_triangular(Lkk, x[k0:k1], lower=True, trans="T" if trans else "N", check_finite=False)
        if trans:
            x[:k0] -= A[k0:k1, :k0].T @ x[k0:k1]
        else:
            x[k1:] -= A[k1:, k0:k1] @ x[k0:k1]
    return x


def inv_lower_blocked(A, nbk):
    """X = L^-1 (lower) into a new array, blockwise: X_kk = L_kk^-1, X_ik = -X_ii sum_{k<=j<i} L_ij X_jk."""
    n = A.shape[0]
    B = _blocks(n, nbk)
    X = np.zeros((n, n), order="F")
    for (k0, k1) in B:
        X[k0:k1, k0:k1] = solve_triangular(np.asfortranarray(A[k0:k1, k0:k1]), np.eye(k1 - k0), lower=True, check_finite=False)
    for k, (k0, k1) in enumerate(B):
        for bi, (i0, i1) in enumerate(B[k + 1:], k + 1):
            T = np.zeros((i1 - i0, k1 - k0))
            for (j0, j1) in B[k:bi]:
                T += A[i0:i1, j0:j1] @ X[j0:j1, k0:k1]
            X[i0:i1, k0:k1] = -(X[i0:i1, i0:i1] @ T)
    return X


def ll_grad_nugget_m32_lean(X, y, theta, alpha, blk=400, nbk=8000):
    n, d = X.shape
    theta = np.asarray(theta, float)
    A = np.empty((n, n), order="F")
    for i0 in range(0, n, blk):  # lower part, row blocks
        i1 = min(n, i0 + blk)
        s = SQRT3 * np.abs(X[i0:i1, None, :] - X[None, :i1, :]) / theta
        A[i0:i1, :i1] = alpha * np.exp(-np.sum(s - np.log1p(s), axis=2))
    A[np.arange(n), np.arange(n)] = 1.0
    chol_blocked(A, nbk)
    sumlog = float(np.sum(np.log(np.diag(A))))
    F = np.ones(n)
    Fstar = trsv_blocked(A, F, nbk)
    ystar = trsv_blocked(A, y, nbk)
    beta = float(Fstar @ ystar) / float(Fstar @ Fstar)  # constant trend: Rstar^-1 Rstar^-T F*' y*
    Estar = trsv_blocked(A, y - F * beta, nbk)
    SSE = float(Estar @ Estar)
    x = trsv_blocked(A, Estar, nbk, trans=True)
    tv = SSE / n
    ll = -0.5 * (n * np.log(2 * np.pi * tv) + 2 * sumlog + n)
    Li = inv_lower_blocked(A, nbk)
    # rcond_1(L) exactly (it can only be below dtrcon's estimate: accepted here means accepted by safe_chol_lower)
    normL = max(float(np.sum(np.abs(A[j:, j]))) for j in range(n))
    normLi = max(float(np.sum(np.abs(Li[j:, j]))) for j in range(n))
    rc = 1.0 / (normL * normLi)
    assert rc * rc >= 1e-18, rc
    # A <- R^-1 = L^-T L^-1 (lower blocks), blockwise
    B = _blocks(n, nbk)
    for bi, (i0, i1) in enumerate(B):
        for (j0, j1) in B[:bi + 1]:
            acc = np.zeros((i1 - i0, j1 - j0))
            for (k0, k1) in B[bi:]:
                acc += Li[k0:k1, i0:i1].T @ Li[k0:k1, j0:j1]
            A[i0:i1, j0:j1] = acc
    del Li
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
    v1, g1, _ = ll_grad_nugget_m32_lean(Xs, ys, gam[:8], 0.9, nbk=400)  # 4 x 4 blocks, like the full-size run
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
    print("grad", [float(t) for t in g], flush=True)
    if n != 40000:
        return  # smaller sizes are a check of this script only; the fixture holds the full-size case
    doc = json.load(open(OUT))
    doc["cases"]["cfg4-oracle"] = dict(
        n=n, d=d, seed=123, kernel="matern3_2", noise_model="nugget", objective="LL", theta=theta, extra=alpha,
        value=v, grad=[float(t) for t in g], eval_s=wall, threads=len(os.sched_getaffinity(0)),
        y_sum=float(np.sum(y)), X_sum=float(np.sum(X)), rcond2=rc2,
        source="oracle (tests/golden/make_golden_cfg4_oracle.py: numpy / LAPACK in place; the reference cannot "
               "allocate this size)")
    json.dump(doc, open(OUT, "w"), indent=1)


if __name__ == "__main__":
    main()
