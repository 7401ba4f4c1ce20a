"""NestedKriging's sub-model fits as a batched many-small-n workload.

The reference fits the p sub-models of a NestedKriging one after the other on the CPU
(src/lib/NestedKriging.cpp:199-275: `km->fit(m_y(idx), m_X.rows(idx), regmodel, false, optim, objective, parameters)`
per group, "kept sequential for now"), then replaces their hyper-parameters by a weighted common prior and re-fits
every sub-model in closed form (unify_hyperparameters, :277-331).  A sub-model is a mid-size factorisation
(n / p rows) whose panel chain is latency-bound and cannot fill 148 SMs: here the p fits run concurrently, one engine
handle (own workspaces, own CUDA streams) and one host thread per fit in flight -- the same batched-occupancy
mechanism as the concurrent multistart rows of Kriging.fit (BASELINE cfg 5).  Every sub-model is the one a sequential loop produces up to
the rounding of the triangular sweeps (overlapping evaluations use the launch-chain sweep kernels: engine.cu,
SweepGate); `concurrent=1` gives the exact sequential loop.  Across GPUs the groups shard like multistart rows
(group g on rank g mod world); only the fitted hyper-parameters (d + 2 doubles per group) are exchanged.

Out of scope (reference control plane): the k-means partition (arma::kmeans with arma's own RNG), the PoE / BCM /
NK aggregations of predict, warped sub-models, the LLVecchia-unified path.
"""
from __future__ import annotations

import numpy as np

from .kriging import Kriging


def random_partition(n: int, nb_groups: int, seed: int = 123):
    """Balanced random split (the reference's Partition::Random / k-means fallback, NestedKriging.cpp:155-159:
    assignment(perm(i)) = i % nb_groups) -- with numpy's PCG64 permutation, not arma's RNG stream."""
    perm = np.random.Generator(np.random.PCG64(seed)).permutation(n)
    assignment = np.empty(n, dtype=np.int64)
    assignment[perm] = np.arange(n) % nb_groups
    return [np.flatnonzero(assignment == g) for g in range(nb_groups)]


def check_groups(n: int, d: int, groups):
    """fit()'s argument checks (NestedKriging.cpp:177-180) on an explicit partition."""
    p = len(groups)
    if p < 1 or p > n // (d + 2):
        raise ValueError("nb_groups should be in [1, n/(d+2)]")
    seen = np.concatenate([np.asarray(g, dtype=np.int64) for g in groups])
    if seen.size != n or np.unique(seen).size != n or seen.min() != 0 or seen.max() != n - 1:
        raise ValueError("groups must partition range(n)")


def fit_submodels(y, X, groups, kernel, regmodel="constant", optim="BFGS", objective="LL", parameters=None, *,
                  concurrent: int | None = None, device: int | None = None, comm=None, backend_factory=None):
    """Step 1 of NestedKriging::fit (NestedKriging.cpp:262-270): one Kriging fit per group, `concurrent` of them in
    flight on the device at a time.  Returns {g: fitted Kriging} for the groups of this rank (all groups when comm is
    None).  Each fit is Kriging.fit(y[idx], X[idx], regmodel, normalize=False, optim, objective, parameters)."""
    from concurrent.futures import ThreadPoolExecutor

    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    if y.size != X.shape[0]:
        raise ValueError("y and X should have the same number of rows")
    check_groups(X.shape[0], X.shape[1], groups)
    mine = list(range(len(groups))) if comm is None else comm.my_starts(len(groups))
    dev = device if device is not None else (comm.device if comm is not None else 0)
    if concurrent is None:
        concurrent = 4
    concurrent = max(1, min(int(concurrent), len(mine) or 1))

    def fit_one(g):
        idx = np.asarray(groups[g], dtype=np.int64)
        k = Kriging(kernel, device=dev, backend_factory=backend_factory, concurrent_starts=1,
                    concurrent_handle=concurrent > 1)
        k.fit(y[idx], X[idx], regmodel, False, optim, objective, parameters)
        return g, k

    if concurrent == 1:
        return dict(fit_one(g) for g in mine)
    with ThreadPoolExecutor(max_workers=concurrent) as ex:
        return dict(ex.map(fit_one, mine))


def gather_hyperparameters(models, groups, comm=None):
    """theta (p x d), sigma2 (p), beta0 (p) of all groups on every rank: the only data that crosses ranks."""
    p = len(groups)
    d = next(iter(models.values())).theta().size if models else 0
    if comm is not None:
        d = int(comm.allreduce_max(d))
    tab = np.zeros((p, d + 2))
    for g, k in models.items():
        tab[g, :d] = k.theta()
        tab[g, d] = k.sigma2()
        tab[g, d + 1] = k.beta()[0]
    if comm is not None:
        tab = comm.allreduce_sum(tab)  # every group is owned by exactly one rank
    return tab[:, :d], tab[:, d], tab[:, d + 1]


def unify_hyperparameters(models, groups, y, X, regmodel="constant", objective="LL", comm=None, concurrent=None):
    """Step 2 of NestedKriging::fit, plain path (NestedKriging.cpp:304-331): weighted geometric mean of the thetas,
    weighted means of sigma2 and (constant trend) beta0 with weights n_g / n, then every sub-model is re-fitted in
    closed form (optim = none) on that common prior.  Returns (theta, sigma2, beta0); `models` is updated in place."""
    from concurrent.futures import ThreadPoolExecutor

    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n = float(X.shape[0])
    thetas, sigma2s, beta0s = gather_hyperparameters(models, groups, comm)
    log_theta = np.zeros(thetas.shape[1])
    sigma2 = 0.0
    beta0 = 0.0
    for g in range(len(groups)):
        w = float(len(groups[g])) / n
        log_theta += w * np.log(thetas[g])
        sigma2 += w * sigma2s[g]
        if regmodel == "constant":
            beta0 += w * beta0s[g]
    theta = np.exp(log_theta)
    fixed = dict(theta=theta[None, :], is_theta_estim=False, sigma2=sigma2, is_sigma2_estim=False)
    if regmodel == "constant":
        fixed.update(beta=np.array([beta0]), is_beta_estim=False)

    def refit(item):
        g, k = item
        idx = np.asarray(groups[g], dtype=np.int64)
        k.fit(y[idx], X[idx], regmodel, False, "none", objective, fixed)

    items = list(models.items())
    nthreads = max(1, min(int(concurrent or 8), len(items) or 1))
    if nthreads == 1:
        for it in items:
            refit(it)
    else:
        with ThreadPoolExecutor(max_workers=nthreads) as ex:
            list(ex.map(refit, items))
    return theta, sigma2, beta0
