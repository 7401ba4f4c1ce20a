"""NestedKriging's sub-model fits as a batched workload (libkriging_b200/nested.py).
CPU: the batch (several fits in flight) equals the sequential loop of the reference (NestedKriging.cpp:262-270) bit for
bit, the common prior follows unify_hyperparameters (:304-331), groups shard over ranks (gloo, world_size 2).
GPU (-m gpu): the same on the device engine, with the batch's wall time beside the sequential loop's."""
import os
import socket
import time

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from libkriging_b200 import nested
from libkriging_b200.kriging import Kriging
from tests.oracle_backend import OracleBackend
from tests.util import relerr, synth


def _data(n=240, d=2, seed=31):
    X, y, _ = synth(n, d, seed, "smooth")
    return X, y


def test_random_partition_is_balanced_and_checked():
    g = nested.random_partition(103, 5, seed=7)
    assert sorted(np.concatenate(g).tolist()) == list(range(103))
    assert max(len(x) for x in g) - min(len(x) for x in g) <= 1
    with pytest.raises(ValueError, match="nb_groups"):
        nested.check_groups(20, 3, nested.random_partition(20, 5))
    with pytest.raises(ValueError, match="partition"):
        nested.check_groups(103, 2, g[:-1])


def test_batched_submodel_fits_equal_sequential_reference_loop():
    X, y = _data()
    groups = nested.random_partition(len(y), 4, seed=3)
    prm = {"theta": np.full((1, 2), 0.5)}
    seq = {}
    for g, idx in enumerate(groups):  # the reference's loop, one Kriging::fit per group
        k = Kriging("matern5_2", backend_factory=OracleBackend)
        k.fit(y[idx], X[idx], "constant", False, "BFGS", "LL", prm)
        seq[g] = k
    bat = nested.fit_submodels(y, X, groups, "matern5_2", optim="BFGS", objective="LL", parameters=prm, concurrent=3,
                               backend_factory=OracleBackend)
    assert sorted(bat) == [0, 1, 2, 3]
    for g in range(4):
        assert np.array_equal(bat[g].theta(), seq[g].theta())
        assert bat[g].sigma2() == seq[g].sigma2() and np.array_equal(bat[g].beta(), seq[g].beta())


def test_unify_hyperparameters_matches_reference_formulas():
    X, y = _data()
    groups = nested.random_partition(len(y), 3, seed=5)
    prm = {"theta": np.full((1, 2), 0.5)}
    models = nested.fit_submodels(y, X, groups, "matern5_2", parameters=prm, concurrent=2, backend_factory=OracleBackend)
    th = np.array([models[g].theta() for g in range(3)])
    s2 = np.array([models[g].sigma2() for g in range(3)])
    b0 = np.array([models[g].beta()[0] for g in range(3)])
    w = np.array([len(g) for g in groups]) / len(y)
    theta, sigma2, beta0 = nested.unify_hyperparameters(models, groups, y, X, concurrent=2)
    assert relerr(theta, np.exp(w @ np.log(th))) < 1e-14
    assert relerr(sigma2, w @ s2) < 1e-14 and relerr(beta0, w @ b0) < 1e-14
    for g in range(3):  # closed-form re-fit on the common prior (optim = none, everything fixed)
        assert np.array_equal(models[g].theta(), theta)
        assert models[g].sigma2() == sigma2 and models[g].beta()[0] == beta0
        mean, sd = models[g].predict(X[groups[g]][:5], True)
        assert np.allclose(mean, y[groups[g]][:5], atol=0.3)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from libkriging_b200.parallel import MultistartComm
        comm = MultistartComm()
        X, y = _data()
        groups = nested.random_partition(len(y), 4, seed=3)
        prm = {"theta": np.full((1, 2), 0.5)}
        models = nested.fit_submodels(y, X, groups, "matern5_2", parameters=prm, concurrent=2, comm=comm,
                                      backend_factory=OracleBackend)
        theta, sigma2, beta0 = nested.unify_hyperparameters(models, groups, y, X, comm=comm)
        q.put((rank, sorted(models), theta.tolist(), sigma2, beta0))
    finally:
        dist.destroy_process_group()


def test_groups_shard_over_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    out = sorted(q.get(timeout=10) for _ in range(2))
    X, y = _data()
    groups = nested.random_partition(len(y), 4, seed=3)
    prm = {"theta": np.full((1, 2), 0.5)}
    models = nested.fit_submodels(y, X, groups, "matern5_2", parameters=prm, concurrent=1, backend_factory=OracleBackend)
    theta, sigma2, beta0 = nested.unify_hyperparameters(models, groups, y, X)
    assert out[0][1] == [0, 2] and out[1][1] == [1, 3]
    for _, _, th, s2, b0 in out:
        assert relerr(th, theta) < 1e-14 and relerr(s2, sigma2) < 1e-14 and relerr(b0, beta0) < 1e-14


@pytest.mark.gpu
def test_batched_submodel_fits_on_device():
    """8 sub-models of 2500 points (n = 20000 split like a NestedKriging): the batch (overlapping evaluations) equals
    the sequential loop bit for bit; each sub-model's objective agrees with the oracle at its theta.  The wall times
    of both are printed."""
    from oracle import kriging_oracle as ko
    n, d, p = 20000, 6, 8
    X, y, _ = synth(n, d, 77, "smooth")
    groups = nested.random_partition(n, p, seed=11)
    prm = {"theta": np.full((1, d), 0.6)}
    for k in nested.fit_submodels(y[:5000], X[:5000], nested.random_partition(5000, 2, seed=1), "matern5_2",
                                  parameters=prm, concurrent=1).values():
        k.close()  # warm-up: kernel modules loaded, allocator primed, before anything is timed
    t0 = time.perf_counter()
    seq = nested.fit_submodels(y, X, groups, "matern5_2", parameters=prm, concurrent=1)
    t_seq = time.perf_counter() - t0
    t0 = time.perf_counter()
    bat = nested.fit_submodels(y, X, groups, "matern5_2", parameters=prm, concurrent=8)
    t_bat = time.perf_counter() - t0
    print(f"\\n8 sub-model fits (n_g = 2500, d = 6): sequential {t_seq:.2f} s, 8 in flight {t_bat:.2f} s")
    for g in range(p):
        # overlapping evaluations return the bits of a lone handle: the batch IS the sequential loop
        assert np.array_equal(bat[g].theta(), seq[g].theta()) and bat[g].sigma2() == seq[g].sigma2()
        assert bat[g].fit_log["objective"] == seq[g].fit_log["objective"]
    for g in (0, p - 1):
        idx = groups[g]
        pb = ko.Problem(X=X[idx], y=y[idx], F=np.ones((len(idx), 1)), kernel="matern5_2")
        ll, _ = ko.log_likelihood(pb, bat[g].theta(), False)
        assert relerr(bat[g].logLikelihood(), ll) < 1e-9
    theta, sigma2, beta0 = nested.unify_hyperparameters(bat, groups, y, X)
    assert np.all(theta > 0) and sigma2 > 0
    for k in list(seq.values()) + list(bat.values()):
        k.close()


def _nested_reference_cases():
    import json
    import os
    from tests.util import GOLDEN
    return json.load(open(os.path.join(GOLDEN, "refgen_nested.json")))["cases"]


def _check_against_reference_nested(c, backend_factory, tol, concurrent):
    """Step 1 + step 2 of NestedKriging::fit (NestedKriging.cpp:262-331) on the partition the reference drew."""
    X, y, _ = synth(c["n"], c["d"], c["seed"], "smooth")
    groups = [np.array(g) for g in c["groups"]]
    prm = {"theta": np.full((1, c["d"]), c["theta0"])} if "theta0" in c else None
    kw = {"backend_factory": backend_factory} if backend_factory else {}
    m = nested.fit_submodels(y, X, groups, c["kernel"], parameters=prm, concurrent=concurrent, **kw)
    theta, sigma2, beta0 = nested.unify_hyperparameters(m, groups, y, X)
    assert relerr(theta, c["theta"]) < tol and relerr(sigma2, c["sigma2"]) < tol and relerr(beta0, c["beta0"]) < tol
    for g, sm in enumerate(c["submodels"]):
        assert relerr(m[g].theta(), sm["theta"]) < tol and relerr(m[g].sigma2(), sm["sigma2"]) < tol
        assert relerr(m[g].beta(), sm["beta"]) < tol
        assert relerr(m[g].logLikelihood(), sm["LL"]) < max(tol, 1e-9)
    for k in m.values():
        k.close()


@pytest.mark.parametrize("c", _nested_reference_cases(), ids=lambda c: c["name"])
def test_nested_matches_reference_nestedkriging_host(c):
    """Host logic (oracle as backend) against the UNMODIFIED reference's NestedKriging (tests/golden/refgen_nested.json,
    generated by oracle/_ref/ref_nested_driver): sub-model fits + unified hyper-parameters within 1e-6."""
    _check_against_reference_nested(c, OracleBackend, 1e-6, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("c", _nested_reference_cases(), ids=lambda c: c["name"])
def test_nested_matches_reference_nestedkriging_device(c):
    """The same with the device engine, all sub-model fits in flight at once."""
    _check_against_reference_nested(c, None, 1e-6, 8)
