"""world_size-2 gloo tests (CPU) of the multistart sharding: starts are handed out statically (s mod G == r, at most
one per rank) or by the store-backed queue (more starts than ranks), one all-reduce picks the reference's argmin, and
the sharded fit equals the single-process fit bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.util import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from libkriging_b200.kriging import Kriging
        from libkriging_b200.parallel import MultistartComm
        from tests.oracle_backend import OracleBackend
        comm = MultistartComm()
        # (a) argmin rule on synthetic results: ties resolved by start order, failed starts ignored
        res = {}
        vals = {0: 5.0, 1: 3.0, 2: 3.0, 3: float("nan"), 4: 4.0}
        for s in comm.my_starts(5):
            ok = s != 3
            res[s] = dict(success=ok, objective_value=vals[s], gamma=np.array([float(s), 10.0 + s]), n_eval=s + 1)
        best, mn, gam, nev = comm.argmin_exchange(res, 5, 2)
        assert best == 1 and mn == 3.0 and gam.tolist() == [1.0, 11.0] and nev == 15
        # (b) sharded fit == single-process fit
        X, y, _ = synth(50, 2, 5, "smooth")
        k = Kriging("matern5_2", backend_factory=OracleBackend)
        k.fit(y, X, optim="BFGS4", comm=comm)
        out = [rank, k.theta().tolist(), k.sigma2(), k.fit_log["best_start"], k.fit_log["local_starts"],
               k.fit_log["n_eval_local"], k.fit_log["n_eval"]]
        # (c) one start per rank: static assignment s mod G == r
        k2 = Kriging("matern5_2", backend_factory=OracleBackend)
        k2.fit(y, X, optim="BFGS2", comm=comm)
        out += [k2.theta().tolist(), k2.fit_log["local_starts"]]
        q.put(tuple(out))
    finally:
        dist.destroy_process_group()


def test_sharded_multistart_equals_single_process():
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
    from libkriging_b200.kriging import Kriging
    from tests.oracle_backend import OracleBackend
    X, y, _ = synth(50, 2, 5, "smooth")
    k = Kriging("matern5_2", backend_factory=OracleBackend, concurrent_starts=1)
    k.fit(y, X, optim="BFGS4")
    k2 = Kriging("matern5_2", backend_factory=OracleBackend, concurrent_starts=1)
    k2.fit(y, X, optim="BFGS2")
    for rank, theta, s2, best, local, nloc, ntot, theta2, local2 in out:
        assert theta == k.theta().tolist() and s2 == k.sigma2() and best == k.fit_log["best_start"]
        assert ntot == k.fit_log["n_eval"]
        assert theta2 == k2.theta().tolist() and local2 == [rank]
    # dynamic queue: the four starts were split between the ranks, each run exactly once
    assert sorted(out[0][4] + out[1][4]) == [0, 1, 2, 3]
    assert out[0][5] + out[1][5] == k.fit_log["n_eval"]
