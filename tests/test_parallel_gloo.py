"""world_size-2 gloo tests (CPU) of the multistart sharding: rank r runs the starts {s : s mod G == r}, one
all_gather + one broadcast pick the reference's argmin, and the sharded fit equals the single-process fit."""
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
        q.put((rank, k.theta().tolist(), k.sigma2(), k.fit_log["best_start"], k.fit_log["local_starts"],
               k._backend.n_objective_calls))
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
    for rank, theta, s2, best, local, ncalls in out:
        assert theta == k.theta().tolist() and s2 == k.sigma2() and best == k.fit_log["best_start"]
        assert local == [s for s in range(4) if s % 2 == rank]
    # the work really was split: each rank evaluated fewer objectives than the single process
    assert all(o[5] < k._backend.n_objective_calls for o in out)
