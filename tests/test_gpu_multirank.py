"""On-device multistart sharding (needs >= 2 GPUs; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multirank.py -m gpu`):
BFGS8 sharded over 2 ranks (one process per GPU, NCCL, dynamic start queue) returns bit for bit the theta / sigma2 /
objective of BFGS8 in one process -- the reference's own order-independence test (tests/KrigingTest.cpp:266-346,
"BFGS20 == best of 20 x BFGS") carried to GPUs.  Also the static assignment (one start per rank)."""
import os
import socket

import numpy as np
import pytest

from tests.util import synth

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _data():
    return synth(1500, 4, 81, "smooth")[:2]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        from libkriging_b200.kriging import Kriging
        from libkriging_b200.parallel import MultistartComm
        comm = MultistartComm()
        X, y = _data()
        out = [rank]
        for optim in ("BFGS8", "BFGS2"):
            k = Kriging("matern5_2", device=rank)
            k.fit(y, X, optim=optim, comm=comm)
            out += [k.theta().tolist(), k.sigma2(), k.fit_log["objective"], k.fit_log["best_start"],
                    k.fit_log["local_starts"], k.fit_log["n_eval"]]
            k.close()
        q.put(tuple(out))
    finally:
        dist.destroy_process_group()


def test_sharded_fit_on_two_gpus_equals_single_process():
    import torch
    import torch.multiprocessing as mp
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0
    out = sorted(q.get(timeout=10) for _ in range(2))
    from libkriging_b200.kriging import Kriging
    X, y = _data()
    ref = {}
    for optim in ("BFGS8", "BFGS2"):
        k = Kriging("matern5_2", device=0)
        k.fit(y, X, optim=optim)
        ref[optim] = (k.theta().tolist(), k.sigma2(), k.fit_log["objective"], k.fit_log["best_start"], k.fit_log["n_eval"])
        k.close()
    for r in out:
        th8, s8, o8, b8, loc8, ne8, th2, s2, o2, b2, loc2, ne2 = r[1:]
        assert (th8, s8, o8, b8, ne8) == ref["BFGS8"]
        assert (th2, s2, o2, b2, ne2) == ref["BFGS2"]
        assert loc2 == [r[0]]  # static: start s on rank s mod 2
    # dynamic queue: the eight starts were split between the two ranks, each run exactly once
    assert sorted(out[0][5] + out[1][5]) == list(range(8))
