"""Multistart sharding over one process per GPU.

Unit of work = one multistart row of theta0 = one full L-BFGS-B run including restarts
(reference optimize_worker, src/lib/Kriging.cpp:1904-2084).  Units share only read-only X, y, F and
bounds, and the reference already guarantees order independence ("BFGS20 == best of 20 x BFGS",
tests/KrigingTest.cpp:266-346); every rank draws the whole start-point stream from the same seed.  No
factorisation is split across GPUs.  Assignment of starts to ranks:
  * multistart <= world: static, rank r takes the starts {s : s mod G == r} (one start per GPU, BASELINE cfg 2 / 4);
  * multistart  > world: a dynamic queue -- a counter in the process group's store hands out start indices in order
    to whichever worker (rank, host thread) is free, so a straggling start does not hold idle the ranks that
    finished their share (BASELINE cfg 5: 64 starts on 8 GPUs, 8 in flight per GPU).  A start's result does not
    depend on who ran it (deterministic kernels on identical devices), so the fit is the same either way.

The only data exchange is the argmin of the reference's sequential loop (Kriging.cpp:2097-2110): one all-reduce
(sum) of a multistart x (3 + d + 1) table in which every start's row -- objective value, success flag, evaluation
count, gamma -- is filled by the one rank that ran it and zero elsewhere; every rank then applies the reference's
tie rule (strict '<' in start order) and rebuilds the committed model locally by one evaluation at gamma*.
One process per GPU is mandatory: the reference's L-BFGS-B and RNG hold process-global state.
"""
from __future__ import annotations

import math
import os

import numpy as np


class MultistartComm:
    def __init__(self, group=None):
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised")
        self.dist, self.torch, self.group = dist, torch, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.backend = dist.get_backend(group)
        if self.backend == "nccl":
            self.device = int(os.environ.get("LOCAL_RANK", self.rank))
            self.tdev = torch.device("cuda", self.device)
        else:
            self.device = int(os.environ.get("LOCAL_RANK", 0)) if torch.cuda.is_available() else 0
            self.tdev = torch.device("cpu")

    def my_starts(self, multistart: int):
        return [s for s in range(multistart) if s % self.world == self.rank]

    def start_queue(self, multistart: int, dynamic: bool | None = None):
        """Iterator-like hand-out of this fit's start indices.  dynamic=None: dynamic iff multistart > world."""
        if dynamic is None:
            dynamic = multistart > self.world and os.environ.get("LKGPU_STATIC_STARTS", "0") != "1"
        if not dynamic:
            return StaticQueue(self.my_starts(multistart))
        MultistartComm._fit_counter += 1
        return StoreQueue(self, multistart, f"lkgpu/startq/{MultistartComm._fit_counter}")

    _fit_counter = 0  # every rank runs the same sequence of fits, so the key of a fit's queue agrees across ranks

    def allreduce_sum(self, arr):
        """Element-wise sum over ranks of a float64 array."""
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self.tdev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy().copy()

    def allreduce_max(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.tdev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def allgather_vec(self, vec):
        """(world, len(vec)) array of every rank's vector (per-rank statistics for bench.py)."""
        t = self.torch.from_numpy(np.ascontiguousarray(vec, dtype=np.float64)).to(self.tdev)
        out = self.torch.empty((self.world, t.numel()), dtype=self.torch.float64, device=self.tdev)
        self.dist.all_gather_into_tensor(out.view(-1), t, group=self.group)
        return out.cpu().numpy().copy()

    def argmin_exchange(self, results: dict, multistart: int, gd: int):
        """results: {start_index: dict(success, objective_value, gamma, n_eval)} for the starts THIS rank ran (any
        subset; every start is run by exactly one rank).  Returns (best_idx, min_objective, gamma*, total number of
        evaluations over all ranks, owner[multistart] = rank that ran each start)."""
        table = np.zeros((multistart, 4 + gd))
        for s, r in results.items():
            ok = bool(r["success"]) and np.isfinite(r["objective_value"])
            table[s, 0] = r["objective_value"] if ok else 0.0
            table[s, 1] = 1.0 if ok else 0.0
            table[s, 2] = float(r.get("n_eval", 0))
            table[s, 3] = float(self.rank + 1)
            if ok:
                table[s, 4:] = np.asarray(r["gamma"], dtype=np.float64)
        table = self.allreduce_sum(table)  # x + 0 + ... + 0 is exact: every rank sees the owner's bits
        best_idx, min_ofn = -1, math.inf
        for s in range(multistart):  # the reference's order: start 0, 1, 2, ...; strict '<'
            if table[s, 1] > 0.5 and table[s, 0] < min_ofn:
                min_ofn, best_idx = float(table[s, 0]), s
        n_eval_total = int(round(float(table[:, 2].sum())))
        self.last_owner = (table[:, 3] - 1).astype(int)
        if best_idx < 0:
            return -1, math.inf, None, n_eval_total
        return best_idx, min_ofn, table[best_idx, 4:].copy(), n_eval_total


class StaticQueue:
    def __init__(self, starts):
        import threading
        self._it, self._lock, self.taken = iter(list(starts)), threading.Lock(), []

    def next(self):
        with self._lock:
            s = next(self._it, None)
            if s is not None:
                self.taken.append(s)
            return s


class StoreQueue:
    """Start indices 0 .. multistart-1 handed out by an atomic counter in the process group's key-value store (the
    rendezvous TCPStore of torch.distributed: no GPU traffic, ~0.1 ms per ticket against minutes per start)."""

    def __init__(self, comm, multistart, key):
        import threading
        from torch.distributed import distributed_c10d as c10d
        self._store = c10d._get_default_store()
        self._key, self._n, self._lock, self.taken = key, multistart, threading.Lock(), []

    def next(self):
        with self._lock:  # one client connection per process: tickets are drawn one at a time
            s = int(self._store.add(self._key, 1)) - 1
            if s >= self._n:
                return None
            self.taken.append(s)
            return s


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / MASTER_*), binding this
    process to cuda:LOCAL_RANK.  Returns a MultistartComm, or None for a single process."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if not dist.is_initialized():
        dist.init_process_group(backend=backend)
    return MultistartComm()
