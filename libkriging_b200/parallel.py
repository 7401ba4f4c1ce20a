"""Multistart sharding over one process per GPU.

Unit of work = one multistart row of theta0 = one full L-BFGS-B run including restarts
(reference optimize_worker, src/lib/Kriging.cpp:1904-2084).  Units share only read-only X, y, F and
bounds, and the reference already guarantees order independence ("BFGS20 == best of 20 x BFGS",
tests/KrigingTest.cpp:266-346), so rank r of G takes the starts {s : s mod G == r}; every rank draws the
whole start-point stream from the same seed before slicing.  No factorisation is split across GPUs.

The only exchange is the argmin of the reference's sequential loop (Kriging.cpp:2097-2110):
  1. all_gather of (objective value, success flag) per start   -- 16 B per start
  2. every rank applies the reference's tie rule (strict '<' in start order)
  3. broadcast of gamma* from the owner of the best start        -- (d+1) * 8 B
after which each rank rebuilds the committed model locally by one evaluation at gamma*.
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

    def allreduce_sum(self, arr):
        """Element-wise sum over ranks of a float64 array (NestedKriging sub-model hyper-parameters: d + 2 doubles per
        group, each group owned by exactly one rank)."""
        t = self.torch.from_numpy(np.ascontiguousarray(arr, dtype=np.float64)).to(self.tdev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM, group=self.group)
        return t.cpu().numpy().copy()

    def allreduce_max(self, value):
        t = self.torch.tensor([float(value)], dtype=self.torch.float64, device=self.tdev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX, group=self.group)
        return float(t.item())

    def argmin_exchange(self, results: dict, multistart: int, gd: int):
        """results: {start_index: dict(success, objective_value, gamma, n_eval)} for this rank's starts.
        Returns (best_idx, min_objective, gamma*, total number of evaluations over all ranks)."""
        torch, dist = self.torch, self.dist
        per = (multistart + self.world - 1) // self.world
        mine = torch.full((per, 3), math.inf, dtype=torch.float64)
        mine[:, 1:] = 0.0
        for k, s in enumerate(self.my_starts(multistart)):
            r = results[s]
            mine[k, 0] = r["objective_value"] if r["success"] else math.inf
            mine[k, 1] = 1.0 if r["success"] else 0.0
            mine[k, 2] = float(r.get("n_eval", 0))
        mine = mine.to(self.tdev)
        gathered = torch.empty((self.world, per, 3), dtype=torch.float64, device=self.tdev)
        dist.all_gather_into_tensor(gathered.view(-1), mine.view(-1), group=self.group)
        g = gathered.cpu().numpy()
        best_idx, min_ofn = -1, math.inf
        for s in range(multistart):  # the reference's order: start 0, 1, 2, ...; strict '<'
            rk, k = s % self.world, s // self.world
            if g[rk, k, 1] > 0.5 and g[rk, k, 0] < min_ofn:
                min_ofn, best_idx = float(g[rk, k, 0]), s
        n_eval_total = int(round(float(g[:, :, 2].sum())))
        if best_idx < 0:
            return -1, math.inf, None, n_eval_total
        owner = best_idx % self.world
        buf = torch.zeros(gd, dtype=torch.float64)
        if self.rank == owner:
            buf[:] = torch.from_numpy(np.asarray(results[best_idx]["gamma"], dtype=np.float64))
        buf = buf.to(self.tdev)
        src = owner if self.group is None else dist.get_global_rank(self.group, owner)
        dist.broadcast(buf, src=src, group=self.group)
        return best_idx, min_ofn, buf.cpu().numpy().copy(), n_eval_total


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
