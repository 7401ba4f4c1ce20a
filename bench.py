#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the lkgpu engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], the configuration the metric is quoted on):
    Kriging('matern5_2') log-likelihood + analytic gradient, synthetic n = 20000, d = 10, fp64, constant trend,
    evaluated at theta_k = 0.5 (well-conditioned point of SURVEY.md §8d), one evaluation = one "step".
    Every n x n buffer is 3.2 GB >> the 126 MB L2, so no L2 flush is needed between steps.
Metric: LL+grad evaluations per second (whole job over all N GPUs) and the wall time of one full fit.
    value  : evaluations / s with X, y, F resident in HBM (lkgpu_objective_fun on a live handle)
    e2e    : the same metric through the host-buffer API: every step uploads X, y, F from pinned host memory
             (lkgpu_set_data), evaluates, and reads value + gradient back
    fit    : wall seconds of Kriging.fit(optim="BFGS<N>", objective="LL"): multistart sharded one start per GPU
N > 1: one process per GPU; each rank evaluates its own multistart point (no data-path collective, weak scaling);
the only exchange is the argmin at the end of the fit (libkriging_b200/parallel.py).
--impl reference: the unmodified reference (oracle/_ref/ref_driver, all host threads) on a bounded sample of the same
workload, extrapolated to n = 20000 with t(n) = a n^3 + b n^2 fitted on two sizes (the reference cannot run n = 20000
in minutes: ~25 min per evaluation and > 50 GB of host memory).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

KERNEL = "matern5_2"
METRIC = "ll_grad_evals_per_sec"
UNIT = "evals/s"


def synth(n, d, seed):
    """Same generator as tests/util.py:synth(..., 'smooth') (kept local: bench must not depend on tests/)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((n, d))
    y = np.sin(3.0 * X[:, 0]) + np.sum(X * X, axis=1) + 0.05 * rng.standard_normal(n)
    return X, y


def workload_name(n, d):
    return f"Kriging('{KERNEL}') LL + analytic gradient, synthetic n={n} d={d} fp64, theta=0.5 (BASELINE configs[1])"


# ------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md: sample nvidia-smi DURING the timed region)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); pw.append(float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        self.f.close()
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # samples under load = the upper half of the SM clock samples is not meaningful; use power as the load marker
        load = [s for s, p_ in zip(sm, pw) if p_ >= 0.5 * max(pw)] or sm
        return {"sm_mhz": float(np.median(load)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference / cpu baseline (oracle/_ref: the unmodified reference compiled by oracle/build_ref.sh)
# ------------------------------------------------------------------------------------------------
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def ref_eval_times(n, d, reps, threads, timeout=1500):
    from oracle import ref
    X, y = synth(n, d, 123)
    r = ref.run(X, y, kernel=KERNEL, mode="eval", objective="LL", theta=np.full((1, d), 0.5), grad=True, reps=reps,
                threads=threads, timeout=timeout)
    return [float(t) for t in r["eval_s_all"]], r


def ref_extrapolate(n_small, t_small, n_big, t_big, n_full):
    """t(n) = a n^3 + b n^2 through the two measured sizes (b clamped at >= 0: pure n^3 then)."""
    A = np.array([[n_small ** 3, n_small ** 2], [n_big ** 3, n_big ** 2]], float)
    a, b = np.linalg.solve(A, np.array([t_small, t_big], float))
    if a <= 0 or b < 0:
        a, b = t_big / n_big ** 3, 0.0
    return float(a * n_full ** 3 + b * n_full ** 2), float(a), float(b)


def cpu_reference_measure(n_full, d, steps, warmup, budget_s):
    """Times the reference on a bounded sample; returns dict(times of the timed steps, extrapolated full-size s)."""
    from oracle import ref
    if not ref.available():
        raise RuntimeError("oracle/_ref/ref_driver missing (run oracle/build_ref.sh in the build container)")
    threads = host_threads()
    n1 = min(1500, n_full)
    t1s, _ = ref_eval_times(n1, d, 2, threads)
    t1 = min(t1s)
    # pick the big sample so that (warmup + steps + 1 populate) evaluations fit the budget (n^3 scaling estimate)
    per = budget_s / (warmup + steps + 1.0)
    n2 = int(min(n_full, 6000, max(2 * n1, n1 * (per / max(t1, 1e-3)) ** (1.0 / 3.0))) // 100 * 100)
    if n2 <= n1:
        n2 = n1
    t2s, _ = ref_eval_times(n2, d, warmup + steps, threads)
    timed = t2s[warmup:]
    t2 = float(np.mean(timed))
    if n2 > n1:
        t_full, a, b = ref_extrapolate(n1, t1, n2, t2, n_full)
    else:
        t_full, a, b = t2, 0.0, 0.0
    return dict(threads=threads, n_small=n1, t_small=t1, n_sample=n2, timed=timed, t_sample=t2, t_full=t_full, a=a, b=b)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    n, d = args.n, args.d
    m = cpu_reference_measure(n, d, args.steps, args.warmup, budget_s=150.0)
    value = 1.0 / m["t_full"]
    sample = (f"reference (oracle/_ref/ref_driver, OpenBLAS, {m['threads']} threads) LL+grad at n={m['n_sample']} d={d}: "
              f"{m['t_sample']:.3f} s/eval (mean of {len(m['timed'])} timed steps) and n={m['n_small']}: {m['t_small']:.3f} s; "
              f"t(n)=a n^3+b n^2 fitted on both and extrapolated to n={n}: {m['t_full']:.1f} s/eval")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["t_full"] * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, d), "n": n, "d": d, "kernel": KERNEL,
                   "sample_n": m["n_sample"], "sample_ms_per_step": m["t_sample"] * 1e3, "extrapolated": True},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": m["threads"], "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def fp64_peak_probe(torch, dev):
    """FP64 peak on this GPU.  MEASURED_PEAKS.json has no FP64 entry, so it is measured here:
    (a) cuBLAS DGEMM 8192^3 best of 10 (library used ONLY as the peak probe);
    (b) dependency-free DMMA.8x8x4 issue-rate probe of the engine (lkgpu_probe_fp64_peak)."""
    from libkriging_b200 import _capi
    out = {}
    try:
        a = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        c = torch.empty_like(a)
        for _ in range(2):
            torch.matmul(a, b, out=c)
        best = 0.0
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b, out=c)
            e1.record()
            e1.synchronize()
            best = max(best, 2.0 * 8192 ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out["cublas_dgemm_tflops"] = best
        del a, b, c
        torch.cuda.empty_cache()
    except Exception as ex:  # pragma: no cover
        out["cublas_dgemm_error"] = str(ex)[:200]
    out["dmma_issue_tflops"] = _capi.probe_fp64_peak(dev.index or 0, 0)
    out["dfma_issue_tflops"] = _capi.probe_fp64_peak(dev.index or 0, 1)
    return out


def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from libkriging_b200 import _capi
    from libkriging_b200.kriging import Kriging
    comm = None
    if world > 1:
        from libkriging_b200 import parallel
        comm = parallel.init_from_env("nccl")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return float(x)
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    n, d, K, W = args.n, args.d, args.steps, max(args.warmup, 3)
    X, y = synth(n, d, 123)
    F = np.ones((n, 1))
    if not args.smooth_y:
        # y = one draw of the GP itself at theta* = 0.5 ("a draw from the GP itself at a known theta*"):
        # y = L z with L = chol(R(theta*)) taken from the engine (untimed set-up), so that the fit has an interior,
        # well-conditioned optimum instead of running into the theta upper bound.
        with _capi.Engine(X, y, F, kernel=KERNEL, device=local) as e0:
            e0.objective("LL", np.full(d, 0.5), False)
            L = e0.export("L")
        z = np.random.Generator(np.random.PCG64(321)).standard_normal(n)
        y = 1.5 + 2.0 * (L @ z)
        del L
    # pinned host staging for the e2e leg (column-major X)
    Xp = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
    yp = torch.from_numpy(y.copy()).pin_memory()
    Fp = torch.from_numpy(np.ascontiguousarray(F.T)).pin_memory()
    X_h, y_h, F_h = Xp.numpy().T, yp.numpy(), Fp.numpy().T  # F-contiguous views of the pinned buffers
    # each rank = one multistart stream: its own evaluation point around theta = 0.5 (rank 0: exactly 0.5)
    theta = np.full(d, 0.5) * (1.0 + 0.01 * rank)

    peaks, peak_src = measured_peaks()
    fp64 = fp64_peak_probe(torch, dev) if rank == 0 else {}

    eng = _capi.Engine(X_h, y_h, F_h, kernel=KERNEL, device=local)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)

    def timed_region(step_fn, nsteps):
        """K steps bracketed by barrier + synchronize; CUDA events on the engine's launch stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        acc = []
        for i in range(nsteps):
            acc.append(step_fn(i))
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1), acc

    # ---- value: inputs resident in HBM ----
    def step_resident(i):
        v, g, info = eng.objective("LL", theta, True, with_info=True)
        return v, g, info

    for i in range(W):
        step_resident(i)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launch_count
    ms_total, acc = timed_region(step_resident, K)
    launches = eng.launch_count - l0
    clocks = sampler.stop()
    ms_step = max_over_ranks(ms_total / K)
    value = world / (ms_step * 1e-3)
    stage_names = _capi.STAGE_NAMES
    stages = {k: float(np.mean([a[2]["stage_ms"][k] for a in acc])) for k in stage_names}
    ll_val, ll_grad, last_info = acc[-1]

    # ---- e2e: host buffers in, value + gradient out, copies inside the timed region ----
    def step_e2e(i):
        eng.set_data(X_h, y_h, F_h)
        return eng.objective("LL", theta, True)

    step_e2e(0)
    ms_e2e_total, acc2 = timed_region(step_e2e, K)
    ms_e2e = max_over_ranks(ms_e2e_total / K)
    e2e_value = world / (ms_e2e * 1e-3)
    h2d = (n * d + n + n) * 8
    d2h = (320 + 1 + 1) * 8 + 20 + 8  # engine's scalar block + beta + info words + the jitter-loop norms
    assert acc2[-1][0] == ll_val, "e2e path must reproduce the resident path bit for bit"

    # ---- fit wall time: Kriging.fit(BFGS<world>), one start per GPU ----
    fit = None
    if not args.no_fit:
        k = Kriging(KERNEL, device=local)
        barrier()
        t0 = time.perf_counter()
        k.fit(y, X, optim=f"BFGS{world}" if world > 1 else "BFGS", objective="LL", comm=comm)
        torch.cuda.synchronize()
        t_fit = max_over_ranks(time.perf_counter() - t0)
        fit = {"wall_s": t_fit, "optim": f"BFGS{world}" if world > 1 else "BFGS", "n_eval_all_ranks": int(k.fit_log["n_eval"]),
               "starts": int(k.fit_log["multistart"]), "best_start": int(k.fit_log["best_start"]),
               "LL_at_fit": float(k.fit_log["objective"]) * -1.0, "theta": [float(t) for t in k.theta()],
               "sigma2": float(k.sigma2())}
        st = getattr(k._backend, "stats", None)
        if st:  # this rank's handle: objective calls, jitter-ladder rungs climbed, device time inside the evaluations
            fit.update(evals_this_rank=int(st["evals"]), jitter_rungs_this_rank=int(st["jitter_rungs"]),
                       device_ms_this_rank=float(st["device_ms"]), rungs_rejected_by_failed_factorisation=int(st["reject_info"]),
                       rungs_rejected_by_rcond=int(st["reject_rcond"]), chol_ms_this_rank=float(st["chol_ms"]),
                       rcond_ms_this_rank=float(st["rcond_ms"]))
        k.close()
    eng.close()

    # ---- Kriging::update, no refit: the last 5 % of the rows appended to a model of the
    #      first 95 %; block extension of the kept factor vs the from-scratch factorisation of all rows ----
    update = None
    if not args.no_update and rank == 0:
        n_u = max(1, n // 20)
        n0 = n - n_u
        with _capi.Engine(X[:n0], y[:n0], F[:n0], kernel=KERNEL, device=local) as eu:
            eu.objective("LL", theta, False)
            eu.commit_model()
            t0 = time.perf_counter()
            eu.append_data(X[n0:], y[n0:], F[n0:])
            t_append = time.perf_counter() - t0
            vu, _, iu = eu.objective("LL", theta, False, with_info=True)
            used = eu.last_eval_was_update
            vs, _, isc = eu.objective("LL", theta, False, with_info=True)   # same point again: from scratch
        update = {"n0": n0, "n_u": n_u, "block_extension": bool(used), "append_s": t_append,
                  "eval_ms_block_extension": iu["stage_ms"]["total"], "chol_ms_block_extension": iu["stage_ms"]["chol"],
                  "eval_ms_from_scratch": isc["stage_ms"]["total"], "chol_ms_from_scratch": isc["stage_ms"]["chol"],
                  "LL_relerr_vs_from_scratch": abs(vu - vs) / abs(vs)}

    total_launches = int(sum_over_ranks(launches))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel: gemm_dmma_kernel (FP64 DMMA tile engine) ----
    # It is the only kernel with O(n^3) work: Cholesky trailing updates + panel TRSM (n^3/3), TRTRI (n^3/3),
    # LAUUM (n^3/3).  achieved = algorithmic n^3 flop per evaluation / device time of those three stages
    # (CUDA events recorded by the engine on its launch stream, inside the timed region).
    gemm_ms = stages["chol"] + stages["trtri"] + stages["lauum"]
    flops = float(n) ** 3
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    peak = fp64.get("cublas_dgemm_tflops") or fp64.get("dmma_issue_tflops")
    peak = max(peak, fp64.get("dmma_issue_tflops", 0.0)) if args.peak == "max" else peak
    roofline = {
        "bound": "tensor", "kernel": "gemm_dmma_kernel (FP64 DMMA.8x8x4 + TMA ring)", "achieved": achieved,
        "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": None,
        "peak_source": "measured in this run: cuBLAS DGEMM 8192^3 best of 10 (MEASURED_PEAKS.json has no FP64 entry)",
        "fp64_probes": fp64,
        "flops_per_eval": flops, "gemm_stage_ms": gemm_ms,
        "per_stage": {
            "chol": {"flops": flops / 3, "ms": stages["chol"], "tflops": flops / 3 / (stages["chol"] * 1e-3) / 1e12},
            "trtri": {"flops": flops / 3, "ms": stages["trtri"], "tflops": flops / 3 / (stages["trtri"] * 1e-3) / 1e12},
            "lauum": {"flops": flops / 3, "ms": stages["lauum"], "tflops": flops / 3 / (stages["lauum"] * 1e-3) / 1e12},
        },
        "whole_eval_tflops": flops / (ms_step * 1e-3) / 1e12,
        "hbm_side_stages": {
            "cov_build": {"bytes": 4.0 * n * n, "ms": stages["cov"],
                          "gbs": 4.0 * n * n / (stages["cov"] * 1e-3) / 1e9 if stages["cov"] > 0 else None},
            "grad_reduce": {"bytes": 4.0 * n * n, "ms": stages["grad"],
                            "gbs": 4.0 * n * n / (stages["grad"] * 1e-3) / 1e9 if stages["grad"] > 0 else None},
            "solves": {"bytes": 3 * 4.0 * n * n, "ms": stages["solves"],
                       "gbs": 12.0 * n * n / (stages["solves"] * 1e-3) / 1e9 if stages["solves"] > 0 else None},
            "hbm_peak_gbs": peaks.get("hbm_gbs"), "hbm_peak_source": peak_src,
        },
    }
    tr = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tr):
        try:
            roofline["traffic"] = json.load(open(tr)).get("gemm_dmma_kernel_bytes_per_eval")
        except Exception:
            pass

    # ---- cpu baseline on this box's host cores (bounded sample) ----
    cpu = None
    if not args.no_cpu:
        try:
            m = cpu_reference_measure(n, d, steps=2, warmup=1, budget_s=25.0)
            cpu = {"value": 1.0 / m["t_full"], "unit": UNIT, "cores": m["threads"], "kind": "reference",
                   "sample": (f"unmodified reference (oracle/_ref/ref_driver, OpenBLAS) LL+grad at n={m['n_sample']} d={d}: "
                              f"{m['t_sample']:.3f} s/eval; n={m['n_small']}: {m['t_small']:.3f} s/eval; t(n)=a n^3+b n^2 "
                              f"extrapolated to n={n}: {m['t_full']:.1f} s/eval")}
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference", "sample": f"failed: {ex}"[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(n, d), "n": n, "d": d, "kernel": KERNEL, "objective": "LL", "regmodel": "constant",
                   "y": "analytic smooth function" if args.smooth_y else "GP draw at theta*=0.5 (y = 1.5 + 2 L z, seed 321)",
                   "l2": "inputs larger than L2 (each n x n fp64 buffer is %.1f GB)" % (8.0 * n * n / 1e9),
                   "parallelism": f"multistart x{world}: one independent evaluation stream per GPU, no data-path collective"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "lkgpu_set_data (pinned host X, y, F) + lkgpu_objective_fun through libkriging_b200._capi"},
        "gpu_launches": total_launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "fit": fit,
        "update": update,
        "stages_ms": stages,
        "result": {"LL": ll_val, "grad_norm": float(np.linalg.norm(ll_grad)), "n_jitter": last_info["n_jitter"],
                   "rcond": last_info["rcond"]},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--d", type=int, default=10)
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-update", action="store_true")
    ap.add_argument("--smooth-y", action="store_true", help="analytic y instead of the GP draw (debug)")
    ap.add_argument("--peak", default="cublas", choices=["cublas", "max"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
