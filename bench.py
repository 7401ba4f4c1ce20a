#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the lkgpu engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Default workload = BASELINE.json configs[1] (--config 2), the configuration the metric is quoted on:
    Kriging('matern5_2') log-likelihood + analytic gradient, synthetic n = 20000, d = 10, fp64, constant trend,
    evaluated at theta_k = 0.5 (well-conditioned point of SURVEY.md §8d); one evaluation = one "step".
    X = U[0,1)^(n x d), y = sin(3 x0) + sum x^2 + 0.05 N(0,1) (PCG64 seed 123) -- the inputs of the full-size
    reference fixture tests/golden/refgen_fullsize.json, so that the timed evaluation itself is compared with the
    reference's value and gradient (`parity_vs_reference`).  Every n x n buffer is 3.2 GB >> the 126 MB L2, so no L2
    flush is needed between steps.
Metric: LL+grad evaluations per second (whole job over all N GPUs) and the wall time of one full fit.
    value   : evaluations / s with X, y, F resident in HBM (lkgpu_objective_fun on a live handle)
    e2e     : the same through the host-buffer API: every step uploads X, y, F from pinned host memory
              (lkgpu_set_data), evaluates, and reads value + gradient back
    fit     : wall seconds of Kriging.fit(optim="BFGS<N>", objective="LL") on a GP-draw y: multistart sharded one
              start per GPU, with per-rank evaluation counts / device time and the balance efficiency
    batched : the batched-occupancy path at BASELINE configs[4]'s shape (gauss, n = 5000, d = 20): 8 handles in flight
              per GPU -- evaluations / s against one handle alone, and a BFGS<8 N> fit (8 starts per GPU, dynamic
              start queue across ranks)
Other configurations of BASELINE.json: --config 1 | 3 | 4 | 5 (same JSON line, their own workload).
N > 1: one process per GPU; each rank evaluates its own multistart point (no data-path collective, weak scaling);
the only exchange is the argmin at the end of the fit (libkriging_b200/parallel.py).
--impl reference: the unmodified reference (oracle/_ref/ref_driver, all host threads) on the same configuration,
MEASURED at full size when the host has the memory for it (cfg 2: 63 GB, about 3 minutes per evaluation on 16 cores:
one populate inside fit(optim="none") + one timed logLikelihoodFun), with the a n^3 + b n^2 extrapolation from two
small sizes beside it; only configurations the reference cannot allocate (cfg 4) are extrapolated, and say so.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # before any CUDA context (libkriging_b200/__init__.py)

UNIT = "evals/s"

# BASELINE.json configs (1-based as in SURVEY.md §8).  theta / alpha: the fixed evaluation point.
CONFIGS = {
    1: dict(kernel="gauss", noise_model="none", objective="LL", n=1000, d=4, theta=0.3, fit="BFGS",
            name="Kriging('gauss') LL + analytic gradient / BFGS fit, synthetic n=1000 d=4 fp64 (BASELINE configs[0])"),
    2: dict(kernel="matern5_2", noise_model="none", objective="LL", n=20000, d=10, theta=0.5, fit="BFGS",
            name="Kriging('matern5_2') LL + analytic gradient, synthetic n=20000 d=10 fp64, theta=0.5 (BASELINE configs[1])"),
    3: dict(kernel="exp", noise_model="none", objective="LOO", n=10000, d=6, theta=0.8, fit=None,
            name="Kriging('exp') LOO + analytic gradient, synthetic n=10000 d=6 fp64, theta=0.8 (BASELINE configs[2])"),
    4: dict(kernel="matern3_2", noise_model="nugget", objective="LL", n=40000, d=8, theta=0.6, alpha=0.9, fit="BFGS",
            name="NuggetKriging('matern3_2') LL + analytic gradient / BFGS<N> fit one start per GPU, synthetic n=40000 d=8 "
                 "fp64, theta=0.6 alpha=0.9 (BASELINE configs[3])"),
    5: dict(kernel="gauss", noise_model="none", objective="LL", n=5000, d=20, theta=1.2, fit="BFGS", handles=8,
            name="Kriging('gauss') LL + analytic gradient, 8 handles in flight per GPU / BFGS<8N> fit, synthetic n=5000 "
                 "d=20 fp64, theta=1.2 (BASELINE configs[4], batched-occupancy path)"),
}


def synth(n, d, seed):
    """Same generator as tests/util.py:synth(..., 'smooth') and tests/golden/make_golden_fullsize.py:synth (kept
    local: bench must not depend on tests/)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((n, d))
    y = np.sin(3.0 * X[:, 0]) + np.sum(X * X, axis=1) + 0.05 * rng.standard_normal(n)
    return X, y


def metric_name(cfg):
    return "loo_grad_evals_per_sec" if cfg["objective"] == "LOO" else "ll_grad_evals_per_sec"


def config_dict(cfg, args):
    """The workload description -- identical in both arms (the driver compares them)."""
    n, d = cfg["n"], cfg["d"]
    c = {"workload": cfg["name"] if (n, d) == (CONFIGS[args.config]["n"], CONFIGS[args.config]["d"]) else
         cfg["name"] + f" [overridden: n={n} d={d}]",
         "n": n, "d": d, "kernel": cfg["kernel"], "noise_model": cfg["noise_model"], "objective": cfg["objective"],
         "regmodel": "constant", "theta": cfg["theta"],
         "X": "U[0,1)^(n x d), PCG64 seed 123", "y": "sin(3 x0) + sum_k x_k^2 + 0.05 N(0,1)",
         "l2": "inputs larger than L2 (each n x n fp64 buffer is %.2f GB)" % (8.0 * n * n / 1e9) if n >= 8192 else
               "n x n buffers of %.0f MB: L2 flushed between steps by a 512 MB device write" % (8.0 * n * n / 1e6)}
    if "alpha" in cfg:
        c["alpha"] = cfg["alpha"]
    return c


def gamma_of(cfg):
    th = np.full(cfg["d"], cfg["theta"])
    return np.append(th, cfg["alpha"]) if cfg["noise_model"] == "nugget" else th


def flops_per_eval(cfg):
    """Algorithmic FP64 work of one evaluation with gradient: POTRF n^3/3 + TRTRI n^3/3 + LAUUM n^3/3;
    the LOO gradient adds one symmetric n x n x n product (n^3) in place of the reference's d dense products."""
    n = float(cfg["n"])
    return n ** 3 * (2.0 if cfg["objective"] == "LOO" else 1.0)


def load_golden(cfg):
    """Reference value + gradient for this exact workload, if tests/golden/refgen_fullsize.json holds it."""
    p = os.path.join(ROOT, "tests", "golden", "refgen_fullsize.json")
    if not os.path.isfile(p):
        return None, None
    try:
        cases = json.load(open(p))["cases"]
    except Exception:
        return None, None
    for name, c in cases.items():
        if (c["n"], c["d"], c["kernel"], c["noise_model"], c["objective"], c["seed"]) == \
           (cfg["n"], cfg["d"], cfg["kernel"], cfg["noise_model"], cfg["objective"], 123) and \
           abs(c["theta"] - cfg["theta"]) < 1e-15 and abs(c.get("extra", cfg.get("alpha", 0.0)) - cfg.get("alpha", 0.0)) < 1e-15:
            return name, c
    return None, None


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
        load = [s for s, p_ in zip(sm, pw) if p_ >= 0.5 * max(pw)] or sm  # power marks the samples under load
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


def host_mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return float(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


def ref_eval(cfg, n, reps, threads, timeout):
    """reps timed evaluations (value + gradient) of the reference at size n of this configuration's workload."""
    from oracle import ref
    X, y = synth(n, cfg["d"], 123)
    th = np.full((1, cfg["d"]), cfg["theta"])
    kw = {}
    if cfg["noise_model"] == "nugget":
        kw["gamma"] = np.append(th[0], cfg["alpha"])
    r = ref.run(X, y, kernel=cfg["kernel"], noise_model=cfg["noise_model"], mode="eval", objective=cfg["objective"],
                theta=th, grad=True, reps=reps, threads=threads, timeout=timeout, **kw)
    return [float(t) for t in r["eval_s_all"]], r


def ref_extrapolate(n_small, t_small, n_big, t_big, n_full):
    """t(n) = a n^3 + b n^2 through the two measured sizes (b clamped at >= 0: pure n^3 then)."""
    A = np.array([[n_small ** 3, n_small ** 2], [n_big ** 3, n_big ** 2]], float)
    a, b = np.linalg.solve(A, np.array([t_small, t_big], float))
    if a <= 0 or b < 0:
        a, b = t_big / n_big ** 3, 0.0
    return float(a * n_full ** 3 + b * n_full ** 2), float(a), float(b)


def ref_host_gb(n, d):
    """Peak resident memory of ref_driver (fit(optim='none') + one objective call): measured 19.6 x 8 n^2 bytes at
    d = 10 (n = 4000, 6000), of which dX is 8 d n^2."""
    return (8.0 * d + 77.0) * n * n / 1e9



def ref_sample(cfg, budget_s, threads):
    """Bounded sample: two sizes, the larger chosen so that ~3 evaluations fit `budget_s`; returns the measurement
    and the extrapolation to the configuration's n."""
    n_full = cfg["n"]
    n1 = min(1500, n_full)
    t1s, _ = ref_eval(cfg, n1, 2, threads, timeout=600)
    t1 = min(t1s)
    per = budget_s / 4.0
    n2 = int(min(n_full, 6000, max(2 * n1, n1 * (per / max(t1, 1e-3)) ** (1.0 / 3.0))) // 100 * 100)
    n2 = max(n2, n1)
    t2s, _ = ref_eval(cfg, n2, 2, threads, timeout=900)
    t2 = float(np.mean(t2s))
    if n2 > n1:
        t_full, a, b = ref_extrapolate(n1, t1, n2, t2, n_full)
    else:
        t_full, a, b = t2, 0.0, 0.0
    return dict(threads=threads, n_small=n1, t_small=t1, n_sample=n2, t_sample=t2, t_full=t_full, a=a, b=b)


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_driver missing (oracle/build_ref.sh)"}))
        return 0
    threads = host_threads()
    n, d = cfg["n"], cfg["d"]
    t_start = time.time()
    samp = ref_sample(cfg, budget_s=60.0, threads=threads)
    need_gb = ref_host_gb(n, d)
    avail_gb = host_mem_available_gb()
    measured, why_not, res = None, None, None
    est_total = 2.2 * samp["t_full"]  # populate inside fit(optim="none") + one timed evaluation
    if 1.0 * d * n * n > 2.0 ** 32 or need_gb > 0.9 * avail_gb:
        why_not = (f"the reference needs ~{need_gb:.0f} GB of host memory at n={n} (dX alone 8 d n^2 bytes; "
                   f"{avail_gb:.0f} GB available" + ("; > 2^32 elements: ARMA_32BIT_WORD" if 1.0 * d * n * n > 2 ** 32 else "") + ")")
    elif est_total > args.ref_budget:
        why_not = f"estimated {est_total:.0f} s for one populate + one evaluation exceeds the {args.ref_budget:.0f} s budget"
    else:
        try:
            # as many timed steps as the budget allows, at most the K the driver asked for
            reps = int(max(1, min(args.steps, (args.ref_budget - 1.2 * samp["t_full"]) // max(samp["t_full"], 1e-3))))
            timed, res = ref_eval(cfg, n, reps, threads, timeout=args.ref_budget * 1.5)
            measured = dict(steps=reps, times=timed, t=float(np.mean(timed)), populate_s=float(res["fit_s"]),
                            value=res["value"], grad=res["grad"])
        except Exception as ex:  # e.g. out of memory on a smaller host
            why_not = f"full-size run failed: {str(ex)[:160]}"
    t_eval = measured["t"] if measured else samp["t_full"]
    value = 1.0 / t_eval
    extrap = (f"t(n)=a n^3+b n^2 fitted on n={samp['n_small']} ({samp['t_small']:.3f} s) and n={samp['n_sample']} "
              f"({samp['t_sample']:.3f} s) extrapolates to {samp['t_full']:.1f} s/eval at n={n}")
    if measured:
        sample = (f"unmodified reference (oracle/_ref/ref_driver, OpenBLAS, {threads} threads) MEASURED at n={n} d={d}: "
                  f"{measured['t']:.1f} s per {cfg['objective']}+grad evaluation ({measured['steps']} timed step(s), after the "
                  f"populate_Model of fit(optim='none'): {measured['populate_s']:.1f} s); for comparison {extrap}")
    else:
        sample = (f"unmodified reference (oracle/_ref/ref_driver, OpenBLAS, {threads} threads) EXTRAPOLATED: {extrap}; "
                  f"not measured at full size because {why_not}")
    steps_done = measured["steps"] if measured else 0
    line = {
        "impl": "reference", "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": args.gpus,
        # steps / warmup: what was actually timed at full size (one evaluation takes minutes); the request is beside it
        "steps": steps_done if measured else args.steps, "warmup": 0 if measured else args.warmup,
        "steps_requested": args.steps, "warmup_requested": args.warmup,
        "steps_measured": steps_done, "extrapolated": measured is None,
        "ms_per_step": t_eval * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": config_dict(cfg, args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "reference_run": {"measured": measured is not None, "why_not_measured": why_not,
                          "eval_s": measured["times"] if measured else None,
                          "populate_s": measured["populate_s"] if measured else None,
                          "extrapolation_s": samp["t_full"], "extrapolation_fit": {k: samp[k] for k in
                                                                                  ("n_small", "t_small", "n_sample", "t_sample", "a", "b")},
                          "host_mem_available_gb": avail_gb, "host_mem_needed_gb": need_gb,
                          "wall_s": time.time() - t_start},
    }
    if measured:
        line["result"] = {"value": measured["value"], "grad": measured["grad"]}
        if args.golden_out:  # tests/golden/refgen_fullsize.json entry of this run (generator: this command line)
            c = dict(n=n, d=d, seed=123, kernel=cfg["kernel"], noise_model=cfg["noise_model"], objective=cfg["objective"],
                     theta=cfg["theta"], value=measured["value"], grad=measured["grad"], eval_s=measured["times"][0],
                     populate_s=measured["populate_s"], threads=threads)
            if "alpha" in cfg:
                c["extra"] = cfg["alpha"]
            X, y = synth(n, d, 123)
            c.update(y_sum=float(np.sum(y)), X_sum=float(np.sum(X)))
            with open(args.golden_out, "w") as f:
                json.dump({f"cfg{args.config}": c}, f, indent=1)
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
# the B200 arm
# ------------------------------------------------------------------------------------------------
def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback of B200_PROFILING.md"


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


def gp_draw(_capi, X, y0, kernel, theta, device, seed=321):
    """y = 1.5 + 2 L z with L = chol(R(theta*)) taken from the engine (untimed set-up): a draw of the GP itself, so
    that the fit has an interior, well-conditioned optimum."""
    n, d = X.shape
    with _capi.Engine(X, y0, np.ones((n, 1)), kernel=kernel, device=device) as e0:
        e0.objective("LL", np.full(d, theta), False)
        L = e0.export("L")
    z = np.random.Generator(np.random.PCG64(seed)).standard_normal(n)
    y = 1.5 + 2.0 * (L @ z)
    del L
    return y


def fit_block(Kriging, cfg, X, y, optim, local, comm, world, max_over_ranks, barrier, torch, concurrent=None,
              ladder_shortcut=None):
    """One Kriging.fit through the host API; returns the `fit` dict incl. per-rank statistics and balance."""
    k = Kriging(cfg["kernel"], cfg["noise_model"], device=local, concurrent_starts=concurrent,
                ladder_shortcut=ladder_shortcut)
    barrier()
    t0 = time.perf_counter()
    k.fit(y, X, optim=optim, objective=cfg["objective"] if cfg["objective"] != "LOO" else "LL", comm=comm)
    torch.cuda.synchronize()
    t_local = time.perf_counter() - t0
    t_fit = max_over_ranks(t_local)
    st = dict(getattr(k._backend, "stats", {}) or {})
    fit = {"wall_s": t_fit, "optim": optim, "n_eval_all_ranks": int(k.fit_log["n_eval"]),
           "starts": int(k.fit_log["multistart"]), "best_start": int(k.fit_log["best_start"]),
           "LL_at_fit": -float(k.fit_log["objective"]), "theta": [float(t) for t in k.theta()],
           "sigma2": float(k.sigma2()), "concurrent_starts_per_gpu": int(k.fit_log.get("concurrent_starts", 1)),
           "start_assignment": "static s mod G" if int(k.fit_log["multistart"]) <= world else
                               "dynamic queue (process-group store counter)" if world > 1 else "in order"}
    if cfg["noise_model"] == "nugget":
        fit["nugget"] = float(k.nugget())
    vec = np.array([st.get("evals", 0), st.get("device_ms", 0.0), t_local, len(k.fit_log["local_starts"]),
                    st.get("jitter_rungs", 0), st.get("reject_rcond", 0), st.get("reject_info", 0),
                    st.get("rungs_skipped", 0), st.get("chol_ms", 0.0), st.get("rcond_ms", 0.0)], dtype=np.float64)
    allv = comm.allgather_vec(vec) if comm is not None else vec[None, :]
    dev_ms = allv[:, 1]
    fit["per_rank"] = {"evals": [int(v) for v in allv[:, 0]], "device_ms": [float(v) for v in dev_ms],
                       "wall_s": [float(v) for v in allv[:, 2]], "starts_run": [int(v) for v in allv[:, 3]]}
    # balance of the sharded fit: 1.0 = every GPU busy for the whole fit (device time inside evaluations)
    fit["balance_eff"] = float(dev_ms.sum() / (len(dev_ms) * dev_ms.max())) if dev_ms.max() > 0 else None
    fit["ladder"] = {"jitter_rungs_accepted_sum": int(allv[:, 4].sum()), "rungs_rejected_by_rcond": int(allv[:, 5].sum()),
                     "rungs_rejected_by_failed_factorisation": int(allv[:, 6].sum()),
                     "rungs_skipped_by_shortcut": int(allv[:, 7].sum()), "chol_ms": float(allv[:, 8].sum()),
                     "rcond_ms": float(allv[:, 9].sum())}
    k.close()
    return fit


def concurrent_throughput(_capi, X, y, cfg, device, handles, reps, stream_of, torch):
    """evaluations / s with `handles` engine handles in flight on this GPU (one host thread each), and the same for
    one handle alone; every result is compared bit for bit with the lone handle's."""
    n, d = X.shape
    F = np.ones((n, 1))
    g = gamma_of(cfg)
    thetas = [g * (1.0 + 0.05 * k) for k in range(3)]
    engines = [_capi.Engine(X, y, F, kernel=cfg["kernel"], noise_model=cfg["noise_model"], device=device)
               for _ in range(handles)]
    ref = []
    for th in thetas:
        v, gr = engines[0].objective(cfg["objective"], th, True)
        ref.append((v, gr.copy()))
    for e in engines[1:]:
        e.objective(cfg["objective"], thetas[0], True)
    # one handle alone
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for r in range(reps):
        for th in thetas:
            engines[0].objective(cfg["objective"], th, True)
    torch.cuda.synchronize()
    t_single = (time.perf_counter() - t0) / (reps * len(thetas))
    bad = [0]
    lock = threading.Lock()

    def worker(e):
        b = 0
        for r in range(reps):
            for k, th in enumerate(thetas):
                v, gr = e.objective(cfg["objective"], th, True)
                if v != ref[k][0] or not np.array_equal(gr, ref[k][1]):
                    b += 1
        with lock:
            bad[0] += b

    ths = [threading.Thread(target=worker, args=(e,)) for e in engines]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    nev = handles * reps * len(thetas)
    for e in engines:
        e.close()
    return {"handles_in_flight": handles, "evals": nev, "wall_s": wall, "evals_per_s": nev / wall,
            "ms_per_eval": 1e3 * wall / nev, "one_handle_ms_per_eval": 1e3 * t_single,
            "speedup_vs_one_handle": t_single / (wall / nev), "mismatching_vs_lone_handle": bad[0],
            "timing": "host wall clock around the batch (device idle before and after)"}


def run_b200_arm(args, cfg):
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
    import scipy.optimize._lbfgsb_py  # noqa: F401  (the Python host's L-BFGS-B: its ~1 s import is not part of a fit)
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

    n, d, K, W = cfg["n"], cfg["d"], args.steps, max(args.warmup, 3)
    obj = cfg["objective"]
    X, y = synth(n, d, 123)
    F = np.ones((n, 1))
    # pinned host staging for the e2e leg (column-major X)
    Xp = torch.from_numpy(np.ascontiguousarray(X.T)).pin_memory()
    yp = torch.from_numpy(y.copy()).pin_memory()
    Fp = torch.from_numpy(np.ascontiguousarray(F.T)).pin_memory()
    X_h, y_h, F_h = Xp.numpy().T, yp.numpy(), Fp.numpy().T  # F-contiguous views of the pinned buffers
    # each rank = one multistart stream: its own evaluation point around the configuration's (rank 0: exactly it)
    gamma = gamma_of(cfg)
    gamma[:d] *= (1.0 + 0.01 * rank)

    peaks, peak_src = measured_peaks()
    fp64 = fp64_peak_probe(torch, dev) if rank == 0 else {}

    eng = _capi.Engine(X_h, y_h, F_h, kernel=cfg["kernel"], noise_model=cfg["noise_model"], device=local)
    stream = torch.cuda.ExternalStream(eng.stream_ptr, device=dev)
    flush = None
    if n < 8192:  # working set comparable with the 126 MB L2: flush it between steps
        flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed_region(step_fn, nsteps):
        """K steps bracketed by barrier + synchronize; CUDA events on the engine's launch stream."""
        barrier()
        if flush is None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            acc = [step_fn(i) for i in range(nsteps)]
            e1.record(stream)
            barrier()
            return e0.elapsed_time(e1), acc
        total, acc = 0.0, []
        for i in range(nsteps):  # per-step events so that the flush stays outside the timed intervals
            with torch.cuda.stream(stream):
                flush.fill_(i & 255)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            acc.append(step_fn(i))
            e1.record(stream)
            e1.synchronize()
            total += e0.elapsed_time(e1)
        barrier()
        return total, acc

    # ---- value: inputs resident in HBM ----
    def step_resident(i):
        return eng.objective(obj, gamma, True, with_info=True)

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
    val, grad, last_info = acc[-1]

    # ---- e2e: host buffers in, value + gradient out, copies inside the timed region ----
    def step_e2e(i):
        eng.set_data(X_h, y_h, F_h)
        return eng.objective(obj, gamma, True)

    step_e2e(0)
    ms_e2e_total, acc2 = timed_region(step_e2e, K)
    ms_e2e = max_over_ranks(ms_e2e_total / K)
    e2e_value = world / (ms_e2e * 1e-3)
    h2d = (n * d + n + n) * 8
    d2h = (320 + 1 + 1) * 8 + 20 + 8  # engine's scalar block + beta + info words + the jitter-loop norms
    assert acc2[-1][0] == val, "e2e path must reproduce the resident path bit for bit"
    eng.close()
    del flush
    torch.cuda.empty_cache()

    # ---- parity of the timed evaluation with the reference (rank 0 evaluates exactly the fixture's point) ----
    parity = None
    gname, gold = load_golden(cfg)
    if rank == 0 and gold is not None:
        gref = np.asarray(gold["grad"], dtype=np.float64)
        parity = {"fixture": f"tests/golden/refgen_fullsize.json:{gname}",
                  "reference": "unmodified libKriging (oracle/_ref/ref_driver), same X, y, theta",
                  "value_relerr": abs(val - gold["value"]) / abs(gold["value"]),
                  "grad_relerr_norm": float(np.linalg.norm(grad - gref) / np.linalg.norm(gref)),
                  "grad_relerr_max": float(np.max(np.abs(grad - gref) / np.maximum(np.abs(gref), 1e-300))),
                  "y_checksum_relerr": abs(float(np.sum(y)) - gold["y_sum"]) / abs(gold["y_sum"]),
                  "tolerance": 1e-10, "reference_eval_s": gold.get("eval_s"), "reference_threads": gold.get("threads")}

    # ---- fit wall time: Kriging.fit(BFGS<world>), one start per GPU, on a GP-draw y ----
    fit = None
    fit_plain = None
    if not args.no_fit and cfg.get("fit"):
        if cfg["noise_model"] == "none":
            y_fit = gp_draw(_capi, X, y, cfg["kernel"], cfg["theta"], local)
            y_desc = f"GP draw at theta*={cfg['theta']} (y = 1.5 + 2 L z, seed 321)"
        else:
            y_fit, y_desc = y, "the evaluation legs' y (smooth function + noise: what the nugget models)"
        starts = world * cfg.get("handles", 1)
        optim = "BFGS" if starts == 1 else f"BFGS{starts}"
        # untimed warm-up of the host path (first use of the optimiser module, the worker threads, the bounds kernels:
        # ~1 s once per process, which would otherwise sit inside a sub-second mid-size fit)
        try:
            kw = Kriging(cfg["kernel"], cfg["noise_model"], device=local, concurrent_starts=min(2, cfg.get("handles") or 1))
            kw.fit(y_fit[:512], X[:512], optim="BFGS2", objective="LL")
            kw.close()
        except Exception:  # pragma: no cover -- the warm-up must never decide the bench
            pass
        fit = fit_block(Kriging, cfg, X, y_fit, optim, local, comm, world, max_over_ranks, barrier, torch,
                        concurrent=cfg.get("handles"))
        fit["y"] = y_desc
        fit["ladder_mode"] = ("shortcut: safe_chol_lower's lowest accepted rung found by a bracketed search from the "
                              "previous evaluation's rung (lkgpu_set_ladder_shortcut, the engine's default)")
        # the same fit with the reference's plain ladder (rung 0, 1, 2, ... on every evaluation): same model expected
        if args.config == 2 and not args.no_plain_ladder:
            fp = fit_block(Kriging, cfg, X, y_fit, optim, local, comm, world, max_over_ranks, barrier, torch,
                           concurrent=cfg.get("handles"), ladder_shortcut=False)
            fit_plain = {"wall_s": fp["wall_s"], "n_eval_all_ranks": fp["n_eval_all_ranks"], "LL_at_fit": fp["LL_at_fit"],
                         "ladder": fp["ladder"], "theta_identical_to_fit": fp["theta"] == fit["theta"],
                         "LL_identical_to_fit": fp["LL_at_fit"] == fit["LL_at_fit"],
                         "n_eval_identical_to_fit": fp["n_eval_all_ranks"] == fit["n_eval_all_ranks"]}

    # ---- the same fit through the C++ host (lkgpu::Kriging: Armadillo API + lbfgsb_cpp loop on the CPU, every
    #      objective evaluation through the C ABI) -- the host north_star names.  N = 1: one process on this GPU.
    #      N > 1: rank 0 launches one C++ driver process per GPU; they share the multistart rows over lkgpu::ShardComm
    #      (TCP star, no NCCL / MPI) while the Python ranks wait on the rendezvous store with their GPUs idle ----
    fit_cpp = None
    if fit is not None and not args.no_cpp_host and cfg["noise_model"] == "none":
        barrier()
        if rank == 0:
            try:
                from libkriging_b200.host import driver as cpp
                if cpp.available():
                    t0 = time.perf_counter()
                    if world == 1:
                        rs = [cpp.run(X, y_fit, kernel=cfg["kernel"], objective="LL", mode="fit", optim="BFGS",
                                      device=local, timeout=900)]
                    else:
                        rs = cpp.run(X, y_fit, kernel=cfg["kernel"], objective="LL", mode="fit", optim=optim, timeout=1500,
                                     world=world, devices=list(range(world)), concurrent_starts=cfg.get("handles"))
                    r = rs[0]
                    fit_cpp = {"host": "libkriging_b200/host/lkgpu_host_driver (C++: Armadillo + lbfgsb_cpp)",
                               "processes": world, "optim": "BFGS" if world == 1 else optim,
                               "wall_s": max(float(q["fit_s"]) for q in rs),
                               "cuda_init_s": max(float(q.get("cuda_init_s", 0.0)) for q in rs),
                               "wall_s_incl_process_start": time.perf_counter() - t0,
                               "n_eval": int(r["n_eval"]), "LL_at_fit": float(r["objective_at_fit"]),
                               "theta": [float(t) for t in r["theta"]], "sigma2": float(r["sigma2"])}
                    if world > 1:
                        fit_cpp["exchange"] = ("lkgpu::ShardComm: tickets for the start queue + one all-gather of a row per "
                                               "start over TCP (rank 0 serves); no GPU traffic")
                        fit_cpp["per_rank"] = [{"rank": q["rank"], "device": q["rank"], "fit_s": float(q["fit_s"]),
                                                "cuda_init_s": q.get("cuda_init_s"), "evals": int(q["local_n_eval"]),
                                                "starts": q["local_starts"]} for q in rs]
                        fit_cpp["all_ranks_same_model"] = all(q["theta"] == r["theta"] and q["sigma2"] == r["sigma2"]
                                                              for q in rs)
                    th_py = np.asarray(fit["theta"])
                    fit_cpp["theta_relerr_vs_python_host"] = float(np.max(np.abs(np.asarray(r["theta"]) - th_py) / th_py))
                    fit_cpp["LL_relerr_vs_python_host"] = abs(fit_cpp["LL_at_fit"] - fit["LL_at_fit"]) / abs(fit["LL_at_fit"])
                else:
                    fit_cpp = {"unavailable": "libkriging_b200/host/_build/lkgpu_host_driver not built"}
            except Exception as ex:  # pragma: no cover
                fit_cpp = {"failed": str(ex)[:300]}
        if world > 1:  # CPU-side wait (an NCCL barrier would keep a spinning kernel on the waiting GPUs)
            from datetime import timedelta
            from torch.distributed import distributed_c10d as c10d
            st = c10d._get_default_store()
            if rank == 0:
                st.set("lkgpu/bench/cpp_fit_done", "1")
            else:
                st.wait(["lkgpu/bench/cpp_fit_done"], timedelta(seconds=3000))
            barrier()

    # ---- batched-occupancy path (BASELINE configs[4] shape) ----
    batched = None
    if not args.no_batched and (args.config == 2 or cfg.get("handles")):
        c5 = dict(CONFIGS[5])
        X5, y5 = synth(c5["n"], c5["d"], 123)
        thr = concurrent_throughput(_capi, X5, y5, c5, local, c5["handles"], 6, None, torch)
        thr_agg = sum_over_ranks(thr["evals"]) / max_over_ranks(thr["wall_s"])
        batched = {"workload": c5["name"], "throughput": thr, "evals_per_s_all_gpus": thr_agg}
        if not args.no_fit and args.config != 5:  # (--config 5: that fit is the line's own `fit` block)
            y5f = gp_draw(_capi, X5, y5, c5["kernel"], c5["theta"], local)
            fb = fit_block(Kriging, c5, X5, y5f, f"BFGS{8 * world}", local, comm, world, max_over_ranks, barrier, torch,
                           concurrent=c5["handles"])
            fb["y"] = f"GP draw at theta*={c5['theta']}"
            batched["fit"] = fb

    # ---- Kriging::update, no refit: the last 5 % of the rows appended to a model of the
    #      first 95 %; block extension of the kept factor vs the from-scratch factorisation of all rows ----
    update = None
    if not args.no_update and rank == 0 and cfg["noise_model"] == "none" and obj == "LL":
        n_u = max(1, n // 20)
        n0 = n - n_u
        with _capi.Engine(X[:n0], y[:n0], F[:n0], kernel=cfg["kernel"], device=local) as eu:
            eu.objective("LL", gamma, False)
            eu.commit_model()
            t0 = time.perf_counter()
            eu.append_data(X[n0:], y[n0:], F[n0:])
            t_append = time.perf_counter() - t0
            vu, _, iu = eu.objective("LL", gamma, False, with_info=True)
            used = eu.last_eval_was_update
            vs, _, isc = eu.objective("LL", gamma, False, with_info=True)   # same point again: from scratch
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
    # LAUUM (n^3/3) (+ the LOO gradient's product).  achieved = algorithmic flop per evaluation / device time of those
    # stages (CUDA events recorded by the engine on its launch stream, inside the timed region).
    flops = flops_per_eval(cfg)
    gemm_ms = stages["chol"] + stages["trtri"] + stages["lauum"] + (stages["grad"] if obj == "LOO" else 0.0)
    achieved = flops / (gemm_ms * 1e-3) / 1e12
    peak = fp64.get("cublas_dgemm_tflops") or fp64.get("dmma_issue_tflops")
    peak = max(peak, fp64.get("dmma_issue_tflops", 0.0)) if args.peak == "max" else peak
    n3 = float(n) ** 3
    roofline = {
        "bound": "tensor", "kernel": "gemm_dmma_kernel (FP64 DMMA.8x8x4 + TMA ring)", "achieved": achieved,
        "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": None,
        "peak_source": "measured in this run: cuBLAS DGEMM 8192^3 best of 10 (MEASURED_PEAKS.json has no FP64 entry)",
        "fp64_probes": fp64,
        "flops_per_eval": flops, "gemm_stage_ms": gemm_ms,
        "per_stage": {k: {"flops": n3 / 3, "ms": stages[k], "tflops": n3 / 3 / (stages[k] * 1e-3) / 1e12}
                      for k in ("chol", "trtri", "lauum") if stages[k] > 0},
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
    if os.path.isfile(tr) and args.config == 2 and n == 20000:
        try:
            tj = json.load(open(tr))
            roofline["traffic"] = tj.get("gemm_dmma_kernel_bytes_per_eval")
            roofline["traffic_source"] = ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum summed "
                                          "over the gemm_dmma_kernel launches of one evaluation, from profiles/traffic.json "
                                          "(" + str(tj.get("source", "ncu capture")) + ")")
        except Exception:
            pass

    # ---- cpu baseline on this box's host cores (bounded sample) ----
    cpu = None
    if not args.no_cpu:
        try:
            m = ref_sample(cfg, budget_s=25.0, threads=host_threads())
            note = ""
            if gold is not None and gold.get("eval_s"):
                note = (f"; measured at full size when the fixture was generated: {gold['eval_s']:.1f} s/eval on "
                        f"{gold.get('threads')} threads (tests/golden/refgen_fullsize.json:{gname})")
            cpu = {"value": 1.0 / m["t_full"], "unit": UNIT, "cores": m["threads"], "kind": "reference",
                   "sample": (f"unmodified reference (oracle/_ref/ref_driver, OpenBLAS) {obj}+grad at n={m['n_sample']} d={d}: "
                              f"{m['t_sample']:.3f} s/eval; n={m['n_small']}: {m['t_small']:.3f} s/eval; t(n)=a n^3+b n^2 "
                              f"extrapolated to n={n}: {m['t_full']:.1f} s/eval" + note)}
        except Exception as ex:
            cpu = {"value": None, "unit": UNIT, "cores": host_threads(), "kind": "reference", "sample": f"failed: {ex}"[:300]}

    cdict = config_dict(cfg, args)
    line = {
        "metric": metric_name(cfg), "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": cdict,
        "parallelism": f"multistart x{world}: one independent evaluation stream per GPU, no data-path collective",
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "lkgpu_set_data (pinned host X, y, F) + lkgpu_objective_fun through libkriging_b200._capi"},
        "gpu_launches": total_launches,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "parity_vs_reference": parity,
        "fit": fit,
        "fit_plain_ladder": fit_plain,
        "fit_cpp_host": fit_cpp,
        "batched": batched,
        "update": update,
        "stages_ms": stages,
        "result": {"value": val, "grad_norm": float(np.linalg.norm(grad)), "n_jitter": last_info["n_jitter"],
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
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=None, help="override the configuration's n (debug)")
    ap.add_argument("--d", type=int, default=None, help="override the configuration's d (debug)")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-update", action="store_true")
    ap.add_argument("--no-batched", action="store_true")
    ap.add_argument("--no-cpp-host", action="store_true")
    ap.add_argument("--no-plain-ladder", action="store_true", help="skip the second fit with the plain jitter ladder")
    ap.add_argument("--peak", default="cublas", choices=["cublas", "max"])
    ap.add_argument("--ref-budget", type=float, default=1300.0,
                    help="reference arm: seconds allowed for the full-size run (the driver's limit is 1800 s)")
    ap.add_argument("--golden-out", default=None, help="reference arm: write the measured value / gradient here")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.n:
        cfg["n"] = args.n
    if args.d:
        cfg["d"] = args.d
    if args.impl == "reference":
        return run_reference_arm(args, cfg)
    return run_b200_arm(args, cfg)


if __name__ == "__main__":
    sys.exit(main())
