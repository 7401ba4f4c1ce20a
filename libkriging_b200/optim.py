"""Optimiser configuration, reparametrisation and the start-point RNG of the reference's fit driver.

* `Optim` statics and their LK_* environment defaults: reference src/lib/Optim.cpp:39-141, 211.
* reparam (gamma = log theta) and the Nugget alpha map: src/lib/Optim.cpp:53-60, src/lib/Kriging.cpp:678-702.
* `parse_method("BFGS20")`: src/lib/Optim.cpp:151-177.
* `Random`: process-global std::mt19937(123) + std::uniform_real_distribution<double>
  (src/lib/Random.cpp:18-43).  libstdc++'s generate_canonical<double, 53> consumes two 32-bit draws per
  double, u = (g1 + g2 * 2^32) / 2^64; start points must be bit-identical to the reference's so that
  a sharded multistart fit visits exactly the reference's starts.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass

import numpy as np


def _env(name, default, cast):
    v = os.environ.get(name)
    if v is None:
        return default
    if cast is bool:
        return v.strip().lower() in ("1", "true", "yes", "on")
    return cast(v)


@dataclass
class OptimConfig:
    reparametrize: bool = True
    theta_lower_factor: float = 0.02
    theta_upper_factor: float = 10.0
    variogram_bounds_heuristic: bool = True
    log_level: int = 0
    max_restart: int = 10
    max_iteration: int = 20
    gradient_tolerance: float = 1e-3
    objective_rel_tolerance: float = 1e-3

    @classmethod
    def from_env(cls):
        return cls(
            reparametrize=_env("LK_REPARAMETRIZE", True, bool),
            theta_lower_factor=_env("LK_THETA_LOWER_FACTOR", 0.02, float),
            theta_upper_factor=_env("LK_THETA_UPPER_FACTOR", 10.0, float),
            variogram_bounds_heuristic=_env("LK_VARIOGRAM_BOUNDS_HEURISTIC", True, bool),
            log_level=_env("LK_LOG_LEVEL", 0, int),
            max_restart=_env("LK_MAX_RESTART", 10, int),
            max_iteration=_env("LK_MAX_ITERATION", 20, int),
            gradient_tolerance=_env("LK_GRADIENT_TOLERANCE", 1e-3, float),
            objective_rel_tolerance=_env("LK_OBJECTIVE_REL_TOLERANCE", 1e-3, float),
        )


NUGGET_ALPHA_LOWER = 1e-3


def parse_method(method: str, prefix: str = "BFGS"):
    """'BFGS' -> ('BFGS', 1); 'BFGS20' -> ('BFGS', 20)."""
    m = re.fullmatch(re.escape(prefix) + r"(\d*)(.*)", method)
    if not m:
        raise ValueError(f"Unsupported optim: {method} (supported are: none, BFGS[#])")
    k = int(m.group(1)) if m.group(1) else 1
    return prefix + m.group(2), max(k, 1)


class Reparam:
    """theta <-> gamma maps for the three noise models (make_fit_objective, Kriging.cpp:1414-1522)."""

    def __init__(self, noise_model: str, d: int, enabled: bool = True):
        self.noise_model, self.d, self.enabled = noise_model, d, enabled

    def to(self, v):
        v = np.array(v, dtype=float)
        if not self.enabled:
            return v
        out = v.copy()
        out[:self.d] = np.log(v[:self.d])
        if self.noise_model == "nugget":
            out[self.d] = -np.log(1.0 + NUGGET_ALPHA_LOWER - v[self.d])
        elif self.noise_model == "hetero":
            out[self.d] = np.log(v[self.d])
        return out

    def frm(self, g):
        g = np.array(g, dtype=float)
        if not self.enabled:
            return g
        out = g.copy()
        out[:self.d] = np.exp(g[:self.d])
        if self.noise_model == "nugget":
            out[self.d] = 1.0 + NUGGET_ALPHA_LOWER - np.exp(-g[self.d])
        elif self.noise_model == "hetero":
            out[self.d] = np.exp(g[self.d])
        return out

    def deriv(self, v, grad):
        """chain rule: d/dgamma = d/dtheta * dtheta/dgamma."""
        if not self.enabled:
            return np.array(grad, dtype=float)
        out = np.array(grad, dtype=float) * np.array(v, dtype=float)
        if self.noise_model == "nugget":
            out[self.d] = grad[self.d] * (1.0 + NUGGET_ALPHA_LOWER - v[self.d])
        return out


class ReferenceRandom:
    """std::mt19937(seed) + std::uniform_real_distribution<double>() of libstdc++, bit for bit."""

    def __init__(self, seed: int = 123):
        self.seed = seed
        self.init()

    def init(self):
        rs = np.random.RandomState(self.seed)  # init_genrand(seed): the same seeding as std::mt19937(seed)
        self._bg = rs._bit_generator

    def randu(self, count: int) -> np.ndarray:
        raw = self._bg.random_raw(2 * count).astype(np.uint64)
        lo, hi = raw[0::2], raw[1::2]
        # (lo + hi * 2^32) / 2^64 in long-double-free exact arithmetic: the sum has < 2^64, a double holds 53 bits;
        # libstdc++ accumulates in double as well (sum += (g - min) * tmp; tmp *= 2^32), so mirror that order.
        s = lo.astype(np.float64) + hi.astype(np.float64) * 4294967296.0
        u = s / 18446744073709551616.0
        u[u >= 1.0] = np.nextafter(1.0, 0.0)
        return u

    def randu_vec(self, n: int) -> np.ndarray:
        return self.randu(n)

    def randu_mat(self, n: int, m: int) -> np.ndarray:
        return self.randu(n * m).reshape((n, m), order="F")  # arma imbue fills column by column
