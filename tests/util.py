"""Shared helpers for the tests: seeded synthetic inputs (must stay identical to
tests/golden/make_golden.py:synth) and fixture loading."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synth(n, d, seed, yfun="prodsin"):
    rng = np.random.Generator(np.random.PCG64(seed))
    X = rng.random((n, d))
    if yfun == "prodsin":
        y = np.prod(np.sin((X - 0.5) ** 2), axis=1)
    elif yfun == "sumsin":
        y = np.sum(np.sin(2 * np.pi * X), axis=1)
    elif yfun == "smooth":
        y = np.sin(3.0 * X[:, 0]) + np.sum(X * X, axis=1) + 0.05 * rng.standard_normal(n)
    else:
        raise ValueError(yfun)
    noise = 0.01 + 0.05 * rng.random(n)
    return X, y, noise


def synth_update(c):
    """Data of an update fixture (tests/golden/make_golden_update.py): n0 + n_u rows of synth(); with c["dup"] the
    first `dup` appended points duplicate kept ones up to 1e-8, which makes the Schur complement of the block
    extension numerically singular (the jitter ladder then runs on it)."""
    n0, n = c["n0"], c["n0"] + c["n_u"]
    X, y, noise = synth(n, c["d"], c["seed"], c.get("yfun", "smooth"))
    dup = c.get("dup", 0)
    if dup:
        X[n0:n0 + dup] = X[:dup] + 1e-8
        y[n0:n0 + dup] = y[:dup] + 1e-3 * np.arange(dup)
    return X, y, noise


def load_reference_vectors():
    with open(os.path.join(GOLDEN, "reference_vectors.json")) as f:
        return json.load(f)


def load_refgen():
    with open(os.path.join(GOLDEN, "refgen_vectors.json")) as f:
        return json.load(f)


def relerr(a, b):
    a = np.atleast_1d(np.asarray(a, float))
    b = np.atleast_1d(np.asarray(b, float))
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def relerr_vec(a, b):
    """Norm-wise relative error (used for gradients whose components can cross zero)."""
    a = np.atleast_1d(np.asarray(a, float))
    b = np.atleast_1d(np.asarray(b, float))
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
