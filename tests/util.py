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
