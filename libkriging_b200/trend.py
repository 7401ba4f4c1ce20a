"""Trend basis F (n x p).  Host-side O(n p) work; F is an INPUT of the device engine.
Mirrors Trend::regressionModelMatrix (reference src/lib/Trend.cpp:34-93), same column order."""
from __future__ import annotations

import numpy as np

REGMODELS = ("none", "constant", "linear", "interactive", "quadratic")


def regression_model_matrix(regmodel: str, X: np.ndarray) -> np.ndarray:
    n, d = X.shape
    if regmodel not in REGMODELS:
        raise ValueError(f"Unsupported regression model: {regmodel}")
    if regmodel == "none":
        return np.ones((n, 0))
    cols = [np.ones(n)]
    if regmodel == "linear":
        cols += [X[:, i] for i in range(d)]
    elif regmodel in ("interactive", "quadratic"):
        for i in range(d):
            cols.append(X[:, i])
            upto = i if regmodel == "interactive" else i + 1
            for j in range(upto):
                cols.append(X[:, i] * X[:, j])
    return np.asfortranarray(np.column_stack(cols))
