"""Data pipeline pieces the benchmarks and drivers share.

* ``takens_embedding``: the layout of the reference's Data_OneStepAhead files -- row k is
  s[2k : 2k+5] (embedding dimension 4, lag 2, next value as target; verified against
  Sunspot/scaled_dataset.txt in the survey).
* ``synthetic_timeseries``: BASELINE configs[3] -- a 100 000-point Mackey-Glass series
  (tau = 17) + N(0, 0.01^2) noise, min-max scaled to [0, 1], first 60 % of the points for
  training (29 998 rows) and the rest for testing (19 998 rows).
* ``synthetic_pendigit``: BASELINE configs[4] -- 16 z-scored features drawn from 10 Gaussian
  class clusters, integer label in the last column (the shape of DATA/PenDigit, C:972-986).
"""
from __future__ import annotations

import numpy as np


def takens_embedding(series: np.ndarray, dim: int = 4, lag: int = 2) -> np.ndarray:
    s = np.asarray(series, dtype=np.float64)
    n = (len(s) - (dim + 1)) // lag + 1
    idx = lag * np.arange(n)[:, None] + np.arange(dim + 1)[None, :]
    return s[idx]


def mackey_glass(n: int, tau: int = 17, beta: float = 0.2, gamma: float = 0.1, power: int = 10,
                 x0: float = 1.2, warmup: int = 500) -> np.ndarray:
    hist = [x0] * (tau + 1)
    out = np.empty(n + warmup)
    x = x0
    for t in range(n + warmup):
        xt = hist[0]
        x = x + beta * xt / (1.0 + xt ** power) - gamma * x
        hist.append(x)
        hist.pop(0)
        out[t] = x
    return out[warmup:]


def synthetic_timeseries(points: int = 100_000, seed: int = 1234, train_fraction: float = 0.6):
    rng = np.random.default_rng(seed)
    s = mackey_glass(points) + rng.normal(0.0, 0.01, size=points)
    s = (s - s.min()) / (s.max() - s.min())
    cut = int(points * train_fraction)
    return takens_embedding(s[:cut]), takens_embedding(s[cut:])


def synthetic_pendigit(n_train: int = 20_000, n_test: int = 5_000, n_features: int = 16, n_classes: int = 10,
                       seed: int = 4321):
    rng = np.random.default_rng(seed)
    centers = rng.normal(0.0, 1.5, size=(n_classes, n_features))

    def draw(n):
        y = rng.integers(0, n_classes, size=n)
        x = centers[y] + rng.normal(0.0, 1.0, size=(n, n_features))
        return x, y

    xtr, ytr = draw(n_train)
    xte, yte = draw(n_test)
    for x in (xtr, xte):                                  # per-split z-scoring, as C:975-982 does
        x -= x.mean(axis=0)
        x /= x.std(axis=0)
    return np.hstack([xtr, ytr[:, None].astype(np.float64)]), np.hstack([xte, yte[:, None].astype(np.float64)])
