"""Shared helpers for the test-suite: fixture loading (tests/golden/*.npz) and oracle glue."""
import os

import numpy as np

from oracle import ptfnn_numpy as on

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["reg_sunspot_lg", "reg_lazer_rw", "reg_mackey_h10", "cls_iris_lg", "cls_cancer_lg", "cls_ions_lg",
         "reg_lorenz_lg", "reg_henon_lg1", "reg_acfin_rw", "reg_rossler_lg", "cls_iris_rw", "cls_pendigit_lg"]
REG_DATASETS = ["Lazer", "Sunspot", "Mackey", "Lorenz", "Rossler", "Henon", "ACFinance"]
CLS_DATASETS = ["Iris", "Cancer", "Ionosphere"]

_cache = {}


def npz(name):
    if name not in _cache:
        _cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False))
    return _cache[name]


def dataset(task, name):
    d = npz("datasets")
    pre = "reg_" if task in (on.REGRESSION, "regression") else "cls_"
    return d[pre + name + "_train"], d[pre + name + "_test"]


def cls_topology(name):
    return tuple(int(x) for x in npz("datasets")["cls_" + name + "_topology"])


def case(name):
    """-> (fixture dict, PTConfig, train, test, Draws)"""
    fx = npz(name)
    task = int(fx["task"])
    cfg = on.PTConfig(task=task, topology=tuple(int(x) for x in fx["topology"]), samples=int(fx["S"]),
                      swap_interval=int(fx["swap_interval"]), use_langevin_gradients=bool(fx["use_lg"]),
                      l_prob=float(fx["l_prob"]), learn_rate=float(fx["learn_rate"]))
    tr, te = dataset(task, str(fx["dataset"]))
    f64 = lambda a: np.asarray(a, dtype=np.float64)   # noqa: E731
    draws = on.Draws(lx=f64(fx["lx"]), z=f64(fx["z"]), z_eta=f64(fx["z_eta"]), u=f64(fx["u"]),
                     u_swap=f64(fx["u_swap"]))
    return fx, cfg, tr, te, draws


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b)))) if a.size else 0.0


def lik_scale(cfg, ref, n_train, temps):
    """[R, S] magnitude of the terms of the proposed log-likelihood (row i+1 = step i); 1 for classification."""
    if cfg.task != on.REGRESSION:
        return np.abs(ref.lik_prop)                          # a plain sum of log-probabilities: relative to itself
    tau = np.maximum(ref.state_tau, 1e-300)                  # row i+1: tau^2 proposed at step i (R:356)
    sse = (ref.rmse_train ** 2) * n_train                    # on accepted rows; carried rows understate it (harmless: a max below)
    t1 = 0.5 * n_train * np.abs(np.log(2.0 * np.pi * tau))
    t2 = 0.5 * sse / tau
    S = ref.lik_prop.shape[1]
    adapt = np.repeat(np.asarray(temps, dtype=np.float64)[:, None], S, axis=1)
    sw = cfg.pt_fraction * cfg.samples                       # R:301-324: adapttemp = 1 from step 0.6 S on, if that is an integer
    if float(sw).is_integer():
        adapt[:, int(sw) + 1:] = 1.0                         # row i+1 = step i
    return np.maximum(np.maximum(t1, t2) / adapt, np.abs(ref.lik_prop))
