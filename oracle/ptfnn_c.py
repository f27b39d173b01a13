"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/ptfnn_oracle.c (float64 C oracle).

Same contract as oracle/ptfnn_numpy.py (takes every random draw as input); used where the NumPy
restatement would take minutes (full-length chains, 30k-row datasets).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import ptfnn_numpy as on

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libptfnn_oracle.so")
_lib = None


class _Cfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("task", "n_in", "n_hidden", "n_out", "n_replicas", "samples",
                                         "swap_interval", "use_langevin")] + \
               [(n, C.c_double) for n in ("l_prob", "learn_rate", "step_w", "step_eta", "sigma_squared",
                                          "nu_1", "nu_2", "pt_fraction")]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ptfnn_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.pto_prior.restype = C.c_double
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _cfg(cfg: on.PTConfig, R: int) -> _Cfg:
    I, H, O = cfg.topology
    return _Cfg(cfg.task, I, H, O, R, cfg.samples, cfg.swap_interval, int(cfg.use_langevin_gradients),
                cfg.l_prob, cfg.learn_rate, cfg.step_w, cfg.step_eta, cfg.sigma_squared, cfg.nu_1, cfg.nu_2,
                cfg.pt_fraction)


def evaluate(task, topology, data, w):
    I, H, O = topology
    data, w = _d(data), _d(w)
    fx = np.zeros(data.shape[0])
    prob = np.zeros((data.shape[0], O))
    lib().pto_evaluate(task, I, H, O, _p(data), data.shape[0], data.shape[1], _p(w), _p(fx), _p(prob))
    return (fx, prob) if task == on.CLASSIFICATION else fx


def langevin_gradient(task, topology, data, w, lr):
    I, H, O = topology
    data, w = _d(data), _d(w)
    out = np.zeros_like(w)
    lib().pto_langevin_gradient(task, I, H, O, _p(data), data.shape[0], data.shape[1], _p(w), C.c_double(lr), _p(out))
    return out


def likelihood(task, topology, data, w, tau_sq=1.0, adapttemp=1.0):
    """-> (loglik/adapttemp, rmse, accuracy)"""
    I, H, O = topology
    data, w = _d(data), _d(w)
    out = np.zeros(3)
    lib().pto_likelihood(task, I, H, O, _p(data), data.shape[0], data.shape[1], _p(w), C.c_double(tau_sq),
                         C.c_double(adapttemp), _p(out))
    return float(out[0]), float(out[1]), float(out[2])


def prior(task, topology, w, sigma_squared=25.0, nu_1=0.0, nu_2=0.0, tausq=1.0):
    I, H, O = topology
    w = _d(w)
    return float(lib().pto_prior(task, I, H, O, _p(w), C.c_double(sigma_squared), C.c_double(nu_1),
                                 C.c_double(nu_2), C.c_double(tausq)))


def swap_sweep(lhood, u_row):
    lhood, u_row = _d(lhood), _d(u_row)
    R = lhood.shape[0]
    src = np.zeros(R, dtype=np.int32)
    sw = np.zeros(max(R - 1, 1), dtype=np.uint8)
    lib().pto_swap_sweep(R, _p(lhood), _p(u_row), _p(src), _p(sw))
    return src, sw[:R - 1].astype(bool)


def total_rounds(cfg: on.PTConfig) -> int:
    c = _cfg(cfg, 1)
    return int(lib().pto_total_rounds(C.byref(c)))


def run_pt(cfg: on.PTConfig, train, test, temperatures, w0, draws: on.Draws, with_state=True) -> on.Traces:
    R, S, P = len(temperatures), cfg.samples, cfg.P
    rounds = cfg.total_rounds()
    tr = on.new_traces(R, S, P, rounds)
    if not with_state:
        tr.state_w = None
    train, test = _d(train), _d(test)
    accepted = np.zeros((R, S), dtype=np.uint8)
    swapped = np.zeros((max(rounds, 1), max(R - 1, 1)), dtype=np.uint8)
    counters = np.zeros(2, dtype=np.int64)
    lx, z, z_eta, u, us = _d(draws.lx), _d(draws.z), _d(draws.z_eta), _d(draws.u), _d(draws.u_swap)
    assert lx.shape == (R, S - 1) and z.shape == (R, S - 1, P) and u.shape == (R, S - 1)
    assert us.size >= rounds * max(R - 1, 0)
    c = _cfg(cfg, R)
    lib().pto_run_pt(C.byref(c), _p(train), train.shape[0], _p(test), test.shape[0], train.shape[1],
                     _p(_d(temperatures)), _p(_d(w0)), _p(lx), _p(z), _p(z_eta), _p(u), _p(us),
                     _p(tr.pos_w), _p(tr.lik_prop), _p(tr.lik_prop_t), _p(tr.prior_prop), _p(tr.diff_prop),
                     _p(tr.mh_prob), _p(tr.rmse_train), _p(tr.rmse_test), _p(tr.acc_train), _p(tr.acc_test),
                     _p(tr.accept_list), _p(accepted), _p(swapped), _p(tr.state_w), _p(tr.state_eta),
                     _p(tr.state_lik), _p(tr.state_prior), _p(tr.state_tau), _p(counters))
    tr.accepted = accepted.astype(bool)
    tr.swapped = swapped[:rounds, :max(R - 1, 0)].astype(bool)
    tr.num_swap, tr.total_swap_proposals = int(counters[0]), int(counters[1])
    if with_state:
        tr.final_w, tr.final_eta = tr.state_w[:, S - 1].copy(), tr.state_eta[:, S - 1].copy()
    return tr
