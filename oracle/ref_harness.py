"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the UNMODIFIED reference scripts from /root/reference (read-only) so that
golden vectors can be generated from the reference's own code in this container.
The reference cannot travel to the GPU box, so everything produced through this
harness is committed as fixtures under tests/golden/ (see oracle/gen_golden.py).

What is stubbed, and why (nothing in the reference's arithmetic is touched):

* ``matplotlib*``      -- imported at module top (R:13-18, C:13-18) but only used for plots;
                          matplotlib is not installed here.
* ``multiprocessing``  -- replaced by a thread-backed shim (Process/Queue/JoinableQueue/Event
                          with the same methods).  ``ptReplica(multiprocessing.Process)`` (R:138,
                          C:157) then runs ``run()`` on a thread of THIS process, which lets the
                          harness give every replica its own recorded random stream.  The
                          handshake code (R:427-437, R:719-752) runs unmodified on the shim.
* ``np`` / ``random``  -- the module-global names are pointed at proxies whose ``random.uniform``
                          / ``normal`` / ``randn`` and ``uniform`` dispatch to a per-thread
                          ``RandomState`` and RECORD every draw (standard normals are rounded to
                          float32-representable values so the device can be fed identical numbers).

R: = multicore-pt-regression/pt_timeseries_regression.py
C: = multicore-pt-classification/pt_classification.py
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import queue
import sys
import threading
import types

import numpy as _np

REFERENCE_ROOT = os.environ.get("PTFNN_REFERENCE_ROOT", "/root/reference")
REG_PATH = os.path.join(REFERENCE_ROOT, "multicore-pt-regression", "pt_timeseries_regression.py")
CLS_PATH = os.path.join(REFERENCE_ROOT, "multicore-pt-classification", "pt_classification.py")


def reference_available() -> bool:
    return os.path.isfile(REG_PATH) and os.path.isfile(CLS_PATH)


# --------------------------------------------------------------------------------------
# stubs
# --------------------------------------------------------------------------------------
def _matplotlib_stubs():
    mods = {}
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mods["matplotlib"] = mpl
    for sub in ("mlab", "pyplot", "patches", "collections"):
        m = types.ModuleType("matplotlib." + sub)
        setattr(mpl, sub, m)
        mods["matplotlib." + sub] = m
    mods["matplotlib.patches"].Polygon = object
    mods["matplotlib.collections"].PatchCollection = object
    return mods


class _ThreadProcess(threading.Thread):
    """multiprocessing.Process look-alike running ``run()`` on a thread."""

    def __init__(self, *a, **k):
        threading.Thread.__init__(self)
        self.daemon = True


class _JoinableQueue(queue.Queue):
    pass


class _CountingEvent:
    """Event look-alike whose k-th ``wait()`` returns after the k-th ``set()``.

    The reference does ``signal_main.set()`` *then* ``event.clear()`` (R:432-433), a lost-wake-up
    race that is vanishingly rare across processes but real across threads.  Counting
    semantics give the intended hand-shake (one release per round) without touching the
    reference's code; ``clear()`` becomes a no-op.
    """

    def __init__(self):
        self._cv = threading.Condition()
        self._sets = 0
        self._seen = 0

    def set(self):
        with self._cv:
            self._sets += 1
            self._cv.notify_all()

    def clear(self):
        pass

    def is_set(self):
        with self._cv:
            return self._sets > self._seen

    def wait(self, timeout=None):
        with self._cv:
            ok = self._cv.wait_for(lambda: self._sets > self._seen, timeout)
            if ok:
                self._seen += 1
            return ok


def _multiprocessing_stub():
    m = types.ModuleType("multiprocessing")
    m.Process = _ThreadProcess
    m.Queue = queue.Queue
    m.JoinableQueue = _JoinableQueue
    m.Event = _CountingEvent
    return m


# --------------------------------------------------------------------------------------
# recorded per-thread random streams
# --------------------------------------------------------------------------------------
class Recorder:
    """Per-thread random streams + a log of every draw, in call order.

    Each thread (replica, or the main/coordinator thread) gets its own RandomState seeded
    from ``(seed, stream_index)``.  Stream 0 is the coordinator; replica k registers as k+1.
    Log entries: ("uniform", value) | ("normal", z_array) | ("randn", array) | ("pyuniform", value)
    """

    def __init__(self, seed: int):
        self.seed = int(seed)
        self._tls = threading.local()
        self.logs: dict[int, list] = {}
        self._lock = threading.Lock()
        self._by_thread: dict[int, int] = {}
        self.register(0)

    def register(self, stream: int):
        self._tls.stream = stream
        self._tls.rs = _np.random.RandomState([self.seed, stream])
        with self._lock:
            self.logs.setdefault(stream, [])

    def _state(self):
        if not hasattr(self._tls, "rs"):
            raise RuntimeError("thread used the recorded RNG without registering a stream")
        return self._tls.rs, self.logs[self._tls.stream]

    # numpy.random API used by the reference ------------------------------------------
    def uniform(self, low=0.0, high=1.0, size=None):
        rs, log = self._state()
        # float32-representable in [0,1): 24 random bits
        n = 1 if size is None else int(_np.prod(size))
        v = (rs.randint(0, 1 << 24, size=n).astype(_np.float64)) / float(1 << 24)
        v = low + (high - low) * v
        log.append(("uniform", v.copy()))
        if size is None:
            return float(v[0])
        return v.reshape(size)

    def _std_normal(self, n):
        rs, _ = self._state()
        return rs.standard_normal(n).astype(_np.float32).astype(_np.float64)

    def normal(self, loc=0.0, scale=1.0, size=None):
        _, log = self._state()
        n = 1 if size is None else int(_np.prod(size))
        z = self._std_normal(n)
        log.append(("normal", z.copy()))
        out = _np.asarray(loc, dtype=_np.float64) + float(scale) * z  # legacy normal = loc + scale*gauss
        if size is None:
            return float(out.reshape(-1)[0])
        return out.reshape(size)

    def randn(self, *shape):
        _, log = self._state()
        n = int(_np.prod(shape)) if shape else 1
        z = self._std_normal(n)
        log.append(("randn", z.copy()))
        if not shape:
            return float(z[0])
        return z.reshape(shape)

    def permutation(self, n):
        rs, _ = self._state()
        return rs.permutation(n)

    # Python `random` API used by the reference (R:387, C:400) --------------------------
    def pyuniform(self, a, b):
        rs, log = self._state()
        v = float(rs.randint(0, 1 << 24)) / float(1 << 24)
        v = a + (b - a) * v
        log.append(("pyuniform", v))
        return v


class _NumpyProxy:
    """Stands in for the module-global ``np`` of the reference; only ``.random`` differs."""

    def __init__(self, rec: Recorder):
        self.random = types.SimpleNamespace(
            uniform=rec.uniform, normal=rec.normal, randn=rec.randn, permutation=rec.permutation,
            seed=lambda *a, **k: None,
        )

    def __getattr__(self, name):
        return getattr(_np, name)


class _RandomProxy:
    def __init__(self, rec: Recorder):
        self.uniform = rec.pyuniform


# --------------------------------------------------------------------------------------
# loading
# --------------------------------------------------------------------------------------
_CACHE: dict[str, types.ModuleType] = {}


def load_reference(which: str) -> types.ModuleType:
    """Import the reference script ``which`` in {"regression","classification"} unmodified."""
    if which in _CACHE:
        return _CACHE[which]
    path = {"regression": REG_PATH, "classification": CLS_PATH}[which]
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    saved = {}
    stubs = _matplotlib_stubs()
    stubs["multiprocessing"] = _multiprocessing_stub()
    for k, v in stubs.items():
        saved[k] = sys.modules.get(k)
        sys.modules[k] = v
    try:
        import warnings
        spec = importlib.util.spec_from_file_location("_ptfnn_reference_" + which, path)
        mod = importlib.util.module_from_spec(spec)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _CACHE[which] = mod
    return mod


@contextlib.contextmanager
def recorded_rng(mod: types.ModuleType, rec: Recorder):
    """Point the reference module's ``np`` / ``random`` globals at recording proxies."""
    old_np, old_random = mod.np, mod.random
    mod.np = _NumpyProxy(rec)
    mod.random = _RandomProxy(rec)
    try:
        yield rec
    finally:
        mod.np, mod.random = old_np, old_random


@contextlib.contextmanager
def silenced():
    old = sys.stdout
    sys.stdout = io.StringIO()
    try:
        yield
    finally:
        sys.stdout = old


# --------------------------------------------------------------------------------------
# datasets (the reference's own files; R:883-909, C:920-957)
# --------------------------------------------------------------------------------------
REG_DATASETS = ("Lazer", "Sunspot", "Mackey", "Lorenz", "Rossler", "Henon", "ACFinance")


def load_regression_dataset(name: str):
    base = os.path.join(REFERENCE_ROOT, "multicore-pt-regression", "Data_OneStepAhead", name)
    return _np.loadtxt(os.path.join(base, "train.txt")), _np.loadtxt(os.path.join(base, "test.txt"))


def load_classification_dataset(name: str, split_seed: int = 0):
    base = os.path.join(REFERENCE_ROOT, "multicore-pt-classification", "DATA")
    if name == "Ionosphere":      # C:942-949
        tr = _np.genfromtxt(os.path.join(base, "Ions/Ions/ftrain.csv"), delimiter=",")[:, :-1]
        te = _np.genfromtxt(os.path.join(base, "Ions/Ions/ftest.csv"), delimiter=",")[:, :-1]
        return tr, te, [34, 50, 2]
    if name == "Cancer":          # C:950-957
        tr = _np.genfromtxt(os.path.join(base, "Cancer/ftrain.txt"), delimiter=" ")[:, :-1]
        te = _np.genfromtxt(os.path.join(base, "Cancer/ftest.txt"), delimiter=" ")[:, :-1]
        return tr, te, [9, 12, 2]
    if name == "Iris":            # C:920-930 + C:1003-1012 (z-score, 70/30 random split)
        data = _np.genfromtxt(os.path.join(base, "iris.csv"), delimiter=";")
        classes = data[:, 4].reshape(data.shape[0], 1) - 1
        features = data[:, 0:4].copy()
        for k in range(4):
            features[:, k] = (features[:, k] - _np.mean(features[:, k])) / _np.std(features[:, k])
        idx = _np.random.RandomState(split_seed).permutation(features.shape[0])
        ntr = int(0.7 * features.shape[0])
        tr = _np.hstack([features[idx[:ntr], :], classes[idx[:ntr], :]])
        te = _np.hstack([features[idx[ntr:], :], classes[idx[ntr:], :]])
        return tr, te, [4, 12, 3]
    if name == "PenDigit":        # C:972-986 (per-split z-score of the 16 features); first 1500 / 600 rows to keep the fixture small
        tr = _np.genfromtxt(os.path.join(base, "PenDigit/train.csv"), delimiter=",")
        te = _np.genfromtxt(os.path.join(base, "PenDigit/test.csv"), delimiter=",")
        for k in range(16):
            tr[:, k] = (tr[:, k] - _np.mean(tr[:, k])) / _np.std(tr[:, k])
            te[:, k] = (te[:, k] - _np.mean(te[:, k])) / _np.std(te[:, k])
        return tr[:1500].copy(), te[:600].copy(), [16, 30, 10]
    raise KeyError(name)


# --------------------------------------------------------------------------------------
# a full, deterministic, in-process run of the reference's ParallelTempering
# --------------------------------------------------------------------------------------
def run_reference_pt(which, traindata, testdata, topology, num_chains, maxtemp, NumSample,
                     swap_interval, use_langevin_gradients, learn_rate, langevin_prob, burn_in,
                     seed, path):
    """Drive the reference exactly as its ``main()`` does (R:995-1007 / C:1080-1092) and record
    every random draw plus the full-precision return value of every likelihood / prior call.

    Returns a dict of recorded draws, per-call values, the 11-tuple of ``run_chains()`` and the
    coordinator's swap counters.  Output files land under ``path`` (caller-provided tmp dir).
    """
    mod = load_reference(which)
    rec = Recorder(seed)
    calls: dict[int, dict[str, list]] = {}
    swaps: list = []

    orig_run = mod.ptReplica.run
    orig_lik = mod.ptReplica.likelihood_func
    orig_prior = mod.ptReplica.prior_likelihood
    orig_lg = mod.Network.langevin_gradient
    orig_swap = mod.ParallelTempering.swap_procedure
    tls = threading.local()

    def run_wrapped(self):
        tls.chain = self._harness_index
        rec.register(self._harness_index + 1)
        calls[self._harness_index] = {"lik": [], "prior": [], "lg_in": [], "lg_out": []}
        return orig_run(self)

    def lik_wrapped(self, fnn, data, w, *a):
        out = orig_lik(self, fnn, data, w, *a)
        if hasattr(tls, "chain"):
            calls[tls.chain]["lik"].append((float(out[0]), float(out[2]), int(data.shape[0]),
                                            float(self.adapttemp)))
        return out

    def prior_wrapped(self, *a):
        out = orig_prior(self, *a)
        if hasattr(tls, "chain"):
            calls[tls.chain]["prior"].append(float(out))
        return out

    def lg_wrapped(self, data, w, depth):
        w_in = _np.array(w, dtype=_np.float64, copy=True)
        out = orig_lg(self, data, w, depth)
        if hasattr(tls, "chain"):
            calls[tls.chain]["lg_in"].append(w_in)
            calls[tls.chain]["lg_out"].append(_np.array(out, dtype=_np.float64, copy=True))
        return out

    def swap_wrapped(self, q1, q2):
        p1, p2, swapped = orig_swap(self, q1, q2)
        swaps.append(bool(swapped))
        return p1, p2, swapped

    mod.ptReplica.run = run_wrapped
    mod.ptReplica.likelihood_func = lik_wrapped
    mod.ptReplica.prior_likelihood = prior_wrapped
    mod.Network.langevin_gradient = lg_wrapped
    mod.ParallelTempering.swap_procedure = swap_wrapped
    try:
        import warnings
        with recorded_rng(mod, rec), silenced(), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if which == "regression":
                pt = mod.ParallelTempering(use_langevin_gradients, learn_rate, traindata, testdata,
                                           topology, num_chains, maxtemp, NumSample, swap_interval,
                                           langevin_prob, path)
            else:
                pt = mod.ParallelTempering(use_langevin_gradients, learn_rate, traindata, testdata,
                                           topology, num_chains, maxtemp, NumSample, swap_interval,
                                           path)
            for d in ("/predictions/", "/posterior", "/results", "/surrogate",
                      "/surrogate/learnsurrogate_data", "/posterior/pos_w",
                      "/posterior/pos_likelihood", "/posterior/surg_likelihood",
                      "/posterior/accept_list"):
                pt.make_directory(path + d)
            pt.initialize_chains(burn_in)
            for k, ch in enumerate(pt.chains):
                ch._harness_index = k
            w0 = _np.stack([_np.array(ch.w, dtype=_np.float64) for ch in pt.chains])
            result = pt.run_chains()
    finally:
        mod.ptReplica.run = orig_run
        mod.ptReplica.likelihood_func = orig_lik
        mod.ptReplica.prior_likelihood = orig_prior
        mod.Network.langevin_gradient = orig_lg
        mod.ParallelTempering.swap_procedure = orig_swap
    return {
        "pt": pt, "w0": w0, "temperatures": _np.array(pt.temperatures, dtype=_np.float64),
        "logs": rec.logs, "calls": calls, "swaps": swaps, "result": result,
        "num_swap": pt.num_swap, "total_swap_proposals": pt.total_swap_proposals,
    }
