"""ctypes binding of the C ABI in include/ptfnn.h (libptfnn.so).

There is no fallback: if the library has not been built (``python __graft_entry__.py``) or no CUDA
device is present, calls raise ``PtfnnError`` -- nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libptfnn.so")

OK, E_INVALID, E_CUDA, E_STATE, E_UNSUPPORTED, E_NOMEM = 0, -1, -2, -3, -4, -5
TASK_REGRESSION, TASK_CLASSIFICATION = 0, 1
SWAP_RULE_AUTO, SWAP_RULE_AFTER_I, SWAP_RULE_BEFORE_I1 = -1, 0, 1
SWAP_KIND_REFERENCE, SWAP_KIND_RATIO_TEMPERATURE = 0, 1      # R:674 | Misc/ldpt_fnn_multi_fixed.py:520 (opt-in)
ABI_VERSION = 2
PEER_HANDLE_BYTES = 64       # cudaIpcMemHandle_t
MAX_HIDDEN = 512             # widest hidden layer a specialisation is generated for

# every symbol include/ptfnn.h declares (tests/test_capi_symbols.py checks the header against this)
SYMBOLS = [
    "ptfnn_abi_version", "ptfnn_build_info", "ptfnn_device_count", "ptfnn_default_config", "ptfnn_last_error",
    "ptfnn_create", "ptfnn_destroy", "ptfnn_set_stream", "ptfnn_set_data", "ptfnn_init_chains",
    "ptfnn_set_state", "ptfnn_get_state", "ptfnn_get_step", "ptfnn_run", "ptfnn_replay", "ptfnn_sync",
    "ptfnn_generate_draws", "ptfnn_swap_uniforms", "ptfnn_get_traces", "ptfnn_traces_begin", "ptfnn_traces_end", "ptfnn_get_swap_stats", "ptfnn_trace_summary", "ptfnn_predictive_summary", "ptfnn_predictive_bands",
    "ptfnn_swap_pending", "ptfnn_swap_export", "ptfnn_swap_plan", "ptfnn_swap_apply",
    "ptfnn_peer_export", "ptfnn_peer_connect", "ptfnn_has_topology", "ptfnn_register_kernels",
    "ptfnn_op_forward_pass", "ptfnn_op_evaluate_proposal", "ptfnn_op_langevin_gradient", "ptfnn_time_langevin_gradient", "ptfnn_op_likelihood", "ptfnn_op_prior",
    "ptfnn_op_swap_sweep", "ptfnn_op_swap_sweep_kind", "ptfnn_op_posterior_predictive", "ptfnn_savetxt", "ptfnn_loadtxt",
]


class PtfnnError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libptfnn error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "abi_version", "task", "n_in", "n_hidden", "n_out", "n_replicas", "n_replicas_global", "replica_offset",
        "samples", "swap_interval", "swap_rule", "use_langevin_gradients", "common_random_numbers",
        "memoize_gradient", "device", "barrier_timeout_ms", "debug_traces", "speculation", "swap_kind", "window_plan")] + \
        [("seed", C.c_uint64)] + \
        [(n, C.c_double) for n in ("l_prob", "learn_rate", "step_w", "step_eta", "sigma_squared", "nu_1", "nu_2",
                                   "pt_fraction")]


class Draws(C.Structure):
    _fields_ = [("lx", C.c_void_p), ("z", C.c_void_p), ("z_eta", C.c_void_p), ("u", C.c_void_p),
                ("u_swap", C.c_void_p), ("n", C.c_int32), ("n_swap_rounds", C.c_int32)]


class Traces(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("pos_w", "lik_prop", "rmse_train", "rmse_test", "acc_train", "acc_test",
                                          "accept_list", "prior_prop", "diff_prop", "mh_prob", "accepted")]


class TraceViews(C.Structure):
    _fields_ = [("pos_w", C.POINTER(C.c_float))] + \
        [(n, C.POINTER(C.c_double)) for n in ("lik_prop", "rmse_train", "rmse_test", "acc_train", "acc_test")] + \
        [("accept_list", C.POINTER(C.c_int32)), ("first", C.c_int32), ("count", C.c_int32)]


class Summary(C.Structure):
    _fields_ = [("n", C.c_int64), ("rmse_train", C.c_double * 4), ("rmse_test", C.c_double * 4),
                ("acc_train", C.c_double * 4), ("acc_test", C.c_double * 4), ("w_mean", C.c_void_p),
                ("w_std", C.c_void_p), ("kernel_ms", C.c_double), ("bytes_read", C.c_int64)]


_lib = None


def load():
    """dlopen libptfnn.so (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise PtfnnError(E_STATE, "%s not found: build it with `python __graft_entry__.py` "
                                  "(there is no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    lib.ptfnn_build_info.restype = C.c_char_p
    lib.ptfnn_last_error.restype = C.c_char_p
    lib.ptfnn_last_error.argtypes = [C.c_void_p]
    lib.ptfnn_default_config.restype = None
    for name in SYMBOLS:
        getattr(lib, name)          # AttributeError here = header / library mismatch
    if lib.ptfnn_abi_version() != ABI_VERSION:
        raise PtfnnError(E_INVALID, "ABI mismatch: library %d, binding %d" % (lib.ptfnn_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(rc, handle=None):
    if rc != OK:
        msg = load().ptfnn_last_error(handle)
        raise PtfnnError(rc, (msg or b"").decode("utf-8", "replace"))


def build_info() -> str:
    return load().ptfnn_build_info().decode()


# ---- topologies compiled on demand -------------------------------------------------------------
_jit_libs = {}


def has_topology(task, topology) -> bool:
    I, H, O = (int(x) for x in topology)
    return bool(load().ptfnn_has_topology(int(task), I, H, O))


def ensure_topology(task, topology, verbose=False):
    """The kernels are compile-time specialisations of [I, H, O] (csrc/ptfnn_topologies.h lists the ones
    built into libptfnn.so).  Any other topology is compiled here, once, from the same sources into
    csrc/build/jit/libptfnn_topo_<name>.so (nvcc, sm_100a, ~20 s; cached on disk) and registered with
    the library.  Needs nvcc; hidden layers wider than MAX_HIDDEN units are not supported."""
    import hashlib
    import subprocess
    import uuid
    task = int(task)
    I, H, O = (int(x) for x in topology)
    lib = load()
    if lib.ptfnn_has_topology(task, I, H, O):
        return
    if task == TASK_REGRESSION and O != 1:
        raise PtfnnError(E_UNSUPPORTED, "regression needs one output (R:132)")
    if H > MAX_HIDDEN or O > 32 or min(I, H, O) < 1:
        raise PtfnnError(E_UNSUPPORTED, "no specialisation for topology [%d,%d,%d] (hidden layers up to %d units, up to 32 outputs: "
                                        "the weights of a layer live in the registers of one CTA)" % (I, H, O, MAX_HIDDEN))
    name = "jit_%s_%d_%d_%d" % ("reg" if task == TASK_REGRESSION else "cls", I, H, O)
    # threads per CTA: 128 (one serial warp + three likelihood warps, or the 128-thread SGD team of a wide hidden layer);
    # 160 for the tcgen05 geometry (H = 256, up to 16 outputs: four epilogue warps + the MMA warp); 256 above that
    nt = 160 if (H == 256 and O <= 16) else 128 if H <= 256 else 256
    minb = 2 if (H > 32 or I > 16) else 4
    csrc = os.path.dirname(LIB_PATH)
    out_dir = os.path.join(csrc, "build", "jit")
    os.makedirs(out_dir, exist_ok=True)
    srcs = [os.path.join(csrc, f) for f in sorted(os.listdir(csrc)) if f.endswith((".cuh", ".h", ".cu"))]
    h = hashlib.sha256()
    for f in srcs:
        h.update(open(f, "rb").read())
    so = os.path.join(out_dir, "libptfnn_topo_%s_%s.so" % (name, h.hexdigest()[:12]))
    if not os.path.exists(so):
        # a name of its own per process: the ranks of a multi-GPU launch may all compile the same new topology at once,
        # and os.replace() must publish a complete file whichever of them finishes first
        tmp = "%s.%d.%s.tmp" % (so, os.getpid(), uuid.uuid4().hex[:8])
        nvcc = os.environ.get("NVCC") or ("/usr/local/cuda/bin/nvcc" if os.path.exists("/usr/local/cuda/bin/nvcc") else "nvcc")
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
               "-shared", "-cudart", "shared", "-DPTFNN_T_STANDALONE",
               "-DPTFNN_T_NAME=%s" % name, "-DPTFNN_T_TASK=%d" % task, "-DPTFNN_T_I=%d" % I, "-DPTFNN_T_H=%d" % H,
               "-DPTFNN_T_O=%d" % O, "-DPTFNN_T_NT=%d" % nt, "-DPTFNN_T_MINB=%d" % minb,
               os.path.join(csrc, "topo_inst.cu"), "-o", tmp]
        if verbose:
            print("[ptfnn] compiling a specialisation for topology [%d,%d,%d]: %s" % (I, H, O, " ".join(cmd)), file=sys.stderr)
        try:
            r = subprocess.run(cmd, capture_output=True, text=True)
        except OSError as e:
            raise PtfnnError(E_UNSUPPORTED, "topology [%d,%d,%d] is not built into libptfnn.so and nvcc is not available "
                                            "to compile it (%s)" % (I, H, O, e))
        if r.returncode != 0:
            if os.path.exists(tmp):
                os.unlink(tmp)
            raise PtfnnError(E_UNSUPPORTED, "compiling topology [%d,%d,%d] failed:\n%s" % (I, H, O, r.stderr[-2000:]))
        os.replace(tmp, so)
    tl = C.CDLL(so)
    tl.ptfnn_topology_kernels.restype = C.c_void_p
    check(lib.ptfnn_register_kernels(C.c_void_p(tl.ptfnn_topology_kernels()), int(tl.ptfnn_topology_registry_version())))
    _jit_libs[(task, I, H, O)] = tl                     # keep the library (its kernels) loaded


def device_count() -> int:
    return int(load().ptfnn_device_count())


def default_config() -> Config:
    c = Config()
    load().ptfnn_default_config(C.byref(c))
    return c


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# ---- single operations (stateless) -------------------------------------------------------------
def op_forward_pass(topology, x, w, device=0):
    """Network.ForwardPass on one row -> (hidout[H], out[O])."""
    I, H, O = topology
    x, w = f64(x).reshape(-1), f64(w)
    hid, out = np.zeros(H), np.zeros(O)
    check(load().ptfnn_op_forward_pass(device, I, H, O, ptr(x), ptr(w), ptr(hid), ptr(out)))
    return hid, out


def op_evaluate_proposal(task, topology, data, w, device=0):
    I, H, O = topology
    ensure_topology(task, topology)
    data, w = f64(data), f64(w)
    fx = np.zeros(data.shape[0])
    prob = np.zeros((data.shape[0], O)) if task == TASK_CLASSIFICATION else None
    check(load().ptfnn_op_evaluate_proposal(device, task, I, H, O, ptr(data), data.shape[0], data.shape[1], ptr(w),
                                            ptr(fx), ptr(prob)))
    return (fx, prob) if task == TASK_CLASSIFICATION else fx


def op_langevin_gradient(task, topology, data, w, learn_rate, depth=1, device=0):
    I, H, O = topology
    ensure_topology(task, topology)
    data, w = f64(data), f64(w)
    out = np.zeros_like(w)
    check(load().ptfnn_op_langevin_gradient(device, task, I, H, O, ptr(data), data.shape[0], data.shape[1], ptr(w),
                                            C.c_double(learn_rate), int(depth), ptr(out)))
    return out


def time_langevin_gradient(task, topology, data, w, learn_rate, depth=1, repeats=3, device=0):
    """Device time (ms, CUDA events, best of ``repeats``) of the langevin_gradient kernel alone."""
    I, H, O = topology
    ensure_topology(task, topology)
    data, w = f64(data), f64(w)
    ms = C.c_double(0.0)
    check(load().ptfnn_time_langevin_gradient(device, task, I, H, O, ptr(data), data.shape[0], data.shape[1], ptr(w),
                                              C.c_double(learn_rate), int(depth), int(repeats), C.byref(ms)))
    return ms.value


def op_likelihood(task, topology, data, w, tau_sq=1.0, adapttemp=1.0, want_fx=True, device=0):
    """-> (loglik/adapttemp, rmse, accuracy, fx)"""
    I, H, O = topology
    ensure_topology(task, topology)
    data, w = f64(data), f64(w)
    out = np.zeros(3)
    fx = np.zeros(data.shape[0]) if want_fx else None
    check(load().ptfnn_op_likelihood(device, task, I, H, O, ptr(data), data.shape[0], data.shape[1], ptr(w),
                                     C.c_double(tau_sq), C.c_double(adapttemp), ptr(out), ptr(fx)))
    return float(out[0]), float(out[1]), float(out[2]), fx


def op_posterior_predictive(task, topology, data, w_samples, device=0):
    """Forward pass of every posterior sample in one batched launch -> (fx_all[n_samples, rows], sums[n_samples, 3]).
    The reference returns these arrays as zeros (R:785-788)."""
    I, H, O = topology
    ensure_topology(task, topology)
    data, w_samples = f64(data), f64(w_samples)
    ns = w_samples.shape[0]
    fx = np.empty((ns, data.shape[0]))
    sums = np.empty((ns, 3))
    check(load().ptfnn_op_posterior_predictive(device, task, I, H, O, ptr(data), data.shape[0], data.shape[1],
                                               ptr(w_samples), ns, ptr(fx), ptr(sums)))
    return fx, sums


def op_prior(task, topology, w, sigma_squared=25.0, nu_1=0.0, nu_2=0.0, tausq=1.0, device=0):
    I, H, O = topology
    w = f64(w)
    out = C.c_double(0.0)
    check(load().ptfnn_op_prior(device, task, I, H, O, ptr(w), C.c_double(sigma_squared), C.c_double(nu_1),
                                C.c_double(nu_2), C.c_double(tausq), C.byref(out)))
    return out.value


def op_swap_sweep(lhood, u_row, device=0, swap_kind=SWAP_KIND_REFERENCE, temperatures=None):
    """The coordinator's sequential sweep (R:741-748) over the lhood fields -> (src, swapped).
    swap_kind: the reference's rule (R:674) or the drafts' temperature-aware one (needs ``temperatures``)."""
    lhood, u_row = f64(lhood), f32(u_row)
    n = lhood.shape[0]
    src = np.zeros(n, dtype=np.int32)
    sw = np.zeros(max(n - 1, 1), dtype=np.uint8)
    t = None if temperatures is None else f64(temperatures)
    check(load().ptfnn_op_swap_sweep_kind(device, n, ptr(lhood), ptr(u_row), int(swap_kind), ptr(t), ptr(src), ptr(sw)))
    return src, sw[:n - 1].astype(bool)


# ---- host side of the result pipeline (no device) ------------------------------------------------
def savetxt(path, X, fmt='%.18e'):
    """np.savetxt(path, X, fmt=fmt) for float arrays of 1 or 2 dimensions, byte for byte (R:454-481)."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 0:
        X = X.reshape(1, 1)
    elif X.ndim == 1:
        X = X.reshape(-1, 1)
    elif X.ndim != 2:
        raise ValueError("Expected 1D or 2D array, got %dD array instead" % X.ndim)
    X = np.ascontiguousarray(X)
    rc = load().ptfnn_savetxt(os.fsencode(path), ptr(X), C.c_int64(X.shape[0]), C.c_int64(X.shape[1]),
                              C.c_int64(X.shape[1]), fmt.encode())
    if rc != OK:
        raise OSError("ptfnn_savetxt(%r, fmt=%r) failed (%d)" % (path, fmt, rc))


def loadtxt(path):
    """np.loadtxt(path) for the files savetxt writes (R:794-831): (rows, cols) -> 2-D, one column -> 1-D."""
    rows, cols = C.c_int64(), C.c_int64()
    try:
        cap = os.path.getsize(path) // 2 + 1            # a value takes at least a digit and a separator
    except OSError as e:
        raise OSError("ptfnn_loadtxt(%r): %s" % (path, e)) from None
    flat = np.empty(cap)                                # (untouched pages cost nothing)
    rc = load().ptfnn_loadtxt(os.fsencode(path), ptr(flat), C.c_int64(cap), C.byref(rows), C.byref(cols))
    if rc != OK:
        raise OSError("ptfnn_loadtxt(%r) failed (%d)" % (path, rc))
    if rows.value == 0:
        return np.empty(0)
    out = flat[:rows.value * cols.value].reshape(rows.value, cols.value).copy()
    if out.shape[1] == 1:
        return out.reshape(()) if out.shape[0] == 1 else out[:, 0]
    return out[0] if out.shape[0] == 1 else out
