"""GPU: the result pipeline on the device traces.  ptfnn_trace_summary pools the burn-in
slice of every replica (R:777, R:797-824) and reduces it on the device; the oracle here is the
reference's own arithmetic for that step -- np.mean / np.std / np.amin / np.amax over the float64
traces (R:1036-1044, C:1130-1136), applied to what ptfnn_get_traces returns for the same run."""
import os

import numpy as np
import pytest

from oracle import ptfnn_numpy as on
from ptnn_b200 import classification as cls
from ptnn_b200 import regression as reg
from ptnn_b200._surface import RESULT_DIRS
from ptnn_b200.sampler import Sampler, geometric_ladder
from tests import common as cm

pytestmark = pytest.mark.gpu
RTOL = 1e-9          # fp64 sums in a different order than NumPy's pairwise / two-pass reductions


def _check(sm, t, first, count):
    sl = slice(first, first + count)
    assert sm["n"] == t["rmse_train"][:, sl].size
    for k in ("rmse_train", "rmse_test", "acc_train", "acc_test"):
        x = t[k][:, sl]
        want = [np.mean(x), np.std(x), np.amin(x), np.amax(x)]
        got = [sm[k][q] for q in ("mean", "std", "min", "max")]
        assert np.allclose(got, want, rtol=RTOL, atol=1e-12), (k, got, want)
    pw = t["pos_w"][:, sl].reshape(-1, t["pos_w"].shape[2])
    assert np.allclose(sm["w_mean"], pw.mean(axis=0), rtol=RTOL, atol=1e-12)
    assert np.allclose(sm["w_std"], pw.std(axis=0), rtol=1e-7, atol=1e-10)


@pytest.mark.parametrize("task,ds,topo,R,S,first,count", [
    (on.REGRESSION, "Sunspot", (4, 5, 1), 10, 120, 60, 60),          # P = 31: eight rows side by side per block
    (on.REGRESSION, "Sunspot", (4, 5, 1), 3, 41, 7, 29),             # ragged: odd counts, slice inside the trace
    (on.REGRESSION, "Sunspot", (4, 5, 1), 1, 12, 11, 1),             # a single pooled row: std = 0
    (on.REGRESSION, "Mackey", (4, 64, 1), 6, 40, 0, 40),             # P = 385: two column tiles per thread; row 0 = ones (Q12)
    (on.CLASSIFICATION, "Ionosphere", (34, 50, 2), 4, 30, 15, 15),   # P = 1852: two blocks of columns (blockIdx.y)
    (on.CLASSIFICATION, "Iris", (4, 12, 3), 5, 50, 25, 25),          # accuracies are filled (C:414)
])
def test_trace_summary_matches_numpy_on_the_traces(task, ds, topo, R, S, first, count):
    tr, te = cm.dataset(task, ds)
    P = topo[0] * topo[1] + topo[1] * topo[2] + topo[1] + topo[2]
    w0 = np.random.RandomState(3).randn(R, P) * 0.5
    with Sampler(task, topo, geometric_ladder(R, 4), S, 5, learn_rate=0.05, l_prob=0.5, seed=9) as s:
        s.set_data(tr, te)
        s.init_chains(w0)
        s.run()
        sm = s.trace_summary(first, count)
        t = s.traces()
        # the same handle again with other grids and slices: series only, a one-row slice, the whole trace
        sm2 = s.trace_summary(first, count, posterior=False)
        sm3 = s.trace_summary(S - 1, 1)
        sm4 = s.trace_summary(0, S, posterior=False)
        sm5 = s.trace_summary(0, S)
    _check(sm, t, first, count)
    assert sm2["w_mean"] is None and sm2["rmse_test"] == sm["rmse_test"] and sm2["acc_train"] == sm["acc_train"]
    _check(sm3, t, S - 1, 1)
    _check(sm5, t, 0, S)
    assert sm4["rmse_train"] == sm5["rmse_train"] and sm4["n"] == R * S
    assert sm["bytes_read"] == R * count * (32 + 4 * P) and sm["kernel_ms"] > 0


def test_trace_summary_rejects_bad_slices():
    from ptnn_b200.capi import PtfnnError
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    with Sampler(on.REGRESSION, (4, 5, 1), geometric_ladder(2, 2), 10, 5) as s:
        s.set_data(tr, te)
        with pytest.raises(PtfnnError):
            s.trace_summary(0, 10)                         # no chains yet
        s.init_chains(np.zeros((2, 31)))
        for first, count in ((-1, 3), (0, 0), (5, 6)):
            with pytest.raises(PtfnnError):
                s.trace_summary(first, count)


def _pt(mod, task, ds, topo, tmp, R, S, swap, seed, **attrs):
    tr, te = cm.dataset(task, ds)
    path = str(tmp)
    args = (True, 0.05, tr, te, list(topo), R, 4, R * S, swap)
    pt = mod.ParallelTempering(*args, 0.5, path) if task == on.REGRESSION else mod.ParallelTempering(*args, path)
    for k, v in attrs.items():
        setattr(pt, k, v)
    for d in RESULT_DIRS:
        pt.make_directory(path + d)
    np.random.seed(seed)
    pt.initialize_chains(0.5)
    return pt


@pytest.mark.parametrize("mod,task,ds,topo", [(reg, on.REGRESSION, "Lazer", (4, 5, 1)),
                                              (cls, on.CLASSIFICATION, "Cancer", (9, 12, 2))])
def test_run_summary_equals_main_statistics_of_run_chains(tmp_path, mod, task, ds, topo):
    """run_summary() (no traces leave the GPU) reports what the reference's main() computes from the
    run_chains() tuple of the same seeded run."""
    a = _pt(mod, task, ds, topo, tmp_path / "a", 4, 80, 10, 6, write_files=False, results_from_files=False)
    res = a.run_chains()
    b = _pt(mod, task, ds, topo, tmp_path / "b", 4, 80, 10, 6)
    sm = b.run_summary()
    rmse_train, rmse_test, acc_train, acc_test, swap_perc, accept_vec = res[3], res[4], res[5], res[6], res[8], res[9]
    for k, x in (("rmse_train", rmse_train), ("rmse_test", rmse_test), ("acc_train", acc_train), ("acc_test", acc_test)):
        want = [np.mean(x), np.std(x), np.amin(x), np.amax(x)]                          # R:1036-1044 / C:1130-1136
        assert np.allclose([sm[k][q] for q in ("mean", "std", "min", "max")], want, rtol=RTOL, atol=1e-12)
    n = accept_vec.shape[1]
    assert sm["accept_per"] == pytest.approx(np.mean(accept_vec[:, n - 1:n] / n) * 100)  # R:1009-1011
    assert sm["swap_perc"] == pytest.approx(swap_perc)
    assert np.allclose(sm["w_mean"], res[0].mean(axis=1), rtol=RTOL, atol=1e-12)        # posterior[P, R*(S-burn)]
    assert a.summary["rmse_test"] == sm["rmse_test"]                                     # run_chains leaves it too
    assert not any(fs for _, _, fs in os.walk(str(tmp_path / "b")))                      # and nothing was written


def test_run_problem_with_device_results(tmp_path):
    root = tmp_path / "data" / "Data_OneStepAhead" / "Sunspot"
    os.makedirs(root)
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    np.savetxt(root / "train.txt", tr); np.savetxt(root / "test.txt", te)
    rows = {}
    for mode in ("host", "device"):
        out = tmp_path / mode
        os.makedirs(out)
        np.random.seed(4)
        rows[mode], _ = reg.run_problem(2, str(tmp_path / "data"), str(out), NumSample=600, num_chains=4,
                                        swap_ratio=0.05, seed=21, results=mode)
        assert len(open(out / "master_result_file.txt").read().split()) == 16            # 15 numbers + run name (R:1052-1061)
    assert np.allclose(rows["host"][:14], rows["device"][:14], rtol=1e-4, atol=1e-7)   # host row: statistics of the %1.8f txt files


@pytest.mark.parametrize("ds,topo,R,S,first,count", [("Sunspot", (4, 5, 1), 4, 40, 20, 20), ("Lazer", (4, 10, 1), 3, 25, 3, 17),
                                                   ("Mackey", (4, 64, 1), 2, 12, 6, 6)])
def test_predictive_moments_from_device_traces_match_oracle(ds, topo, R, S, first, count):
    """ptfnn_predictive_summary == np.mean / np.std over the per-sample predictions the float64 oracle computes for
    every recorded weight vector of the slice (what fx_train_all / fx_test_all would hold, R:785-788, R:809-815)."""
    from oracle import ptfnn_c as oc
    tr, te = cm.dataset(on.REGRESSION, ds)
    P = topo[0] * topo[1] + topo[1] * topo[2] + topo[1] + topo[2]
    w0 = np.random.RandomState(4).randn(R, P) * 0.5
    with Sampler(on.REGRESSION, topo, geometric_ladder(R, 3), S, 5, learn_rate=0.05, l_prob=0.5, seed=3) as s:
        s.set_data(tr, te)
        s.init_chains(w0)
        s.run()
        got = {k: s.predictive_summary(k, first, count) for k in ("train", "test")}
        pw = s.traces()["pos_w"][:, first:first + count].reshape(-1, P)
    for k, data in (("train", tr), ("test", te)):
        fx = np.stack([oc.evaluate(on.REGRESSION, topo, data, w) for w in pw])
        assert np.allclose(got[k]["mean"], fx.mean(axis=0), rtol=1e-4, atol=1e-6)
        assert np.allclose(got[k]["std"], fx.std(axis=0), rtol=1e-3, atol=1e-6)
        assert got[k]["rmse_of_mean"] == pytest.approx(np.sqrt(np.mean((fx.mean(axis=0) - data[:, -1]) ** 2)), rel=1e-4)


def test_predictive_moments_are_regression_only():
    from ptnn_b200.capi import PtfnnError
    tr, te = cm.dataset(on.CLASSIFICATION, "Iris")
    with Sampler(on.CLASSIFICATION, (4, 12, 3), geometric_ladder(2, 2), 10, 5) as s:
        s.set_data(tr, te)
        s.init_chains(np.zeros((2, 99)))
        with pytest.raises(PtfnnError) as e:
            s.predictive_summary("test", 0, 5)
        assert e.value.code == -4                                   # PTFNN_E_UNSUPPORTED


def test_run_chains_predictive_moments_agree_with_the_per_sample_predictions(tmp_path):
    """Two routes to SURVEY 8(f).2 on the same run: pt.posterior_predictive fills fx_train_all / fx_test_all
    (one batched pass, arrays on the host); pt.predictive_moments reduces the same predictions on the device."""
    pt = _pt(reg, on.REGRESSION, "Sunspot", (4, 5, 1), tmp_path, 4, 60, 10, 2, write_files=False, results_from_files=False,
             posterior_predictive=True, predictive_moments=True)
    res = pt.run_chains()
    for k, fx_all, data in (("train", res[1], pt.traindata), ("test", res[2], pt.testdata)):
        fx = fx_all.reshape(-1, data.shape[0])
        assert fx.shape[0] == 4 * 30 and np.any(fx != 0)
        assert np.allclose(pt.predictive[k]["mean"], fx.mean(axis=0), rtol=1e-6, atol=1e-9)
        assert np.allclose(pt.predictive[k]["std"], fx.std(axis=0), rtol=1e-5, atol=1e-9)


@pytest.mark.parametrize("ds,topo,R,S,first,count,q", [("Sunspot", (4, 5, 1), 4, 60, 20, 40, (5.0, 95.0)),
                                                     ("Lazer", (4, 10, 1), 3, 25, 3, 17, (2.5, 97.5)),
                                                     ("Mackey", (4, 64, 1), 2, 12, 6, 6, (0.0, 100.0)),
                                                     ("Henon", (4, 5, 1), 1, 9, 8, 1, (50.0, 50.0))])
def test_predictive_bands_match_numpy_percentile(ds, topo, R, S, first, count, q):
    """the 5 % - 95 % band of the posterior-predictive distribution per data row.  The device selects
    exact order statistics of the [samples, rows] prediction matrix (radix select), so the band equals
    np.percentile (linear interpolation) over the very predictions the batched forward pass returns -- including
    ties (a rejected step repeats the previous row of pos_w), a single sample and the 0 / 100 % ends."""
    from ptnn_b200 import capi
    tr, te = cm.dataset(on.REGRESSION, ds)
    P = topo[0] * topo[1] + topo[1] * topo[2] + topo[1] + topo[2]
    w0 = np.random.RandomState(4).randn(R, P) * 0.5
    with Sampler(on.REGRESSION, topo, geometric_ladder(R, 3) if R > 1 else np.ones(1), S, 5, learn_rate=0.05, l_prob=0.5, seed=3) as s:
        s.set_data(tr, te)
        s.init_chains(w0)
        s.run()
        got = {k: s.predictive_summary(k, first, count, bands=q) for k in ("train", "test")}
        pw = s.traces()["pos_w"][:, first:first + count].reshape(-1, P)
    for k, data in (("train", tr), ("test", te)):
        fx = capi.op_posterior_predictive(on.REGRESSION, topo, data, pw)[0]          # the same float32 predictions, on the host
        lo, hi = np.percentile(fx, q, axis=0)
        assert np.allclose(got[k]["lo"], lo, rtol=1e-12, atol=1e-12), (k, np.max(np.abs(got[k]["lo"] - lo)))
        assert np.allclose(got[k]["hi"], hi, rtol=1e-12, atol=1e-12), (k, np.max(np.abs(got[k]["hi"] - hi)))
        assert np.all(got[k]["lo"] <= got[k]["mean"] + 1e-9) or q[0] > 50
        assert np.all(got[k]["lo"] <= got[k]["hi"])


def test_run_chains_predictive_bands(tmp_path):
    pt = _pt(reg, on.REGRESSION, "Sunspot", (4, 5, 1), tmp_path, 4, 60, 10, 2, write_files=False, results_from_files=False,
             posterior_predictive=True, predictive_moments=True, predictive_bands=(5, 95))
    res = pt.run_chains()
    for k, fx_all, data in (("train", res[1], pt.traindata), ("test", res[2], pt.testdata)):
        fx = fx_all.reshape(-1, data.shape[0])
        lo, hi = np.percentile(fx, (5, 95), axis=0)
        assert np.allclose(pt.predictive[k]["lo"], lo, rtol=1e-6, atol=1e-9)
        assert np.allclose(pt.predictive[k]["hi"], hi, rtol=1e-6, atol=1e-9)
