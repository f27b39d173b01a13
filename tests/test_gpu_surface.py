"""GPU: the drop-in class surface end to end -- per-method parity with the reference's known
answers, the output files of run_chains() (names, shapes, formats), run-to-run
determinism, and free-running statistics inside the reference's ranges."""
import os

import numpy as np
import pytest

from oracle import ptfnn_numpy as on
from ptnn_b200 import classification as cls
from ptnn_b200 import regression as reg
from ptnn_b200._surface import RESULT_DIRS
from tests import common as cm

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def test_regression_methods_match_reference_known_answers():
    ka = cm.npz("known_answers")
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    key = "reg_Sunspot_h5_"
    w, tau = ka[key + "w"], float(ka[key + "tau"])
    net = reg.Network([4, 5, 1], tr, te, 0.1)
    rep = reg.ptReplica(True, 0.1, w, None, None, 10, tr, te, [4, 5, 1], 0.5, 1.25, 5, 0.5, "", None, None, None)
    assert cm.relerr(net.evaluate_proposal(tr, w), ka[key + "fx"]) < RTOL
    lik, fx, rm = rep.likelihood_func(net, tr, w, tau)
    assert np.allclose([lik, rm], ka[key + "lik"][:2], rtol=RTOL)
    assert rep.prior_likelihood(25, 0, 0, w, tau) == pytest.approx(float(ka[key + "prior"]), rel=RTOL)
    wc = w.copy()
    out = net.langevin_gradient(tr, wc, 1)
    assert cm.relerr(out, ka[key + "w_gd"]) < RTOL and np.array_equal(wc, out)       # in-place, like R:101-116
    # ForwardPass / BackwardPass on one row == first row of the SGD epoch
    net.decode(w.copy())
    net.ForwardPass(tr[0, :4])
    ref = on.Network((4, 5, 1), 0.1, on.REGRESSION)
    ref.decode(w.copy()); ref.forward(tr[0, :4])
    assert cm.relerr(net.out, ref.out) < RTOL and cm.relerr(net.hidout, ref.hidout) < RTOL
    net.BackwardPass(tr[0, :4], tr[0, 4:])
    ref.backward(tr[0, :4], tr[0, 4:])
    assert cm.relerr(net.encode(), np.concatenate([ref.W1.ravel(), ref.W2.ravel(), ref.B1, ref.B2])) < RTOL


def test_classification_methods_match_reference_known_answers():
    ka = cm.npz("known_answers")
    tr, te = cm.dataset(on.CLASSIFICATION, "Cancer")
    key = "cls_Cancer_"
    w = ka[key + "w"]
    net = cls.Network([9, 12, 2], tr, te, 0.01)
    rep = cls.ptReplica(True, 0.01, w, None, None, 10, tr, te, [9, 12, 2], 0.5, 2.5, 5, "", None, None, None)
    fx, prob = net.evaluate_proposal(tr, w)
    assert cm.relerr(prob, ka[key + "prob"]) < RTOL
    lik, fx2, rm = rep.likelihood_func(net, tr, w)
    assert lik == pytest.approx(float(ka[key + "lik"][0]), rel=RTOL)
    assert rep.accuracy(fx2, tr[:, 9]) == pytest.approx(float(ka[key + "acc"][0]), abs=0.5)
    assert rep.prior_likelihood(25, 0, 0, w) == pytest.approx(float(ka[key + "prior"]), rel=RTOL)
    assert cm.relerr(net.langevin_gradient(tr, w.copy(), 1), ka[key + "w_gd"]) < RTOL


def _run(mod, task, ds, topo, tmp, R=4, S=60, swap=10, maxtemp=2, use_lg=True, lr=0.1, seed=5, **attrs):
    tr, te = cm.dataset(task, ds)
    path = str(tmp)
    args = (use_lg, lr, tr, te, topo, R, maxtemp, R * S, swap)
    pt = mod.ParallelTempering(*args, 0.5, path) if task == on.REGRESSION else mod.ParallelTempering(*args, path)
    for k, v in attrs.items():
        setattr(pt, k, v)
    for d in RESULT_DIRS:
        pt.make_directory(path + d)
    np.random.seed(seed)
    pt.initialize_chains(0.5)
    return pt, pt.run_chains()


def test_regression_run_chains_outputs_and_files(tmp_path):
    fx = cm.npz("reg_sunspot_lg")
    pt, res = _run(reg, on.REGRESSION, "Sunspot", [4, 5, 1], tmp_path, R=4, S=100, swap=10)
    (pos_w, fx_train, fx_test, rmse_train, rmse_test, acc_train, acc_test, lik, swap_perc, accept_vec, accept) = res
    shapes = [str(getattr(x, "shape", ())) for x in res]
    assert shapes == [str(s) for s in fx["ref_result_shapes"]]                     # same 11-tuple shapes as R:771
    assert accept == 0 and 0 <= swap_perc <= 100                                     # R:860 accept is always 0
    assert pt.total_swap_proposals == int(fx["ref_total_swap_proposals"])            # 10 rounds x 3 pairs (Q9)
    files = sorted(os.path.relpath(os.path.join(dp, f), str(tmp_path)) for dp, _, fs in os.walk(str(tmp_path)) for f in fs)
    assert files == sorted(str(f) for f in fx["ref_files"])                          # same files as the reference run
    T = str(pt.temperatures[1])
    pw = np.loadtxt(os.path.join(str(tmp_path), "posterior/pos_w/chain_%s.txt" % T))
    assert pw.shape == (100, 31) and np.all(pw[0] == 1.0)                            # Q12
    lk = np.loadtxt(os.path.join(str(tmp_path), "posterior/pos_likelihood/chain_%s.txt" % T))
    assert lk.shape == (100, 2) and lk[0, 0] == -100 and np.all(lk[1:, 1] == 0)
    acc = np.loadtxt(os.path.join(str(tmp_path), "posterior/accept_list/chain_%s.txt" % T))
    assert acc.shape == (100,) and np.all(np.diff(acc) >= 0)
    assert open(os.path.join(str(tmp_path), "acceptpercent.txt")).read().strip() == "0.00"
    assert os.path.getsize(os.path.join(str(tmp_path), "num_exchange.txt")) == 0     # R:704
    assert np.all(rmse_train >= 0) and np.all(acc_train == 0)


def test_run_chains_is_deterministic_and_memoisation_is_exact(tmp_path):
    a = _run(reg, on.REGRESSION, "Lazer", [4, 5, 1], tmp_path / "a", seed=3, write_files=False, results_from_files=False)[1]
    b = _run(reg, on.REGRESSION, "Lazer", [4, 5, 1], tmp_path / "b", seed=3, write_files=False, results_from_files=False)[1]
    c = _run(reg, on.REGRESSION, "Lazer", [4, 5, 1], tmp_path / "c", seed=3, write_files=False, results_from_files=False,
             memoize_gradient=False)[1]
    d = _run(reg, on.REGRESSION, "Lazer", [4, 5, 1], tmp_path / "d", seed=4, write_files=False, results_from_files=False)[1]
    for x, y, z in zip(a, b, c):
        assert np.array_equal(np.asarray(x), np.asarray(y))          # same seed -> bit-identical
        assert np.array_equal(np.asarray(x), np.asarray(z))          # memoised langevin_gradient(w) changes nothing
    assert not np.array_equal(a[0], d[0])


def test_classification_run_chains(tmp_path):
    fx = cm.npz("cls_iris_lg")
    pt, res = _run(cls, on.CLASSIFICATION, "Iris", [4, 12, 3], tmp_path, R=4, S=80, swap=8, maxtemp=10, lr=0.01)
    shapes = [str(getattr(x, "shape", ())) for x in res]
    assert shapes == [str(s) for s in fx["ref_result_shapes"]]
    assert pt.total_swap_proposals == int(fx["ref_total_swap_proposals"])
    acc_train = res[5]
    assert acc_train.max() <= 100.0 and acc_train.max() > 30.0
    T = str(pt.temperatures[0])
    rm = open(os.path.join(str(tmp_path), "predictions/rmse_train_chain_%s.txt" % T)).readline().strip()
    assert len(rm.split(".")[1]) == 2                                                 # '%1.2f' (C:473-475)


def test_free_running_statistics_match_reference_spread():
    """Free-running (Philox) mode at the BASELINE config -- Sunspot 4-5-1, 10 replicas, maxtemp 2,
    Langevin l_prob 0.5, swap_interval 50 (2000 samples per replica here).  Acceptance rate, swap rate
    and posterior RMSE must fall inside the run-to-run spread of the reference algorithm, measured
    by the float64 oracle on (a) the very draws the device generated and (b) independent seeds."""
    from oracle import ptfnn_c as oc
    from ptnn_b200.sampler import Sampler
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    R, S, si = 10, 2000, 50
    cfg = on.PTConfig(task=on.REGRESSION, topology=(4, 5, 1), samples=S, swap_interval=si,
                      use_langevin_gradients=True, l_prob=0.5, learn_rate=0.1)
    temps = on.geometric_ladder(R, 2)

    def stats(rmse_train, rmse_test, accept_last, ns, tot):
        b = S // 2
        return np.array([rmse_train[:, b:].mean(), rmse_test[:, b:].mean(), np.mean(accept_last / S) * 100, 100.0 * ns / tot])

    dev, same, other = [], [], []
    for seed in (1, 2, 3):
        w0 = np.random.RandomState(seed).randn(R, cfg.P)
        with Sampler(on.REGRESSION, (4, 5, 1), temps, S, si, learn_rate=0.1, l_prob=0.5, seed=seed) as s:
            s.set_data(tr, te)
            s.init_chains(w0)
            lx, z, ze, u = s.generate_draws(0, S - 1)
            us = np.stack([s.swap_uniforms(r) for r in range(cfg.total_rounds())])
            assert s.run() == S - 1
            t = s.traces(pos_w=False)
            ns, tot, _ = s.swap_stats()
        dev.append(stats(t["rmse_train"], t["rmse_test"], t["accept_list"][:, -1], ns, tot))
        assert np.all(lx == lx[0]) and np.all(z == z[0]) and not np.all(u == u[0])      # SURVEY Q10 semantics
        ref = oc.run_pt(cfg, tr, te, temps, w0, on.Draws(lx=lx, z=z, z_eta=ze, u=u, u_swap=us), with_state=False)
        same.append(stats(ref.rmse_train, ref.rmse_test, ref.accept_list[:, -1], ref.num_swap, ref.total_swap_proposals))
        assert tot == ref.total_swap_proposals
        ind = oc.run_pt(cfg, tr, te, temps, w0, on.random_draws(cfg, R, 100 + seed), with_state=False)
        other.append(stats(ind.rmse_train, ind.rmse_test, ind.accept_list[:, -1], ind.num_swap, ind.total_swap_proposals))
    dev, same, other = np.array(dev), np.array(same), np.array(other)
    # (a) same draws: the device chain IS the oracle chain up to rare near-tie flips
    assert np.all(np.abs(dev[:, :2] - same[:, :2]) < 0.35 * same[:, :2] + 0.01), (dev, same)
    assert np.all(np.abs(dev[:, 2] - same[:, 2]) < 3.0) and np.all(np.abs(dev[:, 3] - same[:, 3]) < 12.0), (dev, same)
    # (b) inside the oracle's run-to-run spread over all six oracle runs (with margin)
    allref = np.vstack([same, other])
    lo, hi = allref.min(axis=0), allref.max(axis=0)
    span = np.maximum(hi - lo, np.array([0.02, 0.02, 2.0, 8.0]))
    assert np.all(dev.mean(axis=0) > lo - span) and np.all(dev.mean(axis=0) < hi + span), (dev, allref)


def test_swap_procedure_object_api():
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    pt = reg.ParallelTempering(True, 0.1, tr, te, [4, 5, 1], 2, 2, 100, 10, 0.5, "")
    P = pt.num_param

    class Q:
        def __init__(self, v): self.v = v
        def get(self): return self.v
    p1 = np.concatenate([np.zeros(P), [0.1], [-50.0], [1.0]])
    p2 = np.concatenate([np.ones(P), [0.2], [50.0], [2.0]])
    np.random.seed(0)
    a, b, swapped = pt.swap_procedure(Q(p1), Q(p2))                                   # l2 - l1 = 100 -> p = 1
    assert swapped and np.array_equal(a, p2) and np.array_equal(b, p1)
    assert (pt.num_swap, pt.total_swap_proposals) == (1, 1)
    a, b, swapped = pt.swap_procedure(Q(p2), Q(p1))                                   # l2 - l1 = -100 -> p ~ 0
    assert not swapped and (pt.num_swap, pt.total_swap_proposals) == (1, 2)


def test_integration_md_binding_stub_runs():
    """The ctypes stub INTEGRATION.md shows a reference maintainer (section 2) is executable as printed: it is
    extracted from the document and driven with a minimal stand-in for ParallelTempering."""
    import re
    import types
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = [b for b in re.findall(r"```python\n(.*?)```", text, flags=re.S) if "def run_chains(self)" in b]
    assert len(blocks) == 1
    ns = {}
    cwd = os.getcwd()
    os.chdir(root)                                      # the stub opens the library by its path inside the repository
    try:
        exec(compile(blocks[0], "INTEGRATION.md", "exec"), ns)
        tr, te = cm.dataset(on.REGRESSION, "Sunspot")
        R, S = 4, 60
        rs = np.random.RandomState(1)
        pt = types.SimpleNamespace(topology=[4, 5, 1], num_chains=R, NumSamples=S, swap_interval=10, use_langevin_gradients=True,
                                   langevin_prob=0.5, learn_rate=0.1, temperatures=list(on.geometric_ladder(R, 2)),
                                   traindata=tr, testdata=te, num_param=31, num_swap=0, total_swap_proposals=0,
                                   chains=[types.SimpleNamespace(w=rs.randn(31)) for _ in range(R)])
        np.random.seed(3)
        ns["run_chains"](pt)
    finally:
        os.chdir(cwd)
    assert pt.total_swap_proposals == 6 * (R - 1) and 0 <= pt.num_swap <= pt.total_swap_proposals   # S/swap_interval rounds (Q9)


def test_classification_cli_like_run_sh(tmp_path, monkeypatch):
    """`python pt_classification.py <swap_ratio>` (run.sh:8-11, C:1039) through ptnn_b200.classification.main on an
    iris.csv laid out like the reference's DATA/iris.csv (';' separated, labels 1..3): default NumSample 50 000,
    10 chains, the result row appended to master_result_file.txt (C:1138-1146)."""
    tr, te = cm.dataset(on.CLASSIFICATION, "Iris")
    data = np.vstack([tr, te])
    data[:, 4] += 1
    os.makedirs(tmp_path / "root" / "DATA")
    np.savetxt(tmp_path / "root" / "DATA" / "iris.csv", data, delimiter=";")
    monkeypatch.setenv("PT_DATA_ROOT", str(tmp_path / "root"))
    monkeypatch.setenv("PT_OUT_ROOT", str(tmp_path / "out"))
    np.random.seed(0)
    cls.main(["0.02", "3"])
    row = open(tmp_path / "out" / "master_result_file.txt").read().split()
    assert len(row) == 16 and row[-1] == "iris_0"
    vals = [float(v) for v in row[:15]]
    assert vals[0] == 3 and vals[1] == 50000 and vals[3] == 100                       # problem, NumSample, swap_interval = 0.02 * 50000 / 10
    assert 50.0 < vals[6] <= 100.0 and 50.0 < vals[9] <= 100.0                        # mean train / test accuracy of the pooled posterior
    files = [f for _, _, fs in os.walk(tmp_path / "out" / "iris_0") for f in fs]
    assert len(files) == 10 * 8 + 5                                                   # 8 files per chain + 4 aggregates + result.txt


def test_free_running_statistics_classification_match_reference_spread():
    """The classification sampler (C:313-448) in free-running mode -- Iris 4-12-3, 10 replicas, maxtemp 10, Langevin
    lr 0.01, swap interval 100 (C:1036-1045; 1000 samples per replica here) -- against the float64 oracle on the
    device's own draws and on independent seeds: acceptance rate, swap rate and the posterior's mean train / test
    accuracy must lie inside the oracle's run-to-run spread."""
    from oracle import ptfnn_c as oc
    from ptnn_b200.sampler import Sampler
    tr, te = cm.dataset(on.CLASSIFICATION, "Iris")
    R, S, si = 10, 1000, 100
    cfg = on.PTConfig(task=on.CLASSIFICATION, topology=(4, 12, 3), samples=S, swap_interval=si,
                      use_langevin_gradients=True, l_prob=0.5, learn_rate=0.01)
    temps = on.geometric_ladder(R, 10)

    def stats(acc_train, acc_test, accept_last, ns, tot):
        b = S // 2
        return np.array([acc_train[:, b:].mean(), acc_test[:, b:].mean(), np.mean(accept_last / S) * 100, 100.0 * ns / max(tot, 1)])

    dev, same, other = [], [], []
    for seed in (1, 2, 3):
        w0 = np.random.RandomState(seed).randn(R, cfg.P)
        with Sampler(on.CLASSIFICATION, (4, 12, 3), temps, S, si, learn_rate=0.01, l_prob=0.5, seed=seed) as s:
            s.set_data(tr, te)
            s.init_chains(w0)
            lx, z, ze, u = s.generate_draws(0, S - 1)
            us = np.stack([s.swap_uniforms(r) for r in range(cfg.total_rounds())])
            assert s.run() == S - 1
            t = s.traces(pos_w=False)
            ns, tot, _ = s.swap_stats()
        dev.append(stats(t["acc_train"], t["acc_test"], t["accept_list"][:, -1], ns, tot))
        ref = oc.run_pt(cfg, tr, te, temps, w0, on.Draws(lx=lx, z=z, z_eta=ze, u=u, u_swap=us), with_state=False)
        same.append(stats(ref.acc_train, ref.acc_test, ref.accept_list[:, -1], ref.num_swap, ref.total_swap_proposals))
        assert tot == ref.total_swap_proposals
        ind = oc.run_pt(cfg, tr, te, temps, w0, on.random_draws(cfg, R, 100 + seed), with_state=False)
        other.append(stats(ind.acc_train, ind.acc_test, ind.accept_list[:, -1], ind.num_swap, ind.total_swap_proposals))
    dev, same, other = np.array(dev), np.array(same), np.array(other)
    # (a) the device's draws replayed by the oracle: the same chain up to rare near-tie flips
    assert np.all(np.abs(dev[:, :2] - same[:, :2]) < 8.0), (dev, same)                 # accuracy, percentage points
    assert np.all(np.abs(dev[:, 2] - same[:, 2]) < 3.0) and np.all(np.abs(dev[:, 3] - same[:, 3]) < 12.0), (dev, same)
    # (b) inside the oracle's run-to-run spread over all six oracle runs
    allref = np.vstack([same, other])
    lo, hi = allref.min(axis=0), allref.max(axis=0)
    span = np.maximum(hi - lo, np.array([5.0, 5.0, 2.0, 8.0]))
    assert np.all(dev.mean(axis=0) > lo - span) and np.all(dev.mean(axis=0) < hi + span), (dev, allref)


def test_sunspot_published_configuration_device_vs_oracle_vs_published_rows():
    """The reference's only published numbers for the north-star configuration are single unseeded runs
    (Res_LG01 / Res_LG001 / Res_RW master_result_file.txt:2 -- Sunspot, 100 000 samples,
    10 replicas, maxtemp 5, swap interval 100, Langevin l_prob 0.5, lr 0.1): train RMSE 0.0199-0.0242, test RMSE
    0.0192-0.0239, acceptance 12.6-18.3 %, swap rate 44.5-48.5 %.

    The checked-in script does not reproduce its own table: the float64 oracle -- bit-identical to the unmodified
    reference in replay (tests/test_oracle_golden.py) -- gives, pooled over the ten chains as R:1036-1044 pools them,
    RMSE ~0.11, acceptance ~4 %, swap rate ~31 % at exactly these settings, because the hot rungs (T up to 5) never
    fit the series; only its T = 1 chain sits at the published RMSE (0.021).  So the statistical bar is: (a) the
    device's pooled statistics lie inside the ORACLE's run-to-run spread at the published configuration, and (b) the
    T = 1 chain's posterior RMSE lies in the published band."""
    from oracle import ptfnn_c as oc
    from ptnn_b200.sampler import Sampler
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    R, S, si = 10, 10000, 100                                          # NumSample 100 000 / 10 chains (R:506)
    cfg = on.PTConfig(task=on.REGRESSION, topology=(4, 5, 1), samples=S, swap_interval=si,
                      use_langevin_gradients=True, l_prob=0.5, learn_rate=0.1)
    temps = on.geometric_ladder(R, 5)
    b = S // 2

    def stats(rmse_train, rmse_test, accept_last, ns, tot):
        return np.array([rmse_train[:, b:].mean(), rmse_test[:, b:].mean(), np.mean(accept_last / S) * 100, 100.0 * ns / tot,
                         rmse_train[0, b:].mean(), rmse_test[0, b:].mean()])

    dev, ora = [], []
    for seed in (1, 2, 3):
        w0 = np.random.RandomState(seed).randn(R, cfg.P)
        with Sampler(on.REGRESSION, (4, 5, 1), temps, S, si, learn_rate=0.1, l_prob=0.5, seed=seed) as s:
            s.set_data(tr, te)
            s.init_chains(w0)
            assert s.run() == S - 1
            t = s.traces(pos_w=False)
            ns, tot, _ = s.swap_stats()
        dev.append(stats(t["rmse_train"], t["rmse_test"], t["accept_list"][:, -1], ns, tot))
        ref = oc.run_pt(cfg, tr, te, temps, w0, on.random_draws(cfg, R, 50 + seed), with_state=False)
        ora.append(stats(ref.rmse_train, ref.rmse_test, ref.accept_list[:, -1], ref.num_swap, ref.total_swap_proposals))
    dev, ora = np.array(dev), np.array(ora)
    lo, hi = ora.min(axis=0), ora.max(axis=0)
    span = np.maximum(hi - lo, np.array([0.03, 0.03, 1.5, 6.0, 0.003, 0.003]))
    assert np.all(dev.mean(axis=0) > lo - span) and np.all(dev.mean(axis=0) < hi + span), (dev, ora)      # (a)
    assert 0.0199 - 0.004 < dev[:, 4].mean() < 0.0242 + 0.004, dev[:, 4]                                   # (b) train, T = 1
    assert 0.0192 - 0.004 < dev[:, 5].mean() < 0.0239 + 0.004, dev[:, 5]                                   # (b) test, T = 1
