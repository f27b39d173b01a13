"""GPU parity, per function: the CUDA path (through the C ABI) against the reference's known
answers (tests/golden/known_answers.npz, recorded from the unmodified reference) and the oracle.
Tolerance: 1e-4 relative (BASELINE north_star, fp32 device arithmetic vs the float64 reference)."""
import numpy as np
import pytest

from oracle import ptfnn_c as oc
from oracle import ptfnn_numpy as on
from ptnn_b200 import capi
from tests import common as cm

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.mark.parametrize("ds", cm.REG_DATASETS)
@pytest.mark.parametrize("H", [5, 10])
def test_regression_ops_match_reference(ds, H):
    ka = cm.npz("known_answers")
    tr, te = cm.dataset(on.REGRESSION, ds)
    key = "reg_%s_h%d_" % (ds, H)
    topo, w, tau = (4, H, 1), ka[key + "w"], float(ka[key + "tau"])
    fx = capi.op_evaluate_proposal(capi.TASK_REGRESSION, topo, tr, w)                    # R:120-134
    assert cm.relerr(fx, ka[key + "fx"]) < RTOL
    lik, rm, _, fx2 = capi.op_likelihood(capi.TASK_REGRESSION, topo, tr, w, tau, 1.25)    # R:200-205
    lik_te, rm_te, _, _ = capi.op_likelihood(capi.TASK_REGRESSION, topo, te, w, tau, 1.25)
    assert np.allclose([lik, rm, lik_te, rm_te], ka[key + "lik"], rtol=RTOL, atol=0)
    assert np.array_equal(fx, fx2)
    assert capi.op_prior(capi.TASK_REGRESSION, topo, w, tausq=tau) == pytest.approx(float(ka[key + "prior"]), rel=RTOL)
    w_gd = capi.op_langevin_gradient(capi.TASK_REGRESSION, topo, tr, w, 0.1)             # R:99-118
    assert cm.relerr(w_gd, ka[key + "w_gd"]) < RTOL


@pytest.mark.parametrize("ds", cm.CLS_DATASETS)
def test_classification_ops_match_reference(ds):
    ka = cm.npz("known_answers")
    tr, te = cm.dataset(on.CLASSIFICATION, ds)
    topo = cm.cls_topology(ds)
    key = "cls_%s_" % ds
    w = ka[key + "w"]
    fx, prob = capi.op_evaluate_proposal(capi.TASK_CLASSIFICATION, topo, tr, w)          # C:134-153
    # argmax may legitimately differ only where two sigmoid outputs tie to within fp32 round-off
    ref_prob = ka[key + "prob"]
    top2 = np.sort(ref_prob, axis=1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-5
    assert np.array_equal(fx[clear], ka[key + "fx"][clear])
    assert cm.relerr(prob, ref_prob) < RTOL
    lik, rm, acc, _ = capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, tr, w, 1.0, 2.5)   # C:209-222
    lik_te, rm_te, acc_te, _ = capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, te, w, 1.0, 2.5)
    assert np.allclose([lik, lik_te], ka[key + "lik"][[0, 2]], rtol=RTOL, atol=0)
    if clear.all():
        assert np.allclose([rm, rm_te], ka[key + "lik"][[1, 3]], rtol=RTOL, atol=0)
        assert np.allclose([acc, acc_te], ka[key + "acc"], rtol=RTOL, atol=0)
    assert capi.op_prior(capi.TASK_CLASSIFICATION, topo, w) == pytest.approx(float(ka[key + "prior"]), rel=RTOL)
    w_gd = capi.op_langevin_gradient(capi.TASK_CLASSIFICATION, topo, tr, w, 0.01)        # C:114-132
    assert cm.relerr(w_gd, ka[key + "w_gd"]) < RTOL


def test_langevin_gradient_depth_and_golden_calls():
    """langevin_gradient inputs/outputs recorded inside real reference runs (first two calls of chain 0)."""
    for name in ("reg_sunspot_lg", "reg_mackey_h10", "cls_iris_lg", "cls_cancer_lg", "cls_ions_lg"):
        fx, cfg, tr, te, _ = cm.case(name)
        for k in range(fx["ref_lg_in"].shape[0]):
            out = capi.op_langevin_gradient(cfg.task, cfg.topology, tr, fx["ref_lg_in"][k], cfg.learn_rate)
            assert cm.relerr(out, fx["ref_lg_out"][k]) < RTOL, (name, k)
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    w = np.random.RandomState(1).randn(31)
    two = capi.op_langevin_gradient(capi.TASK_REGRESSION, (4, 5, 1), tr, w, 0.1, depth=2)   # R:108 epochs
    ref = oc.langevin_gradient(on.REGRESSION, (4, 5, 1), tr, oc.langevin_gradient(on.REGRESSION, (4, 5, 1), tr, w, 0.1), 0.1)
    assert cm.relerr(two, ref) < RTOL
    assert np.array_equal(capi.op_langevin_gradient(capi.TASK_REGRESSION, (4, 5, 1), tr, w, 0.1, depth=0), w.astype(np.float32).astype(np.float64))


def test_ragged_and_tiny_inputs():
    """Row counts that are not multiples of the TMA tile (128), the row block or 4 (y padding)."""
    rs = np.random.RandomState(5)
    for n in (1, 2, 3, 5, 127, 128, 129, 257, 1000):
        data = rs.rand(n, 5)
        w = rs.randn(385) * 0.5
        fx = capi.op_evaluate_proposal(capi.TASK_REGRESSION, (4, 64, 1), data, w)
        assert cm.relerr(fx, oc.evaluate(on.REGRESSION, (4, 64, 1), data, w)) < RTOL
        lik, rm, _, _ = capi.op_likelihood(capi.TASK_REGRESSION, (4, 64, 1), data, w, 0.3, 1.7)
        l2, r2, _ = oc.likelihood(on.REGRESSION, (4, 64, 1), data, w, 0.3, 1.7)
        assert np.allclose([lik, rm], [l2, r2], rtol=RTOL)
        w_gd = capi.op_langevin_gradient(capi.TASK_REGRESSION, (4, 64, 1), data, w, 0.05)
        assert cm.relerr(w_gd, oc.langevin_gradient(on.REGRESSION, (4, 64, 1), data, w, 0.05)) < RTOL
    for n in (1, 7, 130):
        data = np.hstack([rs.randn(n, 16), rs.randint(0, 10, size=(n, 1)).astype(float)])
        w = rs.randn(on.num_params((16, 30, 10))) * 0.3
        fx, prob = capi.op_evaluate_proposal(capi.TASK_CLASSIFICATION, (16, 30, 10), data, w)
        _, prob_ref = oc.evaluate(on.CLASSIFICATION, (16, 30, 10), data, w)
        assert cm.relerr(prob, prob_ref) < RTOL
        w_gd = capi.op_langevin_gradient(capi.TASK_CLASSIFICATION, (16, 30, 10), data, w, 0.01)
        assert cm.relerr(w_gd, oc.langevin_gradient(on.CLASSIFICATION, (16, 30, 10), data, w, 0.01)) < RTOL


def test_wide_hidden_topology_ops():
    """[16,256,10] (BASELINE configs[4] shape), 300 rows."""
    rs = np.random.RandomState(9)
    topo = (16, 256, 10)
    data = np.hstack([rs.randn(300, 16), rs.randint(0, 10, size=(300, 1)).astype(float)])
    w = rs.randn(on.num_params(topo)) * 0.2
    lik, rm, acc, _ = capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, data, w, 1.0, 3.0)
    l2, r2, a2 = oc.likelihood(on.CLASSIFICATION, topo, data, w, 1.0, 3.0)
    assert lik == pytest.approx(l2, rel=RTOL)
    w_gd = capi.op_langevin_gradient(capi.TASK_CLASSIFICATION, topo, data, w, 0.01)
    assert cm.relerr(w_gd, oc.langevin_gradient(on.CLASSIFICATION, topo, data, w, 0.01)) < RTOL


def test_wide_hidden_ragged_rows():
    """The team kernel's tile logic: one row, a pair, odd / even last tiles, exactly one tile, one past it."""
    rs = np.random.RandomState(11)
    topo = (16, 256, 10)
    for n in (1, 2, 3, 127, 128, 129, 257):
        data = np.hstack([rs.randn(n, 16), rs.randint(0, 10, size=(n, 1)).astype(float)])
        w = rs.randn(on.num_params(topo)) * 0.2
        w_gd = capi.op_langevin_gradient(capi.TASK_CLASSIFICATION, topo, data, w, 0.02)
        assert cm.relerr(w_gd, oc.langevin_gradient(on.CLASSIFICATION, topo, data, w, 0.02)) < RTOL, n
        fx, prob = capi.op_evaluate_proposal(capi.TASK_CLASSIFICATION, topo, data, w)
        _, prob_ref = oc.evaluate(on.CLASSIFICATION, topo, data, w)
        assert cm.relerr(prob, prob_ref) < RTOL, n


def test_swap_sweep_matches_reference_rule():
    rs = np.random.RandomState(3)
    lh = [0.0, 10.0, 20.0, 30.0]
    src, sw = capi.op_swap_sweep(lh, [0.99, 0.99, 0.99])                                 # R:674, R:741-748 bubble
    assert src.tolist() == [1, 2, 3, 0] and sw.all()
    for n in (2, 3, 10, 257, 1024):
        lh = rs.randn(n) * 2
        u = rs.randint(0, 1 << 24, size=n - 1).astype(np.float64) / (1 << 24)
        a, b = on.swap_sweep(lh, u)
        c, d = capi.op_swap_sweep(lh, u)
        assert a == c.tolist() and b == d.tolist(), n
    # the 709 clamp (R:674): a huge gap must not overflow into NaN / no-swap
    src, sw = capi.op_swap_sweep([-1e6, 1e6], [0.999])
    assert sw.all() and src.tolist() == [1, 0]
    src, sw = capi.op_swap_sweep([1e6, -1e6], [0.0])
    assert not sw.any()


def test_on_demand_topologies_match_oracle():
    """Topologies outside csrc/ptfnn_topologies.h are compiled on first use (capi.ensure_topology): narrow,
    two-units-per-lane and wide (team kernel, H = 90, 100 and 300) hidden layers, regression and classification."""
    rs = np.random.RandomState(17)
    for task, topo in ((on.REGRESSION, (3, 7, 1)), (on.CLASSIFICATION, (6, 40, 4)), (on.CLASSIFICATION, (5, 100, 3)),
                       (on.CLASSIFICATION, (51, 90, 2)),       # the shape of the reference's Bank set (C:958-971)
                       (on.CLASSIFICATION, (6, 300, 3))):      # wider than 256: three hidden units per team thread
        I, H, O = topo
        n = 300
        y = rs.rand(n, 1) if task == on.REGRESSION else rs.randint(0, O, size=(n, 1)).astype(float)
        data = np.hstack([rs.randn(n, I) * 0.7, y])
        w = rs.randn(on.num_params(topo)) * 0.4
        lik, rm, acc, _ = capi.op_likelihood(task, topo, data, w, 0.4, 1.3)
        l2, r2, a2 = oc.likelihood(task, topo, data, w, 0.4, 1.3)
        assert lik == pytest.approx(l2, rel=RTOL), topo
        w_gd = capi.op_langevin_gradient(task, topo, data, w, 0.05)
        assert cm.relerr(w_gd, oc.langevin_gradient(task, topo, data, w, 0.05)) < RTOL, topo
        # and a short chain against the oracle on the same draws
        cfg = on.PTConfig(task=task, topology=topo, samples=12, swap_interval=4, use_langevin_gradients=True,
                          l_prob=0.5, learn_rate=0.05)
        temps = np.array([1.0, 1.4, 2.0])
        draws = on.random_draws(cfg, 3, 9, common_random_numbers=True)
        w0 = rs.randn(3, cfg.P) * 0.4
        te = data[:97]
        ref = oc.run_pt(cfg, data, te, temps, w0, draws)
        from ptnn_b200.sampler import Sampler
        with Sampler.from_oracle_config(cfg, temps, debug_traces=True) as s:
            s.set_data(data, te)
            s.init_chains(w0)
            s.replay(draws)
            t = s.traces()
        diff = np.argwhere(t["accepted"] != ref.accepted)
        i_star = int(diff[:, 1].min()) - 1 if diff.size else cfg.samples - 1
        assert i_star >= 5, (topo, i_star)
        assert cm.relerr(t["lik_prop"][:, 1:i_star + 2], ref.lik_prop[:, 1:i_star + 2]) < RTOL, topo


def test_posterior_predictive_batch_matches_oracle():
    """ptfnn_op_posterior_predictive: the forward pass of a batch of weight vectors (what the reference leaves as
    zeros in fx_train_all, R:785-788) -- CUDA-core path and the tcgen05 path of the wide-hidden net."""
    rs = np.random.RandomState(23)
    tr, _ = cm.dataset(on.REGRESSION, "Sunspot")
    ws = rs.randn(7, 31) * 0.5
    fx, sums = capi.op_posterior_predictive(capi.TASK_REGRESSION, (4, 5, 1), tr, ws)
    for k in range(7):
        assert cm.relerr(fx[k], oc.evaluate(on.REGRESSION, (4, 5, 1), tr, ws[k])) < RTOL
        assert sums[k, 0] == pytest.approx(np.sum((fx[k] - tr[:, 4]) ** 2), rel=1e-5)
    topo = (16, 256, 10)
    data = np.hstack([rs.randn(200, 16), rs.randint(0, 10, size=(200, 1)).astype(float)])
    ws = rs.randn(5, on.num_params(topo)) * 0.2
    fx, sums = capi.op_posterior_predictive(capi.TASK_CLASSIFICATION, topo, data, ws)
    for k in range(5):
        ref_fx, ref_prob = oc.evaluate(on.CLASSIFICATION, topo, data, ws[k])
        top2 = np.sort(ref_prob, axis=1)[:, -2:]
        clear = (top2[:, 1] - top2[:, 0]) > 1e-5
        assert np.array_equal(fx[k][clear], ref_fx[clear])
        assert sums[k, 0] == pytest.approx(np.sum(np.log(ref_prob[np.arange(200), data[:, 16].astype(int)])), rel=RTOL)


def test_class_labels_outside_the_outputs_are_refused_where_they_are_used():
    """C:217 indexes prob[i, int(y)] and C:73-75 builds the one-hot target from the label: the reference raises
    IndexError for a label >= n_out (and NumPy wraps a negative one).  The device refuses both instead of reading
    out of bounds; evaluate_proposal never looks at the labels (C:134-153) and keeps working."""
    from ptnn_b200.sampler import Sampler
    rs = np.random.RandomState(9)
    topo = (4, 12, 3)
    good = np.hstack([rs.randn(20, 4), rs.randint(0, 3, size=(20, 1)).astype(float)])
    w = rs.randn(on.num_params(topo)) * 0.3
    for label in (3.0, -1.0, float("nan")):
        bad = good.copy()
        bad[7, 4] = label
        for call in (lambda: capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, bad, w, 1.0, 1.0),
                     lambda: capi.op_langevin_gradient(capi.TASK_CLASSIFICATION, topo, bad, w, 0.01)):
            with pytest.raises(capi.PtfnnError) as e:
                call()
            assert e.value.code == capi.E_INVALID and "row 7" in str(e.value)
        fx, prob = capi.op_evaluate_proposal(capi.TASK_CLASSIFICATION, topo, bad, w)
        fx_ref, prob_ref = oc.evaluate(on.CLASSIFICATION, topo, good, w)
        assert np.array_equal(fx, fx_ref) and cm.relerr(prob, prob_ref) < RTOL
        with Sampler(capi.TASK_CLASSIFICATION, topo, [1.0, 2.0], 10, 5) as s:
            with pytest.raises(capi.PtfnnError) as e:
                s.set_data(good, bad)
            assert e.value.code == capi.E_INVALID and "test row 7" in str(e.value)
    frac = good.copy()
    frac[:, 4] += 0.5                                            # int(y) truncates (C:74, C:217)
    assert np.allclose(capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, frac, w, 1.0, 1.0)[0],
                       capi.op_likelihood(capi.TASK_CLASSIFICATION, topo, good, w, 1.0, 1.0)[0])


def test_swap_sweep_under_the_drafts_temperature_rule():
    """the temperature-aware swap rule of the reference's drafts (Misc/ldpt_fnn_multi_fixed.py:520) as
    an opt-in ``swap_kind`` of the same sequential sweep; the temperature field travels with its vector."""
    rs = np.random.RandomState(9)
    for n in (2, 5, 64, 500):
        temps = 1.0 / np.logspace(0, -np.log10(4.0), n)
        lh = -np.abs(rs.randn(n)) * 50 - 1.0                       # log-likelihood fields are negative
        lh[rs.randint(0, n)] = 0.0                                 # the rule's lhood2 == 0 guard
        u = rs.randint(0, 1 << 24, size=n - 1).astype(np.float64) / (1 << 24)
        a, b = on.swap_sweep_ratio_temperature(lh, u, temps)
        c, d = capi.op_swap_sweep(lh, u, swap_kind=capi.SWAP_KIND_RATIO_TEMPERATURE, temperatures=temps)
        assert a == c.tolist() and b == d.tolist(), n
        assert n < 5 or (any(b) and not all(b))
    with pytest.raises(capi.PtfnnError):
        capi.op_swap_sweep([1.0, 2.0], [0.5], swap_kind=capi.SWAP_KIND_RATIO_TEMPERATURE)      # needs the temperatures


def test_chain_runs_under_the_drafts_swap_rule_and_counts_every_proposal():
    from ptnn_b200.sampler import Sampler, geometric_ladder
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    R, S, si = 6, 61, 10
    out = {}
    for kind in (capi.SWAP_KIND_REFERENCE, capi.SWAP_KIND_RATIO_TEMPERATURE):
        with Sampler(on.REGRESSION, (4, 5, 1), geometric_ladder(R, 2), S, si, seed=5, swap_kind=kind) as s:
            s.set_data(tr, te)
            s.init_chains(np.random.RandomState(1).randn(R, 31))
            assert s.run() == S - 1
            ns, tot, sw = s.swap_stats()
            out[kind] = (ns, tot, sw, s.traces()["lik_prop"])
        assert tot == sw.shape[0] * (R - 1) and ns == int(sw.sum())
    # same chains up to the first round, different swap decisions afterwards
    assert np.array_equal(out[0][3][:, :si + 1], out[1][3][:, :si + 1])
    assert not np.array_equal(out[0][2], out[1][2])
    with pytest.raises(capi.PtfnnError) as e:                      # opt-in on a single GPU only
        Sampler(on.REGRESSION, (4, 5, 1), geometric_ladder(R, 2)[:3], S, si, n_replicas_global=R, swap_kind=1)
    assert e.value.code == capi.E_UNSUPPORTED


def test_eta_crosses_a_swap_round_as_float64():
    """R:430-437: the hand-shake moves the float64 eta.  A handle that stops BEFORE the round (its block is half of a
    larger ladder, so the host would complete the round) shows the pre-swap state; the same chains on one handle run
    through the round on the device: every eta afterwards must be one of the pre-swap values BIT FOR BIT, at the slot
    the sweep assigns."""
    from ptnn_b200.sampler import Sampler, geometric_ladder
    tr, te = cm.dataset(on.REGRESSION, "Lazer")
    R, si = 8, 6
    S = si + 3
    temps = geometric_ladder(2 * R, 3)
    w0 = np.random.RandomState(2).randn(R, 31)
    kw = dict(seed=11, learn_rate=0.1, common_random_numbers=False)
    with Sampler(on.REGRESSION, (4, 5, 1), temps[:R], S, si, n_replicas_global=2 * R, **kw) as a:
        a.set_data(tr, te)
        a.init_chains(w0)
        n = a.run()                                                # stops right after the swap step
        assert a.swap_pending()[0] and n == si + 1
        before = a.get_state()
    with Sampler(on.REGRESSION, (4, 5, 1), temps[:R], S, si, **kw) as b:
        b.set_data(tr, te)
        b.init_chains(w0)
        assert b.run(n) == n
        after = b.get_state()
        ns, tot, sw = b.swap_stats()
    assert ns > 0 and tot == R - 1
    src = list(range(R))
    for k in range(R - 1):                                         # the sweep's permutation from its own decisions
        if sw[0, k]:
            src[k], src[k + 1] = src[k + 1], src[k]
    assert src != list(range(R))
    assert np.array_equal(after["eta"], before["eta"][src])        # float64, bit for bit (no float32 round trip)
    assert np.array_equal(after["w"], before["w"][src])
    assert len(set(before["eta"].tolist())) == R                   # (distinct values: the check can tell them apart)


def test_a_wait_that_gives_up_fails_closed():
    """A device-side wait that times out (here: a peer flag that never arrives -- rank 1 is a process that exports its
    swap window and then never runs) must end the launch WITHOUT touching chain state or traces, report PTFNN_E_CUDA
    from every later call, and be cleared by ptfnn_init_chains."""
    import os
    import subprocess
    import sys
    from ptnn_b200.sampler import Sampler, geometric_ladder
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    R, S, si = 4, 30, 5
    temps = geometric_ladder(2 * R, 2)
    w0 = np.random.RandomState(0).randn(R, 31)
    idle = subprocess.Popen([sys.executable, os.path.join(root, "tests", "peer_idle_worker.py")], stdin=subprocess.PIPE,
                            stdout=subprocess.PIPE, text=True, cwd=root)
    try:
        line = idle.stdout.readline()
        assert line.startswith("HANDLES "), line
        theirs = bytes.fromhex(line.split()[1])
        with Sampler(on.REGRESSION, (4, 5, 1), temps[:R], S, si, n_replicas_global=2 * R, seed=3, barrier_timeout_ms=300) as s:
            s.set_data(tr, te)
            s.init_chains(w0)
            s.peer_connect([s.peer_export(), theirs], 0)
            s.run(si + 1)                            # reaches the first round and waits for rank 1, which never publishes
            with pytest.raises(capi.PtfnnError) as e:
                s.sync()
            assert e.value.code == capi.E_CUDA and "timed out" in str(e.value)
            for call in (lambda: s.get_state(), lambda: s.swap_stats(), lambda: s.traces(), lambda: s.run(1)):
                with pytest.raises(capi.PtfnnError) as e:
                    call()
                assert e.value.code == capi.E_CUDA
            s.init_chains(w0)                        # starts over: the failure is cleared, and the first steps run again
            assert s.run(si) == si                   # (stops short of the round)
            s.sync()
            t = s.traces(first=0, count=si + 1)
            assert np.all(np.isfinite(t["lik_prop"]))
    finally:
        try:
            idle.stdin.write("bye\n"); idle.stdin.flush()
        except OSError:
            pass
        idle.wait(timeout=60)


def test_set_data_between_launches_does_not_wait_and_is_stream_ordered():
    """ptfnn_set_data packs into page-locked staging and queues the upload behind the running launch (no wait for the
    device; two slots, reused two calls later).  Data sets swapped between launches without any synchronisation in
    between -- three times, so that a staging slot is reused -- must give exactly the chain that waits after every call,
    also for the tcgen05 topology, whose A tiles are repacked on the device by the same call."""
    from ptnn_b200.sampler import Sampler, geometric_ladder
    from ptnn_b200 import datasets
    for task, topo, lr in ((on.REGRESSION, (4, 5, 1), 0.1), (on.CLASSIFICATION, (16, 256, 10), 0.01)):
        if task == on.REGRESSION:
            tr, te = cm.dataset(on.REGRESSION, "Sunspot")
            sets = [(tr, te), (tr[::-1].copy(), te), (tr[:200], te[:100]), (tr, te[::-1].copy())]
        else:
            tr, te = datasets.synthetic_pendigit(n_train=700, n_test=300)
            sets = [(tr, te), (tr[::-1].copy(), te), (tr[:300], te[:130]), (tr, te[::-1].copy())]
        R, S, n = 4, 4 * 12 + 2, 12
        P = topo[0] * topo[1] + topo[1] * topo[2] + topo[1] + topo[2]
        w0 = np.random.RandomState(2).randn(R, P) * 0.3
        out = []
        for wait in (True, False):
            with Sampler(task, topo, geometric_ladder(R, 2), S, 5, learn_rate=lr, seed=11, debug_traces=True) as s:
                s.set_data(*sets[0])
                s.init_chains(w0)
                for a, b in sets:
                    s.set_data(a, b)
                    if wait:
                        s.sync()
                    s.run(n)
                    if wait:
                        s.sync()
                out.append((s.traces(), s.get_state()))
        for k in out[0][0]:
            assert np.array_equal(out[0][0][k], out[1][0][k]), (topo, k)
        for k in ("w", "eta", "lik", "prior", "tau", "num_accepted"):
            assert np.array_equal(out[0][1][k], out[1][1][k]), (topo, k)
