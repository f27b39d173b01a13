"""Pins the oracles (oracle/ptfnn_numpy.py, oracle/ptfnn_oracle.c) to outputs of the UNMODIFIED
reference (tests/golden/*.npz, produced by oracle/gen_golden.py in the build container).
CPU only."""
import numpy as np
import pytest

from oracle import ptfnn_c as oc
from oracle import ptfnn_numpy as on
from tests import common as cm


def _check_traces(t, fx, cfg, tol):
    assert cm.relerr(t.extra["init_lik"], fx["ref_init_lik"]) < tol if "init_lik" in t.extra else True
    assert cm.relerr(t.lik_prop_t[:, 1:], fx["ref_lik_prop"][:, 1:]) < tol
    assert cm.relerr(t.prior_prop[:, 1:], fx["ref_prior_prop"][:, 1:]) < tol
    assert cm.relerr(t.pos_w, fx["ref_pos_w"]) < tol
    assert np.array_equal(t.accept_list, fx["ref_accept_list"])
    acc = t.accepted
    assert cm.relerr(t.rmse_train * acc, fx["ref_rmse_train"] * acc) < tol
    assert cm.relerr(t.rmse_test * acc, fx["ref_rmse_test"] * acc) < tol
    assert np.array_equal(t.swapped, fx["ref_swapped"])
    assert t.num_swap == int(fx["ref_num_swap"])
    assert t.total_swap_proposals == int(fx["ref_total_swap_proposals"])
    # what the reference writes to posterior/pos_likelihood (%1.4f) and predictions/ (%1.8f | %1.2f)
    assert np.max(np.abs(t.lik_prop - fx["ref_pos_likelihood_file"][:, :, 0])) < 5.01e-5 * max(1.0, np.max(np.abs(t.lik_prop)) * 1e-3)
    dec = 8 if cfg.task == on.REGRESSION else 2
    assert np.max(np.abs(t.rmse_train - fx["ref_rmse_train_file"])) <= 0.5001 * 10.0 ** -dec
    assert np.max(np.abs(t.acc_train - fx["ref_acc_train_file"])) <= 0.5001e-2
    assert np.max(np.abs(t.acc_test - fx["ref_acc_test_file"])) <= 0.5001e-2
    swap_perc = t.num_swap * 100 / t.total_swap_proposals
    assert abs(swap_perc - float(fx["ref_swap_perc"])) < 1e-12
    assert np.array_equal(t.accept_list, fx["ref_accept_vec"])


@pytest.mark.parametrize("name", ["reg_lazer_rw", "reg_mackey_h10", "cls_cancer_lg"])
def test_numpy_oracle_replays_reference(name):
    fx, cfg, tr, te, draws = cm.case(name)
    t = on.run_pt(cfg, tr, te, fx["temperatures"], fx["w0"], draws)
    _check_traces(t, fx, cfg, 1e-12)


@pytest.mark.parametrize("name", cm.CASES)
def test_c_oracle_replays_reference(name):
    fx, cfg, tr, te, draws = cm.case(name)
    t = oc.run_pt(cfg, tr, te, fx["temperatures"], fx["w0"], draws)
    _check_traces(t, fx, cfg, 1e-9)


@pytest.mark.parametrize("name", cm.CASES)
def test_round_count_matches_reference(name):
    fx, cfg, *_ = cm.case(name)
    R = int(fx["R"])
    assert cfg.total_rounds() * (R - 1) == int(fx["ref_total_swap_proposals"])
    assert oc.total_rounds(cfg) == cfg.total_rounds()


@pytest.mark.parametrize("ds", cm.REG_DATASETS)
@pytest.mark.parametrize("H", [5, 10])
def test_regression_known_answers(ds, H):
    ka = cm.npz("known_answers")
    tr, te = cm.dataset(on.REGRESSION, ds)
    key = "reg_%s_h%d_" % (ds, H)
    topo, w, tau = (4, H, 1), ka[key + "w"], float(ka[key + "tau"])
    net = on.Network(topo, 0.1, on.REGRESSION)
    fx = net.evaluate_proposal(tr, w)
    assert np.array_equal(fx, ka[key + "fx"])
    assert float(np.var(fx - tr[:, 4])) == tau
    lik, _, rm = on.likelihood_regression(net, tr, w, tau, 1.25)
    lik_te, _, rm_te = on.likelihood_regression(net, te, w, tau, 1.25)
    assert np.allclose([lik, rm, lik_te, rm_te], ka[key + "lik"], rtol=1e-14, atol=0)
    assert on.prior_regression(25, 0, 0, w, tau, topo) == pytest.approx(float(ka[key + "prior"]), rel=1e-14)
    assert cm.relerr(net.langevin_gradient(tr, w, 1), ka[key + "w_gd"]) < 1e-13
    # C oracle
    assert cm.relerr(oc.evaluate(on.REGRESSION, topo, tr, w), ka[key + "fx"]) < 1e-13
    l, r, _ = oc.likelihood(on.REGRESSION, topo, tr, w, tau, 1.25)
    assert np.allclose([l, r], ka[key + "lik"][:2], rtol=1e-12, atol=0)
    assert oc.prior(on.REGRESSION, topo, w, tausq=tau) == pytest.approx(float(ka[key + "prior"]), rel=1e-13)
    assert cm.relerr(oc.langevin_gradient(on.REGRESSION, topo, tr, w, 0.1), ka[key + "w_gd"]) < 1e-11


@pytest.mark.parametrize("ds", cm.CLS_DATASETS)
def test_classification_known_answers(ds):
    ka = cm.npz("known_answers")
    tr, te = cm.dataset(on.CLASSIFICATION, ds)
    topo = cm.cls_topology(ds)
    key = "cls_%s_" % ds
    w = ka[key + "w"]
    net = on.Network(topo, 0.01, on.CLASSIFICATION)
    fx, prob = net.evaluate_proposal(tr, w)
    assert np.array_equal(fx, ka[key + "fx"]) and np.array_equal(prob, ka[key + "prob"])
    lik, _, rm = on.likelihood_classification(net, tr, w, 2.5)
    lik_te, fx_te, rm_te = on.likelihood_classification(net, te, w, 2.5)
    assert np.allclose([lik, rm, lik_te, rm_te], ka[key + "lik"], rtol=1e-13, atol=0)
    assert on.accuracy(fx, tr[:, topo[0]]) == float(ka[key + "acc"][0])
    assert on.accuracy(fx_te, te[:, topo[0]]) == float(ka[key + "acc"][1])
    assert on.prior_classification(25, 0, 0, w, topo) == pytest.approx(float(ka[key + "prior"]), rel=1e-14)
    assert cm.relerr(net.langevin_gradient(tr, w, 1), ka[key + "w_gd"]) < 1e-13
    # C oracle
    fxc, probc = oc.evaluate(on.CLASSIFICATION, topo, tr, w)
    assert np.array_equal(fxc, ka[key + "fx"]) and cm.relerr(probc, ka[key + "prob"]) < 1e-13
    l, r, a = oc.likelihood(on.CLASSIFICATION, topo, tr, w, 1.0, 2.5)
    assert np.allclose([l, r, a], [ka[key + "lik"][0], ka[key + "lik"][1], ka[key + "acc"][0]], rtol=1e-12, atol=0)
    assert oc.prior(on.CLASSIFICATION, topo, w) == pytest.approx(float(ka[key + "prior"]), rel=1e-13)
    assert cm.relerr(oc.langevin_gradient(on.CLASSIFICATION, topo, tr, w, 0.01), ka[key + "w_gd"]) < 1e-11


def test_ladder_and_swap_rule():
    ka = cm.npz("known_answers")
    assert np.allclose(on.geometric_ladder(10, 2), ka["ladder_10_2"], rtol=1e-15, atol=0)
    assert np.allclose(on.geometric_ladder(7, 10), ka["ladder_7_10"], rtol=1e-15, atol=0)
    # swap rule R:674 incl. the 709 clamp and the 0.5 prefactor; sequential bubble (R:741-748)
    assert on.swap_probability(0.0, 0.0) == 0.5
    assert on.swap_probability(0.0, 800.0) == 1
    assert on.swap_probability(5.0, 5.0 + np.log(2.0) - 1e-9) < 1
    lh = [0.0, 10.0, 20.0, 30.0]                      # every pair swaps: state 0 bubbles to the top
    src, sw = on.swap_sweep(lh, [0.99, 0.99, 0.99])
    assert src == [1, 2, 3, 0] and all(sw)
    srcc, swc = oc.swap_sweep(lh, [0.99, 0.99, 0.99])
    assert srcc.tolist() == src and swc.all()
    rs = np.random.RandomState(3)
    for _ in range(50):
        lh, u = rs.randn(9) * 2, rs.rand(8)
        a, b = on.swap_sweep(lh, u)
        c, d = oc.swap_sweep(lh, u)
        assert a == c.tolist() and b == d.tolist()


@pytest.mark.parametrize("seed", range(6))
def test_c_oracle_equals_numpy_oracle_on_random_configurations(seed):
    """The C restatement (used for full-size replays) against the NumPy one (bit-identical to the reference on the
    golden runs) outside the golden set: random topologies, ladders, swap cadences, Langevin probabilities and data,
    including a round on the last step, 0.6*S integral or not, and both tasks."""
    rs = np.random.RandomState(100 + seed)
    task = on.REGRESSION if seed % 2 == 0 else on.CLASSIFICATION
    I, H = int(rs.randint(1, 7)), int(rs.randint(1, 12))
    O = 1 if task == on.REGRESSION else int(rs.randint(2, 5))
    R, S, si = int(rs.randint(1, 5)), int(rs.choice([10, 15, 17, 20])), int(rs.randint(1, 7))
    cfg = on.PTConfig(task=task, topology=(I, H, O), samples=S, swap_interval=si, use_langevin_gradients=bool(seed % 3),
                      l_prob=float(rs.choice([0.3, 0.5, 1.0])), learn_rate=float(rs.choice([0.01, 0.1])))

    def data(n):
        x = rs.rand(n, I) if task == on.REGRESSION else rs.randn(n, I)
        y = rs.rand(n, 1) if task == on.REGRESSION else rs.randint(0, O, size=(n, 1)).astype(float)
        return np.hstack([x, y])

    tr, te = data(int(rs.randint(3, 40))), data(int(rs.randint(2, 25)))
    temps = on.geometric_ladder(R, float(rs.choice([2, 5, 10]))) if R > 1 else np.ones(1)
    w0 = rs.randn(R, cfg.P) * 0.7
    draws = on.random_draws(cfg, R, seed, common_random_numbers=bool(seed % 2))
    a = on.run_pt(cfg, tr, te, temps, w0, draws)
    b = oc.run_pt(cfg, tr, te, temps, w0, draws)
    assert np.array_equal(a.accepted, b.accepted) and np.array_equal(a.swapped, b.swapped)
    assert (a.num_swap, a.total_swap_proposals) == (b.num_swap, b.total_swap_proposals)
    for k in ("pos_w", "lik_prop", "prior_prop", "diff_prop", "mh_prob", "rmse_train", "rmse_test", "acc_train", "acc_test", "accept_list"):
        assert cm.relerr(getattr(a, k), getattr(b, k)) < 1e-9, k
