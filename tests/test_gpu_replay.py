"""GPU parity of the whole hot path in replay mode: the reference's recorded draws are fed to the
persistent chain kernel and its traces are compared with what the UNMODIFIED reference produced
(tests/golden/*.npz) and with the float64 oracle.

  teacher-forced: the device state is reset to the oracle's pre-step state before every step, so
                  every step's proposed log-likelihood / prior / MH probability is checked
                  independently (1e-4 relative);
  free replay:    only the initial state is shared; accept / swap decisions must be identical
                  up to a documented near-tie (u between the two MH probabilities).
"""
import numpy as np
import pytest

from oracle import ptfnn_c as oc
from oracle import ptfnn_numpy as on
from ptnn_b200.sampler import Sampler
from tests import common as cm

pytestmark = pytest.mark.gpu
RTOL = 1e-4          # per step (teacher-forced): the north-star criterion
# Free replay shares only the initial state, so fp32 rounding compounds from step to step.  The bar stays the
# north-star's 1e-4; what is compared against it is chosen by the CONDITIONING of each quantity, not by what was
# observed: weights and RMSE relative to themselves; the Gaussian log-likelihood (R:204-205)
#     lik = [-N/2 log(2 pi tau^2) - SSE / (2 tau^2)] / T
# relative to the magnitude of its two terms -- where tau^2 ~ MSE they cancel (|lik| << N/2 |log 2 pi tau^2|), and an
# SSE that is right to 1e-5 then moves lik by 1e-4 of ITSELF (reg_henon_lg1: 44 consecutive Langevin steps, weights
# within 5.5e-6, |lik| ~ 25 against terms of ~400).  `lik_scale` is that magnitude, from the oracle's own tau^2.
RTOL_FREE = 1e-4


def _sampler(cfg, temps, **kw):
    kw.setdefault("debug_traces", True)
    return Sampler.from_oracle_config(cfg, temps, **kw)


@pytest.mark.parametrize("name", cm.CASES)
def test_init_chains_matches_reference(name):
    fx, cfg, tr, te, draws = cm.case(name)
    with _sampler(cfg, fx["temperatures"]) as s:
        s.set_data(tr, te)
        s.init_chains(fx["w0"])
        st = s.get_state()
    assert cm.relerr(st["lik"], fx["ref_init_lik"]) < RTOL          # R:284 / C:283
    assert cm.relerr(st["prior"], fx["ref_init_prior"]) < RTOL      # R:280 / C:281
    assert np.array_equal(st["w"], fx["w0"].astype(np.float32).astype(np.float64))


@pytest.mark.parametrize("name", cm.CASES)
def test_teacher_forced_replay(name):
    fx, cfg, tr, te, draws = cm.case(name)
    ref = oc.run_pt(cfg, tr, te, fx["temperatures"], fx["w0"], draws)
    S = cfg.samples
    with _sampler(cfg, fx["temperatures"]) as s:
        s.set_data(tr, te)
        s.init_chains(fx["w0"])
        for i in range(S - 1):
            s.set_state(w=ref.state_w[:, i], eta=ref.state_eta[:, i], lik=ref.state_lik[:, i],
                        prior=ref.state_prior[:, i], tau=ref.state_tau[:, i])
            assert s.replay(draws, n_steps=1) == 1
        t = s.traces()
    # the reference's own numbers (golden) and the oracle's
    assert cm.relerr(t["lik_prop"][:, 1:], ref.lik_prop[:, 1:]) < RTOL
    scale = 1.0 if cfg.task == on.REGRESSION else None
    if scale:
        assert cm.relerr(t["lik_prop"][:, 1:], fx["ref_lik_prop"][:, 1:]) < RTOL
    assert cm.relerr(t["prior_prop"][:, 1:], fx["ref_prior_prop"][:, 1:]) < RTOL
    # diff_prop is a difference of two O(P/2..1000) terms: compare on the scale of the terms
    assert np.max(np.abs(t["diff_prop"] - ref.diff_prop)) < RTOL * max(1.0, np.max(np.abs(ref.diff_prop)), cfg.P)
    # decisions: identical except where u falls between the two MH probabilities (near-tie)
    diff = t["accepted"] != ref.accepted
    lo = np.minimum(t["mh_prob"], ref.mh_prob) - 1e-7
    hi = np.maximum(t["mh_prob"], ref.mh_prob) + 1e-7
    u = np.concatenate([np.zeros((ref.accepted.shape[0], 1)), draws.u], axis=1)
    assert np.all((u[diff] >= lo[diff]) & (u[diff] <= hi[diff]))
    assert diff.mean() < 0.02
    # rows written on acceptance carry the proposal (pos_w) and its rmse / accuracy
    acc = t["accepted"] & ref.accepted
    w_prop_ref = np.where(ref.accepted[..., None], ref.pos_w, 0.0)
    assert cm.relerr(np.where(acc[..., None], t["pos_w"], 0.0), np.where(acc[..., None], w_prop_ref, 0.0)) < RTOL
    assert cm.relerr(t["rmse_train"] * acc, fx["ref_rmse_train"] * acc) < (RTOL if cfg.task == on.REGRESSION else 0.2)
    assert cm.relerr(t["rmse_test"] * acc, fx["ref_rmse_test"] * acc) < (RTOL if cfg.task == on.REGRESSION else 0.2)


def _first_divergence(t, ref, sw_dev, cfg):
    """First step index at which any accept decision or any swap decision differs (S-1 if none)."""
    S = cfg.samples
    bad = np.argwhere(t["accepted"] != ref.accepted)
    i_acc = int(bad[:, 1].min()) - 1 if bad.size else S - 1          # row i+1 holds step i
    n = min(len(sw_dev), len(ref.swapped))
    i_sw = S - 1
    rnd = 0
    for i in range(S - 1):
        if cfg.swap_due(i):
            if rnd < n and not np.array_equal(sw_dev[rnd], ref.swapped[rnd]):
                i_sw = i
                break
            rnd += 1
    return min(i_acc, i_sw), i_acc, i_sw


@pytest.mark.parametrize("name", cm.CASES)
def test_free_replay_decisions_and_traces(name):
    fx, cfg, tr, te, draws = cm.case(name)
    ref = oc.run_pt(cfg, tr, te, fx["temperatures"], fx["w0"], draws)
    S = cfg.samples
    with _sampler(cfg, fx["temperatures"]) as s:
        s.set_data(tr, te)
        s.init_chains(fx["w0"])
        assert s.replay(draws) == S - 1
        t = s.traces()
        ns, tot, sw = s.swap_stats()
    assert tot == int(fx["ref_total_swap_proposals"])                 # rounds incl. the left-over one (Q9)
    i_star, i_acc, i_sw = _first_divergence(t, ref, sw, cfg)
    if i_star < S - 1:
        # documented near-tie: at the first divergent step u lies between the two MH probabilities
        if i_acc <= i_sw:
            r = int(np.argwhere(t["accepted"][:, i_acc + 1] != ref.accepted[:, i_acc + 1])[0, 0])
            u = draws.u[r, i_acc]
            lo, hi = sorted([t["mh_prob"][r, i_acc + 1], ref.mh_prob[r, i_acc + 1]])
            assert lo - 1e-7 <= u <= hi + 1e-7 and (hi - lo) <= 1e-3 * max(hi, 1e-30) + 1e-7, (name, i_acc, r, u, lo, hi)
    rows = slice(0, i_star + 1)     # rows 0..i_star were written by steps < i_star
    tol = RTOL_FREE
    scale = np.maximum(1.0, cm.lik_scale(cfg, ref, tr.shape[0], fx["temperatures"]))
    err = np.abs(t["lik_prop"] - ref.lik_prop) / scale
    assert float(np.max(err[:, rows][:, 1:])) < tol
    assert cm.relerr(t["pos_w"][:, rows], fx["ref_pos_w"][:, rows]) < tol          # the reference's own file
    assert np.array_equal(t["accept_list"][:, rows], fx["ref_accept_list"][:, rows])
    if cfg.task == on.REGRESSION:
        assert cm.relerr(t["rmse_train"][:, rows], ref.rmse_train[:, rows]) < tol
        assert cm.relerr(t["rmse_test"][:, rows], ref.rmse_test[:, rows]) < tol
    if i_star == S - 1:
        assert ns == int(fx["ref_num_swap"]) and np.array_equal(sw, fx["ref_swapped"])
        assert np.array_equal(t["accepted"], ref.accepted)
    # most of every golden case must replay before any near-tie
    assert i_star >= (S - 1) // 2, (name, i_star)


def test_trace_row_zero_and_carried_rows():
    """pos_w row 0 = ones and stays ones until the first acceptance; likelihood row 0
    = -100; rmse rows 0 until the first acceptance; accept_list[i+1] = count BEFORE step i."""
    fx, cfg, tr, te, draws = cm.case("reg_lazer_rw")
    with _sampler(cfg, fx["temperatures"]) as s:
        s.set_data(tr, te)
        s.init_chains(fx["w0"])
        s.replay(draws)
        t = s.traces()
    assert np.all(t["pos_w"][:, 0] == 1.0) and np.all(t["lik_prop"][:, 0] == -100.0)
    for r in range(t["accepted"].shape[0]):
        first = int(np.argmax(t["accepted"][r])) if t["accepted"][r].any() else cfg.samples
        assert np.all(t["pos_w"][r, :first] == 1.0) and np.all(t["rmse_train"][r, :first] == 0.0)
        assert np.array_equal(t["accept_list"][r, 1:], np.cumsum(t["accepted"][r])[:-1])
    assert np.all(t["acc_train"] == 0.0)                                              # R:403-404


@pytest.mark.parametrize("memo", [0, 1])
def test_wide_hidden_chain_replay(memo):
    """[16,256,10] (BASELINE configs[4] shape): the team SGD kernel, langevin_gradient buffers in global
    memory and the wide-hidden likelihood path inside the chain kernel, against the float64 oracle on the
    same draws (Langevin and random-walk steps, swaps, the 60% temperature switch at step 6)."""
    rs = np.random.RandomState(21)
    topo = (16, 256, 10)
    cfg = on.PTConfig(task=on.CLASSIFICATION, topology=topo, samples=10, swap_interval=3,
                      use_langevin_gradients=True, l_prob=0.5, learn_rate=0.01)
    mk = lambda n: np.hstack([rs.randn(n, 16), rs.randint(0, 10, size=(n, 1)).astype(float)])   # noqa: E731
    tr, te = mk(300), mk(131)
    R = 3
    temps = np.array([1.0, 1.5, 2.25])
    draws = on.random_draws(cfg, R, 5, common_random_numbers=True)
    draws.lx[:, ::2] = 0.1                 # Langevin steps
    draws.lx[:, 1::2] = 0.9                # random-walk steps
    w0 = rs.randn(R, cfg.P) * 0.2
    ref = oc.run_pt(cfg, tr, te, temps, w0, draws)
    with _sampler(cfg, temps, memoize_gradient=memo) as s:
        s.set_data(tr, te)
        s.init_chains(w0)
        assert s.replay(draws) == cfg.samples - 1
        t = s.traces()
    diff = np.argwhere(t["accepted"] != ref.accepted)
    i_star = int(diff[:, 1].min()) - 1 if diff.size else cfg.samples - 1
    if diff.size:                          # only a documented near-tie may differ
        r, c = diff[np.argmin(diff[:, 1])]
        lo, hi = sorted([t["mh_prob"][r, c], ref.mh_prob[r, c]])
        assert lo - 1e-7 <= draws.u[r, c - 1] <= hi + 1e-7
    rows = slice(1, i_star + 2)
    assert cm.relerr(t["lik_prop"][:, rows], ref.lik_prop[:, rows]) < RTOL
    assert cm.relerr(t["prior_prop"][:, rows], ref.prior_prop[:, rows]) < RTOL
    assert np.max(np.abs(t["diff_prop"][:, rows] - ref.diff_prop[:, rows])) < RTOL * cfg.P
    assert cm.relerr(t["pos_w"][:, :i_star + 1], ref.pos_w[:, :i_star + 1]) < RTOL
    assert i_star >= 4


@pytest.mark.parametrize("memo", [0, 1])
def test_speculative_windows_are_bit_identical(memo):
    """Small ladders: K CTAs per temperature evaluate the steps ahead at once, each step assuming the earlier
    ones rejected.  Every depth must give the sequential chain bit for bit: traces, swap
    decisions, counters and final state -- through swap rounds, the left-over round and the 60% temperature
    switch (step 120 of 200)."""
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    R, S, si = 10, 200, 20
    from ptnn_b200.sampler import geometric_ladder
    temps = geometric_ladder(R, 2)
    w0 = np.random.RandomState(3).randn(R, 31)
    out = {}
    for depth in (1, 3, 8, 14):
        with Sampler(on.REGRESSION, (4, 5, 1), temps, S, si, use_langevin_gradients=True, l_prob=0.5, learn_rate=0.1,
                     seed=123, common_random_numbers=False, memoize_gradient=memo, debug_traces=True,
                     speculation=depth) as s:
            s.set_data(tr, te)
            s.init_chains(w0)
            assert s.run() == S - 1
            t = s.traces()
            st = s.get_state()
            out[depth] = (t, s.swap_stats(), st)
    t1, sw1, st1 = out[1]
    assert t1["accepted"].sum() > R and sw1[0] > 0            # the run really accepts and swaps
    for depth in (3, 8, 14):
        t, sw, st = out[depth]
        for k in t1:
            assert np.array_equal(t[k], t1[k]), (depth, k)
        assert sw[0] == sw1[0] and sw[1] == sw1[1] and np.array_equal(sw[2], sw1[2]), depth
        for k in ("w", "eta", "lik", "prior", "tau", "num_accepted"):
            assert np.array_equal(st[k], st1[k]), (depth, k)


@pytest.mark.parametrize("task,l_prob,use_lg", [(0, 0.08, True), (0, 1.0, True), (0, 0.5, False), (1, 0.3, True)])
def test_packed_window_plans_are_bit_identical(task, l_prob, use_lg):
    """The window plan (a step belongs to the CTA indexed by the number of Langevin steps before it;
    runs of more than eight random-walk steps move on; at most 32 steps per window) at its corners: long random-walk
    runs (l_prob 0.08), Langevin steps only, random-walk chains (one step per CTA) and a classification chain whose
    swap segments (40 steps) are longer than a window.  Depths 2, 5 and 16 against the sequential chain, bit for bit."""
    if task == 0:
        tr, te = cm.dataset(on.REGRESSION, "Sunspot")
        topo, kind, lr = (4, 5, 1), on.REGRESSION, 0.1
    else:
        tr, te = cm.dataset(on.CLASSIFICATION, "Iris")
        topo, kind, lr = (4, 12, 3), on.CLASSIFICATION, 0.01
    P = topo[0] * topo[1] + topo[1] * topo[2] + topo[1] + topo[2]
    R, S, si = 6, 161, 40
    from ptnn_b200.sampler import geometric_ladder
    temps = geometric_ladder(R, 2)
    w0 = np.random.RandomState(11).randn(R, P) * 0.5
    out = {}
    for depth in (1, 2, 5, 16):
        with Sampler(kind, topo, temps, S, si, use_langevin_gradients=use_lg, l_prob=l_prob, learn_rate=lr, seed=77,
                     common_random_numbers=False, memoize_gradient=depth % 2, debug_traces=True, speculation=depth) as s:
            s.set_data(tr, te)
            s.init_chains(w0)
            assert s.run() == S - 1
            out[depth] = (s.traces(), s.swap_stats(), s.get_state())
    t1, sw1, st1 = out[1]
    assert t1["accepted"].sum() > 0 and sw1[0] > 0            # (an all-Langevin Sunspot chain accepts rarely)
    for depth in (2, 5, 16):
        t, sw, st = out[depth]
        for k in t1:
            assert np.array_equal(t[k], t1[k]), (depth, k)
        assert sw[0] == sw1[0] and sw[1] == sw1[1] and np.array_equal(sw[2], sw1[2]), depth
        for k in ("w", "eta", "lik", "prior", "tau", "num_accepted"):
            assert np.array_equal(st[k], st1[k]), (depth, k)


def test_automatic_window_depth_follows_acceptance_and_changes_nothing():
    """Ladders that leave CTA slots free (here 96 temperatures of the 4-64-1 net: ~10 slots each) get speculative
    windows whose depth follows the acceptance rate observed so far on the run -- launches early in
    the run, where everything is accepted, stay sequential; later ones go deep.  Whatever depth each launch picks, the
    chain is the sequential one bit for bit."""
    from ptnn_b200 import datasets
    tr, te = datasets.synthetic_timeseries()
    tr, te = tr[:3000], te[:1000]
    R, S, si = 96, 121, 10
    from ptnn_b200.sampler import geometric_ladder
    temps = geometric_ladder(R, 2)
    w0 = np.random.RandomState(5).randn(R, 385) * 0.5
    out = {}
    for depth in (1, 0, 6, "0, one call"):  # sequential | automatic | fixed | automatic, the whole chain asked for at once
        with Sampler(on.REGRESSION, (4, 64, 1), temps, S, si, use_langevin_gradients=True, l_prob=0.5, learn_rate=0.01,
                     seed=9, common_random_numbers=True, memoize_gradient=1, debug_traces=True,
                     speculation=depth if isinstance(depth, int) else 0) as s:
            s.set_data(tr, te)
            s.init_chains(w0)
            if isinstance(depth, str):
                assert s.run() == S - 1     # (the library launches it in pieces that end on swap rounds, and steers between them)
            while s.step < S - 1:
                s.run(si)
                s.sync()                    # (gives the asynchronous acceptance sample time to arrive: the depth may change)
            out[depth] = (s.traces(), s.swap_stats(), s.get_state())
    t1, sw1, st1 = out[1]
    acc_rate = t1["accepted"][:, 60:].mean()
    assert acc_rate < 0.6                   # the regime in which the automatic policy opens windows
    for depth in (0, 6, "0, one call"):
        t, sw, st = out[depth]
        for k in t1:
            assert np.array_equal(t[k], t1[k]), (depth, k)
        assert sw[0] == sw1[0] and np.array_equal(sw[2], sw1[2]), depth
        for k in ("w", "eta", "lik", "prior", "tau", "num_accepted"):
            assert np.array_equal(st[k], st1[k]), (depth, k)
