"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden

For each case it
  1. runs the reference's own ParallelTempering / ptReplica / Network classes in-process through
     oracle/ref_harness.py, recording every random draw and the full-precision return value of
     every likelihood_func / prior_likelihood / langevin_gradient call,
  2. converts the recorded draws into the (lx, z, z_eta, u, u_swap) replay arrays,
  3. stores draws + reference outputs as a compressed fixture.
The fixtures are what pins oracle/ptfnn_numpy.py and oracle/ptfnn_oracle.c to the reference
(tests/test_oracle_golden.py) and what the CUDA path is replayed against (tests/test_gpu_*.py).
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_harness as rh          # noqa: E402
from oracle import ptfnn_numpy as on          # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (task, dataset, topology, R, maxtemp, S, swap_interval, use_lg, lr, l_prob, seed)
    "reg_sunspot_lg":  ("regression", "Sunspot", [4, 5, 1], 4, 2, 100, 10, True, 0.1, 0.5, 7),
    "reg_lazer_rw":    ("regression", "Lazer", [4, 5, 1], 3, 4, 57, 7, False, 0.1, 0.5, 11),
    "reg_mackey_h10":  ("regression", "Mackey", [4, 10, 1], 3, 5, 50, 10, True, 0.01, 0.8, 13),
    "cls_iris_lg":     ("classification", "Iris", [4, 12, 3], 4, 10, 80, 8, True, 0.01, 0.5, 17),
    "cls_cancer_lg":   ("classification", "Cancer", [9, 12, 2], 3, 10, 40, 5, True, 0.01, 0.5, 19),
    "cls_ions_lg":     ("classification", "Ionosphere", [34, 50, 2], 2, 10, 25, 6, True, 0.01, 0.5, 23),
    # the rest of the regression suite (BASELINE configs[1]) and random-walk classification
    "reg_lorenz_lg":   ("regression", "Lorenz", [4, 5, 1], 5, 2, 60, 6, True, 0.1, 0.5, 29),       # 0.6*S = 36: the switch fires
    "reg_henon_lg1":   ("regression", "Henon", [4, 5, 1], 3, 4, 45, 9, True, 0.01, 1.0, 31),       # every step Langevin
    "reg_acfin_rw":    ("regression", "ACFinance", [4, 5, 1], 6, 2, 70, 10, False, 0.1, 0.5, 37),
    "reg_rossler_lg":  ("regression", "Rossler", [4, 5, 1], 2, 5, 33, 4, True, 0.1, 0.5, 41),      # a round every 4 steps + the left-over one
    "cls_iris_rw":     ("classification", "Iris", [4, 12, 3], 5, 10, 60, 5, False, 0.01, 0.5, 43),
    "cls_pendigit_lg": ("classification", "PenDigit", [16, 30, 10], 3, 10, 20, 4, True, 0.01, 0.5, 47),   # ten classes, the reference's own data
}


def dataset(task, name):
    if task == "regression":
        tr, te = rh.load_regression_dataset(name)
        return tr, te
    tr, te, _ = rh.load_classification_dataset(name, split_seed=0)
    return tr, te


def logs_to_draws(task, logs, R, S, P, rounds):
    """Recorder logs -> replay arrays.  Per replica the call order is fixed by the reference
    (SURVEY Q5): randn(P) [R:256], 4x randn [R:35-38], then per step uniform [R:327], normal(P)
    [R:331|353], (regression) normal(1) [R:355], python uniform [R:387]."""
    per_step = 4 if task == "regression" else 3
    lx, z = np.zeros((R, S - 1)), np.zeros((R, S - 1, P))
    z_eta, u = np.zeros((R, S - 1)), np.zeros((R, S - 1))
    for k in range(R):
        log = logs[k + 1]
        assert [e[0] for e in log[:5]] == ["randn"] * 5
        body = log[5:]
        assert len(body) == per_step * (S - 1), (len(body), per_step, S)
        for i in range(S - 1):
            e = body[per_step * i: per_step * (i + 1)]
            assert e[0][0] == "uniform" and e[1][0] == "normal" and e[-1][0] == "pyuniform"
            lx[k, i] = e[0][1][0]
            z[k, i] = e[1][1]
            if task == "regression":
                assert e[2][0] == "normal" and e[2][1].size == 1
                z_eta[k, i] = e[2][1][0]
            u[k, i] = e[-1][1]
    main = logs[0]
    assert [e[0] for e in main[:R]] == ["randn"] * R
    us = [float(e[1][0]) for e in main[R:]]
    assert all(e[0] == "uniform" for e in main[R:])
    assert len(us) == rounds * (R - 1), (len(us), rounds, R)
    return on.Draws(lx=lx, z=z, z_eta=z_eta, u=u, u_swap=np.array(us).reshape(rounds, R - 1))


def calls_to_arrays(task, cfg, calls, R, S):
    """Per-call values -> per-step arrays.  likelihood_func call order per replica: init train,
    init test [R:284-285], then per step (switch: train, test [R:322-323])?, train, test [R:360-362]."""
    lik_prop = np.zeros((R, S))
    rmse_tr, rmse_te = np.zeros((R, S)), np.zeros((R, S))
    prior_prop = np.zeros((R, S))
    init = {"lik": np.zeros(R), "rmse_train": np.zeros(R), "rmse_test": np.zeros(R), "prior": np.zeros(R)}
    pt_samples = S * cfg.pt_fraction
    for k in range(R):
        lik, pri = calls[k]["lik"], calls[k]["prior"]
        init["lik"][k], init["rmse_train"][k], init["rmse_test"][k] = lik[0][0], lik[0][1], lik[1][1]
        init["prior"][k] = pri[0]
        j = 2
        for i in range(S - 1):
            if i == pt_samples:
                j += 2
            lik_prop[k, i + 1] = lik[j][0]
            rmse_tr[k, i + 1], rmse_te[k, i + 1] = lik[j][1], lik[j + 1][1]
            prior_prop[k, i + 1] = pri[i + 1]
            j += 2
        assert j == len(lik), (j, len(lik))
        assert len(pri) == S
    return lik_prop, rmse_tr, rmse_te, prior_prop, init


def read_chain_file(path, sub, T, suffix=""):
    return np.loadtxt(os.path.join(path, sub, "chain_" + str(T) + suffix + ".txt"))


def generate_case(name, spec):
    task, ds, topo, R, maxtemp, S, swap, use_lg, lr, l_prob, seed = spec
    tr, te = dataset(task, ds)
    if task == "classification":
        l_prob = 0.5                                     # C:192 hard-codes it
    cfg = on.PTConfig(task=on.REGRESSION if task == "regression" else on.CLASSIFICATION,
                      topology=tuple(topo), samples=S, swap_interval=swap,
                      use_langevin_gradients=use_lg, l_prob=l_prob, learn_rate=lr)
    P = cfg.P
    with tempfile.TemporaryDirectory() as d:
        out = rh.run_reference_pt(task, tr, te, topo, R, maxtemp, R * S, swap, use_lg, lr, l_prob,
                                  0.5, seed, d)
        temps = out["temperatures"]
        pos_w = np.stack([read_chain_file(d, "posterior/pos_w", T) for T in temps])
        accept_list = np.stack([read_chain_file(d, "posterior/accept_list", T) for T in temps])
        pos_lik_file = np.stack([read_chain_file(d, "posterior/pos_likelihood", T) for T in temps])
        rmse_train_file = np.stack([np.loadtxt(os.path.join(d, "predictions", "rmse_train_chain_%s.txt" % T)) for T in temps])
        acc_train_file = np.stack([np.loadtxt(os.path.join(d, "predictions", "acc_train_chain_%s.txt" % T)) for T in temps])
        acc_test_file = np.stack([np.loadtxt(os.path.join(d, "predictions", "acc_test_chain_%s.txt" % T)) for T in temps])
        files = sorted(os.path.relpath(os.path.join(dp, f), d) for dp, _, fs in os.walk(d) for f in fs)
    rounds = out["total_swap_proposals"] // (R - 1)
    assert rounds == cfg.total_rounds(), (rounds, cfg.total_rounds())
    draws = logs_to_draws(task, out["logs"], R, S, P, rounds)
    lik_prop, rmse_tr, rmse_te, prior_prop, init = calls_to_arrays(task, cfg, out["calls"], R, S)
    swapped = np.array(out["swaps"], dtype=bool).reshape(rounds, R - 1)
    res = out["result"]
    # langevin_gradient known answers: first two calls of chain 0 (if any)
    lg_in = np.stack(out["calls"][0]["lg_in"][:2]) if out["calls"][0]["lg_in"] else np.zeros((0, P))
    lg_out = np.stack(out["calls"][0]["lg_out"][:2]) if out["calls"][0]["lg_out"] else np.zeros((0, P))
    fixture = dict(
        task=np.array(cfg.task), dataset=np.array(ds), topology=np.array(topo), R=np.array(R),
        maxtemp=np.array(maxtemp), S=np.array(S), swap_interval=np.array(swap),
        use_lg=np.array(use_lg), learn_rate=np.array(lr), l_prob=np.array(l_prob), seed=np.array(seed),
        temperatures=temps, w0=out["w0"],
        lx=draws.lx.astype(np.float32), z=draws.z.astype(np.float32),
        z_eta=draws.z_eta.astype(np.float32), u=draws.u.astype(np.float32),
        u_swap=draws.u_swap.astype(np.float32),
        ref_pos_w=pos_w, ref_accept_list=accept_list, ref_lik_prop=lik_prop,
        ref_prior_prop=prior_prop, ref_rmse_train=rmse_tr, ref_rmse_test=rmse_te,
        ref_init_lik=init["lik"], ref_init_prior=init["prior"],
        ref_init_rmse_train=init["rmse_train"], ref_init_rmse_test=init["rmse_test"],
        ref_pos_likelihood_file=pos_lik_file, ref_rmse_train_file=rmse_train_file,
        ref_acc_train_file=acc_train_file, ref_acc_test_file=acc_test_file,
        ref_swapped=swapped, ref_num_swap=np.array(out["num_swap"]),
        ref_total_swap_proposals=np.array(out["total_swap_proposals"]),
        ref_swap_perc=np.array(res[8]), ref_accept_vec=np.asarray(res[9]),
        ref_lg_in=lg_in, ref_lg_out=lg_out,
        ref_files=np.array(files),
        ref_result_shapes=np.array([str(getattr(x, "shape", ())) for x in res]),
    )
    # float32 round trip of the draws must be lossless (the recorder only emits f32 values)
    for k in ("lx", "z", "z_eta", "u", "u_swap"):
        assert np.array_equal(fixture[k].astype(np.float64), getattr(draws, k)), k
    np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), **fixture)
    return fixture, cfg, tr, te, draws


def check_oracle(name, fx, cfg, tr, te, draws):
    t = on.run_pt(cfg, tr, te, fx["temperatures"], fx["w0"], draws)
    S = cfg.samples

    def close(a, b, what, tol=1e-9):
        err = np.max(np.abs(a - b) / (1e-300 + np.maximum(1.0, np.abs(b))))
        assert err < tol, (name, what, err)
        return err

    e = [close(t.extra["init_lik"], fx["ref_init_lik"], "init_lik"),
         close(t.extra["init_prior"], fx["ref_init_prior"], "init_prior"),
         close(t.lik_prop_t[:, 1:], fx["ref_lik_prop"][:, 1:], "lik_prop"),
         close(t.prior_prop[:, 1:], fx["ref_prior_prop"][:, 1:], "prior_prop"),
         close(t.pos_w, fx["ref_pos_w"], "pos_w"),
         close(t.accept_list, fx["ref_accept_list"], "accept_list"),
         close(t.rmse_train * t.accepted, fx["ref_rmse_train"] * t.accepted, "rmse_train@accepted")]
    assert np.array_equal(t.swapped, fx["ref_swapped"]), (name, "swapped")
    assert t.num_swap == int(fx["ref_num_swap"]) and t.total_swap_proposals == int(fx["ref_total_swap_proposals"])
    # the file the reference writes holds likeh_list (R:391 tempered / C:404 x adapttemp), %1.4f
    assert np.max(np.abs(t.lik_prop - fx["ref_pos_likelihood_file"][:, :, 0])) < 5.1e-5 + 1e-9 * np.max(np.abs(t.lik_prop))
    print("  %-16s oracle == reference  (max rel err %.2e, swaps %d/%d, accepts %s)" % (
        name, max(e), t.num_swap, t.total_swap_proposals, t.accept_list[:, S - 1].astype(int).tolist()))


def generate_datasets():
    d = {}
    for n in rh.REG_DATASETS:
        tr, te = rh.load_regression_dataset(n)
        d["reg_%s_train" % n], d["reg_%s_test" % n] = tr, te
    for n in ("Iris", "Cancer", "Ionosphere", "PenDigit"):
        tr, te, topo = rh.load_classification_dataset(n, split_seed=0)
        d["cls_%s_train" % n], d["cls_%s_test" % n], d["cls_%s_topology" % n] = tr, te, np.array(topo)
    np.savez_compressed(os.path.join(GOLDEN, "datasets.npz"), **d)
    print("  datasets.npz: %d arrays" % len(d))


def generate_known_answers():
    """Per-function known answers straight from the reference classes (SURVEY Appendix D recipe)."""
    ka = {}
    reg = rh.load_reference("regression")
    cls = rh.load_reference("classification")
    import warnings
    with warnings.catch_warnings(), rh.silenced():
        warnings.simplefilter("ignore")
        for n in rh.REG_DATASETS:
            tr, te = rh.load_regression_dataset(n)
            for H in (5, 10):
                topo = [4, H, 1]
                P = on.num_params(topo)
                w = np.random.RandomState(12345 + H).randn(P)
                net = reg.Network(topo, tr, te, 0.1)
                rep = reg.ptReplica(True, 0.1, w, None, None, 10, tr, te, topo, 0.5, 1.25, 5, 0.5, "", None, None, None)
                fx = net.evaluate_proposal(tr, w)
                tau = float(np.var(fx - tr[:, 4]))
                lik = rep.likelihood_func(net, tr, w, tau)
                lik_te = rep.likelihood_func(net, te, w, tau)
                key = "reg_%s_h%d_" % (n, H)
                ka[key + "w"], ka[key + "fx"], ka[key + "tau"] = w, fx, np.array(tau)
                ka[key + "lik"] = np.array([lik[0], lik[2], lik_te[0], lik_te[2]])
                ka[key + "prior"] = np.array(rep.prior_likelihood(25, 0, 0, w, tau))
                ka[key + "w_gd"] = net.langevin_gradient(tr, w.copy(), 1)
        for n in ("Iris", "Cancer", "Ionosphere"):
            tr, te, topo = rh.load_classification_dataset(n, split_seed=0)
            P = on.num_params(topo)
            w = np.random.RandomState(54321).randn(P)
            net = cls.Network(topo, tr, te, 0.01)
            rep = cls.ptReplica(True, 0.01, w, None, None, 10, tr, te, topo, 0.5, 2.5, 5, "", None, None, None)
            fx, prob = net.evaluate_proposal(tr, w)
            lik = rep.likelihood_func(net, tr, w)
            lik_te = rep.likelihood_func(net, te, w)
            key = "cls_%s_" % n
            ka[key + "w"], ka[key + "fx"], ka[key + "prob"] = w, fx, prob
            ka[key + "lik"] = np.array([lik[0], lik[2], lik_te[0], lik_te[2]])
            ka[key + "acc"] = np.array([rep.accuracy(fx, tr[:, topo[0]]), rep.accuracy(lik_te[1], te[:, topo[0]])])
            ka[key + "prior"] = np.array(rep.prior_likelihood(25, 0, 0, w))
            ka[key + "w_gd"] = net.langevin_gradient(tr, w.copy(), 1)
        # ladder + swap rule (coordinator side)
        pt = reg.ParallelTempering(True, 0.1, tr, te, [4, 5, 1], 10, 2, 1000, 10, 0.5, "")
        pt.assign_temperatures()
        ka["ladder_10_2"] = np.array(pt.temperatures)
        pt = reg.ParallelTempering(True, 0.1, tr, te, [4, 5, 1], 7, 10, 1000, 10, 0.5, "")
        pt.assign_temperatures()
        ka["ladder_7_10"] = np.array(pt.temperatures)
    np.savez_compressed(os.path.join(GOLDEN, "known_answers.npz"), **ka)
    print("  known_answers.npz: %d arrays" % len(ka))


def main(argv):
    os.makedirs(GOLDEN, exist_ok=True)
    if not rh.reference_available():
        raise SystemExit("reference not found under %s" % rh.REFERENCE_ROOT)
    names = argv[1:] or list(CASES) + ["datasets", "known_answers"]
    for n in names:
        if n == "datasets":
            generate_datasets()
        elif n == "known_answers":
            generate_known_answers()
        else:
            print("case", n)
            fx, cfg, tr, te, draws = generate_case(n, CASES[n])
            check_oracle(n, fx, cfg, tr, te, draws)


if __name__ == "__main__":
    main(sys.argv)
