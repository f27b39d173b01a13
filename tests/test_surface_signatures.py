"""CPU: the drop-in modules expose the reference's class surface -- same class names,
same constructor / method parameter lists.  Compared with the live reference when it is present
(build container); otherwise with the parameter lists recorded here from R:/C:."""
import inspect

import pytest

from oracle import ref_harness as rh
from ptnn_b200 import classification as cls
from ptnn_b200 import regression as reg

EXPECTED = {
    "regression": {
        "Network.__init__": ["self", "Topo", "Train", "Test", "learn_rate"],                                  # R:29
        "Network.evaluate_proposal": ["self", "data", "w"],                                                 # R:120
        "Network.langevin_gradient": ["self", "data", "w", "depth"],                                        # R:99
        "Network.ForwardPass": ["self", "X"], "Network.BackwardPass": ["self", "Input", "desired"],
        "Network.decode": ["self", "w"], "Network.encode": ["self"],
        "ptReplica.__init__": ["self", "use_langevin_gradients", "learn_rate", "w", "minlim_param", "maxlim_param",
                               "samples", "traindata", "testdata", "topology", "burn_in", "temperature",
                               "swap_interval", "langevin_prob", "path", "parameter_queue", "main_process", "event"],  # R:140
        "ptReplica.likelihood_func": ["self", "fnn", "data", "w", "tau_sq"],                                # R:200
        "ptReplica.prior_likelihood": ["self", "sigma_squared", "nu_1", "nu_2", "w", "tausq"],              # R:215
        "ptReplica.rmse": ["self", "pred", "actual"], "ptReplica.run": ["self"],
        "ParallelTempering.__init__": ["self", "use_langevin_gradients", "learn_rate", "traindata", "testdata",
                                       "topology", "num_chains", "maxtemp", "NumSample", "swap_interval",
                                       "langevin_prob", "path"],                                          # R:489
        "ParallelTempering.default_beta_ladder": ["self", "ndim", "ntemps", "Tmax"],
        "ParallelTempering.assign_temperatures": ["self"], "ParallelTempering.initialize_chains": ["self", "burn_in"],
        "ParallelTempering.swap_procedure": ["self", "parameter_queue_1", "parameter_queue_2"],
        "ParallelTempering.run_chains": ["self"], "ParallelTempering.show_results": ["self"],
        "ParallelTempering.make_directory": ["self", "directory"],
    },
    "classification": {
        "Network.__init__": ["self", "Topo", "Train", "Test", "learn_rate"],                                  # C:28
        "Network.evaluate_proposal": ["self", "data", "w"], "Network.langevin_gradient": ["self", "data", "w", "depth"],
        "Network.softmax": ["self"],
        "ptReplica.__init__": ["self", "use_langevin_gradients", "learn_rate", "w", "minlim_param", "maxlim_param",
                               "samples", "traindata", "testdata", "topology", "burn_in", "temperature",
                               "swap_interval", "path", "parameter_queue", "main_process", "event"],       # C:159
        "ptReplica.likelihood_func": ["self", "fnn", "data", "w"],                                          # C:209
        "ptReplica.prior_likelihood": ["self", "sigma_squared", "nu_1", "nu_2", "w"],                       # C:224
        "ptReplica.accuracy": ["self", "pred", "actual"],
        "ParallelTempering.__init__": ["self", "use_langevin_gradients", "learn_rate", "traindata", "testdata",
                                       "topology", "num_chains", "maxtemp", "NumSample", "swap_interval", "path"],  # C:499
        "ParallelTempering.swap_procedure": ["self", "parameter_queue_1", "parameter_queue_2"],
        "ParallelTempering.run_chains": ["self"], "ParallelTempering.initialize_chains": ["self", "burn_in"],
    },
}


def _params(mod, dotted):
    c, m = dotted.split(".")
    return list(inspect.signature(getattr(getattr(mod, c), m)).parameters)


@pytest.mark.parametrize("which,mod", [("regression", reg), ("classification", cls)])
def test_surface_matches_recorded_signatures(which, mod):
    for dotted, params in EXPECTED[which].items():
        assert _params(mod, dotted) == params, dotted


@pytest.mark.parametrize("which,mod", [("regression", reg), ("classification", cls)])
def test_surface_matches_live_reference(which, mod):
    if not rh.reference_available():
        pytest.skip("reference checkout not present (GPU box)")
    ref = rh.load_reference(which)
    for dotted in EXPECTED[which]:
        assert _params(mod, dotted) == _params(ref, dotted), dotted
    for cname in ("Network", "ptReplica", "ParallelTempering"):
        ref_methods = {n for n, v in vars(getattr(ref, cname)).items() if callable(v) and not n.startswith("_")}
        ours = {n for n in dir(getattr(mod, cname)) if not n.startswith("_")}
        assert ref_methods <= ours, (cname, sorted(ref_methods - ours))


def test_ladder_matches_reference_known_answer():
    from tests import common as cm
    import numpy as np
    ka = cm.npz("known_answers")
    pt = reg.ParallelTempering.__new__(reg.ParallelTempering)
    pt._init_common(True, 0.1, np.zeros((2, 5)), np.zeros((2, 5)), [4, 5, 1], 10, 2, 1000, 10, 0.5, "")
    pt.assign_temperatures()
    assert np.allclose(pt.temperatures, ka["ladder_10_2"], rtol=1e-15, atol=0)
    assert pt.NumSamples == 100 and pt.num_param == 31
    with pytest.raises(ValueError):
        pt.default_beta_ladder(2, ntemps=10, Tmax=1)                              # R:544-545


def test_classification_loaders_follow_the_reference_lines():
    """load_problem == the reference's own preprocessing lines (C:909-930, C:1001-1012) executed here on its
    data files, including the test-set slice quirk (DESIGN Q17).  Skipped where /root/reference is absent."""
    import os
    import numpy as np
    root = "/root/reference/multicore-pt-classification"
    if not os.path.isdir(os.path.join(root, "DATA")):
        pytest.skip("reference data not present")
    for problem, fname, ip, shift, skip in ((1, "winequality-red.csv", 11, 0, 1), (2, "winequality-white.csv", 11, 0, 1), (3, "iris.csv", 4, 1, 0)):
        np.random.seed(5)
        data = np.genfromtxt(os.path.join(root, "DATA", fname), delimiter=';')[skip:, :]
        classes = data[:, ip].reshape(data.shape[0], 1) - shift
        features = data[:, 0:ip]
        for k in range(ip):                                                        # C:1003-1007
            features[:, k] = (features[:, k] - np.mean(features[:, k])) / np.std(features[:, k])
        indices = np.random.permutation(features.shape[0])
        n = int(0.7 * features.shape[0])
        traindata = np.hstack([features[indices[:n], :], classes[indices[:n], :]])
        testdata = np.hstack([features[indices[n]:, :], classes[indices[n]:, :]])   # C:1012, as written
        np.random.seed(5)
        _, tr, te, topo = cls.load_problem(problem, root)
        assert np.array_equal(tr, traindata) and np.array_equal(te, testdata) and topo[0] == ip
        np.random.seed(5)
        _, tr2, te2, _ = cls.load_problem(problem, root, complementary_test_split=True)
        assert np.array_equal(tr2, traindata) and te2.shape[0] == features.shape[0] - n
    for problem, shape in ((4, (245, 35)), (5, (489, 10)), (7, (7494, 17))):
        assert cls.load_problem(problem, root)[1].shape == shape
    with pytest.raises(ValueError):
        cls.load_problem(6, root)                                                  # Bank: bank-processed.csv is not in the repository
