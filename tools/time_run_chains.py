"""Measurement helper (GPU box): wall time of the drop-in ParallelTempering.run_chains() on the reference's
Sunspot configuration (R:883-1007: FNN 4-5-1, 10 chains, maxtemp 2, NumSample 50 000, swap interval 50,
Langevin prob 0.5, lr 0.1, burn-in 0.5), with and without the reference's text files."""
import os, sys, time, tempfile, io, contextlib
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ptnn_b200.regression import ParallelTempering, RESULT_DIRS

d = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "datasets.npz"))
tr, te = d["reg_Sunspot_train"], d["reg_Sunspot_test"]
for write in (False, False, False, True, True, False):   # the first run pays CUDA context creation and module load
    with tempfile.TemporaryDirectory() as path:
        for sub in RESULT_DIRS:
            os.makedirs(path + sub, exist_ok=True)
        np.random.seed(0)
        pt = ParallelTempering(True, 0.1, tr, te, [4, 5, 1], 10, 2, 50000, 50, 0.5, path)
        pt.write_files = write
        pt.results_from_files = write
        pt.seed = 1
        with contextlib.redirect_stdout(io.StringIO()):
            pt.initialize_chains(0.5)
            t0 = time.perf_counter()
            out = pt.run_chains()
            wall = time.perf_counter() - t0
        rmse_tr, rmse_te = out[3], out[4]
        print("write_files=%s: run_chains() %.2f s wall (sampler %.2f s) = %.0f samples/s; swap %.1f %%; rmse train %.4f test %.4f" % (
            write, wall, pt.last_sampler_seconds, 50000 / wall, out[8], rmse_tr.mean(), rmse_te.mean()))
