"""CPU, world_size 2 (gloo): the ladder partitioned over ranks with boundary swaps reproduces the
single-process run bit for bit.  The exchange logic under test is the product's
(ptnn_b200.distributed.PartitionedLadder); only the chains are oracle-backed."""
import os
import socket
import subprocess
import sys

import pytest

from ptnn_b200.distributed import combine_summaries, partition

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world,R", [(2, 4), (2, 6)])
def test_partitioned_ladder_matches_single_process(world, R):
    env = dict(os.environ, PT_TEST_R=str(R), PT_TEST_S="31", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gloo_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_GLOO_RESULT ok=True" in out.stdout, out.stdout[-2000:]


def test_partition_is_contiguous_and_even():
    assert [partition(1024, 8, r) for r in (0, 3, 7)] == [(0, 128), (384, 128), (896, 128)]
    with pytest.raises(ValueError):
        partition(10, 4, 0)


def test_per_rank_summaries_pool_to_the_whole_ladder():
    """combine_summaries (what PartitionedLadder.summary applies to the all-gathered per-rank results)
    == the reference's statistics over all chains at once (R:1036-1044)."""
    import numpy as np
    rs = np.random.RandomState(2)
    blocks = [rs.rand(n, 7) * (k + 1) + k for k, n in enumerate((40, 25, 1))]       # unequal blocks, one of a single row
    names = ("rmse_train", "rmse_test", "acc_train", "acc_test")

    def summarise(x):
        d = {"n": len(x), "w_mean": x[:, 4:].mean(axis=0), "w_std": x[:, 4:].std(axis=0)}
        for j, k in enumerate(names):
            d[k] = {"mean": x[:, j].mean(), "std": x[:, j].std(), "min": x[:, j].min(), "max": x[:, j].max()}
        return d

    got, want = combine_summaries([summarise(b) for b in blocks] + [None]), summarise(np.vstack(blocks))
    assert got["n"] == want["n"] == 66
    for k in names:
        assert np.allclose([got[k][q] for q in ("mean", "std", "min", "max")], [want[k][q] for q in ("mean", "std", "min", "max")], rtol=1e-12)
    assert np.allclose(got["w_mean"], want["w_mean"], rtol=1e-12) and np.allclose(got["w_std"], want["w_std"], rtol=1e-12)
    parts = [summarise(b) for b in blocks]
    parts[1]["w_mean"] = parts[1]["w_std"] = None
    assert combine_summaries(parts)["w_mean"] is None
