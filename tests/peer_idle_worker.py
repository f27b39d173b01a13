"""Helper of test_a_wait_that_gives_up_fails_closed: a "rank 1" that exports its swap window and then never runs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ptfnn_numpy as on                      # noqa: E402
from ptnn_b200.sampler import Sampler, geometric_ladder    # noqa: E402
from tests import common as cm                            # noqa: E402


def main():
    R, S, si = 4, 30, 5
    tr, te = cm.dataset(on.REGRESSION, "Sunspot")
    temps = geometric_ladder(2 * R, 2)
    with Sampler(on.REGRESSION, (4, 5, 1), temps[R:], S, si, n_replicas_global=2 * R, replica_offset=R, seed=3) as s:
        s.set_data(tr, te)
        s.init_chains(np.random.RandomState(1).randn(R, 31))
        print("HANDLES " + s.peer_export().hex(), flush=True)
        sys.stdin.readline()                              # stay alive (the window stays mapped) until told to leave


if __name__ == "__main__":
    main()
