"""GPU (needs >= 2 devices; skipped otherwise): ladder partitioned over ranks == single-GPU run,
bit for bit, in replay mode and in free-running (Philox) mode, with the swap round completed on the
device through peer memory (NVLink) and with the host-completed round (NCCL)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run_worker(mode, exchange, reinit, spec=1):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    env = dict(os.environ, PT_TEST_MODE=mode, PT_TEST_EXCHANGE=exchange, PT_TEST_REINIT="1" if reinit else "0", PT_TEST_SPEC=str(spec))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "DIST_GPU_RESULT mode=%s exchange=%s ok=True" % (mode, exchange) in out.stdout, out.stdout[-2000:]
    return out.stdout


@pytest.mark.parametrize("exchange", ["peer", "host"])
@pytest.mark.parametrize("mode", ["replay", "free"])
def test_partitioned_ladder_bit_identical_to_one_gpu(mode, exchange):
    _run_worker(mode, exchange, reinit=False)


@pytest.mark.parametrize("exchange", ["peer", "host"])
def test_second_run_on_connected_handles(exchange):
    """ptfnn_init_chains again on peer-connected handles: the arrival flags of the first run (written remotely,
    never reset) must not satisfy the waits of the second -- the flag domain moves on with every init."""
    out = _run_worker("free", exchange, reinit=True)
    assert "DIST_GPU_PASS second" in out and "DIST_GPU_PASS second mode=free exchange=%s ok=True" % exchange in out


@pytest.mark.parametrize("mode", ["replay", "free"])
def test_speculative_windows_on_a_partitioned_ladder(mode):
    """Speculative windows (several CTAs per temperature) together with the peer-memory swap round: every rank runs
    windows of depth 3, the single-GPU comparison runs sequentially -- bit-identical traces, swaps and final state."""
    _run_worker(mode, "peer", reinit=False, spec=3)
