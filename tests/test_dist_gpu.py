"""Multi-GPU parity (needs >= 2 GPUs on the box): launches tests/dist_worker.py under torch.distributed.run."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [1, 2])
def test_distributed_factorisation_parity(world):
    """world = 1 runs the same worker on the single GPU of the driver's test box with the panel schedule forced
    (CHEQ_FACTOR_MODE=1): the multi-GPU factorisation code path (right-looking panels, panel / bulk / main streams,
    look-ahead events) without the NCCL exchange, which needs a second device."""
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    if world == 1:
        env["CHEQ_FACTOR_MODE"] = "1"
        env["CHEQ_DIST_TEST_BIG"] = "1200"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", str(ROOT / "tests" / "dist_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    sys.stdout.write(res.stdout[-4000:])
    sys.stderr.write(res.stderr[-4000:])
    assert res.returncode == 0 and "DIST PARITY PASS" in res.stdout
