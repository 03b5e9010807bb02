"""Multi-GPU parity as a -m gpu test: launches tests/mgpu_check.py (row-sharded solves — l1 / FISTA, AoRR hinge / l2,
EHRM — against the unsharded CPU oracle at 1e-9, plus the sharded test-set metrics) under torchrun on every GPU
count in {2, 4, 8} the box offers.  Skipped on a single-GPU box (the driver's default GPU test lease)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_parity_under_torchrun(world):
    import torch

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    env = dict(os.environ)
    env.pop("RBL_W_MODE", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), os.path.join(ROOT, "tests", "mgpu_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-4000:])
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count("-> OK") >= 5 and "FAIL" not in res.stdout
