"""Multi-GPU parity (run with `-m gpu` on a box with >= 2 GPUs; skipped otherwise).

Shard-count invariance + oracle parity: the same circuits on 2, 4 and 8 GPUs must give the
reference's amplitudes to 1e-12 and its measurement outcomes exactly (tests/sharded_worker.py).
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _gpu_count():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_register_matches_oracle(world):
    if _gpu_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    port = 29600 + world
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "sharded_worker.py")],
                         capture_output=True, text=True, env=env, timeout=1500)
    assert res.returncode == 0, (res.stdout[-4000:] + res.stderr[-4000:])
    assert f"SHARDED_OK {world}" in res.stdout
