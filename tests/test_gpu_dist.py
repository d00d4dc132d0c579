"""GPU: destination-partitioned path (parallel.py).  world=1 runs in every `-m gpu` session; world=2/4/8
need that many visible GPUs (gpurun --gpus N) and are skipped otherwise; the logs of the builder's own
multi-GPU runs are under profiles/ (r02_*_pytest_dist_n*.txt)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, port):
    env = dict(os.environ)
    env.pop("RANK", None), env.pop("WORLD_SIZE", None)
    if world == 1:
        cmd = [sys.executable, os.path.join(HERE, "dist_worker.py")]
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(HERE, "dist_worker.py")]
    res = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert f"world={world} OK" in res.stdout, res.stdout[-2000:]


def test_partitioned_world1_equals_single_gpu():
    _run(1, 29551)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_multi_gpu_equals_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _run(world, 29553 + world)
