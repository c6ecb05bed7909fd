"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): row-sharded operators and
vector blocks, NCCL all-gather / all-reduce inside the library, results bit-identical to the
sequential oracle for any number of ranks."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_run_matches_oracle(lib, world):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and f"MGPU_OK world={world}" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.skipif(not os.environ.get("BLK_TEST_EXPERIMENTAL"), reason="experimental path: set BLK_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("world,K", [(2, 2), (2, 4), (4, 3)])
def test_arrival_order_exchange_matches_oracle(lib, world, K):
    """BLK_COLBLOCKS=K with several GPUs: column-blocked consumers, pieces broadcast as they finish."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29520 + world + K), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, BLK_COLBLOCKS=str(K)))
    assert r.returncode == 0 and f"MGPU_OK world={world}" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.skipif(not os.environ.get("BLK_TEST_EXPERIMENTAL"), reason="experimental path: set BLK_TEST_EXPERIMENTAL=1")
@pytest.mark.parametrize("world,grid", [(2, "2x1"), (2, "1x2"), (4, "2x2"), (4, "auto"), (4, "4x1")])
def test_block_grid_matches_oracle(lib, world, grid):
    """BLK_GRID=PxQ: the loop runs on the P x Q block grid (all-gathers inside grid rows / columns, reduce-scatters
    mod p), everything else on the 1-D blocks after grid_export; specification: tests/test_grid_cpu.py."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29540 + world), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, BLK_GRID=grid))
    assert r.returncode == 0 and f"MGPU_OK world={world}" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
