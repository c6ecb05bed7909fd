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
