"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): row-sharded operators and
vector blocks, exchange of Av and tmp inside the library (stores over NVLink through peer mappings,
or NCCL), results bit-identical to the sequential oracle for any number of ranks, for both process
models: one process per GPU (torchrun, CUDA IPC) and one process for all GPUs (BLK_RANK_ALL)."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
import mgpu_cases

pytestmark = pytest.mark.gpu

MODES = {                      # environment of the exchange modes
    "push": {},                                             # default: k_spmv PUSH for Av, k_push_rows for tmp
    "push_kernel": {"BLK_PUSH_AV": "kernel"},               # k_push_rows for both
    "nccl": {"BLK_EXCHANGE": "nccl"},                       # grouped ncclBroadcast of the pieces
    "ce": {"BLK_EXCHANGE": "ce"},                           # copy-engine pushes
    "norecur": {"BLK_RECUR": "0"},                          # plain all-gathers of v and tmp
    "pieces1": {"BLK_PIECES": "1"},
    "pieces8": {"BLK_PIECES": "8"},                         # shards disagree on the piece count for the small cases
    # degree-sorted labels dealt to the ranks + hot prefix, forced on for test-sized matrices (default: blocks > 96 MB)
    "hot": {"BLK_HOT_MIN_BYTES": "0", "BLK_HOT_BYTES": "8192"},
    "hot_nccl": {"BLK_HOT_MIN_BYTES": "0", "BLK_HOT_BYTES": "8192", "BLK_EXCHANGE": "nccl"},
    # pieces pushed by the bulk-copy engine (k_push_bulk) instead of k_push_rows
    # column-banded products (n_pad <= 4) forced on for the test-sized matrices: every rank's operators in bands
    "bands": {"BLK_BAND_BYTES": "8192"},
    "bulk": {"BLK_PUSH_COPY": "bulk"},
    "bulk_kernel": {"BLK_PUSH_COPY": "bulk", "BLK_PUSH_AV": "kernel", "BLK_PUSH_CTAS": "3"},
}


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _torchrun(world, port, env):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, **env))
    assert r.returncode == 0 and f"MGPU_OK world={world}" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_run_matches_oracle(lib, world):
    """One process per GPU, default exchange (peer mappings through CUDA IPC)."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _torchrun(world, 29500 + world, {})


@pytest.mark.parametrize("mode", ["nccl", "push_kernel", "hot", "bulk_kernel", "bands"])
def test_exchange_modes_one_process_per_gpu(lib, mode):
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    _torchrun(2, 29560 + list(MODES).index(mode), MODES[mode])


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("mode", list(MODES))
def test_group_context_matches_oracle(lib, oracle, monkeypatch, world, mode):
    """One process for all GPUs (blk_params.rank = BLK_RANK_ALL): one host thread per GPU inside the
    library, peer access between the devices of the process."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    if world > 2 and mode not in ("push", "nccl", "pieces8", "hot", "bulk"):
        pytest.skip("mode covered at world 2")
    for k in ("BLK_EXCHANGE", "BLK_PUSH_AV", "BLK_RECUR", "BLK_PIECES", "BLK_HOT_MIN_BYTES", "BLK_HOT_BYTES", "BLK_PUSH_COPY", "BLK_PUSH_CTAS", "BLK_BAND_BYTES", "BLK_BAND_ACC"):
        monkeypatch.delenv(k, raising=False)
    for k, v in MODES[mode].items():
        monkeypatch.setenv(k, v)
    for ci, (M, n, p, right, stop_after) in enumerate(mgpu_cases.cases(lib)):
        with lib.BlockLanczos(M.reduced(p), n=n, prime=p, right=right, rank=lib.BLK_RANK_ALL, world=world) as ctx:
            info = ctx.info()
            assert (info["local_N0"], info["local_N1"]) == (0, M.ncols if right else M.nrows)
            mgpu_cases.check_context(lib, oracle, ctx, M, n, p, right, stop_after, (mode, world, ci), seed=100 + ci)


def test_group_context_runtime_check(lib, oracle, monkeypatch):
    """BLK_CHECK=1 (correctness_tests, sequential/lanczos_modp.c:532-557) on a sharded run."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("BLK_CHECK", "1")
    M, n, p, right, stop_after = mgpu_cases.cases(lib)[0]
    with lib.BlockLanczos(M.reduced(p), n=n, prime=p, right=right, rank=lib.BLK_RANK_ALL, world=2) as ctx:
        mgpu_cases.check_context(lib, oracle, ctx, M, n, p, right, stop_after, "check")
    monkeypatch.setenv("BLK_CHECK_FAULT", "4")
    with lib.BlockLanczos(M.reduced(p), n=n, prime=p, right=right, rank=lib.BLK_RANK_ALL, world=2) as ctx:
        ctx.set_state(oracle.start_block(M.nrows * n, p))
        with pytest.raises(lib.BlkError, match="correctness_tests failed in iteration 4"):
            ctx.iterate(9)


def test_cli_on_several_gpus_is_byte_identical(lib, tmp_path):
    """BLK_GPUS=G: the kept command line on G GPUs writes the same kernel file as on one."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    driver = os.path.join(ROOT, "block-lanczos-algorithm-parallelization_b200", "driver", "lanczos_modp")
    M = lib.synth.powerlaw_rows(2800, 3000, mean=9, seed=21, with_empty_rows=7)
    mtx = str(tmp_path / "m.mtx")
    lib.synth.write_mtx(mtx, M)
    hashes = {}
    for g in [1, 2] + ([4] if _ngpus() >= 4 else []) + ([8] if _ngpus() >= 8 else []):
        out = str(tmp_path / f"k{g}.mtx")
        env = dict(os.environ, BLK_GPUS=str(g))
        r = subprocess.run([driver, "--matrix", mtx, "--prime", "2147483647", "--n", "16", "--right", "--output-file", out],
                           capture_output=True, text=True, env=env, cwd=str(tmp_path))
        assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-2000:])
        assert "OK:    v != 0" in r.stdout and "OK: vt*M == 0" in r.stdout
        hashes[g] = hashlib.sha256(open(out, "rb").read()).hexdigest()
    assert len(set(hashes.values())) == 1, hashes
    # checkpoint + resume on 2 GPUs ends in the same file
    env = dict(os.environ, BLK_GPUS="2")
    args = [driver, "--matrix", mtx, "--prime", "2147483647", "--n", "16", "--right"]
    r = subprocess.run(args + ["--checkpoint", "0", "--stop-after", "20"], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert open(tmp_path / "checkpoint.commit").read().split() == ["complete", "20"]
    out = str(tmp_path / "resumed.mtx")
    r = subprocess.run(args + ["--load-checkpoint", "--output-file", out], capture_output=True, text=True, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-2000:]
    assert hashlib.sha256(open(out, "rb").read()).hexdigest() == hashes[1]


# ---- validated, non-default modes (need BLK_EXPERIMENTAL=1 next to their own switch) ----------------------
@pytest.mark.parametrize("world,K", [(2, 2), (2, 4), (4, 3)])
def test_arrival_order_exchange_matches_oracle(lib, world, K):
    """BLK_COLBLOCKS=K with several GPUs: column-blocked consumers, pieces broadcast as they finish."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _torchrun(world, 29520 + world + K, {"BLK_EXPERIMENTAL": "1", "BLK_COLBLOCKS": str(K)})


@pytest.mark.parametrize("world,grid", [(2, "2x1"), (2, "1x2"), (4, "2x2"), (4, "auto"), (4, "4x1"), (8, "4x2")])
def test_block_grid_matches_oracle(lib, world, grid):
    """BLK_GRID=PxQ: the loop runs on the P x Q block grid (all-gathers inside grid rows / columns, reduce-scatters
    mod p), everything else on the 1-D blocks after grid_export; specification: tests/test_grid_cpu.py."""
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    _torchrun(world, 29540 + world + len(grid), {"BLK_EXPERIMENTAL": "1", "BLK_GRID": grid})


def test_experimental_modes_need_the_opt_in(lib, monkeypatch):
    monkeypatch.delenv("BLK_EXPERIMENTAL", raising=False)
    monkeypatch.setenv("BLK_COLBLOCKS", "2")
    M = lib.synth.uniform_rows(300, 280, 5, seed=8)
    with pytest.raises(lib.BlkError, match="BLK_EXPERIMENTAL"):
        lib.BlockLanczos(M.reduced(65537), n=4, prime=65537)
