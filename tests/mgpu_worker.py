"""Worker for tests/test_gpu_multi.py: one process per GPU (torchrun), NCCL exchange inside the
library.  Every rank checks its results against the CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import blk_lanczos_b200 as B          # noqa: E402
from oracle.oracle import Oracle      # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    O = Oracle()
    cases = [
        (B.synth.powerlaw_rows(3000, 2700, mean=9, seed=3, with_empty_rows=11), 4, 65537, False, 9),
        (B.synth.uniform_nnz(2500, 3100, 30000, seed=4, order="col"), 8, 2147483647, True, 7),
        (B.synth.powerlaw_rows(1500, 1400, mean=6, seed=5), 3, 1073741789, False, -1),      # run to the end
        (B.synth.powerlaw_rows(1200, 1500, mean=7, seed=6), 4, 65537, False, 6),             # Mc > N: tmp keeps old rows
        (B.synth.uniform_nnz(40000, 52000, 600000, seed=7, order="col"), 16, 2147483647, True, 5),   # many tiles: 8 pieces
    ]
    for ci, (M, n, p, right, stop_after) in enumerate(cases):
        Mp = M.reduced(p)
        N = M.ncols if right else M.nrows
        ident = [B.nccl_unique_id() if rank == 0 else None]      # one id per communicator
        dist.broadcast_object_list(ident, src=0)
        ctx = B.BlockLanczos(Mp, n=n, prime=p, right=right, device=local, rank=rank, world=world, nccl_id=ident[0])
        info = ctx.info()
        assert (info["local_N0"], info["local_N1"]) != (0, N) or world == 1
        rng = np.random.default_rng(100 + ci)
        for tr in (False, True):
            cols = M.nrows if tr else M.ncols
            x = rng.integers(0, p, size=cols * n).astype(np.uint32)
            got = ctx.sparse_matrix_vector_product(x, tr)
            want = O.sparse_matrix_vector_product(Mp, x, tr, n, p)
            assert np.array_equal(got, want), ("spmv", ci, tr, rank)
        v0 = O.start_block(N * n, p)
        got = ctx.block_lanczos(v0, stop_after=stop_after, batch=5)
        want = O.lanczos_run(Mp, n, p, right, stop_after=stop_after)
        assert got["iters"] == want["iters"] and got["stopped"] == want["stopped"], (ci, got["iters"], want["iters"])
        for k in ("v", "tmp", "Av", "p"):
            assert np.array_equal(got[k], want[k]), ("loop", ci, k, rank)
        Mc = M.nrows if right else M.ncols
        fc = ctx.final_check()                      # device-side final_check, reduced over the ranks
        assert fc == (bool(want["v"].any()), not O.sparse_matrix_vector_product(Mp, want["v"], not right, n, p).any()), (ci, fc)
        if want["stopped"]:
            assert ctx.check_kernel_block(want["v"][:N * n]) and not ctx.check_kernel_block(v0)
        ctx.close()
        dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
