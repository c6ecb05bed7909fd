"""Worker for tests/test_gpu_multi.py: one process per GPU (torchrun), exchange inside the library
(peer-mapped stores over NVLink or NCCL, see BLK_EXCHANGE).  Every rank checks its results against the
CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import blk_lanczos_b200 as B          # noqa: E402
from oracle.oracle import Oracle      # noqa: E402
import mgpu_cases                     # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    O = Oracle()
    only = os.environ.get("BLK_TEST_CASES")
    for ci, (M, n, p, right, stop_after) in enumerate(mgpu_cases.cases(B)):
        if only and str(ci) not in only.split(","):
            continue
        N = M.ncols if right else M.nrows
        ident = [B.nccl_unique_id() if rank == 0 else None]      # one id per communicator
        dist.broadcast_object_list(ident, src=0)
        ctx = B.BlockLanczos(M.reduced(p), n=n, prime=p, right=right, device=local, rank=rank, world=world, nccl_id=ident[0])
        info = ctx.info()
        assert (info["local_N0"], info["local_N1"]) != (0, N) or world == 1
        want = mgpu_cases.check_context(B, O, ctx, M, n, p, right, stop_after, (ci, rank), seed=100 + ci)
        # blk_get_state_local: only this rank's rows are written
        pad = ctx.pad
        v = np.full(pad, 0xdeadbeef, dtype=np.uint32)
        pb = np.full(pad, 0xdeadbeef, dtype=np.uint32)
        ctx.get_state_local(v=v, p_blk=pb)
        lo, hi = info["local_N0"] * n, info["local_N1"] * n
        assert np.array_equal(v[lo:hi], want["v"][lo:hi]) and np.array_equal(pb[lo:hi], want["p"][lo:hi]), ("local", ci, rank)
        assert (v[:lo] == 0xdeadbeef).all() and (v[hi:N * n] == 0xdeadbeef).all()
        ctx.close()
        dist.barrier()
    if rank == 0:
        print("MGPU_OK world=%d" % world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
