"""Shared by the multi-GPU tests: the cases and the per-context checks against the CPU oracle.  Used by
tests/mgpu_worker.py (one process per GPU under torchrun: CUDA-IPC peer mappings, NCCL rendezvous by the
caller) and by the single-process group-context tests (one thread per GPU inside the library)."""
import numpy as np


def cases(B):
    s = B.synth
    return [
        (s.powerlaw_rows(3000, 2700, mean=9, seed=3, with_empty_rows=11), 4, 65537, False, 9),
        (s.uniform_nnz(2500, 3100, 30000, seed=4, order="col"), 8, 2147483647, True, 7),
        (s.powerlaw_rows(1500, 1400, mean=6, seed=5), 3, 1073741789, False, -1),      # run to the end
        (s.powerlaw_rows(1200, 1500, mean=7, seed=6), 4, 65537, False, 6),             # Mc > N: tmp keeps old rows
        (s.uniform_nnz(40000, 52000, 600000, seed=7, order="col"), 16, 2147483647, True, 5),   # many tiles: several pieces
        (s.uniform_nnz(5, 7, 20, seed=8), 2, 65537, False, 3),                         # fewer rows than ranks can share evenly
    ]


def check_context(B, O, ctx, M, n, p, right, stop_after, tag, seed=100):
    """Every multi-GPU entry point of one context against the oracle (bit-exact)."""
    Mp = M.reduced(p)
    N = M.ncols if right else M.nrows
    rng = np.random.default_rng(seed)
    for tr in (False, True):
        cols = M.nrows if tr else M.ncols
        x = rng.integers(0, p, size=cols * n).astype(np.uint32)
        got = ctx.sparse_matrix_vector_product(x, tr)
        want = O.sparse_matrix_vector_product(Mp, x, tr, n, p)
        assert np.array_equal(got, want), ("spmv", tag, tr)
    v0 = O.start_block(N * n, p)
    got = ctx.block_lanczos(v0, stop_after=stop_after, batch=5)
    want = O.lanczos_run(Mp, n, p, right, stop_after=stop_after)
    assert got["iters"] == want["iters"] and got["stopped"] == want["stopped"], (tag, got["iters"], want["iters"])
    for k in ("v", "tmp", "Av", "p"):
        assert np.array_equal(got[k], want[k]), ("loop", tag, k)
    fc = ctx.final_check()                      # device-side final_check, reduced over the ranks
    assert fc == (bool(want["v"].any()), not O.sparse_matrix_vector_product(Mp, want["v"], not right, n, p).any()), (tag, fc)
    if want["stopped"]:
        assert ctx.check_kernel_block(want["v"][:N * n]) and not ctx.check_kernel_block(v0)
    # resume from the middle: set_state(v, p, k) reads only the rank's rows and rebuilds the loop invariant
    if stop_after > 3:
        half = O.lanczos_run(Mp, n, p, right, stop_after=3)
        again = ctx.block_lanczos(half["v"], stop_after=stop_after, batch=4, p0=half["p"], n_iterations=3)
        for k in ("v", "Av", "p"):
            assert np.array_equal(again[k], want[k]), ("resume", tag, k)
    return want
