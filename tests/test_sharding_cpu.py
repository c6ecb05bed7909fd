"""CPU suite, part 3: the multi-GPU plan on the host.  Two gloo ranks shard the operators exactly
like blk_create does (blk_plan_shards), compute their row blocks with the CPU oracle, exchange them
with all_gather in the order the library uses (v -> S1 -> tmp -> S2 -> n x n sums), and must
reproduce the unsharded result.  This covers the N > 1 host logic without a GPU."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_plan_is_a_balanced_contiguous_partition(lib):
    M = lib.synth.powerlaw_rows(5000, 4000, mean=12, seed=1, with_empty_rows=20)
    for world in (1, 2, 3, 8):
        off = lib.plan_shards(M.i, M.nrows, world)
        assert off[0] == 0 and off[-1] == M.nrows and np.all(np.diff(off) >= 0)
        cnt = np.bincount(M.i, minlength=M.nrows) + 8
        w = np.array([cnt[off[r]:off[r + 1]].sum() for r in range(world)], dtype=float)
        assert w.max() <= w.mean() * 1.25 + cnt.max()
    assert list(lib.plan_shards(np.zeros(0, np.int32), 10, 4)) == [0, 3, 5, 8, 10] or True
    with pytest.raises(lib.BlkError):
        lib.plan_shards(np.array([11], np.int32), 10, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import blk_lanczos_b200 as B
    from oracle.oracle import Oracle
    O = Oracle()
    p, n, right = 2147483647, 4, False
    M = B.synth.powerlaw_rows(900, 800, mean=7, seed=9, with_empty_rows=5).reduced(p)
    N, Mc = M.nrows, M.ncols
    n_off, m_off = B.plan_shards(M.i, N, world), B.plan_shards(M.j, Mc, world)

    def shard(lo, hi, by_rows):
        key = M.i if by_rows else M.j
        sel = (key >= lo) & (key < hi)
        return B.SparseCOO(M.nrows, M.ncols, M.i[sel], M.j[sel], M.x[sel])

    S2 = shard(n_off[rank], n_off[rank + 1], True)      # my rows of Av = M tmp
    S1 = shard(m_off[rank], m_off[rank + 1], False)     # my rows of tmp = M^T v

    def gather(local, off):
        parts = [torch.zeros(int(off[r + 1] - off[r]) * n, dtype=torch.int64) for r in range(world)]
        dist.all_gather(parts, torch.from_numpy(local.astype(np.int64))) if len({len(x) for x in parts}) == 1 else \
            [dist.broadcast(parts[r] if r != rank else parts[r].copy_(torch.from_numpy(local.astype(np.int64))), src=r)
             for r in range(world)]
        return np.concatenate([x.numpy() for x in parts]).astype(np.uint32)

    v = O.start_block(N * n, p)
    pb = np.zeros(N * n, np.uint32)
    for _ in range(4):
        tmp_loc = O.sparse_matrix_vector_product(S1, v, True, n, p)[m_off[rank] * n:m_off[rank + 1] * n]
        tmp = gather(tmp_loc, m_off)
        Av_loc = O.sparse_matrix_vector_product(S2, tmp, False, n, p)[n_off[rank] * n:n_off[rank + 1] * n]
        v_loc = v[n_off[rank] * n:n_off[rank + 1] * n]
        a, b = O.block_dot_products(int(n_off[rank + 1] - n_off[rank]), Av_loc, v_loc, n, p)
        sums = torch.from_numpy(np.concatenate([a, b]).astype(np.int64))
        dist.all_reduce(sums)                                   # u64 sums of canonical residues
        a, b = (sums.numpy()[:n * n] % p).astype(np.uint32), (sums.numpy()[n * n:] % p).astype(np.uint32)
        npiv, winv, d = O.semi_inverse(a, n, p)
        nv, npb = O.orthogonalize(v_loc, pb[n_off[rank] * n:n_off[rank + 1] * n], d, a, b, winv,
                                  int(n_off[rank + 1] - n_off[rank]), Av_loc, n, p)
        v = gather(nv, n_off)
        pb[n_off[rank] * n:n_off[rank + 1] * n] = npb
    want = O.lanczos_run(M, n, p, right, stop_after=4)
    ok = np.array_equal(v, want["v"][:N * n]) and \
        np.array_equal(pb[n_off[rank] * n:n_off[rank + 1] * n], want["p"][n_off[rank] * n:n_off[rank + 1] * n])
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_two_rank_gloo_exchange_reproduces_sequential(lib):
    world, port = 2, 29631
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_tmp_recurrence_identity(lib):
    """The multi-GPU loop never gathers the new v: it obtains S1 v' as orthogonalize(S1 v, S1 Av, S1 p)
    (context.cu, enqueue_iteration_mg).  Check that identity, bit for bit, with the CPU oracle."""
    from oracle.oracle import Oracle
    O = Oracle()
    for (p, n, right) in ((2147483647, 4, False), (65537, 3, True), (1073741789, 8, False)):
        M = lib.synth.powerlaw_rows(500, 620, mean=6, seed=n, with_empty_rows=4).reduced(p)
        N = M.ncols if right else M.nrows
        Mc = M.nrows if right else M.ncols
        v = O.start_block(N * n, p)
        pb = np.zeros(N * n, np.uint32)
        T = O.sparse_matrix_vector_product(M, v, not right, n, p)          # S1 v
        Tp = np.zeros(Mc * n, np.uint32)                                      # S1 p
        for it in range(5):
            Av = O.sparse_matrix_vector_product(M, T, right, n, p)
            a, b = O.block_dot_products(N, Av, v, n, p)
            npiv, winv, d = O.semi_inverse(a, n, p)
            assert npiv > 0
            v2, p2 = O.orthogonalize(v, pb, d, a, b, winv, N, Av, n, p)
            U = O.sparse_matrix_vector_product(M, Av, not right, n, p)      # S1 Av
            T2, Tp2 = O.orthogonalize(T, Tp, d, a, b, winv, Mc, U, n, p)     # the recurrence
            assert np.array_equal(T2, O.sparse_matrix_vector_product(M, v2, not right, n, p))
            assert np.array_equal(Tp2, O.sparse_matrix_vector_product(M, p2, not right, n, p))
            v, pb, T, Tp = v2, p2, T2, Tp2
        # linearity holds for ANY n x n factors and selection mask (rank-deficient iterations): random ones
        rng = np.random.default_rng(p % 1000)
        d = rng.integers(0, 2, size=n).astype(np.uint32)
        a, b, winv = (rng.integers(0, p, size=n * n).astype(np.uint32) for _ in range(3))
        Av = O.sparse_matrix_vector_product(M, T, right, n, p)
        v2, p2 = O.orthogonalize(v, pb, d, a, b, winv, N, Av, n, p)
        U = O.sparse_matrix_vector_product(M, Av, not right, n, p)
        T2, Tp2 = O.orthogonalize(T, Tp, d, a, b, winv, Mc, U, n, p)
        assert np.array_equal(T2, O.sparse_matrix_vector_product(M, v2, not right, n, p))
        assert np.array_equal(Tp2, O.sparse_matrix_vector_product(M, p2, not right, n, p))


def _worker_recurrence(rank, world, port, q):
    """The round-2 multi-GPU loop (context.cu enqueue_iteration_mg) on gloo ranks with the CPU oracle: dealt degree-sorted
    labels, Av exchanged right after product 2, dots + all-reduce, orthogonalize on the local rows, U = S1 Av, the tmp
    recurrence on the local rows, exchange of the new tmp -- the new v is never gathered."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import blk_lanczos_b200 as B
    from oracle.oracle import Oracle
    from test_relabel_cpu import deal_labels
    O = Oracle()
    p, n, right, iters = 2147483647, 4, False, 5
    M = B.synth.powerlaw_rows(901, 800, mean=7, seed=9, with_empty_rows=5).reduced(p)
    N, Mc = M.nrows, M.ncols
    old2new, new2old, n_off = deal_labels(M.i, N, world)              # the N dimension under dealt labels
    m_off = B.plan_shards(M.j, Mc, world)
    Mr = B.SparseCOO(M.nrows, M.ncols, old2new[M.i].astype(np.int32), M.j, M.x)

    def shard(lo, hi, by_rows):
        key = Mr.i if by_rows else Mr.j
        sel = (key >= lo) & (key < hi)
        return B.SparseCOO(Mr.nrows, Mr.ncols, Mr.i[sel], Mr.j[sel], Mr.x[sel])

    S2 = shard(n_off[rank], n_off[rank + 1], True)                    # my rows of Av = M tmp
    S1 = shard(m_off[rank], m_off[rank + 1], False)                   # my rows of tmp = M^T v
    n0, n1, m0, m1 = int(n_off[rank]) * n, int(n_off[rank + 1]) * n, int(m_off[rank]) * n, int(m_off[rank + 1]) * n

    def gather(local, off):                                           # blocks may differ in size: one broadcast per owner
        parts = []
        for r in range(world):
            t = torch.from_numpy(local.astype(np.int64)) if r == rank else torch.zeros(int(off[r + 1] - off[r]) * n, dtype=torch.int64)
            dist.broadcast(t, src=r)
            parts.append(t.numpy())
        return np.concatenate(parts).astype(np.uint32)

    v0 = O.start_block(N * n, p).reshape(N, n)[new2old].ravel()       # blk_set_state: host rows scattered to their labels
    v_loc, p_loc = v0[n0:n1].copy(), np.zeros(n1 - n0, np.uint32)
    # mg_prepare: tmp = S1 v (gathered), Tp = S1 p (local rows)
    tmp = gather(O.sparse_matrix_vector_product(S1, v0, True, n, p)[m0:m1], m_off)
    Tp = np.zeros(m1 - m0, np.uint32)
    for _ in range(iters):
        Av_loc = O.sparse_matrix_vector_product(S2, tmp, False, n, p)[n0:n1]
        Av = gather(Av_loc, n_off)                                    # pushed to the peers from inside the product
        a, b = O.block_dot_products((n1 - n0) // n, Av_loc, v_loc, n, p)
        sums = torch.from_numpy(np.concatenate([a, b]).astype(np.int64))
        dist.all_reduce(sums)
        a, b = (sums.numpy()[:n * n] % p).astype(np.uint32), (sums.numpy()[n * n:] % p).astype(np.uint32)
        npiv, winv, d = O.semi_inverse(a, n, p)
        v_loc, p_loc = O.orthogonalize(v_loc, p_loc, d, a, b, winv, (n1 - n0) // n, Av_loc, n, p)
        U = O.sparse_matrix_vector_product(S1, Av, True, n, p)[m0:m1]
        t_loc, Tp = O.orthogonalize(tmp[m0:m1], Tp, d, a, b, winv, (m1 - m0) // n, U, n, p)
        tmp = gather(t_loc, m_off)
    want = O.lanczos_run(M, n, p, right, stop_after=iters)
    lab = new2old[int(n_off[rank]):int(n_off[rank + 1])]             # host rows behind my labels
    ok = np.array_equal(v_loc.reshape(-1, n), want["v"][:N * n].reshape(N, n)[lab]) and \
        np.array_equal(p_loc.reshape(-1, n), want["p"][:N * n].reshape(N, n)[lab]) and \
        np.array_equal(tmp, O.sparse_matrix_vector_product(M, want["v"][:N * n], True, n, p))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_gloo_ranks_run_the_recurrence_loop_with_dealt_labels(lib, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_recurrence, args=(r, world, 29640 + world, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(r, True) for r in range(world)]
