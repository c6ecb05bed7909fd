"""CPU suite, part 4: the P x Q block grid (the reference's 2-D
decomposition, mpi/lanczos_modp.c:532-547, 590-620).  blk_plan_grid is host logic of the product;
the iteration on the grid is executed here with the CPU oracle on every block -- first with
"virtual ranks" in one process for many grid shapes, then with four gloo ranks and real
all_gather / all_to_all collectives inside row and column groups -- and must reproduce the
unsharded run bit for bit.  This is the specification the device-side grid is built against:

    rank (a, b) stores block (a, b) of the operator and owns piece (a, b) of v, Av, p and piece (b, a) of tmp
    1. all-gather v_a inside grid row a            2. partial tmp_b = S1[b, a] v_a
    3. reduce-scatter mod p inside grid column b   4. all-gather tmp_b inside grid column b
    5. partial Av_a = S2[a, b] tmp_b               6. reduce-scatter mod p inside grid row a
    7. dots on the owned rows, all-reduce of the 2 n^2 u64 sums, semi_inverse, orthogonalize on the owned rows
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_grid_plan_properties(lib):
    M = lib.synth.powerlaw_rows(5000, 4000, mean=12, seed=1, with_empty_rows=20)
    shapes = {1: (1, 1), 2: (2, 1), 4: (2, 2), 6: (3, 2), 8: (4, 2), 7: (7, 1), 16: (4, 4)}      # MPI_Dims_create
    for right in (False, True):
        N, Mc = (M.ncols, M.nrows) if right else (M.nrows, M.ncols)
        for world, (P, Q) in shapes.items():
            g = lib.plan_grid(M, right, world)
            assert (g["P"], g["Q"]) == (P, Q)
            assert g["n_off"][0] == 0 and g["n_off"][-1] == N and np.all(np.diff(g["n_off"]) >= 0)
            assert g["m_off"][0] == 0 and g["m_off"][-1] == Mc and np.all(np.diff(g["m_off"]) >= 0)
            assert g["block_nnz"].sum() == M.nnz
            # every block carries about 1/world of the non-zeros (uniform columns, weight-balanced rows)
            assert g["block_nnz"].max() <= 1.35 * M.nnz / world + 50
            for a in range(P):         # owned pieces tile their block, equal rows
                assert g["n_sub"][a][0] == g["n_off"][a] and g["n_sub"][a][-1] == g["n_off"][a + 1]
                assert np.all(np.diff(g["n_sub"][a]) >= 0) and np.ptp(np.diff(g["n_sub"][a])) <= 1
            for b in range(Q):
                assert g["m_sub"][b][0] == g["m_off"][b] and g["m_sub"][b][-1] == g["m_off"][b + 1]
                assert np.ptp(np.diff(g["m_sub"][b])) <= 1
    g = lib.plan_grid(M, False, 8, (2, 4))
    assert (g["P"], g["Q"]) == (2, 4) and g["n_sub"].shape == (2, 5) and g["m_sub"].shape == (4, 3)
    with pytest.raises(lib.BlkError):
        lib.plan_grid(M, False, 8, (3, 2))
    # exchange volume per rank and iteration, in rows: the figure quoted in include/blk_lanczos.h
    g = lib.plan_grid(M, False, 8)
    P, Q = g["P"], g["Q"]
    grid_rows = 2 * (M.nrows / P * (Q - 1) / Q + M.ncols / Q * (P - 1) / P)
    one_d_rows = (8 - 1) / 8 * (M.nrows + M.ncols)
    assert grid_rows < 0.6 * one_d_rows


def grid_blocks(B, M, right, g):
    """blocks[a][b]: block (a, b) in M's own orientation with block-local indices."""
    iN, iM = (M.j, M.i) if right else (M.i, M.j)
    P, Q = g["P"], g["Q"]
    a_of = np.searchsorted(g["n_off"], iN, side="right") - 1
    b_of = np.searchsorted(g["m_off"], iM, side="right") - 1
    out = [[None] * Q for _ in range(P)]
    for a in range(P):
        for b in range(Q):
            sel = (a_of == a) & (b_of == b)
            ln, lm = (iN[sel] - g["n_off"][a]).astype(np.int32), (iM[sel] - g["m_off"][b]).astype(np.int32)
            sn, sm = int(g["n_off"][a + 1] - g["n_off"][a]), int(g["m_off"][b + 1] - g["m_off"][b])
            out[a][b] = B.SparseCOO(sm, sn, lm, ln, M.x[sel]) if right else B.SparseCOO(sn, sm, ln, lm, M.x[sel])
    return out


def addmod(parts, p):
    s = np.zeros_like(parts[0], dtype=np.uint64)
    for x in parts:
        s += x.astype(np.uint64)
    return (s % np.uint64(p)).astype(np.uint32)


@pytest.mark.parametrize("right", [False, True])
@pytest.mark.parametrize("grid", [(1, 1), (2, 1), (1, 2), (2, 2), (4, 2), (2, 3)])
def test_grid_iteration_virtual_ranks(lib, oracle, grid, right):
    p, n, K = 2147483647, 4, 4
    M = lib.synth.powerlaw_rows(700, 820, mean=6, seed=3, with_empty_rows=6).reduced(p)
    N, Mc = (M.ncols, M.nrows) if right else (M.nrows, M.ncols)
    P, Q = grid
    g = lib.plan_grid(M, right, P * Q, grid)
    blk = grid_blocks(lib, M, right, g)
    ns, ms = g["n_sub"], g["m_sub"]
    v = oracle.start_block(N * n, p)
    # owned pieces, indexed [a][b]
    own = lambda full, a, b: full[ns[a][b] * n:ns[a][b + 1] * n].copy()
    V = [[own(v, a, b) for b in range(Q)] for a in range(P)]
    Pb = [[np.zeros_like(V[a][b]) for b in range(Q)] for a in range(P)]
    for _ in range(K):
        v_a = [np.concatenate(V[a]) for a in range(P)]                                        # 1
        part = [[oracle.sparse_matrix_vector_product(blk[a][b], v_a[a], not right, n, p) for b in range(Q)] for a in range(P)]   # 2
        T = [[None] * P for _ in range(Q)]
        for b in range(Q):                                                                     # 3
            for a in range(P):
                lo, hi = (ms[b][a] - g["m_off"][b]) * n, (ms[b][a + 1] - g["m_off"][b]) * n
                T[b][a] = addmod([part[a2][b][lo:hi] for a2 in range(P)], p)
        t_b = [np.concatenate(T[b]) for b in range(Q)]                                         # 4
        part = [[oracle.sparse_matrix_vector_product(blk[a][b], t_b[b], right, n, p) for b in range(Q)] for a in range(P)]       # 5
        Av = [[None] * Q for _ in range(P)]
        for a in range(P):                                                                     # 6
            for b in range(Q):
                lo, hi = (ns[a][b] - g["n_off"][a]) * n, (ns[a][b + 1] - g["n_off"][a]) * n
                Av[a][b] = addmod([part[a][b2][lo:hi] for b2 in range(Q)], p)
        sums = np.zeros(2 * n * n, dtype=np.uint64)                                            # 7
        for a in range(P):
            for b in range(Q):
                x, y = oracle.block_dot_products(int(ns[a][b + 1] - ns[a][b]), Av[a][b], V[a][b], n, p)
                sums += np.concatenate([x, y]).astype(np.uint64)
        vtAv, vtAAv = ((sums[:n * n] % p).astype(np.uint32), (sums[n * n:] % p).astype(np.uint32))
        npiv, winv, d = oracle.semi_inverse(vtAv, n, p)
        assert npiv > 0
        for a in range(P):
            for b in range(Q):
                V[a][b], Pb[a][b] = oracle.orthogonalize(V[a][b], Pb[a][b], d, vtAv, vtAAv, winv,
                                                         int(ns[a][b + 1] - ns[a][b]), Av[a][b], n, p)
    want = oracle.lanczos_run(M, n, p, right, stop_after=K)
    assert np.array_equal(np.concatenate([np.concatenate(r) for r in V]), want["v"][:N * n])
    assert np.array_equal(np.concatenate([np.concatenate(r) for r in Pb]), want["p"][:N * n])
    assert np.array_equal(np.concatenate([np.concatenate(r) for r in Av]), want["Av"][:N * n])


def _grid_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import blk_lanczos_b200 as B
    from oracle.oracle import Oracle
    O = Oracle()
    p, n, right, K = 1073741789, 3, True, 3
    M = B.synth.powerlaw_rows(640, 510, mean=7, seed=12, with_empty_rows=4).reduced(p)
    N, Mc = (M.ncols, M.nrows) if right else (M.nrows, M.ncols)
    g = B.plan_grid(M, right, world)
    P, Q = g["P"], g["Q"]
    a, b = rank // Q, rank % Q
    # every rank creates every group, in the same order (torch.distributed requirement)
    row_groups = [dist.new_group([a2 * Q + b2 for b2 in range(Q)]) for a2 in range(P)]
    col_groups = [dist.new_group([a2 * Q + b2 for a2 in range(P)]) for b2 in range(Q)]
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_grid_cpu import grid_blocks
    blk = grid_blocks(B, M, right, g)[a][b]
    ns, ms = g["n_sub"], g["m_sub"]

    def t64(x):
        return torch.from_numpy(np.ascontiguousarray(x).astype(np.int64))

    def all_gather(local, bounds, group, members):
        """ragged all-gather inside a group: one broadcast per member (sizes differ by at most one row)"""
        parts = []
        for k, src in enumerate(members):
            buf = t64(local) if src == rank else torch.zeros(int(bounds[k + 1] - bounds[k]) * n, dtype=torch.int64)
            dist.broadcast(buf, src=src, group=group)
            parts.append(buf.numpy())
        return np.concatenate(parts).astype(np.uint32)

    def reduce_scatter(partial, bounds, base, group, members, me):
        """all-to-all of the pieces + local sum mod p (NCCL's u32 sum would overflow at p > 2^30)"""
        got = []
        for k, dst in enumerate(members):
            piece = t64(partial[(bounds[k] - base) * n:(bounds[k + 1] - base) * n])
            bufs = [torch.zeros_like(piece) for _ in members] if dst == rank else None
            dist.gather(piece, bufs, dst=dst, group=group)
            if dst == rank:
                got = [x.numpy() for x in bufs]
        s = np.zeros_like(got[0])
        for x in got:
            s += x
        return (s % p).astype(np.uint32)

    row_members, col_members = [a * Q + b2 for b2 in range(Q)], [a2 * Q + b for a2 in range(P)]
    v_full = O.start_block(N * n, p)
    V = v_full[ns[a][b] * n:ns[a][b + 1] * n].copy()
    Pb = np.zeros_like(V)
    rows_own = int(ns[a][b + 1] - ns[a][b])
    for _ in range(K):
        v_a = all_gather(V, ns[a], row_groups[a], row_members)
        part = O.sparse_matrix_vector_product(blk, v_a, not right, n, p)
        T = reduce_scatter(part, ms[b], g["m_off"][b], col_groups[b], col_members, a)
        t_b = all_gather(T, ms[b], col_groups[b], col_members)
        part = O.sparse_matrix_vector_product(blk, t_b, right, n, p)
        Av = reduce_scatter(part, ns[a], g["n_off"][a], row_groups[a], row_members, b)
        x, y = O.block_dot_products(rows_own, Av, V, n, p)
        sums = t64(np.concatenate([x, y]))
        dist.all_reduce(sums)
        vtAv, vtAAv = (sums.numpy()[:n * n] % p).astype(np.uint32), (sums.numpy()[n * n:] % p).astype(np.uint32)
        npiv, winv, d = O.semi_inverse(vtAv, n, p)
        V, Pb = O.orthogonalize(V, Pb, d, vtAv, vtAAv, winv, rows_own, Av, n, p)
    want = O.lanczos_run(M, n, p, right, stop_after=K)
    lo, hi = ns[a][b] * n, ns[a][b + 1] * n
    ok = np.array_equal(V, want["v"][lo:hi]) and np.array_equal(Pb, want["p"][lo:hi]) and np.array_equal(Av, want["Av"][lo:hi])
    q.put((rank, bool(ok), (P, Q)))
    dist.destroy_process_group()


def test_four_rank_gloo_grid_reproduces_sequential(lib):
    world, port = 4, 29641
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_grid_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = sorted(q.get(timeout=240) for _ in range(world))
    for pr in procs:
        pr.join(timeout=60)
    assert res == [(r, True, (2, 2)) for r in range(world)]
