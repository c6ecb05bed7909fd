"""CPU suite: the degree-sorted relabelling of the Lanczos dimension, restated in numpy and checked with the CPU oracle.

csrc/layout_build.cu `degree_sort_maps` sorts the rows of v / Av / p by decreasing number of entries and, with several
GPUs, deals the sorted sequence round-robin to the ranks' blocks (`k_deal_labels`: sorted position s goes to block
s mod W, place s div W).  The claims this file pins without a GPU:

  * the map is a bijection, the blocks are the equal split the library announces (off[w] = w*(N div W) + min(w, N mod W)),
    every block gets the same number of rows and -- to within one giant row -- the same number of non-zeros, and the first
    `per` labels of every block are its share of the globally hottest rows (what the HOT bit of the column word marks);
  * running the loop on the relabelled matrix with the relabelled start block gives, after undoing the labels, exactly the
    blocks of the unrelabelled run: dots are order-free and orthogonalize is row-wise, so labels are invisible
    (sequential/lanczos_modp.c:443-491).  On the GPU the same statement is tests/test_gpu_parity.py
    ::test_relabelled_layout_is_invisible and the "hot" modes of tests/test_gpu_multi.py.
"""
import numpy as np
import pytest


def deal_labels(idx, dim, world):
    """numpy model of degree_sort_maps: returns (old2new, new2old, block offsets)."""
    cnt = np.bincount(idx, minlength=dim)
    sorted2old = np.argsort(-cnt, kind="stable")            # radix sort of ~cnt is stable: ties keep index order
    s = np.arange(dim)
    if world == 1:
        new2old = sorted2old.copy()
        off = np.array([0, dim])
    else:
        w = s % world
        off_w = w * (dim // world) + np.minimum(w, dim % world)
        lab = off_w + s // world
        new2old = np.empty(dim, np.int64)
        new2old[lab] = sorted2old
        ws = np.arange(world + 1)
        off = np.where(ws == world, dim, ws * (dim // world) + np.minimum(ws, dim % world))
    old2new = np.empty(dim, np.int64)
    old2new[new2old] = s
    return old2new, new2old, off


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_dealt_labels_are_a_balanced_bijection(lib, world):
    M = lib.synth.powerlaw_rows(5003, 4000, mean=12, seed=1, with_empty_rows=20)
    N = M.nrows
    old2new, new2old, off = deal_labels(M.i, N, world)
    assert sorted(old2new.tolist()) == list(range(N)) and np.array_equal(new2old[old2new], np.arange(N))
    sizes = np.diff(off)
    assert sizes.sum() == N and sizes.max() - sizes.min() <= 1
    cnt = np.bincount(M.i, minlength=N)
    per_block = np.array([cnt[new2old[off[w]:off[w + 1]]].sum() for w in range(world)])
    assert per_block.max() - per_block.min() <= cnt.max()              # equal non-zeros to within one giant row
    # inside a block labels are by decreasing degree, and the first `per` labels of all blocks together are the
    # world * per hottest rows of the whole dimension
    for w in range(world):
        assert np.all(np.diff(cnt[new2old[off[w]:off[w + 1]]]) <= 0)
    per = 16
    hot = np.concatenate([new2old[off[w]:off[w] + per] for w in range(world)])
    kth = np.sort(cnt)[::-1][world * per - 1]
    assert cnt[hot].min() >= kth


@pytest.mark.parametrize("world,right", [(1, False), (4, False), (3, True)])
def test_relabelled_run_is_the_same_run(lib, oracle, world, right):
    p, n = 2147483647, 4
    M = lib.synth.powerlaw_rows(700, 640, mean=6, seed=7, with_empty_rows=5).reduced(p)
    N = M.ncols if right else M.nrows
    idxN = M.j if right else M.i
    old2new, new2old, _ = deal_labels(idxN, N, world)
    relabelled = old2new[idxN].astype(np.int32)
    Mr = lib.SparseCOO(M.nrows, M.ncols, M.i if right else relabelled, relabelled if right else M.j, M.x)

    def to_new(block):                                       # rows of an N x n block under the new labels
        out = np.array(block, copy=True)
        rows = block[:N * n].reshape(N, n)
        out[:N * n] = rows[new2old].ravel()
        return out

    want = oracle.lanczos_run(M, n, p, right, stop_after=7)
    start = {k: np.zeros_like(want[k]) for k in ("v", "tmp", "Av", "p")}
    start["v"][:N * n] = oracle.start_block(N * n, p)
    start["v"] = to_new(start["v"])
    start["iters"] = 0
    got = oracle.lanczos_run(Mr, n, p, right, stop_after=7, state=start)
    assert got["iters"] == want["iters"] == 7
    for k in ("v", "Av", "p"):
        assert np.array_equal(got[k], to_new(want[k])), k
    # tmp lives in the other dimension (its labels are untouched) -- but the reference leaves the new v in its first N rows
    # (:652-656), so compare the part that is the product: rows [N, Mc) when Mc > N
    Mc = M.nrows if right else M.ncols
    if Mc > N:
        assert np.array_equal(got["tmp"][N * n:Mc * n], want["tmp"][N * n:Mc * n])
