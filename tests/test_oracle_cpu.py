"""CPU suite, part 1: pin the oracle (oracle/lanczos_oracle.c) against the reference.

(a) the committed golden vectors (tests/golden/*.npz, produced by executing the unmodified
    reference -- see tests/golden/make_golden.py);
(b) where oracle/_ref/libref_seq.so exists (the build container, and the GPU box because the
    built files travel), directly against the reference's object code on fresh random inputs.
"""
import numpy as np
import pytest

from conftest import golden_cases, load_golden
from oracle.oracle import Reference, have_reference, block_pad

PRIMES = [65537, 1073741789, 2147483647, 7]
NS = [1, 2, 3, 4, 8, 16]


@pytest.mark.parametrize("name", golden_cases("cli_"))
def test_oracle_reproduces_reference_cli_kernel(oracle, name):
    z, M = load_golden(name)
    p, n, right = int(z["p"]), int(z["n"]), bool(z["right"])
    st = oracle.lanczos_run(M.reduced(p), n, p, right)
    N = M.ncols if right else M.nrows
    assert st["stopped"] and st["iters"] == int(z["iters"])
    assert np.array_equal(st["v"][:N * n].reshape(N, n), z["kernel"])
    # the reference's own checker verdict on the golden file agrees with the kernel property
    # (it is KO for the small-prime case where the iteration broke down by chance)
    Mp = M.reduced(p)
    in_kernel = st["v"].any() and not oracle.sparse_matrix_vector_product(Mp, st["v"], not right, n, p).any()
    assert bool(z["checker_ok"]) == bool(in_kernel)
    assert bool(z["ok_vtM"]) == (not st["tmp"][:(M.nrows if right else M.ncols) * n].any())


@pytest.mark.parametrize("name", golden_cases("loop_"))
def test_oracle_reproduces_reference_loop_state(oracle, name):
    z, M = load_golden(name)
    p, n, right, K = int(z["p"]), int(z["n"]), bool(z["right"]), int(z["K"])
    N = M.ncols if right else M.nrows
    assert np.array_equal(oracle.start_block(N * n, p), z["v0"][:N * n])
    st = oracle.lanczos_run(M.reduced(p), n, p, right, stop_after=K)
    assert st["iters"] == K and not st["stopped"]
    for k, g in (("v", "v"), ("tmp", "tmp"), ("Av", "Av"), ("p", "pblk")):
        assert np.array_equal(st[k], z[g]), k
    Mp = M.reduced(p)
    # n x n matrices of the last iteration: redo it from the state before
    st2 = oracle.lanczos_run(Mp, n, p, right, stop_after=K - 1) if K > 1 else None
    v = st2["v"] if st2 else z["v0"]
    tmp = oracle.sparse_matrix_vector_product(Mp, v, not right, n, p)
    Av = oracle.sparse_matrix_vector_product(Mp, tmp, right, n, p)
    a, b = oracle.block_dot_products(N, Av, v, n, p)
    npiv, winv, d = oracle.semi_inverse(a, n, p)
    assert np.array_equal(a, z["vtAv"]) and np.array_equal(b, z["vtAAv"])
    assert npiv == int(z["npiv"]) and np.array_equal(winv, z["winv"]) and np.array_equal(d, z["d"])


needs_ref = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (no /root/reference here)")


@needs_ref
@pytest.mark.parametrize("p", PRIMES)
def test_oracle_vs_reference_functions(oracle, p):
    import blk_lanczos_b200 as B
    R = Reference()
    rng = np.random.default_rng(p)
    for n in NS:
        M = B.synth.powerlaw_rows(300, 280, mean=6, seed=n, with_empty_rows=5, order="file").reduced(p)
        for tr in (False, True):
            cols = M.nrows if tr else M.ncols
            x = rng.integers(0, p, size=cols * n).astype(np.uint32)
            assert np.array_equal(oracle.sparse_matrix_vector_product(M, x, tr, n, p),
                                  R.sparse_matrix_vector_product(M, x, tr, n, p))
        N = 301
        v, Av, pb = (rng.integers(0, p, size=N * n).astype(np.uint32) for _ in range(3))
        a, b = oracle.block_dot_products(N, Av, v, n, p), R.block_dot_products(N, Av, v, n, p)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
        for trial in range(12):
            U = rng.integers(0, p, size=(n, n)).astype(np.uint64)
            U = ((U + U.T) % p).astype(np.uint32)
            if trial % 3 == 0 and n > 1:
                U[:, n // 2] = 0; U[n // 2, :] = 0          # rank deficient
            if trial % 5 == 0:
                U[0, 0] = 0                                 # forces a row swap
            if trial == 11:
                U[:] = 0                                    # no pivot at all -> returns 0
            so, sr = oracle.semi_inverse(U.ravel(), n, p), R.semi_inverse(U.ravel(), n, p)
            assert so[0] == sr[0] and np.array_equal(so[1], sr[1]) and np.array_equal(so[2], sr[2])
            vt, vtt = (rng.integers(0, p, size=n * n).astype(np.uint32) for _ in range(2))
            oo = oracle.orthogonalize(v, pb, so[2], vt, vtt, so[1], N, Av, n, p)
            rr = R.orthogonalize(v, pb, so[2], vt, vtt, so[1], N, Av, n, p)
            assert np.array_equal(oo[0], rr[0]) and np.array_equal(oo[1], rr[1])
    assert np.array_equal(oracle.start_block(2000, p), R.start_block(2000, p))
    for a in (1, 2, 3, p - 1, 12345 % p or 1):
        assert oracle.L.orc_invmod(a, p) == R.invmod(a, p)


def test_block_pad_matches_reference_formula():
    # sequential/lanczos_modp.c:594-597
    assert block_pad(20000, 19000, 1, False) == 20000
    assert block_pad(10, 7, 4, False) == 12 * 4
    assert block_pad(10, 17, 4, False) == 20 * 4
    assert block_pad(10, 17, 4, True) == 20 * 4
