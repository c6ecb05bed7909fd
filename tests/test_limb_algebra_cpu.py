"""CPU suite, part 5: the byte-limb algebra behind the tensor-core dense kernels (csrc/dense_mma.cu,
csrc/dense_umma.cu), restated in numpy and checked against the oracle.  The GPU tests prove the kernels;
this proves the formulas the kernels implement, including the overflow bounds quoted in their comments:

  block_dot_products (sequential/lanczos_modp.c:443-453)
      D[(i,a),(j,b)] = sum_r bytes(X)[r, 4i+a] * bytes(Av)[r, 4j+b]          (u8 x u8 -> s32, K = rows)
      C[i,j]         = sum_ab 2^(8(a+b)) D[(i,a),(j,b)]  mod p               flushed every 32 000 rows
  orthogonalize (:478-491)
      B_X[4i+a, 4j+b] = byte b of (2^(8a) X[i,j] mod p)
      D[r, 4j+b]      = sum_k bytes(V)[r, k] * B_X[k, 4j+b]                   (K = 4 n_pad bytes of a row)
      out[r, j]       = sum_b 2^(8b) D[r, 4j+b]  mod p
"""
import numpy as np
import pytest

P_FERMAT, P_CAP, P_MERSENNE = 65537, 1073741789, 2147483647


def row_bytes(X, n):
    """rows x n u32 -> rows x 4n u8: byte 4i+a of a row = limb a of column i (little endian: the memory image)"""
    return np.ascontiguousarray(X.reshape(-1, n).astype("<u4")).view(np.uint8).reshape(-1, 4 * n)


def test_memory_image_is_the_limb_layout():
    X = np.array([[0x04030201, 0xAABBCCDD]], dtype=np.uint32)
    assert row_bytes(X, 2).tolist() == [[1, 2, 3, 4, 0xDD, 0xCC, 0xBB, 0xAA]]


@pytest.mark.parametrize("p", [P_FERMAT, P_CAP, P_MERSENNE, 7])
@pytest.mark.parametrize("n", [8, 16])
def test_dot_products_from_byte_gemms(oracle, n, p):
    rng = np.random.default_rng(n + p % 1000)
    N = 70_000                                   # more than one 32 000-row epoch
    v = rng.integers(0, p, size=N * n).astype(np.uint32)
    Av = rng.integers(0, p, size=N * n).astype(np.uint32)
    v[::7] = p - 1; Av[::5] = p - 1
    vb, ab = row_bytes(v, n).astype(np.int64), row_bytes(Av, n).astype(np.int64)
    w = np.array([pow(2, 8 * a, p) for a in range(4)], dtype=object)
    got = []
    for X in (vb, ab):
        C = np.zeros((n, n), dtype=object)
        for lo in range(0, N, 32_000):          # the accumulators are flushed before they can overflow
            D = X[lo:lo + 32_000].T @ ab[lo:lo + 32_000]
            assert D.max() < 2 ** 31            # 32 000 * 255^2 < 2^31
            D4 = D.reshape(n, 4, n, 4).astype(object)
            for a in range(4):
                for b in range(4):
                    C = (C + D4[:, a, :, b] * int(w[a] * w[b] % p)) % p
        got.append(np.array(C.tolist(), dtype=np.uint32).ravel())
    want = oracle.block_dot_products(N, Av, v, n, p)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
    assert 32_000 * 255 * 255 < 2 ** 31 <= 33_100 * 255 * 255


@pytest.mark.parametrize("p", [P_FERMAT, P_CAP, P_MERSENNE])
def test_orthogonalize_from_byte_gemms(oracle, p):
    n = 16
    rng = np.random.default_rng(p % 977)
    N = 513
    v, Av, pb = (rng.integers(0, p, size=N * n).astype(np.uint32) for _ in range(3))
    v[::3] = p - 1; pb[::4] = p - 1
    U = rng.integers(0, p, size=(n, n)).astype(np.uint64)
    U = ((U + U.T) % p).astype(np.uint32)
    U[:, 5] = 0; U[5, :] = 0                                    # a zero pivot: d mixes 0 and 1
    npiv, winv, d = oracle.semi_inverse(U.ravel(), n, p)
    vtAv, vtAAv = (rng.integers(0, p, size=n * n).astype(np.uint32) for _ in range(2))
    want_v, want_p = oracle.orthogonalize(v, pb, d, vtAv, vtAAv, winv, N, Av, n, p)

    # the n x n coefficients exactly as the reference forms them (:460-475)
    W, A1, A2 = (np.array(x.reshape(n, n).tolist(), dtype=object) for x in (winv, vtAv, vtAAv))
    spliced = np.where(d[None, :] != 0, A2, A1)
    c = (-(W @ spliced)) % p
    vtAvd = np.where(d[None, :] != 0, (-A1) % p, 0)

    def limb_matrix(X):
        """[4i+a, 4j+b] = byte b of (2^(8a) X[i,j] mod p)"""
        B = np.zeros((4 * n, 4 * n), dtype=np.int64)
        for i in range(n):
            for a in range(4):
                for j in range(n):
                    x = int(X[i, j]) * pow(2, 8 * a, p) % p
                    for b in range(4):
                        B[4 * i + a, 4 * j + b] = (x >> (8 * b)) & 0xFF
        return B

    vb, pbb = row_bytes(v, n).astype(np.int64), row_bytes(pb, n).astype(np.int64)
    Dv = vb @ limb_matrix(c) + pbb @ limb_matrix(vtAvd)         # K = 128 bytes per output limb
    Dp = vb @ limb_matrix(W)
    assert Dv.max() <= 128 * 255 * 255 < 2 ** 23                # the bound recombine23() relies on (dense_umma.cu)
    shift = np.array([1, 1 << 8, 1 << 16, 1 << 24], dtype=object)

    def recombine(D):
        return (D.reshape(N, n, 4).astype(object) * shift).sum(axis=2) % p

    V2, A2d, P2 = (np.array(x.reshape(N, n).tolist(), dtype=object) for x in (v, Av, pb))
    out_v = (recombine(Dv) + np.where(d[None, :] != 0, A2d, V2)) % p
    out_p = (recombine(Dp) + np.where(d[None, :] != 0, 0, P2)) % p
    assert np.array_equal(np.array(out_v.tolist(), dtype=np.uint32).ravel(), want_v)
    assert np.array_equal(np.array(out_p.tolist(), dtype=np.uint32).ravel(), want_p)


def test_swizzle_64b_is_a_permutation_of_16_byte_chunks():
    """The XOR pattern the orthogonalize kernel uses to place its B tiles and to address tile rows: chunk' = chunk ^ ((row >> 1) & 3)
    inside a 64-byte row (TMA SWIZZLE_64B = Swizzle<2,4,3> on byte addresses: bits 4-5 ^= bits 7-8)."""
    for row in range(64):
        for chunk in range(4):
            addr = row * 64 + chunk * 16
            swz = addr ^ (((addr >> 7) & 3) << 4)
            assert swz == row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)
    # a quarter-warp (8 consecutive rows, same logical chunk) covers all 32 banks exactly once
    for base in range(0, 64, 8):
        for chunk in range(4):
            banks = set()
            for row in range(base, base + 8):
                a = row * 64 + ((chunk ^ ((row >> 1) & 3)) << 4)
                banks.update(range((a // 4) % 32, (a // 4) % 32 + 4))
            assert len(banks) == 32
