"""Synthetic sparse matrices of the BASELINE.json shapes, and MatrixMarket text I/O.

The reference ships no matrices (they were downloaded by project.py); these generators
follow SURVEY.md section 8(d): deterministic (numpy default_rng(seed)), positive entries in
[1,100) so the loader quirk on negative values (sequential/lanczos_modp.c:238-243, F9) never
triggers, coordinate integer general.

The COO container mirrors struct sparsematrix_t (sequential/lanczos_modp.c:55-62):
0-based int32 row/column indices and u32 values already reduced mod p by the caller.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class SparseCOO:
    nrows: int
    ncols: int
    i: np.ndarray      # int32 [nnz]
    j: np.ndarray      # int32 [nnz]
    x: np.ndarray      # uint32 [nnz]

    @property
    def nnz(self) -> int:
        return int(self.i.shape[0])

    def reduced(self, p: int) -> "SparseCOO":
        """Values mod p, as sparsematrix_mm_load stores them (sequential/lanczos_modp.c:243)."""
        return SparseCOO(self.nrows, self.ncols, self.i, self.j,
                         (self.x.astype(np.uint64) % np.uint64(p)).astype(np.uint32))


def _finish(nrows, ncols, i, j, x, order: str) -> SparseCOO:
    if order == "row":
        k = np.lexsort((j, i))
    elif order == "col":
        k = np.lexsort((i, j))
    elif order == "file":          # keep generation order (unsorted)
        k = slice(None)
    else:
        raise ValueError(order)
    return SparseCOO(nrows, ncols, np.ascontiguousarray(i[k], dtype=np.int32),
                     np.ascontiguousarray(j[k], dtype=np.int32),
                     np.ascontiguousarray(x[k], dtype=np.uint32))


def uniform_rows(nrows: int, ncols: int, per_row: int, seed: int = 0, order: str = "row",
                 vmax: int = 100) -> SparseCOO:
    """`per_row` entries in every row, columns uniform (duplicates allowed: COO semantics add)."""
    rng = np.random.default_rng(seed)
    i = np.repeat(np.arange(nrows, dtype=np.int64), per_row)
    j = rng.integers(0, ncols, size=i.shape[0], dtype=np.int64)
    x = rng.integers(1, vmax, size=i.shape[0], dtype=np.int64)
    return _finish(nrows, ncols, i, j, x, order)


def uniform_nnz(nrows: int, ncols: int, nnz: int, seed: int = 0, order: str = "col",
                vmax: int = 100) -> SparseCOO:
    """`nnz` entries at uniform positions (TF17/TF18-shaped configs; column-sorted like SuiteSparse)."""
    rng = np.random.default_rng(seed)
    i = rng.integers(0, nrows, size=nnz, dtype=np.int64)
    j = rng.integers(0, ncols, size=nnz, dtype=np.int64)
    x = rng.integers(1, vmax, size=nnz, dtype=np.int64)
    return _finish(nrows, ncols, i, j, x, order)


def powerlaw_degrees(nrows: int, mean: float, rng, cap: int = 1_000_000, dmin: int | None = None):
    """Pareto(shape 1.5) row degrees with the requested mean (before capping)."""
    shape = 1.5
    dmin = dmin if dmin is not None else max(1, int(round(mean * (shape - 1) / shape)))
    u = rng.random(nrows)
    d = np.floor(dmin * (1.0 - u) ** (-1.0 / shape)).astype(np.int64)
    return np.minimum(d, cap)


def powerlaw_rows(nrows: int, ncols: int, mean: float = 30.0, seed: int = 0, cap: int = 1_000_000,
                  order: str = "row", vmax: int = 100, with_empty_rows: int = 0) -> SparseCOO:
    """Config-4 shape: heavy-tailed row degrees, uniform columns.  `with_empty_rows` forces that
    many randomly chosen rows to have no entry (edge case for the layout builder)."""
    rng = np.random.default_rng(seed)
    d = powerlaw_degrees(nrows, mean, rng, cap=min(cap, 8 * ncols))
    if with_empty_rows:
        d[rng.choice(nrows, size=with_empty_rows, replace=False)] = 0
    i = np.repeat(np.arange(nrows, dtype=np.int64), d)
    j = rng.integers(0, ncols, size=i.shape[0], dtype=np.int64)
    x = rng.integers(1, vmax, size=i.shape[0], dtype=np.int64)
    return _finish(nrows, ncols, i, j, x, order)


# BASELINE.json configs 1-3 (config 4 is generated on the device by bench.py).
def baseline_config(k: int, seed: int | None = None) -> tuple[SparseCOO, dict]:
    if k == 1:
        M = uniform_rows(20_000, 19_000, 30, seed=11 if seed is None else seed)
        return M, dict(p=65537, n=1, right=False)
    if k == 2:
        M = uniform_nnz(38_132, 48_630, 586_218, seed=17 if seed is None else seed, order="col")
        return M, dict(p=65537, n=4, right=True)
    if k == 3:
        M = uniform_nnz(95_368, 123_412, 1_601_580, seed=18 if seed is None else seed, order="col")
        return M, dict(p=2147483647, n=8, right=True)
    raise ValueError(k)


def write_mtx(path: str, M: SparseCOO) -> None:
    """MatrixMarket coordinate integer general, 1-based, one `i j x` triplet per line -- the only
    flavour sparsematrix_mm_load accepts (sequential/lanczos_modp.c:216-221)."""
    with open(path, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n")
        f.write(f"{M.nrows} {M.ncols} {M.nnz}\n")
        body = np.empty((M.nnz, 3), dtype=np.int64)
        body[:, 0] = M.i.astype(np.int64) + 1
        body[:, 1] = M.j.astype(np.int64) + 1
        body[:, 2] = M.x
        np.savetxt(f, body, fmt="%d")


def read_mtx(path: str) -> SparseCOO:
    with open(path) as f:
        line = f.readline()
        assert line.startswith("%%MatrixMarket"), line
        while True:
            line = f.readline()
            if not line.startswith("%"):
                break
        nrows, ncols, nnz = (int(t) for t in line.split())
        body = np.loadtxt(f, dtype=np.int64, ndmin=2) if nnz else np.zeros((0, 3), dtype=np.int64)
    assert body.shape[0] == nnz
    return SparseCOO(nrows, ncols, (body[:, 0] - 1).astype(np.int32), (body[:, 1] - 1).astype(np.int32),
                     (body[:, 2] & 0xFFFFFFFF).astype(np.uint32))


def read_kernel_block(path: str) -> np.ndarray:
    """Read a kernel file written by save_vector_block (sequential/lanczos_modp.c:673-686):
    dense array, column-major text.  Returns row-major [N, n] uint32."""
    with open(path) as f:
        head = f.readline()
        assert head.startswith("%%MatrixMarket matrix array integer general"), head
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        N, n = (int(t) for t in line.split())
        vals = np.loadtxt(f, dtype=np.int64, ndmin=1)
    return vals.reshape(n, N).T.astype(np.uint32).copy()


def reference_start_block(count: int, p: int) -> np.ndarray:
    """The reference's start block: v[i] = random64() % prime for i row-major over N*n, xoshiro256+
    with its fixed seed (sequential/lanczos_modp.c:64-87, 624-625).  Host-side mirror of what the C
    driver does (driver/lanczos_modp_gpu.c); pure Python integers, ~1 us per value."""
    mask = (1 << 64) - 1
    s0, s1, s2, s3 = 0x1415926535, 0x8979323846, 0x2643383279, 0x5028841971
    out = np.empty(count, dtype=np.uint32)
    for t in range(count):
        x = (s0 + s3) & mask
        r = ((((x << 23) | (x >> 41)) & mask) + s0) & mask
        sh = (s1 << 17) & mask
        s2 ^= s0
        s3 ^= s1
        s1 ^= s2
        s0 ^= s3
        s2 ^= sh
        s3 = ((s3 << 45) | (s3 >> 19)) & mask
        out[t] = r % p
    return out
