"""CPU suite, part 5: the driver's text I/O (driver/mtx_io.c) -- no GPU involved.

The loader must accept what the reference's loader accepts and produce the same COO arrays,
including its quirk on negative entries (fscanf("%d") into a u32, then % prime:
sequential/lanczos_modp.c:238-243); the kernel-block and checkpoint writers must
produce the reference's bytes (save_vector_block :673-686, save_vectors openMP/…:573-589)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "block-lanczos-algorithm-parallelization_b200", "driver")

HARNESS = r'''
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mtx_io.h"
int main(int argc, char **argv) {
    if (!strcmp(argv[1], "load")) {                    /* load <mtx> <prime> <out.bin> */
        struct coo_matrix M; mtx_load(&M, argv[2], strtoull(argv[3], NULL, 10));
        FILE *f = fopen(argv[4], "wb");
        fwrite(&M.nrows, 4, 1, f); fwrite(&M.ncols, 4, 1, f); fwrite(&M.nnz, 8, 1, f);
        fwrite(M.i, 4, M.nnz, f); fwrite(M.j, 4, M.nnz, f); fwrite(M.x, 4, M.nnz, f); fclose(f);
        mtx_free(&M);
    } else if (!strcmp(argv[1], "kernel")) {           /* kernel <in.bin> <nrows> <n> <out> */
        int nrows = atoi(argv[3]), n = atoi(argv[4]);
        uint32_t *v = malloc(4ul * nrows * n); FILE *f = fopen(argv[2], "rb");
        if (fread(v, 4, (size_t)nrows * n, f) != (size_t)nrows * n) return 2; fclose(f);
        kernel_block_save(argv[5], nrows, n, v);
    } else {                                           /* vec <in.bin> <count> <out>, then reload and compare */
        long cnt = atol(argv[3]); uint32_t *v = malloc(4ul * cnt), *w = malloc(4ul * cnt);
        FILE *f = fopen(argv[2], "rb"); if (fread(v, 4, cnt, f) != (size_t)cnt) return 2; fclose(f);
        vector_save(argv[4], cnt, v); vector_load(argv[4], cnt, w);
        return memcmp(v, w, 4ul * cnt) ? 3 : 0;
    }
    return 0;
}
'''


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("mtxio")
    src = d / "h.c"
    src.write_text(HARNESS)
    exe = d / "h"
    subprocess.check_call(["gcc", "-std=c99", "-O2", "-Wall", "-I", DRV, str(src), os.path.join(DRV, "mtx_io.c"), "-o", str(exe)])
    return str(exe)


def _load(harness, mtx, p, tmp_path):
    out = str(tmp_path / "m.bin")
    r = subprocess.run([harness, "load", mtx, str(p), out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    raw = open(out, "rb").read()
    nrows, ncols = np.frombuffer(raw[:8], np.int32)
    nnz = int(np.frombuffer(raw[8:16], np.int64)[0])
    body = np.frombuffer(raw[16:], np.uint32)
    return nrows, ncols, nnz, body[:nnz].view(np.int32), body[nnz:2 * nnz].view(np.int32), body[2 * nnz:]


QUIRKY = """%%MatrixMarket matrix coordinate integer general
% a comment
%another
5 4 7
1 1 3
2 4 -1
5 2   100000
3 3 -65538
4 1 2147483647
 2 2 7
5 4 0
"""


@pytest.mark.parametrize("p", [65537, 2147483647, 7])
def test_loader_matches_reference_loader(harness, tmp_path, p):
    mtx = str(tmp_path / "q.mtx")
    open(mtx, "w").write(QUIRKY)
    nrows, ncols, nnz, i, j, x = _load(harness, mtx, p, tmp_path)
    assert (nrows, ncols, nnz) == (5, 4, 7)
    assert list(i) == [0, 1, 4, 2, 3, 1, 4] and list(j) == [0, 3, 1, 2, 0, 1, 3]
    vals = [3, -1, 100000, -65538, 2147483647, 7, 0]
    assert list(x) == [((v + (1 << 32)) % (1 << 32)) % p for v in vals]          # -k -> (2^32 - k) % p
    ref = os.path.join(ROOT, "oracle", "_ref", "libref_seq.so")
    if os.path.exists(ref):
        L = C.CDLL(ref)

        class Mat(C.Structure):
            _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("nnz", C.c_long), ("i", C.POINTER(C.c_int)),
                        ("j", C.POINTER(C.c_int)), ("x", C.POINTER(C.c_uint32))]
        C.c_uint64.in_dll(L, "prime").value = p
        name = C.c_char_p(mtx.encode())
        C.c_char_p.in_dll(L, "matrix_filename").value = mtx.encode()
        m = Mat()
        L.sparsematrix_mm_load(C.byref(m), name)
        assert (m.nrows, m.ncols, m.nnz) == (5, 4, 7)
        assert [m.i[k] for k in range(7)] == list(i) and [m.j[k] for k in range(7)] == list(j)
        assert [m.x[k] for k in range(7)] == list(x)


def test_loader_rejects_what_the_reference_rejects(harness, tmp_path):
    for head in ("%%MatrixMarket matrix array integer general\n3 3\n",
                 "%%MatrixMarket matrix coordinate real general\n3 3 1\n1 1 0.5\n",
                 "%%MatrixMarket matrix coordinate integer symmetric\n3 3 1\n1 1 1\n", "garbage\n"):
        mtx = str(tmp_path / "bad.mtx")
        open(mtx, "w").write(head)
        r = subprocess.run([harness, "load", mtx, "65537", str(tmp_path / "o.bin")], capture_output=True, text=True)
        assert r.returncode == 1, head
    mtx = str(tmp_path / "short.mtx")
    open(mtx, "w").write("%%MatrixMarket matrix coordinate integer general\n3 3 2\n1 1 1\n")
    r = subprocess.run([harness, "load", mtx, "65537", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 1 and "parse error entry 1" in r.stderr


def test_kernel_block_and_checkpoint_vectors_format(harness, tmp_path):
    rng = np.random.default_rng(3)
    nrows, n = 37, 3
    v = rng.integers(0, 2 ** 31 - 1, size=nrows * n).astype(np.uint32)
    v[5] = 0
    v[6] = 2 ** 31 - 2
    raw = str(tmp_path / "v.bin")
    v.tofile(raw)
    out = str(tmp_path / "k.mtx")
    assert subprocess.run([harness, "kernel", raw, str(nrows), str(n), out], capture_output=True).returncode == 0
    want = "%%MatrixMarket matrix array integer general\n%block of left-kernel vector computed by lanczos_modp\n" \
           f"{nrows} {n}\n" + "".join(f"{v[i * n + j]}\n" for j in range(n) for i in range(nrows))
    assert open(out).read() == want
    vec = str(tmp_path / "v.txt")
    assert subprocess.run([harness, "vec", raw, str(nrows * n), vec], capture_output=True).returncode == 0
    assert open(vec).read() == "".join(f"{t}\n" for t in v)
    assert not os.path.exists(vec + ".tmp")                      # written through rename()


def test_parallel_parser_large_file(harness, tmp_path):
    """The triplet section is parsed by several threads (whole-file read, segments cut at white
    space, integer index -> (entry, field)); result identical to one thread and to numpy for any
    thread count, odd white space included; the first malformed integer is reported by entry."""
    rng = np.random.default_rng(5)
    nnz, nrows, ncols, p = 1_200_000, 70_000, 65_000, 2147483647
    i = rng.integers(1, nrows + 1, nnz); j = rng.integers(1, ncols + 1, nnz)
    x = rng.integers(-50, 2 ** 31 - 1, nnz)
    mtx = str(tmp_path / "big.mtx")
    with open(mtx, "w") as f:
        f.write("%%MatrixMarket matrix coordinate integer general\n% c\n")
        f.write(f"{nrows} {ncols} {nnz}\n")
        rows = [f"{a} {b} {c}" for a, b, c in zip(i[:1000], j[:1000], x[:1000])]
        # odd but legal layouts for fscanf("%d %d %d\n"): tabs, CRLF, an entry split over two lines
        rows[3] = rows[3].replace(" ", "\t")
        rows[5] = rows[5].replace(" ", "   ") + "\r"
        rows[7] = rows[7].replace(" ", "\n", 1)
        f.write("\n".join(rows) + "\n")
        np.savetxt(f, np.stack([i[1000:], j[1000:], x[1000:]], axis=1), fmt="%d")
    want_x = ((x + (1 << 32)) % (1 << 32)) % p
    results = []
    for threads in ("1", "3", "16"):
        env = dict(os.environ, BLK_PARSE_THREADS=threads)
        out = str(tmp_path / f"m{threads}.bin")
        r = subprocess.run([harness, "load", mtx, str(p), out], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        raw = np.fromfile(out, dtype=np.uint32)
        body = raw[4:]
        assert np.array_equal(body[:nnz].view(np.int32), i - 1) and np.array_equal(body[nnz:2 * nnz].view(np.int32), j - 1)
        assert np.array_equal(body[2 * nnz:], want_x.astype(np.uint32))
        results.append(raw)
    assert all(np.array_equal(results[0], r) for r in results[1:])
    # malformed integer deep inside the file, and a truncated file
    txt = open(mtx).read().split("\n")
    bad_line = 3 + 800_000
    txt[bad_line] = txt[bad_line].split(" ")[0] + " 12x4 7"
    bad = str(tmp_path / "bad.mtx")
    open(bad, "w").write("\n".join(txt))
    for threads in ("1", "16"):
        r = subprocess.run([harness, "load", bad, str(p), str(tmp_path / "o.bin")], capture_output=True, text=True,
                           env=dict(os.environ, BLK_PARSE_THREADS=threads))
        # (entry 7 occupies two lines, so file line 3 + 800000 holds entry 799999)
        assert r.returncode == 1 and "parse error entry 799999" in r.stderr, r.stderr[-200:]
    short = str(tmp_path / "short.mtx")
    open(short, "w").write("\n".join(open(mtx).read().split("\n")[:3 + 500_000]) + "\n")
    r = subprocess.run([harness, "load", short, str(p), str(tmp_path / "o.bin")], capture_output=True, text=True,
                       env=dict(os.environ, BLK_PARSE_THREADS="8"))
    assert r.returncode == 1 and "parse error entry" in r.stderr
