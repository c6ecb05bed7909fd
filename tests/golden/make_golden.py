"""Generate tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE (run in the build container,
where /root/reference exists; the GPU box only reads the committed .npz files).

Two kinds of vectors:
  * cli_*:   the reference's sequential CLI (oracle/_ref/lanczos_modp_seq, built by oracle/Makefile
             from /root/reference/sequential with only the prime cap relaxed) run end to end on a
             seeded matrix; the kernel block it wrote, byte for byte (parsed), and whether the
             unmodified checker_modp accepted it.
  * loop_*:  the reference's own object code (oracle/_ref/libref_seq.so) driven function by
             function in the order of block_lanczos (sequential/lanczos_modp.c:631-659) for K
             iterations: the four blocks and the n x n matrices of the last iteration.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import blk_lanczos_b200 as B                                 # noqa: E402  (synth only, no GPU)
from oracle.oracle import Reference, block_pad, REF_DIR      # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
synth = B.synth

CLI_CASES = [
    # name, matrix builder, p, n, right
    ("cli_left_n1_p65537", lambda: synth.uniform_rows(200, 190, 6, seed=3), 65537, 1, False),
    ("cli_left_n2_p65537", lambda: synth.uniform_rows(260, 250, 5, seed=1), 65537, 2, False),
    ("cli_right_n4_p65537", lambda: synth.uniform_nnz(180, 230, 1400, seed=2, order="col"), 65537, 4, True),
    ("cli_right_n8_mersenne", lambda: synth.uniform_nnz(300, 390, 2600, seed=5, order="col"), 2147483647, 8, True),
    ("cli_left_n3_p1073741789", lambda: synth.powerlaw_rows(240, 200, mean=6, seed=7, with_empty_rows=4), 1073741789, 3, False),
    ("cli_left_n16_mersenne", lambda: synth.powerlaw_rows(420, 400, mean=8, seed=9, order="file"), 2147483647, 16, False),
    # small primes: the iteration breaks down by chance (v^T A v == 0) long before a kernel vector is
    # found; the reference then prints KO and writes a block that checker_modp rejects -- must be reproduced as is
    ("cli_left_n1_p251_breakdown", lambda: synth.uniform_rows(300, 280, 5, seed=4), 251, 1, False),
    ("cli_right_n2_p7_breakdown", lambda: synth.uniform_nnz(200, 260, 1500, seed=6, order="col"), 7, 2, True),
]

LOOP_CASES = [
    ("loop_left_n4_p65537_k5", lambda: synth.powerlaw_rows(350, 300, mean=7, seed=21, with_empty_rows=6), 65537, 4, False, 5),
    ("loop_right_n8_mersenne_k4", lambda: synth.uniform_nnz(260, 330, 2200, seed=22, order="col"), 2147483647, 8, True, 4),
    ("loop_left_n1_p1073741789_k9", lambda: synth.uniform_rows(150, 140, 4, seed=23), 1073741789, 1, False, 9),
    ("loop_left_n5_p65537_k3", lambda: synth.uniform_rows(123, 117, 5, seed=24, order="file"), 65537, 5, False, 3),
    ("loop_right_n32_mersenne_k2", lambda: synth.uniform_nnz(200, 270, 2500, seed=25), 2147483647, 32, True, 2),
]


def run_cli(name, M, p, n, right):
    with tempfile.TemporaryDirectory() as td:
        mtx, out = os.path.join(td, "m.mtx"), os.path.join(td, "k.mtx")
        synth.write_mtx(mtx, M)
        side = "--right" if right else "--left"
        r = subprocess.run([os.path.join(REF_DIR, "lanczos_modp_seq"), "--matrix", mtx, "--prime", str(p),
                            "--n", str(n), side, "--output-file", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        iters = int(r.stdout.split("after")[-1].split("iterations")[0])
        chk = subprocess.run([os.path.join(REF_DIR, "checker_modp"), "--matrix", mtx, "--prime", str(p),
                              "--kernel", out, side], capture_output=True, text=True)
        kernel = synth.read_kernel_block(out)
        text = open(out, "rb").read()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), nrows=M.nrows, ncols=M.ncols, i=M.i, j=M.j, x=M.x,
                        p=p, n=n, right=right, kernel=kernel, iters=iters, checker_ok=("OK" in chk.stdout),
                        ok_v=("OK:    v != 0" in r.stdout), ok_vtM=("OK: vt*M == 0" in r.stdout),
                        file_sha256=np.frombuffer(__import__("hashlib").sha256(text).digest(), dtype=np.uint8))
    print(name, "iters", iters, "checker", "OK" in chk.stdout)


def run_loop(name, M, p, n, right, K):
    R = Reference()
    Mp = M.reduced(p)
    N = M.ncols if right else M.nrows
    Mc = M.nrows if right else M.ncols
    pad = block_pad(M.nrows, M.ncols, n, right)
    v = np.zeros(pad, np.uint32); tmp = np.zeros(pad, np.uint32)
    Av = np.zeros(pad, np.uint32); pp = np.zeros(pad, np.uint32)
    v[:N * n] = R.start_block(N * n, p)
    v0 = v.copy()
    small = {}
    for _ in range(K):
        tmp[:Mc * n] = R.sparse_matrix_vector_product(Mp, v, not right, n, p)      # :635
        Av[:N * n] = R.sparse_matrix_vector_product(Mp, tmp, right, n, p)          # :636
        vtAv, vtAAv = R.block_dot_products(N, Av, v, n, p)                         # :640
        npiv, winv, d = R.semi_inverse(vtAv, n, p)                                 # :644
        small = dict(vtAv=vtAv, vtAAv=vtAAv, winv=winv, d=d, npiv=npiv)
        assert npiv > 0
        nv, np_ = R.orthogonalize(v, pp, d, vtAv, vtAAv, winv, N, Av, n, p)        # :652
        tmp[:N * n] = nv; pp[:N * n] = np_
        v[:N * n] = tmp[:N * n]                                                    # :655-656
    np.savez_compressed(os.path.join(OUT, name + ".npz"), nrows=M.nrows, ncols=M.ncols, i=M.i, j=M.j, x=M.x,
                        p=p, n=n, right=right, K=K, v0=v0, v=v, tmp=tmp, Av=Av, pblk=pp, **small)
    print(name, "done")


if __name__ == "__main__":
    only = sys.argv[1:]                      # optional: regenerate just the named cases
    for c in CLI_CASES:
        if not only or c[0] in only:
            run_cli(c[0], c[1](), *c[2:])
    for c in LOOP_CASES:
        if not only or c[0] in only:
            run_loop(c[0], c[1](), *c[2:])
