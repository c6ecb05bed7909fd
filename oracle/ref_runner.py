"""Time the reference's own CPU implementation on a bounded sample -- TEST/BENCH INFRASTRUCTURE.

Run as a subprocess by bench.py (cpu_baseline leg and --impl reference) so that OMP_NUM_THREADS /
OMP_STACKSIZE are in the environment before libgomp starts and the 1 GiB stack limit the OpenMP
reference needs (its SpMV keeps a u64 copy of the whole output block on every thread's stack,
openMP/lanczos_modp.c:336,352, hence setStackLimit :142-164) can be raised for this process only.

It loads oracle/_ref/libref_{omp,seq}.so -- the UNMODIFIED reference compiled by oracle/Makefile
with -Dmain=lanczos_ref_main -- fills its struct sparsematrix_t with an in-memory synthetic
matrix, sets the globals (n, prime, stop_after) and calls the reference's block_lanczos()
(openMP/lanczos_modp.c:911).  Falls back to the oracle port (oracle/liboracle.so, 1 thread) when
oracle/_ref is absent.  Prints one JSON line on stdout.
"""
import argparse
import ctypes as C
import json
import os
import resource
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lib", default="omp", choices=["omp", "seq", "port"])
    ap.add_argument("--rows", type=int, required=True)
    ap.add_argument("--cols", type=int, required=True)
    ap.add_argument("--mean", type=float, default=30.0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--n", type=int, required=True)
    ap.add_argument("--prime", type=int, required=True)
    ap.add_argument("--right", type=int, default=0)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--spmv-reps", type=int, default=0,
                    help="also time sparse_matrix_vector_product alone (both directions), this many calls each")
    a = ap.parse_args()

    try:
        soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
        want = 1 << 30
        if soft != resource.RLIM_INFINITY and soft < want:
            resource.setrlimit(resource.RLIMIT_STACK, (want if hard == resource.RLIM_INFINITY else min(want, hard), hard))
    except Exception:
        pass

    import numpy as np
    import blk_lanczos_b200 as B          # synth only (numpy); no GPU code runs here
    M = B.synth.powerlaw_rows(a.rows, a.cols, mean=a.mean, seed=a.seed).reduced(a.prime)

    # keep the reference's progress prints off our stdout
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    threads = int(os.environ.get("OMP_NUM_THREADS", "1")) if a.lib == "omp" else 1
    kind = "reference"
    path = os.path.join(HERE, "_ref", f"libref_{a.lib}.so")
    times = []
    spmv = None
    if a.lib != "port" and os.path.exists(path):
        L = C.CDLL(path)

        class Mat(C.Structure):
            _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("nnz", C.c_long),
                        ("i", C.c_void_p), ("j", C.c_void_p), ("x", C.c_void_p)]
        m = Mat(M.nrows, M.ncols, M.nnz, M.i.ctypes.data, M.j.ctypes.data, M.x.ctypes.data)
        C.c_long.in_dll(L, "n").value = a.n
        C.c_uint64.in_dll(L, "prime").value = a.prime
        L.block_lanczos.restype = C.c_void_p
        L.block_lanczos.argtypes = [C.c_void_p, C.c_int, C.c_bool]
        libc = C.CDLL(None)
        libc.free.argtypes = [C.c_void_p]
        for count in ([a.warmup] if a.warmup > 0 else []) + [a.iters]:
            C.c_int.in_dll(L, "stop_after").value = count
            C.c_int.in_dll(L, "n_iterations").value = 0
            t = time.perf_counter()
            v = L.block_lanczos(C.byref(m), a.n, bool(a.right))
            times.append(time.perf_counter() - t)
            libc.free(v)
        if a.spmv_reps > 0:
            # the reference's own SpMV (sequential/lanczos_modp.c:266 / openMP/lanczos_modp.c:329), timed alone
            pad = max(-(-M.nrows // a.n), -(-M.ncols // a.n)) * a.n * a.n
            rng = np.random.default_rng(3)
            x = rng.integers(0, a.prime, size=pad, dtype=np.uint64).astype(np.uint32)
            y = np.zeros(pad, dtype=np.uint32)
            # (the OpenMP variant takes block_size_pad as a fifth argument, openMP/lanczos_modp.c:329)
            extra = [C.c_long(pad)] if a.lib == "omp" else []
            L.sparse_matrix_vector_product.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_bool] + ([C.c_long] if extra else [])
            spmv = {}
            for tr in (False, True):
                L.sparse_matrix_vector_product(y.ctypes.data, C.byref(m), x.ctypes.data, tr, *extra)          # warm-up
                t = time.perf_counter()
                for _ in range(a.spmv_reps):
                    L.sparse_matrix_vector_product(y.ctypes.data, C.byref(m), x.ctypes.data, tr, *extra)
                spmv["Mt_x" if tr else "M_x"] = (time.perf_counter() - t) / a.spmv_reps
    else:
        kind, threads = "port", 1
        from oracle.oracle import Oracle
        O = Oracle()
        for count in ([a.warmup] if a.warmup > 0 else []) + [a.iters]:
            t = time.perf_counter()
            O.lanczos_run(M, a.n, a.prime, bool(a.right), stop_after=count)
            times.append(time.perf_counter() - t)
    sys.stdout.flush()
    os.dup2(saved, 1)
    sec = times[-1]
    out = dict(kind=kind, lib=a.lib, threads=threads, host_cores=os.cpu_count(), rows=a.rows, cols=a.cols, nnz=M.nnz, n=a.n,
               prime=a.prime, iters=a.iters, seconds=sec, iters_per_s=a.iters / sec,
               gnnzn_per_s=2.0 * M.nnz * a.n * a.iters / sec / 1e9)
    if spmv:
        out["spmv_seconds"] = spmv
        out["spmv_gnnzn_per_s"] = {k: M.nnz * a.n / t / 1e9 for k, t in spmv.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
