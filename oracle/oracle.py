"""ctypes bindings for the CPU checker -- TEST INFRASTRUCTURE ONLY.

Two libraries are wrapped with the same Python surface:

* ``Oracle()``      -> oracle/liboracle.so, our C restatement (lanczos_oracle.c)
* ``Reference()``   -> oracle/_ref/libref_seq.so, the UNMODIFIED reference
  (sequential/lanczos_modp.c compiled with -Dmain=lanczos_ref_main).
  It keeps ``n`` and ``prime`` as C globals (sequential/lanczos_modp.c:39-40);
  the wrapper sets them before every call.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build(force: bool = False) -> None:
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    if force or not os.path.exists(os.path.join(HERE, "liboracle.so")):
        subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if os.path.isdir("/root/reference/sequential") and (
            force or not os.path.exists(os.path.join(REF_DIR, "libref_seq.so"))):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def have_reference() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "libref_seq.so"))


def block_pad(nrows: int, ncols: int, n: int, right: bool) -> int:
    """block_size_pad of sequential/lanczos_modp.c:594-597 (in u32 elements)."""
    N, Mc = (ncols, nrows) if right else (nrows, ncols)
    up = lambda a: ((a + n - 1) // n) * n
    return max(up(N), up(Mc)) * n


class Oracle:
    """Our plain-C restatement."""

    kind = "port"

    def __init__(self):
        build()
        L = C.CDLL(os.path.join(HERE, "liboracle.so"))
        L.orc_start_block.argtypes = [_u32p, C.c_long, C.c_uint64]
        L.orc_spmv.argtypes = [_u32p, C.c_int, C.c_int, C.c_long, _i32p, _i32p, _u32p, _u32p,
                               C.c_int, C.c_int, C.c_uint64]
        L.orc_block_dot_products.argtypes = [_u32p, _u32p, C.c_long, _u32p, _u32p, C.c_int, C.c_uint64]
        L.orc_semi_inverse.argtypes = [_u32p, _u32p, _u32p, C.c_int, C.c_uint64]
        L.orc_semi_inverse.restype = C.c_int
        L.orc_orthogonalize.argtypes = [_u32p, _u32p, _u32p, _u32p, _u32p, _u32p, _u32p, C.c_long,
                                        _u32p, C.c_int, C.c_uint64]
        L.orc_lanczos_run.argtypes = [C.c_int, C.c_int, C.c_long, _i32p, _i32p, _u32p, C.c_int,
                                      C.c_uint64, C.c_int, C.c_int, _u32p, _u32p, _u32p, _u32p,
                                      C.POINTER(C.c_int)]
        L.orc_lanczos_run.restype = C.c_int
        L.orc_invmod.argtypes = [C.c_uint32, C.c_uint32]
        L.orc_invmod.restype = C.c_uint32
        self.L = L

    # --- reference-named operations (host numpy arrays) -------------------
    def start_block(self, count: int, p: int) -> np.ndarray:
        v = np.zeros(count, dtype=np.uint32)
        self.L.orc_start_block(v, count, p)
        return v

    def sparse_matrix_vector_product(self, M, x, transpose: bool, n: int, p: int) -> np.ndarray:
        rows = M.ncols if transpose else M.nrows
        y = np.zeros(rows * n, dtype=np.uint32)
        self.L.orc_spmv(y, M.nrows, M.ncols, M.nnz, M.i, M.j, M.x, np.ascontiguousarray(x),
                        int(transpose), n, p)
        return y

    def block_dot_products(self, N: int, Av, v, n: int, p: int):
        a = np.zeros(n * n, dtype=np.uint32)
        b = np.zeros(n * n, dtype=np.uint32)
        self.L.orc_block_dot_products(a, b, N, np.ascontiguousarray(Av), np.ascontiguousarray(v), n, p)
        return a, b

    def semi_inverse(self, U, n: int, p: int):
        winv = np.zeros(n * n, dtype=np.uint32)
        d = np.zeros(n, dtype=np.uint32)
        npiv = self.L.orc_semi_inverse(np.ascontiguousarray(U), winv, d, n, p)
        return npiv, winv, d

    def orthogonalize(self, v, p_blk, d, vtAv, vtAAv, winv, N: int, Av, n: int, p: int):
        """Returns (next_v rows [0,N), new p block) without touching the inputs."""
        tmp = np.zeros(N * n, dtype=np.uint32)
        pn = np.array(p_blk[:N * n], dtype=np.uint32, copy=True)
        self.L.orc_orthogonalize(np.ascontiguousarray(v[:N * n]), tmp, pn, np.ascontiguousarray(d),
                                 np.ascontiguousarray(vtAv), np.ascontiguousarray(vtAAv),
                                 np.ascontiguousarray(winv), N, np.ascontiguousarray(Av[:N * n]), n, p)
        return tmp, pn

    def lanczos_run(self, M, n: int, p: int, right: bool, stop_after: int = -1, state=None):
        """Run the main loop.  Returns dict(v,tmp,Av,p,iters,stopped) with padded blocks."""
        pad = block_pad(M.nrows, M.ncols, n, right)
        N = M.ncols if right else M.nrows
        if state is None:
            v = np.zeros(pad, dtype=np.uint32)
            v[:N * n] = self.start_block(N * n, p)
            tmp = np.zeros(pad, dtype=np.uint32)
            Av = np.zeros(pad, dtype=np.uint32)
            pp = np.zeros(pad, dtype=np.uint32)
            it = C.c_int(0)
        else:
            v, tmp, Av, pp = (np.array(state[k], dtype=np.uint32, copy=True) for k in ("v", "tmp", "Av", "p"))
            it = C.c_int(int(state["iters"]))
        stopped = self.L.orc_lanczos_run(M.nrows, M.ncols, M.nnz, M.i, M.j, M.x, n, p, int(right),
                                         stop_after, v, tmp, Av, pp, C.byref(it))
        return dict(v=v, tmp=tmp, Av=Av, p=pp, iters=it.value, stopped=bool(stopped))


class _RefMatrix(C.Structure):
    # struct sparsematrix_t, sequential/lanczos_modp.c:55-62
    _fields_ = [("nrows", C.c_int), ("ncols", C.c_int), ("nnz", C.c_long),
                ("i", C.c_void_p), ("j", C.c_void_p), ("x", C.c_void_p)]


class Reference:
    """The reference's own object code (sequential build), function by function."""

    kind = "reference"

    def __init__(self, lib: str = "libref_seq.so"):
        path = os.path.join(REF_DIR, lib)
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        L = C.CDLL(path)
        self.L = L
        self._n = C.c_long.in_dll(L, "n")
        self._prime = C.c_uint64.in_dll(L, "prime")
        L.semi_inverse.restype = C.c_int
        L.invmod.restype = C.c_uint32
        L.invmod.argtypes = [C.c_uint32, C.c_uint32]
        L.random64.restype = C.c_uint64

    def _set(self, n, p):
        self._n.value = n
        self._prime.value = p

    @staticmethod
    def _mat(M):
        return _RefMatrix(M.nrows, M.ncols, M.nnz, M.i.ctypes.data, M.j.ctypes.data, M.x.ctypes.data)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def sparse_matrix_vector_product(self, M, x, transpose: bool, n: int, p: int) -> np.ndarray:
        self._set(n, p)
        rows = M.ncols if transpose else M.nrows
        y = np.zeros(rows * n, dtype=np.uint32)
        x = np.ascontiguousarray(x, dtype=np.uint32)
        m = self._mat(M)
        self.L.sparse_matrix_vector_product(self._p(y), C.byref(m), self._p(x), C.c_bool(transpose))
        return y

    def block_dot_products(self, N: int, Av, v, n: int, p: int):
        self._set(n, p)
        # the reference reads rows up to ceil(N/n)*n (zero padding), :447-452
        padN = ((N + n - 1) // n) * n
        Avp = np.zeros(padN * n, dtype=np.uint32); Avp[:N * n] = Av[:N * n]
        vp = np.zeros(padN * n, dtype=np.uint32); vp[:N * n] = v[:N * n]
        a = np.zeros(n * n, dtype=np.uint32)
        b = np.zeros(n * n, dtype=np.uint32)
        self.L.block_dot_products(self._p(a), self._p(b), C.c_int(N), self._p(Avp), self._p(vp))
        return a, b

    def semi_inverse(self, U, n: int, p: int):
        self._set(n, p)
        U = np.ascontiguousarray(U, dtype=np.uint32)
        winv = np.zeros(n * n, dtype=np.uint32)
        d = np.zeros(n, dtype=np.uint32)
        npiv = self.L.semi_inverse(self._p(U), self._p(winv), self._p(d))
        return npiv, winv, d

    def orthogonalize(self, v, p_blk, d, vtAv, vtAAv, winv, N: int, Av, n: int, p: int):
        self._set(n, p)
        padN = ((N + n - 1) // n) * n
        def padded(a):
            o = np.zeros(padN * n, dtype=np.uint32); o[:N * n] = a[:N * n]; return o
        vv, pn, av, tmp = padded(v), padded(p_blk), padded(Av), np.zeros(padN * n, dtype=np.uint32)
        d = np.ascontiguousarray(d, dtype=np.uint32)
        args = [np.ascontiguousarray(a, dtype=np.uint32) for a in (vtAv, vtAAv, winv)]
        self.L.orthogonalize(self._p(vv), self._p(tmp), self._p(pn), self._p(d), self._p(args[0]),
                             self._p(args[1]), self._p(args[2]), C.c_int(N), self._p(av))
        return tmp[:N * n].copy(), pn[:N * n].copy()

    def start_block(self, count: int, p: int) -> np.ndarray:
        """v[i] = random64() % prime (sequential/lanczos_modp.c:624-625); resets the RNG state."""
        st = (C.c_uint64 * 4).in_dll(self.L, "rng_state")
        st[0], st[1], st[2], st[3] = 0x1415926535, 0x8979323846, 0x2643383279, 0x5028841971
        out = np.empty(count, dtype=np.uint32)
        r = self.L.random64
        for t in range(count):
            out[t] = r() % p
        return out

    def invmod(self, a, m):
        return self.L.invmod(a, m)


def run_reference_cli(binary: str, mtx: str, p: int, n: int, right: bool, out: str | None = None,
                      stop_after: int | None = None, cwd: str | None = None, env=None, extra=()):
    """Run a reference CLI binary from oracle/_ref (lanczos_modp_seq / _omp / checker_modp)."""
    cmd = [os.path.join(REF_DIR, binary), "--matrix", mtx, "--prime", str(p), "--n", str(n)]
    cmd.append("--right" if right else "--left")
    if out:
        cmd += ["--output-file", out]
    if stop_after:
        cmd += ["--stop-after", str(stop_after)]
    cmd += list(extra)
    return subprocess.run(cmd, capture_output=True, text=True, cwd=cwd, env=env)
