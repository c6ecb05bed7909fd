"""CPU suite, part 2: the C-ABI library builds, loads and exports what include/blk_lanczos.h
declares; the host-side logic that needs no GPU; failure is loud when there is no device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(lib):
    L = lib.load_library()
    header = open(os.path.join(ROOT, "include", "blk_lanczos.h")).read()
    declared = set(re.findall(r"\b(blk_[a-z_0-9]+)\s*\(", header))
    declared -= {"blk_ctx", "blk_params", "blk_info"}
    assert declared == set(lib.ABI_SYMBOLS), declared ^ set(lib.ABI_SYMBOLS)
    for s in declared:
        assert hasattr(L, s), s
    assert L.blk_abi_version() == lib.BLK_ABI_VERSION


def test_struct_layout_matches_header(lib, tmp_path):
    """ctypes mirrors of blk_params / blk_info have the size and field offsets gcc gives the header."""
    import ctypes as C
    import subprocess
    src = tmp_path / "sz.c"
    fields_p = [f for f, _ in lib.blk_params._fields_]
    fields_i = [f for f, _ in lib.blk_info._fields_]
    body = "".join(f'printf("%zu\\n", offsetof(blk_params, {f}));' for f in fields_p)
    body += "".join(f'printf("%zu\\n", offsetof(blk_info, {f}));' for f in fields_i)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "blk_lanczos.h"\nint main(void){'
                   'printf("%zu\\n%zu\\n", sizeof(blk_params), sizeof(blk_info));' + body + 'return 0;}')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    nums = [int(t) for t in subprocess.check_output([str(exe)], text=True).split()]
    assert nums[0] == C.sizeof(lib.blk_params) and nums[1] == C.sizeof(lib.blk_info)
    offs = [getattr(lib.blk_params, f).offset for f in fields_p] + [getattr(lib.blk_info, f).offset for f in fields_i]
    assert nums[2:] == offs


def test_block_pad(lib):
    from oracle.oracle import block_pad
    for (r, c, n, right) in [(20000, 19000, 1, False), (10, 7, 4, False), (10, 17, 4, True), (38132, 48630, 4, True),
                             (95368, 123412, 8, True), (5, 5, 3, False)]:
        assert lib.block_pad(r, c, n, right) == block_pad(r, c, n, right)


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    M = lib.synth.uniform_rows(50, 40, 3)
    with pytest.raises(lib.BlkError, match="no CUDA device"):
        lib.BlockLanczos(M, n=2, prime=65537)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "block-lanczos-algorithm-parallelization_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and "oracle" not in src, (dp, f)


def test_mtx_roundtrip(tmp_path, lib):
    M = lib.synth.powerlaw_rows(60, 50, mean=4, seed=1, with_empty_rows=3)
    path = str(tmp_path / "m.mtx")
    lib.synth.write_mtx(path, M)
    M2 = lib.synth.read_mtx(path)
    assert (M2.nrows, M2.ncols, M2.nnz) == (M.nrows, M.ncols, M.nnz)
    assert np.array_equal(M2.i, M.i) and np.array_equal(M2.j, M.j) and np.array_equal(M2.x, M.x)


def test_python_start_block_is_the_references(lib, oracle):
    for p in (65537, 2147483647, 7):
        assert np.array_equal(lib.synth.reference_start_block(5000, p), oracle.start_block(5000, p))
