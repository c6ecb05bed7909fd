import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_cases(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    import blk_lanczos_b200 as B
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    M = B.SparseCOO(int(z["nrows"]), int(z["ncols"]), z["i"].astype(np.int32), z["j"].astype(np.int32),
                    z["x"].astype(np.uint32))
    return z, M


@pytest.fixture(scope="session")
def oracle():
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def lib():
    """The built CUDA library; building is part of the contract (__graft_entry__.build)."""
    import blk_lanczos_b200 as B
    if not os.path.exists(B.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return B
