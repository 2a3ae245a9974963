import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name + ".npz"))
    return load


def assert_csr_equal(got, want, exact_values=True, rtol=0.0, what=""):
    """got/want = (rowptr, col, val).  rowptr and col bit-exact (the reference's
    check, ref_spgemm.h:79-118); values exact or within rtol relative."""
    grp, gc, gv = got
    wrp, wc, wv = want
    assert np.array_equal(np.asarray(grp, dtype=np.int64), np.asarray(wrp, dtype=np.int64)), f"{what}: rowptrC differs"
    assert gc.shape == wc.shape and np.array_equal(gc, wc), f"{what}: colC differs"
    assert gv.dtype == wv.dtype, f"{what}: value dtype {gv.dtype} vs {wv.dtype}"
    if exact_values:
        assert np.array_equal(gv, wv), f"{what}: valC differs (exact compare)"
    else:
        denom = np.maximum(np.abs(wv), np.finfo(wv.dtype).tiny)
        err = np.abs(gv.astype(np.float64) - wv.astype(np.float64)) / denom
        assert err.size == 0 or err.max() <= rtol, f"{what}: max rel err {err.max()} > {rtol}"
