"""Three-way parity on the GPU box: the reference itself (oracle/_ref, bhSPARSE's own CUDA
kernels compiled for sm_100a) vs the CPU oracle vs this library, same inputs, live.

Covers the reference driver's own workloads (-spgemm 0..4: the 4x6 known-answer case and the
four stock Poisson sizes, main.cu:30-53,149-246), cage4.mtx squared, and the cases of
tests/ref_cases.py that walk the reference's bins up to the global merge path."""
import numpy as np
import pytest

import oracle
import ref_cases
from benchmark_spgemm_using_csr_b200 import spgemm
from benchmark_spgemm_using_csr_b200.generators import CSR
from conftest import assert_csr_equal
from oracle import ref

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (python -m oracle.build_ref)")]

RTOL = {np.float64: 1e-12, np.float32: 1e-5}
CASES = dict(ref_cases.SMALL + ref_cases.LARGE)


def _three_way(A, B, what, exact):
    dt = A.val.dtype.type
    r_rp, r_col, r_val, _ = ref.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
    o = oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
    ours = spgemm(A, B)
    assert_csr_equal((r_rp, r_col, r_val), o, exact_values=exact, rtol=RTOL[dt], what=what + ": reference vs oracle")
    assert_csr_equal(ours, (r_rp, r_col, r_val), exact_values=exact, rtol=RTOL[dt], what=what + ": library vs reference")


@pytest.mark.parametrize("dn", ["f64", "f32"])
@pytest.mark.parametrize("name", list(CASES))
def test_library_equals_reference(name, dn):
    A, B = CASES[name](ref_cases.DTYPES[dn])
    _three_way(A, B, f"{name}.{dn}", exact=not name.endswith("_real"))


def test_cage4_squared_reference(golden):
    g = golden("cage4_sq")
    m = int(g["m"])
    A = CSR(m, m, g["rowptrA"], g["colA"], g["valA"])
    _three_way(A, A, "cage4", exact=False)
