"""BASELINE.json configs at full size: too big for the oracle to finish in
seconds, so they are checked through size-independent properties --
closed-form counts, strictly ascending columns, and the linearity checksum
sum(C) = (1^T A)(B 1), which is exact for the integer-valued inputs."""
import numpy as np
import pytest

from benchmark_spgemm_using_csr_b200 import generators as gen, spgemm

pytestmark = pytest.mark.gpu


def _properties(A, B, rp, col, val, st, P=None, nnzC=None):
    assert rp[0] == 0 and rp[-1] == col.size == val.size == st["nnzC"]
    assert (np.diff(rp.astype(np.int64)) >= 0).all()
    if P is not None:
        assert st["products"] == P
    if nnzC is not None:
        assert st["nnzC"] == nnzC
    assert col.min() >= 0 and col.max() < B.cols
    # strictly ascending inside every row
    d = np.diff(col.astype(np.int64))
    inner = np.ones(col.size - 1, dtype=bool)
    ends = rp[1:-1].astype(np.int64)
    ends = ends[(ends > 0) & (ends < col.size)]
    inner[ends - 1] = False
    assert (d[inner] > 0).all()
    # checksum of checksums: sum_ij C_ij = sum_k colsum(A)_k * rowsum(B)_k
    colsumA = np.bincount(A.col, weights=A.val.astype(np.float64), minlength=A.cols)
    rowsumB = np.add.reduceat(B.val.astype(np.float64), B.rowptr[:-1].astype(np.int64)) * (np.diff(B.rowptr) > 0)
    want = float(np.dot(colsumA, rowsumB))
    got = float(val.astype(np.float64).sum())
    assert abs(got - want) <= 1e-9 * abs(want)
    # per-row checksum on a sample of rows: row i of C sums to A_i . rowsum(B)
    rows = np.linspace(0, A.rows - 1, 2000).astype(np.int64)
    for i in rows:
        a = slice(A.rowptr[i], A.rowptr[i + 1])
        w = float(np.dot(A.val[a].astype(np.float64), rowsumB[A.col[a]]))
        g = float(val[rp[i]:rp[i + 1]].astype(np.float64).sum())
        assert abs(g - w) <= 1e-6 * max(abs(w), 1.0)


def test_config1_poisson5pt_1024():
    A = gen.poisson5pt(1024, 1024)
    rp, col, val, st = spgemm(A, A, return_stats=True)
    _properties(A, A, rp, col, val, st, P=26177544, nnzC=13611012)


def test_config2_poisson27pt_128():
    A = gen.poisson27pt(128, 128, 128)
    rp, col, val, st = spgemm(A, A, return_stats=True)
    _properties(A, A, rp, col, val, st, P=1489355288, nnzC=254840104)


def test_config4_uniform_rect_f32():
    A = gen.uniform_rect(4194304, 1048576, per_row=8, seed=1, dtype=np.float32)
    B = gen.uniform_rect(1048576, 4194304, per_row=8, seed=2, value_seed=3, dtype=np.float32)
    rp, col, val, st = spgemm(A, B, return_stats=True)
    _properties(A, B, rp, col, val, st, P=268435456)
