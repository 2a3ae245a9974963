import os

import numpy as np

from benchmark_spgemm_using_csr_b200 import generators as gen
from benchmark_spgemm_using_csr_b200.mtx import read_mtx


def _check_csr(A):
    assert A.rowptr.dtype == np.int32 and A.col.dtype == np.int32
    assert A.rowptr[0] == 0 and A.rowptr[-1] == A.col.size == A.val.size
    assert (np.diff(A.rowptr) >= 0).all()
    if A.nnz:
        assert A.col.min() >= 0 and A.col.max() < A.cols
        d = np.diff(A.col.astype(np.int64))
        inner = np.ones(A.nnz - 1, dtype=bool)
        ends = A.rowptr[1:-1]
        ends = ends[(ends > 0) & (ends < A.nnz)]
        inner[ends - 1] = False
        assert (d[inner] > 0).all(), "columns must be strictly ascending inside a row"


def test_poisson_counts():
    # CUSP gallery nnz for the reference's stock sizes (main.cu:30-53)
    for f, args, nnz in ((gen.poisson5pt, (256, 256), 326656), (gen.poisson9pt, (256, 256), 586756),
                         (gen.poisson7pt, (51, 51, 51), 912951), (gen.poisson27pt, (51, 51, 51), 3442951)):
        A = f(*args)
        assert A.nnz == nnz
        _check_csr(A)
        assert A.val.min() >= 1 and A.val.max() <= 9 and (A.val == np.round(A.val)).all()


def test_rmat_and_uniform():
    A = gen.rmat(10, 8, seed=5)
    _check_csr(A)
    assert A.rows == 1024 and 0 < A.nnz <= 8 * 1024
    B = gen.rmat(10, 8, seed=5)
    assert np.array_equal(A.col, B.col) and np.array_equal(A.val, B.val)        # reproducible
    U = gen.uniform_rect(1000, 64, per_row=8, seed=2)
    _check_csr(U)
    assert (np.diff(U.rowptr) == 8).all()


def test_random_csr_and_slices():
    A = gen.random_csr(50, 30, np.arange(50) % 31, seed=1)
    _check_csr(A)
    assert (np.diff(A.rowptr) == np.minimum(np.arange(50) % 31, 30)).all()
    S = A.row_slice(10, 20)
    _check_csr(S)
    assert S.rows == 10 and S.nnz == A.rowptr[20] - A.rowptr[10]
    R = gen.real_values(1000, 3)
    assert R.min() > 0 and R.max() <= 1


def test_mtx_reader(tmp_path):
    p = tmp_path / "t.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real symmetric\n% c\n3 3 4\n1 1 2.0\n2 1 3.0\n3 3 1.0\n3 2 -1.5\n")
    A = read_mtx(str(p))
    _check_csr(A)
    D = np.zeros((3, 3))
    for i in range(3):
        D[i, A.col[A.rowptr[i]:A.rowptr[i + 1]]] = A.val[A.rowptr[i]:A.rowptr[i + 1]]
    assert np.array_equal(D, np.array([[2, 3, 0], [3, 0, -1.5], [0, -1.5, 1.0]]))
    q = tmp_path / "p.mtx"
    q.write_text("%%MatrixMarket matrix coordinate pattern general\n2 3 2\n1 3\n2 1\n")
    B = read_mtx(str(q))
    assert B.cols == 3 and B.col.tolist() == [2, 0] and B.val.tolist() == [1.0, 1.0]


def test_rmat_counter_is_identical_in_numpy_and_torch():
    """bench.py generates config 5 (R-MAT scale 24) on the GPU with torch; the CPU side (oracle,
    tests) must see the same matrix from numpy."""
    for scale, ef in ((6, 4), (11, 16)):
        A = gen.rmat_counter(scale, ef, seed=3, value_seed=4)
        rp, col, val = gen.rmat_counter_torch(scale, ef, seed=3, value_seed=4, device="cpu")
        assert np.array_equal(A.rowptr, rp.numpy()) and np.array_equal(A.col, col.numpy())
        assert np.array_equal(A.val, val.numpy())
        assert (np.diff(A.rowptr) >= 0).all() and A.rowptr[-1] == A.col.size
        for i in range(0, A.rows, 97):
            c = A.col[A.rowptr[i]:A.rowptr[i + 1]]
            assert (np.diff(c) > 0).all()
