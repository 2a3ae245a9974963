"""Pins the CPU oracle (oracle/spgemm_oracle.c) against vectors that do NOT come
from the oracle: the reference's known-answer case (main.cu:149-246), cage4^2
and scipy.sparse, and checks the semantics the reference's GPU path has
(explicit zeros kept, columns ascending)."""
import numpy as np
import pytest
import scipy.sparse as sp

import oracle
from benchmark_spgemm_using_csr_b200 import generators as gen
from conftest import assert_csr_equal


def _run(A, B):
    return oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)


def _scipy(A, B):
    SA = sp.csr_matrix((A.val, A.col, A.rowptr), shape=(A.rows, A.cols))
    SB = sp.csr_matrix((B.val, B.col, B.rowptr), shape=(B.rows, B.cols))
    C = (SA @ SB).tocsr()
    C.sort_indices()
    return C.indptr, C.indices.astype(np.int32), C.data


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_kat_small(golden, dt):
    g = golden("kat_small")
    rp, c, v = oracle.spgemm(int(g["m"]), int(g["k"]), int(g["n"]), g["rowptrA"], g["colA"], g["valA"].astype(dt),
                             g["rowptrB"], g["colB"], g["valB"].astype(dt))
    assert_csr_equal((rp, c, v), (g["rowptrC"], g["colC"], g["valC"].astype(dt)), what="kat_small")
    prod, total = oracle.row_products(int(g["m"]), g["rowptrA"], g["colA"], g["rowptrB"])
    assert np.array_equal(prod, g["products"]) and total == 7


def test_cage4_squared(golden):
    g = golden("cage4_sq")
    m = int(g["m"])
    rp, c, v = oracle.spgemm(m, m, m, g["rowptrA"], g["colA"], g["valA"], g["rowptrA"], g["colA"], g["valA"])
    assert rp[-1] == 81
    assert_csr_equal((rp, c, v), (g["rowptrC"], g["colC"], g["valC"]), exact_values=False, rtol=1e-14, what="cage4")
    prod, total = oracle.row_products(m, g["rowptrA"], g["colA"], g["rowptrA"])
    assert np.array_equal(prod, g["products"]) and total == 269


@pytest.mark.parametrize("name,args,nnz,P,nnzC", [
    ("poisson5pt", (256, 256), 326656, 1629192, 846852),        # -spgemm 1 (main.cu:30-35)
    ("poisson9pt", (256, 256), 586756, 5262436, 1623076),       # -spgemm 2
    ("poisson7pt", (51, 51, 51), 912951, 6298245, 3207645),     # -spgemm 3
    ("poisson27pt", (51, 51, 51), 3442951, 90518849, 15438249),  # -spgemm 4
])
def test_stock_workloads_vs_scipy(name, args, nnz, P, nnzC):
    A = getattr(gen, name)(*args)
    assert A.nnz == nnz
    prod, total = oracle.row_products(A.rows, A.rowptr, A.col, A.rowptr)
    assert total == P
    got = _run(A, A)
    assert got[0][-1] == nnzC
    assert_csr_equal(got, _scipy(A, A), what=name)      # integer values: exact


def test_random_rect_vs_scipy():
    A = gen.random_csr(300, 200, np.arange(300) % 37, seed=3)
    B = gen.random_csr(200, 500, (np.arange(200) * 7) % 23, seed=4, value_seed=9)
    assert_csr_equal(_run(A, B), _scipy(A, B), what="random rect")
    Ar, Br = A.astype(np.float64), B.astype(np.float64)
    Ar.val[:] = gen.real_values(A.nnz, 11)
    Br.val[:] = gen.real_values(B.nnz, 12)
    assert_csr_equal(_run(Ar, Br), _scipy(Ar, Br), exact_values=False, rtol=1e-13, what="random rect real")


def test_explicit_zeros_are_kept():
    # [[1,-1]] * [[1],[1]] : scipy drops the zero, the reference's kernels do not
    # (no value test anywhere in bhsparse_cuda.h) -> one stored entry with value 0.
    rp, c, v = oracle.spgemm(1, 2, 1, np.array([0, 2], np.int32), np.array([0, 1], np.int32), np.array([1.0, -1.0]),
                             np.array([0, 1, 2], np.int32), np.array([0, 0], np.int32), np.array([1.0, 1.0]))
    assert rp.tolist() == [0, 1] and c.tolist() == [0] and v.tolist() == [0.0]


def test_empty_rows_and_empty_operands():
    A = gen.random_csr(5, 4, [0, 2, 0, 1, 0], seed=1)
    B = gen.random_csr(4, 6, [0, 0, 0, 0], seed=2)
    rp, c, v = _run(A, B)
    assert rp.tolist() == [0] * 6 and c.size == 0 and v.size == 0


def test_reference_bins():
    # bhsparse.h:377-406
    for cnt, b in [(0, 0), (1, 1), (121, 121), (122, 122), (128, 122), (129, 123), (256, 123), (257, 124),
                   (512, 124), (513, 127), (10 ** 6, 127)]:
        assert oracle.reference_bin(cnt) == b


def test_csr_sort_indices():
    rowptr = np.array([0, 3, 3, 5], np.int32)
    col = np.array([2, 0, 1, 4, 3], np.int32)
    val = np.array([20.0, 0.5, 10.0, 40.0, 30.0])
    oracle.csr_sort_indices(3, rowptr, col, val)
    assert col.tolist() == [0, 1, 2, 3, 4] and val.tolist() == [0.5, 10.0, 20.0, 30.0, 40.0]
