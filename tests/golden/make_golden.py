"""Regenerates tests/golden/*.npz.  Run in the BUILD container (needs
/root/reference for cage4.mtx; the GPU box only reads the committed .npz):

    python tests/golden/make_golden.py

Fixtures:
  kat_small.npz -- the reference's only deterministic known-answer case,
      test_small_spgemm (SpGEMM_cuda/main.cu:149-246): A 4x6, B 6x4 and the result
      derived by hand from those inputs; NOT produced by the
      oracle, so it pins the oracle.
  cage4_sq.npz  -- cage4.mtx (the shipped fixture, SpGEMM_cuda/cage4.mtx) and
      cage4 * cage4 computed with scipy.sparse (all values positive, so scipy's
      zero-dropping cannot bite), columns sorted.  Also independent of the oracle.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from benchmark_spgemm_using_csr_b200.mtx import read_mtx  # noqa: E402


def kat_small():
    # main.cu:153-201
    rowptrA = np.array([0, 1, 4, 5, 6], dtype=np.int32)
    colA = np.array([0, 1, 2, 3, 3, 1], dtype=np.int32)
    valA = np.array([10, 20, 30, 40, 50, 60], dtype=np.float64)      # (i+1)*10
    rowptrB = np.array([0, 1, 3, 5, 5, 5, 7], dtype=np.int32)
    colB = np.array([0, 1, 3, 0, 1, 1, 3], dtype=np.int32)
    valB = np.array([1, 2, 3, 4, 5, 6, 7], dtype=np.float64)         # i+1
    # by hand: row0 = 10*B0 -> (0:10); row1 = 20*B1 + 30*B2 + 40*B3 -> (0:120, 1:40+150=190, 3:60);
    # row2 = 50*B3 -> empty; row3 = 60*B1 -> (1:120, 3:180)
    rowptrC = np.array([0, 1, 4, 4, 6], dtype=np.int32)
    colC = np.array([0, 0, 1, 3, 1, 3], dtype=np.int32)
    valC = np.array([10, 120, 190, 60, 120, 180], dtype=np.float64)
    products = np.array([1, 4, 0, 2], dtype=np.int64)
    np.savez(os.path.join(HERE, "kat_small.npz"), m=4, k=6, n=4, rowptrA=rowptrA, colA=colA, valA=valA,
             rowptrB=rowptrB, colB=colB, valB=valB, rowptrC=rowptrC, colC=colC, valC=valC, products=products)


def cage4():
    A = read_mtx("/root/reference/SpGEMM_cuda/cage4.mtx")
    S = sp.csr_matrix((A.val, A.col, A.rowptr), shape=(A.rows, A.cols))
    assert (A.val > 0).all()
    C = (S @ S).tocsr()
    C.sort_indices()
    np.savez(os.path.join(HERE, "cage4_sq.npz"), m=A.rows, k=A.cols, n=A.cols, rowptrA=A.rowptr, colA=A.col, valA=A.val,
             rowptrC=C.indptr.astype(np.int32), colC=C.indices.astype(np.int32), valC=C.data,
             products=np.array([27, 27, 27, 27, 33, 33, 33, 33, 29], dtype=np.int64))


if __name__ == "__main__":
    kat_small()
    cage4()
    print("golden fixtures written to", HERE)
