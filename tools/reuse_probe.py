import sys, time, numpy as np
sys.path.insert(0, ".")
from benchmark_spgemm_using_csr_b200 import BHSPARSE_CUDA, NUM_PLATFORMS, bhsparse, generators as gen
A = gen.poisson27pt(128, 128, 128)
pl = [False] * NUM_PLATFORMS; pl[BHSPARSE_CUDA] = True
bh = bhsparse(); bh.initPlatform(pl)
rp = np.zeros(A.rows + 1, dtype=np.int32)
bh.initData(A.rows, A.cols, A.cols, A.nnz, A.val, A.rowptr, A.col, A.nnz, A.val, A.rowptr, A.col, rp)
for _ in range(3): bh.spgemm()
bh._lib.bhb200_synchronize(bh._ctx); print("full   ms_total", bh.stats()["ms_total"])
for _ in range(4): bh.spgemm_numeric()
bh._lib.bhb200_synchronize(bh._ctx); st = bh.stats(); print("numeric-only ms_total", st["ms_total"], "numeric", st["ms_numeric"], "launches", st["kernel_launches"])
