"""1-GPU probe: rank 0's share of the 2-GPU weak-scaled problem, alone on one GPU."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from benchmark_spgemm_using_csr_b200 import generators as gen
from benchmark_spgemm_using_csr_b200.dist import CudaEngine, RowBlockSpGEMM
dev = torch.device("cuda", 0)
def run(A, B, aeqb, name):
    eng = CudaEngine(0); eng.use_stream(torch.cuda.current_stream(dev).cuda_stream)
    rb = RowBlockSpGEMM(eng, dev).setup_from_root(A, B, a_equals_b=aeqb)
    for _ in range(3): eng.spgemm()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(8): eng.spgemm()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / 8 * 1e3
    st = eng.stats()
    print(f"{name:45s} {t:8.3f} ms  stages count {st['ms_count']:.2f} sym {st['ms_symbolic']:.2f} scan {st['ms_scan']:.2f} num {st['ms_numeric']:.2f}  m={st['m']} products={st['products']}", flush=True)
    eng.close()
A1 = gen.poisson27pt(128, 128, 128)
run(A1, A1, True, "128^3, A=B (N=1 workload)")
A2 = gen.poisson27pt(128, 128, 256)
half = A2.row_slice(0, A2.rows // 2)
run(half, A2, False, "first half of 128x128x256 rows, B full")
run(A2.row_slice(A2.rows // 2, A2.rows), A2, False, "second half of 128x128x256 rows, B full")
