#!/usr/bin/env python
"""Per-bin work and time of one SpGEMM (needs a GPU): rows, products, nnz(C), ms, products/us.
usage: python tools/bin_report.py <poisson27|poisson27thin|poisson5|rmat|rect|rmatg500> [scale] [f32|f64]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from benchmark_spgemm_using_csr_b200 import capi, generators as gen   # noqa: E402
from benchmark_spgemm_using_csr_b200.dist import CudaEngine, RowBlockSpGEMM   # noqa: E402

name = sys.argv[1]
scale = int(sys.argv[2]) if len(sys.argv) > 2 else 20
dt = np.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else np.float64
if name == "rmat":
    A = gen.rmat(scale, 16, dtype=dt); B = A
elif name == "rmatg500":
    A = gen.rmat(scale, 16, a=0.57, b=0.19, c=0.19, d=0.05, dtype=dt); B = A
elif name == "poisson27":
    A = gen.poisson27pt(128, 128, 128, dtype=dt); B = A
elif name == "poisson27thin":
    A = gen.poisson27pt(48, 48, 1024, dtype=dt); B = A
elif name == "poisson5":
    A = gen.poisson5pt(1024, 1024, dtype=dt); B = A
else:
    A = gen.uniform_rect(4194304, 1048576, 8, seed=1, dtype=np.float32)
    B = gen.uniform_rect(1048576, 4194304, 8, seed=2, value_seed=3, dtype=np.float32)
eng = CudaEngine(0)
eng.set_profiling(True)
rb = RowBlockSpGEMM(eng, torch.device("cuda", 0)).setup_from_root(A, B, a_equals_b=(A is B))
for _ in range(3):
    rb.spgemm()
st = eng.stats()
print(f"{name}: m={st['m']} nnzA={st['nnzA']} products={st['products']} nnzC={st['nnzC']} max_row_products={st['max_row_products']}")
print(f"total {st['ms_total']:.3f} ms  count {st['ms_count']:.3f}  symbolic {st['ms_symbolic']:.3f}  scan {st['ms_scan']:.3f}  numeric {st['ms_numeric']:.3f}  "
      f"GFLOPS {2 * st['products'] / st['ms_total'] / 1e6:.1f}  launches {st['kernel_launches']}")
print("symbolic bins: " + ", ".join(f"{capi.SYM_BIN_NAMES[i]}: {st['sym_bin_rows'][i]} rows {st['ms_sym_bin'][i]:.3f} ms"
                                     for i in range(len(capi.SYM_BIN_NAMES)) if st['sym_bin_rows'][i]))
print(f"{'numeric bin':12s} {'rows':>10s} {'products':>13s} {'nnzC':>13s} {'ms':>9s} {'prod/us':>9s} {'p/row':>8s}")
for i, nm in enumerate(capi.NUM_BIN_NAMES):
    r = st["num_bin_rows"][i]
    if not r:
        continue
    ms = st["ms_num_bin"][i]
    p = st["num_bin_products"][i]
    print(f"{nm:12s} {r:10d} {p:13d} {st['num_bin_nnzC'][i]:13d} {ms:9.3f} {p / ms / 1e3 if ms > 0 else 0:9.0f} {p / r:8.0f}")
