#!/usr/bin/env python
"""Per-bin report (like tools/bin_report.py) for ONE row block of the config-5 style R-MAT that
bench.py runs at N GPUs: `python tools/block_report.py <N> <rank>` generates the scale-(21+log2 N) matrix on
the GPU, takes rank's block of the N-way partition and multiplies it against the whole matrix on one device."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from benchmark_spgemm_using_csr_b200 import capi, generators as gen   # noqa: E402
from benchmark_spgemm_using_csr_b200.dist import CudaEngine, COST_HEAVY_ROW   # noqa: E402

N, rank = int(sys.argv[1]), int(sys.argv[2])
scale = 21 + int(np.log2(N))
dev = torch.device("cuda", 0)
rp, col, val = gen.rmat_counter_torch(scale, 16, 0.45, 0.15, 0.15, 0.25, device=dev)
n = rp.numel() - 1
lenB = (rp[1:] - rp[:-1]).to(torch.int64)
csum = torch.zeros(col.numel() + 1, dtype=torch.int64, device=dev)
torch.cumsum(lenB[col.to(torch.int64)], 0, out=csum[1:])
prods = csum[rp[1:].to(torch.int64)] - csum[rp[:-1].to(torch.int64)]
cost = torch.where(prods > COST_HEAVY_ROW, (prods * 11) // 8, prods)
cpre = torch.zeros(n + 1, dtype=torch.int64, device=dev)
torch.cumsum(cost, 0, out=cpre[1:])
tot = int(cpre[-1])
b = [0] + [int(torch.searchsorted(cpre, torch.tensor([(tot * r) // N], device=dev))[0]) for r in range(1, N)] + [n]
r0, r1 = b[rank], b[rank + 1]
e0, e1 = int(rp[r0]), int(rp[r1])
A = ((rp[r0:r1 + 1] - rp[r0]).contiguous(), col[e0:e1], val[e0:e1])
del lenB, csum, cost, cpre
eng = CudaEngine(0)
eng.set_profiling(True)
eng.set_operands(r1 - r0, n, n, A, (rp, col, val))
for _ in range(3):
    eng.spgemm()
st = eng.stats()
print(f"rmat scale {scale}, block {rank}/{N}: rows [{r0},{r1}) products={st['products']} (of {int(prods.sum())}) nnzC={st['nnzC']} max_row_products={st['max_row_products']}")
print(f"total {st['ms_total']:.3f} ms  count {st['ms_count']:.3f}  symbolic {st['ms_symbolic']:.3f}  scan {st['ms_scan']:.3f}  numeric {st['ms_numeric']:.3f}  launches {st['kernel_launches']}")
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
