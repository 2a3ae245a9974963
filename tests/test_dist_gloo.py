"""Host-side logic of the multi-GPU row-block scheme on CPU: world_size-2 gloo.
Partition on the product prefix sum, broadcast of B, hand-out of A's row blocks,
all-gather of nnz(C) and assembly of the global row pointers.  The per-block
compute engine is injected: here it is the CPU oracle (a checker standing in for
the CUDA engine, which has no CPU fallback); the -m gpu tests run the real one."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from benchmark_spgemm_using_csr_b200 import generators as gen  # noqa: E402
from benchmark_spgemm_using_csr_b200.dist import (LocalResult, RowBlockSpGEMM, partition_rows_by_products,  # noqa: E402
                                                  row_products_host)


class OracleEngine:
    """Test double for CudaEngine (same four methods), CPU tensors in, oracle inside."""

    def set_operands(self, m, k, n, A, B):
        self.m, self.k, self.n, self.A, self.B = m, k, n, A, B

    def spgemm(self):
        import oracle
        a = [t.numpy() for t in self.A]
        b = [t.numpy() for t in self.B]
        self.res = oracle.spgemm(self.m, self.k, self.n, a[0], a[1], a[2], b[0], b[1], b[2])
        return int(self.res[0][-1])

    def result(self):
        return LocalResult(int(self.res[0][-1]), self.res[0], self.res[1], self.res[2])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        if case == "square":
            A = gen.rmat(9, 8, a=0.57, b=0.19, c=0.19, d=0.05, seed=2) if rank == 0 else None
            B, aeqb = A, True
        else:
            A = gen.random_csr(301, 77, np.arange(301) % 19, seed=3) if rank == 0 else None
            B = gen.random_csr(77, 130, 6, seed=4, value_seed=5) if rank == 0 else None
            aeqb = False
        rb = RowBlockSpGEMM(OracleEngine(), torch.device("cpu"))
        rb.setup_from_root(A, B, root=0, a_equals_b=aeqb)
        nnz_local, off, total = rb.spgemm()
        rowptr, col, val = rb.gather_full(off, total)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), rowptr=rowptr, col=col, val=val, bounds=rb.bounds,
                 nnz_local=nnz_local, off=off, total=total)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["square", "rect"])
def test_row_block_world2_gloo(tmp_path, case):
    import oracle
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    if case == "square":
        A = gen.rmat(9, 8, a=0.57, b=0.19, c=0.19, d=0.05, seed=2)
        B = A
    else:
        A = gen.random_csr(301, 77, np.arange(301) % 19, seed=3)
        B = gen.random_csr(77, 130, 6, seed=4, value_seed=5)
    want = oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
    r = [np.load(os.path.join(tmp_path, f"r{i}.npz")) for i in range(world)]
    for g in r:                                   # every rank holds the assembled C
        assert np.array_equal(g["rowptr"], want[0])
        assert np.array_equal(g["col"], want[1])
        assert np.array_equal(g["val"], want[2])
        assert int(g["total"]) == int(want[0][-1])
    assert int(r[0]["off"]) == 0 and int(r[1]["off"]) == int(r[0]["nnz_local"])
    # the split balances products, not rows
    prods = row_products_host(A, B.rowptr)
    b = r[0]["bounds"]
    left = prods[:b[1]].sum()
    assert abs(left - prods.sum() / 2) <= prods.max()


def test_partition_properties():
    rng = np.random.default_rng(0)
    prods = (rng.pareto(1.2, size=5000) * 10).astype(np.int64)
    for world in (1, 2, 4, 8):
        b = partition_rows_by_products(prods, world)
        assert b[0] == 0 and b[-1] == prods.size and (np.diff(b) >= 0).all()
        sums = np.array([prods[b[i]:b[i + 1]].sum() for i in range(world)])
        assert sums.sum() == prods.sum()
        assert sums.max() <= prods.sum() / world + prods.max()
    assert partition_rows_by_products(np.zeros(10, np.int64), 4).tolist()[-1] == 10
    assert partition_rows_by_products(np.zeros(0, np.int64), 2).tolist() == [0, 0, 0]


def test_row_cost_model():
    """The partition's cost of a row (dist.row_cost == csrc/dist_nccl.cu::k_row_cost): its products, times 11/8 beyond the
    largest on-chip capacity, clamped to int32 like the device array; blocks are balanced on this cost, never split a row."""
    from benchmark_spgemm_using_csr_b200.dist import COST_HEAVY_ROW, partition_rows_by_products, row_cost
    p = np.array([0, 1, 12288, 12289, 16000, 2_000_000_000], dtype=np.int64)
    c = row_cost(p)
    assert c.dtype == np.int64 and list(c[:3]) == [0, 1, 12288]
    assert c[3] == (12289 * 11) // 8 and c[4] == 22000 and c[5] == 0x7FFFFFFF
    assert np.all(np.diff(row_cost(np.arange(COST_HEAVY_ROW - 4, COST_HEAVY_ROW + 4))) >= 0)      # monotone across the threshold
    rng = np.random.default_rng(3)
    prods = rng.integers(0, 40000, size=5000)
    cost = row_cost(prods)
    b = partition_rows_by_products(cost, 8)
    assert b[0] == 0 and b[-1] == prods.size and np.all(np.diff(b) >= 0)
    share = np.add.reduceat(cost, b[:-1])[:8]
    assert share.max() - share.min() <= 2 * cost.max()      # equal shares up to one row
