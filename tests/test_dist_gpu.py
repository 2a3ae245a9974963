"""Row-block SpGEMM through the real CUDA engine: one rank (always) and two ranks over
NCCL (when the box has >= 2 GPUs), assembled C compared with the oracle."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _cases():
    from benchmark_spgemm_using_csr_b200 import generators as gen
    A = gen.rmat(12, 16, a=0.57, b=0.19, c=0.19, d=0.05, seed=7)
    R = gen.random_csr(3001, 700, np.arange(3001) % 23, seed=3, dtype=np.float32)
    S = gen.random_csr(700, 1900, 9, seed=4, value_seed=5, dtype=np.float32)
    return {"rmat": (A, A, True), "rect_f32": (R, S, False), "poisson27": (gen.poisson27pt(20, 20, 20),) * 2 + (True,)}


def _check_against_oracle(name, got):
    import oracle
    A, B, _ = _cases()[name]
    want = oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])


def test_row_block_single_rank():
    import torch
    from benchmark_spgemm_using_csr_b200.dist import CudaEngine, RowBlockSpGEMM
    dev = torch.device("cuda", 0)
    for name, (A, B, aeqb) in _cases().items():
        eng = CudaEngine(0)
        rb = RowBlockSpGEMM(eng, dev).setup_from_root(A, B, a_equals_b=aeqb)
        nnz, off, total = rb.spgemm()
        assert off == 0 and nnz == total
        _check_against_oracle(name, rb.gather_full(off, total))
        eng.close()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from benchmark_spgemm_using_csr_b200.dist import CudaEngine, RowBlockSpGEMM
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        for name, (A, B, aeqb) in _cases().items():
            eng = CudaEngine(rank)
            rb = RowBlockSpGEMM(eng, dev)
            rb.setup_from_root(A if rank == 0 else None, B if rank == 0 else None, root=0, a_equals_b=aeqb)
            nnz, off, total = rb.spgemm()
            rp, c, v = rb.gather_full(off, total)
            np.savez(os.path.join(out_dir, f"{name}_r{rank}.npz"), rowptr=rp, col=c, val=v, off=off, nnz=nnz)
            eng.close()
    finally:
        dist.destroy_process_group()


def test_row_block_two_ranks_nccl(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    for name in _cases():
        r = [np.load(os.path.join(tmp_path, f"{name}_r{i}.npz")) for i in range(2)]
        for g in r:
            _check_against_oracle(name, (g["rowptr"], g["col"], g["val"]))
        assert int(r[0]["off"]) == 0 and int(r[1]["off"]) == int(r[0]["nnz"])
