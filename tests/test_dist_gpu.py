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


# ---- the same scheme THROUGH THE C-ABI (bhb200_dist_*, NCCL called by the library) -------------------
def _square_cases():
    from benchmark_spgemm_using_csr_b200 import generators as gen
    return {"rmat": gen.rmat(12, 16, a=0.57, b=0.19, c=0.19, d=0.05, seed=7),
            "rmat_mild_f32": gen.rmat(13, 8, dtype=np.float32),
            "poisson27": gen.poisson27pt(20, 20, 20)}


def _cabi_run(rank, world, dev, out_dir):
    """Every case through NcclRowBlockSpGEMM; this rank's block goes to disk with its global layout."""
    import torch
    from benchmark_spgemm_using_csr_b200.dist import CudaEngine, NcclRowBlockSpGEMM
    eng = CudaEngine(rank)
    rb = NcclRowBlockSpGEMM(eng, dev)
    for name, A in _square_cases().items():
        Bdev = None
        if rank == 0:
            Bdev = tuple(torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in (A.rowptr, A.col, A.val))
        rb.setup_square_from_device_root(Bdev, A.rows)
        for _ in range(2):                                    # twice: the all-gather / offsets are per step
            nnz, _, _ = rb.spgemm()
        r0, r1, off, total = rb.layout()
        grp = rb.global_rowptr().cpu().numpy()
        res = eng.result()
        np.savez(os.path.join(out_dir, f"cabi_{name}_r{rank}.npz"), r0=r0, r1=r1, off=off, total=total, nnz=nnz,
                 grp=grp, col=res.col.cpu().numpy(), val=res.val.cpu().numpy(), products=rb.meta["products"],
                 block_products=np.array(rb.meta["block_products"]))
    eng.close()


def _cabi_check(out_dir, world):
    import oracle
    from benchmark_spgemm_using_csr_b200.dist import partition_rows_by_products, row_cost, row_products_host
    for name, A in _square_cases().items():
        wrp, wcol, wval = oracle.spgemm(A.rows, A.cols, A.cols, A.rowptr, A.col, A.val, A.rowptr, A.col, A.val)
        prods = row_products_host(A, A.rowptr)
        bounds = partition_rows_by_products(row_cost(prods), world)     # the device partition must equal the host one
        for r in range(world):
            g = np.load(os.path.join(out_dir, f"cabi_{name}_r{r}.npz"))
            r0, r1, off = int(g["r0"]), int(g["r1"]), int(g["off"])
            assert (r0, r1) == (int(bounds[r]), int(bounds[r + 1])), f"{name}: partition of rank {r}"
            assert int(g["total"]) == int(wrp[-1]) and int(g["products"]) == int(prods.sum())
            assert int(g["block_products"][r]) == int(prods[r0:r1].sum())
            assert np.array_equal(g["grp"], wrp[r0:r1 + 1]), f"{name}: global row pointers of rank {r}"
            assert off == int(wrp[r0]) and int(g["nnz"]) == int(wrp[r1] - wrp[r0])
            assert np.array_equal(g["col"], wcol[wrp[r0]:wrp[r1]]) and np.array_equal(g["val"], wval[wrp[r0]:wrp[r1]])


def test_cabi_dist_single_rank(tmp_path):
    import torch
    _cabi_run(0, 1, torch.device("cuda", 0), str(tmp_path))
    _cabi_check(str(tmp_path), 1)


def _cabi_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)     # only carries the 128-byte NCCL id
    try:
        _cabi_run(rank, world, dev, out_dir)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_cabi_dist_nccl(tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    mp.spawn(_cabi_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    _cabi_check(str(tmp_path), world)
