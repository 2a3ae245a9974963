"""2-GPU probe: where does the per-step time outside the local pipeline go?"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, ".")
from benchmark_spgemm_using_csr_b200 import generators as gen
from benchmark_spgemm_using_csr_b200.dist import CudaEngine, RowBlockSpGEMM
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
A = gen.poisson27pt(128, 128, 128 * world) if rank == 0 else None
eng = CudaEngine(lr); eng.use_stream(torch.cuda.current_stream(dev).cuda_stream)
rb = RowBlockSpGEMM(eng, dev).setup_from_root(A, A, a_equals_b=True)
def timeit(fn, n=8):
    for _ in range(3): fn()
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / n * 1e3
def local_only():
    eng.spgemm()
def with_gather():
    rb.spgemm()
x = torch.zeros(1, dtype=torch.int64, device=dev); parts = [torch.zeros_like(x) for _ in range(world)]
def gather_only():
    dist.all_gather(parts, x); torch.cat(parts).cpu()
for name, fn in (("local pipeline only", local_only), ("pipeline + all_gather (RowBlockSpGEMM.spgemm)", with_gather), ("all_gather + D2H only", gather_only), ("local pipeline only", local_only)):
    t = timeit(fn)
    if rank == 0: print(f"{name:50s} {t:8.3f} ms", flush=True)
dist.destroy_process_group()
