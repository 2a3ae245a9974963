#!/usr/bin/env python
"""bench.py -- SpGEMM GFLOPS (2 x intermediate products / time) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload poisson27|poisson5|rmat|rect] [--dtype f64|f32]

A "step" is one full C = A*B: all four stages of bhsparse::spgemm
(SpGEMM_cuda/bhsparse.h:297-339) -- upper bound + binning, symbolic, row-pointer
scan + allocation, numeric -- from device-resident A, B to device-resident C,
the span the reference times (bhsparse.h:268-289).

Workload at N=1: BASELINE.json configs[1], Poisson 27-point 128^3, C = A^2,
double.  At N>1 (weak scaling): the same stencil on a 128 x 128 x (128*N) grid,
A split into N row blocks on the prefix sum of the per-row products, B broadcast
once with NCCL in the setup (reported as setup_broadcast_ms), each rank running
the single-GPU pipeline on its block plus the all-gather of nnz(C) offsets.

`--impl reference`: the reference has no CPU SpGEMM (ref_spgemm.h calls CUSP on
the device), so this arm times the CPU oracle (oracle/, a row-wise Gustavson
restatement, OpenMP over all host cores) on the same workload -- kind "port".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "spgemm_gflops_2x_products_per_s"
UNIT = "GFLOPS"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(name: str, n_gpus: int, dtype):
    """Returns (A, B, a_equals_b, description) as host CSR (root only)."""
    from benchmark_spgemm_using_csr_b200 import generators as gen
    if name == "poisson27":
        nz = 128 * n_gpus
        A = gen.poisson27pt(128, 128, nz, dtype=dtype)
        return A, A, True, f"Poisson27pt 3D 128x128x{nz} C=A^2"
    if name == "poisson27thin":          # same stencil, narrow column span (48 x 48 planes)
        A = gen.poisson27pt(48, 48, 1024 * n_gpus, dtype=dtype)
        return A, A, True, f"Poisson27pt 3D 48x48x{1024 * n_gpus} C=A^2"
    if name == "poisson5":
        A = gen.poisson5pt(1024, 1024 * n_gpus, dtype=dtype)
        return A, A, True, f"Poisson5pt 2D 1024x{1024 * n_gpus} C=A^2"
    if name == "rmat":
        scale = 20 + int(np.log2(n_gpus))
        A = gen.rmat(scale, 16, dtype=dtype)
        return A, A, True, f"R-MAT scale-{scale} ef16 (.45,.15,.15,.25) C=A^2"
    if name == "rect":
        A = gen.uniform_rect(4194304 * n_gpus, 1048576, per_row=8, seed=1, dtype=dtype)
        B = gen.uniform_rect(1048576, 4194304, per_row=8, seed=2, value_seed=3, dtype=dtype)
        return A, B, False, f"uniform rect A({4194304 * n_gpus}x1M,8/row)*B(1Mx4M,8/row)"
    raise ValueError(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(args, rank):
    """CPU arm: the oracle (port of the reference's result definition) on host cores."""
    if rank != 0:
        return
    import oracle
    dtype = np.float64 if args.dtype == "f64" else np.float32
    A, B, _, desc = make_workload(args.workload, 1, dtype)     # one rank's share of the weak-scaled problem
    threads = oracle.num_threads()
    _, P = oracle.row_products(A.rows, A.rowptr, A.col, B.rowptr)
    times = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rp, _, _ = oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
    t = float(np.mean(times))
    val = 2.0 * P / t / 1e9
    sample = f"full {desc} ({A.rows} rows, {P} products) per step" + (
        "" if args.gpus == 1 else f"; one rank's share of the {args.gpus}-GPU weak-scaled problem")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": desc, "values": "integers 1..9 (fixed seed)", "nnzC": int(rp[-1])},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="poisson27", choices=["poisson27", "poisson27thin", "poisson5", "rmat", "rect"])
    ap.add_argument("--dtype", default=None, choices=["f64", "f32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.dtype is None:
        args.dtype = "f32" if args.workload == "rect" else "f64"
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from benchmark_spgemm_using_csr_b200 import BHSPARSE_CUDA, NUM_PLATFORMS, bhsparse
    from benchmark_spgemm_using_csr_b200.dist import CudaEngine, RowBlockSpGEMM

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    np_dtype = np.float64 if args.dtype == "f64" else np.float32
    vsz = 8 if args.dtype == "f64" else 4

    # ---- setup (untimed): generate on rank 0, partition, broadcast B over NCCL ----
    A = B = None
    desc = ""
    aeqb = True
    if rank == 0:
        A, B, aeqb, desc = make_workload(args.workload, world, np_dtype)
    if world > 1:
        box = [aeqb, desc]
        dist.broadcast_object_list(box, src=0)
        aeqb, desc = box
    # Everything timed runs on ONE explicit non-default stream: the library's kernels, the
    # torch CUDA events and the NCCL all-gather.  (The default stream's handle is 0, which
    # bhb200_set_stream reads as "use the context's own stream": events on the default stream
    # would then not bracket the kernels.)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    engine = CudaEngine(local_rank)
    engine.use_stream(stream.cuda_stream)
    assert stream.cuda_stream != 0
    rb = RowBlockSpGEMM(engine, dev)
    rb.setup_from_root(A, B, root=0, a_equals_b=aeqb)
    meta = rb.meta
    P_total = meta["products"]
    engine.set_profiling(True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- warm-up ----
    for _ in range(args.warmup):
        rb.spgemm()
    barrier()

    # ---- timed region: K steps, CUDA events on the launching stream ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    from benchmark_spgemm_using_csr_b200.capi import NUM_BINS
    bin_ms_sym = np.zeros(NUM_BINS)
    bin_ms_num = np.zeros(NUM_BINS)
    launches = 0
    barrier()
    wall0 = time.perf_counter()
    ev0.record(stream)
    for _ in range(args.steps):
        nnz_local, off, nnz_total = rb.spgemm()
    ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - wall0) * 1e3 / args.steps
    ms = ev0.elapsed_time(ev1) / args.steps
    # the step has host syncs inside, so device time and wall clock must agree
    if abs(ms - wall_ms) > 0.05 * wall_ms + 0.05:
        raise SystemExit(f"timing inconsistency: CUDA events {ms:.3f} ms/step vs wall clock {wall_ms:.3f} ms/step")
    clocks = sampler.stop() if rank == 0 else None
    st = engine.stats()                                       # last step's per-stage / per-bin times
    launches = st["kernel_launches"] * args.steps
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt)
        launches = int(lt.item())
    value = 2.0 * P_total / (ms * 1e-3) / 1e9

    # ---- per-bin kernel times (profiling events, averaged over a few extra steps) ----
    reps = 3
    for _ in range(reps):
        rb.spgemm()
        s2 = engine.stats()
        bin_ms_sym += np.array(s2["ms_sym_bin"])
        bin_ms_num += np.array(s2["ms_num_bin"])
    bin_ms_sym /= reps
    bin_ms_num /= reps
    # NOTE: the same events are recorded during the timed steps (profiling stays on), so the
    # per-bin times come from the identical launch configuration.

    # ---- e2e: the reference-facing API with HOST buffers, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        a_rp, a_c, a_v = (t.cpu().pin_memory() for t in rb.A)
        b_rp, b_c, b_v = (t.cpu().pin_memory() for t in rb.B)
        m_loc = a_rp.numel() - 1
        rowptrC = torch.empty(m_loc + 1, dtype=torch.int32).pin_memory()
        colC = torch.empty(max(nnz_local, 1), dtype=torch.int32).pin_memory()
        valC = torch.empty(max(nnz_local, 1), dtype=b_v.dtype).pin_memory()
        platforms = [False] * NUM_PLATFORMS
        platforms[BHSPARSE_CUDA] = True
        bh = bhsparse(local_rank)
        assert bh.initPlatform(platforms) == 0

        def e2e_step():
            err = bh.initData(m_loc, meta["k"], meta["n"], a_c.numel(), a_v.numpy(), a_rp.numpy(), a_c.numpy(),
                              b_c.numel(), b_v.numpy(), b_rp.numpy(), b_c.numpy(), rowptrC.numpy())
            err |= bh.spgemm()
            n = bh.get_nnzC()
            err |= bh.get_C(colC.numpy()[:n], valC.numpy()[:n])
            assert err == 0, bh.last_error()
            return n

        e_steps = max(3, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize(dev)
        e_ms = (time.perf_counter() - t0) * 1e3 / e_steps
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        h2d = sum(t.numel() * t.element_size() for t in (a_rp, a_c, a_v, b_rp, b_c, b_v))
        d2h = rowptrC.numel() * 4 + nnz_local * (4 + vsz)
        if world > 1:
            t = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
            dist.all_reduce(t)
            h2d, d2h = int(t[0].item()), int(t[1].item())
        e2e = {"value": 2.0 * P_total / (e_ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "api": "bhsparse.initData + spgemm + get_C, pinned host buffers"}
        bh.free_mem()
        bh.freePlatform()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0's block) ----
    from benchmark_spgemm_using_csr_b200.capi import NUM_BIN_NAMES, SYM_BIN_NAMES
    peak, peak_src = peaks()
    # Dominant kernel: either an ordinary numeric bin, or the direct-mode numeric kernels (they run in
    # stage 2 in place of the symbolic pass; their rows come back as the numeric bin "copy_ct").
    # Algorithmic bytes of a launch = stream-gather model restricted to its rows:
    # read the A rows + two rowptrB words per A entry + one (col,val) of B per product + write the C rows.
    def alg_bytes(bin_idx):
        return (st["num_bin_rows"][bin_idx] * 8 + st["num_bin_nnzA"][bin_idx] * (4 + vsz) + st["num_bin_nnzA"][bin_idx] * 8 +
                st["num_bin_products"][bin_idx] * (4 + vsz) + st["num_bin_nnzC"][bin_idx] * (4 + vsz))
    direct_bins = [i for i in range(len(SYM_BIN_NAMES)) if (st["direct_bin_mask"] >> i) & 1]
    direct_ms = float(sum(bin_ms_sym[i] for i in direct_bins))
    copy_idx = NUM_BIN_NAMES.index("copy_ct")
    num_ms = bin_ms_num.copy()
    num_ms[copy_idx] = 0.0
    b = int(np.argmax(num_ms))
    if direct_ms > num_ms[b]:
        kname = "k_num_direct (direct mode, symbolic bins " + ",".join(SYM_BIN_NAMES[i] for i in direct_bins) + ")"
        rows_b, bytes_b, t_b, tkey = st["num_bin_rows"][copy_idx], alg_bytes(copy_idx), direct_ms * 1e-3, "direct"
    else:
        kname = f"k_num_* bin {NUM_BIN_NAMES[b]}"
        rows_b, bytes_b, t_b, tkey = st["num_bin_rows"][b], alg_bytes(b), num_ms[b] * 1e-3, "num_" + NUM_BIN_NAMES[b]
    achieved = bytes_b / t_b / 1e9 if t_b > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"{args.workload}_{args.dtype}_{tkey}")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "kernel": f"{kname} ({rows_b} rows)", "kernel_ms": t_b * 1e3,
                "algorithmic_bytes": int(bytes_b), "peak_source": peak_src}
    step_alg = st["bytes_algorithmic"] if world == 1 else None
    roof_step = None
    if step_alg:
        roof_step = {"algorithmic_bytes": int(step_alg), "achieved": step_alg / (ms * 1e-3) / 1e9,
                     "frac": step_alg / (ms * 1e-3) / 1e9 / peak,
                     "compulsory_bytes": int(st["bytes_compulsory"]),
                     "frac_compulsory": st["bytes_compulsory"] / (ms * 1e-3) / 1e9 / peak}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        t_best = None
        for _ in range(2):
            t0 = time.perf_counter()
            oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
            dt = time.perf_counter() - t0
            t_best = dt if t_best is None else min(t_best, dt)
        cpu = {"value": 2.0 * P_total / t_best / 1e9, "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
               "sample": f"full workload ({A.rows} rows, {P_total} products), best of 2, {t_best:.3f} s"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "wall_ms_per_step": wall_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": desc, "values": "integers 1..9 (fixed seed)", "m": meta["m"], "nnzA": meta["nnzA"],
                   "products": P_total, "nnzC_rank0": int(nnz_local), "nnzC": int(nnz_total),
                   "l2": "inputs larger than L2 (no flush needed)",
                   "partition": f"{world} row block(s) on the prefix sum of per-row products"},
        "roofline": roofline, "roofline_step": roof_step, "cpu_baseline": cpu, "e2e": e2e,
        "gpu_launches": int(launches), "clocks": clocks,
        "stages_ms": {"count_bin": st["ms_count"], "symbolic": st["ms_symbolic"], "scan_alloc": st["ms_scan"],
                      "numeric": st["ms_numeric"], "total": st["ms_total"]},
        "bins_ms": {"symbolic": {SYM_BIN_NAMES[i]: round(float(bin_ms_sym[i]), 4) for i in range(len(SYM_BIN_NAMES)) if bin_ms_sym[i] > 0},
                    "numeric": {NUM_BIN_NAMES[i]: round(float(bin_ms_num[i]), 4) for i in range(len(NUM_BIN_NAMES)) if bin_ms_num[i] > 0}},
        "direct_mode": {"rows": st["direct_rows"], "retry_rows": st["direct_retry_rows"], "staging_bytes": st["direct_ct_bytes"]},
        "setup_broadcast_ms": rb.timings.get("broadcast_B_s", 0.0) * 1e3,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
