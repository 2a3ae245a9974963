"""The C++ drop-in: include/bhsparse.h (the reference's class API, header-only over
the C-ABI) and the CLI-compatible driver built on it (examples/spgemm_driver.cpp,
mirroring SpGEMM_cuda/main.cu)."""
import os
import subprocess

import pytest

from benchmark_spgemm_using_csr_b200 import build as lib_build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_driver_compiles_against_the_header():
    exe = lib_build.build_driver()
    assert os.path.exists(exe)
    hdr = open(os.path.join(ROOT, "include", "bhsparse.h")).read()
    for method in ("initPlatform", "initData", "spgemm", "warmup", "get_nnzC", "get_C", "freePlatform", "free_mem"):
        assert f" {method}(" in hdr, method           # the eight public methods of bhsparse.h:17-34


def test_driver_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    exe = lib_build.build_driver()
    r = subprocess.run([exe, "-cuda", "-spgemm", "0"], capture_output=True, text=True)
    assert r.returncode != 0 and "Found an err, code = -5" in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("dataset", ["0", "1", "2", "3", "4"])
def test_driver_stock_workloads(dataset):
    """`./spgemm -cuda -spgemm {0,1,2,3,4}` (README.md:34-81 of the reference)."""
    exe = lib_build.build_driver()
    r = subprocess.run([exe, "-cuda", "-spgemm", dataset], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout
    assert "NO PASS" not in out
    for tag in ("nnzC = ", "RowPtrC PASS!", "ColIndC/csrValC PASS!", "Gflops"):
        assert tag in out, out
    if dataset == "0":
        assert "nnzC = 6. PASS!" in out


@pytest.mark.gpu
def test_driver_matrix_market(tmp_path):
    p = tmp_path / "a.mtx"
    p.write_text("%%MatrixMarket matrix coordinate real general\n3 3 5\n1 1 2.0\n1 3 1.0\n2 2 3.0\n3 1 4.0\n3 3 5.0\n")
    exe = lib_build.build_driver()
    r = subprocess.run([exe, "-cuda", "-spgemm", str(p)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "ColIndC/csrValC PASS!" in r.stdout, r.stdout + r.stderr
