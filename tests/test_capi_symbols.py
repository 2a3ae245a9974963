"""CPU-side checks of the drop-in boundary: the library loads and exports every
symbol include/bhsparse_b200.h declares; without a GPU the entry points fail
loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

from benchmark_spgemm_using_csr_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "bhsparse_b200.h")).read()
    return sorted(set(re.findall(r"BHB200_API[^;(]*?\b(bhb200_\w+)\s*\(", text)))


def test_header_and_binding_agree():
    assert _declared() == sorted(capi.EXPORTED)


def test_library_exports_every_declared_symbol():
    lib = capi.load(build_if_missing=True)
    for name in _declared():
        assert hasattr(lib, name), f"{name} missing from {capi.LIB_PATH}"
    assert b"sm_100a" in lib.bhb200_version()


def test_stats_struct_matches_header():
    text = open(os.path.join(ROOT, "include", "bhsparse_b200.h")).read()
    body = re.search(r"typedef struct bhb200_stats \{(.*?)\} bhb200_stats;", text, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(None, 1)[1].split(","):
            names.append(re.sub(r"\[.*\]", "", part).strip())
    assert names == [f[0] for f in capi.Stats._fields_]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load(build_if_missing=True)
    ctx = ctypes.c_void_p(None)
    assert lib.bhb200_create(ctypes.byref(ctx), 0) == capi.ERR_NO_DEVICE
    assert not ctx
    from benchmark_spgemm_using_csr_b200 import generators as gen, spgemm
    A = gen.poisson5pt(4, 4)
    with pytest.raises(capi.BhsparseError):
        spgemm(A, A)
