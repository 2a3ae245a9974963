"""oracle/_ref: the reference itself (bhSPARSE, SpGEMM_cuda/), compiled for sm_100a.

TEST INFRASTRUCTURE ONLY.  `python -m oracle.build_ref` compiles the reference's own
headers (bhsparse.h, bhsparse_cuda.h, common.h, cudatimer.h) WHERE THEY LIE under
/root/reference/SpGEMM_cuda -- `-I`, nothing is copied or patched on disk -- together with
oracle/ref_entry.cu (a C-ABI harness playing main.cu's role without CUSP) into

    oracle/_ref/libbhsparse_ref_f64.so      value_type = double (common.h:31 as shipped)
    oracle/_ref/libbhsparse_ref_f32.so      value_type = float  (README.md:84-86)

The three obstacles SURVEY.md 8c lists are handled by oracle/ref_shim/ (see ref_entry.cu):
stand-ins for the two CUDA-samples headers, and `-include legacy_intrinsics.h` mapping the
pre-Volta __shfl_up to __shfl_up_sync.  The reference's build system is not run.

oracle/_ref/ is git-ignored (binaries stay out of history) but not gpurun-ignored: the .so
files travel to the GPU box, where /root/reference does not exist.  When /root/reference is
absent this recipe only reports what is already built.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/SpGEMM_cuda"
OUT_DIR = os.path.join(HERE, "_ref")
ENTRY = os.path.join(HERE, "ref_entry.cu")
SHIM = os.path.join(HERE, "ref_shim")
LIBS = {"f64": os.path.join(OUT_DIR, "libbhsparse_ref_f64.so"), "f32": os.path.join(OUT_DIR, "libbhsparse_ref_f32.so")}


def _nvcc() -> str | None:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def available() -> dict:
    """Paths of the reference libraries that exist (built here earlier, shipped to the GPU box)."""
    return {k: p for k, p in LIBS.items() if os.path.exists(p)}


def build(force: bool = False) -> dict:
    if not os.path.isdir(REF_SRC):
        return available()          # GPU box: prebuilt files only
    cc = _nvcc()
    if cc is None:
        return available()
    os.makedirs(OUT_DIR, exist_ok=True)
    deps = [ENTRY] + [os.path.join(SHIM, f) for f in os.listdir(SHIM)] + \
           [os.path.join(REF_SRC, f) for f in ("bhsparse.h", "bhsparse_cuda.h", "common.h", "cudatimer.h")]
    newest = max(os.path.getmtime(d) for d in deps)
    ccbin = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    for kind, out in LIBS.items():
        if not force and os.path.exists(out) and os.path.getmtime(out) >= newest:
            continue
        cmd = [cc] + ccbin + ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-m64", "-w", "-lineinfo",
                              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "-shared",
                              "-include", os.path.join(SHIM, "legacy_intrinsics.h"), "-I", SHIM, "-I", REF_SRC]
        if kind == "f32":
            cmd.append("-DBHREF_F32")
        cmd += [ENTRY, "-o", out]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"reference build ({kind}) failed:\n{' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return available()


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
