"""Build recipe for the CPU oracle (test infrastructure only).

`python -m oracle.build` compiles oracle/spgemm_oracle.c into
oracle/liboracle_spgemm.so with gcc + OpenMP.

oracle/_ref/ (the reference itself, compiled for sm_100a) is built by oracle/build_ref.py.
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "spgemm_oracle.c")
OUT = os.path.join(HERE, "liboracle_spgemm.so")

_FLAGS = ["-O3", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wextra"]


def _candidates():
    # the image's $CC wrapper (/opt/gcc) has no libgomp.spec; the system gcc does
    seen = []
    for cc in ("/usr/bin/gcc", shutil.which("gcc"), os.environ.get("CC")):
        if cc and cc not in seen and os.path.exists(cc):
            seen.append(cc)
    return seen


def build(force: bool = False) -> str:
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    errors = []
    for omp in (["-fopenmp"], []):
        for cc in _candidates():
            cmd = [cc] + _FLAGS + omp + ["-o", OUT, SRC]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode == 0:
                return OUT
            errors.append(" ".join(cmd) + "\n" + r.stderr)
    raise RuntimeError("could not build the oracle:\n" + "\n".join(errors))


if __name__ == "__main__":
    print(build(force=True))
