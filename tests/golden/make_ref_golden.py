"""Writes tests/golden/ref_outputs.npz: outputs of THE REFERENCE ITSELF (oracle/_ref =
bhSPARSE's CUDA kernels compiled for sm_100a by oracle/build_ref.py) on the inputs of
tests/ref_cases.py.  Needs a GPU, so it runs on the B200 box:

    python -m oracle.build_ref                      # in the build container (/root/reference is here)
    gpurun -- python tests/golden/make_ref_golden.py   # writes gpurun_out/ref_outputs.npz + .log
    cp gpurun_out/ref_outputs.npz tests/golden/     # commit

SMALL cases store rowptrC/colC/valC in full (f64, and f32 where values are real); the f32 runs
of integer-valued cases and the LARGE cases store nnzC and SHA-256 digests of the three arrays.  The script also checks every reference output against the oracle and reports
(does not hide) disagreements -- the committed file is whatever the reference produced."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from oracle import ref  # noqa: E402
import ref_cases  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    store, report = {}, []
    for group, cases in (("small", ref_cases.SMALL), ("large", ref_cases.LARGE)):
        for name, build in cases:
            for dn, dt in ref_cases.DTYPES.items():
                A, B = build(dt)
                key = f"{name}.{dn}"
                try:
                    rp, col, val, ms = ref.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
                except Exception as e:                      # recorded, not hidden
                    report.append({"case": key, "error": str(e)[:500]})
                    continue
                wrp, wcol, wval = oracle.spgemm(A.rows, A.cols, B.cols, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
                same_rp = bool(np.array_equal(rp.astype(np.int64), wrp))
                same_col = bool(col.shape == wcol.shape and np.array_equal(col, wcol))
                if same_col and val.size:
                    rel = float(np.max(np.abs(val.astype(np.float64) - wval.astype(np.float64)) /
                                       np.maximum(np.abs(wval.astype(np.float64)), 1e-300)))
                else:
                    rel = 0.0 if same_col else float("nan")
                report.append({"case": key, "m": A.rows, "nnzA": A.nnz, "nnzC": int(rp[-1]), "ref_ms": ms,
                               "rowptr_equal_oracle": same_rp, "col_equal_oracle": same_col, "max_rel_vs_oracle": rel})
                store[key + ".nnzC"] = np.int64(rp[-1])
                # full arrays for the small f64 cases and for real-valued ones (tolerance compare);
                # digests elsewhere (integer values: f32 results equal the f64 ones cast to f32)
                if group == "small" and (dn == "f64" or name.endswith("_real")):
                    store[key + ".rowptrC"] = rp
                    store[key + ".colC"] = col
                    store[key + ".valC"] = val
                else:
                    store[key + ".digest"] = np.array([ref_cases.digest(rp), ref_cases.digest(col), ref_cases.digest(val)])
    np.savez_compressed(os.path.join(out_dir, "ref_outputs.npz"), **store)
    with open(os.path.join(out_dir, "ref_outputs_report.json"), "w") as f:
        json.dump(report, f, indent=1)
    bad = [r for r in report if r.get("error") or not (r["rowptr_equal_oracle"] and r["col_equal_oracle"])]
    print(json.dumps({"cases": len(report), "disagree_or_error": bad}, indent=1))


if __name__ == "__main__":
    main()
