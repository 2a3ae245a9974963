"""Seeded random sweep through the C-ABI against the oracle: shapes, row-length
distributions (constant, heavy-tailed, mostly empty, a few dense rows) and both precisions.
Structure bit-exact, integer-valued inputs -> values exact."""
import numpy as np
import pytest

import oracle
from benchmark_spgemm_using_csr_b200 import generators as gen, spgemm
from conftest import assert_csr_equal

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["general", "default"])
def _path(request, monkeypatch):
    """Every test of this module runs twice: with the diagonal-pattern mode switched off (the
    general hash / ESC / range / bitmap kernels these tests were written for) and with the
    library's default (structured operands then take csrc/stage_pattern.cuh)."""
    if request.param == "general":
        monkeypatch.setenv("BHB200_PATTERN", "off")


def _row_lengths(rng, rows, cols, kind):
    if kind == 0:
        ln = np.full(rows, int(rng.integers(1, 12)))
    elif kind == 1:      # heavy tail
        ln = np.minimum((rng.pareto(1.2, size=rows) * 3).astype(np.int64), cols)
    elif kind == 2:      # mostly empty
        ln = np.where(rng.random(rows) < 0.8, 0, rng.integers(1, 40, size=rows))
    else:                # a few dense rows among short ones
        ln = rng.integers(0, 6, size=rows)
        ln[rng.integers(0, rows, size=max(1, rows // 200))] = min(cols, int(rng.integers(200, 3000)))
    return np.minimum(ln, cols)


@pytest.mark.parametrize("seed", range(16))
def test_random_shapes(seed):
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    dt = np.float64 if seed % 2 == 0 else np.float32
    m = int(rng.integers(1, 4000))
    k = int(rng.integers(1, 4000))
    n = int(rng.choice([1, 7, 300, 5000, 200_000, 5_000_000]))
    A = gen.random_csr(m, k, _row_lengths(rng, m, k, seed % 4), seed=seed * 3 + 1, value_seed=seed * 3 + 2, dtype=dt)
    B = gen.random_csr(k, n, _row_lengths(rng, k, n, (seed // 4) % 4), seed=seed * 5 + 1, value_seed=seed * 5 + 2, dtype=dt)
    got = spgemm(A, B)
    want = oracle.spgemm(m, k, n, A.rowptr, A.col, A.val, B.rowptr, B.col, B.val)
    assert_csr_equal(got, want, exact_values=True, rtol=0, what=f"sweep seed {seed} {m}x{k}x{n} {dt.__name__}")
