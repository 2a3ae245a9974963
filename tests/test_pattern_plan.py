"""Host side of the diagonal-pattern mode (csrc/pattern_plan.cu) -- no GPU needed: the sumset of the
operands' diagonals, the product -> output-slot table and the accumulator layout."""
import ctypes

import numpy as np
import pytest

from benchmark_spgemm_using_csr_b200 import capi


def _probe(offsA, offsB, vsize):
    lib = capi.load()
    a = np.ascontiguousarray(offsA, dtype=np.int32)
    b = np.ascontiguousarray(offsB, dtype=np.int32)
    info = np.zeros(8, dtype=np.int32)
    pos = np.zeros(a.size * b.size, dtype=np.uint8)
    offc = np.zeros(256, dtype=np.int32)
    p = lambda x: ctypes.c_void_p(x.ctypes.data)
    assert lib.bhb200_pattern_plan_probe(p(a), a.size, p(b), b.size, vsize, p(info), p(pos), p(offc)) == 0
    return info, pos.reshape(a.size, b.size), offc


def _stencil27(n):
    return [dz * n * n + dy * n + dx for dz in (-1, 0, 1) for dy in (-1, 0, 1) for dx in (-1, 0, 1)]


@pytest.mark.parametrize("vsize,banks", [(8, 16), (4, 32)])
def test_27pt_stencil_gets_a_conflict_free_layout(vsize, banks):
    d = _stencil27(128)
    rng = np.random.default_rng(1)
    info, pos, offc = _probe(rng.permutation(d), rng.permutation(d), vsize)      # (any order in)
    assert info[0] == 1 and info[1] == 125 and info[2] == 4
    assert info[4] == info[5]                                                    # wavefronts == conflict-free minimum
    ds = sorted(d)
    sums = sorted({a + b for a in ds for b in ds})
    assert np.array_equal(offc[:125], sums)
    # the slot of product (ja, jb) depends only on the output diagonal, distinct diagonals -> distinct slots
    slot = {}
    for ja, a in enumerate(ds):
        for jb, b in enumerate(ds):
            assert slot.setdefault(a + b, pos[ja, jb]) == pos[ja, jb]
    assert len(set(slot.values())) == 125 and max(slot.values()) < info[3] <= 256
    # every lane group of every B row hits distinct banks
    for ja in range(27):
        for j0 in range(0, 27, banks):
            g = pos[ja, j0:j0 + banks] % banks
            assert len(set(g.tolist())) == g.size


def test_limits_and_irregular_sets():
    info, _, _ = _probe(range(-32, 33), [0], 8)                    # 65 diagonals in A
    assert info[0] == 0
    info, _, _ = _probe([i * i for i in range(40)], [i * i for i in range(40)], 8)   # > 256 sums
    assert info[0] == 0
    info, pos, offc = _probe([-700, -31, -2, -1, 0, 1, 5, 64, 900], [-1000, -65, -3, 0, 1, 2, 3, 4, 33, 512, 1024], 4)
    assert info[0] == 1 and info[1] == len({a + b for a in [-700, -31, -2, -1, 0, 1, 5, 64, 900]
                                            for b in [-1000, -65, -3, 0, 1, 2, 3, 4, 33, 512, 1024]})
    assert info[4] >= info[5] and len(set(pos.ravel().tolist())) <= info[1]
    info, pos, _ = _probe([-1024, -1, 0, 1, 1024], [-1024, -1, 0, 1, 1024], 8)       # 5-point stencil
    assert info[0] == 1 and info[1] == 13 and info[4] == info[5]
