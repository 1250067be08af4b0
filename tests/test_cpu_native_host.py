"""CPU: the C-ABI library loads, exports every declared symbol, and its host-side numerics
(numpy-compatible floor_divide, per-point cell/Morton key) agree with numpy and the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT
from octreelib_b200 import _native
from oracle.structure import OracleGrid, max_points_criterion


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "octreelib_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)  # declarations only, no comments
    declared = set(re.findall(r"\b(ol_[a-z0-9_]+)\s*\(", header))
    declared -= {"ol_alloc_fn", "ol_free_fn"}
    assert len(declared) >= 25
    assert declared == set(_native.SIGNATURES), declared ^ set(_native.SIGNATURES)
    lib = _native.lib()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ol_abi_version() == 2


def test_floor_divide_matches_numpy():
    lib = _native.lib()
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.normal(0, 100, 4000), rng.integers(-50, 50, 500).astype(float),
                        [0.0, -0.0, 1e-300, -1e-300, 5.0, -5.0, 4.999999999999999, 1e15 + 0.5]])
    for b in [1.0, 2.0, 5.0, 0.5, 0.25, 0.3, 3.7, 1e-3, -2.0]:
        want = np.floor_divide(a, b)
        got = np.array([lib.ol_host_floor_divide(float(x), float(b)) for x in a])
        assert (want == got).all(), b


def test_floor_divide_fast_path_adversarial():
    """The kernels use q = floor(a / b) corrected by the sign of fma(-q, b, a) instead of numpy's fmod-based
    formula (csrc/common.cuh); both must give the same integer, also next to multiples of b, for quotients up
    to 2^50 and for edges that are not exactly representable."""
    lib = _native.lib()
    rng = np.random.default_rng(1)
    # (powers of two take the reciprocal path, npy_floor_divide_inv: a * (1 / b) instead of a / b)
    for b in [1.0, 0.1, 0.3, 1.0 / 3.0, 0.7, 2.5, 1e-3, 1e-6, 123.456, 2.0 ** -20, 3e7, 0.5, 0.25, 4.0, 2.0 ** 30, 2.0 ** -300]:
        k = np.concatenate([rng.integers(-10 ** 6, 10 ** 6, 3000), rng.integers(-2 ** 50, 2 ** 50, 1000), [0, 1, -1, 2, -2]])
        base = k.astype(np.float64) * b  # (rounded) multiples of b
        cand = [base, np.nextafter(base, np.inf), np.nextafter(base, -np.inf), base + b * 0.5, base * (1 + 2.0 ** -52),
                base * (1 - 2.0 ** -52), np.nextafter(np.nextafter(base, np.inf), np.inf)]
        a = np.concatenate(cand)
        a = a[np.isfinite(a)]
        want = np.floor_divide(a, b)
        got = np.array([lib.ol_host_floor_divide(float(x), float(b)) for x in a])
        bad = np.flatnonzero(want != got)
        assert len(bad) == 0, (b, a[bad[:5]], want[bad[:5]], got[bad[:5]])


def _host_key(lib, edge, corner, single, depth, p):
    q = (C.c_int64 * 3)()
    m = C.c_uint64()
    bad = C.c_int32()
    cc = (C.c_double * 3)(*corner)
    pp = (C.c_double * 3)(*p)
    assert lib.ol_host_point_key(float(edge), C.byref(cc), int(single), int(depth), C.byref(pp), C.byref(q), C.byref(m),
                                 C.byref(bad)) == 0
    return [q[0], q[1], q[2]], m.value, bad.value


@pytest.mark.parametrize("edge,offset,f32", [
    (2, (0.0, 0.0, 0.0), True),
    (1, (451234.0, 5412345.0, 120.0), False),      # UTM-like coordinates, full float64 mantissas
    (3, (-98765.0, 12345.0, -4321.0), False),      # negative cells, edge that is not a power of two
    (4, (1.0e7, -1.0e7, 0.0), False),
])
def test_point_key_matches_reference_cell_and_child_routing(edge, offset, f32):
    """cell = ((p - corner) // edge * edge).astype(int) (grid.py:72-76); Morton digits = the child ids the
    oracle's level-by-level routing produces (octree.py:73-75, 94-97).  Far from the origin `p - corner` and
    `corner + edge / 2` round, and the key code (the same source the device kernels are compiled from) has to round
    the same way."""
    lib = _native.lib()
    rng = np.random.default_rng(1)
    pts = rng.random((300, 3)) * 12 - 6 + np.asarray(offset)
    if f32:
        pts = pts.astype(np.float32).astype(np.float64)
    og = OracleGrid(edge)
    og.insert_points(0, pts)
    og.subdivide([max_points_criterion(1)])  # split until every point is alone -> deep paths
    depth = 21
    seen = 0
    for leaf in og.get_leaf_points(0):
        d = int(round(np.log2(edge / leaf.edge)))
        for i in leaf.idx:
            q, m, bad = _host_key(lib, edge, [0, 0, 0], 0, depth, pts[i])
            assert [v * edge for v in q] == list(leaf.cell)
            assert bad == depth
            # rebuild the leaf corner from the first d digits, with the same float ops
            c = np.array(leaf.cell, dtype=np.float64)
            e = float(edge)
            for lvl in range(d):
                dig = (m >> (3 * (depth - 1 - lvl))) & 7
                h = e / 2
                c = c + np.array([(dig >> 2) & 1, (dig >> 1) & 1, dig & 1]) * h
                e = h
            assert (c == np.asarray(leaf.corner, dtype=np.float64)).all() and e == leaf.edge
            seen += 1
    assert seen == len(pts)


def test_point_key_flags_points_outside_a_fixed_cell():
    lib = _native.lib()
    _, _, bad = _host_key(lib, 5, [0, 0, 0], 1, 10, [1.0, 2.0, 6.0])
    assert bad == 0
    _, m, bad = _host_key(lib, 5, [0, 0, 0], 1, 10, [1.0, 2.0, 3.0])
    assert bad == 10 and (m >> 27) & 7 == 0b001
