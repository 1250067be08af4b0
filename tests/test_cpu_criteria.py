"""Criteria folding (octreelib_b200/criteria.py) on the CPU: the count-step detection (a step above the probe
table used to be mistaken for `> 1024`), the rejection of coordinate-dependent criteria (translation-invariant ones
included), and the node-size guarded criteria (north_star "point-count and size thresholds") through the public API on the
oracle-backed forest stand-in, pinned by a fixture recorded from the REAL reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest

from conftest import golden
from fake_forest import FakeForest
from octreelib_b200.criteria import (NO_LEVEL_LIMIT, MaxDepth, MaxPoints, MinEdge, MinPoints, as_threshold, fold_count_criteria,
                                     fold_levels)
from octreelib_b200.grid import Grid, GridConfig


@pytest.mark.parametrize("n", [0, 1, 100, 1023, 1024, 1025, 1500, 2047, 2048, 2049, 5000, 65535, 65536, 10 ** 6, 2 ** 31 - 2])
def test_step_threshold_is_found_wherever_it_lies(n):
    crit = [lambda pts, n=n: len(pts) > n]
    table, beyond = fold_count_criteria(crit, "any", 1024)
    assert as_threshold(table, beyond, crit) == n
    # the two-criteria form of the reference's tests: any() of two steps = the lower one
    crit2 = [lambda pts, n=n: len(pts) > n, lambda pts, n=n: len(pts) > n + 77]
    table, beyond = fold_count_criteria(crit2, "any", 1024)
    assert as_threshold(table, beyond, crit2) == n


def test_step_detection_does_not_invent_thresholds():
    never = [lambda pts: False]
    t, b = fold_count_criteria(never, "any", 1024)
    assert as_threshold(t, b, never) >= 1 << 40
    band = [lambda pts: 3000 < len(pts) < 9000]  # true only on a band above the table: not a step
    t, b = fold_count_criteria(band, "any", 1024)
    assert as_threshold(t, b, band) is None
    low_band = [lambda pts: 10 < len(pts) < 20]
    t, b = fold_count_criteria(low_band, "any", 1024)
    assert as_threshold(t, b, low_band) is None
    # without the criteria themselves a step above the table cannot be located: the caller must use a table
    high = [lambda pts: len(pts) > 1500]
    t, b = fold_count_criteria(high, "any", 1024)
    assert as_threshold(t, b) is None and as_threshold(t, b, high) == 1500


@pytest.mark.parametrize("crit", [
    lambda p: len(p) > 3 and np.ptp(p, axis=0).max() > 0.5,          # extent: translation invariant
    lambda p: len(p) > 3 and p.std(axis=0).max() > 0.1,              # spread
    lambda p: len(p) > 6 and p[:, 0].mean() > 1.0,                   # position
    lambda p: len(p) > 3 and np.linalg.eigvalsh(np.cov(p.T))[0] > 1e-4,  # planarity
])
def test_coordinate_dependent_criteria_are_not_folded_but_evaluated_on_the_host(crit):
    """They never become a count table (no silent wrong answer); `subdivide` / `filter` evaluate them on the host node by
    node, like the reference, and hand the scheme / the keep-masks to the forest - checked here against the oracle, which
    evaluates the same callables the way the reference does."""
    from oracle.structure import OracleGrid

    with pytest.raises(NotImplementedError):
        fold_count_criteria([crit], "any", 64)
    cloud = np.random.default_rng(0).random((300, 3)) * np.array([7.5, 3.0, 2.0])
    grid = Grid(GridConfig(voxel_edge_length=4))
    grid._host._forest = FakeForest(4)
    grid.insert_points(0, cloud)
    og = OracleGrid(4)
    og.insert_points(0, cloud)
    grid.subdivide([crit])
    og.subdivide([crit])

    def table(leaves, corner, edge, pts):
        return [(tuple(np.asarray(corner(l), dtype=float)), float(edge(l)), np.asarray(pts(l)).tolist()) for l in leaves]

    want = table(og.get_leaf_points(0), lambda l: l.corner, lambda l: l.edge, lambda l: l.points)
    got = table(grid.get_leaf_points(0), lambda v: v.corner_min, lambda v: v.edge_length, lambda v: v.get_points())
    assert got == want and grid.n_nodes(0) == og.n_nodes(0)
    grid.filter([crit])
    og.filter([crit])
    assert grid.n_points(0) == og.n_points(0)


def test_level_limits():
    assert MaxPoints(5).level_limit(4.0) == NO_LEVEL_LIMIT
    assert MaxPoints(5, min_edge=0.5).level_limit(4.0) == 3      # edges 4, 2, 1 split; 0.5 does not
    assert MaxPoints(5, min_edge=0.4).level_limit(4.0) == 4      # 0.5 > 0.4 still splits
    assert MaxPoints(5, max_depth=2, min_edge=0.5).level_limit(4.0) == 2
    assert MaxDepth(3).level_limit(1.0) == 3 and MinEdge(1.0).level_limit(8.0) == 3
    lv = fold_levels([MaxPoints(40), MaxPoints(6, max_depth=2)], 4.0, 64)
    assert [l[0] for l in lv] == [0, 2]
    assert as_threshold(lv[0][1], lv[0][2]) == 6 and as_threshold(lv[1][1], lv[1][2]) == 40
    with pytest.raises(ValueError):
        MaxPoints(5, min_edge=0.0)
    with pytest.raises(NotImplementedError):  # a guarded criterion called outside a subdivision has no node to look at
        MaxPoints(5, min_edge=0.5)(np.zeros((9, 3)))


def _check_case(grid, g, tag):
    for p in (0, 1):
        vox = grid.get_leaf_points(p)
        assert (np.array([np.asarray(v.corner_min, dtype=np.float64) for v in vox]).reshape(-1, 3) == g[f"{tag}_p{p}_corner"]).all()
        assert (np.array([float(v.edge_length) for v in vox]) == g[f"{tag}_p{p}_edge"]).all()
        assert (np.array([v.n_points for v in vox], dtype=np.int64) == g[f"{tag}_p{p}_size"]).all()
        pts = np.vstack([np.empty((0, 3))] + [v.get_points() for v in vox])
        assert (pts == g[f"cloud{p}"][g[f"{tag}_p{p}_idx"]]).all()
        assert [grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)] == g[f"{tag}_p{p}_counts"].tolist()


SIZE_CASES = {"a": lambda: [MaxPoints(6, min_edge=0.5)], "b": lambda: [MaxPoints(40), MaxPoints(6, max_depth=2)],
              "c": lambda: [MaxDepth(2)]}


@pytest.mark.parametrize("tag", list(SIZE_CASES))
def test_size_guarded_criteria_match_the_reference(tag):
    g = golden("size_limit_edge4")
    grid = Grid(GridConfig(voxel_edge_length=4))
    grid._host._forest = FakeForest(4)
    for p in (0, 1):
        grid.insert_points(p, g[f"cloud{p}"])
    grid.subdivide(SIZE_CASES[tag]())
    _check_case(grid, g, tag)


def test_size_guarded_criteria_with_a_table_rule():
    """a non-step count criterion next to a guarded one goes through the per-level TABLE form"""
    g = golden("size_limit_edge4")
    grid = Grid(GridConfig(voxel_edge_length=4))
    grid._host._forest = FakeForest(4)
    for p in (0, 1):
        grid.insert_points(p, g[f"cloud{p}"])
    # same decisions as case "b" (no node of the fixture holds exactly 977 points), but not a step in the count any more
    grid.subdivide([lambda pts: len(pts) > 40 and len(pts) != 977, MaxPoints(6, max_depth=2)])
    _check_case(grid, g, "b")


def test_filter_rejects_size_guards():
    grid = Grid(GridConfig(voxel_edge_length=4))
    grid._host._forest = FakeForest(4)
    grid.insert_points(0, np.random.default_rng(0).random((50, 3)))
    with pytest.raises(NotImplementedError):
        grid.filter([MaxPoints(3, min_edge=1.0)])
    grid.filter([MinPoints(1)])


def test_count_thresholds_skip_the_probing():
    """`any(len(points) > n_i)` goes to the device as `count > min(n_i)` without a single probe call."""
    from octreelib_b200._host import ForestHost

    calls = []

    class RecordingForest:
        def subdivide(self, threshold, idx):
            calls.append((threshold, idx))

    host = ForestHost(1.0, (0.0, 0.0, 0.0), single_cell=False)
    host.pose_numbers = [0]
    host._forest = RecordingForest()
    host.subdivide([MaxPoints(100), MaxPoints(40), MaxPoints(70)])
    assert calls == [(40, None)]


def test_folded_lambdas_are_memoised_only_when_self_contained():
    from octreelib_b200 import criteria as C

    C._fold_memo.clear()
    def seventeen():  # the same lambda EXPRESSION evaluated again: a new function object with the same code
        return lambda pts: len(pts) > 17

    t1, b1 = C.fold_count_criteria([seventeen()], "any", 64)
    assert len(C._fold_memo) == 1
    t2, b2 = C.fold_count_criteria([seventeen()], "any", 64)
    assert len(C._fold_memo) == 1 and (t1 == t2).all() and b1 == b2 and C.as_threshold(t2, b2) == 17

    def make(n):
        return lambda pts: len(pts) > n

    ta, _ = C.fold_count_criteria([make(5)], "any", 64)
    tb, _ = C.fold_count_criteria([make(9)], "any", 64)   # same code, different closure value: a different entry
    assert C.as_threshold(ta, True) == 5 and C.as_threshold(tb, True) == 9

    before = len(C._fold_memo)
    C.fold_count_criteria([lambda pts: len(pts) > _MODULE_LEVEL_N], "any", 64)  # reads module state: never memoised
    assert len(C._fold_memo) == before


_MODULE_LEVEL_N = 21
