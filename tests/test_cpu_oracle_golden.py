"""CPU: the oracle reproduces the golden vectors that were generated from the real reference
(tests/golden/make_golden.py).  No GPU, no /root/reference."""
import numpy as np
import pytest

from conftest import golden
from oracle import ransac as oransac
from oracle.structure import OracleGrid, max_points_criterion

STRUCTURE_CASES = ["ref_test_grid_gt2", "ref_test_grid_gt3", "random_3pose_edge2", "random_3pose_edge2_filter",
                   "clustered_2pose_edge4", "lidar_2pose_edge1", "indoor_1pose_edge1", "offset_poses_edge1",
                   "subset_subdivide_edge2", "far_offset_edge1"]
RANSAC_CASES = ["ransac_indoor_h128", "ransac_indoor_ppb1_h64", "ransac_lidar_h64_k3", "ransac_lidar_h1024", "ransac_far_h64", "ransac_degenerate_h64"]


def build_oracle(g):
    edge = float(g["edge"])
    edge = int(edge) if edge == int(edge) else edge
    og = OracleGrid(edge)
    poses = [int(p) for p in g["poses"]]
    for p in poses:
        og.insert_points(p, g[f"cloud{p}"])
    return og, poses


@pytest.mark.parametrize("name", STRUCTURE_CASES)
def test_structure_oracle_matches_reference_golden(name):
    g = golden(name)
    og, poses = build_oracle(g)
    for p in poses:  # before subdivision
        leaves = og.get_leaf_points(p)
        assert (np.array([len(l.idx) for l in leaves]) == g[f"pre_p{p}_size"]).all()
        assert (np.concatenate([l.idx for l in leaves]) == g[f"pre_p{p}_idx"]).all()
    sub = [int(x) for x in g["subdivide_poses"]] or None
    og.subdivide([max_points_criterion(int(g["max_points"]))], sub)
    if int(g["filter_min"]) >= 0:
        og.filter([lambda pts, n=int(g["filter_min"]): len(pts) >= n])
    for p in poses:
        leaves = og.get_leaf_points(p)
        corner = np.array([np.asarray(l.corner, dtype=np.float64) for l in leaves]).reshape(-1, 3)
        assert (corner == g[f"p{p}_corner"]).all()
        assert (np.array([float(l.edge) for l in leaves]) == g[f"p{p}_edge"]).all()
        assert (np.array([len(l.idx) for l in leaves]) == g[f"p{p}_size"]).all()
        idx = np.concatenate([l.idx for l in leaves]) if leaves else np.empty(0, dtype=np.int64)
        assert (idx == g[f"p{p}_idx"]).all()
        assert [og.n_leaves(p), og.n_points(p), og.n_nodes(p)] == g[f"p{p}_counts"].tolist()
        assert (og.get_point_indices(p) == g[f"p{p}_getpoints_idx"]).all()
        assert (np.array(og.pose_cells[p], dtype=np.int64).reshape(-1, 3) == g[f"p{p}_cells"]).all()


def test_late_pose_oracle_matches_reference_golden():
    """Poses inserted after the subdivision follow the existing scheme (octree_manager.py:161-171)."""
    g = golden("late_poses_edge2")
    late = [int(p) for p in g["late"]]
    poses = [int(p) for p in g["poses"]]
    og = OracleGrid(int(g["edge"]))
    for p in poses:
        if p not in late:
            og.insert_points(p, g[f"cloud{p}"])
    og.subdivide([max_points_criterion(int(g["max_points"]))])
    for p in late:
        og.insert_points(p, g[f"cloud{p}"])
    for p in poses:
        leaves = og.get_leaf_points(p)
        corner = np.array([np.asarray(l.corner, dtype=np.float64) for l in leaves]).reshape(-1, 3)
        assert (corner == g[f"p{p}_corner"]).all() and (np.array([float(l.edge) for l in leaves]) == g[f"p{p}_edge"]).all()
        assert (np.concatenate([l.idx for l in leaves]) == g[f"p{p}_idx"]).all()
        assert [og.n_leaves(p), og.n_points(p), og.n_nodes(p)] == g[f"p{p}_counts"].tolist()
        assert (og.get_point_indices(p) == g[f"p{p}_getpoints_idx"]).all()


def _assert_oracle_stage(og, g, poses, prefix):
    for p in poses:
        leaves = og.get_leaf_points(p)
        corner = np.array([np.asarray(l.corner, dtype=np.float64) for l in leaves]).reshape(-1, 3)
        assert (corner == g[f"{prefix}p{p}_corner"]).all(), (prefix, p)
        assert (np.array([float(l.edge) for l in leaves]) == g[f"{prefix}p{p}_edge"]).all()
        assert (np.concatenate([l.idx for l in leaves]) == g[f"{prefix}p{p}_idx"]).all()
        assert [og.n_leaves(p), og.n_points(p), og.n_nodes(p)] == g[f"{prefix}p{p}_counts"].tolist()
        assert (og.get_point_indices(p) == g[f"{prefix}p{p}_getpoints_idx"]).all()


def test_resubdivide_oracle_matches_reference_golden():
    """A second, finer subdivide: every pose octree keeps its leaf list across calls (octree_base.py:48-49,
    octree.py:183-191), so the leaf order is history dependent; a pose inserted between the calls sees the first
    scheme as one pass.  The fixture (real reference) records all three stages."""
    g = golden("resubdivide_deepen_edge4")
    late = [int(p) for p in g["late"]]
    poses = [int(p) for p in g["poses"]]
    early = [p for p in poses if p not in late]
    og = OracleGrid(int(g["edge"]))
    for p in early:
        og.insert_points(p, g[f"cloud{p}"])
    og.subdivide([max_points_criterion(int(g["first_max"]))])
    _assert_oracle_stage(og, g, early, "s1_")
    for p in late:
        og.insert_points(p, g[f"cloud{p}"])
    _assert_oracle_stage(og, g, poses, "s2_")
    og.subdivide([max_points_criterion(int(g["second_max"]))])
    _assert_oracle_stage(og, g, poses, "")
    assert len(g["order_differs_from_one_shot"]) > 0  # the fixture discriminates history order from a fresh build


def test_all_leaves_oracle_matches_reference_golden():
    """get_leaf_points(non_empty=False): empty leaves included (grid.py:217-232 -> octree.py:256-263)."""
    g = golden("all_leaves_edge2")
    og = OracleGrid(int(g["edge"]))
    for p in (0, 1):
        og.insert_points(p, g[f"cloud{p}"])
    og.subdivide([max_points_criterion(int(g["max_points"]))])
    for stage in ("a", "b"):
        if stage == "b":
            og.filter([lambda pts, n=int(g["filter_min"]): len(pts) >= n])
        for p in (0, 1):
            leaves = og.get_leaf_points(p, non_empty=False)
            assert (np.array([np.asarray(l.corner, dtype=np.float64) for l in leaves]).reshape(-1, 3) == g[f"{stage}_p{p}_corner"]).all()
            assert (np.array([float(l.edge) for l in leaves]) == g[f"{stage}_p{p}_edge"]).all()
            assert (np.array([len(l.idx) for l in leaves]) == g[f"{stage}_p{p}_size"]).all()


@pytest.mark.parametrize("name", RANSAC_CASES)
def test_ransac_oracle_matches_reference_golden(name):
    g = golden(name)
    table, thr, K = g["table"], float(g["threshold"]), int(g["K"])
    assert (oransac.make_table(int(g["H"]), K, seed=int(g["seed"])) == table).all()
    for bi in range(int(g["n_batches"])):
        pts, bs = g[f"b{bi}_points"], g[f"b{bi}_block_sizes"]
        res = oransac.ransac_evaluate(pts, bs, table, thr, full=True)
        assert (res["best"] == g[f"b{bi}_best"]).all()
        assert (res["best_count"] == g[f"b{bi}_best_count"]).all()
        assert (res["plane"].view(np.uint32) == g[f"b{bi}_plane"].view(np.uint32)).all()
        assert (res["mask"] == g[f"b{bi}_mask"]).all()
        # tie-aware pin against the reference's own kernel output (CUDASIM)
        ref_mask, choice = g[f"b{bi}_ref_mask"], g[f"b{bi}_ref_choice"]
        for b, (n, s) in enumerate(zip(bs, res["block_start"])):
            if n < K:
                assert not ref_mask[s:s + n].any()
                continue
            t = int(choice[b])
            assert res["counts"][b, t] == res["counts"][b].max()
            m = oransac.mask_for_plane(pts, s, n, res["planes"][b, t], thr)
            assert (m.astype(bool) == ref_mask[s:s + n]).all()


def test_c_oracle_equals_numpy_restatement():
    rng = np.random.default_rng(5)
    sizes = rng.integers(0, 40, size=60).astype(np.int32)
    pts = rng.random((int(sizes.sum()), 3))
    pts[:, 2] = 0.2 * pts[:, 0] - 0.1 * pts[:, 1] + 0.004 * rng.standard_normal(len(pts))
    pts = pts.astype(np.float32).astype(np.float64)
    table = oransac.make_table(128, 6, seed=9)
    a = oransac.ransac_evaluate(pts, sizes, table, 0.01, full=True, threads=4)
    b = oransac.ransac_numpy(pts, sizes, table, 0.01)
    assert (a["counts"] == b["counts"]).all() and (a["best"] == b["best"]).all() and (a["mask"] == b["mask"]).all()
    assert (a["plane"].view(np.uint32) == b["plane"].view(np.uint32)).all()


def test_degenerate_hypothesis_wins():
    """norm == 0 -> plane (0,0,0,0) -> every point is an inlier (util.py:76-78, SURVEY hazard 8)."""
    pts = np.tile(np.array([[1.0, 2.0, 3.0]]), (8, 1))
    res = oransac.ransac_evaluate(pts, np.array([8], dtype=np.int32), oransac.make_table(16, 6, seed=1), 0.01, full=True)
    assert res["best"][0] == 0 and res["best_count"][0] == 8 and (res["plane"][0] == 0).all() and res["mask"].all()


def test_oracle_subdivide_as_matches_the_reference_fixture():
    """tests/golden/subdivide_as_edge16.npz (tests/golden/make_subdivide_as.py, real reference): B copies A's scheme."""
    from oracle.structure import _Tree

    g = golden("subdivide_as_edge16")
    corner, edge = g["corner"], np.float64(g["edge"])
    ta, tb = _Tree(corner, edge), _Tree(corner, edge)
    ta.insert(ta.root, np.arange(len(g["a"])), g["a"])
    ta.subdivide(ta.root, [lambda p: len(p) > int(g["max_points"])])
    tb.insert(tb.root, np.arange(len(g["b"])), g["b"])
    tb.subdivide_as(tb.root, ta.root)
    leaves = [n for n in tb.cache.values() if len(n.idx)]
    assert [len(n.idx) for n in leaves] == g["b_sizes"].tolist()
    assert (np.array([n.corner for n in leaves], dtype=np.float64) == g["b_corner"]).all()
    assert (np.vstack([n.pts for n in leaves]) == g["b_points"]).all()
    assert tb.n_nodes() == int(g["b_n_nodes"])
