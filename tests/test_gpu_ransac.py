"""GPU: per-leaf RANSAC parity.  Inlier counts, chosen hypothesis (lowest index among the maxima),
plane (bit-exact float32; the stated tolerance is 1e-5 relative) and masks against the C oracle and
the golden vectors recorded from the reference's own kernel (tie-aware)."""
import numpy as np
import pytest

from conftest import golden
from octreelib_b200 import _native as N
from octreelib_b200.criteria import MaxPoints
from octreelib_b200.grid import Grid, GridConfig
from octreelib_b200.ransac import CudaRansac
from octreelib_b200.synthetic import indoor_scene, lidar64_scan
from oracle import ransac as oransac
from oracle.structure import OracleGrid, max_points_criterion

pytestmark = pytest.mark.gpu

RANSAC_CASES = ["ransac_indoor_h128", "ransac_indoor_ppb1_h64", "ransac_lidar_h64_k3", "ransac_lidar_h1024", "ransac_far_h64", "ransac_degenerate_h64"]


def _check_against_oracle(pts, bs, table, thr, r: CudaRansac, mask):
    ora = oransac.ransac_evaluate(pts, bs, table, thr, threads=8)
    assert (r.last_best == ora["best"]).all(), "chosen hypothesis differs"
    assert (r.last_best_count == ora["best_count"]).all(), "inlier counts differ"
    assert (r.last_planes.view(np.uint32) == ora["plane"].view(np.uint32)).all(), "planes differ"
    rel = np.abs(r.last_planes - ora["plane"]).max(initial=0) / max(np.abs(ora["plane"]).max(initial=0), 1e-30)
    assert rel <= 1e-5  # north_star tolerance (met with zero error)
    assert (mask == ora["mask"].astype(bool)).all(), "masks differ"


ALL_MODES = [0, N.RANSAC_FLAG_NO_TMA, N.RANSAC_FLAG_EXACT_ONLY, N.RANSAC_FLAG_VERIFY]


@pytest.mark.parametrize("flags", ALL_MODES)
@pytest.mark.parametrize("name", RANSAC_CASES)
def test_evaluate_golden(name, flags):
    g = golden(name)
    H, K, thr = int(g["H"]), int(g["K"]), float(g["threshold"])
    np.random.seed(int(g["seed"]))
    r = CudaRansac(threshold=thr, hypotheses_number=H, initial_points_number=K)
    r.flags = flags
    assert (r.random_hypotheses == g["table"]).all()
    for bi in range(int(g["n_batches"])):
        pts, bs = g[f"b{bi}_points"], g[f"b{bi}_block_sizes"]
        mask = r.evaluate(pts, bs)
        assert mask.dtype == np.bool_ and mask.shape == (len(pts),)
        assert (r.last_best == g[f"b{bi}_best"]).all()
        assert (r.last_best_count == g[f"b{bi}_best_count"]).all()
        assert (r.last_planes.view(np.uint32) == g[f"b{bi}_plane"].view(np.uint32)).all()
        assert (mask == g[f"b{bi}_mask"].astype(bool)).all()
        # against the reference kernel's own output: identical wherever the reference's race picked
        # the lowest-index maximum, and always the same inlier COUNT per block
        ref_mask, choice = g[f"b{bi}_ref_mask"], g[f"b{bi}_ref_choice"]
        starts = np.concatenate([[0], np.cumsum(bs)[:-1]])
        for b, (n, s) in enumerate(zip(bs, starts)):
            assert ref_mask[s:s + n].sum() == mask[s:s + n].sum()
            if choice[b] == r.last_best[b]:
                assert (ref_mask[s:s + n] == mask[s:s + n]).all()


@pytest.mark.parametrize("seed,nblocks,maxn,H,K,thr", [(0, 400, 120, 1024, 6, 0.02), (1, 1500, 40, 256, 6, 0.01),
                                                        (2, 50, 3000, 512, 4, 0.015), (3, 300, 12, 100, 9, 0.03)])
@pytest.mark.parametrize("flags", [0, N.RANSAC_FLAG_VERIFY])
def test_evaluate_random_blocks(seed, nblocks, maxn, H, K, thr, flags):
    rng = np.random.default_rng(seed)
    bs = rng.integers(0, maxn + 1, size=nblocks).astype(np.int32)
    n = int(bs.sum())
    pts = rng.random((n, 3)) * 2 + rng.integers(-40, 40, (n, 1))
    # piecewise planar + noise so that inliers exist and ties are frequent
    pts[:, 2] = 0.1 * pts[:, 0] - 0.05 * pts[:, 1] + rng.normal(0, 0.01, n) + (rng.random(n) < 0.2) * rng.random(n)
    pts = pts.astype(np.float32).astype(np.float64)
    np.random.seed(seed + 100)
    r = CudaRansac(threshold=thr, hypotheses_number=H, initial_points_number=K)
    r.flags = flags
    mask = r.evaluate(pts, bs)
    _check_against_oracle(pts, bs, r.random_hypotheses, thr, r, mask)


def _adversarial_blocks(rng):
    """Blocks built to stress the FP32 pre-filter's error bounds: collinear and duplicated samples,
    exactly planar data, tiny and huge extents, coordinates far from the origin, points sitting on the
    threshold, scan-ring like lines."""
    blocks = []

    def add(p):
        blocks.append(np.asarray(p, dtype=np.float64).reshape(-1, 3))

    for n in (6, 7, 8, 12, 40):
        t = rng.random(n)
        add(np.c_[t, 2 * t, -t] + [5.0, -3.0, 1.0])                                   # exactly collinear
        add(np.c_[t, 2 * t, -t] + rng.normal(0, 1e-7, (n, 3)) + [5.0, -3.0, 1.0])     # almost collinear
        add(np.repeat(rng.random((2, 3)), (n + 1) // 2, axis=0)[:n])                  # two distinct points
        add(np.repeat(rng.random((3, 3)), (n + 2) // 3, axis=0)[:n] * 3)              # three distinct points
        xy = rng.random((n, 2))
        add(np.c_[xy, np.zeros(n)])                                                   # exactly planar, axis aligned
        add(np.c_[xy, 0.3 * xy[:, 0] - 0.7 * xy[:, 1]] + [1000.0, -2000.0, 50.0])     # planar, far from the origin
        add(np.c_[xy, rng.normal(0, 0.01, n)] * 1e-4 + [812.0, 9.5, 0.1])             # tiny extent, far away
        add(np.c_[xy, rng.normal(0, 0.01, n)] * 1e3)                                  # huge extent
        add(np.c_[xy, rng.choice([0.0, 0.01, -0.01, 0.02], n)] + [100.0, 0.0, 0.0])   # distances exactly on thresholds
        rings = np.repeat(np.arange(3), (n + 2) // 3)[:n] * 0.07                       # LiDAR-ring like rows
        add(np.c_[np.arange(n) * 0.033, rings, rng.normal(0, 0.02, n)] + [30.0, 4.0, 0.0])
        add(np.c_[xy * 0.1, rng.normal(0, 0.02, n)] + [1e5, 1e5, 10.0])               # float32 ulp ~ 8 mm of the coordinates
        add(rng.normal(0, 1.0, (n, 3)) * [1.0, 1.0, 1e-3])                            # generic noisy plane
        add(np.zeros((n, 3)))                                                         # all points at the origin
    return blocks


@pytest.mark.parametrize("thr", [0.01, 0.02, 1e-4])
@pytest.mark.parametrize("H,K", [(1024, 6), (300, 3), (64, 9)])
def test_prefilter_bounds_adversarial(H, K, thr):
    """Filtered, exact-only and verify modes give the oracle's answer on inputs that stress the bounds;
    verify mode additionally proves lo <= exact count <= hi for EVERY hypothesis (it raises otherwise)."""
    rng = np.random.default_rng(1234)
    blocks = _adversarial_blocks(rng)
    pts = np.vstack(blocks)
    pts32 = pts.astype(np.float32).astype(np.float64)
    bs = np.array([len(b) for b in blocks], dtype=np.int32)
    for cloud in (pts, pts32):
        ora = None
        for flags in (0, N.RANSAC_FLAG_VERIFY, N.RANSAC_FLAG_EXACT_ONLY):
            np.random.seed(77)
            r = CudaRansac(threshold=thr, hypotheses_number=H, initial_points_number=K)
            r.flags = flags
            mask = r.evaluate(cloud, bs)
            if ora is None:
                ora = oransac.ransac_evaluate(cloud, bs, r.random_hypotheses, thr, threads=8)
            assert (r.last_best == ora["best"]).all(), f"flags {flags}: chosen hypothesis differs"
            assert (r.last_best_count == ora["best_count"]).all()
            assert (r.last_planes.view(np.uint32) == ora["plane"].view(np.uint32)).all()
            assert (mask == ora["mask"].astype(bool)).all()


def test_prefilter_statistics():
    """The pre-filter must actually prune: on a LiDAR-like leaf set only a small share of the hypotheses
    needs the exact float64 evaluation."""
    import ctypes as C
    lib = N.lib()
    out = (C.c_uint64 * 16)()
    N.check(lib.ol_ransac_stats_read(C.byref(out), 1))
    clouds = lidar64_scan(0, seed=0)[::3]
    grid = Grid(GridConfig(voxel_edge_length=1.0))
    grid.insert_points(0, clouds)
    grid.subdivide([MaxPoints(100)])
    np.random.seed(3)
    table = oransac.make_table(1024, 6)
    host = grid._host
    host.forest.ransac(table, 0.02, [0], 10, apply=False, flags=N.RANSAC_FLAG_STATS)
    N.check(lib.ol_ransac_stats_read(C.byref(out), 1))
    blocks, filtered, trivial, exact = out[0], out[1], out[2], out[3]
    assert blocks > 0 and filtered > 0
    assert exact < 0.25 * filtered, (blocks, filtered, trivial, exact)
    res_a = host.forest.export_ransac()
    host.forest.ransac(table, 0.02, [0], 10, apply=False, flags=N.RANSAC_FLAG_VERIFY)
    res_b = host.forest.export_ransac()
    for k in ("best", "best_count", "size"):
        assert (res_a[k] == res_b[k]).all()
    assert (res_a["plane"].view(np.uint32) == res_b["plane"].view(np.uint32)).all()


def test_evaluate_edge_cases():
    np.random.seed(0)
    r = CudaRansac(threshold=0.01, hypotheses_number=64, initial_points_number=6)
    # no blocks / only blocks that are too small
    assert r.evaluate(np.empty((0, 3)), np.empty(0, dtype=np.int32)).shape == (0,)
    m = r.evaluate(np.random.rand(5, 3), np.array([5], dtype=np.int32))
    assert not m.any() and r.last_best[0] == -1
    # coincident points: the degenerate (0,0,0,0) plane wins and keeps everything (SURVEY hazard 8)
    pts = np.tile(np.array([[1.0, 2.0, 3.0]]), (9, 1))
    m = r.evaluate(pts, np.array([9], dtype=np.int32))
    assert m.all() and (r.last_planes[0] == 0).all() and r.last_best[0] == 0
    # one huge block (larger than the shared-memory staging area) next to small ones
    rng = np.random.default_rng(4)
    big = rng.random((9000, 3))
    big[:, 2] = 0.02 * rng.standard_normal(9000)
    small = rng.random((30, 3))
    pts = np.vstack([small, big, small]).astype(np.float32).astype(np.float64)
    bs = np.array([30, 9000, 30], dtype=np.int32)
    mask = r.evaluate(pts, bs)
    _check_against_oracle(pts, bs, r.random_hypotheses, 0.01, r, mask)
    with pytest.raises(ValueError):
        r.evaluate(pts, np.array([30, 30], dtype=np.int32))


@pytest.mark.parametrize("ppb", [10, 1, 2])
def test_grid_ransac_matches_oracle(ppb):
    clouds = {p: indoor_scene(6000, seed=10 + p) for p in range(3)}
    grid, og = Grid(GridConfig(voxel_edge_length=2)), OracleGrid(2)
    for p, c in clouds.items():
        grid.insert_points(p, c)
        og.insert_points(p, c)
    grid.subdivide([MaxPoints(150)])
    og.subdivide([max_points_criterion(150)])
    np.random.seed(42)
    table = oransac.make_table(256, 6)
    np.random.seed(42)
    # run without applying first to compare the per-block tables
    host = grid._host
    host.forest.ransac(table, 0.02, [int(x) for x in host.pose_numbers], ppb, apply=False)
    res = host.forest.export_ransac()
    ora_batches = og.map_leaf_points_ransac(table, threshold=0.02, poses_per_batch=ppb)
    o_best = np.concatenate([b["best"] for b in ora_batches])
    o_cnt = np.concatenate([b["best_count"] for b in ora_batches])
    o_plane = np.vstack([b["plane"] for b in ora_batches])
    o_size = np.concatenate([b["block_sizes"] for b in ora_batches])
    o_mask = np.concatenate([b["mask"] for b in ora_batches])
    assert (res["size"] == o_size).all()
    assert (res["best"] == o_best).all() and (res["best_count"] == o_cnt).all()
    assert (res["plane"].view(np.uint32) == o_plane.view(np.uint32)).all()
    pts = host.forest.export_points(-1, order=0, pose_rank=[int(x) for x in host.pose_numbers], want_mask=True)
    assert (pts["mask"] == o_mask).all()
    host.forest.apply_mask()
    host._counts_cache = None
    for p in clouds:
        got = host.forest.export_points(host.pose_index[p], order=0)
        want = np.concatenate([l.idx for l in og.get_leaf_points(p)])
        assert (got["idx"] == want).all()
        assert [grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)] == [og.n_leaves(p), og.n_points(p), og.n_nodes(p)]


def test_grid_ransac_public_api_and_pose_numbering():
    clouds = {p: lidar64_scan(p, seed=1)[::8] for p in range(2)}
    grid, og = Grid(GridConfig(voxel_edge_length=1.0)), OracleGrid(1.0)
    for p in (1, 0):  # insertion order != pose number order
        grid.insert_points(p, clouds[p])
        og.insert_points(p, clouds[p])
    grid.subdivide([lambda pts: len(pts) > 80])
    og.subdivide([max_points_criterion(80)])
    np.random.seed(5)
    table = oransac.make_table(128, 6)
    np.random.seed(5)
    grid.map_leaf_points_cuda_ransac(poses_per_batch=10, threshold=0.03, hypotheses_number=128)
    # the oracle follows the reference's batch layout: pose numbers 0, 1 (grid.py:149-157)
    og2 = og
    batches = [[0, 1]]
    clouds_b, sizes_b = [], []
    for p in batches[0]:
        leaves = og2.get_leaf_points(p)
        clouds_b.append(np.vstack([l.points for l in leaves]))
        sizes_b.append(np.array([len(l.points) for l in leaves], dtype=np.int32))
    res = oransac.ransac_evaluate(np.vstack(clouds_b), np.concatenate(sizes_b), table, 0.03, threads=8)
    pos = 0
    for p, s in zip(batches[0], sizes_b):
        n = int(s.sum())
        og2.apply_mask(p, res["mask"][pos:pos + n].astype(bool))
        pos += n
    for p in clouds:
        got = grid._host.forest.export_points(grid._host.pose_index[p], order=0)
        want = np.concatenate([l.idx for l in og2.get_leaf_points(p)])
        assert (got["idx"] == want).all()
    g3 = Grid(GridConfig())
    g3.insert_points(0, clouds[0][:100])
    g3.insert_points(5, clouds[1][:100])
    with pytest.raises(KeyError):
        g3.map_leaf_points_cuda_ransac()
