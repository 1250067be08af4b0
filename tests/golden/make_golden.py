"""
Generates tests/golden/*.npz by running the REAL reference (imported from /root/reference, which
exists only in the build container) and checks, while doing so, that the CPU oracle
(`oracle/structure.py`, `oracle/ransac.py`) reproduces it.  Run from the repo root:

    NUMBA_ENABLE_CUDASIM=1 python tests/golden/make_golden.py

Import shim (documented in SURVEY.md appendix B; no reference source is modified or copied):
  * `np.float_` alias   (octreelib/internal/point.py:15-16 predates numpy 2)
  * stub `k3d` module   (octreelib/grid/grid.py:5 imports it; only `visualize` uses it)
  * NUMBA_ENABLE_CUDASIM=1 -- the reference's own CI setting (.github/workflows/test.yml:47-48)
Canonical point order = reference + "stable argsort" runtime shim.
"""
import os
import sys
import types

os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
import numpy as np

np.float_ = np.float64
sys.modules.setdefault("k3d", types.ModuleType("k3d"))
REF = os.environ.get("OCTREELIB_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from octreelib.grid import Grid, GridConfig  # noqa: E402  (the reference)
from octreelib.ransac.cuda_ransac import CudaRansac  # noqa: E402

from oracle.structure import OracleGrid, max_points_criterion  # noqa: E402
from oracle import ransac as oransac  # noqa: E402
from octreelib_b200.synthetic import lidar64_scan, indoor_scene  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


class _StableInv(np.ndarray):
    def argsort(self, *a, **k):
        k.setdefault("kind", "stable")
        return np.asarray(self).argsort(*a, **k)


class stable_order:
    """Makes the reference's `inverse.argsort()` (grid.py:88, octree.py:87) stable, at run time."""

    def __enter__(self):
        self._orig = np.unique

        def _unique(*a, **k):
            r = self._orig(*a, **k)
            return (r[0], r[1].view(_StableInv)) if k.get("return_inverse") else r

        np.unique = _unique

    def __exit__(self, *exc):
        np.unique = self._orig


def _index_of(cloud):
    """Map exact point bytes -> input index (inputs are generated without duplicate points)."""
    d = {}
    for i, p in enumerate(cloud):
        d[p.tobytes()] = i
    assert len(d) == len(cloud), "duplicate points in a golden input"
    return d


def dump_reference(grid, clouds):
    """Flatten the reference grid into arrays (per pose: leaf corners, edges, sizes, point indices)."""
    out = {}
    for pose, cloud in clouds.items():
        lut = _index_of(cloud)
        leaves = grid.get_leaf_points(pose)
        out[f"p{pose}_corner"] = np.array([np.asarray(v.corner_min, dtype=np.float64) for v in leaves]).reshape(-1, 3)
        out[f"p{pose}_edge"] = np.array([float(v.edge_length) for v in leaves])
        out[f"p{pose}_size"] = np.array([v.n_points for v in leaves], dtype=np.int64)
        idx = [np.array([lut[p.tobytes()] for p in v.get_points()], dtype=np.int64) for v in leaves]
        out[f"p{pose}_idx"] = np.concatenate(idx) if idx else np.empty(0, dtype=np.int64)
        out[f"p{pose}_counts"] = np.array([grid.n_leaves(pose), grid.n_points(pose), grid.n_nodes(pose)])
        gp = grid.get_points(pose)
        out[f"p{pose}_getpoints_idx"] = np.array([lut[p.tobytes()] for p in gp], dtype=np.int64)
        cells = grid._Grid__pose_voxel_coordinates[pose]
        out[f"p{pose}_cells"] = np.array([np.asarray(c.corner_min) for c in cells], dtype=np.int64).reshape(-1, 3)
    return out


def dump_oracle(og, clouds):
    out = {}
    for pose in clouds:
        leaves = og.get_leaf_points(pose)
        out[f"p{pose}_corner"] = np.array([np.asarray(l.corner, dtype=np.float64) for l in leaves]).reshape(-1, 3)
        out[f"p{pose}_edge"] = np.array([float(l.edge) for l in leaves])
        out[f"p{pose}_size"] = np.array([len(l.idx) for l in leaves], dtype=np.int64)
        out[f"p{pose}_idx"] = np.concatenate([l.idx for l in leaves]) if leaves else np.empty(0, dtype=np.int64)
        out[f"p{pose}_counts"] = np.array([og.n_leaves(pose), og.n_points(pose), og.n_nodes(pose)])
        out[f"p{pose}_getpoints_idx"] = og.get_point_indices(pose)
        out[f"p{pose}_cells"] = np.array(og.pose_cells[pose], dtype=np.int64).reshape(-1, 3)
    return out


def compare(ref, ora, ordered: bool, tag: str):
    for k in ref:
        a, b = ref[k], ora[k]
        if k.endswith("_idx") and not ordered:
            # as-is reference order is host dependent (unstable argsort): compare per-leaf SETS
            sizes = ref[k.replace("_getpoints_idx", "_size").replace("_idx", "_size")]
            if k.endswith("getpoints_idx"):
                assert sorted(a.tolist()) == sorted(b.tolist()), (tag, k)
                continue
            pos = 0
            for s in sizes:
                assert sorted(a[pos:pos + s].tolist()) == sorted(b[pos:pos + s].tolist()), (tag, k)
                pos += s
            continue
        assert a.shape == b.shape and (a == b).all(), (tag, k, a[:8], b[:8])


def structure_case(name, clouds, edge, max_points, subdivide_poses=None, filter_min=None):
    crit = [lambda pts: len(pts) > max_points]
    # (1) reference as-is: structure + point sets
    g = Grid(GridConfig(voxel_edge_length=edge))
    for pose, c in clouds.items():
        g.insert_points(pose, c)
    pre = dump_reference(g, clouds)
    g.subdivide(crit, subdivide_poses)
    if filter_min is not None:
        g.filter([lambda pts: len(pts) >= filter_min])
    asis = dump_reference(g, clouds)
    # (2) reference + stable shim: canonical order
    with stable_order():
        gs = Grid(GridConfig(voxel_edge_length=edge))
        for pose, c in clouds.items():
            gs.insert_points(pose, c)
        pre_s = dump_reference(gs, clouds)
        gs.subdivide(crit, subdivide_poses)
        if filter_min is not None:
            gs.filter([lambda pts: len(pts) >= filter_min])
        canon = dump_reference(gs, clouds)
    # (3) oracle
    og = OracleGrid(edge)
    for pose, c in clouds.items():
        og.insert_points(pose, c)
    pre_o = dump_oracle(og, clouds)
    og.subdivide([max_points_criterion(max_points)], subdivide_poses)
    if filter_min is not None:
        og.filter([lambda pts: len(pts) >= filter_min])
    ora = dump_oracle(og, clouds)
    compare(pre, pre_o, ordered=False, tag=name + ":pre/as-is")
    compare(pre_s, pre_o, ordered=True, tag=name + ":pre/stable")
    compare(asis, ora, ordered=False, tag=name + ":as-is")
    compare(canon, ora, ordered=True, tag=name + ":stable")
    save = {f"cloud{p}": c for p, c in clouds.items()}
    save.update({"pre_" + k: v for k, v in pre_s.items()})
    save.update(canon)
    save["edge"] = np.float64(edge)
    save["max_points"] = np.int64(max_points)
    save["poses"] = np.array(list(clouds.keys()), dtype=np.int64)
    save["subdivide_poses"] = np.array([] if subdivide_poses is None else subdivide_poses, dtype=np.int64)
    save["filter_min"] = np.int64(-1 if filter_min is None else filter_min)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    n = sum(len(c) for c in clouds.values())
    print(f"[golden] {name}: {n} pts, {len(og.cells)} cells, "
          f"leaves/pose {[int(canon[f'p{p}_counts'][0]) for p in clouds]}  -- oracle == reference OK")


def late_pose_case(name, clouds, late, edge, max_points):
    """Poses inserted AFTER a subdivision follow the existing scheme (octree_manager.py:161-171); new cells stay
    unsplit.  `late` = the pose numbers inserted after the subdivide call."""
    crit = [lambda pts: len(pts) > max_points]
    early = [p for p in clouds if p not in late]

    def run(make_grid, insert, subdivide):
        g = make_grid()
        for p in early:
            insert(g, p, clouds[p])
        subdivide(g)
        for p in late:
            insert(g, p, clouds[p])
        return g

    ref_asis = dump_reference(run(lambda: Grid(GridConfig(voxel_edge_length=edge)), lambda g, p, c: g.insert_points(p, c),
                                  lambda g: g.subdivide(crit)), clouds)
    with stable_order():
        canon = dump_reference(run(lambda: Grid(GridConfig(voxel_edge_length=edge)), lambda g, p, c: g.insert_points(p, c),
                                   lambda g: g.subdivide(crit)), clouds)
    og = run(lambda: OracleGrid(edge), lambda g, p, c: g.insert_points(p, c),
             lambda g: g.subdivide([max_points_criterion(max_points)]))
    ora = dump_oracle(og, clouds)
    compare(ref_asis, ora, ordered=False, tag=name + ":as-is")
    compare(canon, ora, ordered=True, tag=name + ":stable")
    save = {f"cloud{p}": c for p, c in clouds.items()}
    save.update(canon)
    save["edge"], save["max_points"] = np.float64(edge), np.int64(max_points)
    save["poses"] = np.array(list(clouds.keys()), dtype=np.int64)
    save["late"] = np.array(list(late), dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(f"[golden] {name}: late poses {list(late)}, leaves/pose {[int(canon[f'p{p}_counts'][0]) for p in clouds]}"
          "  -- oracle == reference OK")


def resubdivide_case(name, clouds, late, edge, first_max, second_max):
    """A second, finer subdivide ("deepening re-subdivides" are inside the reference's domain).  Every pose
    octree keeps its `_cached_leaves` list across calls (octree_base.py:48-49, octree.py:183-191), so the leaf order
    after the second call depends on WHEN a node was split: leaves created by the first call stay in front of the
    children appended by the second one, and a pose inserted between the calls (`late`) sees one single pass instead.
    The fixture stores the state after each stage and asserts that it really differs from a one-shot subdivision."""
    early = [p for p in clouds if p not in late]

    def run(make_grid, crit):
        stages = []
        g = make_grid()
        for p in early:
            g.insert_points(p, clouds[p])
        g.subdivide(crit(first_max))
        stages.append((g, {p: clouds[p] for p in early}))
        snap1 = (dump_reference if isinstance(g, Grid) else dump_oracle)(g, {p: clouds[p] for p in early})
        for p in late:
            g.insert_points(p, clouds[p])
        snap2 = (dump_reference if isinstance(g, Grid) else dump_oracle)(g, clouds)
        g.subdivide(crit(second_max))
        snap3 = (dump_reference if isinstance(g, Grid) else dump_oracle)(g, clouds)
        return snap1, snap2, snap3

    ref_crit = lambda m: [lambda pts: len(pts) > m]  # noqa: E731
    ora_crit = lambda m: [max_points_criterion(m)]  # noqa: E731
    asis = run(lambda: Grid(GridConfig(voxel_edge_length=edge)), ref_crit)
    with stable_order():
        canon = run(lambda: Grid(GridConfig(voxel_edge_length=edge)), ref_crit)
        fresh = Grid(GridConfig(voxel_edge_length=edge))
        for p in clouds:
            fresh.insert_points(p, clouds[p])
        fresh.subdivide(ref_crit(second_max))
        oneshot = dump_reference(fresh, clouds)
    ora = run(lambda: OracleGrid(edge), ora_crit)
    for i, tag in enumerate(("first", "late", "second")):
        compare(asis[i], ora[i], ordered=False, tag=f"{name}:{tag}/as-is")
        compare(canon[i], ora[i], ordered=True, tag=f"{name}:{tag}/stable")
    differs = [p for p in clouds if canon[2][f"p{p}_corner"].shape != oneshot[f"p{p}_corner"].shape
               or (canon[2][f"p{p}_corner"] != oneshot[f"p{p}_corner"]).any()]
    same_set = all(sorted(map(tuple, canon[2][f"p{p}_corner"])) == sorted(map(tuple, oneshot[f"p{p}_corner"])) for p in clouds)
    assert differs and same_set, "the fixture must exercise the history-dependent leaf order (same leaves, other order)"
    save = {f"cloud{p}": c for p, c in clouds.items()}
    save.update({"s1_" + k: v for k, v in canon[0].items()})
    save.update({"s2_" + k: v for k, v in canon[1].items()})
    save.update(canon[2])
    save["edge"], save["first_max"], save["second_max"] = np.float64(edge), np.int64(first_max), np.int64(second_max)
    save["poses"] = np.array(list(clouds.keys()), dtype=np.int64)
    save["late"] = np.array(list(late), dtype=np.int64)
    save["order_differs_from_one_shot"] = np.array(differs, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(f"[golden] {name}: subdivide(>{first_max}), late poses {list(late)}, subdivide(>{second_max}); leaves/pose "
          f"{[int(canon[2][f'p{p}_counts'][0]) for p in clouds]}, history order != one-shot order for poses {differs}"
          "  -- oracle == reference OK")


def ransac_case(name, clouds, edge, max_points, H, K, threshold, seed, poses_per_batch):
    """Runs the reference's numba kernel under CUDASIM (slow: keep #blocks * H small)."""
    crit = [lambda pts: len(pts) > max_points]
    with stable_order():
        g = Grid(GridConfig(voxel_edge_length=edge))
        for pose, c in clouds.items():
            g.insert_points(pose, c)
        g.subdivide(crit)
    # restate grid.py:149-197 far enough to capture evaluate()'s inputs and output per batch
    np.random.seed(seed)
    ransac = CudaRansac(threshold=threshold, hypotheses_number=H, initial_points_number=K)
    table = ransac._CudaRansac__random_hypotheses_cuda.copy_to_host()
    assert (table == oransac.make_table(H, K, seed=seed)).all()
    nposes = len(clouds)
    batches = [list(range(i, min(i + poses_per_batch, nposes))) for i in range(0, nposes, poses_per_batch)]
    save = {f"cloud{p}": c for p, c in clouds.items()}
    total_ties = 0
    for bi, batch in enumerate(batches):
        pcs, sizes = [], []
        for pose in batch:
            leaves = g.get_leaf_points(pose)
            pcs.append(np.vstack([v.get_points() for v in leaves]))
            sizes.append(np.array([len(v.get_points()) for v in leaves], dtype=np.int32))
        cloud = np.vstack(pcs)
        bs = np.concatenate(sizes)
        ref_mask = ransac.evaluate(cloud, bs)
        ora = oransac.ransac_evaluate(cloud, bs, table, threshold, full=True)
        onp = oransac.ransac_numpy(cloud, bs, table, threshold)
        assert (ora["counts"] == onp["counts"]).all() and (ora["mask"] == onp["mask"]).all()
        # tie-aware pin: the reference's mask must be the mask of SOME member of the tied-max set
        starts = ora["block_start"]
        chosen = np.full(len(bs), -1, dtype=np.int32)
        for b, (n, s) in enumerate(zip(bs, starts)):
            rm = ref_mask[s:s + n]
            if n < K:
                assert not rm.any()
                continue
            cnt = ora["counts"][b]
            tied = np.flatnonzero(cnt == cnt.max())
            total_ties += len(tied) > 1
            for t in tied:
                m = oransac.mask_for_plane(cloud, s, n, ora["planes"][b, t], threshold)
                if (m.astype(bool) == rm).all():
                    chosen[b] = t
                    break
            assert chosen[b] >= 0, (name, "block", b, "reference mask matches no tied-best hypothesis")
            assert rm.sum() == cnt.max()
        save[f"b{bi}_points"] = cloud
        save[f"b{bi}_block_sizes"] = bs
        save[f"b{bi}_ref_mask"] = ref_mask
        save[f"b{bi}_ref_choice"] = chosen
        save[f"b{bi}_best"] = ora["best"]
        save[f"b{bi}_best_count"] = ora["best_count"]
        save[f"b{bi}_plane"] = ora["plane"]
        save[f"b{bi}_mask"] = ora["mask"]
        save[f"b{bi}_poses"] = np.array(batch, dtype=np.int64)
        print(f"[golden] {name} batch {bi}: {len(bs)} blocks, {int((bs >= K).sum())} scored, "
              f"ref picks lowest index in {int((chosen == ora['best']).sum())}/{len(bs)}")
    save["table"] = table
    save["edge"], save["max_points"] = np.float64(edge), np.int64(max_points)
    save["H"], save["K"], save["threshold"], save["seed"] = np.int64(H), np.int64(K), np.float64(threshold), np.int64(seed)
    save["poses_per_batch"] = np.int64(poses_per_batch)
    save["n_batches"] = np.int64(len(batches))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **save)
    print(f"[golden] {name}: blocks with ties {total_ties}  -- oracle == reference (tie-aware) OK")


def resubdivide_fixture():
    """S8: clustered clouds (deep trees) subdivided twice, one pose arriving between the two calls (own RNG stream)."""
    r8 = np.random.default_rng(888)

    def f32(a):
        return a.astype(np.float32).astype(np.float64)

    cl = {}
    centers = r8.random((10, 3)) * 8
    for p in range(3):
        pts = centers[r8.integers(0, 10, 900)] + r8.normal(0, 0.25, (900, 3))
        cl[p] = f32(np.clip(pts, 0.01, 7.99))
    resubdivide_case("resubdivide_deepen_edge4", cl, late=[2], edge=4, first_max=60, second_max=12)


def subset_fixture():
    """S9: subdivide(criteria, pose_numbers=[0, 2]) - the scheme is built from the listed poses only and imposed on
    every pose of the cell (octree_manager.py:36-66).  Every pose lives in every cell (the reference raises KeyError
    for a listed pose that is absent from a cell)."""
    r9 = np.random.default_rng(999)
    clouds = {p: (r9.random((1400, 3)) * 4).astype(np.float32).astype(np.float64) for p in range(3)}
    structure_case("subset_subdivide_edge2", clouds, 2, 30, subdivide_poses=[0, 2])


def far_fixture():
    """S10: UTM-like coordinates with full float64 mantissas: `p - corner` and `corner + edge / 2` round at every level."""
    r10 = np.random.default_rng(1010)
    base = np.array([451234.0, 5412345.0, 120.0])
    clouds = {p: base + r10.random((1500, 3)) * np.array([6.0, 5.0, 2.0]) + np.array([0.7 * p, 0, 0]) for p in range(2)}
    structure_case("far_offset_edge1", clouds, 1, 6)


def ransac_far_fixture():
    """R5: the RANSAC kernel on planes 450 km / 5400 km from the origin (float32 plane offsets lose the threshold's
    precision there, exactly like in the reference: cuda_ransac.py:110-113), two poses, two batches."""
    pl = indoor_scene(500, seed=6) + np.array([451234.0, 5412345.0, 120.0])
    ransac_case("ransac_far_h64", {0: pl[:250], 1: pl[250:]}, 8.0, 70, H=64, K=6, threshold=0.05, seed=15, poses_per_batch=1)


def ransac_degenerate_fixture():
    """R6: leaves whose samples are collinear, coincident or exactly coplanar: a zero normal makes the reference return
    the plane (0, 0, 0, 0) (util.py:76-78), for which every point is an inlier; exact planes give many tied hypotheses."""
    r = np.random.default_rng(66)
    cells = []
    t = r.random(40)
    cells.append(np.c_[t * 7.0, t * 3.0 + 1.0, t * 2.0 + 0.5])                           # a line through cell (0,0,0)
    cells.append(np.tile(np.array([[9.5, 1.5, 2.5]]), (12, 1)) + np.r_[np.zeros((11, 3)), [[0.25, 0.5, 0.125]]])  # 11 copies + 1
    xy = r.integers(0, 32, (60, 2)) / 4.0
    cells.append(np.c_[16.0 + xy[:, 0] * 0.9, xy[:, 1] * 0.9, np.full(60, 3.0)])            # an exact plane z = 3
    cells.append(np.c_[r.random((50, 2)) * 7.5, 8.0 + r.random(50) * 7.5])                   # a generic cloud (cell (0,0,8))
    cloud = np.vstack(cells).astype(np.float32).astype(np.float64)
    # duplicates are part of the case, so the index map of dump_reference is not used here (ransac_case does not need it)
    ransac_case("ransac_degenerate_h64", {0: cloud}, 8.0, 1000, H=64, K=6, threshold=0.02, seed=16, poses_per_batch=10)


class RecordingK3d:
    """Stand-in for the k3d module that records what `Grid.visualize` (grid.py:269-341) draws."""

    class Plot(list):
        def __iadd__(self, item):
            self.append(item)
            return self

        def get_snapshot(self):
            return f"<recorded {len(self)} objects>"

    def __init__(self):
        self.plots = []

    def install(self, module):
        rec = self

        def make_plot():
            plot = RecordingK3d.Plot()
            rec.plots.append(plot)
            return plot

        module.Plot = make_plot
        module.points = lambda positions, point_size, color: ("points", np.array(positions, dtype=np.float64).reshape(-1, 3),
                                                              float(point_size), int(color))
        module.lines = lambda vertices, indices, width, color, indices_type: (
            "lines", np.array(vertices, dtype=np.float64).reshape(-1, 3), float(width), int(color), np.array(indices), indices_type)

    @staticmethod
    def flatten(plot):
        kinds = np.array([0 if o[0] == "points" else 1 for o in plot], dtype=np.int64)
        sizes = np.array([len(o[1]) for o in plot], dtype=np.int64)
        data = np.vstack([np.empty((0, 3))] + [o[1] for o in plot])
        scalar = np.array([o[2] for o in plot], dtype=np.float64)
        color = np.array([o[3] for o in plot], dtype=np.int64)
        return dict(kinds=kinds, sizes=sizes, data=data, scalar=scalar, color=color)


def visualize_fixture():
    """V1: what the reference's `Grid.visualize` hands to k3d (object order, colours drawn from `random.seed(seed)`,
    voxel wire frames), for both visualisation types and with unused voxels; recorded through a stub k3d module."""
    import tempfile

    import octreelib.grid.grid as ref_grid_module
    from octreelib.grid import GridVisualizationType, VisualizationConfig

    rec = RecordingK3d()
    rec.install(ref_grid_module.k3d)
    r = np.random.default_rng(4242)
    clouds = {p: (r.random((160, 3)) * np.array([7.0, 3.5, 3.5]) + np.array([0.4 * p, 0, 0])).astype(np.float32).astype(np.float64)
              for p in range(3)}
    with stable_order():
        g = Grid(GridConfig(voxel_edge_length=4))
        for p, c in clouds.items():
            g.insert_points(p, c)
        g.subdivide([lambda pts: len(pts) > 25])
        # "unused" voxels named the way a user would: ids read off the leaves of pose 1
        picked = [1, 4]
        unused = [g.get_leaf_points(1)[i].id for i in picked]
        save = {f"cloud{p}": c for p, c in clouds.items()}
        save["edge"], save["max_points"] = np.float64(4), np.int64(25)
        save["unused_from_pose"], save["unused_leaf_positions"] = np.int64(1), np.array(picked, dtype=np.int64)
        cases = [("pose", GridVisualizationType.POSE, 3, []), ("voxel", GridVisualizationType.VOXEL, 5, []),
                 ("pose_unused", GridVisualizationType.POSE, 7, unused), ("voxel_unused", GridVisualizationType.VOXEL, 9, unused)]
        with tempfile.TemporaryDirectory() as tmp:
            for tag, vtype, seed, unused_ids in cases:
                cfg = VisualizationConfig(type=vtype, point_size=0.05, line_width_size=0.02, line_color=0x00FF00,
                                          filepath=os.path.join(tmp, tag + ".html"), seed=seed, unused_voxels=list(unused_ids))
                g.visualize(cfg)
                flat = RecordingK3d.flatten(rec.plots[-1])
                save.update({f"{tag}_{k}": v for k, v in flat.items()})
                save[f"{tag}_seed"] = np.int64(seed)
                assert open(cfg.filepath).read().startswith("<recorded")
    np.savez_compressed(os.path.join(OUT, "visualize_edge4.npz"), **save)
    print(f"[golden] visualize_edge4: {[len(p) for p in rec.plots]} k3d objects per call recorded from the reference")


def all_leaves_fixture():
    """S11: `get_leaf_points(pose, non_empty=False)` (grid.py:217-232): every leaf - empty ones included - of the cells in
    which the pose owns an octree, before and after a filter that empties leaves."""
    r = np.random.default_rng(1111)
    clouds = {0: (r.random((500, 3)) * np.array([6.0, 4.0, 2.0])).astype(np.float32).astype(np.float64),
              1: (r.random((300, 3)) * np.array([3.0, 4.0, 2.0]) + np.array([3.5, 0.0, 0.0])).astype(np.float32).astype(np.float64)}
    crit = [lambda pts: len(pts) > 12]

    def table(leaves, corner, edge, npts):
        return dict(corner=np.array([np.asarray(corner(v), dtype=np.float64) for v in leaves]).reshape(-1, 3),
                    edge=np.array([float(edge(v)) for v in leaves]), size=np.array([npts(v) for v in leaves], dtype=np.int64))

    with stable_order():
        g = Grid(GridConfig(voxel_edge_length=2))
        for p, c in clouds.items():
            g.insert_points(p, c)
        g.subdivide(crit)
        ref1 = {p: table(g.get_leaf_points(p, non_empty=False), lambda v: v.corner_min, lambda v: v.edge_length, lambda v: v.n_points)
                for p in clouds}
        g.filter([lambda pts: len(pts) >= 5])
        ref2 = {p: table(g.get_leaf_points(p, non_empty=False), lambda v: v.corner_min, lambda v: v.edge_length, lambda v: v.n_points)
                for p in clouds}
    og = OracleGrid(2)
    for p, c in clouds.items():
        og.insert_points(p, c)
    og.subdivide([max_points_criterion(12)])
    ora1 = {p: table(og.get_leaf_points(p, non_empty=False), lambda l: l.corner, lambda l: l.edge, lambda l: len(l.idx)) for p in clouds}
    og.filter([lambda pts: len(pts) >= 5])
    ora2 = {p: table(og.get_leaf_points(p, non_empty=False), lambda l: l.corner, lambda l: l.edge, lambda l: len(l.idx)) for p in clouds}
    save = {f"cloud{p}": c for p, c in clouds.items()}
    for p in clouds:
        for stage, ref, ora in (("a", ref1, ora1), ("b", ref2, ora2)):
            for k in ("corner", "edge", "size"):
                assert ref[p][k].shape == ora[p][k].shape and (ref[p][k] == ora[p][k]).all(), (p, stage, k)
                save[f"{stage}_p{p}_{k}"] = ref[p][k]
        assert (ref2[p]["size"] == 0).any(), "the fixture must contain empty leaves"
    save["edge"], save["max_points"], save["filter_min"] = np.float64(2), np.int64(12), np.int64(5)
    np.savez_compressed(os.path.join(OUT, "all_leaves_edge2.npz"), **save)
    print(f"[golden] all_leaves_edge2: leaves/pose incl. empty {[len(ref2[p]['size']) for p in clouds]}, empty "
          f"{[int((ref2[p]['size'] == 0).sum()) for p in clouds]}  -- oracle == reference OK")


def size_limit_fixture():
    """S11: size thresholds (BASELINE north_star: "splitting by point-count and size thresholds").  The reference passes a
    criterion nothing but the points (octree/octree.py:26), so the size-guarded criterion objects of
    octreelib_b200/criteria.py read the node under test from the caller's frame (`self` of OctreeNode.subdivide); the very
    same objects drive the UNMODIFIED reference here.  Stored: clustered 2-pose clouds (deep trees) under
    (a) MaxPoints(6, min_edge=0.5), (b) [MaxPoints(40), MaxPoints(6, max_depth=2)] (any() of two criteria with different
    limits), (c) [MaxDepth(2)] (uniform refinement: empty nodes split too)."""
    from octreelib_b200.criteria import MaxDepth, MaxPoints

    r11 = np.random.default_rng(1111)

    def f32(a):
        return a.astype(np.float32).astype(np.float64)

    clouds = {}
    for p in range(2):
        centers = r11.random((5, 3)) * 7 - 3
        clouds[p] = f32(centers[r11.integers(0, 5, 900)] + r11.normal(0, 0.04, (900, 3)))
    edge = 4
    cases = {"a": lambda: [MaxPoints(6, min_edge=0.5)],
             "b": lambda: [MaxPoints(40), MaxPoints(6, max_depth=2, voxel_edge_length=edge)],
             "c": lambda: [MaxDepth(2, voxel_edge_length=edge)]}
    save = {f"cloud{p}": c for p, c in clouds.items()}
    for tag, crit in cases.items():
        g = Grid(GridConfig(voxel_edge_length=edge))
        for pose, c in clouds.items():
            g.insert_points(pose, c)
        g.subdivide(crit())
        asis = dump_reference(g, clouds)
        with stable_order():
            gs = Grid(GridConfig(voxel_edge_length=edge))
            for pose, c in clouds.items():
                gs.insert_points(pose, c)
            gs.subdivide(crit())
            canon = dump_reference(gs, clouds)
        og = OracleGrid(edge)
        for pose, c in clouds.items():
            og.insert_points(pose, c)
        og.subdivide(crit())
        ora = dump_oracle(og, clouds)
        compare(asis, ora, ordered=False, tag=f"size_limit:{tag}/as-is")
        compare(canon, ora, ordered=True, tag=f"size_limit:{tag}/stable")
        # the guard must bite: an unguarded MaxPoints(6) goes deeper
        if tag == "a":
            free = OracleGrid(edge)
            for pose, c in clouds.items():
                free.insert_points(pose, c)
            free.subdivide([max_points_criterion(6)])
            assert min(float(l.edge) for l in free.get_leaf_points(0)) < 0.5 <= min(canon["p0_edge"]), "guard without effect"
        save.update({f"{tag}_{k}": v for k, v in canon.items()})
        print(f"[golden] size_limit_edge4 case {tag}: leaves/pose {[int(canon[f'p{p}_counts'][0]) for p in clouds]}, "
              f"smallest leaf edge {min(canon['p0_edge'])}  -- oracle == reference OK")
    save["edge"] = np.float64(edge)
    save["poses"] = np.array(list(clouds.keys()), dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "size_limit_edge4.npz"), **save)


def main():
    rng = np.random.default_rng(2024)
    only = os.environ.get("GOLDEN_ONLY")  # GOLDEN_ONLY=late regenerates only the late-pose fixture (own RNG stream)
    if only == "size_limit":
        size_limit_fixture()
        return
    if only == "resub":
        resubdivide_fixture()
        return
    if only == "subset":
        subset_fixture()
        return
    if only == "far":
        far_fixture()
        return
    if only == "ransac_far":
        ransac_far_fixture()
        return
    if only == "ransac_degenerate":
        ransac_degenerate_fixture()
        return
    if only == "visualize":
        visualize_fixture()
        return
    if only == "all_leaves":
        all_leaves_fixture()
        return
    if only == "late":
        def f32(a):
            return a.astype(np.float32).astype(np.float64)
        rng7 = np.random.default_rng(777)
        c3 = {p: f32(rng7.random((900, 3)) * np.array([5.0, 4.0, 2.0]) + np.array([0.3 * p, 0, 0])) for p in range(3)}
        c3[3] = f32(rng7.random((700, 3)) * np.array([5.0, 4.0, 2.0]) + np.array([3.0, 2.0, 0.0]))
        late_pose_case("late_poses_edge2", c3, late=[2, 3], edge=2, max_points=12)
        return

    def f32(a):
        return a.astype(np.float32).astype(np.float64)

    # S1: the reference's own tiny fixtures (test/grid/test_grid.py:14-40)
    tg0 = np.array([[0, 0, 1], [0, 0, 2], [0, 0, 3], [9, 9, 8], [9, 9, 9]], dtype=float)
    tg1 = np.array([[1, 0, 1], [4, 0, 2], [0, 2, 3], [5, 9, 9], [9, 3, 8]], dtype=float)
    structure_case("ref_test_grid_gt2", {0: tg0, 1: tg1}, 5, 2)
    structure_case("ref_test_grid_gt3", {0: tg0, 1: tg1}, 5, 3)
    # S2: random multi-pose cloud, negative coordinates, integer edge 2
    clouds = {p: f32(rng.random((1500, 3)) * np.array([9.0, 7.0, 3.0]) - np.array([4.0, 3.0, 1.0])) for p in range(3)}
    structure_case("random_3pose_edge2", clouds, 2, 10)
    structure_case("random_3pose_edge2_filter", clouds, 2, 25, filter_min=4)
    # S3: clustered cloud -> deep trees (depth ~6) with empty children
    cl = []
    for p in range(2):
        centers = rng.random((6, 3)) * 6 - 3
        pts = centers[rng.integers(0, 6, 1200)] + rng.normal(0, 0.03, (1200, 3))
        cl.append(f32(pts))
    structure_case("clustered_2pose_edge4", {0: cl[0], 1: cl[1]}, 4, 12)
    # S4: 64-beam lidar, 2 poses x 12k points (first rings = dense near field), edge 1, 100/leaf
    structure_case("lidar_2pose_edge1", {p: lidar64_scan(p, seed=0)[::10] for p in range(2)}, 1.0, 100)
    # S5: indoor planes, one pose, edge 1
    structure_case("indoor_1pose_edge1", {0: indoor_scene(8000, seed=1)}, 1.0, 40)
    # S6: sparse pose numbering is not needed for structure (poses are dict keys): poses 0 and 1,
    #     second pose lives partly in cells the first never touches
    a = f32(rng.random((600, 3)) * 4)
    b = f32(rng.random((600, 3)) * 4 + np.array([2.0, 0, 0]))
    structure_case("offset_poses_edge1", {0: a, 1: b}, 1, 8)

    # S7: two poses inserted after the subdivision; the last one also opens cells the scheme never saw
    rng7 = np.random.default_rng(777)
    c3 = {p: f32(rng7.random((900, 3)) * np.array([5.0, 4.0, 2.0]) + np.array([0.3 * p, 0, 0])) for p in range(3)}
    c3[3] = f32(rng7.random((700, 3)) * np.array([5.0, 4.0, 2.0]) + np.array([3.0, 2.0, 0.0]))
    late_pose_case("late_poses_edge2", c3, late=[2, 3], edge=2, max_points=12)
    # S8: deepening re-subdivide with a pose inserted between the two calls
    resubdivide_fixture()
    # S9: subdivision driven by a subset of the poses
    subset_fixture()
    # S10: far from the origin
    far_fixture()

    # R1..: RANSAC under CUDASIM (about 2 s per block at H=1024 -> small H / few blocks)
    pl = indoor_scene(700, seed=3)
    ransac_case("ransac_indoor_h128", {0: pl[:350], 1: pl[350:]}, 8.0, 60, H=128, K=6, threshold=0.02, seed=11,
                poses_per_batch=10)
    ransac_case("ransac_indoor_ppb1_h64", {0: pl[:350], 1: pl[350:]}, 8.0, 80, H=64, K=6, threshold=0.02, seed=12,
                poses_per_batch=1)
    li = lidar64_scan(0, seed=5)[::40]
    ransac_case("ransac_lidar_h64_k3", {0: li}, 8.0, 150, H=64, K=3, threshold=0.05, seed=13, poses_per_batch=10)
    ransac_case("ransac_lidar_h1024", {0: li[:400]}, 16.0, 200, H=1024, K=6, threshold=0.03, seed=14,
                poses_per_batch=10)
    ransac_far_fixture()
    ransac_degenerate_fixture()
    visualize_fixture()
    all_leaves_fixture()
    size_limit_fixture()


if __name__ == "__main__":
    main()
