"""Golden fixture for `Octree.subdivide_as` from the REAL reference (run in the build container, where /root/reference
exists; the fixture travels):  python tests/golden/make_subdivide_as.py

Two stand-alone octrees over the same root: A is subdivided by a count criterion, B copies A's scheme
(octree/octree.py:34-53, 222-227).  Recorded: B's leaves in the reference's order (corner, edge, points), its counters, and
a second scenario in which B was split finer before and is COLLAPSED onto A's coarser scheme - there the reference forgets
to re-list the collapsed nodes (octree.py:48-53), so only what survives in its leaf list and the total point count of
`get_points` are recorded.  The oracle (oracle/structure.py) is checked against the same run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np

if not hasattr(np, "float_"):
    np.float_ = np.float64
from oracle import ref_loader  # noqa: E402  (import shim: stub k3d, CUDASIM)

ref = ref_loader.load(cudasim=True)
import octreelib.octree  # noqa: E402,F401  (the reference's)
from oracle.structure import _Tree  # noqa: E402


class _StableInv(np.ndarray):
    def argsort(self, *a, **k):
        k.setdefault("kind", "stable")
        return np.asarray(self).argsort(*a, **k)


class stable_order:
    """Makes the reference's `inverse.argsort()` (octree.py:87) stable at run time: the canonical point order of
    SURVEY.md 8(c), the same shim tests/golden/make_golden.py uses."""

    def __enter__(self):
        self._orig = np.unique

        def _unique(*a, **k):
            r = self._orig(*a, **k)
            return (r[0], r[1].view(_StableInv)) if k.get("return_inverse") else r

        np.unique = _unique

    def __exit__(self, *exc):
        np.unique = self._orig


def clouds():
    rng = np.random.default_rng(20261018)
    a = np.vstack([rng.normal([3, 4, 5], 0.6, (700, 3)), rng.normal([12, 11, 2], 1.0, (500, 3)), rng.uniform(0, 16, (300, 3))])
    b = np.vstack([rng.normal([4, 4, 4], 1.5, (900, 3)), rng.uniform(0, 16, (1100, 3))])
    keep = lambda c: c[((c >= 0) & (c < 16)).all(axis=1)]
    return keep(a), keep(b)


def leaves_of(tree):
    out = tree.get_leaf_points()
    return (np.array([v.corner_min for v in out], dtype=np.float64).reshape(-1, 3), np.array([v.edge_length for v in out], dtype=np.float64),
            [np.asarray(v.get_points(), dtype=np.float64).reshape(-1, 3) for v in out])


def main():
    from octreelib.octree import Octree, OctreeConfig  # the REFERENCE (oracle/_ref is first on sys.path)
    a, b = clouds()
    corner, edge = np.array([0.0, 0.0, 0.0]), np.float64(16.0)
    A = Octree(OctreeConfig(), corner, edge)
    A.insert_points(a)
    A.subdivide([lambda p: len(p) > 40])
    B = Octree(OctreeConfig(), corner, edge)
    B.insert_points(b)
    B.subdivide_as(A)
    bc, be, bp = leaves_of(B)
    # oracle cross-check of the same scenario
    ta, tb = _Tree(corner, edge), _Tree(corner, edge)
    ta.insert(ta.root, np.arange(len(a)), a)
    ta.subdivide(ta.root, [lambda p: len(p) > 40])
    tb.insert(tb.root, np.arange(len(b)), b)
    tb.subdivide_as(tb.root, ta.root)
    ol = [n for n in tb.cache.values() if len(n.idx)]
    assert len(ol) == len(bp)
    for n, c, e, p in zip(ol, bc, be, bp):
        assert (np.asarray(n.corner, dtype=np.float64) == c).all() and float(n.edge) == e and (n.pts == p).all()
    # collapse scenarios.  (1) B split finer than A: the reference RAISES - `_remove_from_cache` of a child that has
    # children of its own (octree.py:52, 197: list.remove of a node that is not a leaf).  (2) one level collapsed onto an
    # unsplit octree: it works, but the collapsed root is not put back into the leaf list.
    B2 = Octree(OctreeConfig(), corner, edge)
    B2.insert_points(b)
    B2.subdivide([lambda p: len(p) > 8])       # finer than A
    try:
        B2.subdivide_as(A)
        deep_error = ""
    except Exception as exc:  # noqa: BLE001
        deep_error = f"{type(exc).__name__}: {exc}"
    A0 = Octree(OctreeConfig(), corner, edge)
    A0.insert_points(a)
    B3 = Octree(OctreeConfig(), corner, edge)
    B3.insert_points(b)
    B3.subdivide([lambda p: len(p) > 1500])    # the root only
    assert B3.n_nodes == 9
    B3.subdivide_as(A0)
    c3, e3, p3 = leaves_of(B3)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "subdivide_as_edge16.npz"), a=a, b=b, corner=corner, edge=edge,
                        max_points=40, b_corner=bc, b_edge=be, b_sizes=np.array([len(p) for p in bp]), b_points=np.vstack(bp),
                        b_n_leaves=B.n_leaves, b_n_nodes=B.n_nodes, b_n_points=B.n_points,
                        collapse_deep_reference_error=deep_error,
                        collapse_one_level_n_nodes=B3.n_nodes, collapse_one_level_points=len(B3.get_points()),
                        collapse_one_level_listed_leaves=len(p3), collapse_one_level_n_points=B3.n_points)
    print("subdivide_as fixture:", len(bp), "leaves,", B.n_nodes, "nodes; deep collapse in the reference:", deep_error or "ok",
          "; one-level collapse: n_nodes", B3.n_nodes, "get_points", len(B3.get_points()), "listed leaves", len(p3))


if __name__ == "__main__":
    with stable_order():
        main()
