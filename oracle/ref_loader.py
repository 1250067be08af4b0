"""
TEST / BENCH INFRASTRUCTURE (not product code): imports the REAL reference package from `oracle/_ref/`
(made by `oracle/make_ref.sh`, a verbatim copy of /root/reference/octreelib; `__graft_entry__.build()` runs the recipe
whenever /root/reference is present).  /root/reference itself is never read from here: it does not exist on the GPU box.

Import shim only - no reference source is modified:
  * `np.float_ = np.float64`   (octreelib/internal/point.py:15-16, octree/octree.py:181 predate numpy 2)
  * a stub `k3d` module        (octreelib/grid/grid.py:5 imports it at load time; only `visualize` uses it)
`cudasim=True` sets NUMBA_ENABLE_CUDASIM=1 (the reference's own CI setting, .github/workflows/test.yml:47-48) - it must
be decided before numba is imported for the first time in the process.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")


def reference_root():
    if os.environ.get("OL_NO_REFERENCE") == "1":  # tests: exercise the port fallback
        return None
    return REF_DIR if os.path.isdir(os.path.join(REF_DIR, "octreelib")) else None


def load(cudasim: bool):
    """Returns the imported reference package `octreelib` (or raises ImportError with the reason)."""
    root = reference_root()
    if root is None:
        raise ImportError("the reference package is not under oracle/_ref (run oracle/make_ref.sh where /root/reference exists)")
    if cudasim:
        os.environ["NUMBA_ENABLE_CUDASIM"] = "1"
    else:
        os.environ.pop("NUMBA_ENABLE_CUDASIM", None)
    import numpy as np

    if not hasattr(np, "float_"):
        np.float_ = np.float64
    sys.modules.setdefault("k3d", types.ModuleType("k3d"))
    if root not in sys.path:
        sys.path.insert(0, root)
    import octreelib  # noqa: F401  (the reference)
    import octreelib.grid  # noqa: F401
    import octreelib.ransac.cuda_ransac  # noqa: F401

    assert os.path.abspath(octreelib.__file__).startswith(os.path.abspath(root)), octreelib.__file__
    return octreelib
