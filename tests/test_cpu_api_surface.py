"""CPU: API surface, config validation (messages pinned by the reference's test_grid.py:148-182),
criteria folding, and the loud failure without a GPU."""
import numpy as np
import pytest

import octreelib_b200
from octreelib_b200.criteria import MaxPoints, MinPoints, as_threshold, fold_count_criteria
from octreelib_b200.grid import Grid, GridConfig, GridConfigBase, GridBase, VisualizationConfig, GridVisualizationType
from octreelib_b200.internal import Voxel, VoxelBase, WithID, PointCloud, Point, T  # noqa: F401
from octreelib_b200.octree import Octree, OctreeBase, OctreeConfig, OctreeConfigBase, OctreeNode, OctreeNodeBase  # noqa: F401
from octreelib_b200.octree_manager import OctreeManager
from octreelib_b200.ransac import CudaRansac  # noqa: F401


def test_invalid_octree_type_messages():
    with pytest.raises(TypeError) as e:
        Grid(GridConfig(octree_manager_type=type(None), octree_config=OctreeConfig(), voxel_edge_length=5))
    assert str(e.value) == ("Cannot use the provided octree manager type NoneType. "
                            "It has to be a subclass of octree_manager.OctreeManager.")
    with pytest.raises(TypeError) as e:
        Grid(GridConfig(octree_manager_type=OctreeManager, octree_type=type(None), octree_config=OctreeConfig(),
                        voxel_edge_length=5))
    assert str(e.value) == ("Cannot use the provided octree type NoneType. "
                            "It has to be a subclass of octree.OctreeBase.")


def test_defaults_match_reference():
    cfg = GridConfig()
    assert cfg.voxel_edge_length == 1 and cfg.debug is False and (cfg.corner == 0).all()
    assert cfg.octree_manager_type is OctreeManager and cfg.octree_type is Octree
    assert OctreeConfig().debug is True
    v = VisualizationConfig()
    assert v.type is GridVisualizationType.VOXEL and v.filepath == "visualization.html"


def test_voxel_identity_is_shared():
    a = Voxel(np.array([0, 0, 0.625]), 0.625)
    b = Voxel(np.array([0.0, 0.0, 0.625]), np.float64(0.625), points=np.zeros((1, 3)))
    c = Voxel(np.array([0, 0, 1.25]), 1.25)
    assert a.id == b.id and a.id != c.id and a == b
    assert len(a.all_corners) == 8 and (a.corner_max == np.array([0.625, 0.625, 1.25])).all()


def test_ransac_argument_validation_messages():
    g = Grid(GridConfig())
    for kwargs, msg in [(dict(threshold=0), "Threshold must be positive"),
                        (dict(hypotheses_number=0), "Number of RANSAC hypotheses must be positive"),
                        (dict(hypotheses_number=1025),
                         "Number of RANSAC hypotheses must be <= 1024 because of the CUDA thread limit.")]:
        with pytest.raises(ValueError) as e:
            g.map_leaf_points_cuda_ransac(**kwargs)
        assert str(e.value) == msg


def test_criteria_folding():
    t, beyond = fold_count_criteria([lambda p: len(p) > 100], "any", 1024)
    assert as_threshold(t, beyond) == 100
    t, beyond = fold_count_criteria([lambda p: p.shape[0] > 7, MaxPoints(30)], "any", 1024)
    assert as_threshold(t, beyond) == 7
    t, beyond = fold_count_criteria([lambda p: False], "any", 64)
    assert as_threshold(t, beyond) >= 1 << 40
    t, beyond = fold_count_criteria([lambda p: len(p) >= 2, MinPoints(4)], "all", 16)
    assert t.tolist() == [0, 0, 0, 0] + [1] * 13 and beyond
    t, beyond = fold_count_criteria([lambda p: len(p) % 2 == 0], "any", 8)
    assert beyond is None and as_threshold(t, beyond) is None and as_threshold(t, False) is None
    with pytest.raises(NotImplementedError):
        fold_count_criteria([lambda p: p.sum() > 1], "any", 8)


def test_install_as_alias():
    octreelib_b200.install_as("octreelib_alias_for_test")
    import octreelib_alias_for_test.grid as g  # noqa

    assert g.Grid is Grid


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Grid(GridConfig()).insert_points(0, np.zeros((4, 3)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        CudaRansac().evaluate(np.zeros((8, 3)), np.array([8]))
