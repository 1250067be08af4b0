"""Helpers shared by the -m gpu tests: raw C-ABI calls on torch buffers, oracle <-> native comparison."""
import ctypes as C

import numpy as np

from octreelib_b200 import _native as N
from octreelib_b200.forest import TorchAllocator


def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "the -m gpu tests need a CUDA device"
    return torch


def sort_pairs(keys: np.ndarray, vals: np.ndarray, begin_bit: int, end_bit: int):
    torch = torch_cuda()
    lib = N.lib()
    dev = torch.device("cuda", 0)
    alloc = TorchAllocator(dev)
    stream = torch.cuda.current_stream(dev)
    if keys.dtype == np.uint64:
        dk = torch.from_numpy(keys.view(np.int64).copy()).to(dev)
        fn = lib.ol_sort_pairs_u64
    else:
        dk = torch.from_numpy(keys.view(np.int32).copy()).to(dev)
        fn = lib.ol_sort_pairs_u32
    dv = torch.from_numpy(vals.view(np.int32).copy()).to(dev)
    N.check(fn(C.c_void_p(stream.cuda_stream), C.c_void_p(dk.data_ptr()), C.c_void_p(dv.data_ptr()), len(keys), begin_bit,
               end_bit, alloc.alloc_cb, alloc.free_cb, None))
    return dk.cpu().numpy().view(keys.dtype), dv.cpu().numpy().view(np.uint32)


def exclusive_scan(a: np.ndarray):
    torch = torch_cuda()
    lib = N.lib()
    dev = torch.device("cuda", 0)
    alloc = TorchAllocator(dev)
    stream = torch.cuda.current_stream(dev)
    d = torch.from_numpy(a.view(np.int32).copy()).to(dev)
    out = torch.empty_like(d)
    total = C.c_uint64(0)
    N.check(lib.ol_exclusive_scan_u32(C.c_void_p(stream.cuda_stream), C.c_void_p(d.data_ptr()), C.c_void_p(out.data_ptr()),
                                      len(a), C.byref(total), alloc.alloc_cb, alloc.free_cb, None))
    return out.cpu().numpy().view(np.uint32), total.value


def compare_grid_with_oracle(grid, og, poses, check_points=True):
    """Bit-exact comparison of a native Grid with an OracleGrid: leaf table (corner, edge, order), leaf
    sizes, point order inside leaves (original indices), counters, get_points order, cell lists."""
    host = grid._host
    forest = host.forest
    blocks = forest.export_blocks()
    leaves = forest.export_leaves()
    cells = forest.export_cells()
    for p in poses:
        pi = host.pose_index[p]
        want = og.get_leaf_points(p)
        sel = np.flatnonzero(blocks["pose"] == pi)
        w_corner = np.array([np.asarray(l.corner, dtype=np.float64) for l in want]).reshape(-1, 3)
        w_edge = np.array([float(l.edge) for l in want])
        w_size = np.array([len(l.idx) for l in want], dtype=np.int64)
        assert len(sel) == len(want), (p, len(sel), len(want))
        lf = blocks["leaf"][sel]
        assert (leaves["corner"][lf] == w_corner).all(), f"pose {p}: leaf corners / order differ"
        assert (leaves["edge"][lf] == w_edge).all(), f"pose {p}: leaf edges differ"
        assert (blocks["size"][sel] == w_size).all(), f"pose {p}: leaf sizes differ"
        got = forest.export_points(pi, order=0)
        w_idx = np.concatenate([l.idx for l in want]) if want else np.empty(0, dtype=np.int64)
        assert (got["idx"] == w_idx).all(), f"pose {p}: point order inside leaves differs"
        if check_points:
            w_pts = np.vstack([np.empty((0, 3))] + [l.points for l in want])
            assert (got["xyz"] == w_pts).all()
        assert [grid.n_leaves(p), grid.n_points(p), grid.n_nodes(p)] == [og.n_leaves(p), og.n_points(p), og.n_nodes(p)]
        # get_points: dict order of cells x depth-first leaves
        w_gp = og.get_point_indices(p)
        g_gp = grid.get_points(p)
        assert len(g_gp) == len(w_gp)
        # cells of the pose, lexicographic
        cp = forest.export_cell_poses()
        my_cells = cp["cell"][cp["pose"] == pi]
        edge = og.edge
        assert (cells["q"][my_cells] * int(edge) == np.array(og.pose_cells[p], dtype=np.int64).reshape(-1, 3)).all()
    return True
