"""
`CudaRansac`: batched per-leaf RANSAC plane segmentation on the B200.

Same surface as the reference class (octreelib/ransac/cuda_ransac.py:18-81): the constructor draws
the hypothesis table from the GLOBAL numpy RNG with the same call (`np.random.random((min(H, 1024),
K))`, cuda_ransac.py:39-41) so that `np.random.seed(s)` right before construction yields the same
table in both libraries; `evaluate(point_cloud, block_sizes)` returns the boolean inlier mask.
The work is done by the hand-written sm_100a kernel in csrc/ransac.cu through the C ABI entry
`ol_ransac_evaluate`; torch tensors are only the device buffers.  Extras (not in the reference):
`last_planes`, `last_best`, `last_best_count` hold the winning plane / hypothesis / inlier count of
every block of the last `evaluate` call.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import numpy.typing as npt

from .. import _native as N
from ..forest import TorchAllocator, require_cuda

__all__ = ["CudaRansac"]

CUDA_THREADS = 1024


class CudaRansac:
    def __init__(self, threshold: float = 0.01, hypotheses_number: int = CUDA_THREADS, initial_points_number: int = 6):
        self._threshold = threshold
        self._hypotheses = min(hypotheses_number, CUDA_THREADS)
        self._initial_points_number = initial_points_number
        # drawn exactly like the reference draws it (cuda_ransac.py:39-41)
        self._table = np.random.random((self._hypotheses, initial_points_number))
        self._table_dev = None
        self.last_planes = self.last_best = self.last_best_count = None
        self.flags = 0

    @property
    def random_hypotheses(self) -> np.ndarray:
        """The (H, K) float64 uniform table the sample indices are derived from."""
        return self._table

    def evaluate(self, point_cloud, block_sizes: npt.NDArray):
        """point_cloud: (N, 3) float64, blocks laid out back to back; block_sizes: (B,) ints.
        Returns the (N,) bool inlier mask of the best plane of every block (all False for blocks with
        fewer than `initial_points_number` points, cuda_ransac.py:96-97)."""
        torch = require_cuda()
        lib = N.lib()
        device = torch.device("cuda", torch.cuda.current_device())
        pts = np.ascontiguousarray(point_cloud, dtype=np.float64).reshape(-1, 3)
        bs = np.ascontiguousarray(block_sizes, dtype=np.int32)
        n, nb = len(pts), len(bs)
        stream = torch.cuda.current_stream(device)
        d_pts = torch.from_numpy(pts).to(device, non_blocking=False)
        d_bs = torch.from_numpy(bs).to(device)
        if self._table_dev is None or self._table_dev.device != device:
            self._table_dev = torch.from_numpy(self._table).to(device)
        d_mask = torch.empty(max(n, 1), dtype=torch.uint8, device=device)
        d_plane = torch.empty((max(nb, 1), 4), dtype=torch.float32, device=device)
        d_best = torch.empty(max(nb, 1), dtype=torch.int32, device=device)
        d_cnt = torch.empty(max(nb, 1), dtype=torch.int32, device=device)
        alloc = TorchAllocator(device)
        H, K = self._table.shape
        N.check(lib.ol_ransac_evaluate(C.c_void_p(stream.cuda_stream), C.c_void_p(d_pts.data_ptr()), n,
                                       C.c_void_p(d_bs.data_ptr()), nb, C.c_void_p(self._table_dev.data_ptr()), H, K,
                                       float(self._threshold), C.c_void_p(d_mask.data_ptr()),
                                       C.c_void_p(d_plane.data_ptr()), C.c_void_p(d_best.data_ptr()),
                                       C.c_void_p(d_cnt.data_ptr()), int(self.flags), alloc.alloc_cb, alloc.free_cb, None))
        self.last_planes = d_plane[:nb].cpu().numpy()
        self.last_best = d_best[:nb].cpu().numpy()
        self.last_best_count = d_cnt[:nb].cpu().numpy()
        return d_mask[:n].cpu().numpy().astype(np.bool_)
