"""Per-leaf plane fitting on the GPU; `CudaRansac` keeps the name and call signature of the reference's class
(octreelib/ransac/cuda_ransac.py:18-81) and binds `ol_ransac_evaluate` of the native library."""
from .cuda_ransac import CudaRansac

__all__ = ["CudaRansac"]
