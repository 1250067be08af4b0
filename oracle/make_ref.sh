#!/bin/sh
# Recipe for oracle/_ref/: a byte-for-byte copy of the REAL reference package (prime-slam/octreelib, pure Python) taken
# from where it lies under /root/reference, so that it travels to the GPU box (oracle/_ref/ is git-ignored but NOT
# gpurun-ignored).  Nothing is modified and nothing from it enters the repository history.  It is used only
#   * by `bench.py --impl reference` / `cpu_baseline` (kind "reference"): the unmodified
#     Grid.insert_points -> subdivide -> get_leaf_points (grid/grid.py:58-109,244-258,217-232) on the host cores, and
#     CudaRansac.evaluate (ransac/cuda_ransac.py:43-81) through numba on the same B200 when numba's driver JIT works there;
#   * by tools/ref_on_gpu.py (parity of ol_ransac_evaluate against the reference's kernel on real hardware).
# The import shim the reference needs on this image (numpy 2: np.float_; k3d absent) lives in oracle/ref_loader.py.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${OCTREELIB_REFERENCE:-/root/reference}"
if [ ! -d "$SRC/octreelib" ]; then
    echo "make_ref: $SRC/octreelib not found (GPU box: the prebuilt oracle/_ref is used as shipped)" >&2
    exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC/octreelib" "$HERE/_ref/octreelib"
find "$HERE/_ref" -name __pycache__ -type d -exec rm -rf {} + 2>/dev/null || true
( cd "$SRC" && find octreelib -name '*.py' | sort | xargs sha256sum ) > "$HERE/_ref/SHA256SUMS"
echo "make_ref: copied $(find "$HERE/_ref/octreelib" -name '*.py' | wc -l) files to oracle/_ref/octreelib"
