// Per-point cell coordinate and in-cell Morton code, host + device (bit-identical on both: plain
// IEEE float64 add / sub / compare / fmod, library built with -fmad=false).
//
// Restates, per point instead of per node:
//   cell key      (points - corner) // edge            /root/reference/octreelib/grid/grid.py:72-76
//   child routing ((p - node.corner) // (edge/2))      /root/reference/octreelib/octree/octree.py:73-75
//   child id      4*ix + 2*iy + iz                     /root/reference/octreelib/octree/octree.py:94-97
//   child corner  node.corner + (0|edge/2, ...)        /root/reference/octreelib/octree/octree.py:181-190
// The reference evaluates the routing level by level with the *rounded* difference p - corner and
// node corners built by repeated `corner + edge/2`; doing exactly the same float operations per
// point makes the digit sequence identical to the reference's even for points that sit within an
// ulp of a node boundary.
#pragma once
#include "common.cuh"

namespace ol {

struct KeyParams {
    double edge;
    double corner[3];
    int single_cell;     // one fixed cell: q = 0, cell corner = corner
    int depth;           // Morton levels, <= OL_MAX_DEPTH
    long long qmin[3];   // packed key = sum (q[a] - qmin[a]) << shift[a]
    int shift[3];
    int pose_bits;       // low bits of the sort key reserved for the pose index (multi-segment input)
    double inv_edge;     // 1 / edge when the edge is a power of two (exact: see npy_floor_divide_inv), else 0
};

constexpr unsigned long long MORTON_BAD_BIT = 1ull << 63;

// integer cell coordinate of p along one axis (grid.py:72-76 without the `* edge` rescale)
__host__ __device__ inline double cell_coord(double p, double corner, double edge) {
    return npy_floor_divide(p - corner, edge);
}
__host__ __device__ inline double cell_coord_inv(double p, double corner, double edge, double inv_edge) {
    return npy_floor_divide_inv(p - corner, edge, inv_edge);
}

// Returns the Morton code (3 bits per level, level 0 in the most significant used bits);
// *bad_level = first level at which the point lies outside its node (depth if none).
__host__ __device__ inline unsigned long long point_morton(const double p[3], const double cell_corner[3], double edge,
                                                           int depth, int* bad_level) {
    double c0 = cell_corner[0], c1 = cell_corner[1], c2 = cell_corner[2];
    double e = edge;
    unsigned long long m = 0;
    int bad = depth;
    for (int d = 0; d < depth; ++d) {
        const double h = e * 0.5;  // edge / np.float_(2)  (octree.py:181)
        const double t0 = p[0] - c0, t1 = p[1] - c1, t2 = p[2] - c2;
        const bool in0 = (t0 >= 0.0) && (t0 < e), in1 = (t1 >= 0.0) && (t1 < e), in2 = (t2 >= 0.0) && (t2 < e);
        if (!(in0 && in1 && in2) && bad == depth) bad = d;
        const unsigned b0 = t0 >= h, b1 = t1 >= h, b2 = t2 >= h;  // floor_divide(t, h) for t in [0, 2h)
        if (b0) c0 = c0 + h;  // corner_min + offset  (octree.py:186)
        if (b1) c1 = c1 + h;
        if (b2) c2 = c2 + h;
        m = (m << 3) | (unsigned long long)((b0 << 2) | (b1 << 1) | b2);
        e = h;
    }
    *bad_level = bad;
    return m;
}

// owner rank of a cell under the hash partition of the multi-GPU grid (partition.cu, exchange.cu)
__host__ __device__ inline uint32_t cell_owner(long long qx, long long qy, long long qz, uint32_t world) {
    unsigned long long h = (unsigned long long)qx * 0x9E3779B97F4A7C15ull;
    h ^= (unsigned long long)qy * 0xC2B2AE3D27D4EB4Full;
    h ^= (unsigned long long)qz * 0x165667B19E3779F9ull;
    h ^= h >> 32;
    h *= 0xD6E8FEB86659FD93ull;
    h ^= h >> 32;
    return (uint32_t)((h >> 16) % world);
}

// corner of the cell with integer coordinates q (float64; exact for integer-valued edges)
__host__ __device__ inline double cell_corner_coord(long long q, double corner, double edge, int single_cell) {
    if (single_cell) return corner;
    return corner + (double)q * edge;
}

}  // namespace ol
