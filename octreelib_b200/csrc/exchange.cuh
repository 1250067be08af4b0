// Fused multi-GPU exchange: every rank routes its local clouds to the ranks that own their grid cells with
// stores into peer-mapped memory over NVLink, and the receiving forest ADOPTS the receive buffer as its point array.
// All coordination between the ranks happens on the device (flags in peer-mapped control blocks); the host synchronises
// once, to learn how many rows of which pose arrived from which rank.  See exchange.cu.
#pragma once
#include <vector>

#include "common.cuh"

namespace ol {

struct Forest;

constexpr int XCHG_MAX_WORLD = 64;
constexpr int XCHG_BINS = 1024;     // bins of a rank's histogram of the leading cell coordinate (slab boundaries)
constexpr int XCHG_TILE = 2048;     // points per CTA of the count / scatter kernels
constexpr int XCHG_STAGES = 3;

// byte offsets inside a rank's control block (peer-mapped; written by its owner, read by everybody; the flags are written
// by the peers)
struct XchgCtrl {
    static constexpr size_t FLAG = 0;                                             // uint64 [XCHG_STAGES][XCHG_MAX_WORLD]
    static constexpr size_t RANGE = FLAG + 8 * XCHG_STAGES * XCHG_MAX_WORLD;      // int64 [2] min / max sampled cell x
    static constexpr size_t HIST = RANGE + 16;                                    // uint32 [XCHG_BINS]
    static constexpr size_t BBOX = HIST + 4 * XCHG_BINS;                          // int64 [6] ordered min xyz, max xyz
    static constexpr size_t ERR = BBOX + 48;                                      // uint32 device error bits (+ pad)
    static constexpr size_t CUBE = ERR + 16;                                      // uint32 [world][n_poses] rows to dst, per pose
    static size_t bytes(int world, int n_poses) { return CUBE + 4 * (size_t)world * (size_t)n_poses; }
};

// page-locked result block the kernels write and the host reads after ONE event wait
struct XchgHost {
    static constexpr size_t HDR = 0;          // int64 [16]: 0 err, 1 largest receive total, 2-4 q lo, 5-7 q hi, 8 timeout, 9 received
    static constexpr size_t BOUNDS = 128;     // int64 [XCHG_MAX_WORLD]
    static constexpr size_t COUNTS = BOUNDS + 8 * XCHG_MAX_WORLD;  // uint32 recv[world][n_poses], then pose_size[world][n_poses]
    static size_t bytes(int world, int n_poses) { return COUNTS + 8 * (size_t)world * (size_t)n_poses; }
};

struct Exchange {
    int world = 1, rank = 0, n_poses = 0, nbuf = 0, device = 0;
    int64_t rows_cap = 0;
    std::vector<unsigned char*> ctrl;  // [world]
    std::vector<double*> data;         // [nbuf][world]
    unsigned long long epoch = 0;
    unsigned char* host = nullptr;     // XchgHost block (cudaMallocHost)
    unsigned char* scratch = nullptr;  // persistent device scratch (bounds, bases, totals, prefix sums, pose sizes)
    size_t scratch_bytes = 0;
    cudaEvent_t ev = nullptr;

    Exchange(int world, int rank, int n_poses, int64_t rows_cap, int nbuf, void* const* ctrl_ptrs, void* const* data_ptrs, int device);
    ~Exchange();
    // info: [0] rows sent to other ranks, [1] rows received, [2] rows kept, [3] largest receive total of any rank
    void run(Forest& f, const double* const* clouds, const int64_t* sizes, const int32_t* poses, int count, int slabs, int buf,
             int64_t* info, int64_t* bounds_out, uint32_t* pose_sizes_out);
};

}  // namespace ol
