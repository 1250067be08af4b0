// Forest: RANSAC orchestration, counters and host exports (included at the end of forest.cu).
namespace ol {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// Reference block order = (pose rank, leaf enumeration order).  The block table is ordered (leaf DFS index, pose), and
// the blocks of one leaf are consecutive, so the leaf part of the order is a permutation of whole leaf segments:
//   A: per leaf (in enumeration order) the number of its blocks, and the first block of every leaf;
//   (exclusive scan over the leaves)
//   B: block b goes to slot off[cache_rank[leaf]] + (b - first block of the leaf); its sort key is only the pose rank.
// A stable radix sort on the pose-rank bits alone then yields the reference order (2 passes of 32-bit keys for 839
// poses instead of 5 passes of 64-bit (pose rank, leaf) keys).
__global__ void leaf_block_count_kernel(uint32_t nb, const uint32_t* __restrict__ blk_leaf, const uint32_t* __restrict__ cache_rank,
                                        uint32_t* __restrict__ cnt_c, uint32_t* __restrict__ first_b) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t leaf = blk_leaf[b];
    if (b == 0 || blk_leaf[b - 1] != leaf) {
        first_b[leaf] = b;
        uint32_t e = b + 1;  // blocks of one leaf = its poses: a short run
        while (e < nb && blk_leaf[e] == leaf) ++e;
        cnt_c[cache_rank[leaf]] = e - b;
    }
}

__global__ void block_arrange_kernel(uint32_t nb, const uint32_t* __restrict__ blk_leaf, const int32_t* __restrict__ blk_pose,
                                     const int32_t* __restrict__ pose_rank, const uint32_t* __restrict__ cache_rank,
                                     const uint32_t* __restrict__ off_c, const uint32_t* __restrict__ first_b,
                                     uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t leaf = blk_leaf[b];
    const uint32_t dst = off_c[cache_rank[leaf]] + (b - first_b[leaf]);
    keys[dst] = (uint32_t)pose_rank[blk_pose[b]];
    vals[dst] = b;
}

// The same arrangement when the grid has been subdivided more than once: the reference's order inside (pose, cell) is then
// (max(epoch of the leaf's parent, epoch of the pose), one-call order) - forest.cuh - so the sort key carries the pose
// rank, the cell and that epoch; the one-call order is the initial arrangement and survives the stable sort.
__global__ void block_arrange_history_kernel(uint32_t nb, const uint32_t* __restrict__ blk_leaf, const int32_t* __restrict__ blk_pose,
                                             const int32_t* __restrict__ pose_rank, const uint32_t* __restrict__ cache_rank,
                                             const uint32_t* __restrict__ off_c, const uint32_t* __restrict__ first_b,
                                             const uint32_t* __restrict__ lcell, const int32_t* __restrict__ lparent,
                                             const uint32_t* __restrict__ iepoch, const int32_t* __restrict__ pose_epoch,
                                             int cell_bits, int epoch_bits, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t leaf = blk_leaf[b];
    const uint32_t dst = off_c[cache_rank[leaf]] + (b - first_b[leaf]);
    const int32_t p = blk_pose[b], par = lparent[leaf];
    const uint32_t node_epoch = par >= 0 ? iepoch[par] : 0u;
    const uint32_t eff = max(node_epoch, (uint32_t)pose_epoch[p]);
    keys[dst] = ((uint64_t)(uint32_t)pose_rank[p] << (cell_bits + epoch_bits)) | ((uint64_t)lcell[leaf] << epoch_bits) | (uint64_t)eff;
    vals[dst] = b;
}

// ---------------------------------------------------------------------------------------------
// Batch layout of the RANSAC kernel WITHOUT sorting the block table (the common case: one subdivide call, at most
// REFSTART_MAX_RANKS poses).  The reference's start index of block (pose p, leaf l) inside its batch is
//     [points of the batch's earlier poses] + [points of pose p in the leaves that precede l in enumeration order]
// (cuda_ransac.py:65-67 over grid.py:173-191).  With the blocks ARRANGED by (leaf enumeration order, pose) - a
// permutation of whole leaf segments, block_arrange above - the second term is the exclusive prefix, over the arranged
// slots, of the sizes of the blocks of pose p: a WEIGHTED stable ranking by pose.  Nothing has to move for that:
//   refstart_count_kernel   per chunk of 4096 arranged slots, the points of every pose rank -> table[rank][chunk]
//   (flat exclusive scan of the table: table[p][c] becomes points of ranks < p + points of rank p in chunks < c)
//   refstart_assign_kernel  the same chunks again: ordered prefix inside the chunk (per-warp counters, like the digit
//                           ranking of the radix sort but weighted with the block sizes) on top of the chunk's start
// instead of two radix passes over (pose rank, block) pairs, a size gather, a scan over the blocks and a search per
// block.  The sorted order itself is only needed when somebody asks for the result TABLE of all blocks in reference
// order (materialize_snapshot): the sort is deferred until then.
// ---------------------------------------------------------------------------------------------
// first and one-past-last block of every leaf (the blocks of a leaf are consecutive in the block table): two coalesced
// compares per block, no loop - the head of a leaf's run writes the one, its tail the other; leaves without blocks keep 0, 0
__global__ void leaf_block_span_kernel(uint32_t nb, const uint32_t* __restrict__ blk_leaf, uint32_t* __restrict__ first_b,
                                       uint32_t* __restrict__ last_b) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t leaf = blk_leaf[b];
    if (b == 0 || blk_leaf[b - 1] != leaf) first_b[leaf] = b;
    if (b + 1 == nb || blk_leaf[b + 1] != leaf) last_b[leaf] = b + 1;
}
struct LeafSpanIn {  // number of blocks of the leaf at enumeration position c
    const uint32_t* leaf_by_cache;
    const uint32_t* first_b;
    const uint32_t* last_b;
    __device__ __forceinline__ uint32_t operator()(size_t c) const {
        const uint32_t leaf = leaf_by_cache[c];
        return last_b[leaf] - first_b[leaf];
    }
};

constexpr int REFSTART_THREADS = 256;
constexpr int REFSTART_ROWS = 16;                                   // rows of 32 slots per warp
constexpr int REFSTART_CHUNK = REFSTART_THREADS * REFSTART_ROWS;    // arranged slots per CTA
constexpr int REFSTART_MAX_RANKS = 2048;

__global__ void block_arrange2_kernel(uint32_t nb, const uint32_t* __restrict__ blk_leaf, const int32_t* __restrict__ blk_pose,
                                      const uint32_t* __restrict__ blk_start, const int32_t* __restrict__ pose_rank,
                                      const uint32_t* __restrict__ cache_rank, const uint32_t* __restrict__ off_c,
                                      const uint32_t* __restrict__ first_b, uint32_t* __restrict__ a_rank,
                                      uint32_t* __restrict__ a_size, uint32_t* __restrict__ a_blk, int32_t* __restrict__ blk_size) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    const uint32_t leaf = blk_leaf[b];
    const uint32_t dst = off_c[cache_rank[leaf]] + (b - first_b[leaf]);
    const uint32_t sz = blk_start[b + 1] - blk_start[b];
    a_rank[dst] = (uint32_t)pose_rank[blk_pose[b]];
    a_size[dst] = sz;
    a_blk[dst] = b;
    blk_size[b] = (int32_t)sz;
}

__global__ void __launch_bounds__(REFSTART_THREADS) refstart_count_kernel(uint32_t nb, const uint32_t* __restrict__ a_rank,
                                                                          const uint32_t* __restrict__ a_size, int n_ranks,
                                                                          uint32_t n_chunks, uint32_t* __restrict__ table,
                                                                          uint32_t* __restrict__ max_size) {
    extern __shared__ uint32_t s_cnt[];
    for (int p = threadIdx.x; p < n_ranks; p += REFSTART_THREADS) s_cnt[p] = 0u;
    __syncthreads();
    const uint32_t base = blockIdx.x * REFSTART_CHUNK;
    uint32_t big = 0u;
    uint32_t rk[REFSTART_ROWS], sz[REFSTART_ROWS];  // all loads of the thread first, then the shared-memory atomics
#pragma unroll
    for (int k = 0; k < REFSTART_ROWS; ++k) {
        const uint32_t j = base + (uint32_t)k * REFSTART_THREADS + threadIdx.x;
        rk[k] = j < nb ? a_rank[j] : 0u;
        sz[k] = j < nb ? a_size[j] : 0u;
    }
#pragma unroll
    for (int k = 0; k < REFSTART_ROWS; ++k) {
        if (sz[k]) atomicAdd(&s_cnt[rk[k]], sz[k]);  // (a block holds at least one point: size 0 = past the end)
        big = sz[k] > big ? sz[k] : big;
    }
    if (max_size) {  // the largest block of the table, on the way (the RANSAC launch sizes its shared memory by it)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const uint32_t t = __shfl_xor_sync(0xffffffffu, big, o);
            big = t > big ? t : big;
        }
        if ((threadIdx.x & 31) == 0 && big) atomicMax(max_size, big);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < n_ranks; p += REFSTART_THREADS) table[(size_t)p * n_chunks + blockIdx.x] = s_cnt[p];
}

// start_by_rank != nullptr: this forest holds only a part of the grid (multi-GPU slab partition) and the host supplied the
// batch-global index of the first local point of every pose rank; otherwise the start of a rank inside its batch of `ppb`
// consecutive ranks comes out of the scanned table itself.
__global__ void __launch_bounds__(REFSTART_THREADS) refstart_assign_kernel(uint32_t nb, uint32_t K, uint32_t ppb,
                                                                           const uint32_t* __restrict__ a_rank,
                                                                           const uint32_t* __restrict__ a_size,
                                                                           const uint32_t* __restrict__ a_blk, int n_ranks,
                                                                           uint32_t n_chunks, const uint32_t* __restrict__ scanned,
                                                                           const long long* __restrict__ start_by_rank,
                                                                           long long* __restrict__ blk_ref_start) {
    extern __shared__ __align__(8) unsigned char rs_smem[];
    long long* s_base = reinterpret_cast<long long*>(rs_smem);                        // [n_ranks] start of the rank in this chunk
    uint32_t* s_w = reinterpret_cast<uint32_t*>(rs_smem + 8 * (size_t)n_ranks);       // [8 warps][n_ranks]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t chunk = blockIdx.x;
    for (int p = threadIdx.x; p < n_ranks; p += REFSTART_THREADS) {
        const uint32_t here = scanned[(size_t)p * n_chunks + chunk];
        if (start_by_rank)
            s_base[p] = start_by_rank[p] + (long long)(here - scanned[(size_t)p * n_chunks]);
        else
            s_base[p] = (long long)(here - scanned[(size_t)(p / ppb * ppb) * n_chunks]);
    }
    for (int i = threadIdx.x; i < 8 * n_ranks; i += REFSTART_THREADS) s_w[i] = 0u;
    __syncthreads();
    uint32_t* wh = s_w + (size_t)warp * n_ranks;
    const uint32_t wbase = chunk * REFSTART_CHUNK + (uint32_t)warp * (32 * REFSTART_ROWS);
    const uint32_t lt = (1u << lane) - 1u;
    // (only the in-chunk prefixes stay in registers between the two phases; ranks and sizes are read again for the
    // stores - 103 registers and 23 % occupancy when all three were kept)
    uint32_t local[REFSTART_ROWS];
    constexpr int RG = 8;  // rows whose loads are issued together, ahead of the warp-synchronous ranking
    static_assert(REFSTART_ROWS % RG == 0, "rows come in groups");
#pragma unroll
    for (int r0 = 0; r0 < REFSTART_ROWS; r0 += RG) {
    uint32_t rks[RG], szs[RG];
#pragma unroll
    for (int g = 0; g < RG; ++g) {
        const uint32_t j = wbase + 32 * (r0 + g) + lane;
        rks[g] = j < nb ? a_rank[j] : 0xffffffffu;
        szs[g] = j < nb ? a_size[j] : 0u;
    }
#pragma unroll
    for (int g = 0; g < RG; ++g) {
        const int r = r0 + g;
        const uint32_t rk = rks[g], sz = szs[g];
        // the lanes with this lane's rank, and the sizes of those before it (slot order = lane order): a few iterations,
        // consecutive arranged slots are (leaf, pose ascending), so a rank repeats about once per leaf
        const uint32_t peers = __match_any_sync(0xffffffffu, rk);
        uint32_t below = peers & lt, mine = 0u;
        while (__any_sync(0xffffffffu, below != 0u)) {
            const int src = below ? (__ffs(below) - 1) : lane;
            const uint32_t v = __shfl_sync(0xffffffffu, sz, src);
            if (below) {
                mine += v;
                below &= below - 1u;
            }
        }
        const int leader = __ffs(peers) - 1, last = 31 - __clz(peers);
        const uint32_t group = __shfl_sync(0xffffffffu, mine + sz, last);  // points of the whole group
        uint32_t old = 0u;
        if (rk != 0xffffffffu && lane == leader) {
            old = wh[rk];
            wh[rk] = old + group;
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        local[r] = old + mine;
        __syncwarp();
    }
    }  // groups of rows
    __syncthreads();
    for (int p = threadIdx.x; p < n_ranks; p += REFSTART_THREADS) {  // exclusive prefix over the eight warps
        uint32_t run = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const uint32_t t = s_w[(size_t)w * n_ranks + p];
            s_w[(size_t)w * n_ranks + p] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < REFSTART_ROWS; ++r) {
        const uint32_t j = wbase + 32 * r + lane;
        if (j < nb && a_size[j] >= K) {
            const uint32_t rk = a_rank[j];
            blk_ref_start[a_blk[j]] = s_base[rk] + (long long)(wh[rk] + local[r]);
        }
    }
}

__global__ void key_to_rank_kernel(uint32_t nb, const uint64_t* __restrict__ keys, int shift, uint32_t* __restrict__ ranks) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nb) ranks[j] = (uint32_t)(keys[j] >> shift);
}

__global__ void block_sizes_ref_kernel(uint32_t nb, const uint32_t* __restrict__ ref_order, const uint32_t* __restrict__ blk_start,
                                       uint32_t* __restrict__ sizes_ref, uint32_t* __restrict__ refpos,
                                       int32_t* __restrict__ blk_size) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    uint32_t b = ref_order[j];
    uint32_t sz = blk_start[b + 1] - blk_start[b];
    sizes_ref[j] = sz;
    refpos[b] = j;
    blk_size[b] = (int32_t)sz;
}

// sharded form: start = (batch-global index of the pose's first local point) + offset among the local points of the pose
__global__ void work_refstart_sharded_kernel(uint32_t nb, int K, const uint32_t* __restrict__ ref_order,
                                             const uint32_t* __restrict__ sorted_rank, const uint32_t* __restrict__ sizes_ref,
                                             const uint32_t* __restrict__ refstart, const uint32_t* __restrict__ rank_base,
                                             const long long* __restrict__ start_by_rank, long long* __restrict__ blk_ref_start) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb || sizes_ref[j] < (uint32_t)K) return;
    const uint32_t rk = sorted_rank[j];
    blk_ref_start[ref_order[j]] = start_by_rank[rk] + ((long long)refstart[j] - (long long)rank_base[rk]);
}

// block sizes in block-table order and in reference order (one pass)
__global__ void block_sizes_kernel(uint32_t nb, const uint32_t* __restrict__ blk_start, const uint32_t* __restrict__ ref_order,
                                   int32_t* __restrict__ blk_size, uint32_t* __restrict__ sizes_ref) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    blk_size[j] = (int32_t)(blk_start[j + 1] - blk_start[j]);
    const uint32_t b = ref_order[j];
    sizes_ref[j] = blk_start[b + 1] - blk_start[b];
}

// Work list of the RANSAC launch as the functors of ONE packed 64-bit scan over the block table: a block with at least
// K points contributes (1 << 32 | size); the exclusive prefix of block b is (work index << 32 | first packed point),
// which the output functor scatters straight into the work list.  Grand total = (n_work << 32 | n_packed).
struct WorkIn {
    const uint32_t* blk_start;
    uint32_t K;
    __device__ unsigned long long operator()(size_t b) const {
        const uint32_t sz = blk_start[b + 1] - blk_start[b];
        return sz >= K ? ((1ull << 32) | (unsigned long long)sz) : 0ull;
    }
};
struct WorkOut {
    uint32_t* work;
    uint32_t* pk_start;
    __device__ void operator()(size_t b, unsigned long long ex, unsigned long long v) const {
        if (v) {
            work[ex >> 32] = (uint32_t)b;
            pk_start[ex >> 32] = (uint32_t)ex;
        }
    }
};

// first point of every batch of `ppb` consecutive pose ranks: the reference positions are sorted by pose rank, so the
// first block of batch t is found by bisection
__global__ void batch_base_search_kernel(int n_batches, int ppb, uint32_t nb, const uint32_t* __restrict__ sorted_rank,
                                         const uint32_t* __restrict__ refstart, uint32_t* __restrict__ batch_base) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_batches) return;
    const uint32_t want = (uint32_t)t * (uint32_t)ppb;
    uint32_t lo = 0, hi = nb;
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (sorted_rank[mid] < want)
            lo = mid + 1;
        else
            hi = mid;
    }
    batch_base[t] = lo < nb ? refstart[lo] : 0u;
}

// reference start index (cuda_ransac.py:65-67, per batch) of the blocks that will be fitted; the others never use it
__global__ void work_refstart_kernel(uint32_t nb, int K, int ppb, const uint32_t* __restrict__ ref_order,
                                     const uint32_t* __restrict__ sorted_rank, const uint32_t* __restrict__ sizes_ref,
                                     const uint32_t* __restrict__ refstart, const uint32_t* __restrict__ batch_base,
                                     long long* __restrict__ blk_ref_start) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb || sizes_ref[j] < (uint32_t)K) return;
    blk_ref_start[ref_order[j]] = (long long)refstart[j] - (long long)batch_base[sorted_rank[j] / (uint32_t)ppb];
}

// K5b: gather ONLY the points of the fitted blocks, block after block, so that every block is one contiguous float64
// run for the TMA staging (most (pose, leaf) blocks of a LiDAR map are smaller than K and never reach the kernel).
// Eight lanes per work item; a flat thread-per-double copy inside it.
__global__ void __launch_bounds__(256) gather_blocks_kernel(uint32_t n_work, const uint32_t* __restrict__ work,
                                                            const uint32_t* __restrict__ blk_start, const int32_t* __restrict__ blk_size,
                                                            const uint32_t* __restrict__ pk_start, const double* __restrict__ xyz,
                                                            const uint32_t* __restrict__ perm, double* __restrict__ out) {
    // 8 lanes per work item (fitted blocks of a LiDAR map hold ~7 points = 21 doubles)
    const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (w >= n_work) return;
    const uint32_t sub = threadIdx.x & 7u;
    const uint32_t b = work[w];
    const uint32_t src0 = blk_start[b], dst0 = pk_start[w];
    const uint32_t n3 = (uint32_t)blk_size[b] * 3u;
    // four elements per lane and round: their rank loads, then their coordinate loads, then the stores (a loop of
    // dependent rank -> coordinate loads paid two trips to memory per element)
    for (uint32_t e0 = sub; e0 < n3; e0 += 32) {
        uint32_t r[4];
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t e = e0 + 8u * u;
            r[u] = e < n3 ? perm[src0 + e / 3u] : 0u;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t e = e0 + 8u * u;
            v[u] = e < n3 ? xyz[(size_t)r[u] * 3 + (e - (e / 3u) * 3u)] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t e = e0 + 8u * u;
            if (e < n3) out[(size_t)dst0 * 3 + e] = v[u];
        }
    }
}

__global__ void ransac_snapshot_kernel(uint32_t nb, const uint32_t* __restrict__ ref_order, const int32_t* __restrict__ blk_pose,
                                       const uint32_t* __restrict__ blk_leaf, const uint32_t* __restrict__ cache_rank,
                                       const int32_t* __restrict__ blk_size, const float* __restrict__ plane,
                                       const int32_t* __restrict__ best, const int32_t* __restrict__ best_count,
                                       int32_t* __restrict__ o_pose, int32_t* __restrict__ o_leaf, int32_t* __restrict__ o_size,
                                       float* __restrict__ o_plane, int32_t* __restrict__ o_best,
                                       int32_t* __restrict__ o_count, int K) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    uint32_t b = ref_order[j];
    o_pose[j] = blk_pose[b];
    o_leaf[j] = (int32_t)cache_rank[blk_leaf[b]];
    o_size[j] = blk_size[b];
    if (plane) {
        const bool fitted = blk_size[b] >= K;  // the result rows of smaller blocks were never written (nor initialised)
        for (int c = 0; c < 4; ++c) o_plane[(size_t)j * 4 + c] = fitted ? plane[(size_t)b * 4 + c] : 0.0f;
        o_best[j] = fitted ? best[b] : -1;
        o_count[j] = fitted ? best_count[b] : 0;
    }
}

// per pose: number of non-empty blocks (= leaves with points of that pose) and number of points.  Millions of blocks fall on
// a few hundred poses, so the counters are privatised in shared memory (one flush of the non-zero ones per CTA); poses
// beyond the shared table go straight to global atomics.
constexpr int POSE_COUNT_SMEM = 2048;
__global__ void __launch_bounds__(256) pose_block_counts_kernel(uint32_t nb, const uint32_t* __restrict__ blk_start,
                                                                const int32_t* __restrict__ blk_pose, int n_poses,
                                                                unsigned long long* __restrict__ counts /*[P][3]*/) {
    __shared__ unsigned int s_blocks[POSE_COUNT_SMEM], s_points[POSE_COUNT_SMEM];
    const int np = n_poses < POSE_COUNT_SMEM ? n_poses : POSE_COUNT_SMEM;
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        s_blocks[p] = 0u;
        s_points[p] = 0u;
    }
    __syncthreads();
    // a CTA never sees more than 2^32 points: its share of the block table is bounded by the point count (< 2^31)
    for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
        const int p = blk_pose[b];
        const uint32_t sz = blk_start[b + 1] - blk_start[b];
        if (p < np) {
            atomicAdd(&s_blocks[p], 1u);
            atomicAdd(&s_points[p], sz);
        } else {
            atomicAdd(&counts[(size_t)p * 3 + 0], 1ull);
            atomicAdd(&counts[(size_t)p * 3 + 1], (unsigned long long)sz);
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < np; p += blockDim.x) {
        if (s_blocks[p]) {
            atomicAdd(&counts[(size_t)p * 3 + 0], (unsigned long long)s_blocks[p]);
            atomicAdd(&counts[(size_t)p * 3 + 1], (unsigned long long)s_points[p]);
        }
    }
}

__global__ void cell_leaf_count_kernel(uint32_t L, const uint32_t* __restrict__ lcell, uint32_t* __restrict__ cell_nl) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < L) atomicAdd(&cell_nl[lcell[k]], 1u);
}

__global__ void pose_node_counts_kernel(uint32_t ncp, const uint32_t* __restrict__ cp_cell, const int32_t* __restrict__ cp_pose,
                                        const uint32_t* __restrict__ cell_nl, unsigned long long* __restrict__ counts) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ncp) return;
    const uint32_t nl = cell_nl[cp_cell[j]];  // leaves = 1 + 7 * internal  ->  nodes = leaves + (leaves - 1) / 7
    atomicAdd(&counts[(size_t)cp_pose[j] * 3 + 2], (unsigned long long)(nl + (nl - 1) / 7));
}

__global__ void cell_export_kernel(uint32_t C, const uint64_t* __restrict__ cell_key, KeyParams kp,
                                   const uint32_t* __restrict__ cell_leaf_begin, long long* __restrict__ q_out,
                                   double* __restrict__ corner_out, long long* __restrict__ nodes_out,
                                   long long* __restrict__ leaf_begin_out) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > C) return;
    if (leaf_begin_out) leaf_begin_out[c] = cell_leaf_begin[c];
    if (c == C) return;
    long long q[3] = {0, 0, 0};
    if (!kp.single_cell) unpack_cell(kp, cell_key[c], q);
    for (int a = 0; a < 3; ++a) {
        if (q_out) q_out[(size_t)c * 3 + a] = q[a];
        if (corner_out) corner_out[(size_t)c * 3 + a] = cell_corner_coord(q[a], kp.corner[a], kp.edge, kp.single_cell);
    }
    if (nodes_out) {
        const long long nl = (long long)cell_leaf_begin[c + 1] - (long long)cell_leaf_begin[c];
        nodes_out[c] = nl + (nl - 1) / 7;
    }
}

// selection flags over positions (export of one pose / all poses)
__global__ void select_pose_kernel(uint32_t n, const uint32_t* __restrict__ blk_of_pos, const int32_t* __restrict__ blk_pose,
                                   int pose, uint32_t* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flags[i] = (pose < 0 || blk_pose[blk_of_pos[i]] == pose) ? 1u : 0u;
}

__global__ void select_blocks_ref_kernel(uint32_t nb, const uint32_t* __restrict__ ref_order, const int32_t* __restrict__ blk_pose,
                                         const uint32_t* __restrict__ blk_start, int pose, uint32_t* __restrict__ sel_sizes) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nb) return;
    uint32_t b = ref_order[j];
    sel_sizes[j] = (pose < 0 || blk_pose[b] == pose) ? blk_start[b + 1] - blk_start[b] : 0u;
}

// writes the selected points; dst = dfs ? scan over positions : block offset in reference order
__global__ void export_points_kernel(uint32_t n, int pose, int dfs, const uint32_t* __restrict__ blk_of_pos,
                                     const int32_t* __restrict__ blk_pose, const uint32_t* __restrict__ blk_start,
                                     const uint32_t* __restrict__ refpos, const uint32_t* __restrict__ ref_off,
                                     const uint32_t* __restrict__ pos_scan, const uint32_t* __restrict__ perm,
                                     const double* __restrict__ xyz, const uint32_t* __restrict__ leaf_of,
                                     const uint32_t* __restrict__ lcell, const uint32_t* __restrict__ seg_start,
                                     const int64_t* __restrict__ seg_first, int n_seg, const uint8_t* __restrict__ mask,
                                     double* __restrict__ o_xyz, long long* __restrict__ o_idx, int32_t* __restrict__ o_cell,
                                     uint8_t* __restrict__ o_mask) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = blk_of_pos[i];
    if (pose >= 0 && blk_pose[b] != pose) return;
    const uint32_t dst = dfs ? pos_scan[i] : ref_off[refpos[b]] + (i - blk_start[b]);
    const uint32_t r = perm[i];
    if (o_xyz) {
        o_xyz[(size_t)dst * 3 + 0] = xyz[(size_t)r * 3 + 0];
        o_xyz[(size_t)dst * 3 + 1] = xyz[(size_t)r * 3 + 1];
        o_xyz[(size_t)dst * 3 + 2] = xyz[(size_t)r * 3 + 2];
    }
    if (o_idx) {
        const int s = seg_of_rank(seg_start, n_seg, r);
        o_idx[dst] = (long long)(r - seg_start[s]) + seg_first[s];
    }
    if (o_cell) o_cell[dst] = (int32_t)lcell[leaf_of[i]];
    if (o_mask) o_mask[dst] = mask ? mask[i] : 0;
}

// ---------------------------------------------------------------------------------------------
// block order of the reference: (pose rank, leaf enumeration order)   grid.py:173-191, 217-232
// ---------------------------------------------------------------------------------------------
void Forest::compute_ref_order(const int32_t* pose_rank_host, DevBuf<uint32_t>& ref_order, DevBuf<int32_t>& d_pose_rank,
                               DevBuf<uint32_t>* sorted_rank) {
    ensure_order();
    ensure_blocks();
    std::vector<int32_t> pr(std::max(n_poses, 1));
    for (int p = 0; p < n_poses; ++p) pr[p] = pose_rank_host ? pose_rank_host[p] : p;
    int max_rank = 0;
    for (int p = 0; p < n_poses; ++p) {
        OL_REQUIRE(pr[p] >= 0, OL_ERR_INVALID, "negative pose rank");
        max_rank = std::max(max_rank, pr[p]);
    }
    d_pose_rank.reset(ctx, pr.size());
    h2d(ctx, d_pose_rank.get(), pr.data(), pr.size());  // pageable source: staged before the call returns
    if (NB == 0) {
        ref_order.reset(ctx, 0);
        return;
    }
    DevBuf<uint32_t> k0(ctx, NB), k1(ctx, NB), v0(ctx, NB), v1(ctx, NB), cnt_c(ctx, L), off_c(ctx, L), first_b(ctx, L);
    cnt_c.zero();
    {
        ProfScope ps(ctx, "ransac_prep");
        leaf_block_count_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, blk_leaf.get(), cache_rank.get(), cnt_c.get(), first_b.get());
        OL_CHECK_LAUNCH();
    }
    exclusive_scan_u32(ctx, cnt_c.get(), off_c.get(), L, nullptr);
    if (history_active()) {
        std::vector<int32_t> pe(std::max(n_poses, 1), 0);
        for (int p = 0; p < n_poses && p < (int)pose_epoch.size(); ++p) pe[p] = pose_epoch[p];
        DevBuf<int32_t> d_pe(ctx, pe.size());
        h2d(ctx, d_pe.get(), pe.data(), pe.size());
        const int cell_bits = bit_length_u64(C ? C - 1 : 0), epoch_bits = bit_length_u64((uint64_t)n_subdivide_calls),
                  rank_bits = bit_length_u64((uint64_t)max_rank);
        OL_REQUIRE(cell_bits + epoch_bits + rank_bits <= 64, OL_ERR_RANGE, "block order key does not fit 64 bits");
        DevBuf<uint64_t> kk0(ctx, NB), kk1(ctx, NB);
        {
            ProfScope ps(ctx, "ransac_prep");
            block_arrange_history_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, blk_leaf.get(), blk_pose.get(), d_pose_rank.get(),
                                                                           cache_rank.get(), off_c.get(), first_b.get(), lcell.get(),
                                                                           lparent.get(), iepoch.get(), d_pe.get(), cell_bits,
                                                                           epoch_bits, kk0.get(), v0.get());
            OL_CHECK_LAUNCH();
        }
        const int w = radix_sort_pairs<uint64_t>(ctx, kk0.get(), kk1.get(), v0.get(), v1.get(), NB, 0, cell_bits + epoch_bits + rank_bits);
        ref_order.swap(w ? v1 : v0);
        if (sorted_rank) {
            key_to_rank_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, w ? kk1.get() : kk0.get(), cell_bits + epoch_bits, k0.get());
            OL_CHECK_LAUNCH();
            sorted_rank->swap(k0);
        }
        return;
    }
    {
        ProfScope ps(ctx, "ransac_prep");
        block_arrange_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, blk_leaf.get(), blk_pose.get(), d_pose_rank.get(), cache_rank.get(),
                                                               off_c.get(), first_b.get(), k0.get(), v0.get());
        OL_CHECK_LAUNCH();
    }
    const int w = radix_sort_pairs<uint32_t>(ctx, k0.get(), k1.get(), v0.get(), v1.get(), NB, 0, bit_length_u64((uint64_t)max_rank));
    ref_order.swap(w ? v1 : v0);
    if (sorted_rank) sorted_rank->swap(w ? k1 : k0);  // pose rank of the block at every reference position
}

// ---------------------------------------------------------------------------------------------
// Grid.map_leaf_points_cuda_ransac (grid.py:124-215)
// ---------------------------------------------------------------------------------------------
void Forest::ransac(const double* table_host, int H, int K, double threshold, const int32_t* pose_rank, int ppb, bool apply,
                    uint32_t flags, const int64_t* pose_start) {
    OL_REQUIRE(threshold > 0, OL_ERR_INVALID, "Threshold must be positive");
    OL_REQUIRE(H >= 1, OL_ERR_INVALID, "Number of RANSAC hypotheses must be positive");
    OL_REQUIRE(H <= 1024, OL_ERR_INVALID, "Number of RANSAC hypotheses must be <= 1024 because of the CUDA thread limit.");
    OL_REQUIRE(K >= 1 && K <= 64, OL_ERR_INVALID, "initial_points_number must be in 1..64");
    OL_REQUIRE(ppb >= 1, OL_ERR_INVALID, "poses_per_batch must be positive");
    DevBuf<uint32_t> ref_order, sorted_rank;
    DevBuf<int32_t> d_pose_rank;
    int max_rank = 0;
    for (int p = 0; p < n_poses; ++p) {
        OL_REQUIRE(!pose_rank || pose_rank[p] >= 0, OL_ERR_INVALID, "negative pose rank");
        max_rank = std::max(max_rank, pose_rank ? pose_rank[p] : p);
    }
    static const bool no_fast_layout = getenv("OL_RANSAC_SORTED_LAYOUT") != nullptr;  // debug / A-B: always sort the block table
    ensure_order();
    ensure_blocks();
    const bool fast_layout = !no_fast_layout && !history_active() && max_rank < REFSTART_MAX_RANKS && NB > 0;
    if (!fast_layout) compute_ref_order(pose_rank, ref_order, d_pose_rank, &sorted_rank);
    drop_snapshot();
    res_n = NB;
    mask.reset(ctx, A);
    mask.zero();
    mask_n = A;
    if (NB == 0) {
        ransac_valid = true;
        last_ransac_work = 0;
        return;
    }
    const int n_batches = max_rank / ppb + 1;
    DevBuf<int32_t> blk_size(ctx, NB);
    DevBuf<long long> blk_ref_start(ctx, NB);
    DevBuf<unsigned long long> d_total(ctx, 1);
    DevBuf<uint32_t> a_rank, a_blk;  // fast layout: the arranged (pose rank, block) tables, kept for a deferred sort
    if (fast_layout) {
        const int n_ranks = max_rank + 1;
        std::vector<int32_t> pr(std::max(n_poses, 1));
        for (int p = 0; p < n_poses; ++p) pr[p] = pose_rank ? pose_rank[p] : p;
        d_pose_rank.reset(ctx, pr.size());
        h2d(ctx, d_pose_rank.get(), pr.data(), pr.size());
        const uint32_t n_chunks = (NB + REFSTART_CHUNK - 1) / REFSTART_CHUNK;
        DevBuf<uint32_t> off_c(ctx, L), span(ctx, (size_t)2 * L), a_size(ctx, NB), table(ctx, (size_t)n_ranks * n_chunks);
        uint32_t* first_b = span.get();
        uint32_t* last_b = span.get() + L;
        DevBuf<long long> d_start;
        a_rank.reset(ctx, NB);
        a_blk.reset(ctx, NB);
        span.zero();
        if (pose_start) {  // multi-GPU: batch-global index of the first local point of every pose rank (see below)
            std::vector<long long> by_rank((size_t)n_ranks, 0);
            for (int p = 0; p < n_poses; ++p) by_rank[pr[p]] = (long long)pose_start[p];
            d_start.reset(ctx, (size_t)n_ranks);
            h2d(ctx, d_start.get(), by_rank.data(), by_rank.size());
        }
        {
            ProfScope ps(ctx, "ransac_prep");
            leaf_block_span_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, blk_leaf.get(), first_b, last_b);
            OL_CHECK_LAUNCH();
        }
        transform_scan<uint32_t>(ctx, LeafSpanIn{leaf_by_cache.get(), first_b, last_b}, ScanPtrOut<uint32_t>{off_c.get()}, L, nullptr);
        {
            ProfScope ps(ctx, "ransac_prep");
            block_arrange2_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, blk_leaf.get(), blk_pose.get(), blk_start.get(), d_pose_rank.get(),
                                                                    cache_rank.get(), off_c.get(), first_b, a_rank.get(),
                                                                    a_size.get(), a_blk.get(), blk_size.get());
            OL_CHECK_LAUNCH();
            refstart_count_kernel<<<n_chunks, REFSTART_THREADS, (size_t)n_ranks * 4, ctx.stream>>>(NB, a_rank.get(), a_size.get(), n_ranks,
                                                                                                  n_chunks, table.get(),
                                                                                                  max_block_known || max_block_enqueued ? nullptr : d_max_block.get());
            OL_CHECK_LAUNCH();
            max_block_enqueued = true;
        }
        exclusive_scan_u32(ctx, table.get(), table.get(), (size_t)n_ranks * n_chunks, nullptr);
        {
            ProfScope ps(ctx, "ransac_prep");
            const size_t smem = (size_t)n_ranks * (8 + 8 * 4);
            static bool attr_set = false;
            if (!attr_set) {
                OL_CUDA(cudaFuncSetAttribute(refstart_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             REFSTART_MAX_RANKS * (8 + 8 * 4)));
                attr_set = true;
            }
            refstart_assign_kernel<<<n_chunks, REFSTART_THREADS, smem, ctx.stream>>>(NB, (uint32_t)K, (uint32_t)ppb, a_rank.get(),
                                                                                    a_size.get(), a_blk.get(), n_ranks, n_chunks,
                                                                                    table.get(), d_start.get(), blk_ref_start.get());
            OL_CHECK_LAUNCH();
        }
    } else {
    DevBuf<uint32_t> sizes_ref(ctx, NB), refstart(ctx, NB), batch_base(ctx, n_batches);
    {
        ProfScope ps(ctx, "ransac_prep");
        block_sizes_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, blk_start.get(), ref_order.get(), blk_size.get(), sizes_ref.get());
        OL_CHECK_LAUNCH();
    }
    exclusive_scan_u32(ctx, sizes_ref.get(), refstart.get(), NB, nullptr);
    if (pose_start) {
        // This forest holds only a PART of the grid (multi-GPU, slab partition: partition.cu).  The reference's start index
        // of a block (cuda_ransac.py:65-67, per batch) = pose_start[pose] - the batch-global index of the first point of
        // that pose held HERE, supplied by the host from one all-gather of the per-rank pose sizes - plus the block's
        // offset among this forest's points of the pose.
        const int n_ranks = max_rank + 1;
        std::vector<long long> by_rank((size_t)n_ranks, 0);
        for (int p = 0; p < n_poses; ++p) by_rank[pose_rank ? pose_rank[p] : p] = (long long)pose_start[p];
        DevBuf<long long> d_start(ctx, (size_t)n_ranks);
        DevBuf<uint32_t> rank_base(ctx, (size_t)n_ranks);
        h2d(ctx, d_start.get(), by_rank.data(), by_rank.size());
        ProfScope ps(ctx, "ransac_prep");
        batch_base_search_kernel<<<nblk((size_t)n_ranks), 256, 0, ctx.stream>>>(n_ranks, 1, NB, sorted_rank.get(), refstart.get(),
                                                                                rank_base.get());
        OL_CHECK_LAUNCH();
        work_refstart_sharded_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, K, ref_order.get(), sorted_rank.get(), sizes_ref.get(),
                                                                       refstart.get(), rank_base.get(), d_start.get(),
                                                                       blk_ref_start.get());
        OL_CHECK_LAUNCH();
    } else {
        ProfScope ps(ctx, "ransac_prep");
        batch_base_search_kernel<<<nblk((size_t)n_batches), 256, 0, ctx.stream>>>(n_batches, ppb, NB, sorted_rank.get(), refstart.get(),
                                                                                  batch_base.get());
        OL_CHECK_LAUNCH();
        work_refstart_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, K, ppb, ref_order.get(), sorted_rank.get(), sizes_ref.get(),
                                                               refstart.get(), batch_base.get(), blk_ref_start.get());
        OL_CHECK_LAUNCH();
    }
    }  // sorted layout
    // work list + packed layout of the fitted blocks' points: one scan, one read-back (with the largest block size)
    const size_t work_max = std::min<size_t>(NB, (size_t)A / (size_t)K);
    DevBuf<uint32_t> work(ctx, work_max + 1), pk_start(ctx, work_max + 1);
    // (the number of fitted blocks, their point total and the largest block size are POSTED by the scan kernel; the
    // result tables are initialised behind it, so the GPU has work while the host reads the mailbox)
    enqueue_max_block();  // (a no-op when the batch layout above already left the maximum in d_max_block)
    const Mail mail = mail_open(MAIL_WORK, max_block_known ? nullptr : d_max_block.get());
    transform_scan<unsigned long long>(ctx, WorkIn{blk_start.get(), (uint32_t)K}, WorkOut{work.get(), pk_start.get()}, NB, d_total.get(),
                                       "ransac_prep", mail);
    DevBuf<double> table(ctx, (size_t)H * K);
    h2d(ctx, table.get(), table_host, (size_t)H * K);
    // per-block results: written by the kernels for exactly the fitted blocks (size >= K) and never read for the others
    // (ransac_snapshot_kernel tells them apart by their size), so the 24 bytes per block are not initialised - on a LiDAR
    // map 93 % of the 41 M blocks are too small to be fitted, and clearing their rows cost a gigabyte of stores per step
    DevBuf<float> plane(ctx, (size_t)NB * 4);
    DevBuf<int32_t> best(ctx, NB), best_count(ctx, NB);
    const MailResult posted = mail_take(mail);
    const unsigned long long packed_total = posted.total;
    if (!max_block_known) {
        max_block = posted.aux;
        max_block_known = true;
    }
    const uint32_t n_work = (uint32_t)(packed_total >> 32);
    const uint64_t n_packed = packed_total & 0xffffffffull;
    DevBuf<double> pleaf(ctx, (size_t)n_packed * 3 + 2);
    if (n_work) {
        ProfScope ps(ctx, "gather_points", (double)n_packed);
        gather_blocks_kernel<<<nblk((size_t)n_work * 8), 256, 0, ctx.stream>>>(n_work, work.get(), blk_start.get(), blk_size.get(),
                                                                                pk_start.get(), P64.get(), perm.get(), pleaf.get());
        OL_CHECK_LAUNCH();
    }
    {
        ProfScope ps(ctx, "ransac_kernel", (double)n_work);
        launch_ransac(ctx, pleaf.get(), (int64_t)n_packed, blk_start.get(), blk_size.get(), blk_ref_start.get(), work.get(), pk_start.get(),
                      n_work, max_block,
                      table.get(), H, K, threshold, mask.get(), plane.get(), best.get(), best_count.get(), flags);
    }
    last_ransac_work = n_work;
    // The per-block result table in reference order (res_*) is only gathered when somebody asks for it
    // (materialize_snapshot): keep the raw inputs of that gather.  The block table itself is consumed by the
    // mask application, so it is moved, not copied, when the mask is applied right away.
    sn_ref_order.swap(ref_order);
    sn_arr_rank.swap(a_rank);  // fast layout: the reference order is sorted out of these when a result table is asked for
    sn_arr_blk.swap(a_blk);
    sn_rank_bits = bit_length_u64((uint64_t)max_rank);
    sn_K = K;
    sn_size.swap(blk_size);
    sn_plane.swap(plane);
    sn_best.swap(best);
    sn_count.swap(best_count);
    if (apply) {
        sn_pose.swap(blk_pose);
        sn_leaf.swap(blk_leaf);
        blocks_valid = false;
    } else {
        sn_pose.reset(ctx, NB);
        sn_leaf.reset(ctx, NB);
        d2d(ctx, sn_pose.get(), blk_pose.get(), NB);
        d2d(ctx, sn_leaf.get(), blk_leaf.get(), NB);
    }
    snap_pending = true;
    ransac_valid = true;
    if (apply) {
        apply_mask();  // its read-back also brings the error word (note_ransac_flags)
    } else {
        note_ransac_flags(read_u32(d_err.get()));
    }
}

// RANSAC bits of the device error word: a sample index that left its block is clamped and only recorded
// (ol_forest_stats.sample_oob_seen); a violated pre-filter interval (verify mode) is an internal error
void Forest::note_ransac_flags(uint32_t e) {
    if (!(e & (DEVERR_SAMPLE_OOB | DEVERR_FILTER_BOUND))) return;
    const bool bound = (e & DEVERR_FILTER_BOUND) != 0;
    if (e & DEVERR_SAMPLE_OOB) sample_oob_seen = true;
    e &= ~(uint32_t)(DEVERR_SAMPLE_OOB | DEVERR_FILTER_BOUND);
    OL_CUDA(cudaMemcpyAsync(d_err.get(), &e, 4, cudaMemcpyHostToDevice, ctx.stream));
    ctx.sync();
    OL_REQUIRE(!bound, OL_ERR_INTERNAL, "RANSAC verify: an exact inlier count left its pre-filter interval");
}

void Forest::apply_mask() {
    OL_REQUIRE(ransac_valid && mask_n == A, OL_ERR_STATE, "no RANSAC mask to apply");
    DevBuf<uint8_t> m;
    m.swap(mask);
    apply_keep(m.get());
    ransac_valid = false;
}

__global__ void keep_from_pose_mask_kernel(uint32_t n, int pose, const uint32_t* __restrict__ blk_of_pos,
                                          const int32_t* __restrict__ blk_pose, const uint32_t* __restrict__ blk_start,
                                          const uint32_t* __restrict__ refpos, const uint32_t* __restrict__ ref_off,
                                          const uint8_t* __restrict__ pose_mask, uint8_t* __restrict__ keep_pos) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = blk_of_pos[i];
    uint8_t keep = 1;
    if (blk_pose[b] == pose) keep = pose_mask[ref_off[refpos[b]] + (i - blk_start[b])];
    keep_pos[i] = keep;
}

// OctreeManager.apply_mask (octree_manager.py:173-180): host mask in the pose's leaf order
void Forest::apply_pose_mask(const int32_t* pose_rank, int pose, const uint8_t* mask_host, int64_t n) {
    OL_REQUIRE(pose >= 0 && pose < n_poses, OL_ERR_POSE, "unknown pose index " + std::to_string(pose));
    DevBuf<uint32_t> ref_order;
    DevBuf<int32_t> d_pose_rank;
    compute_ref_order(pose_rank, ref_order, d_pose_rank);
    if (NB == 0) {
        OL_REQUIRE(n == 0, OL_ERR_INVALID, "mask length does not match the pose's point count");
        return;
    }
    DevBuf<uint32_t> sizes_ref(ctx, NB), sel(ctx, NB), refpos(ctx, NB), ref_off(ctx, NB);
    DevBuf<int32_t> blk_size(ctx, NB);
    DevBuf<unsigned long long> d_total(ctx, 1);
    {
        ProfScope ps(ctx, "ransac_prep");
        block_sizes_ref_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, ref_order.get(), blk_start.get(), sizes_ref.get(), refpos.get(),
                                                                 blk_size.get());
        OL_CHECK_LAUNCH();
    }
    select_blocks_ref_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, ref_order.get(), blk_pose.get(), blk_start.get(), pose,
                                                               sel.get());
    OL_CHECK_LAUNCH();
    exclusive_scan_u32(ctx, sel.get(), ref_off.get(), NB, d_total.get());
    const uint64_t total = read_u64(d_total.get());
    OL_REQUIRE((int64_t)total == n, OL_ERR_INVALID,
               "mask length " + std::to_string(n) + " does not match the pose's point count " + std::to_string(total));
    if (total == 0) return;
    DevBuf<uint8_t> pm(ctx, total), keep_pos(ctx, A);
    h2d(ctx, pm.get(), mask_host, total);
    keep_from_pose_mask_kernel<<<nblk(A), 256, 0, ctx.stream>>>(A, pose, blk_of_pos.get(), blk_pose.get(), blk_start.get(),
                                                                refpos.get(), ref_off.get(), pm.get(), keep_pos.get());
    OL_CHECK_LAUNCH();
    ctx.sync();
    apply_keep(keep_pos.get());
}

// ---------------------------------------------------------------------------------------------
// counters (grid.py:343-362)
// ---------------------------------------------------------------------------------------------
void Forest::pose_counts(int64_t* out_host) {
    ensure_blocks();
    ensure_cell_poses();
    const int P = std::max(n_poses, 1);
    DevBuf<unsigned long long> counts(ctx, (size_t)P * 3);
    counts.zero();
    if (NB) {
        pose_block_counts_kernel<<<std::min<unsigned>(nblk(NB), (unsigned)ctx.num_sms * 8), 256, 0, ctx.stream>>>(
            NB, blk_start.get(), blk_pose.get(), n_poses, counts.get());
        OL_CHECK_LAUNCH();
    }
    if (CP) {
        DevBuf<uint32_t> cell_nl(ctx, C);
        cell_nl.zero();
        cell_leaf_count_kernel<<<nblk(L), 256, 0, ctx.stream>>>(L, lcell.get(), cell_nl.get());
        OL_CHECK_LAUNCH();
        pose_node_counts_kernel<<<nblk(CP), 256, 0, ctx.stream>>>(CP, cp_cell.get(), cp_pose.get(), cell_nl.get(), counts.get());
        OL_CHECK_LAUNCH();
    }
    std::vector<unsigned long long> h((size_t)P * 3);
    d2h(ctx, h.data(), counts.get(), h.size());
    ctx.sync();
    for (int p = 0; p < n_poses; ++p)
        for (int k = 0; k < 3; ++k) out_host[(size_t)p * 3 + k] = (int64_t)h[(size_t)p * 3 + k];
}

// points stored per pose (multi-GPU: the per-rank pose sizes behind the batch-global block starts)
void Forest::pose_point_counts(int64_t* out_host) {
    for (int p = 0; p < n_poses; ++p) out_host[p] = 0;
    if (!any_dead) {  // nothing has been removed yet: the inserted segments say it all, no kernel and no synchronisation
        for (size_t s = 0; s < seg_pose.size(); ++s) out_host[seg_pose[s]] += (int64_t)seg_start[s + 1] - (int64_t)seg_start[s];
        return;
    }
    ensure_blocks();
    const int P = std::max(n_poses, 1);
    DevBuf<unsigned long long> counts(ctx, (size_t)P * 3);
    counts.zero();
    if (NB) {
        pose_block_counts_kernel<<<std::min<unsigned>(nblk(NB), (unsigned)ctx.num_sms * 8), 256, 0, ctx.stream>>>(
            NB, blk_start.get(), blk_pose.get(), n_poses, counts.get());
        OL_CHECK_LAUNCH();
    }
    std::vector<unsigned long long> h((size_t)P * 3);
    d2h(ctx, h.data(), counts.get(), h.size());
    ctx.sync();
    for (int p = 0; p < n_poses; ++p) out_host[p] = (int64_t)h[(size_t)p * 3 + 1];
}

void Forest::stats(ol_forest_stats* s, bool light) {
    if (light) {  // no derived table is (re)built: block fields are -1 when the block table is stale
        ensure_shape();
        if (blocks_valid)
            ensure_max_block();
        else
            ctx.sync();
    } else {
        ensure_blocks();
        ensure_max_block();
        ensure_cell_poses();
    }
    memset(s, 0, sizeof(*s));
    s->n_points_inserted = (int64_t)N;
    s->n_points_alive = A;
    s->n_poses = n_poses;
    s->n_cells = C;
    s->n_cell_poses = cp_valid ? (int64_t)CP : -1;
    s->n_leaves = L;
    s->n_internal = I;
    s->n_blocks = blocks_valid ? (int64_t)NB : -1;
    s->max_block_size = blocks_valid ? (int64_t)max_block : -1;
    s->max_depth_reached = depth_reached;
    s->key_bits = key_bits;
    s->device_bytes_peak = (int64_t)ctx.bytes_peak;
    s->sample_oob_seen = sample_oob_seen ? 1 : 0;
}

std::string Forest::profile_report() {
    ctx.sync();
    struct Acc {
        const char* name;
        int count;
        double ms, units;
    };
    std::vector<Acc> acc;
    for (auto& r : prof.recs) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        bool found = false;
        for (auto& a : acc)
            if (strcmp(a.name, r.name) == 0) {
                a.count += 1;
                a.ms += ms;
                a.units += r.units;
                found = true;
                break;
            }
        if (!found) acc.push_back(Acc{r.name, 1, (double)ms, r.units});
    }
    prof.clear();
    std::string out;
    char line[160];
    for (auto& a : acc) {
        snprintf(line, sizeof(line), "%s %d %.6f %.0f\n", a.name, a.count, a.ms, a.units);
        out += line;
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// exports
// ---------------------------------------------------------------------------------------------
template <typename T>
static void copy_out(Ctx& ctx, T* host, const T* dev, size_t n) {
    if (host && n) d2h(ctx, host, dev, n);
}

void Forest::export_cells(int64_t* q, double* corner, int32_t* first_pose, int64_t* n_nodes, int64_t* leaf_begin) {
    ensure_order();
    ensure_cell_poses();
    DevBuf<long long> dq(ctx, (size_t)C * 3), dn(ctx, C), dl(ctx, (size_t)C + 1);
    DevBuf<double> dc(ctx, (size_t)C * 3);
    cell_export_kernel<<<nblk((size_t)C + 1), 256, 0, ctx.stream>>>(C, cell_key.get(), kp, cell_leaf_begin.get(), dq.get(),
                                                                    dc.get(), dn.get(), dl.get());
    OL_CHECK_LAUNCH();
    copy_out(ctx, (long long*)q, dq.get(), (size_t)C * 3);
    copy_out(ctx, corner, dc.get(), (size_t)C * 3);
    copy_out(ctx, (long long*)n_nodes, dn.get(), C);
    copy_out(ctx, (long long*)leaf_begin, dl.get(), (size_t)C + 1);
    copy_out(ctx, first_pose, cell_first_pose.get(), C);
    ctx.sync();
}

void Forest::export_cell_poses(int32_t* cell, int32_t* pose) {
    build();
    ensure_cell_poses();
    copy_out(ctx, (uint32_t*)cell, cp_cell.get(), CP);
    copy_out(ctx, pose, cp_pose.get(), CP);
    ctx.sync();
}

__global__ void leaf_meta_kernel(uint32_t L, const uint32_t* __restrict__ leaf_by_cache, const uint32_t* __restrict__ lcell,
                                 const uint8_t* __restrict__ ldepth, const int32_t* __restrict__ lparent,
                                 const uint32_t* __restrict__ iepoch, int32_t* __restrict__ o_cell, int32_t* __restrict__ o_depth,
                                 int32_t* __restrict__ o_epoch) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= L) return;
    uint32_t k = leaf_by_cache[j];
    o_cell[j] = (int32_t)lcell[k];
    o_depth[j] = (int32_t)ldepth[k];
    const int32_t par = lparent[k];
    o_epoch[j] = par >= 0 ? (iepoch ? (int32_t)iepoch[par] : 1) : 0;  // subdivide call that split the leaf's parent
}

void Forest::export_leaves(double* corner, double* edge, int32_t* cell, int32_t* depth, int32_t* parent_epoch) {
    ensure_order();
    copy_out(ctx, corner, leaf_corner.get(), (size_t)L * 3);
    copy_out(ctx, edge, leaf_edge.get(), L);
    if ((cell || depth || parent_epoch) && L) {
        DevBuf<int32_t> dc(ctx, L), dd(ctx, L), de(ctx, L);
        leaf_meta_kernel<<<nblk(L), 256, 0, ctx.stream>>>(L, leaf_by_cache.get(), lcell.get(), ldepth.get(), lparent.get(),
                                                          epochs_valid ? iepoch.get() : nullptr, dc.get(), dd.get(), de.get());
        OL_CHECK_LAUNCH();
        copy_out(ctx, cell, dc.get(), L);
        copy_out(ctx, depth, dd.get(), L);
        copy_out(ctx, parent_epoch, de.get(), L);
        ctx.sync();
    }
    ctx.sync();
}

void Forest::export_blocks(const int32_t* pose_rank, int32_t* pose, int32_t* leaf, int32_t* size) {
    DevBuf<uint32_t> ref_order;
    DevBuf<int32_t> d_pose_rank;
    compute_ref_order(pose_rank, ref_order, d_pose_rank);
    if (NB == 0) return;
    DevBuf<uint32_t> sizes_ref(ctx, NB), refpos(ctx, NB);
    DevBuf<int32_t> blk_size(ctx, NB), o_pose(ctx, NB), o_leaf(ctx, NB), o_size(ctx, NB);
    {
        ProfScope ps(ctx, "ransac_prep");
        block_sizes_ref_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, ref_order.get(), blk_start.get(), sizes_ref.get(), refpos.get(),
                                                                 blk_size.get());
        OL_CHECK_LAUNCH();
    }
    {
        ProfScope ps(ctx, "ransac_prep");
        ransac_snapshot_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, ref_order.get(), blk_pose.get(), blk_leaf.get(), cache_rank.get(),
                                                                 blk_size.get(), nullptr, nullptr, nullptr, o_pose.get(), o_leaf.get(),
                                                                 o_size.get(), nullptr, nullptr, nullptr, 0);
        OL_CHECK_LAUNCH();
    }
    copy_out(ctx, pose, o_pose.get(), NB);
    copy_out(ctx, leaf, o_leaf.get(), NB);
    copy_out(ctx, size, o_size.get(), NB);
    ctx.sync();
}

__global__ void scored_flags_kernel(uint32_t n, const int32_t* __restrict__ best, uint32_t* __restrict__ flags) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) flags[j] = best[j] >= 0 ? 1u : 0u;
}

__global__ void scored_compact_kernel(uint32_t n, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ scan_ex,
                                      const int32_t* __restrict__ pose, const int32_t* __restrict__ leaf,
                                      const int32_t* __restrict__ size, const float* __restrict__ plane,
                                      const int32_t* __restrict__ best, const int32_t* __restrict__ count,
                                      int32_t* __restrict__ o_pose, int32_t* __restrict__ o_leaf, int32_t* __restrict__ o_size,
                                      float* __restrict__ o_plane, int32_t* __restrict__ o_best, int32_t* __restrict__ o_count) {
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || !flags[j]) return;
    const uint32_t d = scan_ex[j];
    o_pose[d] = pose[j];
    o_leaf[d] = leaf[j];
    o_size[d] = size[j];
    reinterpret_cast<float4*>(o_plane)[d] = reinterpret_cast<const float4*>(plane)[j];
    o_best[d] = best[j];
    o_count[d] = count[j];
}

void Forest::drop_snapshot() {
    snap_pending = false;
    sn_ref_order.release();
    sn_arr_rank.release();
    sn_arr_blk.release();
    sn_pose.release();
    sn_leaf.release();
    sn_size.release();
    sn_plane.release();
    sn_best.release();
    sn_count.release();
}

// gathers the result table of the last RANSAC run into reference block order (res_*); must run before the
// leaf enumeration order it refers to (cache_rank) is rebuilt
void Forest::materialize_snapshot() {
    if (!snap_pending) return;
    const uint32_t nb = res_n;
    res_pose.reset(ctx, nb);
    res_leaf.reset(ctx, nb);
    res_size.reset(ctx, nb);
    res_plane.reset(ctx, (size_t)nb * 4);
    res_best.reset(ctx, nb);
    res_count.reset(ctx, nb);
    if (nb && !sn_ref_order.get()) {
        // the RANSAC launch took the sort-free batch layout: the reference order of the blocks is produced now, by the
        // stable sort on the pose rank that compute_ref_order does up front on the other path
        OL_REQUIRE(sn_arr_rank.get() && sn_arr_blk.get(), OL_ERR_INTERNAL, "RANSAC snapshot without a block order");
        DevBuf<uint32_t> k1(ctx, nb), v1(ctx, nb);
        const int w = radix_sort_pairs<uint32_t>(ctx, sn_arr_rank.get(), k1.get(), sn_arr_blk.get(), v1.get(), nb, 0, sn_rank_bits);
        sn_ref_order.swap(w ? v1 : sn_arr_blk);
    }
    if (nb) {
        ProfScope ps(ctx, "ransac_prep");
        ransac_snapshot_kernel<<<nblk(nb), 256, 0, ctx.stream>>>(nb, sn_ref_order.get(), sn_pose.get(), sn_leaf.get(), cache_rank.get(),
                                                                 sn_size.get(), sn_plane.get(), sn_best.get(), sn_count.get(),
                                                                 res_pose.get(), res_leaf.get(), res_size.get(), res_plane.get(),
                                                                 res_best.get(), res_count.get(), sn_K);
        OL_CHECK_LAUNCH();
    }
    drop_snapshot();
}

// rows of the last RANSAC run in reference block order; scored_only keeps the blocks that were fitted
int64_t Forest::export_ransac(bool scored_only, bool count_only, int32_t* pose, int32_t* leaf, int32_t* size, float* plane,
                              int32_t* best, int32_t* count) {
    if (!count_only) materialize_snapshot();
    if (!scored_only) {
        if (!count_only) {
            copy_out(ctx, pose, res_pose.get(), res_n);
            copy_out(ctx, leaf, res_leaf.get(), res_n);
            copy_out(ctx, size, res_size.get(), res_n);
            copy_out(ctx, plane, res_plane.get(), (size_t)res_n * 4);
            copy_out(ctx, best, res_best.get(), res_n);
            copy_out(ctx, count, res_count.get(), res_n);
            ctx.sync();
        }
        return res_n;
    }
    if (count_only) return last_ransac_work;
    const uint32_t n = res_n, m = last_ransac_work;
    if (n == 0 || m == 0) return 0;
    DevBuf<uint32_t> flags(ctx, n), scan(ctx, n);
    DevBuf<int32_t> o_pose(ctx, m), o_leaf(ctx, m), o_size(ctx, m), o_best(ctx, m), o_count(ctx, m);
    DevBuf<float> o_plane(ctx, (size_t)m * 4);
    scored_flags_kernel<<<nblk(n), 256, 0, ctx.stream>>>(n, res_best.get(), flags.get());
    OL_CHECK_LAUNCH();
    exclusive_scan_u32(ctx, flags.get(), scan.get(), n, nullptr);
    scored_compact_kernel<<<nblk(n), 256, 0, ctx.stream>>>(n, flags.get(), scan.get(), res_pose.get(), res_leaf.get(),
                                                           res_size.get(), res_plane.get(), res_best.get(), res_count.get(),
                                                           o_pose.get(), o_leaf.get(), o_size.get(), o_plane.get(), o_best.get(),
                                                           o_count.get());
    OL_CHECK_LAUNCH();
    copy_out(ctx, pose, o_pose.get(), m);
    copy_out(ctx, leaf, o_leaf.get(), m);
    copy_out(ctx, size, o_size.get(), m);
    copy_out(ctx, plane, o_plane.get(), (size_t)m * 4);
    copy_out(ctx, best, o_best.get(), m);
    copy_out(ctx, count, o_count.get(), m);
    ctx.sync();
    return m;
}

int64_t Forest::export_points(const int32_t* pose_rank, int pose, int order, double* xyz, int64_t* idx, int32_t* cell,
                              uint8_t* mask_out) {
    OL_REQUIRE(pose < n_poses, OL_ERR_POSE, "unknown pose index " + std::to_string(pose));
    ensure_blocks();
    if (A == 0) return 0;
    const int S = (int)seg_pose.size();
    const bool dfs = order == 1;
    DevBuf<uint32_t> ref_order, refpos, ref_off, pos_scan;
    DevBuf<int32_t> d_pose_rank;
    DevBuf<unsigned long long> d_total(ctx, 1);
    uint64_t total = 0;
    if (dfs) {
        DevBuf<uint32_t> flags(ctx, A);
        pos_scan.reset(ctx, A);
        select_pose_kernel<<<nblk(A), 256, 0, ctx.stream>>>(A, blk_of_pos.get(), blk_pose.get(), pose, flags.get());
        OL_CHECK_LAUNCH();
        exclusive_scan_u32(ctx, flags.get(), pos_scan.get(), A, d_total.get());
        total = read_u64(d_total.get());
    } else {
        compute_ref_order(pose_rank, ref_order, d_pose_rank);
        DevBuf<uint32_t> sizes_ref(ctx, NB), sel(ctx, NB);
        DevBuf<int32_t> blk_size(ctx, NB);
        refpos.reset(ctx, NB);
        ref_off.reset(ctx, NB);
        {
            ProfScope ps(ctx, "ransac_prep");
            block_sizes_ref_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, ref_order.get(), blk_start.get(), sizes_ref.get(),
                                                                     refpos.get(), blk_size.get());
            OL_CHECK_LAUNCH();
        }
        select_blocks_ref_kernel<<<nblk(NB), 256, 0, ctx.stream>>>(NB, ref_order.get(), blk_pose.get(), blk_start.get(), pose,
                                                                   sel.get());
        OL_CHECK_LAUNCH();
        exclusive_scan_u32(ctx, sel.get(), ref_off.get(), NB, d_total.get());
        total = read_u64(d_total.get());
    }
    if (total == 0) return 0;
    const bool want_mask = mask_out && ransac_valid && mask_n == A && !dfs;
    DevBuf<double> o_xyz;
    DevBuf<long long> o_idx;
    DevBuf<int32_t> o_cell;
    DevBuf<uint8_t> o_mask;
    if (xyz) o_xyz.reset(ctx, total * 3);
    if (idx) o_idx.reset(ctx, total);
    if (cell) o_cell.reset(ctx, total);
    if (mask_out) {
        o_mask.reset(ctx, total);
        o_mask.zero();
    }
    export_points_kernel<<<nblk(A), 256, 0, ctx.stream>>>(A, pose, dfs ? 1 : 0, blk_of_pos.get(), blk_pose.get(), blk_start.get(),
                                                          refpos.get(), ref_off.get(), pos_scan.get(), perm.get(), P64.get(),
                                                          leaf_of.get(), lcell.get(), d_seg_start.get(), d_seg_first.get(), S,
                                                          want_mask ? mask.get() : nullptr, o_xyz.get(), o_idx.get(), o_cell.get(),
                                                          o_mask.get());
    OL_CHECK_LAUNCH();
    copy_out(ctx, xyz, o_xyz.get(), total * 3);
    copy_out(ctx, (long long*)idx, o_idx.get(), total);
    copy_out(ctx, cell, o_cell.get(), total);
    copy_out(ctx, mask_out, o_mask.get(), total);
    ctx.sync();
    return (int64_t)total;
}

}  // namespace ol
