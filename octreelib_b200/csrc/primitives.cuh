// Device-wide primitives used by the forest pipeline: exclusive scan (u32), stable LSD radix sort of (key, u32 value)
// pairs (onesweep.cuh below 2^30 pairs, the three-kernel sort in this file above), run segmentation (count + emit).
// All work is enqueued on Ctx::stream; temporaries come from Ctx::alloc.
#pragma once
#include "common.cuh"
#include "onesweep.cuh"

namespace ol {

// =============================================================================================
// exclusive scan (uint32), reduce-then-scan over 4096-element tiles
// =============================================================================================
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) >= o) v += t;
    }
    return v;
}

// exclusive scan of one value per thread across a 256-thread block; returns block total via *total
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t* smem_warp /*[8]*/, uint32_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = warp_inclusive_scan(v);
    if (lane == 31) smem_warp[warp] = inc;
    __syncthreads();
    uint32_t wsum = (lane < 8) ? smem_warp[lane] : 0;
    uint32_t winc = warp_inclusive_scan(wsum);
    uint32_t wbase = __shfl_sync(0xffffffffu, winc - wsum, warp);
    uint32_t tot = __shfl_sync(0xffffffffu, winc, 7);
    __syncthreads();
    if (total) *total = tot;
    return wbase + inc - v;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const uint32_t* __restrict__ in,
                                                                    uint32_t* __restrict__ tile_sums, size_t n) {
    __shared__ uint32_t sw[8];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        size_t i = base + (size_t)j * SCAN_THREADS + threadIdx.x;
        if (i < n) s += in[i];
    }
    uint32_t tot;
    block_exclusive_scan_256(s, sw, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// single block: in-place exclusive scan of the tile sums; writes the grand total (64-bit)
static __global__ void __launch_bounds__(1024) scan_tilesums_kernel(uint32_t* __restrict__ sums, size_t m,
                                                             unsigned long long* __restrict__ total_out) {
    __shared__ unsigned long long swarp[32];
    __shared__ unsigned long long carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t base = 0; base < m; base += 1024) {
        size_t i = base + threadIdx.x;
        unsigned long long v = (i < m) ? sums[i] : 0;
        unsigned long long inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) swarp[warp] = inc;
        __syncthreads();
        unsigned long long ws = swarp[lane];
        unsigned long long wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        unsigned long long wbase = __shfl_sync(0xffffffffu, wi - ws, warp);
        unsigned long long chunk_total = __shfl_sync(0xffffffffu, wi, 31);
        unsigned long long carry = carry_s;
        if (i < m) sums[i] = (uint32_t)(carry + wbase + inc - v);
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry_s;
}

static __global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const uint32_t* __restrict__ in,
                                                                   uint32_t* __restrict__ out,
                                                                   const uint32_t* __restrict__ tile_offsets, size_t n) {
    // coalesced load into padded smem, blocked (consecutive) processing per thread, coalesced store
    __shared__ uint32_t tile[SCAN_TILE + SCAN_TILE / 32];
    __shared__ uint32_t sw[8];
    size_t base = (size_t)blockIdx.x * SCAN_TILE;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int li = j * SCAN_THREADS + threadIdx.x;
        size_t i = base + li;
        tile[li + (li >> 5)] = (i < n) ? in[i] : 0;
    }
    __syncthreads();
    uint32_t v[SCAN_ITEMS];
    uint32_t s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int li = threadIdx.x * SCAN_ITEMS + j;
        v[j] = tile[li + (li >> 5)];
        s += v[j];
    }
    uint32_t ex = block_exclusive_scan_256(s, sw, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int li = threadIdx.x * SCAN_ITEMS + j;
        tile[li + (li >> 5)] = ex;
        ex += v[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        int li = j * SCAN_THREADS + threadIdx.x;
        size_t i = base + li;
        if (i < n) out[i] = tile[li + (li >> 5)];
    }
}

// out[i] = sum_{j<i} in[j]; in == out allowed.  If d_total != nullptr the 64-bit grand total is
// written there (device memory).  No host synchronisation.
inline void exclusive_scan_u32(Ctx& c, const uint32_t* in, uint32_t* out, size_t n, unsigned long long* d_total) {
    if (n == 0) {
        if (d_total) OL_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), c.stream));
        return;
    }
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    DevBuf<uint32_t> sums(c, tiles);
    ProfScope ps(c, "scan", (double)n);
    scan_reduce_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, c.stream>>>(in, sums.get(), n);
    OL_CHECK_LAUNCH();
    scan_tilesums_kernel<<<1, 1024, 0, c.stream>>>(sums.get(), tiles, d_total);
    OL_CHECK_LAUNCH();
    scan_apply_kernel<<<(unsigned)tiles, SCAN_THREADS, 0, c.stream>>>(in, out, sums.get(), n);
    OL_CHECK_LAUNCH();
}

// =============================================================================================
// stable LSD radix sort of (key, u32 value) pairs.
// One pass = digit histogram per 4096-key tile -> exclusive scan of the (digit-major) count
// matrix -> stable scatter with warp-level match ranking.
// =============================================================================================
constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 16;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;
constexpr int SORT_WARPS = SORT_THREADS / 32;

template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS) radix_hist_kernel(const KeyT* __restrict__ keys,
                                                                   uint32_t* __restrict__ counts, uint32_t n,
                                                                   uint32_t num_tiles, int shift, uint32_t nbins) {
    __shared__ uint32_t hist[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < 256; b += SORT_THREADS) hist[b] = 0;
    __syncthreads();
    const uint32_t mask = nbins - 1;
    const uint32_t wbase = blockIdx.x * SORT_TILE + warp * (32 * SORT_ITEMS);
#pragma unroll 4
    for (int j = 0; j < SORT_ITEMS; ++j) {
        uint32_t i = wbase + j * 32 + lane;
        bool valid = i < n;
        uint32_t d = valid ? (uint32_t)((keys[i] >> shift) & mask) : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == (__ffs(peers) - 1)) atomicAdd(&hist[d], __popc(peers));
    }
    __syncthreads();
    for (uint32_t b = threadIdx.x; b < nbins; b += SORT_THREADS) counts[(size_t)b * num_tiles + blockIdx.x] = hist[b];
}

template <typename KeyT>
__global__ void __launch_bounds__(SORT_THREADS) radix_scatter_kernel(const KeyT* __restrict__ keys_in,
                                                                      const uint32_t* __restrict__ vals_in,
                                                                      KeyT* __restrict__ keys_out,
                                                                      uint32_t* __restrict__ vals_out,
                                                                      const uint32_t* __restrict__ offsets, uint32_t n,
                                                                      uint32_t num_tiles, int shift, uint32_t nbins) {
    __shared__ uint32_t whist[SORT_WARPS][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int b = threadIdx.x; b < SORT_WARPS * 256; b += SORT_THREADS) (&whist[0][0])[b] = 0;
    __syncthreads();
    const uint32_t mask = nbins - 1;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t wbase = blockIdx.x * SORT_TILE + warp * (32 * SORT_ITEMS);
    KeyT k[SORT_ITEMS];
    uint32_t rnk[SORT_ITEMS];
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        uint32_t i = wbase + j * 32 + lane;
        bool valid = i < n;
        k[j] = valid ? keys_in[i] : (KeyT)0;
        uint32_t d = valid ? (uint32_t)((k[j] >> shift) & mask) : 0xffffffffu;
        uint32_t peers = __match_any_sync(0xffffffffu, d);
        int leader = __ffs(peers) - 1;
        uint32_t old = 0;
        if (valid && lane == leader) {
            old = whist[warp][d];
            whist[warp][d] = old + __popc(peers);
        }
        old = __shfl_sync(0xffffffffu, old, leader);
        rnk[j] = old + __popc(peers & lt);
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive prefix over the warps of this tile, on top of the tile's global offset
    for (uint32_t b = threadIdx.x; b < nbins; b += SORT_THREADS) {
        uint32_t run = offsets[(size_t)b * num_tiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_WARPS; ++w) {
            uint32_t t = whist[w][b];
            whist[w][b] = run;
            run += t;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SORT_ITEMS; ++j) {
        uint32_t i = wbase + j * 32 + lane;
        if (i < n) {
            uint32_t d = (uint32_t)((k[j] >> shift) & mask);
            uint32_t dst = whist[warp][d] + rnk[j];
            keys_out[dst] = k[j];
            vals_out[dst] = vals_in[i];
        }
    }
}

// Sorts (keys0, vals0) by bits [begin_bit, end_bit) using (keys1, vals1) as the alternate buffer.
// Returns 0 if the result is in buffer 0, 1 if it is in buffer 1.  Onesweep (onesweep.cuh) below 2^30
// pairs; the three-kernel LSD sort in this file above that (and when g_force_legacy_sort is set: tests).
extern bool g_force_legacy_sort;
template <typename KeyT>
inline int radix_sort_pairs(Ctx& c, KeyT* keys0, KeyT* keys1, uint32_t* vals0, uint32_t* vals1, size_t n, int begin_bit,
                            int end_bit, bool main_sort = false) {
    OL_REQUIRE(n < (1ull << 31), OL_ERR_INVALID, "radix_sort_pairs: n must be < 2^31");
    int bits = end_bit - begin_bit;
    if (n <= 1 || bits <= 0) return 0;
    if (n < (1ull << 30) && !g_force_legacy_sort) return onesweep_sort_pairs<KeyT>(c, keys0, keys1, vals0, vals1, n, begin_bit, end_bit, main_sort);
    int passes = (bits + 7) / 8;
    int per = (bits + passes - 1) / passes;
    uint32_t tiles = (uint32_t)((n + SORT_TILE - 1) / SORT_TILE);
    DevBuf<uint32_t> counts(c, (size_t)256 * tiles);
    int cur = 0;
    int bit = begin_bit;
    for (int p = 0; p < passes; ++p) {
        int nb = (end_bit - bit < per) ? (end_bit - bit) : per;
        uint32_t nbins = 1u << nb;
        KeyT* kin = cur ? keys1 : keys0;
        KeyT* kout = cur ? keys0 : keys1;
        uint32_t* vin = cur ? vals1 : vals0;
        uint32_t* vout = cur ? vals0 : vals1;
        {
            ProfScope ps(c, sizeof(KeyT) == 8 ? "radix_hist_u64" : "radix_hist_u32", (double)n);
            radix_hist_kernel<KeyT><<<tiles, SORT_THREADS, 0, c.stream>>>(kin, counts.get(), (uint32_t)n, tiles, bit, nbins);
            OL_CHECK_LAUNCH();
        }
        exclusive_scan_u32(c, counts.get(), counts.get(), (size_t)nbins * tiles, nullptr);
        {
            ProfScope ps(c, sizeof(KeyT) == 8 ? "radix_scatter_u64" : "radix_scatter_u32", (double)n);
            radix_scatter_kernel<KeyT><<<tiles, SORT_THREADS, 0, c.stream>>>(kin, vin, kout, vout, counts.get(), (uint32_t)n,
                                                                               tiles, bit, nbins);
            OL_CHECK_LAUNCH();
        }
        cur ^= 1;
        bit += nb;
    }
    return cur;
}

// =============================================================================================
// run segmentation: maximal runs of equal key(i) over i in [0, n)  (K3 cells, (cell, pose) pairs, (pose, leaf) blocks)
//   runs_count_kernel  heads per 2048-element tile
//   (exclusive scan of the tile counts, total = number of runs)
//   runs_emit_kernel   recomputes the head flags, ranks them with ballots, writes run_of_pos[i] and calls emit(run, i, key)
//                      for every head
// Traffic: the key inputs are read twice and run_of_pos is written once; no flag / scan arrays of length n.
// KeyFn:  __device__ uint64_t operator()(uint32_t i) const;      EmitFn: __device__ void operator()(uint32_t run, uint32_t i, uint64_t key) const
// =============================================================================================
constexpr int RUNS_THREADS = 256;
constexpr int RUNS_ITEMS = 8;
constexpr int RUNS_TILE = RUNS_THREADS * RUNS_ITEMS;

template <typename KeyFn>
__device__ __forceinline__ void runs_warp_flags(const KeyFn& key, uint32_t wbase, uint32_t n, int lane, uint32_t masks[RUNS_ITEMS],
                                                uint64_t keys[RUNS_ITEMS]) {
    // the warp walks RUNS_ITEMS x 32 consecutive elements; element (j, lane) = wbase + 32 j + lane
    uint64_t carry = 0;  // key of the element just before the current group of 32 (valid in every lane)
    bool have_carry = false;
    if (wbase > 0 && wbase < n) {
        carry = key(wbase - 1);
        have_carry = true;
    }
#pragma unroll
    for (int j = 0; j < RUNS_ITEMS; ++j) {
        const uint32_t i = wbase + 32 * j + lane;
        const bool valid = i < n;
        const uint64_t k = valid ? key(i) : 0ull;
        uint64_t prev = __shfl_up_sync(0xffffffffu, k, 1);
        bool has_prev = true;
        if (lane == 0) {
            prev = carry;
            has_prev = have_carry;
        }
        const bool head = valid && (!has_prev || prev != k);
        masks[j] = __ballot_sync(0xffffffffu, head);
        keys[j] = k;
        carry = __shfl_sync(0xffffffffu, k, 31);
        have_carry = true;
    }
}

template <typename KeyFn>
__global__ void __launch_bounds__(RUNS_THREADS) runs_count_kernel(KeyFn key, uint32_t n, uint32_t* __restrict__ tile_counts) {
    __shared__ uint32_t s_w[RUNS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = blockIdx.x * RUNS_TILE + warp * (32 * RUNS_ITEMS);
    uint32_t masks[RUNS_ITEMS];
    uint64_t keys[RUNS_ITEMS];
    runs_warp_flags(key, wbase, n, lane, masks, keys);
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < RUNS_ITEMS; ++j) cnt += __popc(masks[j]);
    if (lane == 0) s_w[warp] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < RUNS_THREADS / 32; ++w) t += s_w[w];
        tile_counts[blockIdx.x] = t;
    }
}

template <typename KeyFn, typename EmitFn>
__global__ void __launch_bounds__(RUNS_THREADS) runs_emit_kernel(KeyFn key, EmitFn emit, uint32_t n,
                                                                 const uint32_t* __restrict__ tile_offsets,
                                                                 uint32_t* __restrict__ run_of_pos) {
    __shared__ uint32_t s_w[RUNS_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t wbase = blockIdx.x * RUNS_TILE + warp * (32 * RUNS_ITEMS);
    uint32_t masks[RUNS_ITEMS];
    uint64_t keys[RUNS_ITEMS];
    runs_warp_flags(key, wbase, n, lane, masks, keys);
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < RUNS_ITEMS; ++j) cnt += __popc(masks[j]);
    if (lane == 0) s_w[warp] = cnt;
    __syncthreads();
    uint32_t run = tile_offsets[blockIdx.x];  // heads before this warp's first element
    for (int w = 0; w < warp; ++w) run += s_w[w];
    const uint32_t lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < RUNS_ITEMS; ++j) {
        const uint32_t i = wbase + 32 * j + lane;
        const bool head = (masks[j] >> lane) & 1u;
        const uint32_t mine = run + __popc(masks[j] & lt) + (head ? 1u : 0u) - 1u;  // index of the run element i belongs to
        if (i < n) {
            if (run_of_pos) run_of_pos[i] = mine;
            if (head) emit(mine, i, keys[j]);
        }
        run += __popc(masks[j]);
    }
}

// number of runs is written to d_total (device, 64-bit); run tables must be sized by the caller after reading it, so
// the emit pass is a separate call
template <typename KeyFn>
inline void segment_runs_count(Ctx& c, KeyFn key, size_t n, DevBuf<uint32_t>& tile_offsets, unsigned long long* d_total) {
    const size_t tiles = (n + RUNS_TILE - 1) / RUNS_TILE;
    tile_offsets.reset(c, tiles ? tiles : 1);
    if (n == 0) {
        OL_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), c.stream));
        return;
    }
    runs_count_kernel<KeyFn><<<(unsigned)tiles, RUNS_THREADS, 0, c.stream>>>(key, (uint32_t)n, tile_offsets.get());
    OL_CHECK_LAUNCH();
    exclusive_scan_u32(c, tile_offsets.get(), tile_offsets.get(), tiles, d_total);
}

template <typename KeyFn, typename EmitFn>
inline void segment_runs_emit(Ctx& c, KeyFn key, EmitFn emit, size_t n, const DevBuf<uint32_t>& tile_offsets, uint32_t* run_of_pos) {
    if (n == 0) return;
    const size_t tiles = (n + RUNS_TILE - 1) / RUNS_TILE;
    runs_emit_kernel<KeyFn, EmitFn><<<(unsigned)tiles, RUNS_THREADS, 0, c.stream>>>(key, emit, (uint32_t)n, tile_offsets.get(), run_of_pos);
    OL_CHECK_LAUNCH();
}

// =============================================================================================
// small elementwise helpers
// =============================================================================================
static __global__ void iota_kernel(uint32_t* out, uint32_t n, uint32_t first) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = first + i;
}

template <typename T>
__global__ void fill_kernel(T* out, size_t n, T v) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

template <typename T>
__global__ void gather_kernel(T* __restrict__ out, const T* __restrict__ in, const uint32_t* __restrict__ idx, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[idx[i]];
}

// highest set bit + 1 (0 for v == 0)
inline int bit_length_u64(uint64_t v) {
    int b = 0;
    while (v) {
        ++b;
        v >>= 1;
    }
    return b;
}

}  // namespace ol
