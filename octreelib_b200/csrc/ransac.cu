// K6: per-leaf batched RANSAC plane segmentation, one CTA per (pose, leaf) block.
//
// Replaces the reference's numba-CUDA kernel
//   /root/reference/octreelib/ransac/cuda_ransac.py:85-155  (kernel body)
//   /root/reference/octreelib/ransac/util.py:16-24, 28-84   (distance, plane fit)
// with the same arithmetic (float64 plane fit without FMA contraction, plane rounded to float32,
// float64 point-plane distance compared against a float64 threshold), so that every hypothesis'
// inlier count - and therefore the arg-max and the mask - is identical to the reference's.
// Differences by design:
//   * the leaf's points are staged in shared memory once (TMA bulk copy + mbarrier) instead of
//     being streamed from global memory by all 1024 threads;
//   * ties between equally good hypotheses are broken deterministically (lowest index) where the
//     reference has a CAS race (cuda_ransac.py:135-146);
//   * the chosen plane, hypothesis index and inlier count are returned per block.
#include "common.cuh"
#include "forest.cuh"

namespace ol {

constexpr int RANSAC_THREADS = 256;
constexpr uint32_t RANSAC_SMEM_POINTS_MAX = 8192;  // 192 KB of float64 xyz
constexpr uint32_t RANSAC_FLAG_NO_TMA = 1u;

struct RansacArgs {
    const double* points;
    long long n_points;
    const uint32_t* blk_start;      // physical first point of block b
    const int32_t* blk_size;        // points in block b
    const long long* blk_ref_start; // block_start_indices[b] of the reference's batch layout
    const uint32_t* work;           // blocks to score (size >= K)
    uint32_t n_work;
    const double* table;            // [H][K]
    int H, K;
    double thr;
    uint8_t* mask;
    float* plane;
    int32_t* best;
    int32_t* best_count;
    uint32_t cap;                   // points that fit the shared-memory staging area
    uint32_t flags;
    uint32_t* err;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// sample index of hypothesis t, draw i (cuda_ransac.py:103-107): float64 arithmetic on the
// reference's batch-global start index, truncated to int32; returned relative to the block.
__device__ __forceinline__ int sample_index(const double* __restrict__ table, int K, int t, int i, int n, long long ref_start,
                                            uint32_t* err) {
    double v = __dadd_rn(__dmul_rn(__ldg(&table[(size_t)t * K + i]), (double)n), (double)ref_start);
    long long j = (long long)(int)v - ref_start;
    if (j < 0 || j >= n) {  // float rounding pushed the draw into a neighbouring block (p ~ 1e-8)
        atomicOr(err, (uint32_t)DEVERR_SAMPLE_OOB);
        j = j < 0 ? 0 : n - 1;
    }
    return (int)j;
}

// util.py:28-84, op for op (no contraction), result rounded to float32 (cuda_ransac.py:110-113)
__device__ __forceinline__ float4 fit_plane(const double* __restrict__ pts, const double* __restrict__ table, int K, int t,
                                            int n, long long ref_start, uint32_t* err) {
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int i = 0; i < K; ++i) {
        const double* p = pts + 3 * sample_index(table, K, t, i, n, ref_start, err);
        cx = __dadd_rn(cx, p[0]);
        cy = __dadd_rn(cy, p[1]);
        cz = __dadd_rn(cz, p[2]);
    }
    const double kd = (double)K;
    cx = __ddiv_rn(cx, kd);
    cy = __ddiv_rn(cy, kd);
    cz = __ddiv_rn(cz, kd);
    double xx = 0.0, xy = 0.0, xz = 0.0, yy = 0.0, yz = 0.0, zz = 0.0;
    for (int i = 0; i < K; ++i) {
        const double* p = pts + 3 * sample_index(table, K, t, i, n, ref_start, err);
        const double rx = __dsub_rn(p[0], cx), ry = __dsub_rn(p[1], cy), rz = __dsub_rn(p[2], cz);
        xx = __dadd_rn(xx, __dmul_rn(rx, rx));
        xy = __dadd_rn(xy, __dmul_rn(rx, ry));
        xz = __dadd_rn(xz, __dmul_rn(rx, rz));
        yy = __dadd_rn(yy, __dmul_rn(ry, ry));
        yz = __dadd_rn(yz, __dmul_rn(ry, rz));
        zz = __dadd_rn(zz, __dmul_rn(rz, rz));
    }
    const double det_x = __dsub_rn(__dmul_rn(yy, zz), __dmul_rn(yz, yz));
    const double det_y = __dsub_rn(__dmul_rn(xx, zz), __dmul_rn(xz, xz));
    const double det_z = __dsub_rn(__dmul_rn(xx, yy), __dmul_rn(xy, xy));
    double ax, ay, az;
    if (det_x > det_y && det_x > det_z) {
        ax = det_x;
        ay = __dsub_rn(__dmul_rn(xz, yz), __dmul_rn(xy, zz));
        az = __dsub_rn(__dmul_rn(xy, yz), __dmul_rn(xz, yy));
    } else if (det_y > det_z) {
        ax = __dsub_rn(__dmul_rn(xz, yz), __dmul_rn(xy, zz));
        ay = det_y;
        az = __dsub_rn(__dmul_rn(xy, xz), __dmul_rn(yz, xx));
    } else {
        ax = __dsub_rn(__dmul_rn(xy, yz), __dmul_rn(xz, yy));
        ay = __dsub_rn(__dmul_rn(xy, xz), __dmul_rn(yz, xx));
        az = det_z;
    }
    const double norm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
    if (norm == 0.0) return make_float4(0.f, 0.f, 0.f, 0.f);
    ax = __ddiv_rn(ax, norm);
    ay = __ddiv_rn(ay, norm);
    az = __ddiv_rn(az, norm);
    const double d = -__dadd_rn(__dadd_rn(__dmul_rn(ax, cx), __dmul_rn(ay, cy)), __dmul_rn(az, cz));
    return make_float4((float)ax, (float)ay, (float)az, (float)d);
}

// util.py:16-24: float32 plane promoted to float64, left-to-right sum, no contraction
__device__ __forceinline__ double plane_distance(double a, double b, double c, double d, const double* __restrict__ p) {
    return fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, p[0]), __dmul_rn(b, p[1])), __dmul_rn(c, p[2])), d));
}

__global__ void __launch_bounds__(RANSAC_THREADS) ransac_kernel(RansacArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: [0,8) mbarrier | [8,16) best key | [16, 16 + H*16) planes | staged points (16 B aligned)
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem_raw + 8);
    float4* s_plane = reinterpret_cast<float4*>(smem_raw + 16);
    double* s_pts_raw = reinterpret_cast<double*>(smem_raw + 16 + (size_t)A.H * 16);

    const uint32_t b = A.work[blockIdx.x];
    const int n = A.blk_size[b];
    const uint32_t ps = A.blk_start[b];
    const long long rs = A.blk_ref_start[b];
    const double* gsrc = A.points + (size_t)ps * 3;
    const double* pts = gsrc;
    const int tid = threadIdx.x;

    if (tid == 0) *s_best = 0ull;
    if ((uint32_t)n <= A.cap) {
        if (!(A.flags & RANSAC_FLAG_NO_TMA)) {
            // TMA bulk copy of the block's contiguous float64 xyz run.  Source must be 16-byte
            // aligned: start from the aligned-down address (the host guarantees the buffer base
            // is 16-byte aligned) and fetch the odd trailing 8 bytes, if any, with a plain load.
            const uintptr_t src_addr = reinterpret_cast<uintptr_t>(gsrc);
            const uint32_t lead = (uint32_t)(src_addr & 15u);  // 0 or 8
            const unsigned char* asrc = reinterpret_cast<const unsigned char*>(src_addr - lead);
            const uint32_t total = lead + (uint32_t)n * 24u;
            const uint32_t bulk = total & ~15u;
            const uint32_t bar = smem_addr(s_bar);
            if (tid == 0) {
                mbar_init(bar, 1);
                asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            }
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(bar, bulk);
                const uint32_t dst0 = smem_addr(s_pts_raw);
                for (uint32_t off = 0; off < bulk; off += 32768u) {
                    const uint32_t sz = (bulk - off) < 32768u ? (bulk - off) : 32768u;
                    tma_bulk_g2s(dst0 + off, asrc + off, sz, bar);
                }
                if (total != bulk) s_pts_raw[bulk / 8] = reinterpret_cast<const double*>(asrc)[bulk / 8];
            }
            mbar_wait(bar, 0);
            __syncthreads();
            pts = s_pts_raw + lead / 8;
        } else {
            for (int i = tid; i < n * 3; i += RANSAC_THREADS) s_pts_raw[i] = gsrc[i];
            __syncthreads();
            pts = s_pts_raw;
        }
    } else {
        __syncthreads();
    }

    // ---- hypotheses: thread t handles t, t + 256, ... ----------------------------------------
    unsigned long long my_best = 0ull;
    for (int t = tid; t < A.H; t += RANSAC_THREADS) {
        const float4 pl = fit_plane(pts, A.table, A.K, t, n, rs, A.err);
        s_plane[t] = pl;
        const double a = (double)pl.x, bb = (double)pl.y, c = (double)pl.z, d = (double)pl.w;
        int cnt = 0;
        for (int i = 0; i < n; ++i) cnt += plane_distance(a, bb, c, d, pts + 3 * i) < A.thr;  // cuda_ransac.py:116-121
        // max count, ties -> lowest hypothesis index
        const unsigned long long key = ((unsigned long long)(uint32_t)cnt << 32) | (unsigned long long)(0xffffffffu - (uint32_t)t);
        my_best = key > my_best ? key : my_best;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, my_best, o);
        my_best = other > my_best ? other : my_best;
    }
    if ((tid & 31) == 0) atomicMax(s_best, my_best);
    __syncthreads();
    const unsigned long long bk = *s_best;
    const int best_t = (int)(0xffffffffu - (uint32_t)(bk & 0xffffffffull));
    const int best_cnt = (int)(bk >> 32);
    const float4 bp = s_plane[best_t];
    if (tid == 0) {
        if (A.plane) reinterpret_cast<float4*>(A.plane)[b] = bp;
        if (A.best) A.best[b] = best_t;
        if (A.best_count) A.best_count[b] = best_cnt;
    }
    // ---- final mask (cuda_ransac.py:149-155) ---------------------------------------------------
    const double a = (double)bp.x, bb = (double)bp.y, c = (double)bp.z, d = (double)bp.w;
    for (int i = tid; i < n; i += RANSAC_THREADS) A.mask[(size_t)ps + i] = plane_distance(a, bb, c, d, pts + 3 * i) < A.thr ? 1 : 0;
}

void launch_ransac(Ctx& c, const double* points, int64_t n_points, const uint32_t* blk_phys_start, const int32_t* blk_size,
                   const long long* blk_ref_start, const uint32_t* work_list, uint32_t n_work, uint32_t max_block,
                   const double* table, int H, int K, double threshold, uint8_t* mask, float* plane, int32_t* best,
                   int32_t* best_count, uint32_t flags) {
    if (n_work == 0) return;
    RansacArgs a{};
    a.points = points;
    a.n_points = n_points;
    a.blk_start = blk_phys_start;
    a.blk_size = blk_size;
    a.blk_ref_start = blk_ref_start;
    a.work = work_list;
    a.n_work = n_work;
    a.table = table;
    a.H = H;
    a.K = K;
    a.thr = threshold;
    a.mask = mask;
    a.plane = plane;
    a.best = best;
    a.best_count = best_count;
    a.cap = max_block < RANSAC_SMEM_POINTS_MAX ? max_block : RANSAC_SMEM_POINTS_MAX;
    a.flags = flags;
    a.err = c.d_err;
    if (reinterpret_cast<uintptr_t>(points) & 15u) a.flags |= RANSAC_FLAG_NO_TMA;
    size_t smem = 16 + (size_t)H * 16 + ((size_t)a.cap * 3 + 2) * 8;
    smem = (smem + 15) & ~(size_t)15;
    OL_CUDA(cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ransac_kernel<<<n_work, RANSAC_THREADS, smem, c.stream>>>(a);
    OL_CHECK_LAUNCH();
}

}  // namespace ol
