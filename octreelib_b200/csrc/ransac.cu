// K6: per-leaf batched RANSAC plane segmentation, one CTA per (pose, leaf) block.
//
// Replaces the reference's numba-CUDA kernel
//   /root/reference/octreelib/ransac/cuda_ransac.py:85-155  (kernel body)
//   /root/reference/octreelib/ransac/util.py:16-24, 28-84   (distance, plane fit)
// and returns exactly what that arithmetic returns under IEEE evaluation WITHOUT contraction (float64 plane fit, plane
// rounded to float32, float64 point-plane distance compared against a float64 threshold): every reported inlier count,
// the arg-max and the mask are bit-identical to the reference run under NUMBA_ENABLE_CUDASIM (its own CI setting, pure
// Python) and to the C oracle.  On real hardware numba compiles the reference kernel through libNVVM, which may contract
// a*b+c into FMAs; that build could not be compared here (its launch fails on the B200:
// profiles/r02_reference_numba_on_b200_v2.json), so the claim is against the un-contracted arithmetic only.
//
// How the work is organised:
//   1. the block's float64 points are staged in shared memory once (TMA bulk copy + mbarrier)
//      instead of being streamed from global memory by all 1024 threads;
//   2. FP32 INTERVAL PRE-FILTER: every hypothesis is first fitted and scored in float32 on the
//      block-local (origin-shifted) coordinates, together with a rigorous bound eps_t on
//      |distance_reference - distance_fp32|.  That gives an interval lo_t <= count_t <= hi_t for the
//      reference's exact inlier count.  Hypotheses whose fit is ill conditioned, whose adjugate-row
//      selection is ambiguous, or whose sample index sits next to a rounding boundary get the
//      trivial interval [0, n];
//   3. only the hypotheses that can still be the winner (hi_t >= max lo, with the lowest-index
//      tie-break taken into account) are evaluated with the reference's exact float64 arithmetic;
//      typically 1-30 of 1024;
//   4. hypotheses are processed in chunks of 256 in index order; as soon as one is certain to have
//      ALL points as inliers, no later hypothesis can win (ties go to the lowest index) and the
//      remaining chunks are skipped.
// Flags select the plain exact evaluation of all hypotheses (RANSAC_FLAG_EXACT_ONLY) or a
// verification mode (RANSAC_FLAG_VERIFY) that evaluates everything exactly AND checks
// lo_t <= count_t <= hi_t for every hypothesis (tests run both and require identical output).
// Differences to the reference by design:
//   * ties between equally good hypotheses are broken deterministically (lowest index) where the
//     reference has a CAS race (cuda_ransac.py:135-146);
//   * the chosen plane, hypothesis index and inlier count are returned per block.
#include <algorithm>

#include "common.cuh"
#include "forest.cuh"

namespace ol {

constexpr int RANSAC_THREADS = 256;
constexpr int RANSAC_MAX_H = 1024;
constexpr int RANSAC_MAX_CHUNKS = RANSAC_MAX_H / RANSAC_THREADS;
constexpr uint32_t RANSAC_SMEM_POINTS_MAX = 4096;  // 96 KB of float64 xyz + 64 KB of float4 local copies
constexpr uint32_t RANSAC_FLAG_NO_TMA = 1u;
constexpr uint32_t RANSAC_FLAG_EXACT_ONLY = 2u;
constexpr uint32_t RANSAC_FLAG_VERIFY = 4u;
constexpr uint32_t RANSAC_FLAG_STATS = 8u;
constexpr int RANSAC_COOP_MIN_POINTS = 96;  // blocks at least this large score candidates warp-cooperatively

// statistics of the pre-filter (only with RANSAC_FLAG_STATS / VERIFY): see ol_ransac_stats_read
__device__ unsigned long long g_ransac_stats[16];

struct RansacArgs {
    const double* points;
    long long n_points;
    const uint32_t* blk_start;      // physical first point of block b
    const int32_t* blk_size;        // points in block b
    const long long* blk_ref_start; // block_start_indices[b] of the reference's batch layout
    const uint32_t* work;           // blocks to score (size >= K), indexed by work item w
    uint32_t n_work;
    const uint32_t* sub;            // optional: the work items this launch handles; NULL = all of them
    uint32_t n_sub;
    const uint32_t* n_sub_dev;      // optional: device-side length of `sub` (warp kernel after the lane passes)
    const uint32_t* pk_start;       // optional [n_work]: first point of work item w inside `points` when `points` only
                                    // holds the fitted blocks back to back; NULL = `points` is indexed by blk_start
    const double* table;            // [H][K] float64 uniform [0,1)
    const uint32_t* r32t;           // [K][H] floor(table * 2^32), transposed (pre-filter)
    const uint32_t* table_ok;       // [1] 1 if every table entry is in [0, 1)
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

// =============================================================================================
// exact path: the reference's arithmetic, op for op
// =============================================================================================
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
// KC > 0: the number of draws is known at compile time (the loops unroll and the K sample indices stay in registers
// instead of being recomputed for the second pass); KC = 0: any K.  SMEM: `pts` points into shared memory (LDS instead
// of generic loads).  The arithmetic and its order are the same in every instantiation.
template <int KC, bool SMEM>
__device__ __forceinline__ float4 fit_plane_impl(const double* __restrict__ pts, const double* __restrict__ table, int K, int t, int n,
                                                 long long ref_start, uint32_t* err) {
    if (SMEM) __builtin_assume(__isShared(pts));
    const int kk = KC ? KC : K;
    int idx[KC ? KC : 1];
    double cx = 0.0, cy = 0.0, cz = 0.0;
#pragma unroll
    for (int i = 0; i < kk; ++i) {
        const int s = sample_index(table, kk, t, i, n, ref_start, err);
        if (KC) idx[i] = s;
        const double* p = pts + 3 * s;
        cx = __dadd_rn(cx, p[0]);
        cy = __dadd_rn(cy, p[1]);
        cz = __dadd_rn(cz, p[2]);
    }
    const double kd = (double)kk;
    cx = __ddiv_rn(cx, kd);
    cy = __ddiv_rn(cy, kd);
    cz = __ddiv_rn(cz, kd);
    double xx = 0.0, xy = 0.0, xz = 0.0, yy = 0.0, yz = 0.0, zz = 0.0;
#pragma unroll
    for (int i = 0; i < kk; ++i) {
        const double* p = pts + 3 * (KC ? idx[i] : sample_index(table, kk, t, i, n, ref_start, err));
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

// generic address space (the CTA kernel streams blocks > cap points from global memory)
__device__ __noinline__ float4 fit_plane(const double* __restrict__ pts, const double* __restrict__ table, int K, int t, int n,
                                         long long ref_start, uint32_t* err) {
    if (K == 6) return fit_plane_impl<6, false>(pts, table, K, t, n, ref_start, err);
    return fit_plane_impl<0, false>(pts, table, K, t, n, ref_start, err);
}
// points staged in shared memory (warp-per-block kernel)
__device__ __noinline__ float4 fit_plane_smem(const double* __restrict__ pts, const double* __restrict__ table, int K, int t, int n,
                                              long long ref_start, uint32_t* err) {
    if (K == 6) return fit_plane_impl<6, true>(pts, table, K, t, n, ref_start, err);
    return fit_plane_impl<0, true>(pts, table, K, t, n, ref_start, err);
}

// util.py:16-24: float32 plane promoted to float64, left-to-right sum, no contraction
__device__ __forceinline__ double plane_distance(double a, double b, double c, double d, const double* __restrict__ p) {
    return fabs(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a, p[0]), __dmul_rn(b, p[1])), __dmul_rn(c, p[2])), d));
}

__device__ __forceinline__ unsigned long long pack_key(int count, int t) {  // max count, ties -> lowest hypothesis index
    return ((unsigned long long)(uint32_t)count << 32) | (unsigned long long)(0xffffffffu - (uint32_t)t);
}

// Exact evaluation of one hypothesis per lane (active lanes only); returns the lane's inlier count
// (cuda_ransac.py:116-121).  Large blocks are scored warp-cooperatively: every active lane's plane
// is broadcast in turn and all 32 lanes stride over the points.
__device__ __forceinline__ int exact_count(const double* __restrict__ pts, int n, bool active, const float4 pl, double thr) {
    int cnt = 0;
    if (n < RANSAC_COOP_MIN_POINTS) {
        if (active) {
            const double a = (double)pl.x, b = (double)pl.y, c = (double)pl.z, d = (double)pl.w;
            for (int i = 0; i < n; ++i) cnt += plane_distance(a, b, c, d, pts + 3 * i) < thr;
        }
        return cnt;
    }
    const int lane = threadIdx.x & 31;
    uint32_t todo = __ballot_sync(0xffffffffu, active);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const double a = (double)__shfl_sync(0xffffffffu, pl.x, src), b = (double)__shfl_sync(0xffffffffu, pl.y, src);
        const double c = (double)__shfl_sync(0xffffffffu, pl.z, src), d = (double)__shfl_sync(0xffffffffu, pl.w, src);
        int part = 0;
        for (int i = lane; i < n; i += 32) part += plane_distance(a, b, c, d, pts + 3 * i) < thr;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (lane == src) cnt = part;
    }
    return cnt;
}

// =============================================================================================
// FP32 interval pre-filter
// =============================================================================================
struct Interval {
    int lo, hi;
};

// Bounds on the reference's exact inlier count of hypothesis t from a float32 fit on the block-local
// coordinates q_j = fl32(p_j - p_0).  With u = 2^-24, Qm >= max_j |q_j|_2, Pm >= max_j |p_j|_1:
//   reference side : the plane is rounded to float32 before scoring (cuda_ransac.py:110-113), which moves
//                    a.x + b.y + c.z + d by at most 2^-25 |p|_1 + 2^-24 |d| <= 1.5 * 2^-24 Pm
//   filter side    : covariance entries carry E_m (rounding of q, of the residuals and of the sums),
//                    adjugate entries E_cof = 4 S E_m + 3 u S^2 (S = trace), so the unit normal is off
//                    by <= ~2 eta with eta = sqrt(3) E_cof / |n|; a point is at most 2 Qm from the centroid.
// Any hypothesis for which a bound cannot be given (eta too large, ambiguous adjugate row, sample index
// next to a rounding boundary, non-finite intermediate) returns [0, n].
__device__ __forceinline__ Interval filter_hypothesis(const float4* __restrict__ q, const uint32_t* __restrict__ r32t, int H,
                                                      int K, int t, int n, float Pm, float Qm, float thr) {
    const float u = 5.9604645e-8f;  // 2^-24
    const uint32_t guard = 0u - ((uint32_t)n + 4096u);  // low word at / above this: floor(R n) may be off by one
    bool ok = true;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    for (int i = 0; i < K; ++i) {
        const uint32_t r = __ldg(&r32t[(size_t)i * H + t]);
        ok = ok && (r * (uint32_t)n < guard);
        const float4 p = q[__umulhi(r, (uint32_t)n)];
        cx += p.x;
        cy += p.y;
        cz += p.z;
    }
    const float inv_k = 1.0f / (float)K;
    cx *= inv_k;
    cy *= inv_k;
    cz *= inv_k;
    float xx = 0.f, xy = 0.f, xz = 0.f, yy = 0.f, yz = 0.f, zz = 0.f;
    for (int i = 0; i < K; ++i) {
        const float4 p = q[__umulhi(__ldg(&r32t[(size_t)i * H + t]), (uint32_t)n)];
        const float rx = p.x - cx, ry = p.y - cy, rz = p.z - cz;
        xx = fmaf(rx, rx, xx);
        xy = fmaf(rx, ry, xy);
        xz = fmaf(rx, rz, xz);
        yy = fmaf(ry, ry, yy);
        yz = fmaf(ry, rz, yz);
        zz = fmaf(rz, rz, zz);
    }
    const float kf = (float)K;
    const float S = (xx + yy + zz) * 1.0001f;
    const float g = (kf + 2.f) * u * Qm;            // centroid error (common to all residuals: second order)
    const float gr = 1.8e-15f * (kf + 1.f) * Pm;    // the reference's own centroid error, 2^-53 (K+1) Pm (x2)
    const float Em = (2.f * kf + 6.f) * u * S + 4.f * u * Qm * sqrtf(kf * S) + 2.f * kf * (g * g + gr * gr);
    const float Ecof = 4.f * S * Em + 3.f * u * S * S;
    const float det_x = fmaf(yy, zz, -(yz * yz));
    const float det_y = fmaf(xx, zz, -(xz * xz));
    const float det_z = fmaf(xx, yy, -(xy * xy));
    // which adjugate row does the reference take (util.py:63-74)?  Only accept a certain answer.
    const float e2 = 2.f * Ecof;
    const bool b1 = (det_x - det_y > e2) && (det_x - det_z > e2);
    const bool nb1 = (det_y - det_x >= e2) || (det_z - det_x >= e2);
    const bool b2 = nb1 && (det_y - det_z > e2);
    const bool b3 = nb1 && (det_z - det_y >= e2);
    ok = ok && (b1 || b2 || b3);
    const float c_xy = fmaf(xz, yz, -(xy * zz));  // shared off-diagonal cofactors
    const float c_xz = fmaf(xy, yz, -(xz * yy));
    const float c_yz = fmaf(xy, xz, -(yz * xx));
    float ax, ay, az;
    if (b1) {
        ax = det_x;
        ay = c_xy;
        az = c_xz;
    } else if (b2) {
        ax = c_xy;
        ay = det_y;
        az = c_yz;
    } else {
        ax = c_xz;
        ay = c_yz;
        az = det_z;
    }
    const float nn = fmaf(ax, ax, fmaf(ay, ay, az * az));
    const float inv = rsqrtf(nn);
    const float eta = 1.7321f * Ecof * inv;
    ok = ok && (eta <= 0.05f) && (S >= 1e-30f);
    const float nx = ax * inv, ny = ay * inv, nz = az * inv;
    const float dn = -fmaf(nx, cx, fmaf(ny, cy, nz * cz));
    float eps = 1.1920929e-7f * Pm                       // 2^-23 Pm: float32 rounding of the reference plane
                + 2.f * Qm * (3.2f * eta + 16.f * u)     // direction error x lever arm, normalisation, dot product
                + 1.7321f * g + 4.f * u * thr;           // centroid error, threshold rounding
    eps *= 1.01f;
    ok = ok && (eps <= 0.25f * thr);
    const float t_lo = thr - eps, t_hi = thr + eps;
    int lo = 0, hi = 0;
    for (int j = 0; j < n; ++j) {
        const float4 p = q[j];
        const float d = fabsf(fmaf(nx, p.x, fmaf(ny, p.y, fmaf(nz, p.z, dn))));
        lo += d < t_lo;
        hi += d < t_hi;
    }
    Interval out;
    out.lo = ok ? lo : 0;
    out.hi = ok ? hi : n;
    return out;
}

// [K][H] transposed 32-bit fixed-point copy of the hypothesis table for the pre-filter
__global__ void ransac_table_prep_kernel(const double* __restrict__ table, int H, int K, uint32_t* __restrict__ r32t,
                                         uint32_t* __restrict__ table_ok) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= H * K) return;
    const int t = e / K, i = e - t * K;
    const double r = table[e];
    uint32_t v = 0;
    if (r >= 0.0 && r < 1.0)
        v = (uint32_t)(unsigned long long)(r * 4294967296.0);  // exact: power-of-two scaling, then truncation
    else
        atomicAnd(table_ok, 0u);
    r32t[(size_t)i * H + t] = v;
}

// shared-memory header (bytes): [0,8) mbarrier | [8,16) best exact key | [16,24) best lower-bound key |
// [24,28) Pmax bits | [28,32) Qmax^2 bits | [32,36) candidate count | [36,48) statistics | [48,64) winning plane
constexpr int RANSAC_SMEM_HEADER = 64;

__global__ void __launch_bounds__(RANSAC_THREADS) ransac_kernel(RansacArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(smem_raw);
    unsigned long long* s_best = reinterpret_cast<unsigned long long*>(smem_raw + 8);
    unsigned long long* s_sel = reinterpret_cast<unsigned long long*>(smem_raw + 16);
    uint32_t* s_pmax = reinterpret_cast<uint32_t*>(smem_raw + 24);
    uint32_t* s_qmax2 = reinterpret_cast<uint32_t*>(smem_raw + 28);
    uint32_t* s_nc = reinterpret_cast<uint32_t*>(smem_raw + 32);
    uint32_t* s_stat = reinterpret_cast<uint32_t*>(smem_raw + 36);  // [0] uncertain [1] violations
    float4* s_bp = reinterpret_cast<float4*>(smem_raw + 48);
    uint16_t* s_cand = reinterpret_cast<uint16_t*>(smem_raw + RANSAC_SMEM_HEADER);  // [1024]
    double* s_pts_raw = reinterpret_cast<double*>(smem_raw + RANSAC_SMEM_HEADER + 2 * RANSAC_MAX_H);
    float4* s_q = reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s_pts_raw) + ((((size_t)A.cap * 3 + 2) * 8 + 15) & ~(size_t)15));

    const uint32_t w = A.sub ? A.sub[blockIdx.x] : blockIdx.x;
    const uint32_t b = A.work[w];
    const int n = A.blk_size[b];
    const uint32_t ps = A.blk_start[b];
    const long long rs = A.blk_ref_start[b];
    const double* gsrc = A.points + (size_t)(A.pk_start ? A.pk_start[w] : ps) * 3;
    const double* pts = gsrc;
    const int tid = threadIdx.x;
    const bool staged = (uint32_t)n <= A.cap;
    const bool verify = (A.flags & RANSAC_FLAG_VERIFY) != 0;
    const bool use_filter = staged && (verify || !(A.flags & RANSAC_FLAG_EXACT_ONLY)) && (__ldg(A.table_ok) != 0u);
    const int nch = (A.H + RANSAC_THREADS - 1) / RANSAC_THREADS;

    if (tid == 0) {
        *s_best = 0ull;
        *s_sel = 0ull;
        *s_pmax = 0u;
        *s_qmax2 = 0u;
        *s_nc = 0u;
        s_stat[0] = 0u;
        s_stat[1] = 0u;
    }
    if (staged) {
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

    // ---- block-local float32 copies + magnitude bounds for the pre-filter ----------------------
    float Pm = 0.f, Qm = 0.f;
    if (use_filter) {
        const double ox = pts[0], oy = pts[1], oz = pts[2];
        float pm = 0.f, q2 = 0.f;
        for (int j = tid; j < n; j += RANSAC_THREADS) {
            const double x = pts[3 * j], y = pts[3 * j + 1], z = pts[3 * j + 2];
            const float qx = (float)(x - ox), qy = (float)(y - oy), qz = (float)(z - oz);
            s_q[j] = make_float4(qx, qy, qz, 0.f);
            pm = fmaxf(pm, (float)(fabs(x) + fabs(y) + fabs(z)));
            q2 = fmaxf(q2, fmaf(qx, qx, fmaf(qy, qy, qz * qz)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            pm = fmaxf(pm, __shfl_xor_sync(0xffffffffu, pm, o));
            q2 = fmaxf(q2, __shfl_xor_sync(0xffffffffu, q2, o));
        }
        if ((tid & 31) == 0) {  // non-negative floats order like their bit patterns; NaN patterns compare above +inf
            atomicMax(s_pmax, __float_as_uint(pm));
            atomicMax(s_qmax2, __float_as_uint(q2));
        }
        __syncthreads();
        Pm = __uint_as_float(*s_pmax) * 1.0001f;
        Qm = sqrtf(__uint_as_float(*s_qmax2)) * 1.0001f;
    }

    // ---- hypotheses in chunks of RANSAC_THREADS, in index order ----------------------------------
    const float thr_f = (float)A.thr;
    Interval iv[RANSAC_MAX_CHUNKS];
    unsigned long long my_key = 0ull;  // this thread's best exact (count, index) key and its plane
    float4 my_pl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (use_filter && !verify) {
        unsigned long long my_sel = 0ull;
        int processed = 0;
        bool stop = false;
#pragma unroll
        for (int k = 0; k < RANSAC_MAX_CHUNKS; ++k) {
            iv[k].lo = 0;
            iv[k].hi = -1;
            if (k < nch && !stop) {  // CTA-uniform
                const int t = k * RANSAC_THREADS + tid;
                bool full = false;
                if (t < A.H) {
                    iv[k] = filter_hypothesis(s_q, A.r32t, A.H, A.K, t, n, Pm, Qm, thr_f);
                    const unsigned long long key = pack_key(iv[k].lo, t);
                    my_sel = key > my_sel ? key : my_sel;
                    full = iv[k].lo == n;
                    if ((A.flags & RANSAC_FLAG_STATS) && iv[k].hi - iv[k].lo == n) atomicAdd(&s_stat[0], 1u);
                }
                processed = k + 1;
                // a hypothesis that certainly keeps every point cannot be beaten by a later index
                stop = __syncthreads_or(full) != 0;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, my_sel, o);
            my_sel = other > my_sel ? other : my_sel;
        }
        if ((tid & 31) == 0) atomicMax(s_sel, my_sel);
        __syncthreads();
        const unsigned long long sel = *s_sel;
        const int M = (int)(sel >> 32);
        const int t0 = (int)(0xffffffffu - (uint32_t)(sel & 0xffffffffull));
        // candidates: everything that can still be the (count desc, index asc) maximum
#pragma unroll
        for (int k = 0; k < RANSAC_MAX_CHUNKS; ++k) {
            const int t = k * RANSAC_THREADS + tid;
            if (k < processed && t < A.H) {
                const bool cand = (t < t0) ? (iv[k].hi >= M) : (t == t0 || iv[k].hi > M);
                if (cand) s_cand[atomicAdd(s_nc, 1u)] = (uint16_t)t;
            }
        }
        __syncthreads();
        const int nc = (int)*s_nc;
        for (int c0 = 0; c0 < nc; c0 += RANSAC_THREADS) {  // CTA-uniform trip count
            const int c = c0 + tid;
            const bool active = c < nc;
            float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
            int t = 0;
            if (active) {
                t = s_cand[c];
                pl = fit_plane(pts, A.table, A.K, t, n, rs, A.err);
            }
            const int cnt = exact_count(pts, n, active, pl, A.thr);
            if (active) {
                const unsigned long long key = pack_key(cnt, t);
                if (key > my_key) {
                    my_key = key;
                    my_pl = pl;
                }
            }
        }
        if (A.flags & RANSAC_FLAG_STATS) {
            if (tid == 0) {
                atomicAdd(&g_ransac_stats[0], 1ull);                                       // blocks
                atomicAdd(&g_ransac_stats[1], (unsigned long long)min(processed * RANSAC_THREADS, A.H));  // filtered
                atomicAdd(&g_ransac_stats[2], (unsigned long long)s_stat[0]);              // trivial intervals
                atomicAdd(&g_ransac_stats[3], (unsigned long long)nc);                     // exact evaluations
                atomicAdd(&g_ransac_stats[4], processed < nch ? 1ull : 0ull);              // early exits
                atomicAdd(&g_ransac_stats[6], (unsigned long long)nc);                     // exact fits
                atomicAdd(&g_ransac_stats[7], (unsigned long long)nc * n + n);             // exact distance evaluations (+ mask)
                atomicAdd(&g_ransac_stats[8], (unsigned long long)min(processed * RANSAC_THREADS, A.H) * n);  // float32 ones
            }
        }
    } else {
        // exact evaluation of every hypothesis (plus interval verification)
        for (int k = 0; k < nch; ++k) {
            const int t = k * RANSAC_THREADS + tid;
            const bool active = t < A.H;
            float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active) pl = fit_plane(pts, A.table, A.K, t, n, rs, A.err);
            const int cnt = exact_count(pts, n, active, pl, A.thr);
            bool full = false;
            if (active) {
                const unsigned long long key = pack_key(cnt, t);
                if (key > my_key) {
                    my_key = key;
                    my_pl = pl;
                }
                full = cnt == n;
                if (use_filter) {  // verify mode
                    const Interval v = filter_hypothesis(s_q, A.r32t, A.H, A.K, t, n, Pm, Qm, thr_f);
                    if (cnt < v.lo || cnt > v.hi) {
                        atomicOr(A.err, (uint32_t)DEVERR_FILTER_BOUND);
                        atomicAdd(&s_stat[1], 1u);
                    }
                    if (v.hi - v.lo == n) atomicAdd(&s_stat[0], 1u);
                }
            }
            // same early exit as the filtered path (not in verify mode: check every hypothesis)
            if (!verify && __syncthreads_or(full)) break;
        }
        if (verify && tid == 0) {
            atomicAdd(&g_ransac_stats[0], 1ull);
            atomicAdd(&g_ransac_stats[1], (unsigned long long)A.H);
            atomicAdd(&g_ransac_stats[2], (unsigned long long)s_stat[0]);
            atomicAdd(&g_ransac_stats[5], (unsigned long long)s_stat[1]);  // interval violations
            atomicAdd(&g_ransac_stats[6], (unsigned long long)A.H);
            atomicAdd(&g_ransac_stats[7], (unsigned long long)A.H * n + n);
            atomicAdd(&g_ransac_stats[8], (unsigned long long)A.H * n);
        }
    }

    // ---- block-wide maximum of the exact keys; its owner publishes the plane -------------------------
    {
        unsigned long long wb = my_key;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, wb, o);
            wb = other > wb ? other : wb;
        }
        if ((tid & 31) == 0) atomicMax(s_best, wb);
    }
    __syncthreads();
    const unsigned long long bk = *s_best;
    if (my_key == bk) {  // keys are unique per hypothesis: exactly one thread
        *s_bp = my_pl;
        if (A.plane) reinterpret_cast<float4*>(A.plane)[b] = my_pl;
        if (A.best) A.best[b] = (int)(0xffffffffu - (uint32_t)(bk & 0xffffffffull));
        if (A.best_count) A.best_count[b] = (int)(bk >> 32);
    }
    __syncthreads();
    const float4 bp = *s_bp;
    // ---- final mask (cuda_ransac.py:149-155) ---------------------------------------------------
    const double a = (double)bp.x, bb = (double)bp.y, c = (double)bp.z, d = (double)bp.w;
    for (int i = tid; i < n; i += RANSAC_THREADS) A.mask[(size_t)ps + i] = plane_distance(a, bb, c, d, pts + 3 * i) < A.thr ? 1 : 0;
}


// =============================================================================================
// small blocks (n <= RS_MAX_POINTS): one WARP per block, persistent warps pulling work items.
// Step 0 evaluates hypotheses 0..31 exactly.  If one of them keeps every point it is the winner (ties
// go to the lowest index) - on LiDAR leaves that ends > 99 % of the blocks after 32 of 1024 hypotheses.
// Otherwise hypotheses 32.. go through the FP32 interval pre-filter, 32 per step, and the surviving
// candidates are evaluated exactly in index order.
// =============================================================================================
constexpr int RS_WARPS = 8;
constexpr int RS_MAX_POINTS = 128;
constexpr int RS_GRAB = 8;  // work items a warp takes per atomic (size of the d_* arrays below)

struct __align__(16) RansacWarpSmem {
    double pts[RS_MAX_POINTS * 3 + 2];  // TMA destination (16-byte aligned run + optional 8-byte lead)
    float4 q[RS_MAX_POINTS];            // block-local float32 copies; reused as the uint16 candidate list
    uint8_t hi[RANSAC_MAX_H];           // upper bound of every filtered hypothesis
    unsigned long long bar;
    unsigned long long pad;
    long long d_rs[RS_GRAB];                 // descriptors of the RS_GRAB work items the warp took with one atomic
    uint32_t d_b[RS_GRAB], d_ps[RS_GRAB], d_pk[RS_GRAB];
    int d_n[RS_GRAB];
};

__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
        v = other > v ? other : v;
    }
    return v;
}

__device__ __forceinline__ int exact_count_serial(const double* __restrict__ pts, int n, const float4 pl, double thr) {
    const double a = (double)pl.x, b = (double)pl.y, c = (double)pl.z, d = (double)pl.w;
    int cnt = 0;
    for (int i = 0; i < n; ++i) cnt += plane_distance(a, b, c, d, pts + 3 * i) < thr;
    return cnt;
}

// STATS: tally what was executed (pre-filter statistics, exact fits, point-distance evaluations) for ol_ransac_stats_read;
// the plain instantiation carries no counters.  With A.sub the kernel walks that list of work items (the blocks the lane
// passes below left undecided) instead of all of them.
template <bool STATS>
__global__ void __launch_bounds__(RS_WARPS * 32, 3) ransac_small_kernel(RansacArgs A, uint32_t* __restrict__ counter) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RansacWarpSmem* sm = reinterpret_cast<RansacWarpSmem*>(smem_raw);
    const int lane = threadIdx.x & 31;
    RansacWarpSmem& W = sm[threadIdx.x >> 5];
    const uint32_t bar = smem_addr(&W.bar);
    const bool verify = (A.flags & RANSAC_FLAG_VERIFY) != 0;
    const bool filter_on = (verify || !(A.flags & RANSAC_FLAG_EXACT_ONLY)) && (__ldg(A.table_ok) != 0u);
    const bool tma = !(A.flags & RANSAC_FLAG_NO_TMA);
    const int nsteps = (A.H + 31) >> 5;
    const float thr_f = (float)A.thr;
    if (tma) {
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    uint32_t parity = 0;
    uint32_t st_blocks = 0, st_filtered = 0, st_trivial = 0, st_exact = 0, st_early = 0, st_viol = 0;  // tallies (STATS only)
    uint32_t st_fit = 0, st_dist = 0, st_fdist = 0;  // exact plane fits, exact and float32 point-distance evaluations
    const uint32_t n_items = A.sub ? (A.n_sub_dev ? __ldg(A.n_sub_dev) : A.n_sub) : A.n_work;
    // items per grab: RS_GRAB when every warp has several grabs' worth of work; fewer when the list is short (what the lane
    // passes leave of a small map), where a grab of 8 would leave most warps idle while a few walk 8 slow items in a row
    const uint32_t per_warp = n_items / (gridDim.x * (uint32_t)RS_WARPS);
    const uint32_t grab = per_warp >= 4u * RS_GRAB ? (uint32_t)RS_GRAB : (per_warp >= 8u ? 4u : (per_warp >= 4u ? 2u : 1u));
    for (;;) {
        // RS_GRAB work items per atomic; their descriptors are fetched by RS_GRAB lanes at once, so the dependent
        // counter -> work[] -> block table round trips are paid once per grab instead of once per block
        uint32_t w0 = 0;
        if (lane == 0) w0 = atomicAdd(counter, grab);
        w0 = __shfl_sync(0xffffffffu, w0, 0);
        if (w0 >= n_items) break;
        __syncwarp();
        if (lane < (int)grab && w0 + lane < n_items) {
            const uint32_t wi = A.sub ? A.sub[w0 + lane] : w0 + lane;
            const uint32_t db = A.work[wi];
            const uint32_t dps = A.blk_start[db];
            W.d_b[lane] = db;
            W.d_n[lane] = A.blk_size[db];
            W.d_ps[lane] = dps;
            W.d_rs[lane] = A.blk_ref_start[db];
            W.d_pk[lane] = A.pk_start ? A.pk_start[wi] : dps;
        }
        __syncwarp();
        const int items = (int)min(grab, n_items - w0);
      for (int it = 0; it < items; ++it) {
        const int n = W.d_n[it];
        if (n > RS_MAX_POINTS) continue;  // handled by the CTA-per-block kernel
        const uint32_t b = W.d_b[it];
        const uint32_t ps = W.d_ps[it];
        const long long rs = W.d_rs[it];
        const uint32_t pk = W.d_pk[it];
        const double* gsrc = A.points + (size_t)pk * 3;
        const double* pts;
        // ---- stage the block's float64 points (TMA bulk copy, completion on the warp's mbarrier) --
        if (tma) {
            const uintptr_t src_addr = reinterpret_cast<uintptr_t>(gsrc);
            const uint32_t lead = (uint32_t)(src_addr & 15u);  // 0 or 8
            const unsigned char* asrc = reinterpret_cast<const unsigned char*>(src_addr - lead);
            const uint32_t total = lead + (uint32_t)n * 24u;
            const uint32_t bulk = total & ~15u;
            if (lane == 0) {
                // order the previous item's generic-proxy reads before the async-proxy overwrite
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, bulk);
                tma_bulk_g2s(smem_addr(W.pts), asrc, bulk, bar);
                if (total != bulk) W.pts[bulk / 8] = reinterpret_cast<const double*>(asrc)[bulk / 8];
            }
            mbar_wait(bar, parity);
            parity ^= 1u;
            __syncwarp();
            pts = W.pts + lead / 8;
        } else {
            for (int i = lane; i < n * 3; i += 32) W.pts[i] = gsrc[i];
            __syncwarp();
            pts = W.pts;
        }

        unsigned long long my_key = 0ull;
        float4 my_pl = make_float4(0.f, 0.f, 0.f, 0.f);
        bool done = false;
        int k = 0;
        // ---- exact steps: step 0 always; every step in EXACT_ONLY / VERIFY mode ---------------------
        float Pm = 0.f, Qm = 0.f;
        bool have_q = false;
        auto prepare_filter = [&]() {
            const double ox = pts[0], oy = pts[1], oz = pts[2];
            float pm = 0.f, q2 = 0.f;
            for (int j = lane; j < n; j += 32) {
                const double x = pts[3 * j], y = pts[3 * j + 1], z = pts[3 * j + 2];
                const float qx = (float)(x - ox), qy = (float)(y - oy), qz = (float)(z - oz);
                W.q[j] = make_float4(qx, qy, qz, 0.f);
                pm = fmaxf(pm, (float)(fabs(x) + fabs(y) + fabs(z)));
                q2 = fmaxf(q2, fmaf(qx, qx, fmaf(qy, qy, qz * qz)));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                // bit-pattern max: NaN patterns order above +inf, so a NaN coordinate poisons the bound (=> trivial intervals)
                pm = __uint_as_float(max(__float_as_uint(pm), __shfl_xor_sync(0xffffffffu, __float_as_uint(pm), o)));
                q2 = __uint_as_float(max(__float_as_uint(q2), __shfl_xor_sync(0xffffffffu, __float_as_uint(q2), o)));
            }
            Pm = pm * 1.0001f;
            Qm = sqrtf(q2) * 1.0001f;
            have_q = true;
            __syncwarp();
        };
        if (verify && filter_on) prepare_filter();
        const int exact_steps = (filter_on && !verify) ? 1 : nsteps;
        for (; k < exact_steps && !done; ++k) {
            const int t = (k << 5) + lane;
            if (t < A.H) {
                const float4 pl = fit_plane_smem(pts, A.table, A.K, t, n, rs, A.err);
                const int cnt = exact_count_serial(pts, n, pl, A.thr);
                const unsigned long long key = pack_key(cnt, t);
                if (key > my_key) {
                    my_key = key;
                    my_pl = pl;
                }
                if (verify && filter_on) {
                    const Interval v = filter_hypothesis(W.q, A.r32t, A.H, A.K, t, n, Pm, Qm, thr_f);
                    if (cnt < v.lo || cnt > v.hi) {
                        atomicOr(A.err, (uint32_t)DEVERR_FILTER_BOUND);
                        if (STATS) ++st_viol;
                    }
                    if (STATS) st_trivial += (v.hi - v.lo == n), ++st_filtered, st_fdist += n;
                }
                if (STATS) ++st_fit, st_dist += n;
            }
            // a hypothesis that keeps every point cannot be beaten by a later index
            if (!verify && __any_sync(0xffffffffu, (int)(my_key >> 32) == n)) done = true;
        }
        // ---- FP32 interval pre-filter for the remaining hypotheses ---------------------------------------
        if (!done && k < nsteps) {
            if (!have_q) prepare_filter();
            const int cnt0 = (int)(warp_max_u64(my_key) >> 32);  // best exact count so far (lower indices)
            unsigned long long my_sel = 0ull;                     // best (lower bound, index) among the filtered ones
            int k_end = nsteps;
            for (int kk = k; kk < nsteps; ++kk) {
                const int t = (kk << 5) + lane;
                bool full = false;
                if (t < A.H) {
                    const Interval v = filter_hypothesis(W.q, A.r32t, A.H, A.K, t, n, Pm, Qm, thr_f);
                    W.hi[t] = (uint8_t)v.hi;
                    const unsigned long long key = pack_key(v.lo, t);
                    my_sel = key > my_sel ? key : my_sel;
                    full = v.lo == n;
                    if (STATS) st_trivial += (v.hi - v.lo == n), ++st_filtered, st_fdist += n;
                }
                if (__any_sync(0xffffffffu, full)) {
                    k_end = kk + 1;
                    if (STATS) st_early += (lane == 0);
                    break;
                }
            }
            const unsigned long long sel = warp_max_u64(my_sel);
            const int M = (int)(sel >> 32);
            const int t0 = (int)(0xffffffffu - (uint32_t)(sel & 0xffffffffull));
            __syncwarp();
            // candidates, in index order, into the (now dead) q area
            uint16_t* list = reinterpret_cast<uint16_t*>(W.q);
            int nc = 0;
            for (int kk = k; kk < k_end; ++kk) {
                const int t = (kk << 5) + lane;
                bool cand = false;
                if (t < A.H) {
                    const int hi = W.hi[t];
                    cand = hi > cnt0 && ((t < t0) ? (hi >= M) : (t == t0 || hi > M));
                }
                const uint32_t m = __ballot_sync(0xffffffffu, cand);
                if (cand) list[nc + __popc(m & ((1u << lane) - 1u))] = (uint16_t)t;
                nc += __popc(m);
            }
            __syncwarp();
            for (int c0 = 0; c0 < nc; c0 += 32) {
                const int c = c0 + lane;
                if (c < nc) {
                    const int t = list[c];
                    const float4 pl = fit_plane_smem(pts, A.table, A.K, t, n, rs, A.err);
                    const int cnt = exact_count_serial(pts, n, pl, A.thr);
                    const unsigned long long key = pack_key(cnt, t);
                    if (key > my_key) {
                        my_key = key;
                        my_pl = pl;
                    }
                    if (STATS) ++st_exact, ++st_fit, st_dist += n;
                }
                if (__any_sync(0xffffffffu, (int)(my_key >> 32) == n)) break;  // later indices cannot win
            }
        } else if (done) {
            if (STATS) st_early += (lane == 0);
        }
        // ---- winner, outputs, mask (cuda_ransac.py:149-155) -----------------------------------------------
        const unsigned long long bk = warp_max_u64(my_key);
        const int src = __ffs(__ballot_sync(0xffffffffu, my_key == bk)) - 1;
        float4 bp;
        bp.x = __shfl_sync(0xffffffffu, my_pl.x, src);
        bp.y = __shfl_sync(0xffffffffu, my_pl.y, src);
        bp.z = __shfl_sync(0xffffffffu, my_pl.z, src);
        bp.w = __shfl_sync(0xffffffffu, my_pl.w, src);
        if (lane == 0) {
            if (A.plane) reinterpret_cast<float4*>(A.plane)[b] = bp;
            if (A.best) A.best[b] = (int)(0xffffffffu - (uint32_t)(bk & 0xffffffffull));
            if (A.best_count) A.best_count[b] = (int)(bk >> 32);
        }
        const double a = (double)bp.x, bb = (double)bp.y, c = (double)bp.z, d = (double)bp.w;
        for (int i = lane; i < n; i += 32) A.mask[(size_t)ps + i] = plane_distance(a, bb, c, d, pts + 3 * i) < A.thr ? 1 : 0;
        if (STATS) st_blocks += (lane == 0), st_dist += (lane == 0) ? n : 0;  // + the mask pass
        __syncwarp();
      }
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            st_filtered += __shfl_xor_sync(0xffffffffu, st_filtered, o);
            st_trivial += __shfl_xor_sync(0xffffffffu, st_trivial, o);
            st_exact += __shfl_xor_sync(0xffffffffu, st_exact, o);
            st_viol += __shfl_xor_sync(0xffffffffu, st_viol, o);
            st_blocks += __shfl_xor_sync(0xffffffffu, st_blocks, o);
            st_early += __shfl_xor_sync(0xffffffffu, st_early, o);
            st_fit += __shfl_xor_sync(0xffffffffu, st_fit, o);
            st_dist += __shfl_xor_sync(0xffffffffu, st_dist, o);
            st_fdist += __shfl_xor_sync(0xffffffffu, st_fdist, o);
        }
        if (lane == 0) {
            atomicAdd(&g_ransac_stats[0], (unsigned long long)st_blocks);
            atomicAdd(&g_ransac_stats[1], (unsigned long long)st_filtered);
            atomicAdd(&g_ransac_stats[2], (unsigned long long)st_trivial);
            atomicAdd(&g_ransac_stats[3], (unsigned long long)st_exact);
            atomicAdd(&g_ransac_stats[4], (unsigned long long)st_early);
            atomicAdd(&g_ransac_stats[5], (unsigned long long)st_viol);
            atomicAdd(&g_ransac_stats[6], (unsigned long long)st_fit);
            atomicAdd(&g_ransac_stats[7], (unsigned long long)st_dist);
            atomicAdd(&g_ransac_stats[8], (unsigned long long)st_fdist);
        }
    }
}

// =============================================================================================
// Lane-per-block passes: hypothesis t for every block that is still undecided, ONE LANE PER BLOCK.
//
// The winner of a block is the lowest-index hypothesis with the maximal inlier count, and no count exceeds n: the first
// hypothesis that keeps all n points ends the block.  On LiDAR leaves (mean 7 points, K = 6 samples) that is hypothesis
// 0 for 57 % of the blocks, one of 0..2 for 95 %, one of 0..7 for 98.9 % (tools/block_histogram.py on the 100 M-point
// workload).  A warp that gives a block 8 or 32 lanes evaluates 8 or 32 exact fits where 1-3 are needed; here pass t
// evaluates hypothesis t with the reference's exact float64 arithmetic for exactly the blocks that passes 0..t-1 left
// undecided (a compacted list), so the number of exact fits per block is its winning index + 1.  A decided block gets
// its outputs at once (plane, index, count = n, mask all ones: every point is an inlier).  What is still undecided after
// RANSAC_LANE_PASSES passes, and every block above RANSAC_LANE_MAX points, goes to the warp-per-block kernel.
// The blocks are read straight from the packed point array (each lane its own contiguous run; neighbouring lanes read
// neighbouring runs), no shared memory.
// =============================================================================================
constexpr int RANSAC_LANE_PASSES = 8;
constexpr int RANSAC_LANE_SINGLE = 3;  // hypotheses 0..2 get a launch each, 3..7 share one (ransac_lane_group_kernel)
constexpr int RANSAC_LANE_MAX = 32;

__global__ void __launch_bounds__(256) ransac_lane_kernel(RansacArgs A, int t, const uint32_t* __restrict__ list_in,
                                                          const uint32_t* __restrict__ n_in_dev, uint32_t n_in_host,
                                                          uint32_t* __restrict__ list_out, uint32_t* __restrict__ n_out) {
    const uint32_t n_in = n_in_dev ? __ldg(n_in_dev) : n_in_host;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool keep = false;  // stays on the list
    uint32_t w = 0;
    int n = 0;
    bool fitted = false;
    if (i < n_in) {
        w = list_in ? list_in[i] : i;
        const uint32_t b = A.work[w];
        n = A.blk_size[b];
        keep = true;
        if (n <= RANSAC_LANE_MAX) {
            const uint32_t ps = A.blk_start[b];
            const double* pts = A.points + (size_t)(A.pk_start ? A.pk_start[w] : ps) * 3;
            const float4 pl = fit_plane(pts, A.table, A.K, t, n, A.blk_ref_start[b], A.err);
            const int cnt = exact_count_serial(pts, n, pl, A.thr);
            fitted = true;
            if (cnt == n) {  // keeps every point: no later hypothesis can win, earlier ones did not (the block was still listed)
                keep = false;
                if (A.plane) reinterpret_cast<float4*>(A.plane)[b] = pl;
                if (A.best) A.best[b] = t;
                if (A.best_count) A.best_count[b] = n;
                for (int j = 0; j < n; ++j) A.mask[(size_t)ps + j] = 1;
            }
        }
    }
    // order-preserving-enough compaction: one atomic per warp (the order of the list does not matter)
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (m) {
        uint32_t base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(n_out, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (keep) list_out[base + __popc(m & ((1u << lane) - 1u))] = w;
    }
    if (A.flags & RANSAC_FLAG_STATS) {
        uint32_t fits = fitted ? 1u : 0u, dist = fitted ? (uint32_t)n : 0u, blocks = (fitted && !keep) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fits += __shfl_xor_sync(0xffffffffu, fits, o);
            dist += __shfl_xor_sync(0xffffffffu, dist, o);
            blocks += __shfl_xor_sync(0xffffffffu, blocks, o);
        }
        if (lane == 0 && fits) {
            atomicAdd(&g_ransac_stats[0], (unsigned long long)blocks);
            atomicAdd(&g_ransac_stats[4], (unsigned long long)blocks);  // decided = early exits
            atomicAdd(&g_ransac_stats[6], (unsigned long long)fits);
            atomicAdd(&g_ransac_stats[7], (unsigned long long)dist);
        }
    }
}

// The tail of the lane passes in ONE launch: after three passes only a few percent of the blocks are still undecided, and a
// launch per hypothesis over them is all launch latency.  Eight lanes per listed block evaluate hypotheses t0 .. t1-1 side
// by side (exact arithmetic, like ransac_lane_kernel); the lowest one that keeps every point decides the block - exactly
// the hypothesis the sequential passes would have stopped at - and a block none of them decides stays on the list.
__global__ void __launch_bounds__(256) ransac_lane_group_kernel(RansacArgs A, int t0, int t1, const uint32_t* __restrict__ list_in,
                                                                const uint32_t* __restrict__ n_in_dev, uint32_t* __restrict__ list_out,
                                                                uint32_t* __restrict__ n_out) {
    const uint32_t n_in = __ldg(n_in_dev);
    const int lane = threadIdx.x & 31, sub = lane & 7;
    const int t = t0 + sub;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    // four listed blocks per warp and round; the list's length only exists on the device, so the warps stride over it
    for (uint32_t first = warp_global * 4u; first < n_in; first += n_warps * 4u) {
    const uint32_t item = first + (uint32_t)(lane >> 3);
    bool listed = item < n_in, decided = false, fitted = false;
    uint32_t w = 0, b = 0, ps = 0;
    int n = 0;
    float4 pl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (listed) {
        w = list_in[item];
        b = A.work[w];
        n = A.blk_size[b];
        if (n <= RANSAC_LANE_MAX && t < t1) {
            ps = A.blk_start[b];
            const double* pts = A.points + (size_t)(A.pk_start ? A.pk_start[w] : ps) * 3;
            pl = fit_plane(pts, A.table, A.K, t, n, A.blk_ref_start[b], A.err);
            decided = exact_count_serial(pts, n, pl, A.thr) == n;
            fitted = true;
        }
    }
    const uint32_t dm = __ballot_sync(0xffffffffu, decided);
    const uint32_t mine = (dm >> (lane & ~7)) & 0xffu;  // decided hypotheses of this lane's block
    if (listed && mine && sub == __ffs(mine) - 1) {  // the lowest deciding hypothesis writes the block's result
        if (A.plane) reinterpret_cast<float4*>(A.plane)[b] = pl;
        if (A.best) A.best[b] = t;
        if (A.best_count) A.best_count[b] = n;
        for (int j = 0; j < n; ++j) A.mask[(size_t)ps + j] = 1;
    }
    const bool keep = listed && !mine && sub == 0;
    const uint32_t m = __ballot_sync(0xffffffffu, keep);
    if (m) {
        uint32_t base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(n_out, (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (keep) list_out[base + __popc(m & ((1u << lane) - 1u))] = w;
    }
    if (A.flags & RANSAC_FLAG_STATS) {
        uint32_t fits = fitted ? 1u : 0u, dist = fitted ? (uint32_t)n : 0u, blocks = (listed && mine && sub == 0) ? 1u : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fits += __shfl_xor_sync(0xffffffffu, fits, o);
            dist += __shfl_xor_sync(0xffffffffu, dist, o);
            blocks += __shfl_xor_sync(0xffffffffu, blocks, o);
        }
        if (lane == 0 && fits) {
            atomicAdd(&g_ransac_stats[0], (unsigned long long)blocks);
            atomicAdd(&g_ransac_stats[4], (unsigned long long)blocks);
            atomicAdd(&g_ransac_stats[6], (unsigned long long)fits);
            atomicAdd(&g_ransac_stats[7], (unsigned long long)dist);
        }
    }
    }  // rounds
}

// work items whose block is too large for the warp kernel -> compact list for the CTA kernel
__global__ void ransac_split_kernel(const uint32_t* __restrict__ work, uint32_t n_work, const int32_t* __restrict__ blk_size,
                                    uint32_t* __restrict__ large, uint32_t* __restrict__ n_large) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool big = i < n_work && blk_size[work[i]] > RS_MAX_POINTS;
    const uint32_t m = __ballot_sync(0xffffffffu, big);
    if (!m) return;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(n_large, (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
    if (big) large[base + __popc(m & ((1u << lane) - 1u))] = i;  // the work item, not the block id
}

void launch_ransac(Ctx& c, const double* points, int64_t n_points, const uint32_t* blk_phys_start, const int32_t* blk_size,
                   const long long* blk_ref_start, const uint32_t* work_list, const uint32_t* pk_start, uint32_t n_work,
                   uint32_t max_block,
                   const double* table, int H, int K, double threshold, uint8_t* mask, float* plane, int32_t* best,
                   int32_t* best_count, uint32_t flags) {
    if (n_work == 0) return;
    OL_REQUIRE(H >= 1 && H <= RANSAC_MAX_H, OL_ERR_INVALID, "hypotheses_number must be in 1..1024");
    DevBuf<uint32_t> r32t(c, (size_t)H * K), table_ok(c, 1);
    {
        const uint32_t one = 1u;
        OL_CUDA(cudaMemcpyAsync(table_ok.get(), &one, 4, cudaMemcpyHostToDevice, c.stream));
        ransac_table_prep_kernel<<<(H * K + 255) / 256, 256, 0, c.stream>>>(table, H, K, r32t.get(), table_ok.get());
        OL_CHECK_LAUNCH();
    }
    RansacArgs a{};
    a.points = points;
    a.n_points = n_points;
    a.blk_start = blk_phys_start;
    a.blk_size = blk_size;
    a.blk_ref_start = blk_ref_start;
    a.work = work_list;
    a.n_work = n_work;
    a.pk_start = pk_start;
    a.table = table;
    a.r32t = r32t.get();
    a.table_ok = table_ok.get();
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
    // ---- lane-per-block passes (default mode): hypotheses 0 .. RANSAC_LANE_PASSES-1, one exact fit per undecided block
    // and pass; the lists live on the device, no host synchronisation between the passes ---------------------------------
    DevBuf<uint32_t> counters(c, 2 + RANSAC_LANE_PASSES), list_a, list_b;
    counters.zero();
    const bool lane_passes = !(flags & (RANSAC_FLAG_EXACT_ONLY | RANSAC_FLAG_VERIFY)) && H > RANSAC_LANE_PASSES;
    if (lane_passes) {
        list_a.reset(c, n_work);
        list_b.reset(c, n_work);
        const unsigned grid = (n_work + 255) / 256;
        const uint32_t* in = nullptr;
        const uint32_t* n_in = nullptr;
        for (int t = 0; t < RANSAC_LANE_SINGLE; ++t) {
            uint32_t* out = (t & 1) ? list_b.get() : list_a.get();
            uint32_t* n_out = counters.get() + 2 + t;
            // pass t > 0 is launched over the upper bound n_work; the threads beyond the list's device-side length retire at once
            ransac_lane_kernel<<<grid, 256, 0, c.stream>>>(a, t, in, n_in, n_work, out, n_out);
            OL_CHECK_LAUNCH();
            in = out;
            n_in = n_out;
        }
        {  // hypotheses RANSAC_LANE_SINGLE .. RANSAC_LANE_PASSES-1 of what is left, eight lanes per block, one launch.  The
           // list is short by now (a few percent of the blocks), but its length only exists on the device: a fixed grid
           // strides over it
            uint32_t* out = (RANSAC_LANE_SINGLE & 1) ? list_b.get() : list_a.get();
            uint32_t* n_out = counters.get() + 2 + RANSAC_LANE_SINGLE;
            const unsigned ggrid = std::min<unsigned>(grid, (unsigned)c.num_sms * 16);
            ransac_lane_group_kernel<<<ggrid, 256, 0, c.stream>>>(a, RANSAC_LANE_SINGLE, RANSAC_LANE_PASSES, in, n_in, out, n_out);
            OL_CHECK_LAUNCH();
            in = out;
            n_in = n_out;
        }
        a.sub = in;
        a.n_sub = n_work;
        a.n_sub_dev = n_in;
    }
    // ---- warp-per-block kernel: persistent warps over the remaining work items (skips the large blocks) ------------
    {
        const size_t smem_small = sizeof(RansacWarpSmem) * RS_WARPS;
        const bool stats = (flags & (RANSAC_FLAG_STATS | RANSAC_FLAG_VERIFY)) != 0;
        auto kernel = stats ? ransac_small_kernel<true> : ransac_small_kernel<false>;
        OL_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_small));
        int per_sm = 4;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RS_WARPS * 32, smem_small);
        if (per_sm < 1) per_sm = 1;
        const unsigned long long want = ((unsigned long long)n_work + RS_WARPS - 1) / RS_WARPS;
        const unsigned grid = (unsigned)std::min<unsigned long long>(want, (unsigned long long)c.num_sms * per_sm);
        kernel<<<grid, RS_WARPS * 32, smem_small, c.stream>>>(a, counters.get());
        OL_CHECK_LAUNCH();
        a.sub = nullptr;
        a.n_sub_dev = nullptr;
    }
    // ---- CTA-per-block kernel for blocks above RS_MAX_POINTS points -------------------------------------------
    if (max_block > (uint32_t)RS_MAX_POINTS) {
        DevBuf<uint32_t> large(c, n_work);
        ransac_split_kernel<<<(n_work + 255) / 256, 256, 0, c.stream>>>(work_list, n_work, blk_size, large.get(), counters.get() + 1);
        OL_CHECK_LAUNCH();
        uint32_t n_large = 0;
        OL_CUDA(cudaMemcpyAsync(&n_large, counters.get() + 1, 4, cudaMemcpyDeviceToHost, c.stream));
        c.sync();
        if (n_large) {
            a.sub = large.get();
            a.n_sub = n_large;
            size_t smem = RANSAC_SMEM_HEADER + 2 * RANSAC_MAX_H + ((((size_t)a.cap * 3 + 2) * 8 + 15) & ~(size_t)15) + (size_t)a.cap * 16;
            smem = (smem + 15) & ~(size_t)15;
            OL_CUDA(cudaFuncSetAttribute(ransac_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ransac_kernel<<<n_large, RANSAC_THREADS, smem, c.stream>>>(a);
            OL_CHECK_LAUNCH();
            c.sync();  // `large` is released on return
        }
    }
}

void ransac_stats_read(unsigned long long out[16], bool reset) {
    OL_CUDA(cudaDeviceSynchronize());
    OL_CUDA(cudaMemcpyFromSymbol(out, g_ransac_stats, sizeof(unsigned long long) * 16));
    if (reset) {
        unsigned long long z[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        OL_CUDA(cudaMemcpyToSymbol(g_ransac_stats, z, sizeof(z)));
    }
}

}  // namespace ol
