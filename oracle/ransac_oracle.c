/*
 * CPU ORACLE (test infrastructure, NOT product code) -- RANSAC half.
 *
 * Plain-C restatement of the reference's per-leaf RANSAC kernel
 *   /root/reference/octreelib/ransac/cuda_ransac.py:85-155   (kernel body)
 *   /root/reference/octreelib/ransac/util.py:16-24           (measure_distance)
 *   /root/reference/octreelib/ransac/util.py:28-84           (get_plane_from_points)
 * in the arithmetic the reference's own CI runs it in (NUMBA_ENABLE_CUDASIM=1: IEEE float64,
 * no FMA contraction, plane rounded to float32 before scoring).  Build with
 * -ffp-contract=off (see oracle/Makefile).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * The reference picks "any" hypothesis with the maximal inlier count (CAS race,
 * cuda_ransac.py:135-146).  The oracle reports, per block, every hypothesis' count so that
 * a tie-aware comparison is possible, and designates best = lowest index among the maxima.
 */
#include <math.h>
#include <pthread.h>
#include <stdatomic.h>
#include <stdint.h>
#include <string.h>

/* util.py:28-84 -- float64 arithmetic, result rounded to float32 by the caller's store
 * (cuda_ransac.py:110-113). */
static void plane_from_points(const double *pts, const int64_t *idx, int K, float plane[4]) {
    double cx = 0.0, cy = 0.0, cz = 0.0;
    for (int i = 0; i < K; ++i) { /* util.py:37-40 */
        const double *p = pts + 3 * idx[i];
        cx += p[0];
        cy += p[1];
        cz += p[2];
    }
    cx /= (double)K; /* util.py:42-44 */
    cy /= (double)K;
    cz /= (double)K;
    double xx = 0.0, xy = 0.0, xz = 0.0, yy = 0.0, yz = 0.0, zz = 0.0;
    for (int i = 0; i < K; ++i) { /* util.py:48-57 */
        const double *p = pts + 3 * idx[i];
        double rx = p[0] - cx, ry = p[1] - cy, rz = p[2] - cz;
        xx += rx * rx;
        xy += rx * ry;
        xz += rx * rz;
        yy += ry * ry;
        yz += ry * rz;
        zz += rz * rz;
    }
    double det_x = yy * zz - yz * yz; /* util.py:59-61 */
    double det_y = xx * zz - xz * xz;
    double det_z = xx * yy - xy * xy;
    double ax, ay, az;
    if (det_x > det_y && det_x > det_z) { /* util.py:63-74 */
        ax = det_x;
        ay = xz * yz - xy * zz;
        az = xy * yz - xz * yy;
    } else if (det_y > det_z) {
        ax = xz * yz - xy * zz;
        ay = det_y;
        az = xy * xz - yz * xx;
    } else {
        ax = xy * yz - xz * yy;
        ay = xy * xz - yz * xx;
        az = det_z;
    }
    double norm = sqrt(ax * ax + ay * ay + az * az); /* util.py:76 */
    if (norm == 0) { /* util.py:77-78 */
        plane[0] = plane[1] = plane[2] = plane[3] = 0.0f;
        return;
    }
    ax /= norm; /* util.py:80-82 */
    ay /= norm;
    az /= norm;
    double d = -(ax * cx + ay * cy + az * cz); /* util.py:83 */
    plane[0] = (float)ax;
    plane[1] = (float)ay;
    plane[2] = (float)az;
    plane[3] = (float)d;
}

/* util.py:16-24 with a float32 plane and float64 point: products promote to float64. */
static inline double distance(const float pl[4], const double *p) {
    return fabs((double)pl[0] * p[0] + (double)pl[1] * p[1] + (double)pl[2] * p[2] + (double)pl[3]);
}

/*
 * points        [n_points][3] float64, blocks laid out back to back (grid.py:173-191)
 * block_sizes   [n_blocks] int32
 * block_starts  [n_blocks] int64  = exclusive cumsum of block_sizes (cuda_ransac.py:65-67)
 * table         [H][K] float64 uniform [0,1) (cuda_ransac.py:39-41)
 * out_mask      [n_points] uint8   (zero-initialised here; cuda_ransac.py:57)
 * out_best      [n_blocks] int32   lowest hypothesis index with the maximal count, -1 if skipped
 * out_best_cnt  [n_blocks] int32
 * out_plane     [n_blocks][4] float32 plane of out_best
 * out_counts    [n_blocks][H] int32 or NULL
 * out_planes    [n_blocks][H][4] float32 or NULL
 * n_threads     host threads (blocks are independent; handed out by an atomic counter)
 * returns the number of sample indices that fell outside [0, n_points) (clamped; the reference
 * would read out of bounds there).
 */
typedef struct {
    const double *points;
    int64_t n_points;
    const int32_t *block_sizes;
    const int64_t *block_starts;
    int64_t n_blocks;
    const double *table;
    int H, K;
    double threshold;
    uint8_t *out_mask;
    int32_t *out_best, *out_best_cnt;
    float *out_plane;
    int32_t *out_counts;
    float *out_planes;
    atomic_llong next;
    atomic_llong oob;
} job_t;

static void one_block(job_t *J, int64_t b) {
    const double *points = J->points;
    const int H = J->H, K = J->K;
    const int32_t n = J->block_sizes[b];
    const int64_t start = J->block_starts[b];
    J->out_best[b] = -1;
    J->out_best_cnt[b] = 0;
    if (J->out_plane) memset(J->out_plane + 4 * b, 0, 4 * sizeof(float));
    if (J->out_counts) memset(J->out_counts + b * (int64_t)H, 0, (size_t)H * sizeof(int32_t));
    if (J->out_planes) memset(J->out_planes + b * (int64_t)H * 4, 0, (size_t)H * 4 * sizeof(float));
    if (n < K) return; /* cuda_ransac.py:96-97 */
    int32_t best = -1, best_cnt = -1;
    float best_plane[4] = {0, 0, 0, 0};
    int64_t idx[64];
    for (int t = 0; t < H; ++t) {
        for (int i = 0; i < K; ++i) { /* cuda_ransac.py:103-107: float64, then int32 truncation */
            double v = J->table[(int64_t)t * K + i] * (double)n + (double)start;
            int64_t j = (int64_t)(int32_t)v;
            if (j < 0 || j >= J->n_points) {
                atomic_fetch_add(&J->oob, 1);
                j = j < 0 ? 0 : J->n_points - 1;
            }
            idx[i] = j;
        }
        float pl[4];
        plane_from_points(points, idx, K, pl);
        int32_t cnt = 0;
        for (int32_t i = 0; i < n; ++i) /* cuda_ransac.py:116-121 */
            if (distance(pl, points + 3 * (start + i)) < J->threshold) ++cnt;
        if (J->out_counts) J->out_counts[b * (int64_t)H + t] = cnt;
        if (J->out_planes) memcpy(J->out_planes + (b * (int64_t)H + t) * 4, pl, sizeof(pl));
        if (cnt > best_cnt) { /* lowest index among the maxima */
            best_cnt = cnt;
            best = t;
            memcpy(best_plane, pl, sizeof(pl));
        }
    }
    J->out_best[b] = best;
    J->out_best_cnt[b] = best_cnt;
    if (J->out_plane) memcpy(J->out_plane + 4 * b, best_plane, sizeof(best_plane));
    for (int32_t i = 0; i < n; ++i) /* cuda_ransac.py:149-155 */
        if (distance(best_plane, points + 3 * (start + i)) < J->threshold) J->out_mask[start + i] = 1;
}

static void *worker(void *arg) {
    job_t *J = (job_t *)arg;
    for (;;) {
        int64_t b0 = atomic_fetch_add(&J->next, 16);
        if (b0 >= J->n_blocks) break;
        int64_t b1 = b0 + 16 < J->n_blocks ? b0 + 16 : J->n_blocks;
        for (int64_t b = b0; b < b1; ++b) one_block(J, b);
    }
    return 0;
}

int64_t ol_oracle_ransac(const double *points, int64_t n_points, const int32_t *block_sizes,
                         const int64_t *block_starts, int64_t n_blocks, const double *table, int H, int K,
                         double threshold, uint8_t *out_mask, int32_t *out_best, int32_t *out_best_cnt,
                         float *out_plane, int32_t *out_counts, float *out_planes, int n_threads) {
    job_t J = {points, n_points, block_sizes, block_starts, n_blocks, table, H, K, threshold, out_mask,
               out_best, out_best_cnt, out_plane, out_counts, out_planes, 0, 0};
    memset(out_mask, 0, (size_t)n_points);
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) {
        worker(&J);
    } else {
        pthread_t th[256];
        int started = 0;
        for (int i = 0; i < n_threads; ++i)
            if (pthread_create(&th[started], 0, worker, &J) == 0) ++started;
        if (started == 0) worker(&J);
        for (int i = 0; i < started; ++i) pthread_join(th[i], 0);
    }
    return (int64_t)atomic_load(&J.oob);
}

/* mask of one block for an arbitrary plane (used for the tie-aware comparison). */
void ol_oracle_mask_for_plane(const double *points, int64_t start, int32_t n, const float plane[4],
                              double threshold, uint8_t *out_mask) {
    for (int32_t i = 0; i < n; ++i) out_mask[i] = distance(plane, points + 3 * (start + i)) < threshold;
}

int ol_oracle_max_k(void) { return 64; }
