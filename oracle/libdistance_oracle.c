/*
 * oracle/libdistance_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * CPU restatement (plain C, single thread, like the reference) of the
 * msmbuilder.libdistance vector-metric primitives and of the npass==0 branch
 * of the C-Clustering-Library k-medoids that MiniBatchKMedoids drives.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product path (msmbuilder_b200)
 * never touches it.
 *
 * Pinned against the real reference compiled from /root/reference
 * (oracle/_ref/libref.so, see oracle/build_oracle.py) bit-for-bit in
 * tests/test_oracle_libdistance.py, and against scipy.spatial.distance the way
 * the reference's own tests do (msmbuilder/tests/test_libdistance.py:28-196).
 *
 * Reference arithmetic followed here (all paths relative to /root/reference):
 *   metric kernels        msmbuilder/libdistance/src/distance_kernels.h:41-242
 *                         - f32 inputs: the element difference/sum is formed in
 *                           f32, then widened; every accumulation is f64.
 *   metric-name dispatch  distance_kernels.h:245-293
 *   dist                  msmbuilder/libdistance/src/dist.hpp:4-80
 *   assign_nearest        msmbuilder/libdistance/src/assign.hpp:6-91
 *                         (strict '<' => lowest centre index wins ties; inertia
 *                         is the running f64 sum of the per-row minima)
 *   cdist                 msmbuilder/libdistance/src/cdist.hpp:4-50
 *   pdist (+X_indices)    msmbuilder/libdistance/src/pdist.hpp:4-96
 *   sumdist               msmbuilder/libdistance/src/sumdist.hpp:3-46
 *   kmedoids (npass==0)   msmbuilder/cluster/src/kmedoids.cc:60-69,74-260,264-309
 *   contigify_ids         msmbuilder/cluster/src/kmedoids.cc:386-401
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t idx_t; /* npy_intp on LP64 */

enum {
    M_EUCLIDEAN = 0, M_SQEUCLIDEAN, M_CITYBLOCK, M_CHEBYSHEV,
    M_CANBERRA, M_BRAYCURTIS, M_HAMMING, M_JACCARD, M_COUNT
};

static const char *const kNames[M_COUNT] = {
    "euclidean", "sqeuclidean", "cityblock", "chebyshev",
    "canberra", "braycurtis", "hamming", "jaccard"
};

int oracle_metric_id(const char *name)
{
    for (int m = 0; m < M_COUNT; ++m)
        if (strcmp(name, kNames[m]) == 0) return m;
    return -1; /* distance_kernels.h:268/292 returns NULL */
}

/*
 * One generic body per scalar type.  T is the storage type; "T-arithmetic"
 * below means the expression is evaluated in T before being widened, which is
 * what `double d = u[i] - v[i];` does for float operands in the reference.
 */
#define DEFINE_METRIC(T, SUF)                                                   \
static double metric_##SUF(int metric, const T *u, const T *v, idx_t n)         \
{                                                                               \
    double a = 0.0, b = 0.0;                                                    \
    switch (metric) {                                                           \
    case M_EUCLIDEAN:                                                           \
    case M_SQEUCLIDEAN:                                                         \
        for (idx_t i = 0; i < n; ++i) {                                         \
            T df = u[i] - v[i];            /* T-arithmetic */                   \
            double d = (double)df;                                              \
            a += d * d;                                                         \
        }                                                                       \
        return metric == M_EUCLIDEAN ? sqrt(a) : a;                             \
    case M_CITYBLOCK:                                                           \
        for (idx_t i = 0; i < n; ++i) {                                         \
            T df = u[i] - v[i];                                                 \
            a = a + fabs((double)df);                                           \
        }                                                                       \
        return a;                                                               \
    case M_CHEBYSHEV:                                                           \
        for (idx_t i = 0; i < n; ++i) {                                         \
            T df = u[i] - v[i];                                                 \
            double d = fabs((double)df);                                        \
            if (d > a) a = d;                                                   \
        }                                                                       \
        return a;                                                               \
    case M_CANBERRA:                                                            \
        for (idx_t i = 0; i < n; ++i) {                                         \
            T df = u[i] - v[i];                                                 \
            double num = fabs((double)df);                                      \
            T au = u[i] < 0 ? -u[i] : u[i], av = v[i] < 0 ? -v[i] : v[i];       \
            T dn = au + av;                /* T-arithmetic: C++ fabs(float) */  \
            double den = (double)dn;                                            \
            if (den > 0.0) a += num / den;                                      \
        }                                                                       \
        return a;                                                               \
    case M_BRAYCURTIS:                                                          \
        for (idx_t i = 0; i < n; ++i) {                                         \
            T df = u[i] - v[i];                                                 \
            T sm = u[i] + v[i];                                                 \
            a += fabs((double)df);                                              \
            b += fabs((double)sm);                                              \
        }                                                                       \
        return a / b;                                                           \
    case M_HAMMING:                                                             \
        for (idx_t i = 0; i < n; ++i) a += (u[i] != v[i]);                      \
        return a / (double)n;                                                   \
    case M_JACCARD:                                                             \
        for (idx_t i = 0; i < n; ++i) {                                         \
            int nz = (u[i] != 0) | (v[i] != 0);                                 \
            a += (u[i] != v[i]) & nz;                                           \
            b += nz;                                                            \
        }                                                                       \
        return a / b;                                                           \
    default:                                                                    \
        return NAN;                                                             \
    }                                                                           \
}                                                                               \
                                                                                \
/* dist.hpp: out[i] = metric(X[row(i)], y) */                                   \
int oracle_dist_##SUF(const T *X, const T *y, int metric, idx_t n, idx_t m,     \
                      const idx_t *rows, idx_t n_rows, double *out)             \
{                                                                               \
    if (metric < 0 || metric >= M_COUNT) return -1;                             \
    idx_t count = rows ? n_rows : n;                                            \
    for (idx_t i = 0; i < count; ++i) {                                         \
        idx_t r = rows ? rows[i] : i;                                           \
        out[i] = metric_##SUF(metric, X + r * m, y, m);                         \
    }                                                                           \
    return 0;                                                                   \
}                                                                               \
                                                                                \
/* assign.hpp: argmin over centres with strict '<', inertia = sum of minima */  \
double oracle_assign_nearest_##SUF(const T *X, const T *Y, int metric,          \
                                   const idx_t *rows, idx_t n_X, idx_t n_Y,     \
                                   idx_t m, idx_t n_rows, idx_t *assign)        \
{                                                                               \
    if (metric < 0 || metric >= M_COUNT) return -1.0;                           \
    idx_t count = rows ? n_rows : n_X;                                          \
    double inertia = 0.0;                                                       \
    for (idx_t i = 0; i < count; ++i) {                                         \
        const T *x = X + (rows ? rows[i] : i) * m;                              \
        double best = DBL_MAX;                                                  \
        for (idx_t j = 0; j < n_Y; ++j) {                                       \
            double d = metric_##SUF(metric, x, Y + j * m, m);                   \
            if (d < best) { best = d; assign[i] = j; }                          \
        }                                                                       \
        inertia += best;                                                        \
    }                                                                           \
    return inertia;                                                             \
}                                                                               \
                                                                                \
/* cdist.hpp: row-major (na, nb) */                                             \
int oracle_cdist_##SUF(const T *XA, const T *XB, int metric, idx_t na,          \
                       idx_t nb, idx_t m, double *out)                          \
{                                                                               \
    if (metric < 0 || metric >= M_COUNT) return -1;                             \
    for (idx_t i = 0; i < na; ++i)                                              \
        for (idx_t j = 0; j < nb; ++j)                                          \
            out[i * nb + j] = metric_##SUF(metric, XA + i * m, XB + j * m, m);  \
    return 0;                                                                   \
}                                                                               \
                                                                                \
/* pdist.hpp: condensed upper triangle, optionally over gathered rows */        \
int oracle_pdist_##SUF(const T *X, int metric, idx_t n, idx_t m,                \
                       const idx_t *rows, idx_t n_rows, double *out)            \
{                                                                               \
    if (metric < 0 || metric >= M_COUNT) return -1;                             \
    idx_t count = rows ? n_rows : n, k = 0;                                     \
    for (idx_t a = 0; a < count; ++a) {                                         \
        const T *u = X + (rows ? rows[a] : a) * m;                              \
        for (idx_t b = a + 1; b < count; ++b) {                                 \
            const T *v = X + (rows ? rows[b] : b) * m;                          \
            out[k++] = metric_##SUF(metric, u, v, m);                           \
        }                                                                       \
    }                                                                           \
    return 0;                                                                   \
}                                                                               \
                                                                                \
/* sumdist.hpp: sum over explicit (i, j) pairs */                               \
double oracle_sumdist_##SUF(const T *X, int metric, idx_t n, idx_t m,           \
                            const idx_t *pairs, idx_t p)                        \
{                                                                               \
    (void)n;                                                                    \
    if (metric < 0 || metric >= M_COUNT) return -1.0;                           \
    double s = 0.0;                                                             \
    for (idx_t i = 0; i < p; ++i)                                               \
        s += metric_##SUF(metric, X + m * pairs[2 * i],                         \
                          X + m * pairs[2 * i + 1], m);                         \
    return s;                                                                   \
}

DEFINE_METRIC(float, f32)
DEFINE_METRIC(double, f64)

/* ------------------------------------------------------------------------ */
/* k-medoids on a condensed distance matrix, npass == 0 branch only.        */
/* ------------------------------------------------------------------------ */

/* kmedoids.cc:60-69 -- position of unordered pair (i, j), i != j, in pdist order */
static idx_t condensed_index(idx_t i, idx_t j, idx_t n)
{
    if (i > j) { idx_t t = i; i = j; j = t; }
    return n * i - i * (i + 1) / 2 + j - 1 - i;
}

idx_t oracle_condensed_index(idx_t i, idx_t j, idx_t n)
{
    return condensed_index(i, j, n);
}

/* kmedoids.cc:264-309 -- medoid of each cluster = member with the smallest
 * (early-terminated) sum of distances to the other members; first wins ties. */
static void cluster_medoids(idx_t k, idx_t n, const double *dm,
                            const idx_t *cid, idx_t *medoid, double *cost)
{
    for (idx_t c = 0; c < k; ++c) cost[c] = DBL_MAX;
    for (idx_t i = 0; i < n; ++i) {
        idx_t c = cid[i];
        double d = 0.0;
        for (idx_t o = 0; o < n; ++o) {
            if (o == i || cid[o] != c) continue;
            d += dm[condensed_index(i, o, n)];
            if (d > cost[c]) break;
        }
        if (d < cost[c]) { cost[c] = d; medoid[c] = i; }
    }
}

/*
 * kmedoids.cc:74-260 with npass == 0: alternate medoid update / reassignment
 * starting from the caller's clusterid until the total stops decreasing or a
 * periodically saved assignment recurs.  On return clusterid[i] is the element
 * index of i's medoid.  Returns ifound (1 on success, 0 if k > n, -1 alloc).
 */
int oracle_kmedoids(idx_t k, idx_t n, const double *dm, idx_t *clusterid,
                    double *error)
{
    if (n < k) return 0;
    idx_t *saved = (idx_t *)malloc(sizeof(idx_t) * (size_t)n);
    idx_t *medoid = (idx_t *)malloc(sizeof(idx_t) * (size_t)k);
    idx_t *work = (idx_t *)malloc(sizeof(idx_t) * (size_t)n);
    double *cost = (double *)malloc(sizeof(double) * (size_t)k);
    if (!saved || !medoid || !cost || !work) {
        free(saved); free(medoid); free(cost); free(work);
        return -1;
    }
    /* npass <= 1: the reference iterates in place on clusterid (tclusterid ==
     * clusterid, kmedoids.cc:165-166); the final "differs from input" test at
     * :238-249 therefore compares clusterid with centroids[clusterid]. */
    memcpy(work, clusterid, sizeof(idx_t) * (size_t)n);

    double total = DBL_MAX;
    idx_t counter = 0, period = 10;
    *error = DBL_MAX;
    for (;;) {
        double previous = total;
        total = 0.0;
        if (counter % period == 0) {
            memcpy(saved, work, sizeof(idx_t) * (size_t)n);
            if (period < INT64_MAX / 2) period *= 2;
        }
        ++counter;

        cluster_medoids(k, n, dm, work, medoid, cost);

        for (idx_t i = 0; i < n; ++i) {
            double best = DBL_MAX;
            for (idx_t c = 0; c < k; ++c) {
                idx_t j = medoid[c];
                if (i == j) { best = 0.0; work[i] = c; break; }
                double d = dm[condensed_index(i, j, n)];
                if (d < best) { best = d; work[i] = c; }
            }
            total += best;
        }
        if (total >= previous) break;
        idx_t i = 0;
        while (i < n && saved[i] == work[i]) ++i;
        if (i == n) break;
    }

    int ifound = -1;
    idx_t i = 0;
    for (; i < n; ++i) {
        if (work[i] != medoid[work[i]]) {
            if (total < *error) {
                ifound = 1;
                *error = total;
                for (idx_t j = 0; j < n; ++j) work[j] = medoid[work[j]];
            }
            break;
        }
    }
    if (i == n) ifound += 1;
    memcpy(clusterid, work, sizeof(idx_t) * (size_t)n);
    free(saved); free(medoid); free(cost); free(work);
    return ifound;
}

/*
 * kmedoids.cc:386-401: relabel ids to 0..n_unique-1 in order of first
 * appearance.  keys[r] receives the original id that became label r;
 * returns the number of distinct ids.
 */
idx_t oracle_contigify_ids(idx_t *ids, idx_t length, idx_t *keys)
{
    idx_t n_keys = 0;
    for (idx_t i = 0; i < length; ++i) {
        idx_t r = 0;
        while (r < n_keys && keys[r] != ids[i]) ++r;
        if (r == n_keys) keys[n_keys++] = ids[i];
        ids[i] = r;
    }
    return n_keys;
}
