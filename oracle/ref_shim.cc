/*
 * oracle/ref_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Thin extern "C" doorway onto the UNMODIFIED reference sources.  Nothing from
 * /root/reference is copied into this repository: the reference headers and
 * kmedoids.cc are #included from where they lie (the include roots are given on
 * the command line by oracle/build_oracle.py) and the resulting object code
 * lands only in oracle/_ref/libref.so (git-ignored, shipped to the GPU box).
 *
 * Used (a) to pin oracle/libdistance_oracle.c bit-for-bit, (b) as the
 * "reference" CPU baseline in bench.py (cpu_baseline.kind == "reference").
 *
 * Wrapped reference entry points:
 *   msmbuilder/libdistance/src/dist.hpp:4-80     dist_{double,float}[_X_indices]
 *   msmbuilder/libdistance/src/assign.hpp:6-91   assign_nearest_{double,float}
 *   msmbuilder/libdistance/src/cdist.hpp:4-50    cdist_{double,float}
 *   msmbuilder/libdistance/src/pdist.hpp:4-96    pdist_{double,float}[_X_indices]
 *   msmbuilder/libdistance/src/sumdist.hpp:3-46  sumdist_{double,float}
 *   msmbuilder/cluster/src/kmedoids.cc:74-260    kmedoids (npass==0 => random unused)
 *   msmbuilder/cluster/src/kmedoids.cc:386-401   contigify_ids
 */
#include <cstdio>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <map>
#include <Python.h>
#include <numpy/npy_common.h>

/* the libdistance headers, verbatim (paths resolved through -I) */
#include "assign.hpp"
#include "dist.hpp"
#include "cdist.hpp"
#include "pdist.hpp"
#include "sumdist.hpp"

/* kmedoids.cc verbatim; its one py2-only symbol is mapped by
 * -DPyInt_AsLong=PyLong_AsLong on the command line. */
/* kmedoids.cc:47-58 `initialize_numpy` is declared `static int` under Python 3 but has
 * no return statement; g++ >= 8 treats falling off its end as unreachable and the call
 * crashes.  Without touching the source: pull in the NumPy header first (its include guard
 * makes the later #include a no-op) and let the `import_array()` statement itself return. */
#include <numpy/arrayobject.h>
#undef import_array
#define import_array() do { return _import_array() < 0 ? 0 : 1; } while (0)
#include "kmedoids.cc"

extern "C" {

void ref_dist_f32(const float *X, const float *y, const char *metric, int64_t n,
                  int64_t m, const int64_t *rows, int64_t n_rows, double *out)
{
    if (rows) dist_float_X_indices(X, y, metric, n, m, (const npy_intp *)rows, n_rows, out);
    else dist_float(X, y, metric, n, m, out);
}

void ref_dist_f64(const double *X, const double *y, const char *metric, int64_t n,
                  int64_t m, const int64_t *rows, int64_t n_rows, double *out)
{
    if (rows) dist_double_X_indices(X, y, metric, n, m, (const npy_intp *)rows, n_rows, out);
    else dist_double(X, y, metric, n, m, out);
}

double ref_assign_nearest_f32(const float *X, const float *Y, const char *metric,
                              const int64_t *rows, int64_t n_X, int64_t n_Y,
                              int64_t m, int64_t n_rows, int64_t *assign)
{
    return assign_nearest_float(X, Y, metric, (const npy_intp *)rows, n_X, n_Y, m,
                                n_rows, (npy_intp *)assign);
}

double ref_assign_nearest_f64(const double *X, const double *Y, const char *metric,
                              const int64_t *rows, int64_t n_X, int64_t n_Y,
                              int64_t m, int64_t n_rows, int64_t *assign)
{
    return assign_nearest_double(X, Y, metric, (const npy_intp *)rows, n_X, n_Y, m,
                                 n_rows, (npy_intp *)assign);
}

void ref_cdist_f32(const float *A, const float *B, const char *metric, int64_t na,
                   int64_t nb, int64_t m, double *out)
{
    cdist_float(A, B, metric, na, nb, m, out);
}

void ref_cdist_f64(const double *A, const double *B, const char *metric, int64_t na,
                   int64_t nb, int64_t m, double *out)
{
    cdist_double(A, B, metric, na, nb, m, out);
}

void ref_pdist_f32(const float *X, const char *metric, int64_t n, int64_t m,
                   const int64_t *rows, int64_t n_rows, double *out)
{
    if (rows) pdist_float_X_indices(X, metric, n, m, (const npy_intp *)rows, n_rows, out);
    else pdist_float(X, metric, n, m, out);
}

void ref_pdist_f64(const double *X, const char *metric, int64_t n, int64_t m,
                   const int64_t *rows, int64_t n_rows, double *out)
{
    if (rows) pdist_double_X_indices(X, metric, n, m, (const npy_intp *)rows, n_rows, out);
    else pdist_double(X, metric, n, m, out);
}

double ref_sumdist_f32(const float *X, const char *metric, int64_t n, int64_t m,
                       const int64_t *pairs, int64_t p)
{
    return sumdist_float(X, metric, n, m, (const npy_intp *)pairs, p);
}

double ref_sumdist_f64(const double *X, const char *metric, int64_t n, int64_t m,
                       const int64_t *pairs, int64_t p)
{
    return sumdist_double(X, metric, n, m, (const npy_intp *)pairs, p);
}

/* npass is fixed at 0 (what MiniBatchKMedoids passes, minibatchkmedoids.py:116-118),
 * so the RandomState argument is never dereferenced. */
int64_t ref_kmedoids_npass0(int64_t k, int64_t n, double *dm, int64_t *clusterid,
                            double *error)
{
    npy_intp ifound = 0;
    kmedoids(k, n, dm, 0, (npy_intp *)clusterid, NULL, error, &ifound);
    return ifound;
}

/* any npass: `random` is a borrowed numpy RandomState whose binomial/shuffle methods
 * kmedoids.cc:314-383 calls.  Must be entered with the GIL held (ctypes.PyDLL). */
int64_t ref_kmedoids_npass(int64_t k, int64_t n, double *dm, int64_t npass,
                           int64_t *clusterid, PyObject *random, double *error)
{
    npy_intp ifound = 0;
    kmedoids(k, n, dm, npass, (npy_intp *)clusterid, random, error, &ifound);
    return ifound;
}

int64_t ref_contigify_ids(int64_t *ids, int64_t length, int64_t *keys)
{
    std::map<npy_intp, npy_intp> m = contigify_ids((npy_intp *)ids, length);
    for (auto &kv : m) keys[kv.second] = kv.first;
    return (int64_t)m.size();
}

} /* extern "C" */
