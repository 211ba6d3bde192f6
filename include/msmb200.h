/*
 * msmb200.h -- C ABI of the B200-native hot path for MSMBuilder pipelines.
 *
 * One shared library (msmbuilder_b200/csrc/libmsmb200.so, sm_100a only) exports
 * exactly these symbols.  Plain pointers and sizes; no torch / numpy / Python
 * types.  Every DEVICE pointer is borrowed for the duration of the call's
 * stream work; `stream` is a cudaStream_t passed as void* (NULL = default
 * stream).  Every function returns 0 on success or a nonzero MSMB200_E_* code
 * and never throws; msmb200_last_error() gives the text.  There is NO CPU
 * fallback anywhere behind this interface.
 *
 * The reference (msmbuilder @ 515fd5c, paths relative to /root/reference) has no
 * FFI for this path: the boundary is its Python estimator protocol plus the
 * Cython module msmbuilder.libdistance.  Each entry point below names the
 * reference function it stands in for; INTEGRATION.md shows the ctypes binding a
 * maintainer would add on the reference side.
 */
#ifndef MSMB200_H
#define MSMB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSMB200_ABI_VERSION 1

/* ---- status codes ------------------------------------------------------- */
enum {
    MSMB200_OK = 0,
    MSMB200_E_INVALID = 1,   /* bad argument (shape, alignment, metric id ...) */
    MSMB200_E_CUDA = 2,      /* a CUDA runtime / driver call failed            */
    MSMB200_E_UNSUPPORTED = 3, /* valid request this build cannot serve        */
    MSMB200_E_NODEVICE = 4   /* no sm_100 device visible                       */
};

/* ---- metric ids: msmbuilder/libdistance/src/distance_kernels.h:245-293 ----
 * (string dispatch there; the Python wrapper raises ValueError for an unknown
 * name exactly like libdistance.pyx:122-124 before the id is formed)          */
enum {
    MSMB200_EUCLIDEAN = 0, MSMB200_SQEUCLIDEAN = 1, MSMB200_CITYBLOCK = 2,
    MSMB200_CHEBYSHEV = 3, MSMB200_CANBERRA = 4, MSMB200_BRAYCURTIS = 5,
    MSMB200_HAMMING = 6, MSMB200_JACCARD = 7,
    MSMB200_N_VECTOR_METRICS = 8
};

/* ---- element types of a frame matrix ------------------------------------- */
enum { MSMB200_F32 = 0, MSMB200_F64 = 1 };

/* ---- library ---------------------------------------------------------- */
int msmb200_abi_version(void);
const char *msmb200_last_error(void);           /* thread-local, never NULL */
/* number of CUDA kernels this library has launched in this process (bench.py's
 * gpu_launches is a difference of two readings) */
uint64_t msmb200_launch_count(void);
/* sm count / compute capability of `device`; E_NODEVICE when none. */
int msmb200_device_info(int device, int *sm_count, int *cc_major, int *cc_minor,
                        size_t *total_mem);

/* ======================================================================== *
 *  K1  tICA sufficient statistics                                          *
 *  replaces tICA._fit, msmbuilder/decomposition/tica.py:401-424            *
 * ======================================================================== *
 * Packed accumulator (DEVICE, float64, length 3*D*D + 3*D + 2):
 *   [ C_tau  = sum_t x_t x_{t+tau}^T      (tica.py:417  _outer_0_to_T_lagged)
 *   | C_00   = sum_{t<n-tau} x_t x_t^T    (tica.py:421  _outer_0_to_TminusTau)
 *   | C_tt   = sum_{t>=tau}  x_t x_t^T    (tica.py:422  _outer_offset_to_T)
 *   | S_0    = sum_{t<n-tau} x_t          (tica.py:418  _sum_0_to_TminusTau)
 *   | S_tau  = sum_{t>=tau}  x_t          (tica.py:419  _sum_tau_to_T)
 *   | S      = sum_t x_t                  (tica.py:420  _sum_0_to_T)
 *   | n_observations | n_sequences ]      (tica.py:414-415)
 * All matrices row-major D x D.  The call ADDS this batch's statistics.
 * Sequences with n_rows <= lag are skipped and not counted (tica.py:410-412).
 */
size_t msmb200_tica_acc_len(int n_features);

/* Precision / engine selector for K1. */
enum {
    MSMB200_TICA_AUTO = 0,      /* tcgen05 3xF16 when the shape allows, else SIMT f64 */
    MSMB200_TICA_SIMT_F64 = 1,  /* CUDA-core float64, any D, any lag, f32 or f64 input  */
    MSMB200_TICA_UMMA_3XTF32 = 2, /* tcgen05.mma kind::tf32, error-compensated 3-term split */
    MSMB200_TICA_UMMA_TF32 = 3, /* tcgen05.mma kind::tf32, single pass                   */
    MSMB200_TICA_UMMA_3XBF16 = 4, /* kind::f16 bf16 h/m split, 3 products (~2^-16, K = 16) */
    MSMB200_TICA_UMMA_6XBF16 = 5, /* bf16 h/m/l split, 6 products (~2^-24)                  */
    MSMB200_TICA_UMMA_3XF16 = 6   /* kind::f16 fp16 h/l split of the per-feature power-of-two
                                     scaled frame, 3 products (~2^-22, K = 16); a value outside
                                     fp16's range reruns the call as 6XBF16 on the stream   */
};

/* Bytes of DEVICE scratch the call may need for (n_features, engine). */
size_t msmb200_tica_workspace_bytes(int n_features, int engine);

/*
 * seq_ptrs / seq_rows are HOST arrays of length n_seq: device base pointer of
 * each sequence (row-major, `ld` elements between rows, 16-byte aligned for the
 * tcgen05 engine) and its frame count.  dtype = MSMB200_F32 | MSMB200_F64.
 */
int msmb200_tica_accumulate(const void *const *seq_ptrs, const int64_t *seq_rows,
                            int n_seq, int n_features, int64_t ld, int dtype,
                            int lag, int engine, double *acc, void *workspace,
                            size_t workspace_bytes, void *stream);

/* tICA.transform, tica.py:330-336: out[n,k] (f64) = (X - means) @ comps^T [* scale].
 * means (D), comps (k x D row-major), scale (k, may be NULL) are DEVICE f64. */
int msmb200_tica_transform(const void *X, int64_t n, int n_features, int64_t ld,
                           int dtype, const double *means, const double *comps,
                           const double *scale, int k, double *out, void *stream);

/* ======================================================================== *
 *  K2  one k-centers pass: distance to one centre + running min + arg-max  *
 *  replaces kcenters.py:92-97  ==  libdistance.dist (dist.hpp:4-60)        *
 *           + the four NumPy passes (mask, two masked stores, argmax)      *
 * ======================================================================== *
 * struct written by the pass (DEVICE): the farthest frame after the update,
 * first index winning ties like np.argmax (kcenters.py:97), and that frame's
 * row so the next pass (or an all-gather across ranks) can start from it.   */
typedef struct msmb200_candidate {
    double  value;      /* max_i distances[i] after this pass                */
    int64_t index;      /* row_offset + argmax (lowest index among ties)     */
    /* followed in memory by `row_elems` elements of the frame (same dtype as X) */
} msmb200_candidate;

size_t msmb200_candidate_bytes(int row_elems, int dtype);
size_t msmb200_kcenters_workspace_bytes(int device);

/*
 * X: n x d frames (dtype), ld elements between rows.
 * center: DEVICE pointer to d elements (same dtype) -- typically the payload of
 *         the previous pass' candidate.
 * distances (f64, n) / labels (i32, n): running minimum and its centre number;
 *         initialise to +inf / 0 (kcenters.py:86-88).  Strict '<' update.
 * out:    DEVICE msmb200_candidate (+ payload) for this shard.
 */
int msmb200_kcenters_pass(const void *X, int64_t n, int d, int64_t ld, int dtype,
                          int metric, const void *center, int32_t center_label,
                          double *distances, int32_t *labels, int64_t row_offset,
                          msmb200_candidate *out, void *workspace,
                          size_t workspace_bytes, void *stream);

/* Pick the winner among `n_cand` gathered candidates (stride `stride_bytes`):
 * max value, then lowest index; copies it (with payload) to `out`.           */
int msmb200_candidate_select(const void *cands, int n_cand, size_t stride_bytes,
                             int row_elems, int dtype, msmb200_candidate *out,
                             void *stream);

/* Fill `out` from an explicit frame index (the seed centre, kcenters.py:84). */
int msmb200_candidate_from_row(const void *X, int64_t row, int d, int64_t ld,
                               int dtype, int64_t row_offset,
                               msmb200_candidate *out, void *stream);

/* ------------------------------------------------------------------------ *
 *  K2b  k-centers with look-ahead (csrc/kcenters_lookahead.cu)
 *  Same results as k calls of msmb200_kcenters_pass (kcenters.py:91-97), in
 *  (number of certified chains + 1) reads of the frames instead of k:
 *    multi_pass  applies the J pending centres (labels label0 .. label0+J-1) in
 *                ONE streaming read -- float32 filter, reference arithmetic for
 *                every (frame, centre) pair that could lower the frame's minimum
 *                -- and records each lane's three largest minima (the two
 *                largest with their rows);
 *    select      turns those into this shard's candidate set: up to t_cap
 *                frames (value, global row, the row itself; a lane contributes
 *                its two largest) and the bound tau on every frame that is NOT
 *                a candidate (largest third-best of a lane, or the cut);
 *    chain       replays the reference's arg-max / update loop on the candidate
 *                sets of all shards (all-gathered by the caller) and emits the
 *                centres it can certify (value > tau; the first pick is the true
 *                arg-max and always certified): the next multi_pass' input.
 *  float32, euclidean / sqeuclidean, 16-byte aligned rows (..._supported()).
 *  Buffers are opaque DEVICE blobs sized by the *_bytes() helpers; a centres
 *  blob is {int32 n; int32 cap; 24 bytes pad; int64 ids[cap]; float rows[cap][d]}.
 * ------------------------------------------------------------------------ */
int msmb200_kcenters_lookahead_supported(int d, int64_t ld, int dtype, int metric);
size_t msmb200_kcenters_lane_bytes(int device);
size_t msmb200_kcenters_set_bytes(int d, int t_cap);
size_t msmb200_kcenters_centers_bytes(int d, int j_cap);
int msmb200_kcenters_multi_pass(const void *X, int64_t n, int d, int64_t ld, int dtype,
                                int metric, const void *centers, int n_centers, int j_cap,
                                int32_t label0, int first, double *distances,
                                int32_t *labels, int64_t row_offset, void *lane_buf,
                                size_t lane_bytes, void *stream);
int msmb200_kcenters_select(const void *X, int64_t n, int d, int64_t ld, int64_t row_offset,
                            const void *lane_buf, int t_cap, void *set_out, void *stream);
int msmb200_kcenters_chain(void *sets, int n_sets, size_t set_stride, int d, int metric,
                           int k_remaining, int j_cap, void *centers_out, void *stream);

/* ======================================================================== *
 *  K3  assign_nearest                                                      *
 *  replaces libdistance.assign_nearest (libdistance.pyx:82-131 ->          *
 *           assign.hpp:6-91)                                               *
 * ======================================================================== *
 * labels (i32, n_out): argmin_j metric(X[i], Y[j]), lowest j on exact ties.
 * min_dist (f64, n_out, may be NULL): the winning distance.
 * inertia (DEVICE f64[1]): sum of the winning distances.
 * rows (DEVICE i64, may be NULL): optional gather X_indices; n_out = n_rows.
 */
size_t msmb200_assign_workspace_bytes(int64_t n_out, int k, int d);
/* which filter engine float32 (sq)euclidean assign_nearest takes for this shape (contiguous,
 * 16-byte aligned frames, no row gather): 0 = tcgen05 with the centres resident in shared
 * memory, 1 = tcgen05 with streamed centre chunks, 2 = SIMT float32 filter.  Labels are the
 * exact engine's in every case (ambiguous frames are re-scanned in float64). */
int msmb200_assign_engine(int64_t n_out, int k, int d);
int msmb200_assign_nearest(const void *X, int64_t n, int d, int64_t ld, int dtype,
                           const void *Y, int k, int metric, const int64_t *rows,
                           int64_t n_rows, int32_t *labels, double *min_dist,
                           double *inertia, void *workspace, size_t workspace_bytes,
                           void *stream);

/* ======================================================================== *
 *  K4  dist / cdist / pdist / sumdist                                      *
 *  replaces libdistance.{dist,cdist,pdist,sumdist}                         *
 *           (dist.hpp, cdist.hpp, pdist.hpp:73-96, sumdist.hpp)            *
 * ======================================================================== */
int msmb200_dist(const void *X, int64_t n, int d, int64_t ld, int dtype,
                 const void *y, int metric, const int64_t *rows, int64_t n_rows,
                 double *out, void *stream);
int msmb200_cdist(const void *XA, int64_t na, const void *XB, int64_t nb, int d,
                  int dtype, int metric, double *out, void *stream);
/* condensed upper triangle in pdist order; rows may be NULL (all of X) */
int msmb200_pdist(const void *X, int64_t n, int d, int64_t ld, int dtype,
                  int metric, const int64_t *rows, int64_t n_rows, double *out,
                  void *stream);
/* pairs: DEVICE i64 (p x 2); out: DEVICE f64[1] */
int msmb200_sumdist(const void *X, int64_t n, int d, int64_t ld, int dtype,
                    int metric, const int64_t *pairs, int64_t p, double *out,
                    void *stream);

/* ======================================================================== *
 *  K5/K6  RMSD metric (QCP) -- replaces the mdtraj calls made from          *
 *  libdistance.pyx:316-370,464-499,527-562 and cluster/base.py:68           *
 * ======================================================================== *
 * xyz: n x n_atoms x 3 float32, C-contiguous.                               */
/* in-place centring + G = sum |r|^2 (float32), one value per frame */
int msmb200_rmsd_center(float *xyz, int64_t n, int n_atoms, float *traces,
                        void *stream);
int msmb200_rmsd_kcenters_pass(const float *xyz, const float *traces, int64_t n,
                               int n_atoms, const float *center /* n_atoms*3 + 1: coords then G */,
                               int32_t center_label, double *distances,
                               int32_t *labels, int64_t row_offset,
                               msmb200_candidate *out, void *workspace,
                               size_t workspace_bytes, void *stream);
/* The same pass, skipping the frames the triangle inequality rules out (RMSD is a metric):
 * a frame whose centre c satisfies d(c, new) >= 2 d(frame, c) + margin cannot move to the new
 * centre, so only its 16 bytes of (distance, label, trace) are read.  centre_slots: candidate
 * slots 0 .. center_label, slot_stride bytes apart (slot j = centre j as written by pass j - 1;
 * payload n_atoms*3 coords + G at byte 16); labels must hold the centre index of every frame
 * (true after pass 0).  Writes exactly what msmb200_rmsd_kcenters_pass writes (kcenters.py:91-97).
 * workspace: msmb200_rmsd_pass_workspace_bytes(n, k) bytes, zeroed once before the first pass. */
int msmb200_rmsd_kcenters_pass_pruned(const float *xyz, const float *traces, int64_t n,
                                      int n_atoms, const void *centre_slots, size_t slot_stride,
                                      int32_t center_label, double *distances, int32_t *labels,
                                      int64_t row_offset, msmb200_candidate *out, void *workspace,
                                      size_t workspace_bytes, void *stream);
size_t msmb200_rmsd_pass_workspace_bytes(int64_t n, int32_t max_centres);
int msmb200_rmsd_assign_nearest(const float *xyz, const float *traces, int64_t n,
                                int n_atoms, const float *Y, const float *Y_traces,
                                int k, const int64_t *rows, int64_t n_rows,
                                int32_t *labels, double *min_dist, double *inertia,
                                void *stream);
int msmb200_rmsd_dist(const float *xyz, const float *traces, int64_t n, int n_atoms,
                      const float *y, float y_trace, const int64_t *rows,
                      int64_t n_rows, double *out, void *stream);
int msmb200_rmsd_pdist(const float *xyz, const float *traces, int64_t n, int n_atoms,
                       const int64_t *rows, int64_t n_rows, double *out, void *stream);

/* ======================================================================== *
 *  Scans that follow the distance kernels (scan_kernels.cu)                *
 * ======================================================================== */
/* RegularSpatial (cluster/regularspatial.py:70-77): smallest index i >= start with
 * values[i] > threshold written to *out_index (device), -1 if none.  `values` is the
 * running minimum msmb200_kcenters_pass keeps, so "all distances to the centres so far
 * exceed d_min" is one comparison per frame. */
int msmb200_first_above(const double *values, int64_t n, int64_t start, double threshold,
                        int64_t *out_index, void *stream);

/* msm/core.py:487-602 `_transition_counts` on device-resident integer labels.
 * label_bytes: 4 (int32, what the assignment kernels write) or 8 (int64; INT64_MIN marks
 * a missing label = the reference's NaN / None, core.py:576-580).
 *  label_range     -> out_min_max[2] (device): min and max label (INT64_MAX, INT64_MIN if none)
 *  label_presence  -> flags[span] (device, uint8): flags[l - lo] = 1 for every label present;
 *                     its non-zero entries in order are np.unique (core.py:544)
 *  transition_counts: counts[from * n_states + to] += 1 for every t with t + lag inside the same
 *                     sequence; states are remap[label - remap_lo] (negative = skip) or the
 *                     label itself when remap == NULL.  (sliding_window=False is the same call
 *                     on strided sequences at lag 1, exactly core.py:540-542.)  seq_offsets:
 *                     n_seq + 1 device int64 row offsets.  counts (device int64, n_states^2)
 *                     is accumulated into. */
int msmb200_label_range(const void *labels, int64_t n, int label_bytes, int64_t *out_min_max,
                        void *stream);
int msmb200_label_presence(const void *labels, int64_t n, int label_bytes, int64_t lo,
                           int64_t span, uint8_t *flags, void *stream);
int msmb200_transition_counts(const void *labels, int label_bytes, const int64_t *seq_offsets,
                              int64_t n_seq, int64_t n_total, int64_t lag,
                              const int32_t *remap, int64_t remap_lo, int64_t remap_len,
                              int32_t n_states, int64_t *counts, void *stream);

/* LandmarkAgglomerative.predict, cluster/agglomerative.py:234-269: dists is the (n, n_landmarks)
 * float64 row-major output of msmb200_cdist against the landmarks, the landmarks ordered by
 * cluster (group_offsets: n_clusters + 1 device int32 column offsets).  pool: 0 average, 1 complete,
 * 2 single, 3 ward (POOLING_FUNCTIONS, agglomerative.py:31-43; ward needs the device float64
 * arrays cardinality[n_clusters] and sqsum[n_clusters] = squared_distances_within_cluster_).
 * labels (int32, n): arg-min over clusters of the pooled distance, first cluster wins ties, empty
 * clusters skipped; pooled (f64, n, may be NULL): that minimum; any_negative (device int32, may
 * be NULL) is OR-ed with 1 when a pooled distance was negative (the reference warns). */
int msmb200_pooled_assign(const double *dists, int64_t n, int n_landmarks,
                          const int32_t *group_offsets, int n_clusters, int pool,
                          const double *cardinality, const double *sqsum, int32_t *labels,
                          double *pooled, int32_t *any_negative, void *stream);

/* ======================================================================== *
 *  Host k-medoids on a condensed distance matrix (npass == 0 branch)       *
 *  replaces _kmedoids.kmedoids / contigify_ids                             *
 *           (cluster/_kmedoids.pyx:23-117 -> cluster/src/kmedoids.cc)      *
 *  HOST pointers.  Sequential, data-dependent, m = k + batch points: stays *
 *  on the host by design.                           *
 * ======================================================================== */
int msmb200_kmedoids(int64_t n_clusters, int64_t n_elements, const double *distmatrix,
                     int64_t *clusterid /* in: labels, out: medoid element ids */,
                     double *error, int64_t *ifound);
/* n_pass >= 2 random restarts (kmedoids.cc:160-250 as called by cluster/kmedoids.py:92-94).
 * starts: (n_pass, n_elements) initial labels in [0, n_clusters), drawn by the caller with
 * the RandomState calls of kmedoids.cc:314-383; clusterid in: the solution to beat
 * (zeros from _kmedoids.pyx:97), out: medoid element ids of the best pass. */
int msmb200_kmedoids_restarts(int64_t n_clusters, int64_t n_elements, const double *distmatrix,
                              int64_t n_pass, const int64_t *starts, int64_t *clusterid,
                              double *error, int64_t *ifound);
/* ids relabelled in place in order of first appearance; keys[r] = old id of r */
int msmb200_contigify_ids(int64_t *ids, int64_t length, int64_t *keys,
                          int64_t *n_keys);

#ifdef __cplusplus
}
#endif
#endif /* MSMB200_H */
