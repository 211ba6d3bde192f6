// dist_kernels.cu -- vector-metric distance kernels of libmsmb200 (sm_100a).
//
//   K2  kcenters_pass      one Gonzalez pass: d(x_i, c), strict running min,
//                          label update and the arg-max that names the next
//                          centre, fused in ONE streaming read of X.
//                          (replaces kcenters.py:92-97 + dist.hpp:44-60)
//   K3  assign_nearest     arg-min over k centres, lowest index on ties
//                          (replaces assign.hpp:50-91)
//   K4  dist/cdist/pdist/sumdist (dist.hpp, cdist.hpp, pdist.hpp, sumdist.hpp)
//
// All of these are HBM-bound integer/byte-style streaming work (4*D + 8 bytes
// per frame per pass): coalesced 16-byte loads, a sub-warp of G lanes per frame,
// warp-shuffle reductions, grid sized as a multiple of the SM count.  No tensor
// cores here by design.
#include "common.cuh"

namespace msmb {
// K3 filter on the tensor cores (assign_umma.cu)
bool assign_umma_supported(int64_t n_out, int d, int64_t ld, int k, const void *X, bool has_rows);
int assign_umma_mode(int d, int k);
int assign_umma_filter(const float *X, int64_t n, int d, int64_t ld, const float *Y, int k,
                       int32_t *labels, int *amb_list, int *amb_count, cudaStream_t st);
}
namespace msmb {

static constexpr int kThreads = 256;

// A sub-warp "group" of G lanes (G = power of two <= 32) owns one frame / pair.
// The loops below are written so that every lane of a warp executes the same
// number of iterations (full-mask shuffles inside group_distance).
#define MSMB_GROUP_SETUP()                                                             \
    const int lane_in_group = threadIdx.x & (G - 1);                                   \
    const int groups_per_warp = 32 / G;                                                \
    const int group_in_warp = (threadIdx.x & 31) / G;                                  \
    const long long warp_group0 =                                                      \
        (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * groups_per_warp;   \
    const long long n_groups = ((long long)gridDim.x * blockDim.x) / G

struct BlockCand {
    double v;
    long long i;
};

// Workspace layout for the pass: [unsigned counter | pad to 16 | BlockCand[grid]]
static inline int pass_grid() { return sm_count() * 8; }


// Block arg-max, per-block candidate, last-block-done reduction; the winner (value,
// global index, frame) is published in `out`.  First index wins ties == np.argmax
// (kcenters.py:97).  Must be reached by every thread of the block.
template <typename T>
__device__ __forceinline__ void pass_publish(ArgMax best, const T *__restrict__ X, long long n,
                                             int d, long long ld, long long row_offset,
                                             BlockCand *__restrict__ block_cands,
                                             unsigned *__restrict__ counter,
                                             msmb200_candidate *__restrict__ out)
{
    __shared__ ArgMax s_warp[kThreads / 32];
    __shared__ bool s_is_last;
    best = argmax_warp(best);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        ArgMax b = (threadIdx.x < kThreads / 32) ? s_warp[threadIdx.x]
                                                  : ArgMax{-INFINITY, 0x7fffffffffffffffLL};
        b = argmax_warp(b);
        if (threadIdx.x == 0) {
            block_cands[blockIdx.x].v = b.v;
            block_cands[blockIdx.x].i = b.i;
            __threadfence();
            unsigned ticket = atomicInc(counter, gridDim.x - 1);   // wraps to 0: self-resetting
            s_is_last = (ticket == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_is_last) return;

    // last block: reduce the per-block candidates, publish winner + its row
    __threadfence();
    ArgMax w{-INFINITY, 0x7fffffffffffffffLL};
    for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
        ArgMax c;
        c.v = __ldcg(&block_cands[b].v);
        c.i = __ldcg(&block_cands[b].i);
        w = argmax_merge(w, c);
    }
    w = argmax_warp(w);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x < 32) {
        ArgMax b = (threadIdx.x < kThreads / 32) ? s_warp[threadIdx.x]
                                                  : ArgMax{-INFINITY, 0x7fffffffffffffffLL};
        b = argmax_warp(b);
        if (threadIdx.x == 0) s_warp[0] = b;
    }
    __syncthreads();
    w = s_warp[0];
    if (w.i == 0x7fffffffffffffffLL) w.i = 0;   // n == 0 or all-NaN shard
    if (threadIdx.x == 0) {
        out->value = w.v;
        out->index = row_offset + w.i;
    }
    T *payload = reinterpret_cast<T *>(out + 1);
    if (n > 0)
        for (int j = threadIdx.x; j < d; j += blockDim.x) payload[j] = X[w.i * ld + j];
}


// ---------------------------------------------------------------------------
// K2
// ---------------------------------------------------------------------------
template <typename T, int METRIC, bool VEC>
__global__ void __launch_bounds__(kThreads)
kcenters_pass_kernel(const T *__restrict__ X, long long n, int d, long long ld,
                     const T *__restrict__ center, int label,
                     double *__restrict__ dist, int *__restrict__ labels,
                     long long row_offset, BlockCand *__restrict__ block_cands,
                     unsigned *__restrict__ counter, msmb200_candidate *__restrict__ out,
                     int G)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_center = reinterpret_cast<T *>(smem_raw);
    for (int j = threadIdx.x; j < d; j += blockDim.x) s_center[j] = center[j];
    __syncthreads();

    MSMB_GROUP_SETUP();

    ArgMax best;
    best.v = -INFINITY;
    best.i = 0x7fffffffffffffffLL;

    // warp-uniform trip count: every lane joins every shuffle; out-of-range
    // groups recompute the last row and discard it.
    for (long long r0 = warp_group0; r0 < n; r0 += n_groups) {
        const long long r_raw = r0 + group_in_warp;
        const bool valid = r_raw < n;
        const long long r = valid ? r_raw : n - 1;
        double dv = group_distance<METRIC, T, VEC>(X + r * ld, s_center, d, lane_in_group, G);
        if (valid && lane_in_group == 0) {
            double cur = dist[r];
            if (dv < cur) {          // strict: kcenters.py:93
                cur = dv;
                dist[r] = dv;
                labels[r] = label;
            }
            if (cur > best.v) {      // rows visited in increasing order per thread
                best.v = cur;
                best.i = r;
            }
        }
    }

    pass_publish<T>(best, X, n, d, ld, row_offset, block_cands, counter, out);
}

// ---------------------------------------------------------------------------
// K2 fast path (float32 frames, 16-byte vector loads, d/4 == ITERS * G):
// each sub-warp group streams R frames per iteration with all R*ITERS 16-byte
// loads issued before the first use (bytes in flight per warp: R * 4 * d), the
// centre lives in registers, the R float64 group sums share one split reduction
// (group_reduce_split: 6 instead of 20 shuffle rounds at R = 4, G = 32), each frame's
// running-minimum update belongs to one lane (its distances[] value is prefetched
// before the arithmetic) and only that lane's sqrt is taken.  Same arithmetic,
// bit for bit, as the generic kernel.
// ---------------------------------------------------------------------------
template <int METRIC, int ITERS, int R>
__global__ void __launch_bounds__(kThreads)
kcenters_pass_fast_kernel(const float *__restrict__ X, long long n, int d, long long ld,
                          const float *__restrict__ center, int label,
                          double *__restrict__ dist, int *__restrict__ labels,
                          long long row_offset, BlockCand *__restrict__ block_cands,
                          unsigned *__restrict__ counter, msmb200_candidate *__restrict__ out,
                          int G)
{
    typedef Metric<METRIC, float> M;
    const int lane_in_group = threadIdx.x & (G - 1);
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long NG = ((long long)gridDim.x * blockDim.x) / G;
    const long long warp_gid0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / G);
    const long long ld4 = ld >> 2;
    const float4 *X4 = reinterpret_cast<const float4 *>(X);

    float4 c[ITERS];
#pragma unroll
    for (int i = 0; i < ITERS; ++i)
        c[i] = reinterpret_cast<const float4 *>(center)[lane_in_group + i * G];

    ArgMax best{-INFINITY, 0x7fffffffffffffffLL};

    // frame of an iteration whose sum ends up in this lane (group_reduce_split), and whether
    // this lane is the one of the G/R holders that does the running-minimum update
    int jsel = 0;
    {
        int off = G >> 1;
#pragma unroll
        for (int m = R; m > 1; m >>= 1, off >>= 1)
            if (lane_in_group & off) jsel += m >> 1;
    }
    const bool owner = (lane_in_group & (G / R - 1)) == 0;

    // one iteration = R frames per group; FULL iterations (every row of every group of the
    // warp in range) skip the row clamps and validity tests
    const long long stride_j = NG * ld4;                        // float4 units from frame j to j + 1
    const float4 *pfull = X4 + gid * ld4 + lane_in_group;       // frame 0 of the current full iteration
    auto iteration = [&](long long it, auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        long long rr[R];
        float4 x[R][ITERS];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            rr[j] = (it * R + j) * NG + gid;
            const long long rc = (FULL || rr[j] < n) ? rr[j] : n - 1;
            // full iterations walk a running pointer (adds only, no 64-bit multiplies)
            const float4 *p = FULL ? pfull + j * stride_j : X4 + rc * ld4 + lane_in_group;
#pragma unroll
            for (int i = 0; i < ITERS; ++i) x[j][i] = ldg_stream(p + i * G);
        }
        long long myrow = -1;
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (owner && jsel == j && (FULL || rr[j] < n)) myrow = rr[j];
        double cur = INFINITY;
        if (myrow >= 0) cur = __ldcg(dist + myrow);

        double va[R], vb[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                M::acc(a, b, x[j][i].x, c[i].x);
                M::acc(a, b, x[j][i].y, c[i].y);
                M::acc(a, b, x[j][i].z, c[i].z);
                M::acc(a, b, x[j][i].w, c[i].w);
            }
            va[j] = a;
            vb[j] = b;
        }
        // R group sums for the price of ~one: the lanes split the frames while they reduce
        // (same addition tree as group_combine, so the values are bit-identical to it), then
        // ONE finishing step (the sqrt) per lane instead of R
        const double ra = group_reduce_split<M::kIsMax, R>(va, G, lane_in_group);
        const double rb = M::kTwoAcc ? group_reduce_split<false, R>(vb, G, lane_in_group) : 0.0;
        const double mine = M::fin(ra, rb, d);
        if (myrow >= 0) {
            if (mine < cur) {            // strict: kcenters.py:93
                cur = mine;
                dist[myrow] = mine;
                labels[myrow] = label;
            }
            if (cur > best.v) {          // a lane's rows increase with `it`
                best.v = cur;
                best.i = myrow;
            }
        }
    };
    // rows of iteration `it` span [(it*R)*NG + warp_gid0, (it*R + R-1)*NG + warp_gid0 + 32/G)
    long long it = 0;
    for (; (it * R + R - 1) * NG + warp_gid0 + (32 / G) <= n; ++it) {
        iteration(it, std::true_type());
        pfull += R * stride_j;
    }
    for (; (it * R) * NG + warp_gid0 < n; ++it)
        iteration(it, std::false_type());
    pass_publish<float>(best, X, n, d, ld, row_offset, block_cands, counter, out);
}

template <typename T>
__global__ void candidate_select_kernel(const unsigned char *__restrict__ cands, int n_cand,
                                        size_t stride, int row_elems,
                                        msmb200_candidate *__restrict__ out)
{
    __shared__ int s_win;
    if (threadIdx.x == 0) {
        int win = 0;
        const msmb200_candidate *c0 = reinterpret_cast<const msmb200_candidate *>(cands);
        double bv = c0->value;
        long long bi = c0->index;
        for (int r = 1; r < n_cand; ++r) {
            const msmb200_candidate *c =
                reinterpret_cast<const msmb200_candidate *>(cands + (size_t)r * stride);
            if (c->value > bv || (c->value == bv && c->index < bi)) {
                bv = c->value;
                bi = c->index;
                win = r;
            }
        }
        s_win = win;
        out->value = bv;
        out->index = bi;
    }
    __syncthreads();
    const T *src = reinterpret_cast<const T *>(cands + (size_t)s_win * stride + sizeof(msmb200_candidate));
    T *dst = reinterpret_cast<T *>(out + 1);
    for (int j = threadIdx.x; j < row_elems; j += blockDim.x) dst[j] = src[j];
}

template <typename T>
__global__ void candidate_from_row_kernel(const T *__restrict__ X, long long row, int d,
                                          long long ld, long long row_offset,
                                          msmb200_candidate *__restrict__ out)
{
    if (threadIdx.x == 0) {
        out->value = INFINITY;
        out->index = row_offset + row;
    }
    T *dst = reinterpret_cast<T *>(out + 1);
    for (int j = threadIdx.x; j < d; j += blockDim.x) dst[j] = X[row * ld + j];
}

// ---------------------------------------------------------------------------
// K3 (exact engine): one sub-warp per frame scans all k centres in float64,
// exactly the reference arithmetic; strict '<' keeps the lowest centre index.
// ---------------------------------------------------------------------------
template <typename T, int METRIC, bool VEC>
__global__ void __launch_bounds__(kThreads)
assign_exact_kernel(const T *__restrict__ X, long long n_out, int d, long long ld,
                    const T *__restrict__ Y, int k, const long long *__restrict__ rows,
                    int *__restrict__ labels, double *__restrict__ min_dist,
                    double *__restrict__ block_sums, int G)
{
    MSMB_GROUP_SETUP();
    double local = 0.0;
    for (long long i0 = warp_group0; i0 < n_out; i0 += n_groups) {
        const long long i_raw = i0 + group_in_warp;
        const bool valid = i_raw < n_out;
        const long long i = valid ? i_raw : n_out - 1;
        const long long r = rows ? rows[i] : i;
        const T *u = X + r * ld;
        double best = 1.7976931348623157e308;   // DBL_MAX, assign.hpp:66
        int arg = 0;
        for (int j = 0; j < k; ++j) {
            double dv = group_distance<METRIC, T, VEC, false>(u, Y + (long long)j * d, d,
                                                              lane_in_group, G);
            if (dv < best) {
                best = dv;
                arg = j;
            }
        }
        if (valid && lane_in_group == 0) {
            labels[i] = arg;
            if (min_dist) min_dist[i] = best;
            local += best;
        }
    }
    // deterministic inertia: fixed-order block partials, summed by a second kernel
    __shared__ double s_sum[kThreads / 32];
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) s += s_sum[w];
        block_sums[blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------
// K3 fast engine (float32 frames, euclidean / sqeuclidean): filter + refine.
//   1. assign_filter_kernel: register-tiled (8 frames x 4 centres per thread) squared
//      distances with the REFERENCE's float32 difference but float32 accumulation;
//      tracks best and second-best.  |float32 sum - exact sum| <= ~(d+2) 2^-24 relative,
//      so whenever second > best * (1 + margin) the float32 arg-min IS the exact arg-min.
//   2. the (rare) frames inside the margin -- including exact ties -- are re-scanned by
//      assign_exact_kernel's float64 arithmetic (lowest index wins, assign.hpp:69).
//   3. assign_mindist_kernel recomputes the winning distance in float64 exactly like
//      distance_kernels.h:54-65 for min_dist / inertia.
// Labels are therefore identical to the exact engine; FP32-pipe bound (3 k d flop/frame).
// ---------------------------------------------------------------------------
constexpr int AF_TR = 128, AF_TC = 64, AF_DK = 16;

__global__ void __launch_bounds__(256)
assign_filter_kernel(const float *__restrict__ X, long long n_out, int d, long long ld,
                     const float *__restrict__ Y, int k, const long long *__restrict__ rows,
                     float margin, int *__restrict__ labels, int *__restrict__ amb_list,
                     int *__restrict__ amb_count)
{
    __shared__ __align__(16) float Xs[AF_DK][AF_TR + 4];
    __shared__ __align__(16) float Cs[AF_DK][AF_TC + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * AF_TR;

    float best[8], second[8];
    int arg[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { best[u] = INFINITY; second[u] = INFINITY; arg[u] = 0; }

    for (int c0 = 0; c0 < k; c0 += AF_TC) {
        float acc[8][4];
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int v = 0; v < 4; ++v) acc[u][v] = 0.f;
        for (int d0 = 0; d0 < d; d0 += AF_DK) {
            __syncthreads();
            // stage X tile (AF_TR rows x AF_DK cols) and C tile (AF_TC x AF_DK), transposed
            for (int e = tid; e < AF_TR * AF_DK; e += 256) {
                const int r = e / AF_DK, c = e % AF_DK;
                const long long i = row0 + r;
                float v = 0.f;
                if (i < n_out && d0 + c < d) {
                    const long long src = rows ? rows[i] : i;
                    v = X[src * ld + d0 + c];
                }
                Xs[c][r] = v;
            }
            for (int e = tid; e < AF_TC * AF_DK; e += 256) {
                const int r = e / AF_DK, c = e % AF_DK;
                float v = 0.f;
                if (c0 + r < k && d0 + c < d) v = Y[(long long)(c0 + r) * d + d0 + c];
                Cs[c][r] = v;
            }
            __syncthreads();
#pragma unroll
            for (int kk = 0; kk < AF_DK; ++kk) {
                const float4 xa = *reinterpret_cast<const float4 *>(&Xs[kk][ty * 8]);
                const float4 xb = *reinterpret_cast<const float4 *>(&Xs[kk][ty * 8 + 4]);
                const float4 cc = *reinterpret_cast<const float4 *>(&Cs[kk][tx * 4]);
                const float xv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
                const float cv[4] = {cc.x, cc.y, cc.z, cc.w};
#pragma unroll
                for (int u = 0; u < 8; ++u)
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float df = xv[u] - cv[v];
                        acc[u][v] = fmaf(df, df, acc[u][v]);
                    }
            }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int j = c0 + tx * 4 + v;
            if (j < k) {
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float a = acc[u][v];
                    if (a < best[u]) { second[u] = best[u]; best[u] = a; arg[u] = j; }
                    else if (a < second[u]) second[u] = a;
                }
            }
        }
    }
    // merge the 16 tx lanes that share a frame (lanes of a half-warp)
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int off = 8; off > 0; off >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best[u], off);
            const float os = __shfl_xor_sync(0xffffffffu, second[u], off);
            const int oa = __shfl_xor_sync(0xffffffffu, arg[u], off);
            // second = smallest of {larger of the two bests, both seconds}
            const float hi = fmaxf(best[u], ob);
            float ns = fminf(fminf(second[u], os), hi);
            if (ob < best[u] || (ob == best[u] && oa < arg[u])) { best[u] = ob; arg[u] = oa; }
            second[u] = ns;
        }
        const long long i = row0 + ty * 8 + u;
        if (tx == 0 && i < n_out) {
            labels[i] = arg[u];
            // inside the rounding margin (or an exact tie, or NaN): exact re-scan needed
            if (!(second[u] > best[u] * (1.f + margin) + 1e-37f)) {
                const int slot = atomicAdd(amb_count, 1);
                amb_list[slot] = (int)i;
            }
        }
    }
}

// exact float64 re-scan of the frames listed in amb_list (one sub-warp group per frame)
__global__ void __launch_bounds__(kThreads)
assign_refine_kernel(const float *__restrict__ X, int d, long long ld, const float *__restrict__ Y,
                     int k, const long long *__restrict__ rows, const int *__restrict__ amb_list,
                     const int *__restrict__ amb_count, int *__restrict__ labels, int G, int sq)
{
    const int n_amb = *amb_count;
    MSMB_GROUP_SETUP();
    for (long long a0 = warp_group0; a0 < n_amb; a0 += n_groups) {
        const long long a_raw = a0 + group_in_warp;
        const bool valid = a_raw < n_amb;
        const long long a = valid ? a_raw : n_amb - 1;
        const long long i = amb_list[a];
        const long long r = rows ? rows[i] : i;
        const float *u = X + r * ld;
        double bestd = 1.7976931348623157e308;
        int arg = 0;
        for (int j = 0; j < k; ++j) {
            const double dv = sq
                ? group_distance<MSMB200_SQEUCLIDEAN, float, false, false>(u, Y + (long long)j * d, d, lane_in_group, G)
                : group_distance<MSMB200_EUCLIDEAN, float, false, false>(u, Y + (long long)j * d, d, lane_in_group, G);
            if (dv < bestd) { bestd = dv; arg = j; }
        }
        if (valid && lane_in_group == 0) labels[i] = arg;
    }
}

// The same re-scan for many centres (k >= 64): one WARP per listed frame.  A float32 scan first (lane =
// centre, no reductions: 8 instead of 40 warp instructions per (frame, centre)), its values kept in
// shared memory; only the centres within the float32 error bound of the smallest value are then
// recomputed in float64 -- by the same group_distance call as above, in ascending centre order with
// strict '<' -- so the label is the one assign_refine_kernel finds.  With k = 2000 centres and ~1 % of
// ambiguous frames the full float64 re-scan took 8.4 of the 26 ms of assign_nearest
// (profiles/r2f_launches_assign_stream.csv).
__global__ void __launch_bounds__(kThreads)
assign_refine_wide_kernel(const float *__restrict__ X, int d, long long ld, const float *__restrict__ Y,
                          int k, const long long *__restrict__ rows, const int *__restrict__ amb_list,
                          const int *__restrict__ amb_count, int *__restrict__ labels, int G, int sq,
                          float rel_margin)
{
    extern __shared__ float s_ref[];                    // per warp: x[d] | s[k]
    const int n_amb = *amb_count;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_per_block = blockDim.x >> 5;
    float *sx = s_ref + (size_t)warp * (d + k);
    float *sv = sx + d;
    const int lig = lane & (G - 1);
    // 16-byte loads when every centre row and every warp's staging row is 16-byte aligned
    const bool vec4 = (d & 3) == 0 && (k & 3) == 0 && (reinterpret_cast<uintptr_t>(Y) & 15u) == 0;
    for (long long a = (long long)blockIdx.x * warps_per_block + warp; a < n_amb;
         a += (long long)gridDim.x * warps_per_block) {
        const long long i = amb_list[a];
        const long long r = rows ? rows[i] : i;
        const float *u = X + r * ld;
        __syncwarp();
        for (int e = lane; e < d; e += 32) sx[e] = u[e];
        __syncwarp();
        // float32 scan: lane takes centres lane, lane + 32, ...
        float m = INFINITY;
        bool odd = false;                               // a NaN somewhere: leave the frame to the full scan
        for (int j = lane; j < k; j += 32) {
            const float *c = Y + (long long)j * d;
            float acc = 0.f;
            if (vec4) {
                const float4 *c4 = reinterpret_cast<const float4 *>(c);
                const float4 *x4 = reinterpret_cast<const float4 *>(sx);
                for (int e = 0; e < (d >> 2); ++e) {
                    const float4 cv = __ldg(c4 + e), xv = x4[e];
                    float t = xv.x - cv.x; acc = fmaf(t, t, acc);
                    t = xv.y - cv.y; acc = fmaf(t, t, acc);
                    t = xv.z - cv.z; acc = fmaf(t, t, acc);
                    t = xv.w - cv.w; acc = fmaf(t, t, acc);
                }
            } else {
                for (int e = 0; e < d; ++e) {
                    const float t = sx[e] - __ldg(c + e);
                    acc = fmaf(t, t, acc);
                }
            }
            sv[j] = acc;
            if (acc != acc) odd = true;
            m = fminf(m, acc);
        }
        for (int off = 16; off > 0; off >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, off));
        odd = __any_sync(0xffffffffu, odd) || !(m < INFINITY);
        __syncwarp();
        const float thr = fmaxf(m * (1.f + rel_margin), 1e-30f);
        double bestd = 1.7976931348623157e308;
        int arg = 0;
        for (int j0 = 0; j0 < k; j0 += 32) {
            const int j = j0 + lane;
            const bool cand = j < k && (odd || sv[j] <= thr);
            unsigned mask = __ballot_sync(0xffffffffu, cand);
            while (mask) {
                const int jj = j0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const double dv = sq
                    ? group_distance<MSMB200_SQEUCLIDEAN, float, false, false>(u, Y + (long long)jj * d, d, lig, G)
                    : group_distance<MSMB200_EUCLIDEAN, float, false, false>(u, Y + (long long)jj * d, d, lig, G);
                if (dv < bestd) { bestd = dv; arg = jj; }
            }
        }
        if (lane == 0) labels[i] = arg;
    }
}

// exact float64 distance of every frame to its assigned centre (+ block partial sums)
__global__ void __launch_bounds__(kThreads)
assign_mindist_kernel(const float *__restrict__ X, long long n_out, int d, long long ld,
                      const float *__restrict__ Y, const long long *__restrict__ rows,
                      const int *__restrict__ labels, double *__restrict__ min_dist,
                      double *__restrict__ block_sums, int G, int sq)
{
    MSMB_GROUP_SETUP();
    double local = 0.0;
    for (long long i0 = warp_group0; i0 < n_out; i0 += n_groups) {
        const long long i_raw = i0 + group_in_warp;
        const bool valid = i_raw < n_out;
        const long long i = valid ? i_raw : n_out - 1;
        const long long r = rows ? rows[i] : i;
        const float *c = Y + (long long)labels[i] * d;
        const double dv = sq
            ? group_distance<MSMB200_SQEUCLIDEAN, float, false, true>(X + r * ld, c, d, lane_in_group, G)
            : group_distance<MSMB200_EUCLIDEAN, float, false, true>(X + r * ld, c, d, lane_in_group, G);
        if (valid && lane_in_group == 0) {
            if (min_dist) min_dist[i] = dv;
            local += dv;
        }
    }
    __shared__ double s_sum[kThreads / 32];
    for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
    if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sacc = 0.0;
        for (int w = 0; w < kThreads / 32; ++w) sacc += s_sum[w];
        block_sums[blockIdx.x] = sacc;
    }
}

__global__ void sum_partials_kernel(const double *__restrict__ partials, int n, double *out)
{
    __shared__ double s[32];
    double v = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s[w];
        *out = t;
    }
}

// ---------------------------------------------------------------------------
// K4: dist (one-to-many), pair kernels (cdist / pdist / sumdist)
// ---------------------------------------------------------------------------
template <typename T, int METRIC, bool VEC>
__global__ void __launch_bounds__(kThreads)
dist_kernel(const T *__restrict__ X, long long n_out, int d, long long ld,
            const T *__restrict__ y, const long long *__restrict__ rows,
            double *__restrict__ out, int G)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *s_y = reinterpret_cast<T *>(smem_raw);
    for (int j = threadIdx.x; j < d; j += blockDim.x) s_y[j] = y[j];
    __syncthreads();
    MSMB_GROUP_SETUP();
    for (long long i0 = warp_group0; i0 < n_out; i0 += n_groups) {
        const long long i_raw = i0 + group_in_warp;
        const bool valid = i_raw < n_out;
        const long long i = valid ? i_raw : n_out - 1;
        const long long r = rows ? rows[i] : i;
        double dv = group_distance<METRIC, T, VEC>(X + r * ld, s_y, d, lane_in_group, G);
        if (valid && lane_in_group == 0) out[i] = dv;
    }
}

// MODE 0: cdist  (pair p -> (p / nb, p % nb), A rows from XA, B rows from XB)
// MODE 1: pdist  (pair p -> condensed (i, j), i < j, optional row gather)
// MODE 2: sumdist (explicit pairs; per-block partial sums)
template <typename T, int METRIC, bool VEC, int MODE>
__global__ void __launch_bounds__(kThreads)
pair_kernel(const T *__restrict__ XA, const T *__restrict__ XB, long long n_pairs,
            long long na, long long nb, int d, long long ld,
            const long long *__restrict__ idx, double *__restrict__ out,
            double *__restrict__ block_sums, int G)
{
    MSMB_GROUP_SETUP();
    double local = 0.0;
    for (long long p0 = warp_group0; p0 < n_pairs; p0 += n_groups) {
        const long long p_raw = p0 + group_in_warp;
        const bool valid = p_raw < n_pairs;
        const long long p = valid ? p_raw : n_pairs - 1;
        long long ia, ib;
        if (MODE == 0) {
            ia = p / nb;
            ib = p % nb;
        } else if (MODE == 1) {
            // invert p = m*i - i*(i+1)/2 + (j - i - 1), m = na (pdist.hpp:84-95 order)
            const double m = (double)na;
            long long i = (long long)floor(((2.0 * m - 1.0) - sqrt((2.0 * m - 1.0) * (2.0 * m - 1.0) - 8.0 * (double)p)) * 0.5);
            if (i < 0) i = 0;
            // fix floating-point off-by-one
            while (i > 0 && na * i - i * (i + 1) / 2 > p) --i;
            while (na * (i + 1) - (i + 1) * (i + 2) / 2 <= p) ++i;
            long long j = p - (na * i - i * (i + 1) / 2) + i + 1;
            ia = idx ? idx[i] : i;
            ib = idx ? idx[j] : j;
        } else {
            ia = idx[2 * p];
            ib = idx[2 * p + 1];
        }
        double dv = group_distance<METRIC, T, VEC, false>(XA + ia * ld, XB + ib * ld, d,
                                                          lane_in_group, G);
        if (valid && lane_in_group == 0) {
            if (MODE == 2) local += dv;
            else out[p] = dv;
        }
    }
    if (MODE == 2) {
        __shared__ double s_sum[kThreads / 32];
        for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
        if ((threadIdx.x & 31) == 0) s_sum[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < kThreads / 32; ++w) s += s_sum[w];
            block_sums[blockIdx.x] = s;
        }
    }
}

template <typename T>
static bool vec_ok(const void *a, const void *b, int d, long long ld)
{
    constexpr int N = Vec<T>::N;
    return (d % N == 0) && (ld % N == 0) && aligned16(a) && (b == nullptr || aligned16(b));
}

static inline int grid_for(long long groups_needed, int G)
{
    long long per_block = kThreads / G;
    long long blocks = (groups_needed + per_block - 1) / per_block;
    long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace msmb

using namespace msmb;

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" size_t msmb200_candidate_bytes(int row_elems, int dtype)
{
    size_t es = dtype == MSMB200_F64 ? 8 : 4;
    size_t b = sizeof(msmb200_candidate) + (size_t)row_elems * es;
    return (b + 15) & ~(size_t)15;
}

extern "C" size_t msmb200_kcenters_workspace_bytes(int device)
{
    (void)device;
    // counter (16 B) + one BlockCand per block; generous upper bound on the grid
    return 16 + sizeof(BlockCand) * (size_t)(8 * 256);
}

extern "C" int msmb200_kcenters_pass(const void *X, int64_t n, int d, int64_t ld, int dtype,
                                     int metric, const void *center, int32_t center_label,
                                     double *distances, int32_t *labels, int64_t row_offset,
                                     msmb200_candidate *out, void *workspace,
                                     size_t workspace_bytes, void *stream)
{
    MSMB_REQUIRE(n >= 0 && d > 0 && ld >= d, "kcenters_pass: bad shape n=%lld d=%d ld=%lld",
                 (long long)n, d, (long long)ld);
    MSMB_REQUIRE(X && center && distances && labels && out && workspace,
                 "kcenters_pass: null pointer");
    const int grid_cap = pass_grid();
    MSMB_REQUIRE(workspace_bytes >= 16 + sizeof(BlockCand) * (size_t)grid_cap,
                 "kcenters_pass: workspace too small (%zu)", workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    unsigned *counter = reinterpret_cast<unsigned *>(workspace);
    BlockCand *cands = reinterpret_cast<BlockCand *>(reinterpret_cast<unsigned char *>(workspace) + 16);
    return dispatch(dtype, metric, [&](auto t, auto m) -> int {
        typedef decltype(t) T;
        constexpr int METRIC = decltype(m)::value;
        const bool vec = vec_ok<T>(X, nullptr, d, ld);
        const int G = lanes_per_row(d, Vec<T>::N, vec);
        int grid = grid_for(n, G);
        if (grid > grid_cap) grid = grid_cap;
        // the self-resetting ticket counter needs a stable grid per workspace
        // only within one launch, so any grid size is fine.
        size_t smem = (size_t)d * sizeof(T);
        smem = (smem + 15) & ~(size_t)15;
        if (std::is_same<T, float>::value && vec && G >= 4 && (d / 4) % G == 0 &&
            aligned16(center)) {
            const int iters = (d / 4) / G;
            bool done = true;
#define MSMB_FAST(I, RR)                                                                       \
            kcenters_pass_fast_kernel<METRIC, I, RR><<<grid, kThreads, 0, st>>>(              \
                (const float *)X, n, d, ld, (const float *)center, center_label, distances,    \
                labels, row_offset, cands, counter, out, G)
            if (iters == 1) MSMB_FAST(1, 4);
            else if (iters == 2) MSMB_FAST(2, 4);
            else if (iters == 4) MSMB_FAST(4, 2);
            else done = false;
#undef MSMB_FAST
            if (done) {
                MSMB_LAUNCH_CHECK();
                return MSMB200_OK;
            }
        }
        if (vec)
            kcenters_pass_kernel<T, METRIC, true><<<grid, kThreads, smem, st>>>(
                (const T *)X, n, d, ld, (const T *)center, center_label, distances, labels,
                row_offset, cands, counter, out, G);
        else
            kcenters_pass_kernel<T, METRIC, false><<<grid, kThreads, smem, st>>>(
                (const T *)X, n, d, ld, (const T *)center, center_label, distances, labels,
                row_offset, cands, counter, out, G);
        MSMB_LAUNCH_CHECK();
        return MSMB200_OK;
    });
}

extern "C" int msmb200_candidate_select(const void *cands, int n_cand, size_t stride_bytes,
                                        int row_elems, int dtype, msmb200_candidate *out,
                                        void *stream)
{
    MSMB_REQUIRE(cands && out && n_cand > 0 && row_elems >= 0, "candidate_select: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSMB200_F64)
        candidate_select_kernel<double><<<1, 128, 0, st>>>((const unsigned char *)cands, n_cand,
                                                          stride_bytes, row_elems, out);
    else
        candidate_select_kernel<float><<<1, 128, 0, st>>>((const unsigned char *)cands, n_cand,
                                                         stride_bytes, row_elems, out);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_candidate_from_row(const void *X, int64_t row, int d, int64_t ld,
                                          int dtype, int64_t row_offset,
                                          msmb200_candidate *out, void *stream)
{
    MSMB_REQUIRE(X && out && row >= 0 && d > 0, "candidate_from_row: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == MSMB200_F64)
        candidate_from_row_kernel<double><<<1, 128, 0, st>>>((const double *)X, row, d, ld,
                                                            row_offset, out);
    else
        candidate_from_row_kernel<float><<<1, 128, 0, st>>>((const float *)X, row, d, ld,
                                                           row_offset, out);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

// workspace: [block partial sums (8*256 doubles) | amb_count (int, 256 B) | amb_list (n_out ints)]
extern "C" size_t msmb200_assign_workspace_bytes(int64_t n_out, int k, int d)
{
    (void)k; (void)d;
    return sizeof(double) * (size_t)(8 * 256) + 256 + sizeof(int) * (size_t)(n_out > 0 ? n_out : 0) + 256;
}

extern "C" int msmb200_assign_engine(int64_t n_out, int k, int d)
{
    static const float aligned_dummy[4] __attribute__((aligned(16))) = {0.f, 0.f, 0.f, 0.f};
    if (!assign_umma_supported(n_out, d, d, k, aligned_dummy, false)) return 2;
    return assign_umma_mode(d, k) == 1 ? 1 : 0;
}

extern "C" int msmb200_assign_nearest(const void *X, int64_t n, int d, int64_t ld, int dtype,
                                      const void *Y, int k, int metric, const int64_t *rows,
                                      int64_t n_rows, int32_t *labels, double *min_dist,
                                      double *inertia, void *workspace, size_t workspace_bytes,
                                      void *stream)
{
    MSMB_REQUIRE(d > 0 && k > 0 && ld >= d && n >= 0, "assign_nearest: bad shape");
    MSMB_REQUIRE(X && Y && labels && workspace, "assign_nearest: null pointer");
    const int64_t n_out = rows ? n_rows : n;
    cudaStream_t st = (cudaStream_t)stream;
    double *partials = reinterpret_cast<double *>(workspace);
    if (n_out == 0) {
        if (inertia) MSMB_CUDA(cudaMemsetAsync(inertia, 0, sizeof(double), st));
        return MSMB200_OK;
    }
    // fast engine: float32 euclidean family, at least 2 centres, index fits int32
    const size_t fast_ws = sizeof(double) * (size_t)(8 * 256) + 256 + sizeof(int) * (size_t)n_out;
    if (dtype == MSMB200_F32 && (metric == MSMB200_EUCLIDEAN || metric == MSMB200_SQEUCLIDEAN) &&
        k >= 2 && n_out < 0x7fffffffLL && workspace_bytes >= fast_ws && !getenv("MSMB200_ASSIGN_EXACT")) {
        unsigned char *wsb = reinterpret_cast<unsigned char *>(workspace);
        int *amb_count = reinterpret_cast<int *>(wsb + sizeof(double) * (size_t)(8 * 256));
        int *amb_list = reinterpret_cast<int *>(wsb + sizeof(double) * (size_t)(8 * 256) + 256);
        const int sq = metric == MSMB200_SQEUCLIDEAN;
        const float margin = 4.0f * (float)(d + 4) * 5.9604645e-8f;     // 4 (d+4) 2^-24
        MSMB_CUDA(cudaMemsetAsync(amb_count, 0, sizeof(int), st));
        if (assign_umma_supported(n_out, d, ld, k, X, rows != nullptr)) {
            // tensor-core filter (assign_umma.cu): same contract -- labels + ambiguity list
            const int rc = assign_umma_filter((const float *)X, n_out, d, ld, (const float *)Y, k, labels,
                                              amb_list, amb_count, st);
            if (rc != MSMB200_OK) return rc;
        } else {
            const long long blocks = (n_out + AF_TR - 1) / AF_TR;
            assign_filter_kernel<<<(unsigned)blocks, 256, 0, st>>>(
                (const float *)X, n_out, d, ld, (const float *)Y, k, (const long long *)rows, margin,
                labels, amb_list, amb_count);
            MSMB_LAUNCH_CHECK();
        }
        const int G = lanes_per_row(d, 1, false);
        const size_t wide_smem = sizeof(float) * (size_t)(d + k) * (kThreads / 32);
        if (k >= 64 && wide_smem <= 200 * 1024 && !getenv("MSMB200_ASSIGN_REFINE_FULL")) {
            // float32 pre-scan, float64 only for the centres within its error bound (relative margin of the
            // squared distance: twice the 4 (d + 4) 2^-24 of the SIMT filter, both values carry the error)
            if (wide_smem > 48 * 1024)
                MSMB_CUDA(cudaFuncSetAttribute(assign_refine_wide_kernel,
                                               cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wide_smem));
            assign_refine_wide_kernel<<<sm_count() * 2, kThreads, wide_smem, st>>>(
                (const float *)X, d, ld, (const float *)Y, k, (const long long *)rows, amb_list,
                amb_count, labels, G, sq, 4.0f * margin);
        } else {
            assign_refine_kernel<<<sm_count() * 4, kThreads, 0, st>>>(
                (const float *)X, d, ld, (const float *)Y, k, (const long long *)rows, amb_list,
                amb_count, labels, G, sq);
        }
        MSMB_LAUNCH_CHECK();
        if (min_dist || inertia) {
            const int grid = grid_for(n_out, G);
            assign_mindist_kernel<<<grid, kThreads, 0, st>>>(
                (const float *)X, n_out, d, ld, (const float *)Y, (const long long *)rows, labels,
                min_dist, partials, G, sq);
            MSMB_LAUNCH_CHECK();
            if (inertia) {
                sum_partials_kernel<<<1, 256, 0, st>>>(partials, grid, inertia);
                MSMB_LAUNCH_CHECK();
            }
        }
        return MSMB200_OK;
    }
    return dispatch(dtype, metric, [&](auto t, auto m) -> int {
        typedef decltype(t) T;
        constexpr int METRIC = decltype(m)::value;
        const bool vec = vec_ok<T>(X, Y, d, ld);
        const int G = lanes_per_row(d, Vec<T>::N, vec);
        const int grid = grid_for(n_out, G);
        MSMB_REQUIRE(workspace_bytes >= sizeof(double) * (size_t)grid,
                     "assign_nearest: workspace too small");
        if (vec)
            assign_exact_kernel<T, METRIC, true><<<grid, kThreads, 0, st>>>(
                (const T *)X, n_out, d, ld, (const T *)Y, k, (const long long *)rows, labels,
                min_dist, partials, G);
        else
            assign_exact_kernel<T, METRIC, false><<<grid, kThreads, 0, st>>>(
                (const T *)X, n_out, d, ld, (const T *)Y, k, (const long long *)rows, labels,
                min_dist, partials, G);
        MSMB_LAUNCH_CHECK();
        if (inertia) {
            sum_partials_kernel<<<1, 256, 0, st>>>(partials, grid, inertia);
            MSMB_LAUNCH_CHECK();
        }
        return MSMB200_OK;
    });
}

extern "C" int msmb200_dist(const void *X, int64_t n, int d, int64_t ld, int dtype,
                            const void *y, int metric, const int64_t *rows, int64_t n_rows,
                            double *out, void *stream)
{
    MSMB_REQUIRE(d > 0 && ld >= d && n >= 0 && X && y && out, "dist: bad args");
    const int64_t n_out = rows ? n_rows : n;
    cudaStream_t st = (cudaStream_t)stream;
    return dispatch(dtype, metric, [&](auto t, auto m) -> int {
        typedef decltype(t) T;
        constexpr int METRIC = decltype(m)::value;
        const bool vec = vec_ok<T>(X, nullptr, d, ld);
        const int G = lanes_per_row(d, Vec<T>::N, vec);
        const int grid = grid_for(n_out, G);
        size_t smem = ((size_t)d * sizeof(T) + 15) & ~(size_t)15;
        if (vec)
            dist_kernel<T, METRIC, true><<<grid, kThreads, smem, st>>>(
                (const T *)X, n_out, d, ld, (const T *)y, (const long long *)rows, out, G);
        else
            dist_kernel<T, METRIC, false><<<grid, kThreads, smem, st>>>(
                (const T *)X, n_out, d, ld, (const T *)y, (const long long *)rows, out, G);
        MSMB_LAUNCH_CHECK();
        return MSMB200_OK;
    });
}

template <int MODE>
static int launch_pairs(const void *XA, const void *XB, long long n_pairs, long long na,
                        long long nb, int d, long long ld, int dtype, int metric,
                        const int64_t *idx, double *out, cudaStream_t st)
{
    return dispatch(dtype, metric, [&](auto t, auto m) -> int {
        typedef decltype(t) T;
        constexpr int METRIC = decltype(m)::value;
        const bool vec = vec_ok<T>(XA, XB, d, ld);
        const int G = lanes_per_row(d, Vec<T>::N, vec);
        const int grid = grid_for(n_pairs, G);
        double *partials = nullptr;
        if (MODE == 2) {
            MSMB_CUDA(cudaMallocAsync(&partials, sizeof(double) * grid, st));
        }
        if (vec)
            pair_kernel<T, METRIC, true, MODE><<<grid, kThreads, 0, st>>>(
                (const T *)XA, (const T *)XB, n_pairs, na, nb, d, ld, (const long long *)idx,
                out, partials, G);
        else
            pair_kernel<T, METRIC, false, MODE><<<grid, kThreads, 0, st>>>(
                (const T *)XA, (const T *)XB, n_pairs, na, nb, d, ld, (const long long *)idx,
                out, partials, G);
        MSMB_LAUNCH_CHECK();
        if (MODE == 2) {
            sum_partials_kernel<<<1, 256, 0, st>>>(partials, grid, out);
            MSMB_LAUNCH_CHECK();
            MSMB_CUDA(cudaFreeAsync(partials, st));
        }
        return MSMB200_OK;
    });
}

extern "C" int msmb200_cdist(const void *XA, int64_t na, const void *XB, int64_t nb, int d,
                             int dtype, int metric, double *out, void *stream)
{
    MSMB_REQUIRE(XA && XB && out && d > 0 && na >= 0 && nb >= 0, "cdist: bad args");
    if (na == 0 || nb == 0) return MSMB200_OK;
    return launch_pairs<0>(XA, XB, na * nb, na, nb, d, d, dtype, metric, nullptr, out,
                           (cudaStream_t)stream);
}

extern "C" int msmb200_pdist(const void *X, int64_t n, int d, int64_t ld, int dtype,
                             int metric, const int64_t *rows, int64_t n_rows, double *out,
                             void *stream)
{
    MSMB_REQUIRE(X && out && d > 0 && ld >= d, "pdist: bad args");
    const long long m = rows ? n_rows : n;
    if (m < 2) return MSMB200_OK;
    return launch_pairs<1>(X, X, m * (m - 1) / 2, m, m, d, ld, dtype, metric, rows, out,
                           (cudaStream_t)stream);
}

extern "C" int msmb200_sumdist(const void *X, int64_t n, int d, int64_t ld, int dtype,
                               int metric, const int64_t *pairs, int64_t p, double *out,
                               void *stream)
{
    (void)n;
    MSMB_REQUIRE(X && out && pairs && d > 0 && ld >= d && p >= 0, "sumdist: bad args");
    if (p == 0) {
        MSMB_CUDA(cudaMemsetAsync(out, 0, sizeof(double), (cudaStream_t)stream));
        return MSMB200_OK;
    }
    return launch_pairs<2>(X, X, p, 0, 0, d, ld, dtype, metric, pairs, out,
                           (cudaStream_t)stream);
}
