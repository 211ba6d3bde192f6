// scan_kernels.cu -- the light scans that sit directly after the distance kernels:
//
//   first_above ......... RegularSpatial's "is this frame farther than d_min from every
//                         centre so far" test (cluster/regularspatial.py:70-77) on the
//                         running-minimum array the k-centers pass maintains
//   label_range / label_presence / transition_counts
//                         msm/core.py:487-602 `_transition_counts` on the label arrays the
//                         assignment kernels leave on the device
//
// All HBM-bound on 4..8 bytes per frame; the histogram is bound by atomic throughput,
// which is why equal bins are merged inside a warp before they touch memory.
#include "common.cuh"
#include <limits.h>

namespace msmb {

constexpr int SCAN_THREADS = 256;
constexpr long long LABEL_MISSING = LLONG_MIN;   // host maps NaN / None to this

// ------------------------------------------------------------------ first_above
__global__ void __launch_bounds__(SCAN_THREADS)
first_above_kernel(const double *__restrict__ v, int64_t n, int64_t start, double threshold,
                   unsigned long long *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long mine = ULLONG_MAX;
    for (int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        if (v[i] > threshold) {      // NaN compares false, as in np.all(d > d_min)
            mine = (unsigned long long)i;
            break;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, mine, off);
        mine = o < mine ? o : mine;
    }
    if ((threadIdx.x & 31) == 0 && mine != ULLONG_MAX) atomicMin(out, mine);
}

// ------------------------------------------------------------------ label scans
template <typename L>
__device__ __forceinline__ long long load_label(const L *p, int64_t i)
{
    return (long long)p[i];
}

__global__ void label_range_init_kernel(long long *out)
{
    out[0] = LLONG_MAX;
    out[1] = LLONG_MIN;
}

template <typename L>
__global__ void __launch_bounds__(SCAN_THREADS)
label_range_kernel(const L *__restrict__ labels, int64_t n, long long *__restrict__ out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    long long lo = LLONG_MAX, hi = LLONG_MIN;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long x = load_label(labels, i);
        if (sizeof(L) == 8 && x == LABEL_MISSING) continue;
        lo = x < lo ? x : lo;
        hi = x > hi ? x : hi;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const long long a = __shfl_xor_sync(0xffffffffu, lo, off);
        const long long b = __shfl_xor_sync(0xffffffffu, hi, off);
        lo = a < lo ? a : lo;
        hi = b > hi ? b : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(out, lo);
        atomicMax(out + 1, hi);
    }
}

template <typename L>
__global__ void __launch_bounds__(SCAN_THREADS)
label_presence_kernel(const L *__restrict__ labels, int64_t n, long long lo, int64_t span,
                      unsigned char *__restrict__ flags)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const long long x = load_label(labels, i);
        if (sizeof(L) == 8 && x == LABEL_MISSING) continue;
        const long long r = x - lo;
        if (r >= 0 && r < span && !flags[r]) flags[r] = 1;   // benign race: all writers store 1
    }
}

// ------------------------------------------------------------------ transition counts
struct CountArgs {
    const void *labels;
    const int64_t *offsets;      // n_seq + 1 row offsets into labels
    int64_t n_seq, n_total, lag;
    const int32_t *remap;        // label - remap_lo -> state, or < 0; NULL: labels are states
    long long remap_lo;
    int64_t remap_len;
    int n_states;
    unsigned long long *counts;  // n_states x n_states, row = from-state
};

template <typename L>
__device__ __forceinline__ int state_of(const CountArgs &A, int64_t p)
{
    const long long x = load_label((const L *)A.labels, p);
    if (A.remap) {
        if (sizeof(L) == 8 && x == LABEL_MISSING) return -1;
        const long long r = x - A.remap_lo;
        return (r >= 0 && r < A.remap_len) ? A.remap[r] : -1;
    }
    return (x >= 0 && x < A.n_states) ? (int)x : -1;
}

// largest s with offsets[s] <= p
__device__ __forceinline__ int64_t sequence_of(const int64_t *__restrict__ offsets, int64_t n_seq,
                                               int64_t p)
{
    int64_t lo = 0, hi = n_seq;               // invariant: offsets[lo] <= p < offsets[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) <= p) lo = mid; else hi = mid;
    }
    return lo;
}

// SHARED: the n_states^2 bins live in shared memory per block (32-bit), flushed once.
template <typename L, bool SHARED>
__global__ void __launch_bounds__(SCAN_THREADS)
transition_counts_kernel(CountArgs A)
{
    extern __shared__ unsigned int bins[];
    const int n_bins = A.n_states * A.n_states;
    if (SHARED) {
        for (int b = threadIdx.x; b < n_bins; b += blockDim.x) bins[b] = 0u;
        __syncthreads();
    }
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count so that match_any sees the whole warp
    const int64_t first = (int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31);
    for (int64_t base = first; base < A.n_total; base += stride) {
        const int64_t p = base + lane;
        long long bin = -1;
        if (p < A.n_total) {
            const int64_t s = sequence_of(A.offsets, A.n_seq, p);
            const int64_t end = __ldg(A.offsets + s + 1);
            if (p + A.lag < end) {
                const int from = state_of<L>(A, p);
                const int to = state_of<L>(A, p + A.lag);
                if (from >= 0 && to >= 0) bin = (long long)from * A.n_states + to;
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, bin);
        if (bin >= 0 && lane == __ffs(peers) - 1) {
            const unsigned c = __popc(peers);
            if (SHARED) atomicAdd(&bins[bin], c);
            else atomicAdd(&A.counts[bin], (unsigned long long)c);
        }
    }
    if (SHARED) {
        __syncthreads();
        for (int b = threadIdx.x; b < n_bins; b += blockDim.x)
            if (bins[b]) atomicAdd(&A.counts[b], (unsigned long long)bins[b]);
    }
}

static int scan_grid(int64_t n, int per_sm)
{
    int64_t blocks = (n + SCAN_THREADS - 1) / SCAN_THREADS;
    const int64_t cap = (int64_t)sm_count() * per_sm;
    if (blocks > cap) blocks = cap;
    return blocks < 1 ? 1 : (int)blocks;
}

}  // namespace msmb

using namespace msmb;

extern "C" int msmb200_first_above(const double *values, int64_t n, int64_t start,
                                   double threshold, int64_t *out_index, void *stream)
{
    MSMB_REQUIRE(values && out_index && n >= 0 && start >= 0, "first_above: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    MSMB_CUDA(cudaMemsetAsync(out_index, 0xFF, sizeof(int64_t), st));   // ULLONG_MAX == "none" (-1)
    if (start >= n) return MSMB200_OK;
    first_above_kernel<<<scan_grid(n - start, 8), SCAN_THREADS, 0, st>>>(
        values, n, start, threshold, (unsigned long long *)out_index);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_label_range(const void *labels, int64_t n, int label_bytes,
                                   int64_t *out_min_max, void *stream)
{
    MSMB_REQUIRE(out_min_max && n >= 0 && (label_bytes == 4 || label_bytes == 8),
                 "label_range: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    label_range_init_kernel<<<1, 1, 0, st>>>((long long *)out_min_max);
    MSMB_LAUNCH_CHECK();
    if (n == 0) return MSMB200_OK;
    MSMB_REQUIRE(labels, "label_range: null labels");
    if (label_bytes == 4)
        label_range_kernel<int32_t><<<scan_grid(n, 8), SCAN_THREADS, 0, st>>>(
            (const int32_t *)labels, n, (long long *)out_min_max);
    else
        label_range_kernel<long long><<<scan_grid(n, 8), SCAN_THREADS, 0, st>>>(
            (const long long *)labels, n, (long long *)out_min_max);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_label_presence(const void *labels, int64_t n, int label_bytes, int64_t lo,
                                      int64_t span, uint8_t *flags, void *stream)
{
    MSMB_REQUIRE(flags && n >= 0 && span > 0 && (label_bytes == 4 || label_bytes == 8),
                 "label_presence: bad args");
    cudaStream_t st = (cudaStream_t)stream;
    MSMB_CUDA(cudaMemsetAsync(flags, 0, (size_t)span, st));
    if (n == 0) return MSMB200_OK;
    MSMB_REQUIRE(labels, "label_presence: null labels");
    if (label_bytes == 4)
        label_presence_kernel<int32_t><<<scan_grid(n, 8), SCAN_THREADS, 0, st>>>(
            (const int32_t *)labels, n, lo, span, flags);
    else
        label_presence_kernel<long long><<<scan_grid(n, 8), SCAN_THREADS, 0, st>>>(
            (const long long *)labels, n, lo, span, flags);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

// ---------------------------------------------------------------------------------------
// LandmarkAgglomerative.predict (cluster/agglomerative.py:234-269): given the (n, L) float64
// distances of n frames to the L landmarks -- columns ordered so that the landmarks of cluster c
// occupy [group_offsets[c], group_offsets[c + 1]) -- pool each cluster's distances
// (POOLING_FUNCTIONS, agglomerative.py:31-43) and give the frame to the cluster with the smallest
// pooled value; clusters are visited in order with a strict '<', empty clusters are skipped
// (agglomerative.py:256-266).  One warp per frame: coalesced reads of its row.
//   pool 0 = average (mean), 1 = complete (max), 2 = single (min),
//        3 = ward: (card * sum(x^2) - sqsum_c) / (card * (card + 1) / 2)
__global__ void __launch_bounds__(256)
pooled_assign_kernel(const double *__restrict__ dists, long long n, int L,
                     const int *__restrict__ group_offsets, int n_clusters, int pool,
                     const double *__restrict__ cardinality, const double *__restrict__ sqsum,
                     int32_t *__restrict__ labels, double *__restrict__ pooled_out,
                     int *__restrict__ any_negative)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = warp; r < n; r += n_warps) {
        const double *row = dists + r * (long long)L;
        double best = INFINITY;
        int best_c = 0;
        bool neg = false;
        for (int c = 0; c < n_clusters; ++c) {
            const int lo = group_offsets[c], hi = group_offsets[c + 1];
            if (hi <= lo) continue;
            double a = pool == 1 ? -INFINITY : pool == 2 ? INFINITY : 0.0;
            for (int j = lo + lane; j < hi; j += 32) {
                const double x = row[j];
                a = pool == 0 ? a + x : pool == 1 ? fmax(a, x) : pool == 2 ? fmin(a, x) : a + x * x;
            }
            for (int off = 16; off > 0; off >>= 1) {
                const double o = __shfl_xor_sync(0xffffffffu, a, off);
                a = (pool == 0 || pool == 3) ? a + o : pool == 1 ? fmax(a, o) : fmin(a, o);
            }
            double d;
            if (pool == 0) d = a / (double)(hi - lo);
            else if (pool == 3) {
                const double card = cardinality[c];
                d = (card * a - sqsum[c]) / (card * (card + 1.0) / 2.0);
            } else d = a;
            neg |= d < 0.0;
            if (d < best) { best = d; best_c = c; }
        }
        if (lane == 0) {
            labels[r] = best_c;
            if (pooled_out) pooled_out[r] = best;
            if (neg && any_negative) atomicOr(any_negative, 1);
        }
    }
}

extern "C" int msmb200_pooled_assign(const double *dists, int64_t n, int n_landmarks,
                                     const int32_t *group_offsets, int n_clusters, int pool,
                                     const double *cardinality, const double *sqsum,
                                     int32_t *labels, double *pooled, int32_t *any_negative,
                                     void *stream)
{
    MSMB_REQUIRE(n >= 0 && n_landmarks > 0 && n_clusters > 0 && pool >= 0 && pool <= 3,
                 "pooled_assign: bad args n=%lld L=%d k=%d pool=%d", (long long)n, n_landmarks,
                 n_clusters, pool);
    if (n == 0) return MSMB200_OK;
    MSMB_REQUIRE(dists && group_offsets && labels, "pooled_assign: null pointer");
    MSMB_REQUIRE(pool != 3 || (cardinality && sqsum), "pooled_assign: ward needs cardinality and sqsum");
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = (n + 7) / 8;
    const long long cap = 16LL * sm_count();
    if (blocks > cap) blocks = cap;
    pooled_assign_kernel<<<(unsigned)blocks, 256, 0, st>>>(dists, n, n_landmarks, group_offsets,
                                                          n_clusters, pool, cardinality, sqsum,
                                                          labels, pooled, any_negative);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_transition_counts(const void *labels, int label_bytes,
                                         const int64_t *seq_offsets, int64_t n_seq, int64_t n_total,
                                         int64_t lag, const int32_t *remap,
                                         int64_t remap_lo, int64_t remap_len, int32_t n_states,
                                         int64_t *counts, void *stream)
{
    MSMB_REQUIRE(label_bytes == 4 || label_bytes == 8, "transition_counts: labels must be int32 or int64");
    MSMB_REQUIRE(n_seq >= 0 && n_total >= 0 && lag >= 1, "transition_counts: bad lag/shape");
    MSMB_REQUIRE(n_states >= 0 && n_states <= 46340, "transition_counts: n_states out of range");
    if (n_states == 0 || n_total == 0 || n_seq == 0) return MSMB200_OK;
    MSMB_REQUIRE(labels && seq_offsets && counts, "transition_counts: null pointer");
    MSMB_REQUIRE(!remap || remap_len > 0, "transition_counts: empty remap table");
    cudaStream_t st = (cudaStream_t)stream;
    CountArgs A{labels, seq_offsets, n_seq, n_total, lag, remap, (long long)remap_lo,
                remap_len, (int)n_states, (unsigned long long *)counts};
    const int n_bins = n_states * n_states;
    const bool shared = n_bins <= 8192;
    const size_t smem = shared ? sizeof(unsigned int) * (size_t)n_bins : 0;
    const int grid = scan_grid(n_total, shared ? 4 : 8);
    if (label_bytes == 4) {
        if (shared) transition_counts_kernel<int32_t, true><<<grid, SCAN_THREADS, smem, st>>>(A);
        else transition_counts_kernel<int32_t, false><<<grid, SCAN_THREADS, 0, st>>>(A);
    } else {
        if (shared) transition_counts_kernel<long long, true><<<grid, SCAN_THREADS, smem, st>>>(A);
        else transition_counts_kernel<long long, false><<<grid, SCAN_THREADS, 0, st>>>(A);
    }
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}
