// tica_simt.cu -- K1 on CUDA cores in float64: the exact engine of libmsmb200.
//
// Computes, for every sequence and every t in [0, n - lag):
//     C_tau += x_t x_{t+lag}^T      C_00 += x_t x_t^T      C_tt += x_{t+lag} x_{t+lag}^T
//     S_0 += x_t                    S_tau += x_{t+lag}     S += x_t (all t < n)
// exactly the six accumulations of tICA._fit (msmbuilder/decomposition/tica.py:417-422),
// in float64 like the reference (tica.py:402), for ANY n_features / lag / input
// dtype.  It is (a) the engine for shapes the tcgen05 kernel does not take,
// (b) the on-device float64 yardstick the tensor-core engine is tested against
// at sizes no CPU oracle finishes.  FP64-pipe bound (6*D^2 flop/frame).
#include "common.cuh"
#include <vector>

namespace msmb {

struct TicaItem {
    const void *base;   // sequence base pointer
    long long t0;       // first pair index handled by this item
    int count;          // number of pair indices (t0 .. t0+count)
    int pad;
};

static constexpr int TS = 64;     // output tile edge
static constexpr int RS = 16;     // rows staged per step
static constexpr int kChunk = 4096;

template <typename T>
__global__ void __launch_bounds__(256)
tica_outer_kernel(const TicaItem *__restrict__ items, int n_items, int D, long long ld, int lag,
                  double *__restrict__ acc, const int *__restrict__ run_if)
{
    // rescue launch of the tensor-core engine: uniform over the grid
    if (run_if != nullptr && *reinterpret_cast<const volatile int *>(run_if) == 0) return;
    __shared__ double sA0[RS][TS], sB0[RS][TS], sAt[RS][TS], sBt[RS][TS];
    const int ci = blockIdx.z * TS;   // row block of the output (features i)
    const int cj = blockIdx.y * TS;   // col block of the output (features j)
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    double c_tau[4][4] = {}, c_00[4][4] = {}, c_tt[4][4] = {};

    // a block walks over items blockIdx.x, blockIdx.x + gridDim.x, ... and adds its sums once at the
    // end: the grid stays a few blocks per SM whatever the number of frames (the guarded rescue
    // launch of the tensor-core engine then costs microseconds, not the 0.5 ms of 200k empty blocks)
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const TicaItem it = items[item];
    const T *X = reinterpret_cast<const T *>(it.base);
    for (int r0 = 0; r0 < it.count; r0 += RS) {
        // stage RS rows x 64 columns of the four panels, widening to double
        for (int e = threadIdx.x; e < RS * TS; e += 256) {
            const int r = e / TS, c = e % TS;
            const bool row_ok = (r0 + r) < it.count;
            const long long t = it.t0 + r0 + r;
            double a0 = 0.0, b0 = 0.0, at = 0.0, bt = 0.0;
            if (row_ok) {
                if (ci + c < D) {
                    a0 = (double)X[t * ld + ci + c];
                    at = (double)X[(t + lag) * ld + ci + c];
                }
                if (cj + c < D) {
                    b0 = (double)X[t * ld + cj + c];
                    bt = (double)X[(t + lag) * ld + cj + c];
                }
            }
            sA0[r][c] = a0; sB0[r][c] = b0; sAt[r][c] = at; sBt[r][c] = bt;
        }
        __syncthreads();
#pragma unroll 4
        for (int r = 0; r < RS; ++r) {
            double a0[4], b0[4], at[4], bt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                a0[u] = sA0[r][ty * 4 + u];
                at[u] = sAt[r][ty * 4 + u];
                b0[u] = sB0[r][tx * 4 + u];
                bt[u] = sBt[r][tx * 4 + u];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    c_tau[u][v] = fma(a0[u], bt[v], c_tau[u][v]);
                    c_00[u][v] = fma(a0[u], b0[v], c_00[u][v]);
                    c_tt[u][v] = fma(at[u], bt[v], c_tt[u][v]);
                }
        }
        __syncthreads();
    }
    }

    const size_t DD = (size_t)D * D;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int i = ci + ty * 4 + u;
        if (i >= D) continue;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int j = cj + tx * 4 + v;
            if (j >= D) continue;
            atomicAdd(&acc[(size_t)i * D + j], c_tau[u][v]);
            atomicAdd(&acc[DD + (size_t)i * D + j], c_00[u][v]);
            atomicAdd(&acc[2 * DD + (size_t)i * D + j], c_tt[u][v]);
        }
    }
}

// Column sums S_0, S_tau, S.  One block per item; thread per column (strided).
// For the last item of a sequence `tail` > 0 extra rows [t0+count, t0+count+tail)
// exist only in S (and in S_tau through the +lag shift).
template <typename T>
__global__ void __launch_bounds__(256)
tica_sums_kernel(const TicaItem *__restrict__ items, int D, long long ld, int lag,
                 double *__restrict__ acc, const int *__restrict__ run_if)
{
    if (run_if != nullptr && *reinterpret_cast<const volatile int *>(run_if) == 0) return;
    const TicaItem it = items[blockIdx.x];
    const T *X = reinterpret_cast<const T *>(it.base);
    double *S0 = acc + 3 * (size_t)D * D;
    double *St = S0 + D;
    double *S = St + D;
    const bool first = (it.t0 == 0);
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        double s0 = 0.0, st = 0.0;
        for (int r = 0; r < it.count; ++r) {
            const long long t = it.t0 + r;
            s0 += (double)X[t * ld + c];
            st += (double)X[(t + lag) * ld + c];
        }
        // S = sum over ALL rows = (rows < n-lag, i.e. s0 pieces) + (last lag rows).
        // The last lag rows are rows t+lag for the final lag pair indices; instead
        // of special-casing, use S = S_tau + (first lag rows): rows [0, lag).
        double s = st;
        if (first) {
            for (int r = 0; r < lag; ++r) s += (double)X[(long long)r * ld + c];
        }
        atomicAdd(&S0[c], s0);
        atomicAdd(&St[c], st);
        atomicAdd(&S[c], s);
    }
}

__global__ void tica_counts_kernel(double *acc, int D, double n_obs, double n_seq, const int *run_if)
{
    if (run_if != nullptr && *reinterpret_cast<const volatile int *>(run_if) == 0) return;
    double *tail = acc + 3 * (size_t)D * D + 3 * (size_t)D;
    tail[0] += n_obs;
    tail[1] += n_seq;
}

// host side of the item table: one item per <= kChunk pair indices of a usable sequence
size_t tica_simt_items(const void *const *seq_ptrs, const int64_t *seq_rows, int n_seq, int lag,
                       void *items_out /* NULL: count only */, double *n_obs_out, double *n_used_out)
{
    TicaItem *out = reinterpret_cast<TicaItem *>(items_out);
    size_t n_items = 0;
    double n_obs = 0.0, n_used = 0.0;
    for (int s = 0; s < n_seq; ++s) {
        const long long n = seq_rows[s];
        if (!(n > lag)) continue;   // tica.py:410-412: skipped, not counted
        n_obs += (double)n;
        n_used += 1.0;
        const long long pairs = n - lag;
        for (long long t0 = 0; t0 < pairs; t0 += kChunk) {
            if (out) {
                TicaItem it;
                it.base = seq_ptrs[s];
                it.t0 = t0;
                it.count = (int)((pairs - t0) < kChunk ? (pairs - t0) : kChunk);
                it.pad = 0;
                out[n_items] = it;
            }
            ++n_items;
        }
    }
    if (n_obs_out) *n_obs_out = n_obs;
    if (n_used_out) *n_used_out = n_used;
    return n_items;
}
size_t tica_simt_item_bytes() { return sizeof(TicaItem); }

// the three launches on a device-resident item table; with `run_if` they only work when *run_if != 0
int tica_simt_launch(const void *d_items, size_t n_items, int D, int64_t ld, int dtype, int lag,
                     double n_obs, double n_used, double *acc, const int *run_if, cudaStream_t st)
{
    if (n_items == 0) return MSMB200_OK;
    const TicaItem *items = reinterpret_cast<const TicaItem *>(d_items);
    const int tiles = (D + TS - 1) / TS;
    // ~4 resident blocks of this kernel per SM (36 KB of shared memory each)
    size_t gx = ((size_t)sm_count() * 4 + (size_t)tiles * tiles - 1) / ((size_t)tiles * tiles);
    if (gx > n_items) gx = n_items;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, tiles, tiles);
    if (dtype == MSMB200_F64) {
        tica_outer_kernel<double><<<grid, 256, 0, st>>>(items, (int)n_items, D, ld, lag, acc, run_if);
        MSMB_LAUNCH_CHECK();
        tica_sums_kernel<double><<<(unsigned)n_items, 256, 0, st>>>(items, D, ld, lag, acc, run_if);
    } else {
        tica_outer_kernel<float><<<grid, 256, 0, st>>>(items, (int)n_items, D, ld, lag, acc, run_if);
        MSMB_LAUNCH_CHECK();
        tica_sums_kernel<float><<<(unsigned)n_items, 256, 0, st>>>(items, D, ld, lag, acc, run_if);
    }
    MSMB_LAUNCH_CHECK();
    tica_counts_kernel<<<1, 1, 0, st>>>(acc, D, n_obs, n_used, run_if);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

int tica_simt_accumulate(const void *const *seq_ptrs, const int64_t *seq_rows, int n_seq,
                         int D, int64_t ld, int dtype, int lag, double *acc, cudaStream_t st)
{
    double n_obs = 0.0, n_used = 0.0;
    const size_t n_items = tica_simt_items(seq_ptrs, seq_rows, n_seq, lag, nullptr, nullptr, nullptr);
    if (n_items == 0) return MSMB200_OK;
    std::vector<TicaItem> items(n_items);
    tica_simt_items(seq_ptrs, seq_rows, n_seq, lag, items.data(), &n_obs, &n_used);
    TicaItem *d_items = nullptr;
    MSMB_CUDA(cudaMallocAsync(&d_items, sizeof(TicaItem) * n_items, st));
    MSMB_CUDA(cudaMemcpyAsync(d_items, items.data(), sizeof(TicaItem) * n_items,
                              cudaMemcpyHostToDevice, st));
    // the host vector must outlive the (pageable => staged) copy
    MSMB_CUDA(cudaStreamSynchronize(st));
    const int rc = tica_simt_launch(d_items, n_items, D, ld, dtype, lag, n_obs, n_used, acc, nullptr, st);
    MSMB_CUDA(cudaFreeAsync(d_items, st));
    return rc;
}

// ---------------------------------------------------------------------------
// tICA.transform (tica.py:330-336): out = (X - mu) @ comps^T [* scale], float64.
// HBM-bound skinny product: one sub-warp per frame, k outputs per frame.
// ---------------------------------------------------------------------------
template <typename T, int KMAX>
__global__ void __launch_bounds__(256)
tica_transform_kernel(const T *__restrict__ X, long long n, int D, long long ld,
                      const double *__restrict__ mu, const double *__restrict__ comps,
                      const double *__restrict__ scale, int k, int k0,
                      double *__restrict__ out)
{
    // each warp handles one frame at a time; lanes stride over features
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int kk = (k - k0) < KMAX ? (k - k0) : KMAX;
    for (long long r = warp; r < n; r += n_warps) {
        double accv[KMAX];
#pragma unroll
        for (int c = 0; c < KMAX; ++c) accv[c] = 0.0;
        for (int j = lane; j < D; j += 32) {
            // reference order: (X - means) first, in float64 (tica.py:331-333)
            const double xc = (double)X[r * ld + j] - mu[j];
#pragma unroll
            for (int c = 0; c < KMAX; ++c)
                if (c < kk) accv[c] = fma(xc, comps[(size_t)(k0 + c) * D + j], accv[c]);
        }
#pragma unroll
        for (int c = 0; c < KMAX; ++c) {
            double v = accv[c];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0 && c < kk) {
                if (scale) v *= scale[k0 + c];
                out[r * (long long)k + k0 + c] = v;
            }
        }
    }
}

}  // namespace msmb

using namespace msmb;

extern "C" int msmb200_tica_transform(const void *X, int64_t n, int n_features, int64_t ld,
                                      int dtype, const double *means, const double *comps,
                                      const double *scale, int k, double *out, void *stream)
{
    MSMB_REQUIRE(X && means && comps && out && n >= 0 && n_features > 0 && k > 0 &&
                 ld >= n_features, "tica_transform: bad args");
    if (n == 0) return MSMB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    long long blocks = (n + 7) / 8;
    long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    constexpr int KMAX = 8;
    for (int k0 = 0; k0 < k; k0 += KMAX) {
        if (dtype == MSMB200_F64)
            tica_transform_kernel<double, KMAX><<<(unsigned)blocks, 256, 0, st>>>(
                (const double *)X, n, n_features, ld, means, comps, scale, k, k0, out);
        else
            tica_transform_kernel<float, KMAX><<<(unsigned)blocks, 256, 0, st>>>(
                (const float *)X, n, n_features, ld, means, comps, scale, k, k0, out);
        MSMB_LAUNCH_CHECK();
    }
    return MSMB200_OK;
}
