// rmsd.cu -- K5/K6: minimal-RMSD metric on (n_frames, n_atoms, 3) float32
// coordinates by the quaternion-characteristic-polynomial (QCP) method.
//
// The reference does not contain this arithmetic: libdistance calls mdtraj's
// libtheobald (msmbuilder/libdistance/libdistance.pyx:67-72).  What is matched
// here is the reference's CALL CONTRACT:
//   * frames are centred and G = sum|r|^2 cached as float32 per frame
//     (cluster/base.py:68,133-134; libdistance.pyx:336-341)           -> K5
//   * rmsd = sqrtf(msd(x_i, y_j, G_i, G_j)) as a float widened to double
//     (libdistance.pyx:350-351,555-556)                               -> K6
//   * strict '<' running minimum / first-index tie-breaks as for the vector
//     metrics (libdistance.pyx:352-354; kcenters.py:93-97).
// QCP (Theobald 2005; Liu, Agrafiotis & Theobald 2010): with M = sum_a x_a y_a^T,
// lambda_max of the 4x4 key matrix K(M) is the largest root of
//   P(l) = l^4 + c2 l^2 + c1 l + c0,  c2 = -2 |M|_F^2,  c1 = -8 det M,  c0 = det K,
// found by Newton from l0 = (G_a + G_b)/2;  msd = (G_a + G_b - 2 l)/n_atoms,
// clamped at 0.  The 3x3 products are accumulated in float64.
// HBM-bound for the one-centre pass: 12*n_atoms + 4 + 8 bytes per frame.
#include "common.cuh"

namespace msmb {

static constexpr int kRThreads = 256;

struct Mat3 { double m[9]; };

// warp-cooperative M = sum_a x_a (x) y_a ; every lane gets the full sums
__device__ __forceinline__ Mat3 inner_products(const float *__restrict__ x,
                                               const float *__restrict__ y, int n_atoms,
                                               int lane)
{
    Mat3 r;
#pragma unroll
    for (int q = 0; q < 9; ++q) r.m[q] = 0.0;
    for (int a = lane; a < n_atoms; a += 32) {
        const double x0 = x[3 * a], x1 = x[3 * a + 1], x2 = x[3 * a + 2];
        const double y0 = y[3 * a], y1 = y[3 * a + 1], y2 = y[3 * a + 2];
        r.m[0] = fma(x0, y0, r.m[0]); r.m[1] = fma(x0, y1, r.m[1]); r.m[2] = fma(x0, y2, r.m[2]);
        r.m[3] = fma(x1, y0, r.m[3]); r.m[4] = fma(x1, y1, r.m[4]); r.m[5] = fma(x1, y2, r.m[5]);
        r.m[6] = fma(x2, y0, r.m[6]); r.m[7] = fma(x2, y1, r.m[7]); r.m[8] = fma(x2, y2, r.m[8]);
    }
#pragma unroll
    for (int q = 0; q < 9; ++q)
        for (int off = 16; off > 0; off >>= 1)
            r.m[q] += __shfl_xor_sync(0xffffffffu, r.m[q], off);
    return r;
}

// G-lane cooperative M = sum_a x_a (x) y_a for the one-centre pass: y lives in shared memory
// already widened to double; when n_atoms % 4 == 0 each lane streams whole 4-atom groups
// (three 16-byte loads = 48 contiguous bytes), otherwise one atom at a time.
template <int G>
__device__ __forceinline__ Mat3 inner_products_group(const float *__restrict__ x,
                                                     const double *__restrict__ sy, int n_atoms,
                                                     int lane_in_group)
{
    Mat3 r;
#pragma unroll
    for (int q = 0; q < 9; ++q) r.m[q] = 0.0;
    if ((n_atoms & 3) == 0) {
        const float4 *x4 = reinterpret_cast<const float4 *>(x);
        const int n_quads = n_atoms >> 2;
        for (int qd = lane_in_group; qd < n_quads; qd += G) {
            const float4 v0 = ldg_stream(x4 + 3 * qd), v1 = ldg_stream(x4 + 3 * qd + 1),
                         v2 = ldg_stream(x4 + 3 * qd + 2);
            const double xs[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
            const double *y = sy + 12 * qd;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const double x0 = xs[3 * a], x1 = xs[3 * a + 1], x2 = xs[3 * a + 2];
                const double y0 = y[3 * a], y1 = y[3 * a + 1], y2 = y[3 * a + 2];
                r.m[0] = fma(x0, y0, r.m[0]); r.m[1] = fma(x0, y1, r.m[1]); r.m[2] = fma(x0, y2, r.m[2]);
                r.m[3] = fma(x1, y0, r.m[3]); r.m[4] = fma(x1, y1, r.m[4]); r.m[5] = fma(x1, y2, r.m[5]);
                r.m[6] = fma(x2, y0, r.m[6]); r.m[7] = fma(x2, y1, r.m[7]); r.m[8] = fma(x2, y2, r.m[8]);
            }
        }
    } else {
        for (int a = lane_in_group; a < n_atoms; a += G) {
            const double x0 = x[3 * a], x1 = x[3 * a + 1], x2 = x[3 * a + 2];
            const double y0 = sy[3 * a], y1 = sy[3 * a + 1], y2 = sy[3 * a + 2];
            r.m[0] = fma(x0, y0, r.m[0]); r.m[1] = fma(x0, y1, r.m[1]); r.m[2] = fma(x0, y2, r.m[2]);
            r.m[3] = fma(x1, y0, r.m[3]); r.m[4] = fma(x1, y1, r.m[4]); r.m[5] = fma(x1, y2, r.m[5]);
            r.m[6] = fma(x2, y0, r.m[6]); r.m[7] = fma(x2, y1, r.m[7]); r.m[8] = fma(x2, y2, r.m[8]);
        }
    }
#pragma unroll
    for (int q = 0; q < 9; ++q)
#pragma unroll
        for (int off = G >> 1; off > 0; off >>= 1)
            r.m[q] += __shfl_xor_sync(0xffffffffu, r.m[q], off);
    return r;
}

__device__ __forceinline__ double det3(double a, double b, double c, double d, double e,
                                       double f, double g, double h, double i)
{
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

// rmsd (as the float the reference would hold, widened) from M and the traces
__device__ __forceinline__ double qcp_rmsd(const Mat3 &M, double Ga, double Gb, int n_atoms)
{
    const double Sxx = M.m[0], Sxy = M.m[1], Sxz = M.m[2];
    const double Syx = M.m[3], Syy = M.m[4], Syz = M.m[5];
    const double Szx = M.m[6], Szy = M.m[7], Szz = M.m[8];
    double fro = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) fro = fma(M.m[q], M.m[q], fro);
    const double c2 = -2.0 * fro;
    const double c1 = -8.0 * det3(Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz);
    // symmetric key matrix
    const double k00 = Sxx + Syy + Szz, k01 = Syz - Szy, k02 = Szx - Sxz, k03 = Sxy - Syx;
    const double k11 = Sxx - Syy - Szz, k12 = Sxy + Syx, k13 = Szx + Sxz;
    const double k22 = -Sxx + Syy - Szz, k23 = Syz + Szy;
    const double k33 = -Sxx - Syy + Szz;
    // det K by expansion along the first row
    const double c0 =
          k00 * det3(k11, k12, k13, k12, k22, k23, k13, k23, k33)
        - k01 * det3(k01, k12, k13, k02, k22, k23, k03, k23, k33)
        + k02 * det3(k01, k11, k13, k02, k12, k23, k03, k13, k33)
        - k03 * det3(k01, k11, k12, k02, k12, k22, k03, k13, k23);
    double l = 0.5 * (Ga + Gb);
    for (int it = 0; it < 50; ++it) {
        const double l2 = l * l;
        const double p = (l2 + c2) * l2 + c1 * l + c0;
        const double dp = 4.0 * l2 * l + 2.0 * c2 * l + c1;
        if (dp == 0.0) break;
        const double step = p / dp;
        l -= step;
        if (fabs(step) <= 1e-11 * fabs(l)) break;
    }
    double msd = (Ga + Gb - 2.0 * l) / (double)n_atoms;
    if (!(msd > 0.0)) msd = 0.0;                 // clamp (reference: NaN from sqrtf of -eps)
    const float msd_f = (float)msd;              // msd_atom_major returns float
    return (double)sqrtf(msd_f);                 // libdistance.pyx:350: float sqrt
}

// ---- K5: centre each frame in place, G = sum |r|^2 -----------------------------
__global__ void __launch_bounds__(kRThreads)
rmsd_center_kernel(float *__restrict__ xyz, long long n, int n_atoms, float *__restrict__ traces)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long f = warp; f < n; f += n_warps) {
        float *x = xyz + f * (long long)n_atoms * 3;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0;
        for (int a = lane; a < n_atoms; a += 32) {
            s0 += (double)x[3 * a]; s1 += (double)x[3 * a + 1]; s2 += (double)x[3 * a + 2];
        }
        for (int off = 16; off > 0; off >>= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
            s2 += __shfl_xor_sync(0xffffffffu, s2, off);
        }
        const double inv = 1.0 / (double)n_atoms;
        const double m0 = s0 * inv, m1 = s1 * inv, m2 = s2 * inv;
        double g = 0.0;
        for (int a = lane; a < n_atoms; a += 32) {
            const float c0 = (float)((double)x[3 * a] - m0);
            const float c1 = (float)((double)x[3 * a + 1] - m1);
            const float c2 = (float)((double)x[3 * a + 2] - m2);
            x[3 * a] = c0; x[3 * a + 1] = c1; x[3 * a + 2] = c2;
            g += (double)c0 * c0 + (double)c1 * c1 + (double)c2 * c2;
        }
        for (int off = 16; off > 0; off >>= 1) g += __shfl_xor_sync(0xffffffffu, g, off);
        if (lane == 0) traces[f] = (float)g;
    }
}

struct BlockCandR { double v; long long i; };

// ---- K6a: one k-centers pass under RMSD (same contract as kcenters_pass) ---------
__global__ void __launch_bounds__(kRThreads)
rmsd_pass_kernel(const float *__restrict__ xyz, const float *__restrict__ traces, long long n,
                 int n_atoms, const float *__restrict__ center, int label,
                 double *__restrict__ dist, int *__restrict__ labels, long long row_offset,
                 BlockCandR *__restrict__ block_cands, unsigned *__restrict__ counter,
                 msmb200_candidate *__restrict__ out)
{
    constexpr int G = 8;                            // lanes per frame: 4 frames per warp in flight
    extern __shared__ __align__(16) double s_cd[];  // n_atoms*3 centre coordinates, widened
    const int n3 = n_atoms * 3;
    for (int j = threadIdx.x; j < n3; j += blockDim.x) s_cd[j] = (double)center[j];
    __syncthreads();
    const double Gc = (double)center[n3];
    const int lane = threadIdx.x & 31;
    const int lane_in_group = lane & (G - 1), group_in_warp = lane / G;
    const long long warp_group0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / G);
    const long long n_groups = ((long long)gridDim.x * blockDim.x) / G;

    ArgMax best{-INFINITY, 0x7fffffffffffffffLL};
    for (long long f0 = warp_group0; f0 < n; f0 += n_groups) {
        const long long f_raw = f0 + group_in_warp;
        const bool valid = f_raw < n;
        const long long f = valid ? f_raw : n - 1;
        const float *x = xyz + f * (long long)n3;
        double cur = INFINITY, Gx = 0.0;
        if (valid && lane_in_group == 0) {           // prefetch before the arithmetic
            cur = __ldcg(dist + f);
            Gx = (double)traces[f];
        }
        Mat3 M = inner_products_group<G>(x, s_cd, n_atoms, lane_in_group);
        if (valid && lane_in_group == 0) {
            const double dv = qcp_rmsd(M, Gx, Gc, n_atoms);
            if (dv < cur) {
                cur = dv;
                dist[f] = dv;
                labels[f] = label;
            }
            if (cur > best.v) {
                best.v = cur;
                best.i = f;
            }
        }
    }
    __shared__ ArgMax s_warp[kRThreads / 32];
    __shared__ bool s_is_last;
    best = argmax_warp(best);
    if (lane == 0) s_warp[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        ArgMax b = (threadIdx.x < kRThreads / 32) ? s_warp[threadIdx.x]
                                                   : ArgMax{-INFINITY, 0x7fffffffffffffffLL};
        b = argmax_warp(b);
        if (threadIdx.x == 0) {
            block_cands[blockIdx.x].v = b.v;
            block_cands[blockIdx.x].i = b.i;
            __threadfence();
            unsigned ticket = atomicInc(counter, gridDim.x - 1);
            s_is_last = (ticket == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (!s_is_last) return;
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
    if (lane == 0) s_warp[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x < 32) {
        ArgMax b = (threadIdx.x < kRThreads / 32) ? s_warp[threadIdx.x]
                                                   : ArgMax{-INFINITY, 0x7fffffffffffffffLL};
        b = argmax_warp(b);
        if (threadIdx.x == 0) s_warp[0] = b;
    }
    __syncthreads();
    w = s_warp[0];
    if (w.i == 0x7fffffffffffffffLL) w.i = 0;
    if (threadIdx.x == 0) {
        out->value = w.v;
        out->index = row_offset + w.i;
    }
    float *payload = reinterpret_cast<float *>(out + 1);
    if (n > 0) {
        for (int j = threadIdx.x; j < n_atoms * 3; j += blockDim.x)
            payload[j] = xyz[w.i * (long long)n_atoms * 3 + j];
        if (threadIdx.x == 0) payload[n_atoms * 3] = traces[w.i];
    }
}

// ---- K6b: assign_nearest / dist / pdist under RMSD --------------------------------
// MODE 0: assign (each warp: one frame vs all k centres)
// MODE 1: dist   (each warp: one frame vs y)
// MODE 2: pdist  (each warp: one condensed pair)
template <int MODE>
__global__ void __launch_bounds__(kRThreads)
rmsd_multi_kernel(const float *__restrict__ xyz, const float *__restrict__ traces,
                  long long n_items, int n_atoms, const float *__restrict__ Y,
                  const float *__restrict__ Y_traces, int k,
                  const long long *__restrict__ rows, long long m, int *__restrict__ labels,
                  double *__restrict__ out, double *__restrict__ block_sums)
{
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const long long fs = (long long)n_atoms * 3;
    double local = 0.0;
    for (long long it = warp; it < n_items; it += n_warps) {
        if (MODE == 0) {
            const long long r = rows ? rows[it] : it;
            const float *x = xyz + r * fs;
            const double Gx = (double)traces[r];
            float best = 3.402823466e+38f;   // FLT_MAX, libdistance.pyx:347
            int arg = 0;
            for (int j = 0; j < k; ++j) {
                Mat3 M = inner_products(x, Y + (long long)j * fs, n_atoms, lane);
                const float dv = (float)qcp_rmsd(M, Gx, (double)Y_traces[j], n_atoms);
                if (dv < best) {
                    best = dv;
                    arg = j;
                }
            }
            if (lane == 0) {
                labels[it] = arg;
                if (out) out[it] = (double)best;
                local += (double)best;
            }
        } else if (MODE == 1) {
            const long long r = rows ? rows[it] : it;
            Mat3 M = inner_products(xyz + r * fs, Y, n_atoms, lane);
            const double dv = qcp_rmsd(M, (double)traces[r], (double)Y_traces[0], n_atoms);
            if (lane == 0) out[it] = dv;
        } else {
            const double mm = (double)m;
            long long i = (long long)floor(((2.0 * mm - 1.0) - sqrt((2.0 * mm - 1.0) * (2.0 * mm - 1.0) - 8.0 * (double)it)) * 0.5);
            if (i < 0) i = 0;
            while (i > 0 && m * i - i * (i + 1) / 2 > it) --i;
            while (m * (i + 1) - (i + 1) * (i + 2) / 2 <= it) ++i;
            const long long j = it - (m * i - i * (i + 1) / 2) + i + 1;
            const long long ra = rows ? rows[i] : i, rb = rows ? rows[j] : j;
            Mat3 M = inner_products(xyz + ra * fs, xyz + rb * fs, n_atoms, lane);
            const double dv = qcp_rmsd(M, (double)traces[ra], (double)traces[rb], n_atoms);
            if (lane == 0) out[it] = dv;
        }
    }
    if (MODE == 0) {
        __shared__ double s_sum[kRThreads / 32];
        for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
        if (lane == 0) s_sum[threadIdx.x >> 5] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < kRThreads / 32; ++w) s += s_sum[w];
            block_sums[blockIdx.x] = s;
        }
    }
}

__global__ void rmsd_sum_partials(const double *__restrict__ partials, int n, double *out)
{
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < n; ++i) t += partials[i];
        *out = t;
    }
}

static inline int warp_grid(long long items)
{
    long long blocks = (items + (kRThreads / 32) - 1) / (kRThreads / 32);
    long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace msmb

using namespace msmb;

extern "C" int msmb200_rmsd_center(float *xyz, int64_t n, int n_atoms, float *traces,
                                   void *stream)
{
    MSMB_REQUIRE(xyz && traces && n >= 0 && n_atoms > 0, "rmsd_center: bad args");
    if (n == 0) return MSMB200_OK;
    rmsd_center_kernel<<<warp_grid(n), kRThreads, 0, (cudaStream_t)stream>>>(xyz, n, n_atoms, traces);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_rmsd_kcenters_pass(const float *xyz, const float *traces, int64_t n,
                                          int n_atoms, const float *center,
                                          int32_t center_label, double *distances,
                                          int32_t *labels, int64_t row_offset,
                                          msmb200_candidate *out, void *workspace,
                                          size_t workspace_bytes, void *stream)
{
    MSMB_REQUIRE(xyz && traces && center && distances && labels && out && workspace && n >= 0 &&
                 n_atoms > 0, "rmsd_kcenters_pass: bad args");
    int grid = warp_grid((n + 3) / 4);          // 4 frames per warp iteration
    MSMB_REQUIRE(workspace_bytes >= 16 + sizeof(BlockCandR) * (size_t)grid,
                 "rmsd_kcenters_pass: workspace too small");
    unsigned *counter = reinterpret_cast<unsigned *>(workspace);
    BlockCandR *cands = reinterpret_cast<BlockCandR *>(reinterpret_cast<unsigned char *>(workspace) + 16);
    size_t smem = (((size_t)n_atoms * 3) * sizeof(double) + 15) & ~(size_t)15;
    if (smem > 48 * 1024)
        MSMB_CUDA(cudaFuncSetAttribute(rmsd_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rmsd_pass_kernel<<<grid, kRThreads, smem, (cudaStream_t)stream>>>(
        xyz, traces, n, n_atoms, center, center_label, distances, labels, row_offset, cands,
        counter, out);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_rmsd_assign_nearest(const float *xyz, const float *traces, int64_t n,
                                           int n_atoms, const float *Y, const float *Y_traces,
                                           int k, const int64_t *rows, int64_t n_rows,
                                           int32_t *labels, double *min_dist, double *inertia,
                                           void *stream)
{
    MSMB_REQUIRE(xyz && traces && Y && Y_traces && labels && k > 0 && n_atoms > 0,
                 "rmsd_assign_nearest: bad args");
    const int64_t n_out = rows ? n_rows : n;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = warp_grid(n_out);
    double *partials = nullptr;
    MSMB_CUDA(cudaMallocAsync(&partials, sizeof(double) * grid, st));
    rmsd_multi_kernel<0><<<grid, kRThreads, 0, st>>>(xyz, traces, n_out, n_atoms, Y, Y_traces, k,
                                                    (const long long *)rows, 0, labels, min_dist,
                                                    partials);
    MSMB_LAUNCH_CHECK();
    if (inertia) {
        rmsd_sum_partials<<<1, 32, 0, st>>>(partials, grid, inertia);
        MSMB_LAUNCH_CHECK();
    }
    MSMB_CUDA(cudaFreeAsync(partials, st));
    return MSMB200_OK;
}

extern "C" int msmb200_rmsd_dist(const float *xyz, const float *traces, int64_t n, int n_atoms,
                                 const float *y, float y_trace, const int64_t *rows,
                                 int64_t n_rows, double *out, void *stream)
{
    MSMB_REQUIRE(xyz && traces && y && out && n_atoms > 0, "rmsd_dist: bad args");
    const int64_t n_out = rows ? n_rows : n;
    if (n_out == 0) return MSMB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    float *d_trace = nullptr;
    MSMB_CUDA(cudaMallocAsync(&d_trace, sizeof(float), st));
    MSMB_CUDA(cudaMemcpyAsync(d_trace, &y_trace, sizeof(float), cudaMemcpyHostToDevice, st));
    rmsd_multi_kernel<1><<<warp_grid(n_out), kRThreads, 0, st>>>(
        xyz, traces, n_out, n_atoms, y, d_trace, 1, (const long long *)rows, 0, nullptr, out,
        nullptr);
    MSMB_LAUNCH_CHECK();
    MSMB_CUDA(cudaFreeAsync(d_trace, st));
    return MSMB200_OK;
}

extern "C" int msmb200_rmsd_pdist(const float *xyz, const float *traces, int64_t n, int n_atoms,
                                  const int64_t *rows, int64_t n_rows, double *out, void *stream)
{
    MSMB_REQUIRE(xyz && traces && out && n_atoms > 0, "rmsd_pdist: bad args");
    const long long m = rows ? n_rows : n;
    if (m < 2) return MSMB200_OK;
    const long long pairs = m * (m - 1) / 2;
    rmsd_multi_kernel<2><<<warp_grid(pairs), kRThreads, 0, (cudaStream_t)stream>>>(
        xyz, traces, pairs, n_atoms, nullptr, nullptr, 0, (const long long *)rows, m, nullptr,
        out, nullptr);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}
