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

static int rmsd_pass_tiles(const float *xyz, const float *traces, int64_t n, int n_atoms,
                           const float *center, const unsigned char *slots, size_t slot_stride,
                           int n_prev, int label, double *distances, int32_t *labels,
                           int64_t row_offset, msmb200_candidate *out, void *workspace,
                           size_t workspace_bytes, cudaStream_t st);

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

// =========================================================================================
// Reference-class arithmetic (default engine for n_atoms <= kTileMaxAtoms): what the reference CALLS
// (libdistance.pyx:350-351: rmsd = sqrtf(msd_atom_major(...)), mdtraj's libtheobald) restated from
// its published structure -- M accumulated in FLOAT32 four atoms at a time (one SIMD lane per
// atom % 4: a rounded multiply and a rounded add per atom, (s0 + s1) + (s2 + s3) at the end), the
// quartic and the Newton iteration of qcprot.c in double (E0 = (G_a + G_b) / 2, evalprec 1e-11,
// <= 50 steps, msd = |2 (E0 - lambda) / n| returned as a float), sqrtf.  Every double operation is
// written with a non-contracting intrinsic so that oracle/rmsd_oracle.py (rmsd_theobald_f32) can
// follow it operation by operation: the GPU tests compare BIT FOR BIT against that restatement.
//
// Work layout: ONE THREAD PER FRAME.  A warp stages up to 32 frames in shared memory with cp.async
// (16-byte copies when frames are 16-byte aligned), then every lane walks its own frame against
// the centre (shared memory, broadcast reads) -- no cross-lane reduction, all 32 lanes run the
// Newton iteration at once (the old kernel had 1 lane in 8 doing float64 work).
// =========================================================================================
static constexpr int kTileMaxAtoms = 128;        // 32 frames x 1.5 KB = 48 KB of shared memory per warp
static constexpr int kTileWarps = 4;
static constexpr int kTileThreads = 32 * kTileWarps;

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }
__device__ __forceinline__ double det3s(double a, double b, double c, double d, double e, double f,
                                        double g, double h, double i)
{
    // (a (e i - f h) - b (d i - f g)) + c (d h - e g), every operation rounded
    return dadd(dsub(dmul(a, dsub(dmul(e, i), dmul(f, h))), dmul(b, dsub(dmul(d, i), dmul(f, g)))),
                dmul(c, dsub(dmul(d, h), dmul(e, g))));
}
// float rmsd from the float32 M and traces (oracle: rmsd_oracle.qcp_strict)
__device__ __forceinline__ float qcp_rmsd_strict(const float *Mf, float Ga, float Gb, int n_atoms)
{
    double m[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) m[q] = (double)Mf[q];
    const double Sxx = m[0], Sxy = m[1], Sxz = m[2], Syx = m[3], Syy = m[4], Syz = m[5],
                 Szx = m[6], Szy = m[7], Szz = m[8];
    double fro = 0.0;
#pragma unroll
    for (int q = 0; q < 9; ++q) fro = dadd(fro, dmul(m[q], m[q]));
    const double c2 = dmul(-2.0, fro);
    const double c1 = dmul(-8.0, det3s(Sxx, Sxy, Sxz, Syx, Syy, Syz, Szx, Szy, Szz));
    const double k00 = dadd(dadd(Sxx, Syy), Szz), k01 = dsub(Syz, Szy), k02 = dsub(Szx, Sxz),
                 k03 = dsub(Sxy, Syx);
    const double k11 = dsub(dsub(Sxx, Syy), Szz), k12 = dadd(Sxy, Syx), k13 = dadd(Szx, Sxz);
    const double k22 = dsub(dsub(Syy, Sxx), Szz), k23 = dadd(Syz, Szy);
    const double k33 = dsub(Szz, dadd(Sxx, Syy));
    const double d0 = det3s(k11, k12, k13, k12, k22, k23, k13, k23, k33);
    const double d1 = det3s(k01, k12, k13, k02, k22, k23, k03, k23, k33);
    const double d2 = det3s(k01, k11, k13, k02, k12, k23, k03, k13, k33);
    const double d3 = det3s(k01, k11, k12, k02, k12, k22, k03, k13, k23);
    const double c0 = dsub(dadd(dsub(dmul(k00, d0), dmul(k01, d1)), dmul(k02, d2)), dmul(k03, d3));
    const double e0 = dmul(0.5, dadd((double)Ga, (double)Gb));
    double lam = e0;
    for (int it = 0; it < 50; ++it) {
        const double old = lam;
        const double x2 = dmul(lam, lam);
        const double b = dmul(dadd(x2, c2), lam);
        const double a = dadd(b, c1);
        const double num = dadd(dmul(a, lam), c0);
        const double den = dadd(dadd(dmul(dmul(2.0, x2), lam), b), a);
        lam = dsub(lam, __ddiv_rn(num, den));
        if (fabs(dsub(lam, old)) < fabs(dmul(1e-11, lam))) break;
    }
    const double msd = fabs(__ddiv_rn(dmul(2.0, dsub(e0, lam)), (double)n_atoms));
    return __fsqrt_rn((float)msd);
}

// One atom into SIMD lane L of the 4 x 9 float32 accumulators (rounded multiply, rounded add)
#define RMSD_ATOM(L, x0, x1, x2, y0, y1, y2)                                              \
    acc[L][0] = __fadd_rn(acc[L][0], __fmul_rn(x0, y0));                                 \
    acc[L][1] = __fadd_rn(acc[L][1], __fmul_rn(x0, y1));                                 \
    acc[L][2] = __fadd_rn(acc[L][2], __fmul_rn(x0, y2));                                 \
    acc[L][3] = __fadd_rn(acc[L][3], __fmul_rn(x1, y0));                                 \
    acc[L][4] = __fadd_rn(acc[L][4], __fmul_rn(x1, y1));                                 \
    acc[L][5] = __fadd_rn(acc[L][5], __fmul_rn(x1, y2));                                 \
    acc[L][6] = __fadd_rn(acc[L][6], __fmul_rn(x2, y0));                                 \
    acc[L][7] = __fadd_rn(acc[L][7], __fmul_rn(x2, y1));                                 \
    acc[L][8] = __fadd_rn(acc[L][8], __fmul_rn(x2, y2));

// M of one frame pair, both operands addressed as float arrays of n_atoms * 3 (any address space)
template <bool VEC4>
__device__ __forceinline__ void inner_products_simd4(const float *__restrict__ x,
                                                     const float *__restrict__ y, int n_atoms,
                                                     float *M)
{
    float acc[4][9];
#pragma unroll
    for (int l = 0; l < 4; ++l)
#pragma unroll
        for (int q = 0; q < 9; ++q) acc[l][q] = 0.f;
    const int n_quads = n_atoms >> 2;
    for (int g = 0; g < n_quads; ++g) {
        float xs[12], ys[12];
        if (VEC4) {
            const float4 *x4 = reinterpret_cast<const float4 *>(x) + 3 * g;
            const float4 *y4 = reinterpret_cast<const float4 *>(y) + 3 * g;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 a = x4[c], b = y4[c];
                xs[4 * c] = a.x; xs[4 * c + 1] = a.y; xs[4 * c + 2] = a.z; xs[4 * c + 3] = a.w;
                ys[4 * c] = b.x; ys[4 * c + 1] = b.y; ys[4 * c + 2] = b.z; ys[4 * c + 3] = b.w;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 12; ++c) { xs[c] = x[12 * g + c]; ys[c] = y[12 * g + c]; }
        }
        RMSD_ATOM(0, xs[0], xs[1], xs[2], ys[0], ys[1], ys[2])
        RMSD_ATOM(1, xs[3], xs[4], xs[5], ys[3], ys[4], ys[5])
        RMSD_ATOM(2, xs[6], xs[7], xs[8], ys[6], ys[7], ys[8])
        RMSD_ATOM(3, xs[9], xs[10], xs[11], ys[9], ys[10], ys[11])
    }
    // the last n_atoms % 4 atoms (the SIMD code pads with zero atoms: adding +0 changes nothing)
    const int rem = n_atoms & 3;
    const float *xr = x + 12 * n_quads, *yr = y + 12 * n_quads;
    if (rem > 0) { RMSD_ATOM(0, xr[0], xr[1], xr[2], yr[0], yr[1], yr[2]) }
    if (rem > 1) { RMSD_ATOM(1, xr[3], xr[4], xr[5], yr[3], yr[4], yr[5]) }
    if (rem > 2) { RMSD_ATOM(2, xr[6], xr[7], xr[8], yr[6], yr[7], yr[8]) }
#pragma unroll
    for (int q = 0; q < 9; ++q)
        M[q] = __fadd_rn(__fadd_rn(acc[0][q], acc[1][q]), __fadd_rn(acc[2][q], acc[3][q]));
}

// frame stride of the shared-memory tile, in floats: odd number of 16-byte units (VEC4) or of
// words (scalar), so that the 32 lanes' loads of "element m of my frame" never share a bank
__host__ __device__ inline int tile_stride(int n_atoms, bool vec4)
{
    const int n3 = n_atoms * 3;
    if (vec4) {
        int u = n3 / 4;
        if ((u & 1) == 0) ++u;
        return 4 * u;
    }
    return n3 | 1;
}

// stage frame `src` (n3 floats in global memory) into this warp's tile slot; all lanes take part
template <bool VEC4>
__device__ __forceinline__ void tile_copy_frame(float *dst, const float *__restrict__ src, int n3, int lane)
{
    if (VEC4) {
        for (int c = lane; c < n3 / 4; c += 32)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;"
                         :: "r"((uint32_t)__cvta_generic_to_shared(dst + 4 * c)), "l"(src + 4 * c) : "memory");
    } else {
        for (int c = lane; c < n3; c += 32)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;"
                         :: "r"((uint32_t)__cvta_generic_to_shared(dst + c)), "l"(src + c) : "memory");
    }
}
// one frame = one TMA bulk copy (16-byte aligned, a multiple of 16 bytes), completion on an mbarrier
__device__ __forceinline__ void tile_bulk_frame(float *dst, const float *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
                    "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void tile_copy_wait()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncwarp();
}

// ---- K6c: k-centers pass with triangle-inequality pruning ---------------------------------------
// d(x, c_new) >= d(c_lab, c_new) - d(x, c_lab) for a metric, so a frame whose current centre c_lab
// satisfies  d(c_lab, c_new) >= 2 d(x, c_lab) + margin  cannot move to the new centre: the strict
// `dv < cur` of the pass (kcenters.py:93-95) is false for it.  The margin covers the rounding of
// the three computed distances (each is within 1.6e-3 sqrt(G / n) of the exact value: float32 M
// and traces), so the pruned pass writes exactly what the full pass writes.
// Scan: 16 bytes per frame (distance, label, trace); frames that survive go to a compact list.
static constexpr double kPruneAbs = 0.02;        // margin = kPruneAbs * sqrt((G_x + G_c) / (2 n)) + kPruneRel * d_cc
static constexpr double kPruneRel = 1e-5;

__global__ void __launch_bounds__(256)
rmsd_dcc_kernel(const unsigned char *__restrict__ slots, size_t slot_stride, int n_prev, int n_atoms,
                double *__restrict__ dcc)
{
    // thread j: distance between the new centre (slot n_prev) and centre j
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_prev) return;
    const int n3 = n_atoms * 3;
    const float *cn = reinterpret_cast<const float *>(slots + (size_t)n_prev * slot_stride + 16);
    const float *cj = reinterpret_cast<const float *>(slots + (size_t)j * slot_stride + 16);
    float M[9];
    inner_products_simd4<false>(cj, cn, n_atoms, M);
    dcc[j] = (double)qcp_rmsd_strict(M, cj[n3], cn[n3], n_atoms);
}

struct PassScratch {
    unsigned counter;        // last-block-done ticket of the compute kernel
    unsigned list_count;     // frames that survived the scan
    unsigned scan_blocks;    // block candidates written by the scan kernel
    unsigned pad;
};

__global__ void __launch_bounds__(256)
rmsd_scan_kernel(const double *__restrict__ dist, const int *__restrict__ labels,
                 const float *__restrict__ traces, long long n, int n_atoms,
                 const double *__restrict__ dcc, const unsigned char *__restrict__ slots,
                 size_t slot_stride, int n_prev, PassScratch *__restrict__ ps,
                 int *__restrict__ list, BlockCandR *__restrict__ block_cands)
{
    const int n3 = n_atoms * 3;
    const double Gc = (double)reinterpret_cast<const float *>(slots + (size_t)n_prev * slot_stride + 16)[n3];
    const int lane = threadIdx.x & 31;
    ArgMax best{-INFINITY, 0x7fffffffffffffffLL};
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long f0 = (long long)blockIdx.x * blockDim.x; f0 < n; f0 += stride) {
        const long long f = f0 + threadIdx.x;
        bool need = false;
        if (f < n) {
            const double d = dist[f];
            const double dc = dcc[labels[f]];
            const double R = sqrt(0.5 * ((double)traces[f] + Gc) / (double)n_atoms);
            need = !(dc - 2.0 * d >= kPruneAbs * R + kPruneRel * dc);
            if (!need && d > best.v) { best.v = d; best.i = f; }       // its distance stays: arg-max candidate
        }
        // warp-aggregated append (order inside the list does not matter: results are per frame)
        const unsigned m = __ballot_sync(0xffffffffu, need);
        if (m) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(&ps->list_count, (unsigned)__popc(m));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (need) list[base + __popc(m & ((1u << lane) - 1u))] = (int)f;
        }
    }
    __shared__ ArgMax s_warp[8];
    best = argmax_warp(best);
    if (lane == 0) s_warp[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        ArgMax b = threadIdx.x < 8 ? s_warp[threadIdx.x] : ArgMax{-INFINITY, 0x7fffffffffffffffLL};
        b = argmax_warp(b);
        if (threadIdx.x == 0) {
            block_cands[blockIdx.x].v = b.v;
            block_cands[blockIdx.x].i = b.i;
            if (blockIdx.x == 0) ps->scan_blocks = gridDim.x;
        }
    }
}

// The pass proper.  list == NULL: every frame (dense: tiles of 32 consecutive frames); otherwise the
// frames named by list[0 .. ps->list_count).  block_cands[0 .. scan_blocks) were written by the scan.
//
// Two kinds of warps (r2f ncu: with the Newton iteration behind the float32 sums in one warp, the
// single warp a scheduler could hold -- 38 KB of tile each -- issued 1 cycle in 4, waiting on its own
// dependent float64 chain):
//   tile warps (kTileWarps)   stage 32 frames with cp.async, every lane forms the float32 M of its
//                             frame and hands {M, G_x, frame} to its solver warp through a
//                             double-buffered record in shared memory (mbarrier full / empty);
//   solver warps (kSolvers per tile warp) no tile of their own: the double Newton iteration (a chain
//                             of ~400 dependent float64 operations, ~16k cycles per 32 frames: pure
//                             latency, r2h ncu), the strict running minimum and the arg-max, while
//                             the tile warp is already copying / summing the next frames.  Records
//                             go round robin to the tile warp's solvers.
struct SolveRecord { float M[9]; float Gx; long long f; };          // 48 bytes per lane
static constexpr int kSolvers = 4;                                  // solver warps per tile warp
static constexpr int kPassThreads = 32 * kTileWarps * (1 + kSolvers);

__device__ __forceinline__ void mbar_init_(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
                 :: "r"((uint32_t)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}

// a waiting solver warp must not compete for issue slots with the tile warp of its scheduler: poll,
// then sleep (r2j ncu: 136 M spins of the plain try_wait loop took 60 % of all stall samples and the
// four solvers per tile warp bought nothing)
__device__ __forceinline__ void mbar_wait_sleep_(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = (uint32_t)__cvta_generic_to_shared(bar);
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity), "r"(2000u) : "memory");
        if (ok) break;
        __nanosleep(500);
    }
}

template <bool VEC4>
__global__ void __launch_bounds__(kPassThreads)
rmsd_tile_pass_kernel(const float *__restrict__ xyz, const float *__restrict__ traces, long long n,
                      int n_atoms, const float *__restrict__ center, int label,
                      double *__restrict__ dist, int *__restrict__ labels, long long row_offset,
                      const int *__restrict__ list, PassScratch *__restrict__ ps, int scan_blocks_cap,
                      BlockCandR *__restrict__ block_cands, msmb200_candidate *__restrict__ out)
{
    extern __shared__ __align__(16) float s_tile[];
    __shared__ uint64_t s_full[kTileWarps][kSolvers], s_empty[kTileWarps][kSolvers];
    __shared__ uint64_t s_landed[kTileWarps];                  // the tile warp's frames have arrived (TMA bytes)
    __shared__ ArgMax s_warp[kTileWarps * kSolvers];
    __shared__ bool s_is_last;
    const int n3 = n_atoms * 3;
    const int stride = tile_stride(n_atoms, VEC4);
    float *s_center = s_tile;                                  // n3 floats (padded to 16 bytes)
    const int c_pad = (n3 + 3) & ~3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool solver = warp >= kTileWarps;
    const int pair = solver ? (warp - kTileWarps) % kTileWarps : warp;     // the tile warp it belongs to
    const int sk = solver ? (warp - kTileWarps) / kTileWarps : 0;           // which of its solvers
    float *my_tile = s_tile + c_pad + (size_t)pair * 32 * stride;
    SolveRecord *records = reinterpret_cast<SolveRecord *>(s_tile + c_pad + (size_t)kTileWarps * 32 * stride)
                           + (size_t)pair * kSolvers * 32;                 // [solver][lane]
    for (int j = threadIdx.x; j < n3; j += blockDim.x) s_center[j] = center[j];
    if (threadIdx.x < kTileWarps * kSolvers) {
        mbar_init_(&s_full[threadIdx.x / kSolvers][threadIdx.x % kSolvers], 1);
        mbar_init_(&s_empty[threadIdx.x / kSolvers][threadIdx.x % kSolvers], 1);
        if (threadIdx.x < kTileWarps) mbar_init_(&s_landed[threadIdx.x], 1);
    }
    __syncthreads();
    const float Gc = center[n3];
    const long long total = list ? (long long)ps->list_count : n;
    const long long n_chunks = (total + 31) / 32;
    ArgMax best{-INFINITY, 0x7fffffffffffffffLL};
    if (!solver) {
        // ------------------------------------------------ tile warp: copy, float32 sums, hand over
        long long it = 0;
        uint32_t landed_phase = 0;
        for (long long ch = (long long)blockIdx.x * kTileWarps + pair; ; ch += (long long)gridDim.x * kTileWarps) {
            const bool more = ch < n_chunks;
            long long f = -2;                                  // -2: no more chunks (every lane)
            float M[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) M[q] = 0.f;
            float Gx = 0.f;
            if (more) {
                const long long e = ch * 32 + lane;
                f = e < total ? (list ? (long long)list[e] : e) : -1;
                if (VEC4) {
                    // every lane fetches its own frame with ONE bulk copy (the per-chunk cp.async loop
                    // was half of the warp's instructions, and one warp per scheduler is issue bound)
                    const uint32_t bytes = (uint32_t)n3 * 4u;
                    const unsigned have = __ballot_sync(0xffffffffu, f >= 0);
                    if (lane == 0)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                     :: "r"((uint32_t)__cvta_generic_to_shared(&s_landed[pair])),
                                        "r"(bytes * (uint32_t)__popc(have)) : "memory");
                    __syncwarp();
                    if (f >= 0)
                        tile_bulk_frame(my_tile + (size_t)lane * stride, xyz + f * (long long)n3, bytes, &s_landed[pair]);
                    if (f >= 0) Gx = traces[f];
                    mbar_wait_(&s_landed[pair], landed_phase);
                    landed_phase ^= 1;
                } else {
                    for (int j = 0; j < 32; ++j) {
                        const long long fj = __shfl_sync(0xffffffffu, f, j);
                        if (fj < 0) break;
                        tile_copy_frame<VEC4>(my_tile + (size_t)j * stride, xyz + fj * (long long)n3, n3, lane);
                    }
                    if (f >= 0) Gx = traces[f];
                    tile_copy_wait();
                }
                if (f >= 0) inner_products_simd4<VEC4>(my_tile + (size_t)lane * stride, s_center, n_atoms, M);
            }
            // round robin over this warp's solvers; the end marker goes to every one of them
            for (int rep = 0; rep < (more ? 1 : kSolvers); ++rep, ++it) {
                const int k = (int)(it % kSolvers);
                mbar_wait_(&s_empty[pair][k], (uint32_t)(((it / kSolvers) & 1) ^ 1));
                SolveRecord &r = records[k * 32 + lane];
#pragma unroll
                for (int q = 0; q < 9; ++q) r.M[q] = M[q];
                r.Gx = Gx;
                r.f = f;
                __syncwarp();
                if (lane == 0) mbar_arrive_(&s_full[pair][k]);
            }
            if (!more) break;
        }
    } else {
        // ------------------------------------------------ solver warp: Newton, minimum, arg-max
        for (long long it = 0; ; ++it) {
            mbar_wait_sleep_(&s_full[pair][sk], (uint32_t)(it & 1));
            const SolveRecord r = records[sk * 32 + lane];
            __syncwarp();
            if (lane == 0) mbar_arrive_(&s_empty[pair][sk]);
            if (r.f == -2) break;
            if (r.f >= 0) {
                double cur = __ldcg(dist + r.f);
                const double dv = (double)qcp_rmsd_strict(r.M, r.Gx, Gc, n_atoms);
                if (dv < cur) {
                    cur = dv;
                    dist[r.f] = dv;
                    labels[r.f] = label;
                }
                if (cur > best.v || (cur == best.v && r.f < best.i)) { best.v = cur; best.i = r.f; }
            }
        }
        best = argmax_warp(best);
        if (lane == 0) s_warp[warp - kTileWarps] = best;
    }
    __syncthreads();
    BlockCandR *mine = block_cands + scan_blocks_cap;          // the scan kernel's slots come first
    if (threadIdx.x == 0) {
        ArgMax b = s_warp[0];
        for (int w = 1; w < kTileWarps * kSolvers; ++w) b = argmax_merge(b, s_warp[w]);
        mine[blockIdx.x].v = b.v;
        mine[blockIdx.x].i = b.i;
        __threadfence();
        const unsigned ticket = atomicInc(&ps->counter, gridDim.x - 1);
        s_is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_is_last) return;
    __threadfence();
    ArgMax w{-INFINITY, 0x7fffffffffffffffLL};
    const int n_scan = list ? (int)ps->scan_blocks : 0;
    for (int b = threadIdx.x; b < n_scan + (int)gridDim.x; b += blockDim.x) {
        const BlockCandR *src = b < n_scan ? block_cands + b : mine + (b - n_scan);
        ArgMax c;
        c.v = __ldcg(&src->v);
        c.i = __ldcg(&src->i);
        w = argmax_merge(w, c);
    }
    w = argmax_warp(w);
    __syncthreads();
    __shared__ ArgMax s_fin[kTileWarps * (1 + kSolvers)];
    if (lane == 0) s_fin[warp] = w;
    __syncthreads();
    if (threadIdx.x == 0) {
        ArgMax b = s_fin[0];
        for (int q = 1; q < kTileWarps * (1 + kSolvers); ++q) b = argmax_merge(b, s_fin[q]);
        if (b.i == 0x7fffffffffffffffLL) b.i = 0;
        s_warp[0] = b;
        out->value = b.v;
        out->index = row_offset + b.i;
        ps->list_count = 0;                                    // ready for the next pass's scan
        ps->scan_blocks = 0;
    }
    __syncthreads();
    w = s_warp[0];
    float *payload = reinterpret_cast<float *>(out + 1);
    if (n > 0) {
        for (int j = threadIdx.x; j < n3; j += blockDim.x) payload[j] = xyz[w.i * (long long)n3 + j];
        if (threadIdx.x == 0) payload[n3] = traces[w.i];
    }
}

// ---- K6d: assign_nearest / dist on the same tiles (one thread per frame, loop over the centres) ---
// MODE 0: assign (k centres Y, labels + optional minimum + block sums for the inertia); MODE 1: dist.
template <int MODE, bool VEC4>
__global__ void __launch_bounds__(kTileThreads)
rmsd_tile_multi_kernel(const float *__restrict__ xyz, const float *__restrict__ traces,
                       long long n_items, int n_atoms, const float *__restrict__ Y,
                       const float *__restrict__ Y_traces, int k, const long long *__restrict__ rows,
                       int *__restrict__ labels, double *__restrict__ out,
                       double *__restrict__ block_sums)
{
    extern __shared__ __align__(16) float s_tile[];
    const int n3 = n_atoms * 3;
    const int stride = tile_stride(n_atoms, VEC4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float *my_tile = s_tile + (size_t)warp * 32 * stride;
    const long long n_chunks = (n_items + 31) / 32;
    double local = 0.0;
    for (long long ch = (long long)blockIdx.x * kTileWarps + warp; ch < n_chunks;
         ch += (long long)gridDim.x * kTileWarps) {
        const long long e = ch * 32 + lane;
        const bool valid = e < n_items;
        const long long f = valid ? (rows ? rows[e] : e) : -1;
        for (int j = 0; j < 32; ++j) {
            const long long fj = __shfl_sync(0xffffffffu, f, j);
            if (fj < 0) break;
            tile_copy_frame<VEC4>(my_tile + (size_t)j * stride, xyz + fj * (long long)n3, n3, lane);
        }
        const float Gx = valid ? traces[f] : 0.f;
        tile_copy_wait();
        if (valid) {
            const float *x = my_tile + (size_t)lane * stride;
            if (MODE == 0) {
                float bestd = 3.402823466e+38f;   // FLT_MAX, libdistance.pyx:347
                int arg = 0;
                for (int j = 0; j < k; ++j) {
                    float M[9];
                    inner_products_simd4<VEC4>(x, Y + (long long)j * n3, n_atoms, M);
                    const float dv = qcp_rmsd_strict(M, Gx, Y_traces[j], n_atoms);
                    if (dv < bestd) { bestd = dv; arg = j; }
                }
                labels[e] = arg;
                if (out) out[e] = (double)bestd;
                local += (double)bestd;
            } else {
                float M[9];
                inner_products_simd4<VEC4>(x, Y, n_atoms, M);
                out[e] = (double)qcp_rmsd_strict(M, Gx, Y_traces[0], n_atoms);
            }
        }
        __syncwarp();
    }
    if (MODE == 0) {
        __shared__ double s_sum[kTileWarps];
        for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
        if (lane == 0) s_sum[warp] = local;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < kTileWarps; ++w) t += s_sum[w];
            block_sums[blockIdx.x] = t;
        }
    }
}

// ---- pdist: one thread per condensed pair (m <= k + batch rows: small) ---------------------------
__global__ void __launch_bounds__(128)
rmsd_pdist_strict_kernel(const float *__restrict__ xyz, const float *__restrict__ traces, int n_atoms,
                         const long long *__restrict__ rows, long long m, long long pairs,
                         double *__restrict__ out)
{
    const long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (it >= pairs) return;
    const double mm = (double)m;
    long long i = (long long)floor(((2.0 * mm - 1.0) - sqrt((2.0 * mm - 1.0) * (2.0 * mm - 1.0) - 8.0 * (double)it)) * 0.5);
    if (i < 0) i = 0;
    while (i > 0 && m * i - i * (i + 1) / 2 > it) --i;
    while (m * (i + 1) - (i + 1) * (i + 2) / 2 <= it) ++i;
    const long long j = it - (m * i - i * (i + 1) / 2) + i + 1;
    const long long ra = rows ? rows[i] : i, rb = rows ? rows[j] : j;
    const long long fs = (long long)n_atoms * 3;
    float M[9];
    inner_products_simd4<false>(xyz + ra * fs, xyz + rb * fs, n_atoms, M);
    out[it] = (double)qcp_rmsd_strict(M, traces[ra], traces[rb], n_atoms);
}

static inline bool rmsd_tile_engine(int n_atoms)
{
    static int forced = -1;
    if (forced < 0) {
        const char *v = getenv("MSMB200_RMSD_F64");
        forced = (v && atoi(v) != 0) ? 1 : 0;
    }
    return !forced && n_atoms <= kTileMaxAtoms;
}
static inline size_t tile_smem_bytes(int n_atoms, bool vec4, bool with_center)
{
    const size_t c_pad = with_center ? (size_t)((n_atoms * 3 + 3) & ~3) : 0;
    return sizeof(float) * (c_pad + (size_t)kTileWarps * 32 * tile_stride(n_atoms, vec4));
}
// persistent grid: as many blocks as fit the SMs at once (shared memory decides)
static inline int tile_grid(long long items, size_t smem)
{
    long long blocks = (items + 32 * kTileWarps - 1) / (32 * kTileWarps);
    long long per_sm = (long long)((220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long cap = (long long)sm_count() * per_sm;
    if (cap > 2 * 1024) cap = 2 * 1024;          // block candidate slots of the pass workspace
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

static inline int warp_grid(long long items)
{
    long long blocks = (items + (kRThreads / 32) - 1) / (kRThreads / 32);
    long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

// workspace of the tiled pass: [PassScratch | block candidates (scan + pass) | dcc | list]
static constexpr int kScanBlocksCap = 1024;
static size_t pass_ws_bytes(int64_t n, int max_centres)
{
    return 64 + sizeof(BlockCandR) * (size_t)(kScanBlocksCap + 2 * 1024) + sizeof(double) * (size_t)(max_centres + 1)
           + sizeof(int) * (size_t)(n > 0 ? n : 0) + 256;
}

// dense (slots == NULL or n_prev == 0) or pruned k-centers pass in the reference-class arithmetic
static int rmsd_pass_tiles(const float *xyz, const float *traces, int64_t n, int n_atoms,
                           const float *center, const unsigned char *slots, size_t slot_stride,
                           int n_prev, int label, double *distances, int32_t *labels,
                           int64_t row_offset, msmb200_candidate *out, void *workspace,
                           size_t workspace_bytes, cudaStream_t st)
{
    const bool prune = slots != nullptr && n_prev > 0;
    const size_t need = prune ? pass_ws_bytes(n, n_prev) : 64 + sizeof(BlockCandR) * (size_t)(kScanBlocksCap + 2 * 1024);
    MSMB_REQUIRE(workspace_bytes >= need, "rmsd_kcenters_pass: workspace too small (%zu < %zu)",
                 workspace_bytes, need);
    unsigned char *w = reinterpret_cast<unsigned char *>(workspace);
    PassScratch *ps = reinterpret_cast<PassScratch *>(w);
    BlockCandR *cands = reinterpret_cast<BlockCandR *>(w + 64);
    double *dcc = reinterpret_cast<double *>(w + 64 + sizeof(BlockCandR) * (size_t)(kScanBlocksCap + 2 * 1024));
    int *list = reinterpret_cast<int *>(reinterpret_cast<unsigned char *>(dcc) + sizeof(double) * (size_t)(n_prev + 1));
    const bool vec4 = (n_atoms & 3) == 0 && (reinterpret_cast<uintptr_t>(xyz) & 15u) == 0;
    const size_t smem = tile_smem_bytes(n_atoms, vec4, true) + sizeof(SolveRecord) * (size_t)kTileWarps * kSolvers * 32;
    // function attributes belong to a device: the threads of a multi-GPU fit each set their own
    static bool attr_done[64][2] = {};
    int dev = 0;
    MSMB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev][vec4 ? 1 : 0]) {
        if (vec4) MSMB_CUDA(cudaFuncSetAttribute(rmsd_tile_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        else MSMB_CUDA(cudaFuncSetAttribute(rmsd_tile_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        if (dev >= 0 && dev < 64) attr_done[dev][vec4 ? 1 : 0] = true;
    }
    const int *d_list = nullptr;
    if (prune) {
        MSMB_REQUIRE(n <= 0x7fffffffLL, "rmsd_kcenters_pass: shard too long for the pruned pass");
        rmsd_dcc_kernel<<<(n_prev + 255) / 256, 256, 0, st>>>(slots, slot_stride, n_prev, n_atoms, dcc);
        MSMB_LAUNCH_CHECK();
        long long sb = (n + 255) / 256;
        if (sb > kScanBlocksCap) sb = kScanBlocksCap;
        if (sb > (long long)sm_count() * 4) sb = (long long)sm_count() * 4;
        if (sb < 1) sb = 1;
        rmsd_scan_kernel<<<(unsigned)sb, 256, 0, st>>>(distances, labels, traces, n, n_atoms, dcc, slots,
                                                       slot_stride, n_prev, ps, list, cands);
        MSMB_LAUNCH_CHECK();
        d_list = list;
    }
    const int grid = tile_grid(n, smem);
    if (vec4)
        rmsd_tile_pass_kernel<true><<<grid, kPassThreads, smem, st>>>(
            xyz, traces, n, n_atoms, center, label, distances, labels, row_offset, d_list, ps,
            kScanBlocksCap, cands, out);
    else
        rmsd_tile_pass_kernel<false><<<grid, kPassThreads, smem, st>>>(
            xyz, traces, n, n_atoms, center, label, distances, labels, row_offset, d_list, ps,
            kScanBlocksCap, cands, out);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

}  // namespace msmb

using namespace msmb;

extern "C" size_t msmb200_rmsd_pass_workspace_bytes(int64_t n, int32_t max_centres)
{
    return pass_ws_bytes(n, max_centres > 0 ? max_centres : 0);
}

extern "C" int msmb200_rmsd_kcenters_pass_pruned(const float *xyz, const float *traces, int64_t n,
                                                 int n_atoms, const void *centre_slots,
                                                 size_t slot_stride, int32_t center_label,
                                                 double *distances, int32_t *labels,
                                                 int64_t row_offset, msmb200_candidate *out,
                                                 void *workspace, size_t workspace_bytes, void *stream)
{
    MSMB_REQUIRE(xyz && traces && centre_slots && distances && labels && out && workspace && n >= 0 &&
                 n_atoms > 0 && center_label >= 0 && slot_stride >= 16 + sizeof(float) * ((size_t)n_atoms * 3 + 1),
                 "rmsd_kcenters_pass_pruned: bad args");
    const unsigned char *slots = reinterpret_cast<const unsigned char *>(centre_slots);
    const float *center = reinterpret_cast<const float *>(slots + (size_t)center_label * slot_stride + 16);
    if (!rmsd_tile_engine(n_atoms))     // wide frames / float64 engine: the plain pass
        return msmb200_rmsd_kcenters_pass(xyz, traces, n, n_atoms, center, center_label, distances,
                                          labels, row_offset, out, workspace, workspace_bytes, stream);
    return rmsd_pass_tiles(xyz, traces, n, n_atoms, center, slots, slot_stride, center_label,
                           center_label, distances, labels, row_offset, out, workspace,
                           workspace_bytes, (cudaStream_t)stream);
}

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
    if (rmsd_tile_engine(n_atoms))
        return rmsd_pass_tiles(xyz, traces, n, n_atoms, center, nullptr, 0, 0, center_label, distances,
                               labels, row_offset, out, workspace, workspace_bytes, (cudaStream_t)stream);
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
    const bool tiles = rmsd_tile_engine(n_atoms);
    const bool vec4 = (n_atoms & 3) == 0 && (reinterpret_cast<uintptr_t>(xyz) & 15u) == 0 &&
                      (reinterpret_cast<uintptr_t>(Y) & 15u) == 0;
    const size_t smem = tile_smem_bytes(n_atoms, vec4, false);
    const int grid = tiles ? tile_grid(n_out, smem) : warp_grid(n_out);
    double *partials = nullptr;
    MSMB_CUDA(cudaMallocAsync(&partials, sizeof(double) * grid, st));
    if (tiles) {
        if (vec4) {
            MSMB_CUDA(cudaFuncSetAttribute(rmsd_tile_multi_kernel<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            rmsd_tile_multi_kernel<0, true><<<grid, kTileThreads, smem, st>>>(
                xyz, traces, n_out, n_atoms, Y, Y_traces, k, (const long long *)rows, labels, min_dist, partials);
        } else {
            MSMB_CUDA(cudaFuncSetAttribute(rmsd_tile_multi_kernel<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            rmsd_tile_multi_kernel<0, false><<<grid, kTileThreads, smem, st>>>(
                xyz, traces, n_out, n_atoms, Y, Y_traces, k, (const long long *)rows, labels, min_dist, partials);
        }
    } else {
        rmsd_multi_kernel<0><<<grid, kRThreads, 0, st>>>(xyz, traces, n_out, n_atoms, Y, Y_traces, k,
                                                        (const long long *)rows, 0, labels, min_dist,
                                                        partials);
    }
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
    if (rmsd_tile_engine(n_atoms)) {
        const bool vec4 = (n_atoms & 3) == 0 && (reinterpret_cast<uintptr_t>(xyz) & 15u) == 0 &&
                          (reinterpret_cast<uintptr_t>(y) & 15u) == 0;
        const size_t smem = tile_smem_bytes(n_atoms, vec4, false);
        if (vec4) {
            MSMB_CUDA(cudaFuncSetAttribute(rmsd_tile_multi_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            rmsd_tile_multi_kernel<1, true><<<tile_grid(n_out, smem), kTileThreads, smem, st>>>(
                xyz, traces, n_out, n_atoms, y, d_trace, 1, (const long long *)rows, nullptr, out, nullptr);
        } else {
            MSMB_CUDA(cudaFuncSetAttribute(rmsd_tile_multi_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            rmsd_tile_multi_kernel<1, false><<<tile_grid(n_out, smem), kTileThreads, smem, st>>>(
                xyz, traces, n_out, n_atoms, y, d_trace, 1, (const long long *)rows, nullptr, out, nullptr);
        }
    } else {
        rmsd_multi_kernel<1><<<warp_grid(n_out), kRThreads, 0, st>>>(
            xyz, traces, n_out, n_atoms, y, d_trace, 1, (const long long *)rows, 0, nullptr, out,
            nullptr);
    }
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
    if (rmsd_tile_engine(n_atoms))
        rmsd_pdist_strict_kernel<<<(unsigned)((pairs + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
            xyz, traces, n_atoms, (const long long *)rows, m, pairs, out);
    else
        rmsd_multi_kernel<2><<<warp_grid(pairs), kRThreads, 0, (cudaStream_t)stream>>>(
            xyz, traces, pairs, n_atoms, nullptr, nullptr, 0, (const long long *)rows, m, nullptr,
            out, nullptr);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}
