// common.cuh -- shared device/host helpers for libmsmb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <type_traits>
#include "../../include/msmb200.h"

namespace msmb {

// ---- error plumbing (C ABI never throws; thread-local last-error text) ----
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MSMB_CUDA(call)                                                        \
    do {                                                                       \
        cudaError_t e__ = (call);                                              \
        if (e__ != cudaSuccess)                                                \
            return ::msmb::cuda_fail(e__, #call, __FILE__, __LINE__);          \
    } while (0)

#define MSMB_REQUIRE(cond, ...)                                                \
    do {                                                                       \
        if (!(cond)) {                                                         \
            ::msmb::set_error(__VA_ARGS__);                                    \
            return MSMB200_E_INVALID;                                          \
        }                                                                      \
    } while (0)

void count_launch();
// every kernel launch of the library is followed by this (also feeds msmb200_launch_count)
#define MSMB_LAUNCH_CHECK()                                                    \
    do {                                                                       \
        ::msmb::count_launch();                                                \
        MSMB_CUDA(cudaGetLastError());                                         \
    } while (0)

int sm_count();   // of the current device (cached per device)

// ---- small device utilities --------------------------------------------------
__device__ __forceinline__ double shfl_xor_f64(double v, int mask, int width = 32)
{
    return __shfl_xor_sync(0xffffffffu, v, mask, width);
}

// Sum / max of a double over the `G` consecutive lanes of a sub-warp group
// (G a power of two <= 32).  Every lane of the group receives the result.
template <bool IS_MAX>
__device__ __forceinline__ double group_combine(double v, int G)
{
    for (int off = G >> 1; off > 0; off >>= 1) {
        double o = __shfl_xor_sync(0xffffffffu, v, off);
        v = IS_MAX ? fmax(v, o) : v + o;
    }
    return v;
}

// R sums (maxima) over the G lanes of a group at once.  While the lanes fold (xor offsets
// G/2, G/4, ...) they also split the R values between them, so level l moves R/2^l doubles
// instead of R; after log2(R) levels every lane carries ONE value, finished by a plain
// butterfly.  Lane `lig` returns the total of value j = sum over the first log2(R) levels of
// (bit `off` of lig set ? m/2 : 0); the G/R lanes that differ only in lower bits all hold it.
// The additions pair up exactly like group_combine's, so the results are bit-identical to it.
// Needs R a power of two, G >= R.
template <bool IS_MAX, int R>
__device__ __forceinline__ double group_reduce_split(double (&v)[R], int G, int lig)
{
    int off = G >> 1;
#pragma unroll
    for (int m = R; m > 1; m >>= 1, off >>= 1) {
        const bool up = (lig & off) != 0;
#pragma unroll
        for (int i = 0; i < m / 2; ++i) {
            const double send = up ? v[i] : v[i + m / 2];
            const double keep = up ? v[i + m / 2] : v[i];
            const double o = __shfl_xor_sync(0xffffffffu, send, off);
            v[i] = IS_MAX ? fmax(keep, o) : keep + o;
        }
    }
    double r = v[0];
    for (; off > 0; off >>= 1) {
        const double o = __shfl_xor_sync(0xffffffffu, r, off);
        r = IS_MAX ? fmax(r, o) : r + o;
    }
    return r;
}

// (value, index) arg-max with lowest-index tie-break == np.argmax first-max.
struct ArgMax {
    double v;
    long long i;
};
__device__ __forceinline__ ArgMax argmax_merge(ArgMax a, ArgMax b)
{
    // a wins unless b is strictly larger, or equal with a lower index
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}
__device__ __forceinline__ ArgMax argmax_warp(ArgMax a)
{
    for (int off = 16; off > 0; off >>= 1) {
        ArgMax o;
        o.v = __shfl_xor_sync(0xffffffffu, a.v, off);
        o.i = __shfl_xor_sync(0xffffffffu, a.i, off);
        a = argmax_merge(a, o);
    }
    return a;
}

// ---- the eight vector metrics -------------------------------------------------
// Arithmetic contract (msmbuilder/libdistance/src/distance_kernels.h:41-242):
// the element difference (and, for canberra/braycurtis, the element sum) is
// formed in the INPUT precision T, widened to double, and every accumulation is
// double.  acc() folds one element pair into the two running doubles (a, b);
// fin() turns them into the distance.
template <int METRIC, typename T>
struct Metric {
    static constexpr bool kIsMax = (METRIC == MSMB200_CHEBYSHEV);
    static constexpr bool kTwoAcc = (METRIC == MSMB200_BRAYCURTIS || METRIC == MSMB200_JACCARD);

    __device__ __forceinline__ static void acc(double &a, double &b, T x, T c)
    {
        if (METRIC == MSMB200_EUCLIDEAN || METRIC == MSMB200_SQEUCLIDEAN) {
            T df = x - c;
            double d = (double)df;
            // float input: d*d is exact in double, so fma == mul-then-add.
            // double input: keep the reference's two roundings (no contraction).
            if (sizeof(T) == 4) a = fma(d, d, a);
            else a = __dadd_rn(a, __dmul_rn(d, d));
        } else if (METRIC == MSMB200_CITYBLOCK) {
            T df = x - c;
            a += fabs((double)df);
        } else if (METRIC == MSMB200_CHEBYSHEV) {
            T df = x - c;
            a = fmax(a, fabs((double)df));
        } else if (METRIC == MSMB200_CANBERRA) {
            T df = x - c;
            T dn = (x < 0 ? -x : x) + (c < 0 ? -c : c);   // T arithmetic
            double den = (double)dn;
            if (den > 0.0) a += fabs((double)df) / den;
        } else if (METRIC == MSMB200_BRAYCURTIS) {
            T df = x - c;
            T sm = x + c;
            a += fabs((double)df);
            b += fabs((double)sm);
        } else if (METRIC == MSMB200_HAMMING) {
            a += (x != c) ? 1.0 : 0.0;
        } else if (METRIC == MSMB200_JACCARD) {
            bool nz = (x != (T)0) || (c != (T)0);
            a += ((x != c) && nz) ? 1.0 : 0.0;
            b += nz ? 1.0 : 0.0;
        }
    }
    __device__ __forceinline__ static double fin(double a, double b, int n)
    {
        if (METRIC == MSMB200_EUCLIDEAN) return sqrt(a);
        if (METRIC == MSMB200_BRAYCURTIS || METRIC == MSMB200_JACCARD) return a / b;
        if (METRIC == MSMB200_HAMMING) return a / (double)n;
        return a;
    }
};

// Vector width for 16-byte loads.
template <typename T> struct Vec;
template <> struct Vec<float> { typedef float4 type; static constexpr int N = 4; };
template <> struct Vec<double> { typedef double2 type; static constexpr int N = 2; };

__device__ __forceinline__ void unpack(const float4 &v, float (&e)[4])
{
    e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
}
__device__ __forceinline__ void unpack(const double2 &v, double (&e)[2])
{
    e[0] = v.x; e[1] = v.y;
}

// Streaming (read-once) 16-byte global load that does not pollute L1.
__device__ __forceinline__ float4 ldg_stream(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ double2 ldg_stream(const double2 *p)
{
    double2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                 : "=d"(r.x), "=d"(r.y) : "l"(p));
    return r;
}

// Distance between row u (global) and row c (shared or global) cooperatively by
// the G lanes of a group; every lane returns the full value.
template <int METRIC, typename T, bool VEC, bool STREAM = true>
__device__ __forceinline__ double group_distance(const T *__restrict__ u,
                                                 const T *__restrict__ c, int d,
                                                 int lane_in_group, int G)
{
    double a = 0.0, b = 0.0;
    if (VEC) {
        typedef typename Vec<T>::type V;
        constexpr int N = Vec<T>::N;
        const V *u4 = reinterpret_cast<const V *>(u);
        const V *c4 = reinterpret_cast<const V *>(c);
        const int dv = d / N;
#pragma unroll 2
        for (int j = lane_in_group; j < dv; j += G) {
            V xv = STREAM ? ldg_stream(u4 + j) : u4[j];
            V cv = c4[j];
            T xe[N], ce[N];
            unpack(xv, xe);
            unpack(cv, ce);
#pragma unroll
            for (int e = 0; e < N; ++e) Metric<METRIC, T>::acc(a, b, xe[e], ce[e]);
        }
    } else {
        for (int j = lane_in_group; j < d; j += G) Metric<METRIC, T>::acc(a, b, u[j], c[j]);
    }
    a = group_combine<Metric<METRIC, T>::kIsMax>(a, G);
    if (Metric<METRIC, T>::kTwoAcc) b = group_combine<false>(b, G);
    return Metric<METRIC, T>::fin(a, b, d);
}

// Lanes per row: smallest power of two covering the row's 16-byte vectors, <= 32.
inline int lanes_per_row(int d, int elems_per_vec, bool vec)
{
    int units = vec ? d / elems_per_vec : d;
    int g = 1;
    while (g < units && g < 32) g <<= 1;
    return g;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Compile-time dispatch over (dtype, metric): F must be a generic lambda taking
// (T tag, std::integral_constant<int, METRIC>).
template <typename T, typename F>
inline int dispatch_metric(int metric, F &&f)
{
    switch (metric) {
    case MSMB200_EUCLIDEAN:   return f(T(), std::integral_constant<int, MSMB200_EUCLIDEAN>());
    case MSMB200_SQEUCLIDEAN: return f(T(), std::integral_constant<int, MSMB200_SQEUCLIDEAN>());
    case MSMB200_CITYBLOCK:   return f(T(), std::integral_constant<int, MSMB200_CITYBLOCK>());
    case MSMB200_CHEBYSHEV:   return f(T(), std::integral_constant<int, MSMB200_CHEBYSHEV>());
    case MSMB200_CANBERRA:    return f(T(), std::integral_constant<int, MSMB200_CANBERRA>());
    case MSMB200_BRAYCURTIS:  return f(T(), std::integral_constant<int, MSMB200_BRAYCURTIS>());
    case MSMB200_HAMMING:     return f(T(), std::integral_constant<int, MSMB200_HAMMING>());
    case MSMB200_JACCARD:     return f(T(), std::integral_constant<int, MSMB200_JACCARD>());
    default:
        set_error("unknown metric id %d", metric);
        return MSMB200_E_INVALID;
    }
}
template <typename F>
inline int dispatch(int dtype, int metric, F &&f)
{
    if (dtype == MSMB200_F32) return dispatch_metric<float>(metric, f);
    if (dtype == MSMB200_F64) return dispatch_metric<double>(metric, f);
    set_error("unknown dtype id %d", dtype);
    return MSMB200_E_INVALID;
}

}  // namespace msmb
