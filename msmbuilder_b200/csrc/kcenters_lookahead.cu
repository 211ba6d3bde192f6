// kcenters_lookahead.cu -- K2b: Gonzalez k-centers with look-ahead (sm_100a).
//
// The reference (kcenters.py:91-97) makes ONE pass over all frames per centre because centre
// i+1 = argmax_x min_{j<=i} d(x, c_j) is only known after pass i.  A pass is HBM bound
// (4*D + 8 bytes per frame), so k centres cost k reads of the data set.
//
// Look-ahead: while a pass streams the frames it also keeps, per owning lane, the THREE largest
// running minima it has seen (the two largest with their rows).  After the pass
//   * the CANDIDATES are the lanes' two largest values above the T-th largest of all of them (plus
//     the true arg-max), with their frames gathered;
//   * every frame that is NOT a candidate has a running minimum <= tau = max(largest lane third
//     best, T-th largest entry), and running minima only ever decrease.
// (Round 2 kept two values and one candidate per lane: with 37 888 lanes two of the ~250 largest
// frames share a lane -- the birthday bound -- so tau was the 221st largest value of the bench data
// and the first chain stopped after 4 of the 7 centres it could have certified,
// profiles/r2u_lookahead_diag_top2.log.  Three of the top frames in one lane need ~2000 of them.)
// A chain kernel then plays the reference's loop on the candidates alone: pick the arg-max
// (lowest index on ties, np.argmax), lower the candidates' minima by their distance to it, pick
// again ...  A pick whose value is strictly above tau is provably the arg-max over ALL frames,
// i.e. exactly the centre the reference would choose; the chain stops at the first pick it cannot
// certify.  The J centres found this way are applied by ONE fused pass (strict '<' in centre
// order == the reference's k sequential updates), so k centres cost (number of chains + 1) reads
// instead of k.  With nothing to certify it degrades to the reference's one centre per pass.
//
// The fused pass evaluates the J distances of a frame in float32 first (one FSUB + FFMA per
// element); a centre that provably cannot beat the frame's current minimum (relative margin
// eps = 4 (d + 8) 2^-24 on the squared distance) is skipped, every other (frame, centre) pair is
// recomputed with the reference arithmetic (float difference, float64 square-accumulate, sqrt:
// distance_kernels.h:54-77) in exactly the summation order of kcenters_pass_fast_kernel, so
// distances_, labels_ and the centre ids are those of the pass-per-centre path bit for bit.
//
// float32 frames, euclidean / sqeuclidean, 16-byte aligned rows with (d/4) % G == 0; everything
// else keeps the pass-per-centre path (msmb200_kcenters_pass).
#include "common.cuh"
#include <stdlib.h>

namespace msmb {

static constexpr int kThreads = 256;
static constexpr int kStages = 3;            // per-warp frame ring of the fused passes (4 KB per stage)
static constexpr long long kNoRow = 0x7fffffffffffffffLL;

struct LaneCand {           // one per (group, frame slot) of a pass
    double v1;              // largest running minimum seen (-inf: no rows)
    long long i1;           // its GLOBAL row (lowest row among equals)
    double v2;              // second largest (ties with v1 included)
    long long i2;           // its GLOBAL row
    double v3;              // third largest (ties included): bounds every other frame of the lane
    long long pad;
};
// The owning lane's record lives in SHARED memory while a pass runs (ten registers otherwise: the first
// pass then either spills or drops to three resident blocks per SM, and HBM latency needs four -- r2v:
// 8.0 -> 8.3 ms).  The hot path is one 8-byte load and a compare; an update is rare once the lane has
// seen a few frames.  A lane's rows increase from iteration to iteration and the comparisons are
// strict: the first row wins ties.  Every record has exactly one owner: no synchronisation.
struct LaneTop {
    LaneCand *p;
    __device__ __forceinline__ void init(LaneCand *slot, bool owner)
    {
        p = slot;
        if (owner) {
            p->v1 = p->v2 = p->v3 = -INFINITY;
            p->i1 = p->i2 = kNoRow;
            p->pad = 0;
        }
    }
    __device__ __forceinline__ void add(double cur, long long row)
    {
        if (cur > p->v3) {
            const double a1 = p->v1, a2 = p->v2;
            if (cur > a1) { p->v3 = a2; p->v2 = a1; p->i2 = p->i1; p->v1 = cur; p->i1 = row; }
            else if (cur > a2) { p->v3 = a2; p->v2 = cur; p->i2 = row; }
            else p->v3 = cur;
        }
    }
    __device__ __forceinline__ void store(LaneCand *out, long long row_offset) const
    {
        LaneCand lc = *p;
        if (lc.i1 != kNoRow) lc.i1 += row_offset;
        if (lc.i2 != kNoRow) lc.i2 += row_offset;
        *out = lc;
    }
};
// shared-memory bytes of the lane records of one block: one per owner = (frame slot of a group)
static inline size_t lane_top_bytes(int G, int R) { return sizeof(LaneCand) * (size_t)(kThreads / G) * R; }
struct LaneHeader {         // first 32 bytes of the lane buffer
    long long n_slots;
    long long pad[3];
};
struct SetHeader {          // candidate set of one rank: header | double val[cap] | int64 idx[cap] | float rows[cap][d]
    int count;
    int cap;
    double tau;
    long long pad[2];
};
struct CentersHeader {      // pending centres: header | int64 ids[cap] | float rows[cap][d]
    int n;
    int cap;
    long long pad[3];
};

static inline size_t set_bytes(int d, int cap)
{
    return sizeof(SetHeader) + (size_t)cap * 16 + (size_t)cap * d * sizeof(float);
}
static inline size_t centers_bytes(int d, int cap)
{
    return sizeof(CentersHeader) + (size_t)cap * 8 + (size_t)cap * d * sizeof(float);
}
static inline int pass_grid() { return sm_count() * 8; }

__device__ __forceinline__ double *set_val(unsigned char *s) { return reinterpret_cast<double *>(s + sizeof(SetHeader)); }
__device__ __forceinline__ long long *set_idx(unsigned char *s, int cap)
{
    return reinterpret_cast<long long *>(s + sizeof(SetHeader) + (size_t)cap * 8);
}
__device__ __forceinline__ float *set_rows(unsigned char *s, int cap)
{
    return reinterpret_cast<float *>(s + sizeof(SetHeader) + (size_t)cap * 16);
}

// float flavour of group_reduce_split (common.cuh): V sums over the G lanes of a group at once;
// lane `lig` ends up with the total of value lig / (G / V)
template <int V>
__device__ __forceinline__ float group_reduce_split_f32(float (&v)[V], int G, int lig)
{
    int off = G >> 1;
#pragma unroll
    for (int m = V; m > 1; m >>= 1, off >>= 1) {
        const bool up = (lig & off) != 0;
#pragma unroll
        for (int i = 0; i < m / 2; ++i) {
            const float send = up ? v[i] : v[i + m / 2];
            const float keep = up ? v[i + m / 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float r = v[0];
    for (; off > 0; off >>= 1) r += __shfl_xor_sync(0xffffffffu, r, off);
    return r;
}

// ---------------------------------------------------------------------------------------
// The very first pass (one centre, every running minimum is +inf): the body of
// kcenters_pass_fast_kernel (dist_kernels.cu) -- R frames per group and iteration straight from
// global memory, reference arithmetic, one split reduction -- plus the lane records.  Nothing is
// read back from `dist`, so it runs at the HBM roof (4 d bytes per frame).
// ---------------------------------------------------------------------------------------
template <int METRIC, int ITERS, int R>
__global__ void __launch_bounds__(kThreads, 4)
kcenters_first_pass_kernel(const float *__restrict__ X, long long n, int d, long long ld,
                           const float *__restrict__ center, int label0,
                           double *__restrict__ dist, int *__restrict__ labels,
                           long long row_offset, unsigned char *__restrict__ lane_buf, int G)
{
    typedef Metric<METRIC, float> M;
    const int lane_in_group = threadIdx.x & (G - 1);
    const long long gid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const long long NG = ((long long)gridDim.x * blockDim.x) / G;
    const long long warp_gid0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * (32 / G);
    const long long ld4 = ld >> 2;
    const float4 *X4 = reinterpret_cast<const float4 *>(X);
    const int lanes_per_f = G / R;
    const int fsel = lane_in_group / lanes_per_f;    // the split reduction peels the frame bits first
    const bool owner = (lane_in_group & (lanes_per_f - 1)) == 0;

    float4 c_first[ITERS];                           // the one centre lives in registers
#pragma unroll
    for (int i = 0; i < ITERS; ++i)
        c_first[i] = reinterpret_cast<const float4 *>(center)[lane_in_group + i * G];

    extern __shared__ float4 s_dyn[];
    LaneTop top;
    top.init(reinterpret_cast<LaneCand *>(s_dyn) + (threadIdx.x / G) * R + fsel, owner);
    const long long stride_j = NG * ld4;
    const float4 *pfull = X4 + gid * ld4 + lane_in_group;
    long long rbase = gid;                           // it * R * NG + gid
    auto iteration = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        long long rr[R];
        float4 x[R][ITERS];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            rr[j] = rbase + j * NG;
            const long long rc = (FULL || rr[j] < n) ? rr[j] : n - 1;
            const float4 *p = FULL ? pfull + j * stride_j : X4 + rc * ld4 + lane_in_group;
#pragma unroll
            for (int i = 0; i < ITERS; ++i) x[j][i] = ldg_stream(p + i * G);
        }
        long long myrow = -1;
#pragma unroll
        for (int j = 0; j < R; ++j)
            if (fsel == j && (FULL || rr[j] < n)) myrow = rr[j];
        double va[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            double a = 0.0, b = 0.0;
#pragma unroll
            for (int i = 0; i < ITERS; ++i) {
                M::acc(a, b, x[j][i].x, c_first[i].x);
                M::acc(a, b, x[j][i].y, c_first[i].y);
                M::acc(a, b, x[j][i].z, c_first[i].z);
                M::acc(a, b, x[j][i].w, c_first[i].w);
            }
            va[j] = a;
        }
        const double ra = group_reduce_split<false, R>(va, G, lane_in_group);
        const double dv = M::fin(ra, 0.0, d);
        double cur = INFINITY;
        if (owner && myrow >= 0) {
            if (dv < cur) {                  // false for NaN, like the reference's mask
                cur = dv;
                dist[myrow] = cur;
                labels[myrow] = label0;
            }
            top.add(cur, myrow);
        }
    };
    const long long full_span = (long long)(R - 1) * NG + warp_gid0 + (32 / G);   // + it * R * NG <= n: full
    const long long step_rows = (long long)R * NG;
    long long wrow = 0;                              // it * R * NG
    for (; wrow + full_span <= n; wrow += step_rows) {
        iteration(std::true_type());
        pfull += R * stride_j;
        rbase += step_rows;
    }
    for (; wrow + warp_gid0 < n; wrow += step_rows) {
        iteration(std::false_type());
        rbase += step_rows;
    }
    if (owner)
        top.store(reinterpret_cast<LaneCand *>(lane_buf + sizeof(LaneHeader)) + (gid * R + fsel), row_offset);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        reinterpret_cast<LaneHeader *>(lane_buf)->n_slots = NG * R;
}

// ---------------------------------------------------------------------------------------
// Fused pass over J pending centres (labels label0 .. label0 + J - 1), never the first pass.
// Centres are taken JB at a time: R x JB float32 squared distances per group and iteration (packed
// f32x2 add / fma), ONE split reduction for all of them (its shuffle latency is paid once per JB
// centres), then lane (f, jj) decides whether centre jj can possibly lower the running minimum of
// frame f; the few pairs that can are redone in the reference arithmetic by the whole group, in
// centre order.  Every lane whose reduction slot belongs to frame f carries that frame's running
// minimum (identical copies), one of them (`owner`) writes it back and keeps the lane's top three.
// Frames arrive through a per-warp cp.async ring (HBM latency overlaps the arithmetic without
// holding the data in registers).
// This is the second generation of the kernel: same arithmetic and outputs as the first, bit for
// bit; what changed is everything AROUND the arithmetic.  The r2o profile of the first version
// (7.5 G warp instructions for 4 centres x 50M frames, issue slots 70 % busy, 18.1 M cycles against
// the 15.3 M of the HBM stream) showed ~540 instructions per 4-frame iteration of which only ~270
// were the float32 filter and its reduction: the cp.async prefetch evaluated the clamped tail
// addressing (four 64-bit row * pitch products) next to the fast one on EVERY iteration and picked
// with SEL, the loop bounds (warp_gid0, full_span) were rematerialised from %tid / %ctaid per
// iteration, and every row index was rebuilt from `it`.  Here
//   * the number of full and of existing iterations of the warp is computed once; the loop is two
//     counted loops, the prefetch takes a warp-uniform branch to a fast path that is 8 LDGSTS + 3
//     pointer bumps;
//   * a lane carries ONE row index and ONE pointer into distances (its frame slot of the current
//     iteration), advanced by a constant;
//   * the ring is addressed by 32-bit shared addresses with immediate offsets.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(unsigned dst, const void *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned addr)
{
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}

template <int METRIC, int ITERS, int R, int JB>
__global__ void __launch_bounds__(kThreads, 2)
kcenters_fused_pass_kernel(const float *__restrict__ X, long long n, int d, long long ld,
                           const float *__restrict__ centers, int J, int label0,
                           double *__restrict__ dist, int *__restrict__ labels,
                           long long row_offset, unsigned char *__restrict__ lane_buf, int G,
                           float one_minus_eps)
{
    typedef Metric<METRIC, float> M;
    constexpr int V = R * JB;                        // float32 sums per group and chunk
    constexpr int STAGES = kStages;
    constexpr unsigned SLOT_BYTES = 32 * 16;         // one 16-byte request of every lane of the warp
    constexpr unsigned STAGE_BYTES = R * ITERS * SLOT_BYTES;
    // stored negated: x + (-c) is x - c bit for bit, and the float32 filter can then use the packed
    // f32x2 add / fma of sm_100 (two elements per instruction); the padding centres of the last
    // chunk are copies of the last real one and never looked at
    extern __shared__ float4 s_c[];                  // [J rounded up to JB][d / 4], NEGATED
    const int d4 = d >> 2;
    const int Jpad = (J + JB - 1) / JB * JB;
    {
        const float4 *c4 = reinterpret_cast<const float4 *>(centers);
        for (int i = threadIdx.x; i < Jpad * d4; i += blockDim.x) {
            const int jc = i / d4;
            const float4 v = c4[(jc < J ? jc : J - 1) * d4 + (i - jc * d4)];
            s_c[i] = make_float4(-v.x, -v.y, -v.z, -v.w);
        }
    }
    __syncthreads();

    const int lane_in_group = threadIdx.x & (G - 1);
    const long long tid_global = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long gid = tid_global / G;
    const long long NG = ((long long)gridDim.x * blockDim.x) / G;
    const long long warp_gid0 = (tid_global >> 5) * (32 / G);
    const long long ld4 = ld >> 2;
    const float4 *X4 = reinterpret_cast<const float4 *>(X);

    // reduction slots: lane -> value index vsel = lig / (G / V) = f * JB + jj; frame slot
    // f = lig / (G / R) (the split reduction peels the frame bits first)
    const int lanes_per_f = G / R;
    const int fsel = lane_in_group / lanes_per_f;
    const int vsel = lane_in_group / (G / V);
    const int jjsel = vsel - fsel * JB;
    const bool owner = (lane_in_group & (lanes_per_f - 1)) == 0;
    int slot_shift = 0;
    while ((1 << slot_shift) < G / V) ++slot_shift;
    const unsigned slot_lanes = (G / V >= 32) ? 0xffffffffu : ((1u << (G / V)) - 1u);

    // iterations of this warp: [0, n_full) touch only rows < n, [n_full, n_iter) are clamped
    const long long step_rows = (long long)R * NG;
    const long long full_span = (long long)(R - 1) * NG + warp_gid0 + (32 / G);
    const long long n_full = n >= full_span ? (n - full_span) / step_rows + 1 : 0;
    const long long n_iter = n > warp_gid0 ? (n - warp_gid0 + step_rows - 1) / step_rows : 0;

    const unsigned ring = (unsigned)__cvta_generic_to_shared(
        s_c + (size_t)Jpad * d4 + (size_t)(threadIdx.x >> 5) * (STAGES * R * ITERS * 32) + (threadIdx.x & 31));
    unsigned off_p = 0, off_c = 0;                   // byte offset of the stage being produced / consumed
    const long long stride_j = NG * ld4;             // float4 units between the R frames of a group
    const float4 *ppre = X4 + gid * ld4 + lane_in_group;
    long long pit = 0;                               // iteration being prefetched
    auto prefetch = [&]() {
        if (pit < n_full) {                          // warp uniform
            const float4 *p = ppre;
#pragma unroll
            for (int j = 0; j < R; ++j) {
#pragma unroll
                for (int i = 0; i < ITERS; ++i)
                    cp_async16(ring + off_p + (unsigned)(j * ITERS + i) * SLOT_BYTES, p + i * G);
                p += stride_j;
            }
            ppre = p;
        } else if (pit < n_iter) {
            const long long row0 = pit * step_rows + gid;
#pragma unroll 1
            for (int j = 0; j < R; ++j) {
                const long long row = row0 + j * NG;
                const float4 *p = X4 + (row < n ? row : n - 1) * ld4 + lane_in_group;
#pragma unroll
                for (int i = 0; i < ITERS; ++i)
                    cp_async16(ring + off_p + (unsigned)(j * ITERS + i) * SLOT_BYTES, p + i * G);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");      // (possibly empty: keeps the count uniform)
        ++pit;
        off_p = off_p + STAGE_BYTES == STAGES * STAGE_BYTES ? 0u : off_p + STAGE_BYTES;
    };
#pragma unroll 1
    for (int s = 0; s < STAGES - 1; ++s) prefetch();

    LaneTop top;                                     // records: behind the frame ring
    top.init(reinterpret_cast<LaneCand *>(s_c + (size_t)Jpad * d4 + (size_t)(kThreads / 32) * (STAGES * R * ITERS * 32)) +
                 (threadIdx.x / G) * R + fsel, owner);
    long long myrow = (long long)fsel * NG + gid;    // this lane's frame slot in the current iteration
    double *dptr = dist + myrow;
    // the frame's running minimum is requested one iteration ahead (its load latency would
    // otherwise sit in front of every iteration: 30 % of the stall samples of the first version)
    double cur_pre = myrow < n ? __ldcg(dptr) : INFINITY;

    auto iteration = [&](auto full_tag) {
        constexpr bool FULL = decltype(full_tag)::value;
        prefetch();
        asm volatile("cp.async.wait_group %0;" :: "n"(STAGES - 1) : "memory");
        float4 x[R][ITERS];
#pragma unroll
        for (int j = 0; j < R; ++j)
#pragma unroll
            for (int i = 0; i < ITERS; ++i)
                x[j][i] = lds128(ring + off_c + (unsigned)(j * ITERS + i) * SLOT_BYTES);
        const bool valid = FULL || myrow < n;
        double cur = valid ? cur_pre : INFINITY;
        int lab = -1;
        cur_pre = myrow + step_rows < n ? __ldcg(dptr + step_rows) : INFINITY;
        // float upper bound of what the float32 sums are compared with
        float bound = __double2float_ru(METRIC == MSMB200_EUCLIDEAN ? cur * cur : cur);
        for (int j0 = 0; j0 < J; j0 += JB) {
            float s[V];
#pragma unroll
            for (int jj = 0; jj < JB; ++jj) {
                if (j0 + jj >= J) {                              // padding centre of the last chunk (warp uniform):
#pragma unroll
                    for (int f = 0; f < R; ++f) s[f * JB + jj] = 0.f;   // its slots are never looked at
                    continue;
                }
                float4 c[ITERS];
#pragma unroll
                for (int i = 0; i < ITERS; ++i) c[i] = s_c[(j0 + jj) * d4 + lane_in_group + i * G];
#pragma unroll
                for (int f = 0; f < R; ++f) {
                    float2 a2 = make_float2(0.f, 0.f);
#pragma unroll
                    for (int i = 0; i < ITERS; ++i) {
                        float2 t = __fadd2_rn(make_float2(x[f][i].x, x[f][i].y), make_float2(c[i].x, c[i].y));
                        a2 = __ffma2_rn(t, t, a2);
                        t = __fadd2_rn(make_float2(x[f][i].z, x[f][i].w), make_float2(c[i].z, c[i].w));
                        a2 = __ffma2_rn(t, t, a2);
                    }
                    s[f * JB + jj] = a2.x + a2.y;
                }
            }
            const float tot = group_reduce_split_f32<V>(s, G, lane_in_group);
            // certainly not below the current minimum -> the reference's mask is false (the relative
            // margin needs float32's normal range: a sum below 1e-30 always goes to the refine step; so
            // does a sum that overflowed float32 -- its float64 value may still be below the minimum)
            const bool need = valid && (j0 + jjsel) < J &&
                              !(tot * one_minus_eps >= bound && tot >= 1e-30f && tot <= 3.0e38f);
            const unsigned need_mask = __ballot_sync(0xffffffffu, need);
            if (need_mask == 0) continue;
            unsigned m = need_mask;
            for (int g = G; g < 32; g <<= 1) m |= m >> g;
            if (G < 32) m &= (1u << G) - 1u;
            while (m) {
                const int q = (__ffs(m) - 1) >> slot_shift;
                m &= ~(slot_lanes << (q << slot_shift));
                const int f = q / JB, jj = q - f * JB;
                const float4 *cn = s_c + (j0 + jj) * d4 + lane_in_group;
                double a = 0.0, b = 0.0;
#pragma unroll
                for (int ff = 0; ff < R; ++ff) {
                    if (ff != f) continue;                  // warp uniform; x[] stays in registers
#pragma unroll
                    for (int i = 0; i < ITERS; ++i) {
                        const float4 c = cn[i * G];
                        M::acc(a, b, x[ff][i].x, -c.x);
                        M::acc(a, b, x[ff][i].y, -c.y);
                        M::acc(a, b, x[ff][i].z, -c.z);
                        M::acc(a, b, x[ff][i].w, -c.w);
                    }
                }
                a = group_combine<false>(a, G);
                if (fsel == f && valid) {
                    const double dv = M::fin(a, 0.0, d);
                    if (dv < cur) {                         // strict: kcenters.py:93
                        cur = dv;
                        lab = label0 + j0 + jj;
                        bound = __double2float_ru(METRIC == MSMB200_EUCLIDEAN ? a : dv);
                    }
                }
            }
        }
        if (owner && valid) {
            if (lab >= 0) {
                *dptr = cur;
                labels[myrow] = lab;
            }
            top.add(cur, myrow);
        }
        myrow += step_rows;
        dptr += step_rows;
        off_c = off_c + STAGE_BYTES == STAGES * STAGE_BYTES ? 0u : off_c + STAGE_BYTES;
    };
    long long it = 0;
#pragma unroll 1
    for (; it < n_full; ++it) iteration(std::true_type());
#pragma unroll 1
    for (; it < n_iter; ++it) iteration(std::false_type());
    asm volatile("cp.async.wait_group 0;" ::: "memory");

    if (owner)
        top.store(reinterpret_cast<LaneCand *>(lane_buf + sizeof(LaneHeader)) + (gid * R + fsel), row_offset);
    if (blockIdx.x == 0 && threadIdx.x == 0)
        reinterpret_cast<LaneHeader *>(lane_buf)->n_slots = NG * R;
}

// (A "lane owns its frames" formulation of this pass -- 32 or 64 consecutive frames per warp, one or two
// per lane, float32 sums without any reduction, the exact step done by the whole warp from L2 -- was
// built, passed the same bit-for-bit tests and was SLOWER: 13.2 ms (one frame per lane: the warp-wide
// broadcast reads of the centre chunks saturate the shared-memory pipe) and 16.7 ms (two frames per
// lane, 128-byte row segments per stage: DRAM page locality) against 11.9 ms for the 7-centre pass
// of the bench; profiles/r2x_lane_owns_frame_experiment.log.  Removed.)

// ---------------------------------------------------------------------------------------
// Candidate selection: one block.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long cand_key(double v)
{
    // running minima are >= 0; -inf (empty lane), zeros and NaN never qualify
    return (v > 0.0) ? (unsigned long long)__double_as_longlong(v) : 0ull;
}

__global__ void __launch_bounds__(1024)
kcenters_select_kernel(const unsigned char *__restrict__ lane_buf, const float *__restrict__ X,
                       long long n, int d, long long ld, long long row_offset, int t_cap,
                       unsigned char *__restrict__ set_out)
{
    __shared__ ArgMax s_arg[32];
    __shared__ double s_v3[32];
    __shared__ unsigned long long s_cnt[32];
    __shared__ unsigned long long s_total;
    __shared__ int s_count;
    __shared__ ArgMax s_best;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long n_slots = reinterpret_cast<const LaneHeader *>(lane_buf)->n_slots;
    const LaneCand *lc = reinterpret_cast<const LaneCand *>(lane_buf + sizeof(LaneHeader));

    // (1) global arg-max of v1 (lowest row among equals) and the largest third best
    ArgMax best{-INFINITY, kNoRow};
    double maxv3 = -INFINITY;
    for (long long s = tid; s < n_slots; s += blockDim.x) {
        ArgMax c{lc[s].v1, lc[s].i1};
        if (c.i != kNoRow) best = argmax_merge(best, c);
        maxv3 = fmax(maxv3, lc[s].v3);
    }
    best = argmax_warp(best);
    for (int off = 16; off > 0; off >>= 1) maxv3 = fmax(maxv3, __shfl_xor_sync(0xffffffffu, maxv3, off));
    if (lane == 0) { s_arg[warp] = best; s_v3[warp] = maxv3; }
    __syncthreads();
    if (warp == 0) {
        ArgMax b = s_arg[lane];
        double m3 = s_v3[lane];
        b = argmax_warp(b);
        for (int off = 16; off > 0; off >>= 1) m3 = fmax(m3, __shfl_xor_sync(0xffffffffu, m3, off));
        if (lane == 0) { s_best = b; s_v3[0] = m3; }
    }
    __syncthreads();
    best = s_best;
    maxv3 = s_v3[0];

    // (2) key of the t_cap-th largest ENTRY (every lane contributes its two largest values; 0 if fewer
    //     than t_cap are positive): radix select, 8 bits per round from the top -- 8 passes over the
    //     lane records instead of the 63 of a bitwise search (this block is the serial section
    //     between two fused passes)
    unsigned long long K = 0;
    {
        __shared__ unsigned s_hist[256];
        __shared__ unsigned long long s_prefix;
        __shared__ unsigned s_rank;
        // positive keys in all
        unsigned long long cnt = 0;
        for (long long s = tid; s < n_slots; s += blockDim.x)
            cnt += (cand_key(lc[s].v1) != 0ull) + (cand_key(lc[s].v2) != 0ull);
        for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
        if (lane == 0) s_cnt[warp] = cnt;
        __syncthreads();
        if (warp == 0) {
            unsigned long long c = s_cnt[lane];
            for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xffffffffu, c, off);
            if (lane == 0) { s_total = c; s_prefix = 0ull; s_rank = (unsigned)t_cap; }
        }
        __syncthreads();
        if (s_total >= (unsigned long long)t_cap) {
            for (int shift = 56; shift >= 0; shift -= 8) {
                if (tid < 256) s_hist[tid] = 0u;
                __syncthreads();
                const unsigned long long prefix = s_prefix;
                for (long long s = tid; s < n_slots; s += blockDim.x) {
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const unsigned long long key = cand_key(e ? lc[s].v2 : lc[s].v1);
                        // keys that agree with the digits chosen so far (a zero key -- no entry -- only
                        // ever lands in bin 0 of a prefix that is still all zeros, below every rank asked for)
                        if (shift == 56 || (key >> (shift + 8)) == (prefix >> (shift + 8)))
                            atomicAdd(&s_hist[(unsigned)(key >> shift) & 0xFFu], 1u);
                    }
                }
                __syncthreads();
                if (tid == 0) {
                    // the digit whose bin holds the s_rank-th largest of the remaining keys
                    unsigned rank = s_rank, above = 0u;
                    int b = 255;
                    for (; b > 0; --b) {
                        if (above + s_hist[b] >= rank) break;
                        above += s_hist[b];
                    }
                    s_prefix = prefix | ((unsigned long long)b << shift);
                    s_rank = rank - above;
                }
                __syncthreads();
            }
            K = s_prefix;
        }
        __syncthreads();
    }

    // (3) candidates: entries strictly above that value (< t_cap of them), plus the arg-max
    SetHeader *hdr = reinterpret_cast<SetHeader *>(set_out);
    double *val = set_val(set_out);
    long long *idx = set_idx(set_out, t_cap);
    float *rows = set_rows(set_out, t_cap);
    if (tid == 0) s_count = 0;
    __syncthreads();
    bool have_best = false;
    for (long long s = tid; s < n_slots; s += blockDim.x) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const double v = e ? lc[s].v2 : lc[s].v1;
            const long long i = e ? lc[s].i2 : lc[s].i1;
            if (cand_key(v) > K) {
                const int slot = atomicAdd(&s_count, 1);
                val[slot] = v;
                idx[slot] = i;
                if (i == best.i) have_best = true;
            }
        }
    }
    const int any_best = __syncthreads_or(have_best);
    if (tid == 0 && !any_best && best.i != kNoRow) {
        const int slot = s_count++;
        val[slot] = best.v;
        idx[slot] = best.i;
    }
    __syncthreads();
    const int count = s_count;
    if (tid == 0) {
        hdr->count = count;
        hdr->cap = t_cap;
        // frames that are not candidates: neither of their lane's two largest (<= its third best <= the
        // largest third best) or an entry at or below the cut (<= its value; 0 when the cut is
        // "everything positive")
        const double cut = K ? __longlong_as_double((long long)K) : 0.0;
        hdr->tau = fmax(maxv3, cut);
    }
    // (4) their frames
    for (int c = warp; c < count; c += blockDim.x >> 5) {
        const float *src = X + (idx[c] - row_offset) * ld;
        for (int j = lane; j < d; j += 32) rows[(size_t)c * d + j] = src[j];
    }
}

// ---------------------------------------------------------------------------------------
// The reference's loop on the candidates of all ranks: one block.
// ---------------------------------------------------------------------------------------
template <int METRIC>
__global__ void __launch_bounds__(1024)
kcenters_chain_kernel(unsigned char *__restrict__ sets, int n_sets, size_t set_stride, int d,
                      int G, int k_remaining, int j_cap, unsigned char *__restrict__ centers_out)
{
    __shared__ ArgMax s_arg[32];
    __shared__ int s_slot[32];
    __shared__ ArgMax s_best;
    __shared__ int s_best_slot;
    __shared__ int s_prefix[65];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    CentersHeader *out_hdr = reinterpret_cast<CentersHeader *>(centers_out);
    long long *out_ids = reinterpret_cast<long long *>(centers_out + sizeof(CentersHeader));
    float *out_rows = reinterpret_cast<float *>(centers_out + sizeof(CentersHeader) + (size_t)j_cap * 8);

    if (tid == 0) {
        int p = 0;
        for (int s = 0; s < n_sets; ++s) {
            s_prefix[s] = p;
            p += reinterpret_cast<const SetHeader *>(sets + s * set_stride)->count;
        }
        s_prefix[n_sets] = p;
    }
    __syncthreads();
    const int total = s_prefix[n_sets];
    double tau = -INFINITY;
    for (int s = 0; s < n_sets; ++s)
        tau = fmax(tau, reinterpret_cast<const SetHeader *>(sets + s * set_stride)->tau);
    // flat candidate c -> (set, local slot)
    auto locate = [&](int c, int &set, int &slot) {
        int s = 0;
        while (c >= s_prefix[s + 1]) ++s;
        set = s;
        slot = c - s_prefix[s];
    };

    int steps = 0;
    const int max_steps = k_remaining < j_cap ? k_remaining : j_cap;
    const int lig = tid & (G - 1);
    const int groups = blockDim.x / G;
    while (steps < max_steps) {
        // arg-max over the candidates' current minima (lowest global row among equals)
        ArgMax best{-INFINITY, kNoRow};
        int best_c = -1;
        for (int c = tid; c < total; c += blockDim.x) {
            int set, slot;
            locate(c, set, slot);
            unsigned char *sb = sets + set * set_stride;
            const int cap = reinterpret_cast<const SetHeader *>(sb)->cap;
            ArgMax a{set_val(sb)[slot], set_idx(sb, cap)[slot]};
            if (a.v > best.v || (a.v == best.v && a.i < best.i)) { best = a; best_c = c; }
        }
        for (int off = 16; off > 0; off >>= 1) {
            ArgMax o;
            o.v = __shfl_xor_sync(0xffffffffu, best.v, off);
            o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
            const int oc = __shfl_xor_sync(0xffffffffu, best_c, off);
            if (o.v > best.v || (o.v == best.v && o.i < best.i)) { best = o; best_c = oc; }
        }
        if (lane == 0) { s_arg[warp] = best; s_slot[warp] = best_c; }
        __syncthreads();
        if (warp == 0) {
            best = s_arg[lane];
            best_c = s_slot[lane];
            for (int off = 16; off > 0; off >>= 1) {
                ArgMax o;
                o.v = __shfl_xor_sync(0xffffffffu, best.v, off);
                o.i = __shfl_xor_sync(0xffffffffu, best.i, off);
                const int oc = __shfl_xor_sync(0xffffffffu, best_c, off);
                if (o.v > best.v || (o.v == best.v && o.i < best.i)) { best = o; best_c = oc; }
            }
            if (lane == 0) { s_best = best; s_best_slot = best_c; }
        }
        __syncthreads();
        best = s_best;
        best_c = s_best_slot;
        // the first pick is the true arg-max of the pass; later ones must beat every frame that
        // is not a candidate
        if (best_c < 0 || (steps > 0 && !(best.v > tau))) break;
        {
            int set, slot;
            locate(best_c, set, slot);
            unsigned char *sb = sets + set * set_stride;
            const int cap = reinterpret_cast<const SetHeader *>(sb)->cap;
            const float *src = set_rows(sb, cap) + (size_t)slot * d;
            float *dst = out_rows + (size_t)steps * d;
            for (int j = tid; j < d; j += blockDim.x) dst[j] = src[j];
            if (tid == 0) out_ids[steps] = best.i;
        }
        __syncthreads();
        ++steps;
        if (steps == max_steps) break;
        // lower the candidates' minima (same arithmetic and summation order as the pass)
        const float *cen = out_rows + (size_t)(steps - 1) * d;
        for (int c0 = 0; c0 < total; c0 += groups) {
            const int c = c0 + tid / G;
            const int cc = c < total ? c : total - 1;
            int set, slot;
            locate(cc, set, slot);
            unsigned char *sb = sets + set * set_stride;
            const int cap = reinterpret_cast<const SetHeader *>(sb)->cap;
            const float *row = set_rows(sb, cap) + (size_t)slot * d;
            const double dv = group_distance<METRIC, float, true, false>(row, cen, d, lig, G);
            if (c < total && lig == 0) {
                double *v = set_val(sb) + slot;
                if (dv < *v) *v = dv;
            }
        }
        __syncthreads();
    }
    if (tid == 0) {
        out_hdr->n = steps;
        out_hdr->cap = j_cap;
    }
}

static bool lookahead_shape_ok(const void *X, int d, long long ld, int dtype, int metric, int *G_out,
                               int *iters_out)
{
    if (dtype != MSMB200_F32) return false;
    if (metric != MSMB200_EUCLIDEAN && metric != MSMB200_SQEUCLIDEAN) return false;
    if (d % 4 != 0 || ld % 4 != 0 || (X && !aligned16(X))) return false;
    const int G = lanes_per_row(d, 4, true);
    if (G < 4 || (d / 4) % G != 0) return false;
    const int iters = (d / 4) / G;
    if (iters != 1 && iters != 2 && iters != 4) return false;
    if (G_out) *G_out = G;
    if (iters_out) *iters_out = iters;
    return true;
}

}  // namespace msmb

using namespace msmb;

extern "C" int msmb200_kcenters_lookahead_supported(int d, int64_t ld, int dtype, int metric)
{
    return lookahead_shape_ok(nullptr, d, ld, dtype, metric, nullptr, nullptr) ? 1 : 0;
}

extern "C" size_t msmb200_kcenters_lane_bytes(int device)
{
    (void)device;
    return sizeof(LaneHeader) + sizeof(LaneCand) * (size_t)(8 * 256) * kThreads;
}

extern "C" size_t msmb200_kcenters_set_bytes(int d, int t_cap) { return set_bytes(d, t_cap); }

extern "C" size_t msmb200_kcenters_centers_bytes(int d, int j_cap) { return centers_bytes(d, j_cap); }

extern "C" int msmb200_kcenters_multi_pass(const void *X, int64_t n, int d, int64_t ld, int dtype,
                                           int metric, const void *centers, int n_centers, int j_cap,
                                           int32_t label0, int first, double *distances,
                                           int32_t *labels, int64_t row_offset, void *lane_buf,
                                           size_t lane_bytes, void *stream)
{
    int G = 0, iters = 0;
    MSMB_REQUIRE(n >= 0 && d > 0 && ld >= d, "kcenters_multi_pass: bad shape");
    MSMB_REQUIRE(X && centers && distances && labels && lane_buf, "kcenters_multi_pass: null pointer");
    MSMB_REQUIRE(lookahead_shape_ok(X, d, ld, dtype, metric, &G, &iters),
                 "kcenters_multi_pass: unsupported shape/metric (d=%d ld=%lld dtype=%d metric=%d)", d,
                 (long long)ld, dtype, metric);
    MSMB_REQUIRE(n_centers >= 1 && n_centers <= j_cap && (!first || n_centers == 1),
                 "kcenters_multi_pass: bad centre count %d (cap %d, first %d)", n_centers, j_cap, first);
    cudaStream_t st = (cudaStream_t)stream;
    long long per_block = kThreads / G;
    long long blocks = (n + per_block - 1) / per_block;
    if (blocks > pass_grid()) blocks = pass_grid();
    if (blocks < 1) blocks = 1;
    const int grid = (int)blocks;
    const int R = iters == 4 ? 2 : 4;
    MSMB_REQUIRE(lane_bytes >= sizeof(LaneHeader) + sizeof(LaneCand) * (size_t)grid * kThreads / G * R,
                 "kcenters_multi_pass: lane buffer too small");
    const float *crow = reinterpret_cast<const float *>(
        reinterpret_cast<const unsigned char *>(centers) + sizeof(CentersHeader) + (size_t)j_cap * 8);
    const float om_eps = 1.0f - 4.0f * (float)(d + 8) * 5.9604645e-8f;
#define MSMB_FIRST(METRIC, I, RR)                                                                 \
    kcenters_first_pass_kernel<METRIC, I, RR><<<grid, kThreads, lane_top_bytes(G, RR), st>>>(    \
        (const float *)X, n, d, ld, crow, label0, distances, labels, row_offset,                  \
        (unsigned char *)lane_buf, G)
#define MSMB_FUSED(METRIC, I, RR, JBV)                                                            \
    do {                                                                                          \
        auto kern = kcenters_fused_pass_kernel<METRIC, I, RR, JBV>;                               \
        const size_t smem = (size_t)((n_centers + JBV - 1) / JBV * JBV) * d * sizeof(float) +     \
            (size_t)(kThreads / 32) * kStages * RR * I * 32 * sizeof(float4) +                    \
            lane_top_bytes(G, RR);                                                                \
        MSMB_REQUIRE(smem <= 113 * 1024, "kcenters_multi_pass: %d centres of %d floats exceed "   \
                     "the shared memory of two resident blocks", n_centers, d);                   \
        if (smem > 48 * 1024)                                                                     \
            MSMB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                           (int)smem));                                           \
        kern<<<grid, kThreads, smem, st>>>((const float *)X, n, d, ld, crow, n_centers, label0,   \
                                           distances, labels, row_offset,                         \
                                           (unsigned char *)lane_buf, G, om_eps);                 \
    } while (0)
    // centres per chunk (JB): R * JB float32 sums are reduced at once and must not outnumber the lanes
#define MSMB_PASS_M(METRIC, I, RR)                                                                \
    do {                                                                                          \
        const int jb = G / RR >= 4 ? 4 : G / RR;                                                  \
        if (first) MSMB_FIRST(METRIC, I, RR);                                                     \
        else if (jb == 4) MSMB_FUSED(METRIC, I, RR, 4);                                           \
        else if (jb == 2) MSMB_FUSED(METRIC, I, RR, 2);                                           \
        else MSMB_FUSED(METRIC, I, RR, 1);                                                        \
    } while (0)
#define MSMB_PASS(I, RR)                                                                          \
    do {                                                                                          \
        if (metric == MSMB200_EUCLIDEAN) MSMB_PASS_M(MSMB200_EUCLIDEAN, I, RR);                   \
        else MSMB_PASS_M(MSMB200_SQEUCLIDEAN, I, RR);                                             \
    } while (0)
    if (iters == 1) MSMB_PASS(1, 4);
    else if (iters == 2) MSMB_PASS(2, 4);
    else MSMB_PASS(4, 2);
#undef MSMB_PASS
#undef MSMB_PASS_M
#undef MSMB_FUSED
#undef MSMB_FIRST
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_kcenters_select(const void *X, int64_t n, int d, int64_t ld, int64_t row_offset,
                                       const void *lane_buf, int t_cap, void *set_out, void *stream)
{
    MSMB_REQUIRE(X && lane_buf && set_out && t_cap >= 2, "kcenters_select: bad arguments");
    kcenters_select_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(
        (const unsigned char *)lane_buf, (const float *)X, n, d, ld, row_offset, t_cap,
        (unsigned char *)set_out);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}

extern "C" int msmb200_kcenters_chain(void *sets, int n_sets, size_t set_stride, int d, int metric,
                                      int k_remaining, int j_cap, void *centers_out, void *stream)
{
    MSMB_REQUIRE(sets && centers_out && n_sets >= 1 && n_sets <= 64 && k_remaining >= 1 && j_cap >= 1,
                 "kcenters_chain: bad arguments");
    int G = 0;
    MSMB_REQUIRE(lookahead_shape_ok(nullptr, d, d, MSMB200_F32, metric, &G, nullptr),
                 "kcenters_chain: unsupported shape/metric");
    if (metric == MSMB200_EUCLIDEAN)
        kcenters_chain_kernel<MSMB200_EUCLIDEAN><<<1, 1024, 0, (cudaStream_t)stream>>>(
            (unsigned char *)sets, n_sets, set_stride, d, G, k_remaining, j_cap,
            (unsigned char *)centers_out);
    else
        kcenters_chain_kernel<MSMB200_SQEUCLIDEAN><<<1, 1024, 0, (cudaStream_t)stream>>>(
            (unsigned char *)sets, n_sets, set_stride, d, G, k_remaining, j_cap,
            (unsigned char *)centers_out);
    MSMB_LAUNCH_CHECK();
    return MSMB200_OK;
}
