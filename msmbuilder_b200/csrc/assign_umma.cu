// assign_umma.cu -- K3 on Blackwell tensor cores: the filter stage of assign_nearest for float32
// (sq)euclidean (libdistance.pyx:82-131 -> assign.hpp:50-91) as a tcgen05 GEMM.
//
//   arg min_j |x_i - c_j|^2  =  arg min_j ( |c_j|^2 - 2 x_i . c_j )
//
// The inner products S = X C^T (n x k, contraction over the d features) run on the tensor cores with
// the same error-compensated fp16 split as K1: every frame is scaled by its own power of two
// (largest magnitude -> [1, 2)), the centres by one global power of two, both are split x' = h + l
// into fp16 parts and three products h ch + h cl + l ch are accumulated in fp32 TMEM (|error| <=
// 2^-19 |x'||c'| + 2^-22 d, derived in DESIGN.md).  The epilogue forms |c_j|^2 - 2 S_ij from TMEM,
// tracks the best and second best centre of each frame, and every frame whose gap is inside the
// error bound -- exact ties included -- goes to the ambiguity list that assign_refine_kernel
// (dist_kernels.cu) re-scans with the reference's float64 arithmetic, lowest index first
// (assign.hpp:69).  Labels are therefore those of the exact engine; the winning distance is
// recomputed in float64 by assign_mindist_kernel as before.
//
// One CTA (cta_group::1, M = 128 frames, N <= 256 centres per UMMA, K = 16 features per UMMA):
//   warps 2-9  converters: 32-byte segments of the frame tile straight from global memory (coalesced
//              16-byte loads, four segments per thread in flight), per-frame scale from a butterfly over
//              the frame's segments, fp16 h / l K-major chunks into a double-buffered operand stage;
//   warp 1     UMMA issuer (one elected lane), accumulators double buffered in TMEM (2 x 256 columns);
//   warps 10-13 epilogue, one per TMEM lane quarter: lane = frame.
// The centres' fp16 tiles are prepared once per call by assign_umma_prep_kernel.  When they fit 96 KB
// (k_pad * d_pad * 4 bytes) they stay resident in shared memory (assign_umma_kernel); larger centre
// tables -- every d > 64, k = 2000 at d = 128 -- are STREAMED (assign_umma_stream_kernel): the frame
// tile's operand stays in shared memory while chunks of NS centres arrive from L2 through a two-slot
// ring of TMA bulk copies (a loader lane, mbarrier complete_tx), one accumulator buffer per chunk.
#include "common.cuh"

namespace msmb {

namespace {

constexpr int AU_M = 128;                 // frames per tile
constexpr int AU_NT = 256;                // centres per UMMA / accumulator buffer
constexpr int AU_CONV_WARPS = 8;
constexpr int AU_EPI_WARPS = 4;
constexpr int AU_FIRST_CONV = 2;
constexpr int AU_FIRST_EPI = AU_FIRST_CONV + AU_CONV_WARPS;       // 10: quarter = warp & 3 (any 4 consecutive warps)
constexpr int AU_THREADS = 32 * (AU_FIRST_EPI + AU_EPI_WARPS + 2);   // 16 warps (14: loader of the streamed kernel)
constexpr int AU_CONV_BATCH = 4;          // segments a converter thread has in flight
constexpr size_t AU_B_LIMIT = 96 * 1024;

__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = s_u32(bar);
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity), "r"(1000u) : "memory");
        if (ok) break;
    }
}
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred = 0, laneid = 0;
    asm volatile(
        "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %2;\n\t"
        "@%%px mov.s32 %1, 1;\n\t"
        "mov.s32 %0, %%rx;\n\t}"
        : "+r"(laneid), "+r"(pred) : "r"(0xFFFFFFFFu));
    return pred;
}
// K-major, no swizzle: 16-byte chunks of 8 K elements, core matrix = 8 rows x 16 bytes
__device__ __forceinline__ uint64_t kdesc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
// (h, l) fp16 pair split of two floats: h = fp16(x), l = fp16(x - h)
__device__ __forceinline__ void split_h2(float a0, float a1, uint32_t &h, uint32_t &l)
{
    h = pack_h2(a0, a1);
    float l0, l1;
    asm("{\n\t.reg .f16 lo, hi, m1;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
        "fma.rn.f32.f16 %0, lo, m1, %3;\n\tfma.rn.f32.f16 %1, hi, m1, %4;\n\t}"
        : "=f"(l0), "=f"(l1) : "r"(h), "f"(a0), "f"(a1));
    l = pack_h2(l0, l1);
}
#define AU_TMEM_LD32(v, taddr) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
    : "r"(taddr) : "memory")

// power of two 2^-e with 2^e <= m < 2^(e+1) (1 for m = 0 or non-finite: such frames end up ambiguous)
__device__ __forceinline__ float pow2_scale(float m, float &inv)
{
    if (!(m > 0.f) || !(m < INFINITY)) { inv = 1.f; return 1.f; }
    int e = ilogbf(m);
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    inv = ldexpf(1.f, e);
    return ldexpf(1.f, -e);
}

struct AuPrep {              // written by the prep kernel, read by the main kernel
    float c_inv_scale;       // 2^E
    float c_norm_max;        // max_j |c'_j|_2 (scaled centres)
    float cn_max;            // max_j |c_j|^2
    int pad;
};

}  // namespace

// centres -> fp16 h / l tiles [n_tile][k chunk][nt_rows rows][16 B] (zero padded), |c_j|^2, global scale.
// One block; the centre table is small (<= 96 KB of tiles).
__global__ void __launch_bounds__(256)
assign_umma_prep_kernel(const float *__restrict__ Y, int k, int d, int d_pad, int k_pad, int nt_rows,
                        unsigned char *__restrict__ tiles_h, unsigned char *__restrict__ tiles_l,
                        float *__restrict__ cn, AuPrep *__restrict__ prep)
{
    __shared__ float s_red[256];
    const int tid = threadIdx.x;
    float m = 0.f;
    for (int e = tid; e < k * d; e += 256) m = fmaxf(m, fabsf(Y[e]));
    s_red[tid] = m;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) s_red[tid] = fmaxf(s_red[tid], s_red[tid + off]);
        __syncthreads();
    }
    float inv;
    const float scale = pow2_scale(s_red[0], inv);
    __syncthreads();
    const int nc = d_pad / 8;
    float nmax = 0.f, cnmax = 0.f;
    for (int j = tid; j < k_pad; j += 256) {
        double s2 = 0.0;
        float sp2 = 0.f;
        const int nt = j / nt_rows, r = j % nt_rows;
        for (int c = 0; c < nc; ++c) {
            float a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int f = 8 * c + q;
                const float v = (j < k && f < d) ? Y[(size_t)j * d + f] : 0.f;
                s2 += (double)v * (double)v;
                a[q] = v * scale;
                sp2 = fmaf(a[q], a[q], sp2);
            }
            uint4 hw, lw;
            split_h2(a[0], a[1], hw.x, lw.x);
            split_h2(a[2], a[3], hw.y, lw.y);
            split_h2(a[4], a[5], hw.z, lw.z);
            split_h2(a[6], a[7], hw.w, lw.w);
            const size_t o = (((size_t)nt * nc + c) * nt_rows + r) * 16;
            *reinterpret_cast<uint4 *>(tiles_h + o) = hw;
            *reinterpret_cast<uint4 *>(tiles_l + o) = lw;
        }
        // padding centres can never win: +inf squared norm
        cn[j] = j < k ? (float)s2 : INFINITY;
        if (j < k) { nmax = fmaxf(nmax, sqrtf(sp2)); cnmax = fmaxf(cnmax, (float)s2); }
    }
    s_red[tid] = nmax;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) s_red[tid] = fmaxf(s_red[tid], s_red[tid + off]);
        __syncthreads();
    }
    const float nm = s_red[0];
    __syncthreads();
    s_red[tid] = cnmax;
    __syncthreads();
    for (int off = 128; off > 0; off >>= 1) {
        if (tid < off) s_red[tid] = fmaxf(s_red[tid], s_red[tid + off]);
        __syncthreads();
    }
    if (tid == 0) {
        prep->c_inv_scale = inv;
        prep->c_norm_max = nm;
        prep->cn_max = s_red[0];
        prep->pad = 0;
    }
}

// One frame tile (128 frames x d_pad features) -> fp16 h / l K-major chunks [nc][128][16 B] plus the
// per-frame multiplier and ambiguity margin; shared by the resident and the streamed kernel.  A thread
// takes 32-byte segments s = fr * nc + c (nc = d_pad / 8, a power of two: the segments of a frame sit
// in adjacent lanes) and keeps AU_CONV_BATCH of them in flight -- with one segment per round trip
// the 32 segments a thread owns at d = 256 cost 30 us of pure load latency per tile (r2t: 22.8 ms
// for 10M x 256 frames against 8 centres).
__device__ __forceinline__ void au_convert_tile(const float *__restrict__ X, long long n, int d, long long ld,
                                                int nc, int d_pad, long long row0, unsigned char *ah,
                                                unsigned char *al, int ct, float *f_mul, float *f_margin,
                                                float c_inv, float c_nmax, float cn_max)
{
    constexpr int STRIDE = 32 * AU_CONV_WARPS;
    const int segs = AU_M * nc;                              // a multiple of STRIDE (nc >= 2)
    const int nc_shift = 31 - __clz(nc);
    for (int s0 = ct; s0 < segs; s0 += AU_CONV_BATCH * STRIDE) {
        float4 u[AU_CONV_BATCH], w[AU_CONV_BATCH];
#pragma unroll
        for (int q = 0; q < AU_CONV_BATCH; ++q) {
            const int s = s0 + q * STRIDE;
            const int fr = s >> nc_shift, c = s & (nc - 1);
            const long long row = row0 + fr;
            u[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            w[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (s < segs && row < n && 8 * c < d) {
                const float4 *src = reinterpret_cast<const float4 *>(X + row * ld + 8 * c);
                u[q] = __ldg(src);
                if (8 * c + 4 < d) w[q] = __ldg(src + 1);
            }
        }
#pragma unroll
        for (int q = 0; q < AU_CONV_BATCH; ++q) {
            const int s = s0 + q * STRIDE;
            if (s >= segs) break;                            // block uniform
            const int fr = s >> nc_shift, c = s & (nc - 1);
            const float a[8] = {u[q].x, u[q].y, u[q].z, u[q].w, w[q].x, w[q].y, w[q].z, w[q].w};
            // largest magnitude and squared norm of the frame: butterfly over its nc segments
            float m = 0.f, s2 = 0.f;
#pragma unroll
            for (int e = 0; e < 8; ++e) { m = fmaxf(m, fabsf(a[e])); s2 = fmaf(a[e], a[e], s2); }
            for (int off = 1; off < nc && off < 32; off <<= 1) {
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, off));
                s2 += __shfl_xor_sync(0xffffffffu, s2, off);
            }
            float inv;
            const float scale = pow2_scale(m, inv);
            uint4 hw, lw;
            split_h2(a[0] * scale, a[1] * scale, hw.x, lw.x);
            split_h2(a[2] * scale, a[3] * scale, hw.y, lw.y);
            split_h2(a[4] * scale, a[5] * scale, hw.z, lw.z);
            split_h2(a[6] * scale, a[7] * scale, hw.w, lw.w);
            const size_t o = ((size_t)c * AU_M + fr) * 16;
            *reinterpret_cast<uint4 *>(ah + o) = hw;
            *reinterpret_cast<uint4 *>(al + o) = lw;
            if (c == 0) {
                // S = S' * 2^(e + E);  value = |c|^2 - 2 S.  Error of S' (DESIGN.md section 4, K3):
                //   operands  3 * 2^-22 |x'| |c'|  (h + l is x' to 2^-22, the l cl product is dropped)
                //             + 2^-23 d            (fp16 subnormal low parts, |x'|, |c'| <= 2)
                //   fp32 TMEM accumulation, one truncation per UMMA: (3 d / 16) 2^-22 |x'| |c'|
                // and of the value: 2 * 2^(e + E) * that, + 2^-24 |c|^2 (float |c|^2) + the fma's own
                // rounding; two values are compared, so the margin is twice the sum (x 1.5 for slack).
                const float unit = inv * c_inv;
                f_mul[fr] = -2.f * unit;
                const float nxp = sqrtf(s2) * scale;
                const float coef = 2.3841858e-7f * (3.f + 3.f * (float)d_pad * 0.0625f);     // 2^-22 (3 + 3 d / 16)
                float eS = coef * nxp * c_nmax + 1.1920929e-7f * (float)d_pad;
                float mg = 3.f * (2.f * unit * eS + 1.7881393e-7f * (cn_max + 2.f * unit * nxp * c_nmax));
                if (!(m < INFINITY) || !(mg < INFINITY)) mg = INFINITY;      // non-finite frame: exact path
                f_margin[fr] = mg;
            }
        }
    }
}

struct AuSmem {
    uint64_t a_full[2];      // converters -> issuer (128 arrivals... one per converter warp)
    uint64_t a_empty[2];     // issuer (commit) -> converters
    uint64_t acc_full[2];    // issuer (commit) -> epilogue
    uint64_t acc_empty[2];   // epilogue (4 warps) -> issuer
    uint32_t tmem_base;
    // per frame, ring of 4 tiles (the converters run at most two tiles ahead of the epilogue, whose
    // read of tile t is over before the UMMAs of tile t + 2 -- hence the conversion of t + 4 -- start)
    float f_mul[4][AU_M];    // -2 * 2^(e_i + E)
    float f_margin[4][AU_M]; // ambiguity margin in |c|^2 - 2 S units
};

__global__ void __launch_bounds__(AU_THREADS, 1)
assign_umma_kernel(const float *__restrict__ X, long long n, int d, long long ld, int d_pad, int k,
                   int k_pad, const unsigned char *__restrict__ tiles_h,
                   const unsigned char *__restrict__ tiles_l, const float *__restrict__ cn,
                   const AuPrep *__restrict__ prep, int *__restrict__ labels,
                   int *__restrict__ amb_list, int *__restrict__ amb_count)
{
    extern __shared__ __align__(1024) unsigned char au_smem[];
    const int nc = d_pad / 8;                               // 16-byte K chunks per row
    const int n_nt = k_pad / AU_NT + ((k_pad % AU_NT) ? 1 : 0);
    const uint32_t a_tile = (uint32_t)AU_M * d_pad * 2;     // bytes of one fp16 component of a frame tile
    const uint32_t b_tile = (uint32_t)AU_NT * d_pad * 2;    // bytes of one component of one centre tile
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(au_smem) + 1023) & ~(uintptr_t)1023);
    unsigned char *sB_h = base;                                       // [n_nt][nc][256][16]
    unsigned char *sB_l = sB_h + (size_t)n_nt * b_tile;
    unsigned char *sA = sB_l + (size_t)n_nt * b_tile;                 // [2 stages][h | l][nc][128][16]
    float *s_cn = reinterpret_cast<float *>(sA + 4 * (size_t)a_tile); // [n_nt * 256]
    AuSmem *ctl = reinterpret_cast<AuSmem *>(s_cn + (size_t)n_nt * AU_NT);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            bar_init(&ctl->a_full[s], AU_CONV_WARPS);
            bar_init(&ctl->a_empty[s], 1);
            bar_init(&ctl->acc_full[s], 1);
            bar_init(&ctl->acc_empty[s], AU_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // the centres' tiles and squared norms: resident for the whole kernel
    {
        const uint4 *gh = reinterpret_cast<const uint4 *>(tiles_h), *gl = reinterpret_cast<const uint4 *>(tiles_l);
        uint4 *dh = reinterpret_cast<uint4 *>(sB_h), *dl = reinterpret_cast<uint4 *>(sB_l);
        const int total = n_nt * (int)(b_tile / 16);
        for (int i = tid; i < total; i += AU_THREADS) { dh[i] = gh[i]; dl[i] = gl[i]; }
        for (int i = tid; i < n_nt * AU_NT; i += AU_THREADS) s_cn[i] = i < k_pad ? cn[i] : INFINITY;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(s_u32(&ctl->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    const long long n_tiles = (n + AU_M - 1) / AU_M;
    const long long my_first = blockIdx.x, tile_step = gridDim.x;
    const float c_inv = prep->c_inv_scale, c_nmax = prep->c_norm_max, cn_max = prep->cn_max;

    if (warp == 1) {
        // ================================ UMMA issuer
        const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
        const uint32_t a_addr = __shfl_sync(0xffffffffu, s_u32(sA), 0);
        const uint32_t bh_addr = __shfl_sync(0xffffffffu, s_u32(sB_h), 0);
        const uint32_t bl_addr = __shfl_sync(0xffffffffu, s_u32(sB_l), 0);
        long long it = 0;                                   // frame tiles done by this CTA
        long long acc_it = 0;                               // accumulator buffers handed out
        for (long long t = my_first; t < n_tiles; t += tile_step, ++it) {
            const int st = (int)(it & 1);
            bar_wait(&ctl->a_full[st], (uint32_t)((it >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int nt = 0; nt < n_nt; ++nt, ++acc_it) {
                const int ab = (int)(acc_it & 1);
                bar_wait(&ctl->acc_empty[ab], (uint32_t)(((acc_it >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;");
                const int n_cols = min(AU_NT, k_pad - nt * AU_NT);
                uint32_t idesc = 0;
                idesc |= 1u << 4;                                   // f32 accumulate, f16 x f16
                idesc |= (uint32_t)(n_cols >> 3) << 17;
                idesc |= (uint32_t)(AU_M >> 4) << 24;
                if (elect_one()) {
                    const uint32_t ah = a_addr + (uint32_t)st * 2 * a_tile, al = ah + a_tile;
                    const uint32_t bh = bh_addr + (uint32_t)nt * b_tile, bl = bl_addr + (uint32_t)nt * b_tile;
                    for (int ks = 0; ks < d_pad / 16; ++ks) {
                        // K step = two 16-byte chunks: A chunks are 128 rows * 16 B apart, B chunks 256 * 16 B
                        const uint64_t dAh = kdesc(ah + (uint32_t)ks * 2 * (AU_M * 16), AU_M * 16, 128);
                        const uint64_t dAl = kdesc(al + (uint32_t)ks * 2 * (AU_M * 16), AU_M * 16, 128);
                        const uint64_t dBh = kdesc(bh + (uint32_t)ks * 2 * (AU_NT * 16), AU_NT * 16, 128);
                        const uint64_t dBl = kdesc(bl + (uint32_t)ks * 2 * (AU_NT * 16), AU_NT * 16, 128);
                        mma_f16(tmem + (uint32_t)ab * AU_NT, dAh, dBh, idesc, ks ? 1u : 0u);
                        mma_f16(tmem + (uint32_t)ab * AU_NT, dAh, dBl, idesc, 1u);
                        mma_f16(tmem + (uint32_t)ab * AU_NT, dAl, dBh, idesc, 1u);
                    }
                    mma_commit(&ctl->acc_full[ab]);
                    if (nt == n_nt - 1) mma_commit(&ctl->a_empty[st]);     // the stage may be refilled
                }
                __syncwarp();
            }
        }
    } else if (warp >= AU_FIRST_CONV && warp < AU_FIRST_EPI) {
        // ================================ converters (256 threads): au_convert_tile
        const int ct = tid - 32 * AU_FIRST_CONV;
        long long it = 0;
        for (long long t = my_first; t < n_tiles; t += tile_step, ++it) {
            const int st = (int)(it & 1);
            bar_wait(&ctl->a_empty[st], (uint32_t)(((it >> 1) & 1) ^ 1));
            unsigned char *ah = sA + (size_t)st * 2 * a_tile, *al = ah + a_tile;
            au_convert_tile(X, n, d, ld, nc, d_pad, t * AU_M, ah, al, ct, ctl->f_mul[it & 3], ctl->f_margin[it & 3],
                            c_inv, c_nmax, cn_max);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(&ctl->a_full[st]);
        }
    } else if (warp >= AU_FIRST_EPI && warp < AU_FIRST_EPI + AU_EPI_WARPS) {
        // ================================ epilogue: lane = frame, scan the centres of every accumulator
        const int quarter = warp & 3;
        const int fr = quarter * 32 + lane;
        const uint32_t tmem = ctl->tmem_base;
        long long it = 0, acc_it = 0;
        for (long long t = my_first; t < n_tiles; t += tile_step, ++it) {
            float best = INFINITY, second = INFINITY;
            int arg = 0;
            float mul = 0.f, margin = 0.f;
            for (int nt = 0; nt < n_nt; ++nt, ++acc_it) {
                const int ab = (int)(acc_it & 1);
                bar_wait(&ctl->acc_full[ab], (uint32_t)((acc_it >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (nt == 0) { mul = ctl->f_mul[it & 3][fr]; margin = ctl->f_margin[it & 3][fr]; }
                const int n_cols = min(AU_NT, k_pad - nt * AU_NT);
                for (int c0 = 0; c0 < n_cols; c0 += 32) {
                    uint32_t v[32];
                    AU_TMEM_LD32(v, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * AU_NT + c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float *cnp = s_cn + nt * AU_NT + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float val = fmaf(__uint_as_float(v[j]), mul, cnp[j]);
                        if (val < best) { second = best; best = val; arg = nt * AU_NT + c0 + j; }
                        else if (val < second) second = val;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) bar_arrive(&ctl->acc_empty[ab]);
            }
            const long long row = t * AU_M + fr;
            if (row < n) {
                labels[row] = arg;
                // inside the error bound (or a tie, or NaN): exact re-scan
                if (!(second - best > margin)) {
                    const int slot = atomicAdd(amb_count, 1);
                    amb_list[slot] = (int)row;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(ctl->tmem_base), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// Streamed centres: the same three-product filter when the centre table does not fit shared memory.
//   warp 14    loader (one lane): chunk ch of the centres' h / l tiles -> ring slot, two TMA bulk copies
//              (the tiles of a chunk are contiguous in global memory), completion on b_full[slot];
//   warp 1     issuer: per frame tile, per chunk: wait b_full + acc_empty, d_pad / 16 K steps of three
//              UMMAs (M = 128 frames, N = NS centres), commit -> acc_full, b_empty (and a_empty after
//              the tile's last chunk);
//   warps 2-9  converters, warps 10-13 epilogue: as in assign_umma_kernel.
// NS (centres per chunk) and the number of frame-operand stages follow the shared-memory budget:
// d_pad 256 -> NS 32, 128 -> 128, <= 64 -> 256; the frame operand is single buffered above d_pad 64
// (the UMMAs of a tile then take several times longer than its conversion).  |c_j|^2 of every centre
// stays resident (4 k bytes).
// ------------------------------------------------------------------------------------------------
struct AsSmem {
    uint64_t a_full[2], a_empty[2];
    uint64_t b_full[2], b_empty[2];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
    float f_mul[4][AU_M];
    float f_margin[4][AU_M];
};
constexpr int AS_LOADER_WARP = AU_FIRST_EPI + AU_EPI_WARPS;          // 14

__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(s_u32(dst)), "l"(src), "r"(bytes), "r"(s_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(AU_THREADS, 1)
assign_umma_stream_kernel(const float *__restrict__ X, long long n, int d, long long ld, int d_pad,
                          int n_ch, int ns, int a_stages, const unsigned char *__restrict__ tiles_h,
                          const unsigned char *__restrict__ tiles_l, const float *__restrict__ cn,
                          const AuPrep *__restrict__ prep, int *__restrict__ labels,
                          int *__restrict__ amb_list, int *__restrict__ amb_count)
{
    extern __shared__ __align__(1024) unsigned char au_smem[];
    const int nc = d_pad / 8;                               // 16-byte K chunks per row
    const uint32_t a_tile = (uint32_t)AU_M * d_pad * 2;     // bytes of one fp16 component of a frame tile
    const uint32_t b_tile = (uint32_t)ns * d_pad * 2;       // bytes of one component of one centre chunk
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(au_smem) + 1023) & ~(uintptr_t)1023);
    unsigned char *sB = base;                                         // [2 slots][h | l][nc][ns][16]
    unsigned char *sA = sB + 4 * (size_t)b_tile;                      // [a_stages][h | l][nc][128][16]
    float *s_cn = reinterpret_cast<float *>(sA + (size_t)a_stages * 2 * a_tile);   // [n_ch * ns]
    AsSmem *ctl = reinterpret_cast<AsSmem *>(s_cn + (size_t)n_ch * ns);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            bar_init(&ctl->a_full[s], AU_CONV_WARPS);
            bar_init(&ctl->a_empty[s], 1);
            bar_init(&ctl->b_full[s], 1);
            bar_init(&ctl->b_empty[s], 1);
            bar_init(&ctl->acc_full[s], 1);
            bar_init(&ctl->acc_empty[s], AU_EPI_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int i = tid; i < n_ch * ns; i += AU_THREADS) s_cn[i] = cn[i];
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(s_u32(&ctl->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");

    const long long n_tiles = (n + AU_M - 1) / AU_M;
    const long long my_first = blockIdx.x, tile_step = gridDim.x;
    const float c_inv = prep->c_inv_scale, c_nmax = prep->c_norm_max, cn_max = prep->cn_max;

    if (warp == AS_LOADER_WARP) {
        // ================================ loader: centre chunks, once per frame tile (L2 resident)
        // (the whole warp walks the loop, one lane issues: no lane is left behind at the final barrier)
        long long bit = 0;
        for (long long t = my_first; t < n_tiles; t += tile_step) {
            for (int ch = 0; ch < n_ch; ++ch, ++bit) {
                const int slot = (int)(bit & 1);
                bar_wait(&ctl->b_empty[slot], (uint32_t)(((bit >> 1) & 1) ^ 1));
                if (lane == 0) {
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                                 :: "r"(s_u32(&ctl->b_full[slot])), "r"(2u * b_tile) : "memory");
                    unsigned char *dst = sB + (size_t)slot * 2 * b_tile;
                    bulk_g2s(dst, tiles_h + (size_t)ch * b_tile, b_tile, &ctl->b_full[slot]);
                    bulk_g2s(dst + b_tile, tiles_l + (size_t)ch * b_tile, b_tile, &ctl->b_full[slot]);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ================================ UMMA issuer
        const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
        const uint32_t a_addr = __shfl_sync(0xffffffffu, s_u32(sA), 0);
        const uint32_t b_addr = __shfl_sync(0xffffffffu, s_u32(sB), 0);
        uint32_t idesc = 0;
        idesc |= 1u << 4;                                   // f32 accumulate, f16 x f16
        idesc |= (uint32_t)(ns >> 3) << 17;
        idesc |= (uint32_t)(AU_M >> 4) << 24;
        long long it = 0, cit = 0;                          // frame tiles / chunks done by this CTA
        for (long long t = my_first; t < n_tiles; t += tile_step, ++it) {
            const int st = (int)(it % a_stages);
            bar_wait(&ctl->a_full[st], (uint32_t)((it / a_stages) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int ch = 0; ch < n_ch; ++ch, ++cit) {
                const int ab = (int)(cit & 1);              // accumulator buffer and ring slot advance together
                bar_wait(&ctl->b_full[ab], (uint32_t)((cit >> 1) & 1));
                bar_wait(&ctl->acc_empty[ab], (uint32_t)(((cit >> 1) & 1) ^ 1));
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (elect_one()) {
                    const uint32_t ah = a_addr + (uint32_t)st * 2 * a_tile, al = ah + a_tile;
                    const uint32_t bh = b_addr + (uint32_t)ab * 2 * b_tile, bl = bh + b_tile;
                    for (int ks = 0; ks < d_pad / 16; ++ks) {
                        const uint64_t dAh = kdesc(ah + (uint32_t)ks * 2 * (AU_M * 16), AU_M * 16, 128);
                        const uint64_t dAl = kdesc(al + (uint32_t)ks * 2 * (AU_M * 16), AU_M * 16, 128);
                        const uint64_t dBh = kdesc(bh + (uint32_t)ks * 2 * ((uint32_t)ns * 16), (uint32_t)ns * 16, 128);
                        const uint64_t dBl = kdesc(bl + (uint32_t)ks * 2 * ((uint32_t)ns * 16), (uint32_t)ns * 16, 128);
                        mma_f16(tmem + (uint32_t)ab * (uint32_t)ns, dAh, dBh, idesc, ks ? 1u : 0u);
                        mma_f16(tmem + (uint32_t)ab * (uint32_t)ns, dAh, dBl, idesc, 1u);
                        mma_f16(tmem + (uint32_t)ab * (uint32_t)ns, dAl, dBh, idesc, 1u);
                    }
                    mma_commit(&ctl->acc_full[ab]);
                    mma_commit(&ctl->b_empty[ab]);                          // the slot may be refilled
                    if (ch == n_ch - 1) mma_commit(&ctl->a_empty[st]);      // ... and the frame stage
                }
                __syncwarp();
            }
        }
    } else if (warp >= AU_FIRST_CONV && warp < AU_FIRST_EPI) {
        // ================================ converters (256 threads): au_convert_tile
        const int ct = tid - 32 * AU_FIRST_CONV;
        long long it = 0;
        for (long long t = my_first; t < n_tiles; t += tile_step, ++it) {
            const int st = (int)(it % a_stages);
            bar_wait(&ctl->a_empty[st], (uint32_t)(((it / a_stages) & 1) ^ 1));
            unsigned char *ah = sA + (size_t)st * 2 * a_tile, *al = ah + a_tile;
            au_convert_tile(X, n, d, ld, nc, d_pad, t * AU_M, ah, al, ct, ctl->f_mul[it & 3], ctl->f_margin[it & 3],
                            c_inv, c_nmax, cn_max);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) bar_arrive(&ctl->a_full[st]);
        }
    } else if (warp >= AU_FIRST_EPI && warp < AU_FIRST_EPI + AU_EPI_WARPS) {
        // ================================ epilogue: lane = frame, scan the centres of every chunk
        const int quarter = warp & 3;
        const int fr = quarter * 32 + lane;
        const uint32_t tmem = ctl->tmem_base;
        long long it = 0, cit = 0;
        for (long long t = my_first; t < n_tiles; t += tile_step, ++it) {
            float best = INFINITY, second = INFINITY;
            int arg = 0;
            float mul = 0.f, margin = 0.f;
            for (int ch = 0; ch < n_ch; ++ch, ++cit) {
                const int ab = (int)(cit & 1);
                bar_wait(&ctl->acc_full[ab], (uint32_t)((cit >> 1) & 1));
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (ch == 0) { mul = ctl->f_mul[it & 3][fr]; margin = ctl->f_margin[it & 3][fr]; }
                for (int c0 = 0; c0 < ns; c0 += 32) {
                    uint32_t v[32];
                    AU_TMEM_LD32(v, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(ab * ns + c0));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float *cnp = s_cn + ch * ns + c0;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float val = fmaf(__uint_as_float(v[j]), mul, cnp[j]);
                        if (val < best) { second = best; best = val; arg = ch * ns + c0 + j; }
                        else if (val < second) second = val;
                    }
                }
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0) bar_arrive(&ctl->acc_empty[ab]);
            }
            const long long row = t * AU_M + fr;
            if (row < n) {
                labels[row] = arg;
                if (!(second - best > margin)) {
                    const int slot = atomicAdd(amb_count, 1);
                    amb_list[slot] = (int)row;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(ctl->tmem_base), "r"(512));
}

// ------------------------------------------------------------------------------------------------
static int next_pow2_16(int d)
{
    int p = 16;
    while (p < d) p <<= 1;
    return p;
}

// which kernel takes (d, k): resident centre tiles, streamed chunks, or neither (-> SIMT filter)
struct AuPlan {
    bool ok, stream;
    int d_pad, k_pad;        // k_pad: padded centre count of the tile table (n_nt * 256 or n_ch * ns)
    int rows;                // rows per tile of the table (256 resident, ns streamed)
    int n_tiles;             // n_nt or n_ch
    int a_stages;
    size_t smem;
};
static AuPlan au_plan(int d, int k)
{
    AuPlan p = {};
    p.d_pad = next_pow2_16(d);
    const int k16 = (k + 15) / 16 * 16;
    const int n_nt = (k16 + AU_NT - 1) / AU_NT;
    const size_t a_tile = (size_t)AU_M * p.d_pad * 2;
    if ((size_t)n_nt * AU_NT * p.d_pad * 4 <= AU_B_LIMIT && !getenv("MSMB200_ASSIGN_STREAM")) {
        p.ok = true;
        p.stream = false;
        p.rows = AU_NT;
        p.n_tiles = n_nt;
        p.k_pad = k16;
        p.a_stages = 2;
        p.smem = 1024 + 2 * (size_t)n_nt * AU_NT * p.d_pad * 2 + 4 * a_tile + sizeof(float) * (size_t)n_nt * AU_NT
                 + sizeof(AuSmem) + 64;
        // all 512 TMEM columns belong to one CTA: ask for more than half of the shared memory so that a
        // second CTA is never scheduled on the same SM (its tcgen05.alloc would wait for the first to end)
        if (p.smem < 120 * 1024) p.smem = 120 * 1024;
        return p;
    }
    p.stream = true;
    p.rows = p.d_pad >= 256 ? 32 : (p.d_pad == 128 ? 128 : 256);
    p.n_tiles = (k + p.rows - 1) / p.rows;
    p.k_pad = p.n_tiles * p.rows;
    p.a_stages = p.d_pad <= 64 ? 2 : 1;
    p.smem = 1024 + 4 * (size_t)p.rows * p.d_pad * 2 + (size_t)p.a_stages * 2 * a_tile
             + sizeof(float) * (size_t)p.k_pad + sizeof(AsSmem) + 64;
    if (p.smem < 120 * 1024) p.smem = 120 * 1024;
    p.ok = p.smem <= 225 * 1024;
    return p;
}

bool assign_umma_supported(int64_t n_out, int d, int64_t ld, int k, const void *X, bool has_rows)
{
    if (has_rows || getenv("MSMB200_ASSIGN_SIMT")) return false;
    if (d > 256 || (d % 4) != 0 || k < 2 || (ld % 4) != 0 || (reinterpret_cast<uintptr_t>(X) & 15u)) return false;
    if (n_out < 4096) return false;                         // latency bound below that: the SIMT filter is fine
    return au_plan(d, k).ok;
}

// 1 when the streamed-centres kernel would take (d, k), 0 for the resident one, -1 for neither
int assign_umma_mode(int d, int k)
{
    const AuPlan p = au_plan(d, k);
    return !p.ok ? -1 : (p.stream ? 1 : 0);
}

size_t assign_umma_scratch_bytes(int d, int k)
{
    const AuPlan p = au_plan(d, k);
    const size_t rows = (size_t)p.n_tiles * p.rows;
    return 2 * rows * p.d_pad * 2 + sizeof(float) * rows + 256;
}

// filter stage on the tensor cores: labels + ambiguity list (amb_count zeroed by the caller)
int assign_umma_filter(const float *X, int64_t n, int d, int64_t ld, const float *Y, int k,
                       int32_t *labels, int *amb_list, int *amb_count, cudaStream_t st)
{
    const AuPlan p = au_plan(d, k);
    MSMB_REQUIRE(p.ok, "assign_umma_filter: unsupported shape (d=%d k=%d)", d, k);
    const size_t rows = (size_t)p.n_tiles * p.rows;
    const size_t comp_bytes = rows * p.d_pad * 2;            // one fp16 component of the whole table
    unsigned char *scratch = nullptr;
    MSMB_CUDA(cudaMallocAsync(&scratch, assign_umma_scratch_bytes(d, k), st));
    unsigned char *tiles_h = scratch, *tiles_l = scratch + comp_bytes;
    float *cn = reinterpret_cast<float *>(tiles_l + comp_bytes);
    AuPrep *prep = reinterpret_cast<AuPrep *>(cn + rows);
    assign_umma_prep_kernel<<<1, 256, 0, st>>>(Y, k, d, p.d_pad, (int)rows, p.rows, tiles_h, tiles_l, cn, prep);
    MSMB_LAUNCH_CHECK();
    static bool attr_set[64] = {false};
    int dev = 0;
    MSMB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        MSMB_CUDA(cudaFuncSetAttribute(assign_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        MSMB_CUDA(cudaFuncSetAttribute(assign_umma_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[dev] = true;
    }
    const long long n_tiles = (n + AU_M - 1) / AU_M;
    long long grid = sm_count();
    if (grid > n_tiles) grid = n_tiles;
    if (p.stream)
        assign_umma_stream_kernel<<<(unsigned)grid, AU_THREADS, p.smem, st>>>(
            X, n, d, ld, p.d_pad, p.n_tiles, p.rows, p.a_stages, tiles_h, tiles_l, cn, prep, labels, amb_list,
            amb_count);
    else
        assign_umma_kernel<<<(unsigned)grid, AU_THREADS, p.smem, st>>>(
            X, n, d, ld, p.d_pad, k, p.k_pad, tiles_h, tiles_l, cn, prep, labels, amb_list, amb_count);
    MSMB_LAUNCH_CHECK();
    MSMB_CUDA(cudaFreeAsync(scratch, st));
    return MSMB200_OK;
}

}  // namespace msmb
