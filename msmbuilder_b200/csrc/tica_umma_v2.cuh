// tica_umma_v2.cuh -- K1, second generation of the fp16 3-product engine (included by tica_umma.cu).
//
// Same mathematics as tica_umma_kernel<UM_KIND_F16> (centred, power-of-two scaled frame x' = h + l in
// fp16; raw moments rebuilt in float64 by the finalize kernel), reorganised around what the round-1
// profile showed (tensor pipe 75 % active, 25 % lost to the TMEM drain, the
// converters as slow as the MMAs, narrow inputs paying 256-wide tiles):
//
//  * 5 products instead of 6.  C_00 = sum x x^T is symmetric, so only G = h (h/2)^T + h l^T is
//    accumulated and C_00 = G + G^T (= h h^T + h l^T + l h^T) is formed by the finalize kernel.
//    C_tau keeps its three products h b^T + h bl^T + l b^T.  Operand tiles per stage:
//    a = h, al = l, ah = h / 2 (exact), b, bl  (8 KB each, K-major, no swizzle).
//  * The drain is hidden.  TMEM holds TWO accumulator regions, C_tau and G (N columns each: every UMMA
//    is the full M x N, the only shape that amortises the ~40 cycles a tcgen05.mma costs on top of
//    N / 2, profiles/r2d_probe5_mma_shapes.log).  Their slab boundaries are half a slab apart, so
//    at most one region is being drained at a time; its MMAs are deferred (the operand ring is 4
//    deep) while the tensor pipe keeps working on the other one, and the region catches up as soon
//    as its accumulators have been read out.
//  * Dedicated drain warps (8: two per TMEM lane quarter, half of a region's columns each), so the
//    converters never stop: tcgen05.ld of 64 columns at a time -> fire-and-forget
//    red.global.add.f32 into this CTA's float32 level (L2 resident); the region is released as soon
//    as its last columns are in registers.  After every drain the warp moves ONE 16-column group
//    from the float32 level into a float-float (hi, lo) pair with an error-free TwoSum (round
//    robin: a group is folded every 8 (4) slabs), so the float32 level never sums more than that
//    and no drain event takes long.  The finalize path adds hi + lo + level in float64.  No FP64
//    instruction runs in this kernel while the tensor pipe is busy (it stalls for hundreds of
//    cycles there, profiles/r1_k1_issue.txt).
//  * Converters: a thread owns one feature and the 8 frames of one K chunk of one operand; packed
//    f32x2 math (FFMA2 / FADD2 of sm_100), h = cvt.rn.f16x2, l = fp16(x' - h): ~40 instructions per
//    8 values instead of ~70.  Only the feature blocks that exist are converted (D < 128 per CTA).
//  * template <CG>: CG = 2 is the CTA-pair kernel (cta_group::2, M = 256, N = 256, D in (128, 256]);
//    CG = 1 is the single-CTA kernel for D <= 128 (cta_group::1, M = 128, N = 128 or 64, 148
//    independent CTAs): narrow inputs no longer pay 256-wide tiles (config 2, 10M x 64).
//  * Bit-reproducible like v1: every address has one writer, all sums are taken in a fixed order.
//
// MN-major mode (P.mn; D = 256, 128 or 64, lag <= 32; profiles/r2z_probe7_mn_major_sw128.log).
// The K-major chunks above (8 frames of one feature per 16 bytes) cannot serve both operands: the lagged
// operand's chunks start `lag` frames later, so every frame is TMA-loaded and converted twice and the
// gather that builds a chunk costs eight 4-byte shared loads.  tcgen05 also takes MN-major fp16 operands
// (features contiguous, one frame = one 128-byte row of 64 features, SWIZZLE_128B, LBO = distance of the
// next 64 features, SBO = 8 rows), and the swizzle is a function of the absolute shared-memory address,
// so a descriptor may start at ANY row: the lagged operand is the same converted buffer, `lag` rows on.
// Per CTA three windows H (h), L (l), Q (h / 2) of [2 blocks of 64 features][S ring tiles + 1 mirror tile
// of 32 rows][128 B] (S = 5: 144 KB, inside the operand ring of the K-major mode); a frame is loaded once and
// converted once: two 16-byte loads of the raw row, three 16-byte stores.  Tile t's UMMAs read rows
// [32 t, 32 t + 32 + lag): they wait for ring tiles t and t + 1, a group converts one tile more than it
// multiplies, ring tile 0 is mirrored behind the last one.  Tiles cover ALL rows of a sequence plus `lag` rows
// of zeros (TMA fills them), so G sums x x^T over every row and the finalize kernel takes the tail rows
// out of C_00 and the head rows out of C_tautau (the float64 edge terms it already has).
#pragma once

namespace msmb {

constexpr int V2_TILE = UM_KT * UM_F * 2;           // one fp16 component tile: 8 KB
constexpr int V2_NTILES = 5;                        // a, al, ah, b, bl
constexpr int V2_STAGE_BYTES = V2_NTILES * V2_TILE; // 40 KB
constexpr int V2_RAW_STAGES = 2;                    // raw ring depth with full 4-block boxes (64 KB in all)
constexpr int V2_RAW_MAX = 8;                       // ... 4 / 8 stages when a box only brings 2 / 1 feature blocks
constexpr int V2_OP_STAGES = 4;
constexpr int V2_MAX_STAGES = 6;                    // barrier slots (the MN-major mode runs 4 or 5 ring tiles)
constexpr int V2_CONV_WARPS = 8;
constexpr int V2_DRAIN_WARPS = 8;
constexpr int V2_UNITS_PER_WARP = 4;                // 8 * 4 feature blocks = 32 units per tile at most
// w0 TMA producer (+ TMEM allocation), w1 MMA issuer, w2-3 idle, w4-11 converters, w12-19 drain.
// Registers are allocated to warps in groups of 4: 20 warps = 640 threads at 96 registers.
constexpr int V2_FIRST_CONV_WARP = 4;
constexpr int V2_FIRST_DRAIN_WARP = V2_FIRST_CONV_WARP + V2_CONV_WARPS;   // 12..19: two warps per lane quarter
constexpr int V2_THREADS = 32 * (V2_FIRST_DRAIN_WARP + V2_DRAIN_WARPS);   // 640
constexpr int V2_CONV_TID0 = 32 * V2_FIRST_CONV_WARP;
constexpr int V2_REGIONS = 2;                       // C_tau, G
constexpr int V2_T_A = 0, V2_T_AL = 1, V2_T_AH = 2, V2_T_B = 3, V2_T_BL = 4;
constexpr int V2_MAX_GROUPS = 192;                  // CTA pairs (CG = 2) or CTAs (CG = 1)
constexpr uint32_t V2_RANGE_LIMIT = 0x5400u;        // fp16 bit pattern of 64.0 = 2^6

struct V2Params {
    const CUtensorMap *mapsA;     // [n_seq] unlagged
    const CUtensorMap *mapsB;     // [n_seq] base shifted by lag rows
    const int *tile_prefix;       // [n_seq + 1]
    const int *seq_pairs;         // [n_seq]
    int n_seq;
    int n_tiles;
    int n_groups;                 // CTA pairs (CG = 2) or CTAs (CG = 1)
    int slab_tiles;               // 32-frame tiles per TMEM slab
    int D;                        // real feature count (32k)
    int box_blocks;               // 32-feature blocks one TMA box brings (4; D / 32 for the single-CTA kernel)
    int dbg_mode;                 // 1: converters skip their work, 2: no drain (timing experiments)
    int mn;                       // 0, or the number of ring tiles (4 / 5) of the MN-major rolling-window mode
    int lag;                      // (MN-major mode) rows between the two operands
    const float *shift;           // [UM_D]
    const float *scale;           // [UM_D]
    int *overflow;                // set to 1 when a scaled value reached 2^6 (or was not finite): float64 rescue
    float *lvl1;                  // [n_ctas][2][RW][128] float32 level (red.add target), RW = 128 * CG
    float *hi, *lo;               // same shape: float-float second level
    double *sums;                 // [n_groups][UM_D] column sums of x' (unscaled)
    long long *dbg;
};

struct V2Smem {
    uint64_t raw_full[V2_RAW_MAX];
    uint64_t raw_empty[V2_RAW_MAX];
    uint64_t conv[V2_MAX_STAGES];      // leader's copy is used; CG arrivals
    uint64_t empty[V2_MAX_STAGES];     // local; one commit arrival
    uint64_t acc_full[V2_REGIONS];     // local; one commit arrival
    uint64_t acc_empty[V2_REGIONS];    // leader's copy is used; 8 * CG arrivals
    uint32_t tmem_base;
    int valid_rows[V2_RAW_MAX];
    float sc[UM_F];                    // this CTA's per-feature scale and -shift * scale
    float nsh[UM_F];
};

__device__ __forceinline__ uint32_t mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return __shfl_sync(0xffffffffu, ok, 0);          // one answer for the whole warp
}
// wait of a warp that has nothing else to do for thousands of cycles (drain warps, the TMA producer):
// poll with a suspend hint, then sleep -- a plain try_wait loop issues BRA / SYNCS / YIELD around the
// clock and takes issue slots from the converters of the same scheduler (r2i profile: 5.7 G of the
// kernel's 10.7 G warp instructions were such spins)
template <bool SLEEP = true>
__device__ __forceinline__ void mbar_wait_idle(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(addr), "r"(parity), "r"(1000u) : "memory");
        if (ok) break;
        if (SLEEP) __nanosleep(64);
    }
}
__device__ __forceinline__ void mbar_arrive_local(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}

// One UMMA (kind::f16, fp16 x fp16 -> fp32).  MODE picks the collector hint for the A operand:
// 0 none, 1 fill, 2 use, 3 lastuse -- the caller strings together the MMAs that share A.
#define V2_MMA_ASM(CGS, VEC, QUAL) asm volatile( \
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" \
    "tcgen05.mma.cta_group::" CGS ".kind::f16" QUAL " [%0], %1, %2, %3, " VEC ", p;\n\t}" \
    :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory")
#define V2_VEC8 "{%5, %5, %5, %5, %5, %5, %5, %5}"
#define V2_VEC4 "{%5, %5, %5, %5}"
template <int CG, int MODE>
__device__ __forceinline__ void v2_mma(uint32_t tmem_d, uint64_t da, uint64_t db,
                                       uint32_t idesc, uint32_t acc)
{
    if constexpr (CG == 2) {
        if constexpr (MODE == 1) V2_MMA_ASM("2", V2_VEC8, ".collector::a::fill");
        else if constexpr (MODE == 2) V2_MMA_ASM("2", V2_VEC8, ".collector::a::use");
        else if constexpr (MODE == 3) V2_MMA_ASM("2", V2_VEC8, ".collector::a::lastuse");
        else V2_MMA_ASM("2", V2_VEC8, "");
    } else {
        if constexpr (MODE == 1) V2_MMA_ASM("1", V2_VEC4, ".collector::a::fill");
        else if constexpr (MODE == 2) V2_MMA_ASM("1", V2_VEC4, ".collector::a::use");
        else if constexpr (MODE == 3) V2_MMA_ASM("1", V2_VEC4, ".collector::a::lastuse");
        else V2_MMA_ASM("1", V2_VEC4, "");
    }
}
// The MMA warp runs its bookkeeping warp-uniformly and issues the tcgen05 instructions from ONE lane
// inside `if (elect_one_sync())`: nvcc recognises this form (it is CUTLASS's), knows that a single
// lane is active in the region and moves the operands to uniform registers without a waterfall.  A
// per-instruction lane predicate instead gave an ELECT / BRA.U.ANY loop around every UTCHMMA (~120
// cycles of issue per MMA, r2c profile).
__device__ __forceinline__ uint32_t elect_one_sync()
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
// arrive on `bar` (in every CTA of the group) when all MMAs issued so far have completed
template <int CG>
__device__ __forceinline__ void v2_commit(uint64_t *bar)
{
    if constexpr (CG == 2) {
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
    } else {
        asm volatile(
            "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
            :: "r"(smem_u32(bar)) : "memory");
    }
}
// (fp16 x fp16) -> f32, K-major A and B, M = 128 * CG, N = n_cols
template <int CG>
__device__ __forceinline__ uint32_t v2_idesc(int n_cols)
{
    uint32_t d = 0;
    d |= 1u << 4;                                   // D format F32; A, B format F16 = 0
    d |= (uint32_t)(n_cols >> 3) << 17;             // N
    d |= (uint32_t)((128 * CG) >> 4) << 24;         // M
    return d;
}

// packed float32 pairs (sm_100: FFMA2 / FADD2, two elements per issue slot)
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// explicit shared-state-space accesses with 32-bit addresses: the ring pointers come out of an
// integer round-up, so the compiler only sees generic pointers (LD / ST instead of LDS / STS)
template <int IMM>
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
}
__device__ __forceinline__ float4 lds_v4f(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_v4r(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// MN-major windows: rows of 128 bytes (64 features); S ring tiles of 32 rows + the mirror of tile 0 behind them.
// S = 4 or 5 (5 fits the operand ring of the K-major mode: 3 windows x 2 blocks x 192 rows x 128 B = 144 KB)
__host__ __device__ constexpr int v3_block_bytes(int S) { return (32 * S + 32) * 128; }   // LBO: next 64 features
__host__ __device__ constexpr int v3_win_bytes(int S) { return 2 * v3_block_bytes(S); }   // H, L, Q follow each other
__host__ __device__ constexpr int v3_mirror(int S) { return 32 * S * 128; }               // ring tile 0 again
static_assert(3 * v3_win_bytes(5) <= V2_OP_STAGES * V2_STAGE_BYTES, "the windows live in the operand ring");

template <int IMM>
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    asm volatile("st.shared.v4.b32 [%0+%1], {%2, %3, %4, %5};"
                 :: "r"(addr), "n"(IMM), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// fp16 pair -> the two floats it holds (exact)
__device__ __forceinline__ uint64_t h2_to_f2(uint32_t h)
{
    float a, b;
    asm("{\n\t.reg .f16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f32.f16 %0, lo;\n\tcvt.f32.f16 %1, hi;\n\t}"
        : "=f"(a), "=f"(b) : "r"(h));
    return f2_pack(a, b);
}

// ---------------------------------------------------------------------------------------
// MN: 0 = K-major operand ring (4 stages), 4 / 5 = MN-major rolling window with that many ring tiles.
// (A template parameter: with the stage count a run-time value the issuer of the narrow kernel -- which is
// issue bound -- lost 35 %: config 2 went from 1.90 to 2.56 ms, profiles/r2f_bench_1gpu_runtime_stages.json.)
template <int CG, int MN>
__global__ void __launch_bounds__(V2_THREADS, 1) tica_umma_v2_kernel(const V2Params P)
{
    constexpr int RW = 128 * CG;                     // columns reserved per region (N <= RW)
    constexpr int TMEM_COLS = V2_REGIONS * RW;       // 512 (CG = 2) or 256 (CG = 1)
    constexpr bool mn = MN != 0;
    constexpr int S = mn ? MN : V2_OP_STAGES;        // operand stages / ring tiles
    constexpr int CW = RW / 2;                       // columns of a region one drain warp owns
    constexpr int NG = CW / 16;                      // its 16-column fold groups
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *ring = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *raw_ring = ring;                                        // [2][A raw | B raw]
    unsigned char *op_ring = ring + V2_RAW_STAGES * UM_RAW_BYTES;          // [4][a|al|ah|b|bl]
    V2Smem *ctl = reinterpret_cast<V2Smem *>(op_ring + V2_OP_STAGES * V2_STAGE_BYTES);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    uint32_t cta_rank = 0;
    if constexpr (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const int group = blockIdx.x / CG;
    const int cta_global = blockIdx.x;

    const long long t_begin = (long long)P.n_tiles * group / P.n_groups;
    const long long t_end = (long long)P.n_tiles * (group + 1) / P.n_groups;
    const int my_tiles = (int)(t_end - t_begin);
    const int ST = P.slab_tiles;
    // features of this CTA that exist, in blocks of 32 (TMA zero-fills the rest of the raw tile; the
    // operand rows of the missing blocks are zeroed once, below, and never written again)
    int d_local = P.D - UM_F * (int)cta_rank;
    d_local = d_local < 0 ? 0 : (d_local > UM_F ? UM_F : d_local);
    const int nfb = d_local / 32;
    // raw ring: a stage holds the two boxes of a tile (unlagged, lagged), box_blocks * 4 KB each; the
    // ring always spans 64 KB, so narrow inputs get a deeper ring (the TMA latency showed as 240 cycles
    // of wait per tile at D = 64 with two stages)
    const uint32_t raw_op_bytes = (uint32_t)P.box_blocks * (UM_KT * 128);
    const uint32_t raw_stage_bytes = mn ? raw_op_bytes : 2 * raw_op_bytes;      // MN-major: one box per tile
    // MN-major: one ring tile more is converted than multiplied (tile t's lagged rows reach into t + 1)
    const int conv_tiles = (mn && my_tiles > 0) ? my_tiles + 1 : my_tiles;
    const int n_raw = V2_RAW_STAGES * UM_RAW_BYTES / (int)raw_stage_bytes > V2_RAW_MAX
                          ? V2_RAW_MAX : V2_RAW_STAGES * UM_RAW_BYTES / (int)raw_stage_bytes;
    // N of every UMMA: all the columns the group has (a single CTA with <= 64 features runs N = 64)
    const int n_cols = CG == 2 ? 256 : (P.D > 64 ? 128 : 64);
    // slab boundaries: staggered across groups (so the drains of the whole chip do not coincide)
    // and by half a slab between the two regions of a group
    const int slab_off = (int)(((long long)group * ST) / P.n_groups);
    auto region_off = [&](int q) { return (slab_off + (q * ST) / V2_REGIONS) % ST; };

    if (tid == 0) {
        for (int s = 0; s < V2_RAW_MAX; ++s) {
            mbar_init(&ctl->raw_full[s], 1);
            mbar_init(&ctl->raw_empty[s], 1);
        }
        for (int s = 0; s < V2_MAX_STAGES; ++s) {
            mbar_init(&ctl->conv[s], CG);
            mbar_init(&ctl->empty[s], 1);
        }
        for (int q = 0; q < V2_REGIONS; ++q) {
            mbar_init(&ctl->acc_full[q], 1);
            mbar_init(&ctl->acc_empty[q], V2_DRAIN_WARPS * CG);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < UM_F) {
        const float sc = P.scale[UM_F * cta_rank + tid];
        ctl->sc[tid] = sc;
        ctl->nsh[tid] = -P.shift[UM_F * cta_rank + tid] * sc;      // exact (power-of-two scale)
    }
    if (nfb < 4) {
        // operand rows of feature blocks that do not exist: zero (an uninitialised row could hold a
        // NaN pattern and trip the range check through its accumulators)
        uint4 *p = reinterpret_cast<uint4 *>(op_ring);
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int i = tid; i < V2_OP_STAGES * V2_STAGE_BYTES / 16; i += V2_THREADS) p[i] = z;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 0) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(&ctl->tmem_base)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(&ctl->tmem_base)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = ctl->tmem_base;

    if (warp == 0) {
        // ================================ TMA producer (one lane, every CTA) =============
        if (lane == 0 && my_tiles > 0) {
            int s = 0;
            {
                int lo = 0, hi = P.n_seq;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (P.tile_prefix[mid] <= t_begin) lo = mid; else hi = mid;
                }
                s = lo;
            }
            int stage = 0;
            uint32_t phase = 0;
            const uint64_t policy = l2_evict_first_policy();
            for (long long t = t_begin; t < t_begin + conv_tiles; ++t) {
                if (t >= P.n_tiles) {
                    // (MN-major) behind the last tile of the call: a tile of zeros, nothing to load
                    mbar_wait_idle<false>(&ctl->raw_empty[stage], phase ^ 1);
                    ctl->valid_rows[stage] = 0;
                    mbar_arrive_local(&ctl->raw_full[stage]);
                    if (++stage == n_raw) { stage = 0; phase ^= 1; }
                    continue;
                }
                while (t >= P.tile_prefix[s + 1]) ++s;
                const int row0 = (int)(t - P.tile_prefix[s]) * UM_KT;
                int valid = P.seq_pairs[s] + (mn ? P.lag : 0) - row0;      // MN-major: all rows of the sequence
                if (valid > UM_KT) valid = UM_KT;
                if (valid < 0) valid = 0;
                mbar_wait_idle<false>(&ctl->raw_empty[stage], phase ^ 1);   // hint only: the TMA issue is latency critical
                ctl->valid_rows[stage] = valid;
                unsigned char *st = raw_ring + (size_t)stage * raw_stage_bytes;
                if constexpr (mn) {
                    mbar_expect_tx(&ctl->raw_full[stage], P.box_blocks * (UM_KT * 128));
                    tma_load_3d(st, &P.mapsA[s], &ctl->raw_full[stage], 0, row0, 4 * (int)cta_rank, policy);
                } else {
                    mbar_expect_tx(&ctl->raw_full[stage], 2 * P.box_blocks * (UM_KT * 128));
                    tma_load_3d(st, &P.mapsA[s], &ctl->raw_full[stage], 0, row0, 4 * (int)cta_rank, policy);
                    tma_load_3d(st + raw_op_bytes, &P.mapsB[s], &ctl->raw_full[stage], 0, row0,
                                4 * (int)cta_rank, policy);
                }
                if (++stage == n_raw) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA; warp-uniform, one elected lane issues)
        if (cta_rank == 0 && my_tiles > 0) {
            // (MN-major mode: bits 15 / 16 = A / B are MN-major)
            const uint32_t idesc = v2_idesc<CG>(n_cols) | (mn ? ((1u << 15) | (1u << 16)) : 0u);
            // provably warp-uniform operands (a value loaded from shared memory is not, to the compiler)
            const uint32_t tmem = __shfl_sync(0xffffffffu, ctl->tmem_base, 0);
            const uint32_t ring_addr = __shfl_sync(0xffffffffu, smem_u32(op_ring), 0);
            const bool dbg_on = P.dbg != nullptr && group == 0;
            const long long d_start = clock64();
            long long d_idle = 0;
            int d_deferred = 0, d_fast = 0;
            int nt[V2_REGIONS];                      // next tile of each region
            int nf[V2_REGIONS];                      // tile at which the region's next slab starts
            uint32_t acc_ph = 0;                     // phase bit of acc_empty[q] in bit q
#pragma unroll
            for (int q = 0; q < V2_REGIONS; ++q) {
                nt[q] = 0;
                nf[q] = ST - region_off(q);
            }
            // descriptor of (tile, byte offset) in the stage whose low word is `base_lo`: only the
            // 14-bit start-address field changes (shared addresses stay below 2^18: no carry)
            auto desc = [](uint32_t base_lo, uint32_t byte_off) -> uint64_t {
                uint64_t d;
                asm("mov.b64 %0, {%1, %2};" : "=l"(d)
                    : "r"(base_lo + (byte_off >> 4)), "r"((uint32_t)((UM_SBO >> 4) | (1u << 14))));
                return d;
            };
            // MN-major mode: window `win` (0 H, 1 L, 2 Q), row `row` of the ring: SWIZZLE_128B (layout type 2 in
            // bits 61-63), LBO = next 64 features, SBO = 8 rows; any row may start a descriptor
            const uint32_t mn_win = (uint32_t)v3_win_bytes(S), mn_lbo = (uint32_t)(v3_block_bytes(S) >> 4) << 16;
            auto desc_mn = [&](uint32_t ring, int win, int row) -> uint64_t {
                const uint32_t a = ring + (uint32_t)win * mn_win + (uint32_t)row * 128;
                uint64_t d;
                asm("mov.b64 %0, {%1, %2};" : "=l"(d)
                    : "r"(((a >> 4) & 0x3FFFu) | mn_lbo),
                      "r"((uint32_t)((1024 >> 4) | (1u << 14) | (2u << 29))));
                return d;
            };
            const int lag = P.lag;
            int conv_done = 0;                       // tiles whose conversion has been observed
            int released = 0;                        // tiles whose operand stage has been handed back
            while (released < my_tiles) {
                while (conv_done < conv_tiles && conv_done < released + S &&
                       mbar_test(&ctl->conv[conv_done % S], (uint32_t)((conv_done / S) & 1)))
                    ++conv_done;
                bool progressed = false;
                // every tile index some region is waiting at: batch the regions that can go
                // (MN-major mode: tile tt reads ring tiles tt and tt + 1)
                const int ready = mn ? conv_done - 1 : conv_done;
                for (int tt = released; tt < ready; ++tt) {
                    uint32_t mask = 0, firsts = 0;
#pragma unroll
                    for (int q = 0; q < V2_REGIONS; ++q) {
                        if (nt[q] != tt) continue;
                        if (tt == nf[q]) {
                            // a new slab: the region's previous slab must have left TMEM
                            if (!mbar_test(&ctl->acc_empty[q], (acc_ph >> q) & 1u)) {
                                ++d_deferred;
                                continue;
                            }
                            acc_ph ^= 1u << q;
                            nf[q] += ST;
                            firsts |= 1u << q;
                        }
                        mask |= 1u << q;
                    }
                    if (!mask) continue;
                    if (tt == 0) firsts = mask;
                    progressed = true;
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    const uint32_t st = ring_addr + (uint32_t)(tt % S) * V2_STAGE_BYTES;
                    const uint32_t base_lo = ((st >> 4) & 0x3FFFu) | ((uint32_t)((UM_LBO >> 4) & 0x3FFF) << 16);
                    if (mask == 3u) ++d_fast;
                    if (elect_one_sync()) {
                        const uint32_t f0 = (firsts & 1u) ? 0u : 1u, f1 = (firsts & 2u) ? 0u : 1u;
                        // operand descriptors of K step ks: A = h, A = l, B = b (lagged h), B = bl (lagged l),
                        // B = h / 2, B = l (both of the unlagged frames)
                        const int ring_row = (tt % S) * UM_KT;
                        auto opd = [&](int which, int ks) -> uint64_t {
                            if constexpr (mn) {
                                const int ra = ring_row + 16 * ks, rb = ra + lag;
                                switch (which) {
                                case 0: return desc_mn(ring_addr, 0, ra);
                                case 1: return desc_mn(ring_addr, 1, ra);
                                case 2: return desc_mn(ring_addr, 0, rb);
                                case 3: return desc_mn(ring_addr, 1, rb);
                                case 4: return desc_mn(ring_addr, 2, ra);
                                default: return desc_mn(ring_addr, 1, ra);
                                }
                            }
                            const uint32_t off = ks * 2 * UM_LBO;
                            switch (which) {
                            case 0: return desc(base_lo, V2_T_A * V2_TILE + off);
                            case 1: return desc(base_lo, V2_T_AL * V2_TILE + off);
                            case 2: return desc(base_lo, V2_T_B * V2_TILE + off);
                            case 3: return desc(base_lo, V2_T_BL * V2_TILE + off);
                            case 4: return desc(base_lo, V2_T_AH * V2_TILE + off);
                            default: return desc(base_lo, V2_T_AL * V2_TILE + off);
                            }
                        };
                        if (mask == 3u) {
                            // both regions: 5 UMMAs per K step, A = h kept in the collector for 4 of them
#pragma unroll
                            for (int ks = 0; ks < UM_KT / 16; ++ks) {
                                const uint64_t dA = opd(0, ks);
                                const uint64_t dB = opd(2, ks);
                                v2_mma<CG, 1>(tmem, dA, dB, idesc, ks == 0 ? f0 : 1u);
                                v2_mma<CG, 2>(tmem, dA, opd(3, ks), idesc, 1u);
                                v2_mma<CG, 2>(tmem + RW, dA, opd(4, ks), idesc, ks == 0 ? f1 : 1u);
                                v2_mma<CG, 3>(tmem + RW, dA, opd(5, ks), idesc, 1u);
                                v2_mma<CG, 0>(tmem, opd(1, ks), dB, idesc, 1u);
                            }
                        } else if (mask == 1u) {
                            // C_tau alone (G is being drained, or C_tau is catching up)
#pragma unroll
                            for (int ks = 0; ks < UM_KT / 16; ++ks) {
                                const uint64_t dA = opd(0, ks);
                                const uint64_t dB = opd(2, ks);
                                v2_mma<CG, 1>(tmem, dA, dB, idesc, ks == 0 ? f0 : 1u);
                                v2_mma<CG, 3>(tmem, dA, opd(3, ks), idesc, 1u);
                                v2_mma<CG, 0>(tmem, opd(1, ks), dB, idesc, 1u);
                            }
                        } else {
                            // G alone
#pragma unroll
                            for (int ks = 0; ks < UM_KT / 16; ++ks) {
                                const uint64_t dA = opd(0, ks);
                                v2_mma<CG, 1>(tmem + RW, dA, opd(4, ks), idesc, ks == 0 ? f1 : 1u);
                                v2_mma<CG, 3>(tmem + RW, dA, opd(5, ks), idesc, 1u);
                            }
                        }
#pragma unroll
                        for (int q = 0; q < V2_REGIONS; ++q)
                            if ((mask & (1u << q)) && (tt + 1 == nf[q] || tt + 1 == my_tiles))
                                v2_commit<CG>(&ctl->acc_full[q]);
                    }   // elected lane
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < V2_REGIONS; ++q)
                        if (mask & (1u << q)) nt[q] = tt + 1;
                    if (mask != 3u) break;           // re-evaluate from `released`: a deferred region comes first
                }
                const int low = nt[0] < nt[1] ? nt[0] : nt[1];
                if (released < low) {
                    if (elect_one_sync())
                        for (int r = released; r < low; ++r) v2_commit<CG>(&ctl->empty[r % S]);
                    __syncwarp();
                    released = low;
                }
                if (!progressed) {
                    if (dbg_on) { const long long c = clock64(); __nanosleep(20); d_idle += clock64() - c; }
                    else __nanosleep(20);
                }
            }
            if (dbg_on && lane == 0) {
                P.dbg[0] = clock64() - d_start;
                P.dbg[1] = d_idle;
                P.dbg[2] = d_deferred;
                P.dbg[3] = my_tiles;
                P.dbg[13] = d_fast;
            }
        }
    } else if (warp >= V2_FIRST_CONV_WARP && warp < V2_FIRST_DRAIN_WARP) {
        // ================================ converters (512 threads, every CTA) ==========
        // Work unit = (operand, K chunk of 8 frames, block of 32 features): 8 * nfb units per tile,
        // dealt to the 16 warps (unit cw and, when there are more than 16, unit cw + 16).  lane =
        // feature inside the block.  A thread gathers the 8 frames of its feature with conflict-free
        // 4-byte shared loads from the swizzled raw tile (one 128-byte row segment per warp load:
        // this is the transpose), centres / scales / splits them and stores 16-byte K-major chunks
        // (512 contiguous bytes per warp store).
        const int cw = warp - V2_FIRST_CONV_WARP;
        constexpr int U = V2_UNITS_PER_WARP;
        const int n_units = 8 * nfb;
        int u_kq[U], u_fl[U];
        float u_sc[U], u_nsh[U];
        bool u_on[U], u_a[U];              // unit exists / belongs to operand 0 (the unlagged frames)
        uint32_t u_src[U], u_dst[U];       // byte offsets of the unit inside a raw stage / an operand stage
        const int chunk = lane >> 2, within = (lane & 3) * 4;
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const int u = cw + V2_CONV_WARPS * k;
            u_on[k] = u < n_units;
            const int uu = u_on[k] ? u : 0;
            const int nf = nfb > 0 ? nfb : 1;
            const int op = uu / (4 * nf);
            u_a[k] = op == 0;
            u_kq[k] = (uu / nf) & 3;
            u_fl[k] = 32 * (uu % nf) + lane;
            u_sc[k] = ctl->sc[u_fl[k]];
            u_nsh[k] = ctl->nsh[u_fl[k]];
            // raw tile of block fb: [frame][128 B], 16-byte chunks XOR-swizzled by frame & 7.  Frame
            // 8 kq + i sits at  base + i * 128  with the chunk bits flipped by i: one XOR with an
            // immediate per load, the row offset folds into the instruction
            u_src[k] = (uint32_t)(op * raw_op_bytes + (u_fl[k] >> 5) * (UM_KT * 128)
                                  + u_kq[k] * 1024 + (chunk << 4) + within);
            u_dst[k] = (uint32_t)((op == 0 ? V2_T_A : V2_T_B) * V2_TILE + u_fl[k] * 16 + u_kq[k] * UM_LBO);
        }
        const uint32_t raw_s = smem_u32(raw_ring), op_s = smem_u32(op_ring);
        float sAh[U], sAl[U];                            // column sums (operand 0 units) as float pairs
#pragma unroll
        for (int k = 0; k < U; ++k) sAh[k] = sAl[k] = 0.f;
        uint32_t hmax = 0;                               // largest |h| seen, as two fp16 bit patterns
        int stage = 0, ostage = 0;
        uint32_t phase = 0, ophase = 0;
        const bool dbg_on = P.dbg != nullptr && group == 0 && tid == V2_CONV_TID0 && cta_rank == 0;
        long long d_raw = 0, d_empty = 0, d_comp = 0, d_sync = 0;
        // -------- MN-major mode: a thread owns 8 consecutive features (32 bytes of a raw row) of two rows of
        // every tile.  Thread map (ct = 0..255): cq = ct & 3 (which 8 of a 32-feature block), fbl / fbh = low /
        // high bit of the block, swp, rp: rows 2 rp + (fbl ^ swp) and + 16.  Within a quarter warp (fixed
        // fbh, rp, swp) the lanes with fbl = 1 take the other row of the pair: the 16-byte loads of the
        // swizzled raw rows and the 16-byte stores into the swizzled windows are both bank-conflict free.
        float mn_sh[8], mn_sl[8];                        // column sums of this thread's 8 features (float pairs)
        float mn_sc[8];
        // (64 features per CTA -- D = 64, two feature blocks: no fbh, rp = 0..15 covers the 32 rows, ONE row per thread)
        const int mct = tid - V2_CONV_TID0;
        const bool m_narrow = nfb == 2;
        const int m_cq = mct & 3, m_fbl = (mct >> 2) & 1, m_swp = (mct >> 3) & 1;
        const int m_fbh = m_narrow ? 0 : (mct >> 4) & 1, m_rp = m_narrow ? (mct >> 4) : (mct >> 5);
        const int m_nu = m_narrow ? 1 : 2;                         // rows per thread and tile
        const int m_f0 = 32 * (2 * m_fbh + m_fbl) + 8 * m_cq;     // first feature (inside the CTA)
        if constexpr (mn) {
            const int row_lo = 2 * m_rp + (m_fbl ^ m_swp);
            const uint32_t swz = (uint32_t)(row_lo & 7);           // (row_lo + 16) & 7 is the same
            const uint32_t src_a = (uint32_t)((2 * m_fbh + m_fbl) * (UM_KT * 128) + row_lo * 128) + (((2u * m_cq) ^ swz) << 4);
            const uint32_t src_b = (uint32_t)((2 * m_fbh + m_fbl) * (UM_KT * 128) + row_lo * 128) + (((2u * m_cq + 1u) ^ swz) << 4);
            const uint32_t win_bytes = (uint32_t)v3_win_bytes(S), mirror = (uint32_t)v3_mirror(S);
            const uint32_t dst0 = (uint32_t)(m_fbh * v3_block_bytes(S) + row_lo * 128) + ((((uint32_t)(4 * m_fbl + m_cq)) ^ swz) << 4);
            uint64_t sc2[4], nsh2[4], ps[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                sc2[i] = f2_pack(ctl->sc[m_f0 + 2 * i], ctl->sc[m_f0 + 2 * i + 1]);
                nsh2[i] = f2_pack(ctl->nsh[m_f0 + 2 * i], ctl->nsh[m_f0 + 2 * i + 1]);
                ps[i] = f2_pack(0.f, 0.f);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) { mn_sh[i] = mn_sl[i] = 0.f; mn_sc[i] = ctl->sc[m_f0 + i]; }
            auto fold_sums = [&]() {                     // float32 partial sums (<= 8 values each) -> float pairs
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float p0, p1;
                    f2_unpack(ps[i], p0, p1);
                    const float pv[2] = {p0, p1};
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float ts = pv[e], h0 = mn_sh[2 * i + e];
                        const float tt = h0 + ts, bp = tt - h0;
                        mn_sl[2 * i + e] += (h0 - (tt - bp)) + (ts - bp);
                        mn_sh[2 * i + e] = tt;
                    }
                    ps[i] = f2_pack(0.f, 0.f);
                }
            };
            for (int t = 0; t < conv_tiles; ++t) {
                long long q0 = dbg_on ? clock64() : 0;
                mbar_wait(&ctl->raw_full[stage], phase);
                long long q1 = dbg_on ? clock64() : 0;
                mbar_wait(&ctl->empty[ostage], ophase ^ 1);          // the UMMAs that read this ring tile are done
                long long q2 = dbg_on ? clock64() : 0;
                const int valid = ctl->valid_rows[stage];
                const uint32_t rawst = raw_s + (uint32_t)stage * raw_stage_bytes;
                const uint32_t st = op_s + (uint32_t)ostage * (UM_KT * 128) + dst0;
                const bool count = t < my_tiles;                     // the extra tile belongs to the next group's sums
                float4 va[2], vb[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u < m_nu) {
                        va[u] = lds_v4f(rawst + src_a + (uint32_t)u * (16 * 128));
                        vb[u] = lds_v4f(rawst + src_b + (uint32_t)u * (16 * 128));
                    }
                }
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (u >= m_nu) break;                             // CTA uniform
                    const bool live = row_lo + 16 * u < valid;        // rows behind the sequence: zeros
                    uint64_t as2[4];
                    as2[0] = f2_fma(f2_pack(va[u].x, va[u].y), sc2[0], nsh2[0]);
                    as2[1] = f2_fma(f2_pack(va[u].z, va[u].w), sc2[1], nsh2[1]);
                    as2[2] = f2_fma(f2_pack(vb[u].x, vb[u].y), sc2[2], nsh2[2]);
                    as2[3] = f2_fma(f2_pack(vb[u].z, vb[u].w), sc2[3], nsh2[3]);
                    uint32_t hw[4], lw[4], qw[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (!live) as2[i] = f2_pack(0.f, 0.f);
                        float a0, a1;
                        f2_unpack(as2[i], a0, a1);
                        hw[i] = pack_f16(a0, a1);
                        float l0, l1;
                        asm("{\n\t.reg .f16 lo, hi, m1;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
                            "fma.rn.f32.f16 %0, lo, m1, %3;\n\tfma.rn.f32.f16 %1, hi, m1, %4;\n\t}"
                            : "=f"(l0), "=f"(l1) : "r"(hw[i]), "f"(a0), "f"(a1));
                        lw[i] = pack_f16(l0, l1);
                        asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(qw[i]) : "r"(hw[i]), "r"(0x38003800u));
                        if (count) ps[i] = f2_add(ps[i], as2[i]);
                    }
                    {
                        uint32_t m01, m23;
                        asm("max.u16x2 %0, %1, %2;" : "=r"(m01) : "r"(hw[0] & 0x7FFF7FFFu), "r"(hw[1] & 0x7FFF7FFFu));
                        asm("max.u16x2 %0, %1, %2;" : "=r"(m23) : "r"(hw[2] & 0x7FFF7FFFu), "r"(hw[3] & 0x7FFF7FFFu));
                        asm("max.u16x2 %0, %1, %2;" : "=r"(m01) : "r"(m01), "r"(m23));
                        asm("max.u16x2 %0, %1, %2;" : "=r"(hmax) : "r"(hmax), "r"(m01));
                    }
                    const uint32_t d = st + (uint32_t)u * (16 * 128);
                    sts_v4r(d, hw[0], hw[1], hw[2], hw[3]);
                    sts_v4r(d + win_bytes, lw[0], lw[1], lw[2], lw[3]);
                    sts_v4r(d + 2 * win_bytes, qw[0], qw[1], qw[2], qw[3]);
                    if (ostage == 0) {                               // ring tile 0 again behind the last tile
                        sts_v4r(d + mirror, hw[0], hw[1], hw[2], hw[3]);
                        sts_v4r(d + mirror + win_bytes, lw[0], lw[1], lw[2], lw[3]);
                        sts_v4r(d + mirror + 2 * win_bytes, qw[0], qw[1], qw[2], qw[3]);
                    }
                }
                if ((t & 3) == 3) fold_sums();
                long long q3 = dbg_on ? clock64() : 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("bar.sync 1, %0;" :: "n"(32 * V2_CONV_WARPS) : "memory");
                if (tid == V2_CONV_TID0) {
                    mbar_arrive_local(&ctl->raw_empty[stage]);
                    if constexpr (CG == 2) mbar_arrive_cluster(&ctl->conv[ostage], 0);
                    else mbar_arrive_local(&ctl->conv[ostage]);
                }
                if (dbg_on) { long long q4 = clock64(); d_raw += q1 - q0; d_empty += q2 - q1; d_comp += q3 - q2; d_sync += q4 - q3; }
                if (++stage == n_raw) { stage = 0; phase ^= 1; }
                if (++ostage == S) { ostage = 0; ophase ^= 1; }
            }
            fold_sums();
        } else
        for (int t = 0; t < my_tiles; ++t) {
            long long q0 = dbg_on ? clock64() : 0;
            mbar_wait(&ctl->raw_full[stage], phase);
            long long q1 = dbg_on ? clock64() : 0;
            mbar_wait(&ctl->empty[ostage], ophase ^ 1);
            long long q2 = dbg_on ? clock64() : 0;
            const int valid = ctl->valid_rows[stage];
            const uint32_t rawst = raw_s + (uint32_t)stage * raw_stage_bytes;
            const uint32_t st = op_s + (uint32_t)ostage * V2_STAGE_BYTES;
            // FAST: all four feature blocks exist (D = 256 per pair / 128 per single CTA): every warp has
            // four units, the first two of operand 0 and the last two of operand 1 -- known at compile
            // time, so the unit loop carries no branch and the loads of two units are in flight
            // together (with the runtime `continue`s each unit waited for its own 8 loads: the first
            // FFMA2 after the loads was the converters' top stall, r2i profile)
            // (two feature blocks, D = 64: two units per warp, one of each operand -- the same trick)
            auto convert_tile = [&](auto full_tag, auto fast_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
                constexpr int NU = decltype(fast_tag)::value;       // 0: runtime unit table; 4 or 2: static
                constexpr bool FAST = NU != 0;
#pragma unroll
                for (int k0 = 0; k0 < (FAST ? NU : U); k0 += 2) {
                    // two units at a time: 16 loads in flight, then the arithmetic (all four at once
                    // spilled and was 40 % slower, r2k)
                    float v[2][8];
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int k = k0 + kk;
                        if (!FAST && !u_on[k]) continue;
                        const uint32_t b = rawst + u_src[k];
                        v[kk][0] = lds_f32<0 * 128>(b);
                        v[kk][1] = lds_f32<1 * 128>(b ^ (1u << 4));
                        v[kk][2] = lds_f32<2 * 128>(b ^ (2u << 4));
                        v[kk][3] = lds_f32<3 * 128>(b ^ (3u << 4));
                        v[kk][4] = lds_f32<4 * 128>(b ^ (4u << 4));
                        v[kk][5] = lds_f32<5 * 128>(b ^ (5u << 4));
                        v[kk][6] = lds_f32<6 * 128>(b ^ (6u << 4));
                        v[kk][7] = lds_f32<7 * 128>(b ^ (7u << 4));
                    }
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const int k = k0 + kk;
                        if (!FAST && !u_on[k]) continue;
                        const uint64_t sc2 = f2_pack(u_sc[k], u_sc[k]);
                        const uint64_t nsh2 = f2_pack(u_nsh[k], u_nsh[k]);
                        uint32_t hw[4], lw[4];
                        uint64_t as2[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            // (v - shift) * scale in one rounding: the scale is a power of two, so this is
                            // exactly scale * fl32(v - shift), the x' of the float64 edge kernel
                            as2[i] = f2_fma(f2_pack(v[kk][2 * i], v[kk][2 * i + 1]), sc2, nsh2);
                            if (!FULL) {
                                float a0, a1;
                                f2_unpack(as2[i], a0, a1);
                                const int r = 8 * u_kq[k] + 2 * i;
                                as2[i] = f2_pack(r < valid ? a0 : 0.f, r + 1 < valid ? a1 : 0.f);
                            }
                            float a0, a1;
                            f2_unpack(as2[i], a0, a1);
                            hw[i] = pack_f16(a0, a1);               // h: round to nearest fp16 (Inf beyond 65504)
                            // l = x' - h, exact in fp32: one mixed-precision FMA per value (h * -1 + x'),
                            // full rate -- the unpack + packed subtract it replaces ran at half rate
                            // (profiles/r2e_probe6_instruction_rates.log)
                            float l0, l1;
                            asm("{\n\t.reg .f16 lo, hi, m1;\n\tmov.b32 {lo, hi}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
                                "fma.rn.f32.f16 %0, lo, m1, %3;\n\tfma.rn.f32.f16 %1, hi, m1, %4;\n\t}"
                                : "=f"(l0), "=f"(l1) : "r"(hw[i]), "f"(a0), "f"(a1));
                            lw[i] = pack_f16(l0, l1);
                        }
                        // range check: largest |h| of this thread (Inf and NaN sort above every finite value)
                        {
                            uint32_t m01, m23;
                            asm("max.u16x2 %0, %1, %2;" : "=r"(m01) : "r"(hw[0] & 0x7FFF7FFFu), "r"(hw[1] & 0x7FFF7FFFu));
                            asm("max.u16x2 %0, %1, %2;" : "=r"(m23) : "r"(hw[2] & 0x7FFF7FFFu), "r"(hw[3] & 0x7FFF7FFFu));
                            asm("max.u16x2 %0, %1, %2;" : "=r"(m01) : "r"(m01), "r"(m23));
                            asm("max.u16x2 %0, %1, %2;" : "=r"(hmax) : "r"(hmax), "r"(m01));
                        }
                        const uint32_t dst = st + u_dst[k];
                        if (FAST ? (k < NU / 2) : u_a[k]) {
                            sts_v4<0>(dst, hw[0], hw[1], hw[2], hw[3]);                            // a
                            sts_v4<(V2_T_AL - V2_T_A) * V2_TILE>(dst, lw[0], lw[1], lw[2], lw[3]);  // al
                            // h / 2 (one exact fp16 multiply per pair): the second factor of the G product
                            uint32_t h0, h1, h2, h3;
                            asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(h0) : "r"(hw[0]), "r"(0x38003800u));
                            asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(h1) : "r"(hw[1]), "r"(0x38003800u));
                            asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(h2) : "r"(hw[2]), "r"(0x38003800u));
                            asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(h3) : "r"(hw[3]), "r"(0x38003800u));
                            sts_v4<(V2_T_AH - V2_T_A) * V2_TILE>(dst, h0, h1, h2, h3);              // ah
                            // column sum (operand 0 only): TwoSum of the 8-frame sum into the float pair
                            float s0, s1;
                            f2_unpack(f2_add(f2_add(as2[0], as2[1]), f2_add(as2[2], as2[3])), s0, s1);
                            const float ts = s0 + s1;
                            const float tt = sAh[k] + ts, bp = tt - sAh[k];
                            sAl[k] += (sAh[k] - (tt - bp)) + (ts - bp);
                            sAh[k] = tt;
                        } else {
                            sts_v4<0>(dst, hw[0], hw[1], hw[2], hw[3]);                            // b
                            sts_v4<(V2_T_BL - V2_T_B) * V2_TILE>(dst, lw[0], lw[1], lw[2], lw[3]);  // bl
                        }
                    }
                }
            };
            if (P.dbg_mode & 1) { /* timing experiment: no conversion traffic */ }
            else if (nfb == 4) {
                if (valid == UM_KT) convert_tile(std::true_type(), std::integral_constant<int, 4>());
                else convert_tile(std::false_type(), std::integral_constant<int, 4>());
            } else if (nfb == 2) {
                if (valid == UM_KT) convert_tile(std::true_type(), std::integral_constant<int, 2>());
                else convert_tile(std::false_type(), std::integral_constant<int, 2>());
            } else {
                if (valid == UM_KT) convert_tile(std::true_type(), std::integral_constant<int, 0>());
                else convert_tile(std::false_type(), std::integral_constant<int, 0>());
            }
            long long q3 = dbg_on ? clock64() : 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("bar.sync 1, %0;" :: "n"(32 * V2_CONV_WARPS) : "memory");
            if (tid == V2_CONV_TID0) {
                mbar_arrive_local(&ctl->raw_empty[stage]);
                if constexpr (CG == 2) mbar_arrive_cluster(&ctl->conv[ostage], 0);
                else mbar_arrive_local(&ctl->conv[ostage]);
            }
            if (dbg_on) { long long q4 = clock64(); d_raw += q1 - q0; d_empty += q2 - q1; d_comp += q3 - q2; d_sync += q4 - q3; }
            if (++stage == n_raw) { stage = 0; phase ^= 1; }
            if (++ostage == S) { ostage = 0; ophase ^= 1; }
        }
        if (dbg_on) { P.dbg[4] = d_raw; P.dbg[5] = d_empty; P.dbg[6] = d_comp; P.dbg[7] = d_sync; }
        // Range check.  The scale puts the largest magnitude of the sample into [1, 2).  A value 2^6
        // times larger still fits fp16, but its 22-bit split is then coarse next to everything else
        // (the bulk of the feature drowns in the rounding of the outlier's square), and beyond 65504
        // h is Inf.  Either way the whole call is redone by the float64 engine (guarded launches
        // behind this kernel; non-finite input takes the same detour and comes out as in float64).
        if (max(hmax & 0xFFFFu, hmax >> 16) >= V2_RANGE_LIMIT) atomicOr(P.overflow, 1);
        // column sums: the 4 threads (one per K chunk) that share a feature combine through shared
        // memory in a fixed order (the raw ring is idle: every TMA load has landed and been
        // converted); the doubles appear only here, after this CTA's last tile
        if constexpr (mn) {
            // 16 threads (swp, rp) share a feature (32 with 64 features per CTA): fixed order through shared memory
            double *s_sum = reinterpret_cast<double *>(raw_ring);        // [16 or 32][UM_F]
            const int contrib = m_swp + 2 * m_rp;
#pragma unroll
            for (int i = 0; i < 8; ++i)
                s_sum[contrib * UM_F + m_f0 + i] = ((double)mn_sh[i] + (double)mn_sl[i]) / (double)mn_sc[i];
            asm volatile("bar.sync 1, %0;" :: "n"(32 * V2_CONV_WARPS) : "memory");
            const int f = tid - V2_CONV_TID0;
            if (f < UM_F) {
                double tot = 0.0;
                if (f < d_local)
                    for (int c = 0; c < (m_narrow ? 32 : 16); ++c) tot += s_sum[c * UM_F + f];
                P.sums[(size_t)group * UM_D + UM_F * cta_rank + f] = tot;
            }
        } else {
            double *s_sum = reinterpret_cast<double *>(raw_ring);        // [4][UM_F]
#pragma unroll
            for (int k = 0; k < U; ++k)
                if (u_on[k] && u_a[k])
                    s_sum[u_kq[k] * UM_F + u_fl[k]] = ((double)sAh[k] + (double)sAl[k]) / (double)u_sc[k];
            asm volatile("bar.sync 1, %0;" :: "n"(32 * V2_CONV_WARPS) : "memory");
            const int f = tid - V2_CONV_TID0;
            if (f < UM_F)
                P.sums[(size_t)group * UM_D + UM_F * cta_rank + f] = f < d_local
                    ? ((s_sum[f] + s_sum[UM_F + f]) + s_sum[2 * UM_F + f]) + s_sum[3 * UM_F + f] : 0.0;
        }
    } else if (warp >= V2_FIRST_DRAIN_WARP) {
        // ================================ drain warps (two per TMEM lane quarter, every CTA) ======
        const int quarter = warp & 3;                            // TMEM lanes this warp may touch
        const int half = (warp - V2_FIRST_DRAIN_WARP) >> 2;      // its half of a region's columns
        const int row = quarter * 32 + lane;                     // accumulator row = feature in this CTA
        const size_t cta_base = (size_t)cta_global * V2_REGIONS * RW * UM_F;
        // columns [c_lo, c_hi) of the region are this warp's; none when the UMMAs are narrower
        const int c_lo = half * CW, c_hi = min(n_cols, (half + 1) * CW);
        int next_end[V2_REGIONS], slabs_done[V2_REGIONS];
        uint32_t full_ph[V2_REGIONS];
        const bool dbg_on = P.dbg != nullptr && group == 0 && cta_rank == 0 && warp == V2_FIRST_DRAIN_WARP && lane == 0;
        long long d_wait = 0, d_ld = 0, d_red = 0, d_fold = 0, n_events = 0;
#pragma unroll
        for (int q = 0; q < V2_REGIONS; ++q) {
            int e = ST - 1 - region_off(q);
            if (e > my_tiles - 1) e = my_tiles - 1;
            next_end[q] = my_tiles > 0 ? e : 0x7fffffff;
            slabs_done[q] = 0;
            full_ph[q] = 0;
        }
        while (true) {
            const int q = next_end[1] < next_end[0] ? 1 : 0;
            const int e = next_end[q];
            if (e == 0x7fffffff) break;
            const bool last = e == my_tiles - 1;
            const long long c0 = dbg_on ? clock64() : 0;
            mbar_wait_idle(&ctl->acc_full[q], full_ph[q]);
            full_ph[q] ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;");
            const long long c1 = dbg_on ? clock64() : 0;
            float *l1 = P.lvl1 + cta_base + (size_t)q * RW * UM_F + row;
            long long t_ld = 0;
            bool released = false;
            if (!(P.dbg_mode & 2)) {
#pragma unroll 1
                for (int c = c_lo; c < c_hi; c += 64) {
                    uint32_t v0[32], v1[32];
                    const long long a0 = dbg_on ? clock64() : 0;
                    UM_TMEM_LD32(v0, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(RW * q + c));
                    if (c + 32 < c_hi)
                        UM_TMEM_LD32(v1, tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(RW * q + c + 32));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c + 64 >= c_hi) {
                        // the warp's share of the region is in registers: hand TMEM back before the
                        // reductions are issued
                        asm volatile("tcgen05.fence::before_thread_sync;");
                        __syncwarp();
                        if (lane == 0 && !last) {
                            if constexpr (CG == 2) mbar_arrive_cluster(&ctl->acc_empty[q], 0);
                            else mbar_arrive_local(&ctl->acc_empty[q]);
                        }
                        released = true;
                    }
                    if (dbg_on) t_ld += clock64() - a0;
                    float *dst = l1 + (size_t)c * UM_F;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;"
                                     :: "l"(dst + (size_t)j * UM_F), "f"(__uint_as_float(v0[j])) : "memory");
                    if (c + 32 < c_hi) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;"
                                         :: "l"(dst + (size_t)(32 + j) * UM_F), "f"(__uint_as_float(v1[j])) : "memory");
                    }
                }
            }
            if (!released) {
                asm volatile("tcgen05.fence::before_thread_sync;");
                __syncwarp();
                if (lane == 0 && !last) {
                    if constexpr (CG == 2) mbar_arrive_cluster(&ctl->acc_empty[q], 0);
                    else mbar_arrive_local(&ctl->acc_empty[q]);
                }
            }
            const long long c2 = dbg_on ? clock64() : 0;
            if (!last && !(P.dbg_mode & 2) && c_lo + 16 * (slabs_done[q] % NG) < c_hi) {
                // one 16-column group of this warp's columns: float32 level -> float-float pair,
                // error-free (TwoSum), then clear the level.  This warp is the only one that ever
                // touches these addresses; its own reductions above are ordered before these loads
                // (same thread, same address, gpu scope).  The finalize path adds what is left in the level.
                const int c = c_lo + 16 * (slabs_done[q] % NG);
                float *hi = P.hi + cta_base + (size_t)q * RW * UM_F + row;
                float *lo = P.lo + cta_base + (size_t)q * RW * UM_F + row;
                float s[16], H[16], L[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];"
                                 : "=f"(s[j]) : "l"(l1 + (size_t)(c + j) * UM_F) : "memory");
                    H[j] = __ldcg(hi + (size_t)(c + j) * UM_F);
                    L[j] = __ldcg(lo + (size_t)(c + j) * UM_F);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const float t = H[j] + s[j], bp = t - H[j];
                    const float err = (H[j] - (t - bp)) + (s[j] - bp);
                    __stcg(hi + (size_t)(c + j) * UM_F, t);
                    __stcg(lo + (size_t)(c + j) * UM_F, L[j] + err);
                    __stcg(l1 + (size_t)(c + j) * UM_F, 0.f);
                }
            }
            if (dbg_on) {
                const long long c3 = clock64();
                d_wait += c1 - c0; d_ld += t_ld; d_red += (c2 - c1) - t_ld; d_fold += c3 - c2; ++n_events;
            }
            ++slabs_done[q];
            if (last) next_end[q] = 0x7fffffff;
            else next_end[q] = e + ST < my_tiles - 1 ? e + ST : my_tiles - 1;
        }
        if (dbg_on) { P.dbg[8] = d_wait; P.dbg[9] = d_ld; P.dbg[10] = d_red; P.dbg[11] = d_fold; P.dbg[12] = n_events; }
    }

    // teardown: nobody may still be using the peer's barriers / TMEM
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();
    if (warp == 0) {
        if constexpr (CG == 2)
            asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS));
        else
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(TMEM_COLS));
    }
}

// Sum the CTA-level accumulators (float32 level + float-float pair) over the groups, in group order:
// R[cta_rank][region][col][row] in float64 (1 MB at most).  Coalesced: consecutive threads read
// consecutive rows.  Skipped when the fp16 range check tripped (the v1 rescue owns the memory then).
template <int CG>
__global__ void __launch_bounds__(256)
tica_umma_v2_reduce_kernel(const float *__restrict__ lvl1, const float *__restrict__ hi,
                           const float *__restrict__ lo, int n_groups,
                           const int *__restrict__ rescued, double *__restrict__ R)
{
    constexpr int RW = 128 * CG;
    constexpr size_t PER_CTA = (size_t)V2_REGIONS * RW * UM_F;
    if (*rescued != 0) return;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= CG * PER_CTA) return;
    const size_t rank = idx / PER_CTA, within = idx % PER_CTA;
    double s = 0.0;
    for (int g = 0; g < n_groups; ++g) {
        const size_t a = ((size_t)g * CG + rank) * PER_CTA + within;
        s += ((double)hi[a] + (double)lo[a]) + (double)lvl1[a];
    }
    R[idx] = s;
}

// symmetrise C_00 = G + G^T, add the float64 edge terms, undo scale and shift, add into `acc`
template <int CG>
__global__ void __launch_bounds__(256)
tica_umma_v2_finalize_kernel(const double *__restrict__ R, const double *__restrict__ sums,
                             int n_groups, const double *__restrict__ E, const double *__restrict__ es,
                             const float *__restrict__ shift, const float *__restrict__ scale,
                             const int *__restrict__ rescued, double n_pairs_total, double n_obs,
                             double n_seq, int Dr, int all_rows, double *__restrict__ acc)
{
    constexpr int D = UM_D;
    constexpr int RW = 128 * CG;
    if (*rescued != 0) return;
    const size_t DD = (size_t)D * D;
    const size_t RR = (size_t)Dr * Dr;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int)RR) return;
    const int i = idx / Dr, j = idx % Dr;
    // element (row r, feature column c) of matrix m (0 = C_tau, 1 = G): CTA r / 128 of the group,
    // region m, column c (UMMA column n = CTA n / 128, feature n % 128 = global feature n), row r % 128
    auto at = [&](int m, int r, int c) -> double {
        return R[(((size_t)(r / UM_F) * V2_REGIONS + (size_t)m) * RW + (size_t)c) * UM_F + (size_t)(r % UM_F)];
    };
    const double inv = 1.0 / ((double)scale[i] * (double)scale[j]);
    double ctau = at(0, i, j) * inv;
    double c00 = (at(1, i, j) + at(1, j, i)) * inv;        // C_00 = G + G^T
    const size_t pidx = (size_t)i * D + j;
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
    double esi[3] = {0.0, 0.0, 0.0}, esj[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < UM_EDGE_SLOTS; ++c) {
        const double *Ec = E + (size_t)c * 4 * DD;
        const double *ec = es + (size_t)c * 3 * D;
        e0 += Ec[pidx];
        e1 += Ec[DD + pidx];
        e2 += Ec[2 * DD + pidx];
        e3 += Ec[3 * DD + pidx];
        esi[0] += ec[i];
        esi[1] += ec[D + i];
        esi[2] += ec[2 * D + i];
        esj[0] += ec[j];
        esj[1] += ec[D + j];
        esj[2] += ec[2 * D + j];
    }
    double sum_i = 0.0, sum_j = 0.0;
    for (int p = 0; p < n_groups; ++p) {
        sum_i += sums[(size_t)p * D + i];
        sum_j += sums[(size_t)p * D + j];
    }
    ctau += e0;
    double ctt, S0i, S0j, Sti, Stj, tail_j;
    if (all_rows) {
        // MN-major mode: G and the column sums run over EVERY row of a sequence; the tail rows (E[3],
        // es[2]) leave C_00 / S_0, the head rows (E[2], es[2] - es[1]) leave C_tautau / S_tau
        ctt = c00 - e2;
        c00 -= e3;
        S0i = sum_i - esi[2];
        S0j = sum_j - esj[2];
        Sti = sum_i - (esi[2] - esi[1]);
        Stj = sum_j - (esj[2] - esj[1]);
        tail_j = esj[2];
    } else {
        c00 += e1;
        ctt = c00 - e2 + e3;
        S0i = sum_i + esi[0];
        S0j = sum_j + esj[0];
        Sti = sum_i + esi[1];
        Stj = sum_j + esj[1];
        tail_j = esj[2];
    }
    const double si = (double)shift[i], sj = (double)shift[j];
    const double Np = n_pairs_total;
    acc[idx] += ctau + S0i * sj + si * Stj + Np * si * sj;
    acc[RR + idx] += c00 + S0i * sj + si * S0j + Np * si * sj;
    acc[2 * RR + idx] += ctt + Sti * sj + si * Stj + Np * si * sj;
    if (i == 0) {
        const double S0 = S0j, St = Stj;
        const double Sall = S0 + tail_j;
        acc[3 * RR + j] += S0 + Np * sj;
        acc[3 * RR + Dr + j] += St + Np * sj;
        acc[3 * RR + 2 * Dr + j] += Sall + n_obs * sj;
        if (j == 0) {
            acc[3 * RR + 3 * Dr] += n_obs;
            acc[3 * RR + 3 * Dr + 1] += n_seq;
        }
    }
}

}  // namespace msmb
