// tica_umma.cu -- K1 on Blackwell tensor cores: TMA-fed tcgen05.mma (cta_group::2, kind::f16 or
// kind::tf32) with TMEM accumulators.  sm_100a only.
//
// What it computes (per call, for D = 32k <= 256 float32 features):
//   C_tau' = sum_t x'_t x'_{t+lag}^T       C_00' = sum_t x'_t x'_t^T        S_0' = sum_t x'_t
// over the pair indices t of every sequence, where x' = fl32(x - shift) is the frame
// re-centred by a provisional per-feature mean (covariances are shift invariant;
// the raw moments of tica.py:417-422 are reconstructed exactly in float64 by
// tica_umma_finalize).  The 2*lag head/tail frames per sequence that distinguish C_00 from
// C_tautau and S_0 from S_tau are handled in float64 by tica_umma_edges.
//
// Design:
//  * A CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns a range of 32-frame tiles.
//    CTA r holds features [128r, 128r+128): as the M-half of A (unlagged frames)
//    and as the N-half of B (lagged frames) of one M=256 x N=256 UMMA.
//  * TMEM per CTA: 128 lanes x 512 columns fp32 = C_tau rows (cols 0..255) and
//    C_00 rows (cols 256..511) of this CTA's 128 features: all of TMEM.
//  * Operands are K-major, no swizzle: 16-byte chunks = 8 (2-byte kinds) or 4 (tf32) consecutive
//    frames of one feature, core matrix = 8 features x 16 bytes (LBO = 2048 B between K chunks,
//    SBO = 128 B between 8-feature groups).
//  * TMA: one 3-D tensor map per sequence and operand, dims (32 feat, rows, D/32
//    blocks), box (32, KT, 4), SWIZZLE_128B -> smem [block][frame][128 B]: full
//    128-byte rows, so every 32-byte sector fetched is used.  The lagged operand has its own
//    map whose base is shifted by lag rows, so any lag works and no operand needs an unaligned
//    descriptor; rows past the last pair index are zero-filled by TMA.
//  * Engines (precision of the operand split; the products always accumulate in fp32 TMEM over a
//    bounded slab of frames, then drain into float64 partials):
//      3xF16  (default) x'*2^e = h + l in fp16 (per-feature power-of-two scale, 11-bit parts),
//             products hh' + hl' + lh' (~2^-22); a value outside fp16's range is detected in the
//             drain (Inf/NaN accumulators) and the call redoes itself as 6xBF16 on the stream;
//      6xBF16 x' = h + m + l in bf16, six products (~2^-24, full fp32 range); 3xBF16 (~2^-16);
//      3xTF32 / TF32: kind::tf32, K = 8.
//  * Warp roles (640 threads): w0 TMA producer, w1 MMA issuer (leader CTA; warp-uniform loop,
//    one elected lane issues), w2 TMEM allocator, w4-19 converters: a thread owns one feature and
//    8 frames of a tile -- gathers them with conflict-free 4-byte shared loads (the transpose),
//    centres/scales/splits them, stores K-major 16-byte chunks; the same 16 warps drain TMEM
//    between slabs: tcgen05.ld -> red.global.add.f32 into the pair's float32 second-level partials
//    (L2 resident), which each warp folds into its float64 partials every 16 slabs.
//  * Bit-reproducible: every address of the partials has one writer, column sums and edge terms
//    are stored per pair / per slot and added in a fixed order by tica_umma_finalize.
//  * No FP64 instruction runs while the tensor pipe is busy: on B200 it stalls for hundreds of
//    cycles (column sums are carried as float pairs and folded during the drain).
#include "common.cuh"
#include <cuda.h>
#include <vector>
#include <unordered_map>
#include <mutex>

namespace msmb {

constexpr int UM_D = 256;                       // features handled by this kernel
constexpr int UM_F = 128;                       // features per CTA
constexpr int UM_KT = 32;                       // frames per tile
constexpr int UM_RB = UM_KT / 4;                // 4-frame row-blocks per tile
constexpr int UM_STAGES = 2;                    // raw ring and operand ring depth
constexpr int UM_TILE_BYTES = UM_KT * UM_F * 4; // 16 KB
constexpr int UM_RAW_BYTES = 2 * UM_TILE_BYTES;     // raw A, raw B (TMA destinations)
constexpr int UM_STAGE_BYTES = 4 * UM_TILE_BYTES;   // A_hi, A_lo, B_hi, B_lo (UMMA operands)
constexpr int UM_BF_TILE_BYTES = UM_KT * UM_F * 2;  // bf16 component tile, 8 KB (6 of them fit a stage)
constexpr int UM_LBO = UM_F * 16;               // 2048: next row-block
constexpr int UM_SBO = 128;                     // next 8-feature group
constexpr int UM_CONV_WARPS = 16;               // converter warps per CTA (4 per scheduler: the
                                                // conversion is latency bound with fewer)
constexpr int UM_THREADS = 32 * (4 + UM_CONV_WARPS);   // w0 TMA, w1 MMA, w2 TMEM alloc, w3 idle, w4.. converters
constexpr int UM_FLUSH_WARPS = UM_CONV_WARPS;   // the converters also drain TMEM (4 per lane quarter)
constexpr int UM_EDGE_SLOTS = 8;                // blocks (= output slots) of the float64 edge kernel per row band
constexpr int UM_SLAB_TILES_DEFAULT = 32;       // 1024 frames of fp32 TMEM accumulation per flush

struct UmmaParams {
    const CUtensorMap *mapsA;     // [n_seq] unlagged
    const CUtensorMap *mapsB;     // [n_seq] base shifted by lag rows
    const int *tile_prefix;       // [n_seq + 1] tiles before sequence s
    const int *seq_pairs;         // [n_seq] pair indices of sequence s (n_s - lag)
    int n_seq;
    int n_tiles;
    int n_pairs;                  // clusters launched
    int slab_tiles;
    int passes;                   // tf32: 3 = hi/lo split, 1 = plain; bf16: 3 = h/m products, 6 = h/m/l
    int collector;                // reuse the A operand through the collector buffer
    int flush_red;                // 2: red.f32 into float32 second-level partials, folded into the float64 ones
                                  // every `fold_every` slabs; 1: red.f64 straight into the float64 partials;
                                  // 0: load + add + store
    float *partials32;            // [n_pairs][2][col][row] float32 second level (flush_red == 2)
    int fold_every;
    uint32_t h_add, h_mask;       // fp16 engine: integer rounding of the h component (significand width)
    int dbg_mode;                 // MSMB200_UMMA_DBGMODE (timing experiments, wrong results): 1 = converters
                                  // skip their loads/stores, 2 = flush skips its drain, 4 = one MMA per K step
    const float *shift;           // [D]
    const float *scale;           // [D] power-of-two per-feature scale (f16 engine), else unused
    int *overflow;                // f16 engine: set to 1 when a scaled value leaves fp16's range
    const int *run_if;            // if non-NULL the kernel only runs when *run_if != 0 (rescue launch)
    double *partials;             // [n_pairs][2][col][row]  (C_tau', C_00'), column-major so a
                                  // warp (32 rows) touches 256 contiguous bytes per column
    double *sums;                 // [n_pairs][D]  S_0' of every pair (plain stores, summed in order by finalize)
    long long *dbg;               // optional cycle counters of pair 0 (MSMB200_UMMA_DEBUG=1), else NULL
};

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// arrive on the same-named barrier of CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t *bar, uint32_t cta)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
    // same form as CUTLASS ClusterBarrier::arrive(cta_id): default .release at CTA scope --
    // the smem writes it publishes were already made visible by fence.proxy.async + bar.sync,
    // and .release.cluster would cost a MEMBAR.ALL.GPU + ERRBAR (~500 cycles) per arrival
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// frames are read once: mark them evict-first so the float64 partials stay L2 resident
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c0, int c1, int c2, uint64_t policy)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1, {%3, %4, %5}], [%2], %6;"
        :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "l"(policy)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;"
                 ::: "memory");
}
// K-major, no swizzle, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((UM_LBO >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((UM_SBO >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// (tf32 x tf32 | bf16 x bf16) -> f32, K-major A and B, M = 256 (pair), N = 256
constexpr int UM_KIND_TF32 = 0, UM_KIND_BF16 = 1, UM_KIND_F16 = 2;
template <int KIND>
__device__ __forceinline__ uint32_t umma_idesc()
{
    constexpr uint32_t fmt = KIND == UM_KIND_TF32 ? 2u : KIND == UM_KIND_BF16 ? 1u : 0u;
    uint32_t d = 0;
    d |= 1u << 4;                     // D format F32
    d |= fmt << 7;                    // A format: kind::tf32 TF32 = 2; kind::f16 BF16 = 1, F16 = 0
    d |= fmt << 10;                   // B format
    d |= (uint32_t)(256 >> 3) << 17;  // N
    d |= (uint32_t)(256 >> 4) << 24;  // M
    return d;
}
// The MMA warp runs its loop warp-uniformly and predicates the tcgen05 instructions on ONE elected
// lane (`mma_leader`, in scope at every use): under a divergent `if (lane == 0)` the compiler cannot
// prove the descriptors uniform and wraps every UTCHMMA in an ELECT / R2UR.BROADCAST waterfall loop,
// which serialised issue and execution (178 instead of 128 cycles per MMA, profiles/r1_k1_issue.txt).
__device__ __forceinline__ uint32_t elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(pred));
    return pred;
}
#define UMMA_KIND_PAIR(KIND, QUAL, tmem_d, da, db, idesc, accumulate) asm volatile( \
    "{\n\t.reg .pred p, q;\n\tsetp.ne.b32 p, %4, 0;\n\tsetp.ne.b32 q, %6, 0;\n\t" \
    "@q tcgen05.mma.cta_group::2.kind::" KIND QUAL " [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" \
    :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0u), "r"(mma_leader) : "memory")
#define UMMA_TF32_PAIR(QUAL, tmem_d, da, db, idesc, accumulate) \
    UMMA_KIND_PAIR("tf32", QUAL, tmem_d, da, db, idesc, accumulate)
#define umma_bf16_pair(tmem_d, da, db, idesc, accumulate) UMMA_KIND_PAIR("f16", "", tmem_d, da, db, idesc, accumulate)
#define umma_tf32_pair(tmem_d, da, db, idesc, accumulate) UMMA_TF32_PAIR("", tmem_d, da, db, idesc, accumulate)
// arrive (when all prior MMAs of this thread have completed) on `bar` in BOTH CTAs
#define umma_commit_pair(bar) asm volatile( \
    "{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t" \
    "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" \
    :: "r"(smem_u32(bar)), "h"((uint16_t)3), "r"(mma_leader) : "memory")
// round-to-nearest (ties away) to tf32's 10-bit mantissa with two full-rate integer ops
// (cvt.rna.tf32.f32 goes through the slow conversion pipe: 16k conversions per tile)
__device__ __forceinline__ float tf32_rn(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// same for bfloat16's 8-bit significand; the value stays in a float (low 16 bits zero)
__device__ __forceinline__ float bf16_rn(float x)
{
    return __uint_as_float((__float_as_uint(x) + 0x8000u) & 0xFFFF0000u);
}
// two bf16-valued floats -> one 32-bit word (a in the low half = the earlier frame)
__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    return (__float_as_uint(a) >> 16) | (__float_as_uint(b) & 0xFFFF0000u);
}

// two floats -> packed fp16 pair, round to nearest even (a in the low half = the earlier frame)
__device__ __forceinline__ uint32_t pack_f16(float a, float b)
{
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

#define UM_TMEM_LD32(v, taddr) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
    : "r"(taddr) : "memory")

struct UmmaSmem {
    uint64_t raw_full[UM_STAGES];  // TMA landed (tx bytes), local
    uint64_t raw_empty[UM_STAGES]; // converters done reading the raw stage (1 arrival), local
    uint64_t conv[UM_STAGES];      // tile converted; the LEADER's copy is used, 2 arrivals (1 per CTA)
    uint64_t empty[UM_STAGES];     // MMAs done reading the operand stage (multicast commit), local
    uint64_t acc_full;             // slab finished (multicast commit), local
    uint64_t acc_empty;            // accumulators drained; the LEADER's copy is used (24 arrivals)
    uint32_t tmem_base;
    int valid_rows[UM_STAGES];
};


// Pull this warp's share of the pair's float64 partials into L2 ahead of the drain: the
// read-modify-write below is otherwise a chain of dependent DRAM round trips (the partials are
// evicted between flushes by the frame stream), measured at ~20k cycles per flush.
__device__ __forceinline__ void flush_prefetch(const UmmaParams &P, int pair, uint32_t cta_rank,
                                               int quarter, int part, int nparts, int lane)
{
    const int row = UM_F * cta_rank + quarter * 32;          // 32 rows = 256 B per column
    const double *pc = P.partials + (size_t)pair * 2 * UM_D * UM_D + row;
    for (int c0 = 32 * part; c0 < 512; c0 += 32 * nparts) {
        // lane -> column c0 + lane: one 256-byte row segment = two 128-byte lines
        const double *line = pc + (size_t)(c0 + lane) * UM_D;
        asm volatile("prefetch.global.L2 [%0];" :: "l"(line));
        asm volatile("prefetch.global.L2 [%0];" :: "l"(line + 16));
    }
}

// Drain this warp's share of the TMEM accumulators into the pair's float64 partials.
// quarter = warp % 4 (the TMEM lanes a warp may touch); the 16 column chunks of 32 are
// dealt round-robin to the `nparts` warps that share a quarter.
__device__ __forceinline__ void flush_share(const UmmaParams &P, uint32_t tmem, int pair,
                                            uint32_t cta_rank, int quarter, int part, int nparts,
                                            int lane, uint32_t &absmax, bool fold, float *stage)
{
    const int row = UM_F * cta_rank + quarter * 32 + lane;
    double *pc = P.partials + (size_t)pair * 2 * UM_D * UM_D + row;
    float *pc32 = P.partials32 + (size_t)pair * 2 * UM_D * UM_D + row;
    // flush_red 3 / 4: the float32 level is laid out per (CTA, lane quarter, 32-column chunk) as one
    // contiguous 4 KB block, [col][row] (3: TMA bulk reduce from shared memory) or
    // [col / 4][row][col % 4] (4: 16-byte vector reductions)
    float *blk32 = P.partials32 + (size_t)pair * 2 * UM_D * UM_D
                   + (size_t)((cta_rank * 4 + quarter) * 16) * 1024;
#pragma unroll 1
    for (int c0 = 32 * part; c0 < 512; c0 += 32 * nparts) {
        uint32_t v[32];
        UM_TMEM_LD32(v, tmem + ((uint32_t)(quarter * 32) << 16) + c0);
        if (P.flush_red == 3) {
            float *g = blk32 + (size_t)(c0 >> 5) * 1024;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // the bulk reduction that last read this warp's 2 KB staging buffer must be done with it
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    absmax = max(absmax, v[16 * h + j] & 0x7FFFFFFFu);
                    stage[j * 32 + lane] = __uint_as_float(v[16 * h + j]);
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                                 :: "l"(g + h * 512), "r"(smem_u32(stage)), "r"(2048) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            continue;
        }
        if (P.flush_red == 4) {
            float *g = blk32 + (size_t)(c0 >> 5) * 1024 + lane * 4;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                absmax = max(max(absmax, v[4 * q] & 0x7FFFFFFFu), v[4 * q + 1] & 0x7FFFFFFFu);
                absmax = max(max(absmax, v[4 * q + 2] & 0x7FFFFFFFu), v[4 * q + 3] & 0x7FFFFFFFu);
                asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                             :: "l"(g + q * 128), "f"(__uint_as_float(v[4 * q])),
                                "f"(__uint_as_float(v[4 * q + 1])), "f"(__uint_as_float(v[4 * q + 2])),
                                "f"(__uint_as_float(v[4 * q + 3])) : "memory");
            }
            continue;
        }
        // element (row, col) of matrix m lives at ((m*256 + col) * 256 + row); c0 runs over
        // [C_tau cols 0..255 | C_00 cols 0..255] = m*256 + col directly
        double *dst = pc + (size_t)c0 * UM_D;
        if (P.flush_red == 2) {
            // float32 second level: half the atomic sectors of the float64 reduction, and the
            // 38 MB of all pairs stay L2 resident between drains (the 77 MB of float64 partials did
            // not: 47 GB of DRAM writes per 50M frames).  <= fold_every slabs are summed here with
            // round-to-nearest float adds (unbiased, ~1e-7) before they move on to float64.
            float *d32 = pc32 + (size_t)c0 * UM_D;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                absmax = max(absmax, v[j] & 0x7FFFFFFFu);      // Inf / NaN sort above every finite value
                asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;"
                             :: "l"(d32 + (size_t)j * UM_D), "f"(__uint_as_float(v[j])) : "memory");
            }
            continue;
        }
        if (P.flush_red) {
            // fire-and-forget reductions at L2: no read round trip, half the SM <-> L2 bytes
            // (this warp is the only writer of these addresses, so the sums stay deterministic)
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                absmax = max(absmax, v[j] & 0x7FFFFFFFu);
                asm volatile("red.relaxed.gpu.global.add.f64 [%0], %1;"
                             :: "l"(dst + (size_t)j * UM_D), "d"((double)__uint_as_float(v[j])) : "memory");
            }
            continue;
        }
        double cur[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = __ldcg(dst + (size_t)(q * 8 + j) * UM_D);
            if (q == 0) asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                absmax = max(absmax, v[q * 8 + j] & 0x7FFFFFFFu);
                __stcg(dst + (size_t)(q * 8 + j) * UM_D,
                       cur[j] + (double)__uint_as_float(v[q * 8 + j]));
            }
        }
    }
    if (P.flush_red >= 3 && fold) {
        if (P.flush_red == 3) {
            // every bulk reduction of this warp has been performed before its sums are read back
            if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncwarp();
            asm volatile("fence.proxy.async;" ::: "memory");
        }
#pragma unroll 1
        for (int c0 = 32 * part; c0 < 512; c0 += 32 * nparts) {
            float *g = blk32 + (size_t)(c0 >> 5) * 1024;
            double *dst = pc + (size_t)c0 * UM_D;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float s[16];
                double cur[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int jj = q * 16 + j;
                    float *a = P.flush_red == 3 ? g + jj * 32 + lane
                                                : g + ((jj >> 2) * 32 + lane) * 4 + (jj & 3);
                    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];" : "=f"(s[j]) : "l"(a) : "memory");
                    cur[j] = __ldcg(dst + (size_t)jj * UM_D);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int jj = q * 16 + j;
                    float *a = P.flush_red == 3 ? g + jj * 32 + lane
                                                : g + ((jj >> 2) * 32 + lane) * 4 + (jj & 3);
                    __stcg(dst + (size_t)jj * UM_D, cur[j] + (double)s[j]);
                    __stcg(a, 0.f);
                }
            }
        }
    }
    if (P.flush_red == 2 && fold) {
        // move this warp's share of the float32 level into the float64 partials and clear it.  The
        // warp is the only one that ever touches these addresses; its own reductions above are
        // ordered before these loads (same thread, same address, gpu scope).
#pragma unroll 1
        for (int c0 = 32 * part; c0 < 512; c0 += 32 * nparts) {
            float *d32 = pc32 + (size_t)c0 * UM_D;
            double *dst = pc + (size_t)c0 * UM_D;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                float s[16];
                double cur[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    asm volatile("ld.relaxed.gpu.global.f32 %0, [%1];"
                                 : "=f"(s[j]) : "l"(d32 + (size_t)(q * 16 + j) * UM_D) : "memory");
                    cur[j] = __ldcg(dst + (size_t)(q * 16 + j) * UM_D);
                }
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    __stcg(dst + (size_t)(q * 16 + j) * UM_D, cur[j] + (double)s[j]);
                    __stcg(d32 + (size_t)(q * 16 + j) * UM_D, 0.f);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
template <int KIND>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(UM_THREADS, 1)
tica_umma_kernel(const UmmaParams P)
{
    constexpr bool BF16 = KIND != UM_KIND_TF32;      // 2-byte operands, K = 16 (kind::f16)
    constexpr bool F16 = KIND == UM_KIND_F16;
    // rescue launch (the f16 engine's fp16 range check tripped): uniform over the grid
    if (P.run_if != nullptr && *reinterpret_cast<const volatile int *>(P.run_if) == 0) return;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1 KB-aligned operand ring, control block behind it
    unsigned char *ring = reinterpret_cast<unsigned char *>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *raw_ring = ring;                                   // [UM_STAGES][A raw | B raw]
    unsigned char *op_ring = ring + UM_STAGES * UM_RAW_BYTES;         // [UM_STAGES][A_hi|A_lo|B_hi|B_lo]
    UmmaSmem *ctl = reinterpret_cast<UmmaSmem *>(op_ring + UM_STAGES * UM_STAGE_BYTES);
    // 2 KB per converter warp: staging of the TMA bulk-reduce drain (flush_red == 3)
    float *drain_stage = reinterpret_cast<float *>(op_ring + UM_STAGES * UM_STAGE_BYTES + 1024);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // provably warp-uniform (see elect_one)
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    const int pair = blockIdx.x >> 1;

    // this pair's tile range (identical in every role of both CTAs)
    const long long t_begin = (long long)P.n_tiles * pair / P.n_pairs;
    const long long t_end = (long long)P.n_tiles * (pair + 1) / P.n_pairs;
    const int my_tiles = (int)(t_end - t_begin);
    // Slab boundaries are staggered across pairs (pair p's first slab is shorter) so the
    // float64 flushes of the 74 pairs do not hit L2/HBM in the same microsecond.
    const int slab_off = (int)(((long long)pair * P.slab_tiles) / P.n_pairs);   // in [0, slab_tiles)
    const int first_len = P.slab_tiles - slab_off;                              // tiles in slab 0
    const int n_slabs = (my_tiles <= first_len) ? 1
                        : 1 + (my_tiles - first_len + P.slab_tiles - 1) / P.slab_tiles;

    if (tid == 0) {
        for (int s = 0; s < UM_STAGES; ++s) {
            mbar_init(&ctl->raw_full[s], 1);
            mbar_init(&ctl->raw_empty[s], 1);
            mbar_init(&ctl->conv[s], 2);
            mbar_init(&ctl->empty[s], 1);
        }
        mbar_init(&ctl->acc_full, 1);
        mbar_init(&ctl->acc_empty, 2 * UM_FLUSH_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(&ctl->tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = ctl->tmem_base;

    if (warp == 0) {
        // ================================ TMA producer (one lane, both CTAs) ============
        if (lane == 0 && my_tiles > 0) {
            int s = 0;
            {   // locate the sequence of the first tile
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
            for (long long t = t_begin; t < t_end; ++t) {
                while (t >= P.tile_prefix[s + 1]) ++s;
                const int row0 = (int)(t - P.tile_prefix[s]) * UM_KT;
                int valid = P.seq_pairs[s] - row0;
                if (valid > UM_KT) valid = UM_KT;
                mbar_wait(&ctl->raw_empty[stage], phase ^ 1);
                ctl->valid_rows[stage] = valid;
                mbar_expect_tx(&ctl->raw_full[stage], 2 * UM_TILE_BYTES);
                unsigned char *st = raw_ring + stage * UM_RAW_BYTES;
                tma_load_3d(st, &P.mapsA[s], &ctl->raw_full[stage], 0, row0, 4 * (int)cta_rank, policy);
                tma_load_3d(st + UM_TILE_BYTES, &P.mapsB[s], &ctl->raw_full[stage], 0, row0,
                            4 * (int)cta_rank, policy);
                if (++stage == UM_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================ MMA issuer (leader CTA, one lane) =============
        if (cta_rank == 0 && my_tiles > 0) {
            const uint32_t mma_leader = elect_one();
            const uint32_t idesc = umma_idesc<KIND>();
            const uint32_t ring_addr = smem_u32(op_ring);
            const bool dbg_on = P.dbg != nullptr && pair == 0;
            long long d_wait_conv = 0, d_wait_acc = 0;
            const long long d_start = clock64();
            int stage = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int t = 0; t < my_tiles; ++t) {
                const bool slab_first = t == 0 || ((t + slab_off) % P.slab_tiles) == 0;
                const bool slab_last = ((t + slab_off + 1) % P.slab_tiles) == 0 || t + 1 == my_tiles;
                long long c0 = dbg_on ? clock64() : 0;
                mbar_wait(&ctl->conv[stage], phase);
                long long c1 = dbg_on ? clock64() : 0;
                if (slab_first && t > 0) {
                    mbar_wait(&ctl->acc_empty, acc_phase);
                    acc_phase ^= 1;
                }
                if (dbg_on) { long long c2 = clock64(); d_wait_conv += c1 - c0; d_wait_acc += c2 - c1; }
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t stage_addr = ring_addr + stage * UM_STAGE_BYTES;
                if constexpr (!BF16) {
                    const uint32_t a_hi = stage_addr;
                    const uint32_t a_lo = a_hi + UM_TILE_BYTES;
                    const uint32_t b_hi = a_hi + 2 * UM_TILE_BYTES;
                    const uint32_t b_lo = a_hi + 3 * UM_TILE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < UM_KT / 8; ++ks) {
                        const uint32_t off = ks * 2 * UM_LBO;
                        const uint32_t acc = (slab_first && ks == 0) ? 0u : 1u;
                        const uint64_t dAh = umma_desc(a_hi + off), dAl = umma_desc(a_lo + off);
                        const uint64_t dBh = umma_desc(b_hi + off), dBl = umma_desc(b_lo + off);
                        if (P.passes == 3 && P.collector) {
                            // the four MMAs that share A = hi, then the two that share A = lo, keep A
                            // in the collector buffer: 2 instead of 6 A-tile fetches from shared memory
                            UMMA_TF32_PAIR(".collector::a::fill", tmem, dAh, dBh, idesc, acc);           // C_tau
                            UMMA_TF32_PAIR(".collector::a::use", tmem, dAh, dBl, idesc, 1u);
                            UMMA_TF32_PAIR(".collector::a::use", tmem + 256, dAh, dAh, idesc, acc);      // C_00
                            UMMA_TF32_PAIR(".collector::a::lastuse", tmem + 256, dAh, dAl, idesc, 1u);
                            UMMA_TF32_PAIR(".collector::a::fill", tmem, dAl, dBh, idesc, 1u);
                            UMMA_TF32_PAIR(".collector::a::lastuse", tmem + 256, dAl, dAh, idesc, 1u);
                        } else {
                            umma_tf32_pair(tmem, dAh, dBh, idesc, acc);            // C_tau
                            if (P.passes == 3) {
                                umma_tf32_pair(tmem, dAh, dBl, idesc, 1u);
                                umma_tf32_pair(tmem, dAl, dBh, idesc, 1u);
                            }
                            umma_tf32_pair(tmem + 256, dAh, dAh, idesc, acc);      // C_00
                            if (P.passes == 3) {
                                umma_tf32_pair(tmem + 256, dAh, dAl, idesc, 1u);
                                umma_tf32_pair(tmem + 256, dAl, dAh, idesc, 1u);
                            }
                        }
                    }
                } else {
                    // bf16 components h, m, l of A (unlagged) and B (lagged), 8 KB each:
                    // x ~ h + m (+ l);  3 products: hh' + hm' + mh' (~2^-16 relative, unbiased);
                    // 6 products add mm' + hl' + lh' (~2^-24).  K = 16 frames per instruction.
                    // fp16 engine: the same three products on components h, l of the SCALED
                    // value (11-bit significands: ~2^-22), l sits in the m slot.
                    const uint32_t A0 = stage_addr, B0 = stage_addr + 3 * UM_BF_TILE_BYTES;
#pragma unroll
                    for (int ks = 0; ks < UM_KT / 16; ++ks) {
                        const uint32_t off = ks * 2 * UM_LBO;
                        const uint32_t acc = (slab_first && ks == 0) ? 0u : 1u;
                        const uint64_t dAh = umma_desc(A0 + off), dAm = umma_desc(A0 + UM_BF_TILE_BYTES + off),
                                       dAl = umma_desc(A0 + 2 * UM_BF_TILE_BYTES + off);
                        const uint64_t dBh = umma_desc(B0 + off), dBm = umma_desc(B0 + UM_BF_TILE_BYTES + off),
                                       dBl = umma_desc(B0 + 2 * UM_BF_TILE_BYTES + off);
                        if (P.dbg_mode & 4) {
                            umma_bf16_pair(tmem, dAh, dBh, idesc, acc);
                            continue;
                        }
                        if (P.passes == 3 && P.collector) {
                            // A = h feeds four MMAs, A = m (l for fp16) two: 2 instead of 6 A-tile
                            // fetches from shared memory per K step
                            UMMA_KIND_PAIR("f16", ".collector::a::fill", tmem, dAh, dBh, idesc, acc);
                            UMMA_KIND_PAIR("f16", ".collector::a::use", tmem, dAh, dBm, idesc, 1u);
                            UMMA_KIND_PAIR("f16", ".collector::a::use", tmem + 256, dAh, dAh, idesc, acc);
                            UMMA_KIND_PAIR("f16", ".collector::a::lastuse", tmem + 256, dAh, dAm, idesc, 1u);
                            UMMA_KIND_PAIR("f16", ".collector::a::fill", tmem, dAm, dBh, idesc, 1u);
                            UMMA_KIND_PAIR("f16", ".collector::a::lastuse", tmem + 256, dAm, dAh, idesc, 1u);
                            continue;
                        }
                        umma_bf16_pair(tmem, dAh, dBh, idesc, acc);                // C_tau
                        umma_bf16_pair(tmem, dAh, dBm, idesc, 1u);
                        umma_bf16_pair(tmem, dAm, dBh, idesc, 1u);
                        umma_bf16_pair(tmem + 256, dAh, dAh, idesc, acc);          // C_00
                        umma_bf16_pair(tmem + 256, dAh, dAm, idesc, 1u);
                        umma_bf16_pair(tmem + 256, dAm, dAh, idesc, 1u);
                        if (P.passes == 6) {
                            umma_bf16_pair(tmem, dAm, dBm, idesc, 1u);
                            umma_bf16_pair(tmem, dAh, dBl, idesc, 1u);
                            umma_bf16_pair(tmem, dAl, dBh, idesc, 1u);
                            umma_bf16_pair(tmem + 256, dAm, dAm, idesc, 1u);
                            umma_bf16_pair(tmem + 256, dAh, dAl, idesc, 1u);
                            umma_bf16_pair(tmem + 256, dAl, dAh, idesc, 1u);
                        }
                    }
                }
                umma_commit_pair(&ctl->empty[stage]);      // stage may be rewritten (both CTAs)
                if (slab_last) umma_commit_pair(&ctl->acc_full);
                if (++stage == UM_STAGES) { stage = 0; phase ^= 1; }
            }
            if (dbg_on && lane == 0) {
                P.dbg[0] = clock64() - d_start;      // MMA thread: total issue-loop cycles
                P.dbg[1] = d_wait_conv;              //   of which waiting for converted operands
                P.dbg[2] = d_wait_acc;               //   of which waiting for the TMEM flush
                P.dbg[3] = my_tiles;
            }
        }
    } else if (warp >= 4 && warp < 4 + UM_CONV_WARPS) {
        // ================================ converters (256 threads, both CTAs) ==========
        // Thread map: warp cw & 3 -> 32-feature block of this CTA, lane -> feature inside
        // the block; cw >> 2 -> which quarter (8 frames) of a tile.  A thread gathers the
        // frames of ITS feature 4 (tf32) or 8 (2-byte kinds) at a time with 4-byte
        // shared loads (a warp reads one 128-byte swizzled row segment per load: all 32
        // banks, conflict free) and stores one 16-byte K-major chunk (a warp writes 512
        // contiguous bytes): the fp32 -> tf32 hi/lo conversion transposes for free.
        const int cw = warp - 4;
        const int fb = cw & 3;                           // feature block (32 features)
        const int kq = cw >> 2;                          // 8-frame quarter of the tile handled by this warp
        const int f_local = 32 * fb + lane;              // feature inside the CTA (0..127)
        const int chunk = lane >> 2, within = (lane & 3) * 4;
        const float sh = P.shift[UM_F * cta_rank + f_local];
        const float sc = F16 ? P.scale[UM_F * cta_rank + f_local] : 1.f;
        const float nsh = -sh * sc;                      // exact (power-of-two scale)
        const uint32_t h_add = P.h_add, h_mask = P.h_mask;
        const bool split = P.passes == 3;
        double sumA = 0.0;                               // column sum (unlagged rows) of feature 128*rank + f_local,
                                                         // in units of 1/sc; the lagged sum follows from it and the
                                                         // head/tail rows (tica_umma_edges_kernel)
        float sAh = 0.f, sAl = 0.f;                      //   its float-pair front end (see below)
        int stage = 0;
        uint32_t phase = 0;
        const bool dbg_on = P.dbg != nullptr && pair == 0 && tid == 128 && cta_rank == 0;
        long long d_raw = 0, d_empty = 0, d_comp = 0, d_sync = 0, d_flush = 0;
        // the converters have nothing to convert while the accumulators are being drained (the UMMA
        // pipe stalls, so no operand stage is released): they do the draining, 4 warps per TMEM quarter
        int next_flush = 0;
        uint32_t acc_phase = 0;
        uint32_t absmax = 0;                             // largest |accumulator| bit pattern drained so far
        auto help_flush = [&]() {
            const long long f0 = dbg_on ? clock64() : 0;
            if (!P.flush_red) flush_prefetch(P, pair, cta_rank, cw & 3, cw >> 2, 4, lane);
            mbar_wait(&ctl->acc_full, acc_phase);
            acc_phase ^= 1;
            asm volatile("tcgen05.fence::after_thread_sync;");
            sumA += (double)sAh + (double)sAl;           // the tensor pipe is idle here: FP64 is cheap
            sAh = sAl = 0.f;
            if (!(P.dbg_mode & 2))
                flush_share(P, tmem, pair, cta_rank, cw & 3, cw >> 2, 4, lane, absmax,
                            ((next_flush + 1) % P.fold_every) == 0 || next_flush + 1 == n_slabs,
                            drain_stage + cw * 512);
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncwarp();
            if (lane == 0 && next_flush + 1 < n_slabs) mbar_arrive_cluster(&ctl->acc_empty, 0);
            ++next_flush;
            if (dbg_on) d_flush += clock64() - f0;
        };
        for (int t = 0; t < my_tiles; ++t) {
            // slab j ends with tile min(first_len + j*slab_tiles, my_tiles) - 1; by the time this warp is
            // about to convert tile t >= end + 2 both operand stages are full, so it may as well drain
            while (next_flush < n_slabs) {
                int end = first_len + next_flush * P.slab_tiles;     // one past the slab's last tile
                if (end > my_tiles) end = my_tiles;
                if (end - 1 + 2 > t) break;
                help_flush();
            }
            long long q0 = dbg_on ? clock64() : 0;
            mbar_wait(&ctl->raw_full[stage], phase);
            long long q1 = dbg_on ? clock64() : 0;
            mbar_wait(&ctl->empty[stage], phase ^ 1);      // UMMA finished with this operand stage
            long long q2 = dbg_on ? clock64() : 0;
            const int valid = ctl->valid_rows[stage];
            const unsigned char *rawst = raw_ring + stage * UM_RAW_BYTES;
            unsigned char *st = op_ring + stage * UM_STAGE_BYTES;
            float tsA = 0.f;
            // full tiles (all but the last of a sequence) skip the per-value row test
            auto convert_tile = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
                for (int op = 0; op < 2; ++op) {
                    const unsigned char *raw = rawst + op * UM_TILE_BYTES + fb * (UM_KT * 128) + within;
                    if constexpr (!BF16) {
                        unsigned char *hi_buf = st + op * 2 * UM_TILE_BYTES + f_local * 16;
                        unsigned char *lo_buf = hi_buf + UM_TILE_BYTES;
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int rb = 2 * kq + k;
                            float a[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int r = 4 * rb + i;          // frame inside the tile
                                const float v = *reinterpret_cast<const float *>(
                                    raw + r * 128 + ((chunk ^ (r & 7)) << 4));
                                a[i] = (FULL || r < valid) ? v - sh : 0.f;
                            }
                            const float s4 = (a[0] + a[1]) + (a[2] + a[3]);
                            if (op == 0) tsA += s4;
                            float4 hv;
                            hv.x = tf32_rn(a[0]); hv.y = tf32_rn(a[1]); hv.z = tf32_rn(a[2]); hv.w = tf32_rn(a[3]);
                            *reinterpret_cast<float4 *>(hi_buf + rb * UM_LBO) = hv;
                            if (split) {
                                float4 l;
                                l.x = tf32_rn(a[0] - hv.x); l.y = tf32_rn(a[1] - hv.y);
                                l.z = tf32_rn(a[2] - hv.z); l.w = tf32_rn(a[3] - hv.w);
                                *reinterpret_cast<float4 *>(lo_buf + rb * UM_LBO) = l;
                            }
                        }
                    } else if constexpr (F16) {
                        // fp16 h/l split of the scaled value (x - shift) * 2^e: h = the value rounded
                        // to 11 significant bits (two integer ops; exact in fp16's normal range),
                        // l = fp16(value - h).  Below 2^-14 both roundings are absolute, <= 2^-25.
                        unsigned char *h_buf = st + op * 3 * UM_BF_TILE_BYTES + f_local * 16;
                        unsigned char *l_buf = h_buf + UM_BF_TILE_BYTES;
                        {
                            const int kb = kq;                      // block of 8 frames
                            float h[8], l[8];
                            float s8 = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = 8 * kb + i;
                                const float v = *reinterpret_cast<const float *>(
                                    raw + r * 128 + ((chunk ^ (r & 7)) << 4));
                                // (v - sh) * sc in one rounding: the scale is a power of two, so this is
                            // exactly sc * fl32(v - sh), the x' of the float64 edge kernel
                            const float as = (FULL || r < valid) ? fmaf(v, sc, nsh) : 0.f;
                            s8 += as;
                            h[i] = __uint_as_float((__float_as_uint(as) + h_add) & h_mask);
                            l[i] = as - h[i];                 // exact in fp32
                        }
                        if (op == 0) tsA += s8;
                        uint4 hw, lw;
                            hw.x = pack_f16(h[0], h[1]); hw.y = pack_f16(h[2], h[3]);
                            hw.z = pack_f16(h[4], h[5]); hw.w = pack_f16(h[6], h[7]);
                            lw.x = pack_f16(l[0], l[1]); lw.y = pack_f16(l[2], l[3]);
                            lw.z = pack_f16(l[4], l[5]); lw.w = pack_f16(l[6], l[7]);
                            *reinterpret_cast<uint4 *>(h_buf + kb * UM_LBO) = hw;
                            *reinterpret_cast<uint4 *>(l_buf + kb * UM_LBO) = lw;
                        }
                    } else {
                        // 16-byte chunk = 8 consecutive frames of this thread's feature
                        unsigned char *h_buf = st + op * 3 * UM_BF_TILE_BYTES + f_local * 16;
                        unsigned char *m_buf = h_buf + UM_BF_TILE_BYTES;
                        unsigned char *l_buf = h_buf + 2 * UM_BF_TILE_BYTES;
                        const bool six = P.passes == 6;
                        {
                            const int kb = kq;                      // block of 8 frames
                            float a[8], h[8], m[8];
                            float s8 = 0.f;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int r = 8 * kb + i;
                                const float v = *reinterpret_cast<const float *>(
                                    raw + r * 128 + ((chunk ^ (r & 7)) << 4));
                                a[i] = (FULL || r < valid) ? v - sh : 0.f;
                                s8 += a[i];
                                h[i] = bf16_rn(a[i]);
                                m[i] = bf16_rn(a[i] - h[i]);      // a - h is exact in fp32
                            }
                            if (op == 0) tsA += s8;
                            uint4 hw, mw;
                            hw.x = pack_bf16(h[0], h[1]); hw.y = pack_bf16(h[2], h[3]);
                            hw.z = pack_bf16(h[4], h[5]); hw.w = pack_bf16(h[6], h[7]);
                            mw.x = pack_bf16(m[0], m[1]); mw.y = pack_bf16(m[2], m[3]);
                            mw.z = pack_bf16(m[4], m[5]); mw.w = pack_bf16(m[6], m[7]);
                            *reinterpret_cast<uint4 *>(h_buf + kb * UM_LBO) = hw;
                            *reinterpret_cast<uint4 *>(m_buf + kb * UM_LBO) = mw;
                            if (six) {
                                float l[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) l[i] = bf16_rn((a[i] - h[i]) - m[i]);
                                uint4 lw;
                                lw.x = pack_bf16(l[0], l[1]); lw.y = pack_bf16(l[2], l[3]);
                                lw.z = pack_bf16(l[4], l[5]); lw.w = pack_bf16(l[6], l[7]);
                                *reinterpret_cast<uint4 *>(l_buf + kb * UM_LBO) = lw;
                            }
                        }
                    }
                }
            };
            if (P.dbg_mode & 1) { /* timing experiment: no conversion traffic */ }
            else if (valid == UM_KT) convert_tile(std::true_type());
            else convert_tile(std::false_type());
            // column sums: error-free float pair (TwoSum) per tile, folded into the doubles only while
            // the accumulators drain -- an FP64 instruction issued while the tensor pipe is busy
            // stalls for hundreds of cycles on B200 (stall_math on this DADD was the top stall of
            // the converter warps, profiles/r1_k1_issue.txt)
            {
                const float t = sAh + tsA, bp = t - sAh;
                sAl += (sAh - (t - bp)) + (tsA - bp);
                sAh = t;
            }
            long long q3 = dbg_on ? clock64() : 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // st.shared -> UMMA (async proxy)
            asm volatile("bar.sync 1, %0;" :: "n"(32 * UM_CONV_WARPS) : "memory");   // all converter warps of this CTA
            if (tid == 128) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];"
                             :: "r"(smem_u32(&ctl->raw_empty[stage])) : "memory");
                mbar_arrive_cluster(&ctl->conv[stage], 0);
            }
            if (dbg_on) { long long q4 = clock64(); d_raw += q1 - q0; d_empty += q2 - q1; d_comp += q3 - q2; d_sync += q4 - q3; }
            if (++stage == UM_STAGES) { stage = 0; phase ^= 1; }
        }
        while (next_flush < n_slabs) help_flush();
        if (dbg_on) { P.dbg[4] = d_raw; P.dbg[5] = d_empty; P.dbg[6] = d_comp; P.dbg[7] = d_sync;
                      P.dbg[8] = d_flush; P.dbg[9] = n_slabs; }
        // column sums without atomics (bit-reproducible runs): the 4 warps that share a feature
        // combine through shared memory in a fixed order (the raw ring is idle by now: every TMA
        // load of this CTA has landed and been converted), one plain store per pair and feature
        {
            double *s_sum = reinterpret_cast<double *>(raw_ring);        // [4][UM_F]
            sumA += (double)sAh + (double)sAl;
            s_sum[kq * UM_F + f_local] = sumA / (double)sc;
            asm volatile("bar.sync 1, %0;" :: "n"(32 * UM_CONV_WARPS) : "memory");
            if (kq == 0)
                P.sums[(size_t)pair * UM_D + UM_F * cta_rank + f_local] =
                    ((s_sum[f_local] + s_sum[UM_F + f_local]) + s_sum[2 * UM_F + f_local]) + s_sum[3 * UM_F + f_local];
        }
        // fp16 tops out at 65504: a scaled value beyond it became Inf, and Inf (or the NaN of
        // Inf * 0) is then in the accumulators of its feature -- the whole call is redone by the
        // bf16 engine (rescue launch).  Non-finite input takes the same, harmless, detour.
        if (F16 && absmax >= 0x7F800000u) atomicOr(P.overflow, 1);
    }

    // teardown: nobody may still be using the peer's barriers / TMEM
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    cluster_sync_all();
    if (warp == 2)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// ---------------------------------------------------------------------------------------
// provisional per-feature mean -> float32 shift, and (fp16 engine) a power-of-two scale 2^-e that
// brings the largest centred magnitude m of the sample into [1, 2): fp16 then has 2^15 of headroom
// above the sample and 11-bit components of everything down to 2^-14 of it.  The sample is
// UM_SAMPLE_ROWS rows spread evenly over ALL frames of the call (row j of the sample = frame
// floor(j * total / rows) of the concatenated sequences), so a drifting or bursty feature is seen over
// its whole range, not only at the start of the first sequence.  m is floored at |mean| / 256 so that
// a feature which is (nearly) constant in the sample cannot be blown up into fp16's ceiling by a
// later excursion (an excursion beyond 2^15 m still trips the range check -> bf16 rescue).
// Any shift / scale gives the same moments up to rounding: they only set where the roundings fall.
constexpr int UM_SAMPLE_ROWS = 1024;
constexpr int UM_SAMPLE_BLOCKS = 32;            // 32 sample rows per block
struct EdgeSeq {
    const float *base;
    long long n;
};
// The host resolves the sample rows to pointers (sample[j] = first feature of frame
// floor(j * total / rows) of the concatenated call); three tiny launches then take the mean and the
// largest centred magnitude over the sample with every partial result added in a fixed order.
__global__ void __launch_bounds__(UM_D)
tica_sample_sum_kernel(const float *const *__restrict__ sample, int rows, int D,
                       double *__restrict__ psum /* [UM_SAMPLE_BLOCKS][UM_D] */)
{
    const int c = threadIdx.x, b = blockIdx.x;
    double s = 0.0;
    if (c < D)
        for (int j = b * 32; j < rows && j < b * 32 + 32; ++j) s += (double)sample[j][c];
    psum[b * UM_D + c] = s;
}
__global__ void __launch_bounds__(UM_D)
tica_sample_max_kernel(const float *const *__restrict__ sample, int rows, int D,
                       const double *__restrict__ psum, float *__restrict__ pmax)
{
    const int c = threadIdx.x, b = blockIdx.x;
    double s = 0.0;
    for (int k = 0; k < UM_SAMPLE_BLOCKS; ++k) s += psum[k * UM_D + c];
    const float sh = (float)(s / (double)rows);
    float m = 0.f;
    if (c < D)
        for (int j = b * 32; j < rows && j < b * 32 + 32; ++j) m = fmaxf(m, fabsf(sample[j][c] - sh));
    pmax[b * UM_D + c] = m;
}
__global__ void __launch_bounds__(UM_D)
tica_shift_kernel(const double *__restrict__ psum, const float *__restrict__ pmax, int rows, int D,
                  float *__restrict__ shift, float *__restrict__ scale)
{
    // shift[] is padded to UM_D entries; features >= D do not exist (TMA zero-fills them)
    const int c = threadIdx.x;
    double s = 0.0;
    float m = 0.f;
    for (int k = 0; k < UM_SAMPLE_BLOCKS; ++k) {
        s += psum[k * UM_D + c];
        m = fmaxf(m, pmax[k * UM_D + c]);
    }
    const float sh = c < D ? (float)(s / (double)rows) : 0.f;
    shift[c] = sh;
    m = fmaxf(m, fabsf(sh) * (1.f / 256.f));
    int e = 0;
    if (m > 0.f && m < INFINITY) e = ilogbf(m);
    e = e < -100 ? -100 : (e > 100 ? 100 : e);
    scale[c] = ldexpf(1.f, -e);
}

// rescue path of the fp16 engine: when the range check tripped, forget what that launch wrote
__global__ void tica_umma_rescue_clear_kernel(const int *__restrict__ flag, double *__restrict__ buf,
                                              size_t n)
{
    if (*flag == 0) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x)
        buf[i] = 0.0;
}

// float64 edge terms, in the SAME centred coordinates x' = fl32(x - shift):
//   E[0] += sum over the remainder pair indices t in [4*floor(P/4), P) of x'_t x'_{t+lag}^T
//   E[1] += the same t of x'_t x'_t^T
//   E[2] += sum_{t < lag} x'_t x'_t^T            (head: in C_00, not in C_tautau)
//   E[3] += sum_{t >= n-lag} x'_t x'_t^T         (tail: in C_tautau, not in C_00)
//   es[0..2] (D each): remainder sum of x'_t; (that + tail rows - head rows) = what turns the
//   unlagged column sum of the tensor-core part into the lagged one; the tail-row sum
// grid (n_blocks, D/16): block (b, it) accumulates rows 16*it..16*it+15 of the outputs
// over sequences b, b + n_blocks, ...
__global__ void __launch_bounds__(256)
tica_umma_edges_kernel(const EdgeSeq *__restrict__ seqs, int n_seq, long long ld, int lag, int Dr,
                       const float *__restrict__ shift, double *__restrict__ E,
                       double *__restrict__ es)
{
    constexpr int D = UM_D;       // padded width of every scratch array; Dr <= D real features
    constexpr int RB = 4;         // rows staged per barrier pair (one row per pair made the kernel 0.7 ms of
                                  // pure barrier latency at 500 sequences; the sums keep their row order)
    __shared__ double sx[RB][D], sy[D];
    const int tid = threadIdx.x;
    const int i0 = blockIdx.y * 16;
    double acc[4][16];
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int u = 0; u < 16; ++u) acc[m][u] = 0.0;
    double s0 = 0.0, shead = 0.0, stail = 0.0; // column `tid` sums (only blockIdx.y == 0 publishes)

    // block x owns the sequences x, x + gridDim.x, ... and writes ITS slot of E / es with plain
    // stores (slots are pre-zeroed; finalize adds them in slot order: bit-reproducible)
    E += (size_t)blockIdx.x * 4 * D * D;
    es += (size_t)blockIdx.x * 3 * D;
    for (int s = blockIdx.x; s < n_seq; s += gridDim.x) {
        const float *X = seqs[s].base;
        const long long n = seqs[s].n;
        const long long Pn = n - lag;
        const long long R = Pn;       // every pair index goes through the tensor cores (TMA zero-fills the tail)
        // remainder pairs (none while R == Pn), one row at a time
        for (long long t = R; t < Pn; ++t) {
            __syncthreads();
            sx[0][tid] = tid < Dr ? (double)(X[t * ld + tid] - shift[tid]) : 0.0;
            sy[tid] = tid < Dr ? (double)(X[(t + lag) * ld + tid] - shift[tid]) : 0.0;
            __syncthreads();
            const double xj = sx[0][tid], yj = sy[tid];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const double xi = sx[0][i0 + u];
                acc[0][u] = fma(xi, yj, acc[0][u]);
                acc[1][u] = fma(xi, xj, acc[1][u]);
            }
            s0 += xj;
        }
        // head rows (t < lag), then tail rows (t >= n - lag): RB rows per barrier pair, folded in row order
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const long long t0 = part ? n - lag : 0;
            for (int r0 = 0; r0 < lag; r0 += RB) {
                const int nb = lag - r0 < RB ? lag - r0 : RB;
                __syncthreads();
                for (int b = 0; b < nb; ++b)
                    sx[b][tid] = tid < Dr ? (double)(X[(t0 + r0 + b) * ld + tid] - shift[tid]) : 0.0;
                __syncthreads();
                for (int b = 0; b < nb; ++b) {
                    const double xj = sx[b][tid];
                    if (part == 0) {
#pragma unroll
                        for (int u = 0; u < 16; ++u) acc[2][u] = fma(sx[b][i0 + u], xj, acc[2][u]);
                        shead += xj;
                    } else {
#pragma unroll
                        for (int u = 0; u < 16; ++u) acc[3][u] = fma(sx[b][i0 + u], xj, acc[3][u]);
                        stail += xj;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 4; ++m)
#pragma unroll
        for (int u = 0; u < 16; ++u)
            if (tid < Dr) E[(size_t)m * D * D + (size_t)(i0 + u) * D + tid] = acc[m][u];
    if (blockIdx.y == 0 && tid < Dr) {
        es[tid] = s0;
        es[D + tid] = s0 + stail - shead;
        es[2 * D + tid] = stail;
    }
}

// reduce the pair partials, add the float64 edge terms, undo the shift, add into `acc`
__global__ void __launch_bounds__(256)
tica_umma_finalize_kernel(const double *__restrict__ partials, int n_pairs,
                          const double *__restrict__ sums, const double *__restrict__ E,
                          const double *__restrict__ es, const float *__restrict__ shift,
                          const float *__restrict__ scale /* NULL: partials are unscaled */,
                          const int *__restrict__ rescued /* != 0: the bf16 rescue wrote them */,
                          int only_if_rescued /* v2 engine: this kernel only finishes a rescued call */,
                          double n_pairs_total /* sum_s (n_s - lag) */, double n_obs, double n_seq,
                          int Dr, double *__restrict__ acc)
{
    if (only_if_rescued && *rescued == 0) return;
    constexpr int D = UM_D;                              // padded scratch width
    const size_t DD = (size_t)D * D;                     // scratch matrix size
    const size_t RR = (size_t)Dr * Dr;                   // output matrix size
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;       // element of a Dr x Dr matrix
    if (idx >= (int)RR) return;
    const int i = idx / Dr, j = idx % Dr;
    double ctau = 0.0, c00 = 0.0;
    const size_t tidx = (size_t)j * D + i;               // partials are column-major, padded
    const size_t pidx = (size_t)i * D + j;               // edge terms are row-major, padded
    for (int p = 0; p < n_pairs; ++p) {
        ctau += partials[(size_t)p * 2 * DD + tidx];
        c00 += partials[(size_t)p * 2 * DD + DD + tidx];
    }
    if (scale != nullptr && *rescued == 0) {
        // fp16 engine: the tensor cores summed (s_i x'_i)(s_j x'_j); powers of two, exact to undo
        const double inv = 1.0 / ((double)scale[i] * (double)scale[j]);
        ctau *= inv;
        c00 *= inv;
    }
    // edge terms and column sums: fixed-order sums over the edge kernel's slots / the pairs
    double e0 = 0.0, e1 = 0.0, e2 = 0.0, e3 = 0.0;
    double esi[2] = {0.0, 0.0}, esj[3] = {0.0, 0.0, 0.0};
    for (int c = 0; c < UM_EDGE_SLOTS; ++c) {
        const double *Ec = E + (size_t)c * 4 * DD;
        const double *ec = es + (size_t)c * 3 * D;
        e0 += Ec[pidx];
        e1 += Ec[DD + pidx];
        e2 += Ec[2 * DD + pidx];
        e3 += Ec[3 * DD + pidx];
        esi[0] += ec[i];
        esi[1] += ec[D + i];
        esj[0] += ec[j];
        esj[1] += ec[D + j];
        esj[2] += ec[2 * D + j];
    }
    double sum_i = 0.0, sum_j = 0.0;
    for (int p = 0; p < n_pairs; ++p) {
        sum_i += sums[(size_t)p * D + i];
        sum_j += sums[(size_t)p * D + j];
    }
    ctau += e0;
    c00 += e1;
    const double ctt = c00 - e2 + e3;
    const double si = (double)shift[i], sj = (double)shift[j];
    const double S0i = sum_i + esi[0], S0j = sum_j + esj[0];
    // sum over the pair rows of x'_{t+lag} = the same sum of x'_t, minus the head rows, plus the tail rows
    const double Sti = sum_i + esi[1], Stj = sum_j + esj[1];
    const double Np = n_pairs_total;
    acc[idx] += ctau + S0i * sj + si * Stj + Np * si * sj;
    acc[RR + idx] += c00 + S0i * sj + si * S0j + Np * si * sj;
    acc[2 * RR + idx] += ctt + Sti * sj + si * Stj + Np * si * sj;
    if (i == 0) {
        // vectors and counters (one thread per column j)
        const double S0 = S0j, St = Stj;
        const double Sall = S0 + esj[2];              // all rows = pair rows + last `lag` rows
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
#include "tica_umma_v2.cuh"
namespace msmb {

// float64 engine (tica_simt.cu): the rescue of the second-generation fp16 engine
size_t tica_simt_items(const void *const *seq_ptrs, const int64_t *seq_rows, int n_seq, int lag,
                       void *items_out, double *n_obs_out, double *n_used_out);
size_t tica_simt_item_bytes();
int tica_simt_launch(const void *d_items, size_t n_items, int D, int64_t ld, int dtype, int lag,
                     double n_obs, double n_used, double *acc, const int *run_if, cudaStream_t st);

// ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool tica_umma_supported(int D, int64_t ld, int dtype, int lag)
{
    // any D = 32k <= 256: TMA zero-fills the feature blocks a sequence does not have, so narrower
    // inputs ride the same 256-wide tiles (at the 256-wide cost: worthwhile from D = 64 up,
    // lib.cu picks the float64 CUDA-core engine below that)
    return D >= 32 && D <= UM_D && (D % 32) == 0 && dtype == MSMB200_F32 && (ld % 4) == 0 && lag >= 1;
}

static constexpr int UM_MAX_PAIRS = 96;

// caller-provided scratch: [shift | sums | E | es | partials for up to UM_MAX_PAIRS pairs]
static size_t ws_fixed_bytes(int D)
{
    const size_t DD = (size_t)D * D;
    // shift, scale, flag | sample partials | sums | E | es | R (v2 group sums) + alignment slack
    return 4096 + (sizeof(double) + sizeof(float)) * (size_t)UM_SAMPLE_BLOCKS * UM_D
           + sizeof(double) * ((size_t)V2_MAX_GROUPS * D + (size_t)UM_EDGE_SLOTS * (4 * DD + 3 * D) + 2 * DD)
           + 4096;
}
size_t tica_umma_workspace_bytes(int D)
{
    if (D > UM_D || (D % 32) != 0) return 0;
    // fixed part | float64 partials | float32 second-level partials
    return ws_fixed_bytes(UM_D) + (sizeof(double) + sizeof(float)) * 2 * (size_t)UM_D * UM_D * UM_MAX_PAIRS;
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v ? atoi(v) : dflt;
}

// ---- host-side staging: the per-call tables travel through a small ring of PINNED buffers (one
// cudaMemcpyAsync per call, no stream synchronisation), and encoded tensor maps are cached by
// (base, rows, pitch, width): a partial_fit stream of sequences that live in one FrameStore, or a
// bench step over the same frames, encodes each map once.
struct MapKey {
    const void *base;
    long long rows, ld;
    int D, box_blocks;
    bool operator==(const MapKey &o) const
    {
        return base == o.base && rows == o.rows && ld == o.ld && D == o.D && box_blocks == o.box_blocks;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey &k) const
    {
        size_t h = reinterpret_cast<size_t>(k.base) * 0x9E3779B97F4A7C15ull;
        h ^= (size_t)k.rows * 0xC2B2AE3D27D4EB4Full + (size_t)k.ld * 0x165667B19E3779F9ull + (size_t)k.D
             + ((size_t)k.box_blocks << 20);
        return h ^ (h >> 29);
    }
};
static std::mutex g_host_mu;
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_map_cache;

struct PinnedSlot {
    void *host = nullptr;
    size_t bytes = 0;
    cudaEvent_t done = nullptr;
    bool in_flight = false;
};
constexpr int UM_PINNED_SLOTS = 8;
constexpr int UM_MAX_DEVICES = 64;
static PinnedSlot g_pinned[UM_MAX_DEVICES][UM_PINNED_SLOTS];     // events belong to a device
static int g_pinned_next[UM_MAX_DEVICES] = {0};
static bool g_pool_tuned[UM_MAX_DEVICES] = {false};              // one-time set-up, per device
static bool g_attr_set[UM_MAX_DEVICES] = {false};

static void *g_pinned_arena[UM_MAX_DEVICES] = {nullptr};         // ONE page-locked allocation per device, cut into slots

// a pinned buffer of >= bytes whose previous copy (if any) has left the host; call with g_host_mu held.
// All slots of a device are carved out of one allocation made by the first call (and remade only when a call
// needs more than a slot holds): cudaMallocHost page-locks memory, and with one allocation per slot every
// fourth of the first eight calls of a process paid 20-70 ms for it (the driver grows its pinned pool in
// steps) -- inside the timed steps of the bench, profiles/r2h_bench_pinned_alloc_outliers.json.
static int pinned_acquire(int dev, size_t bytes, PinnedSlot **out)
{
    PinnedSlot &s = g_pinned[dev][g_pinned_next[dev]];
    g_pinned_next[dev] = (g_pinned_next[dev] + 1) % UM_PINNED_SLOTS;
    if (s.bytes < bytes) {
        for (int i = 0; i < UM_PINNED_SLOTS; ++i) {
            PinnedSlot &o = g_pinned[dev][i];
            if (o.in_flight) {
                MSMB_CUDA(cudaEventSynchronize(o.done));
                o.in_flight = false;
            }
        }
        if (g_pinned_arena[dev]) MSMB_CUDA(cudaFreeHost(g_pinned_arena[dev]));
        g_pinned_arena[dev] = nullptr;
        size_t cap = 1 << 18;
        while (cap < bytes) cap <<= 1;
        MSMB_CUDA(cudaMallocHost(&g_pinned_arena[dev], cap * UM_PINNED_SLOTS));
        for (int i = 0; i < UM_PINNED_SLOTS; ++i) {
            g_pinned[dev][i].host = reinterpret_cast<unsigned char *>(g_pinned_arena[dev]) + cap * (size_t)i;
            g_pinned[dev][i].bytes = cap;
        }
    }
    if (s.in_flight) {
        MSMB_CUDA(cudaEventSynchronize(s.done));
        s.in_flight = false;
    }
    if (!s.done) MSMB_CUDA(cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    *out = &s;
    return MSMB200_OK;
}

int tica_umma_accumulate(const void *const *seq_ptrs, const int64_t *seq_rows, int n_seq_in,
                         int D, int64_t ld, int lag, int passes, double *acc, void *workspace,
                         size_t workspace_bytes, cudaStream_t st)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return MSMB200_E_CUDA;
    }
    // usable sequences (tica.py:410-412: n <= lag is skipped and not counted)
    std::vector<EdgeSeq> seqs;
    double n_obs = 0.0, n_pairs_total = 0.0;
    for (int s = 0; s < n_seq_in; ++s) {
        if (!(seq_rows[s] > lag)) continue;
        if ((reinterpret_cast<uintptr_t>(seq_ptrs[s]) & 15u) != 0) {
            set_error("tcgen05 engine needs 16-byte aligned sequences");
            return MSMB200_E_UNSUPPORTED;
        }
        EdgeSeq e;
        e.base = reinterpret_cast<const float *>(seq_ptrs[s]);
        e.n = seq_rows[s];
        seqs.push_back(e);
        n_obs += (double)e.n;
        n_pairs_total += (double)(e.n - lag);
    }
    const int n_seq = (int)seqs.size();
    if (n_seq == 0) return MSMB200_OK;

    // per-call tables, laid out once and built straight into a pinned staging buffer:
    //   [mapsA | mapsB | tile_prefix | seq_pairs | edge seqs | sample row pointers]
    auto align_up = [](size_t v, size_t a) { return (v + a - 1) / a * a; };
    const long long total_rows = (long long)n_obs;
    const int sample_rows = (int)(total_rows < UM_SAMPLE_ROWS ? total_rows : UM_SAMPLE_ROWS);
    size_t off = 0;
    const size_t o_mapsA = off; off = align_up(off + sizeof(CUtensorMap) * n_seq, 128);
    const size_t o_mapsB = off; off = align_up(off + sizeof(CUtensorMap) * n_seq, 128);
    const size_t o_prefix = off; off = align_up(off + sizeof(int) * (n_seq + 1), 128);
    const size_t o_blocks = off; off = align_up(off + sizeof(int) * n_seq, 128);
    const size_t o_eseq = off; off = align_up(off + sizeof(EdgeSeq) * n_seq, 128);
    const size_t o_sample = off; off = align_up(off + sizeof(const float *) * UM_SAMPLE_ROWS, 128);
    // item table of the float64 engine (the v2 engine's rescue runs on the stream, guarded by the flag)
    const bool f16 = passes == 23;                  // 23 = 3xF16 (see lib.cu)
    const bool bf16 = passes >= 10;                 // 13 = 3xBF16, 16 = 6xBF16: 2-byte operands, K = 16
    // second-generation fp16 engine (tica_umma_v2.cuh): single-CTA tiles for D <= 128, CTA pairs above
    const bool v2 = f16 && env_int("MSMB200_UMMA_V1", 0) == 0;
    const int v2_cg = D <= UM_F ? 1 : 2;
    // MN-major rolling-window operands (tica_umma_v2.cuh, "MN-major mode"): every CTA has its four feature
    // blocks and the lag fits the mirror tile
    // (MSMB200_UMMA_MN=0 keeps the K-major mode; MSMB200_UMMA_MN_STAGES = ring tiles, 4 or 5)
    const bool v3 = v2 && (D == UM_D || D == UM_F || (D == 64 && env_int("MSMB200_UMMA_MN64", 0) != 0)) &&
                    lag <= UM_KT && env_int("MSMB200_UMMA_MN", 1) != 0;
    int v3_stages = env_int("MSMB200_UMMA_MN_STAGES", 5);
    v3_stages = v3_stages < 4 ? 4 : (v3_stages > 5 ? 5 : v3_stages);
    const size_t n_items = v2 ? tica_simt_items(seq_ptrs, seq_rows, n_seq_in, lag, nullptr, nullptr, nullptr) : 0;
    const size_t o_items = off; off = align_up(off + tica_simt_item_bytes() * n_items, 128);
    int dev = 0;
    MSMB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= UM_MAX_DEVICES) {
        set_error("device ordinal %d out of range", dev);
        return MSMB200_E_UNSUPPORTED;
    }
    std::unique_lock<std::mutex> host_lock(g_host_mu);
    PinnedSlot *slot = nullptr;
    {
        int rc = pinned_acquire(dev, off, &slot);
        if (rc != MSMB200_OK) return rc;
    }
    unsigned char *hb = reinterpret_cast<unsigned char *>(slot->host);
    CUtensorMap *mapsA = reinterpret_cast<CUtensorMap *>(hb + o_mapsA);
    CUtensorMap *mapsB = reinterpret_cast<CUtensorMap *>(hb + o_mapsB);
    int *tile_prefix = reinterpret_cast<int *>(hb + o_prefix);
    int *seq_pairs = reinterpret_cast<int *>(hb + o_blocks);
    memcpy(hb + o_eseq, seqs.data(), sizeof(EdgeSeq) * n_seq);
    double simt_n_obs = 0.0, simt_n_used = 0.0;
    if (n_items) tica_simt_items(seq_ptrs, seq_rows, n_seq_in, lag, hb + o_items, &simt_n_obs, &simt_n_used);
    {
        // sample row j = frame floor(j * total / rows) of the concatenated call
        const float **sample = reinterpret_cast<const float **>(hb + o_sample);
        int q = 0;
        long long before = 0;
        for (int j = 0; j < sample_rows; ++j) {
            const long long g = (long long)(((__int128)j * total_rows) / sample_rows);
            while (g >= before + seqs[q].n) { before += seqs[q].n; ++q; }
            sample[j] = seqs[q].base + (size_t)(g - before) * (size_t)ld;
        }
        for (int j = sample_rows; j < UM_SAMPLE_ROWS; ++j) sample[j] = nullptr;
    }
    if (g_map_cache.size() > (1u << 16)) g_map_cache.clear();
    // 32-feature blocks per TMA box: the single-CTA v2 kernel only brings the blocks that exist (the
    // zero-filled ones would still cost shared-memory write bandwidth, the kernel's scarcest resource)
    const int box_blocks = (v2 && v2_cg == 1) ? D / 32 : 4;
    long long tiles = 0;
    for (int s = 0; s < n_seq; ++s) {
        const long long Pn = seqs[s].n - lag;           // pair indices (>= 1)
        if (Pn > 0x7fffffffLL) {
            set_error("sequence too long for the tcgen05 engine");
            return MSMB200_E_UNSUPPORTED;
        }
        seq_pairs[s] = (int)Pn;
        tile_prefix[s] = (int)tiles;
        // MN-major mode: tiles cover every row of the sequence and `lag` rows of zeros behind it
        tiles += v3 ? (seqs[s].n + lag + UM_KT - 1) / UM_KT : (Pn + UM_KT - 1) / UM_KT;
        if (tiles > 0x7fffffffLL) {
            set_error("too many tiles");
            return MSMB200_E_UNSUPPORTED;
        }
        for (int which = 0; which < (v3 ? 1 : 2); ++which) {
            // (32 features, Pn rows, D/32 blocks); rows >= Pn read as zeros
            // (MN-major mode: ONE map over all n rows; the lagged operand is the same converted rows)
            void *base = (void *)(seqs[s].base + (which ? (size_t)lag * ld : 0));
            CUtensorMap *dst = which ? &mapsB[s] : &mapsA[s];
            const long long map_rows = v3 ? seqs[s].n : Pn;
            const MapKey key{base, map_rows, (long long)ld, D, box_blocks};
            auto hit = g_map_cache.find(key);
            if (hit != g_map_cache.end()) {
                *dst = hit->second;
                continue;
            }
            cuuint64_t dims[3] = {32, (cuuint64_t)map_rows, (cuuint64_t)(D / 32)};
            cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
            cuuint32_t box[3] = {32, UM_KT, (cuuint32_t)box_blocks};
            cuuint32_t es[3] = {1, 1, 1};
            CUresult r = enc(dst, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base,
                             dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled failed (%d) for sequence %d", (int)r, s);
                return MSMB200_E_CUDA;
            }
            g_map_cache.emplace(key, *dst);
        }
    }
    tile_prefix[n_seq] = (int)tiles;

    // v1 CTA pairs
    int n_pairs = sm_count() / 2;
    n_pairs = env_int("MSMB200_UMMA_PAIRS", n_pairs);
    if (n_pairs > UM_MAX_PAIRS) n_pairs = UM_MAX_PAIRS;
    if (tiles < n_pairs) n_pairs = (int)(tiles > 0 ? tiles : 1);
    // v2 groups: CTA pairs (CG = 2) or single CTAs (CG = 1)
    int n_groups = sm_count() / v2_cg;
    n_groups = env_int("MSMB200_UMMA_GROUPS", n_groups);
    if (n_groups > V2_MAX_GROUPS) n_groups = V2_MAX_GROUPS;
    if (n_groups > UM_MAX_PAIRS * 2 / v2_cg) n_groups = UM_MAX_PAIRS * 2 / v2_cg;   // fits the partial area
    if (tiles < n_groups) n_groups = (int)(tiles > 0 ? tiles : 1);
    const size_t DD = (size_t)UM_D * UM_D;          // scratch is always 256 wide
    const size_t v1_part_bytes = (sizeof(double) + sizeof(float)) * 2 * DD * n_pairs;
    const size_t v2_per_cta = (size_t)V2_REGIONS * (128 * v2_cg) * UM_F;           // floats per array
    const size_t v2_part_bytes = v2 ? 3 * sizeof(float) * v2_per_cta * (size_t)(n_groups * v2_cg) : 0;
    const size_t part_bytes = v2 ? v2_part_bytes : v1_part_bytes;
    const size_t need = ws_fixed_bytes(UM_D) + part_bytes;
    if (!workspace || workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255u)) {
        set_error("tica_accumulate: workspace too small or misaligned (%zu < %zu); size it with "
                  "msmb200_tica_workspace_bytes", workspace_bytes, need);
        return MSMB200_E_INVALID;
    }
    if (!g_pool_tuned[dev]) {   // keep the small stream-ordered table allocations cached
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
        }
        g_pool_tuned[dev] = true;
    }

    // device copy of the tables: stream-ordered allocation
    unsigned char *scratch = nullptr;
    MSMB_CUDA(cudaMallocAsync(&scratch, off, st));
    // big zero-initialised part: caller's workspace
    //   [shift | scale | psum | pmax | flag | sums | E | es | R | partials]
    unsigned char *wsb = reinterpret_cast<unsigned char *>(workspace);
    size_t woff = 0;
    const size_t w_shift = woff; woff = align_up(woff + sizeof(float) * UM_D, 256);
    const size_t w_scale = woff; woff = align_up(woff + sizeof(float) * UM_D, 256);
    const size_t w_psum = woff; woff = align_up(woff + sizeof(double) * UM_SAMPLE_BLOCKS * UM_D, 256);
    const size_t w_pmax = woff; woff = align_up(woff + sizeof(float) * UM_SAMPLE_BLOCKS * UM_D, 256);
    const size_t w_flag = woff; woff = align_up(woff + sizeof(int), 256);       // zeroed with the rest
    const size_t w_sums = woff; woff = align_up(woff + sizeof(double) * (size_t)V2_MAX_GROUPS * UM_D, 256);
    const size_t w_E = woff; woff = align_up(woff + sizeof(double) * UM_EDGE_SLOTS * 4 * DD, 256);
    const size_t w_es = woff; woff = align_up(woff + sizeof(double) * UM_EDGE_SLOTS * 3 * UM_D, 256);
    const size_t w_R = woff; woff = align_up(woff + sizeof(double) * 2 * DD, 256);
    const size_t w_part = woff; woff += v2 ? v2_part_bytes : v1_part_bytes;
    MSMB_CUDA(cudaMemsetAsync(wsb + w_flag, 0, woff - w_flag, st));

    // ONE asynchronous copy from pinned memory; the slot is reusable once `done` has passed
    MSMB_CUDA(cudaMemcpyAsync(scratch, hb, off, cudaMemcpyHostToDevice, st));
    MSMB_CUDA(cudaEventRecord(slot->done, st));
    slot->in_flight = true;
    host_lock.unlock();

    float *d_shift = reinterpret_cast<float *>(wsb + w_shift);
    float *d_scale = reinterpret_cast<float *>(wsb + w_scale);
    int *d_flag = reinterpret_cast<int *>(wsb + w_flag);
    {
        const float *const *d_sample = reinterpret_cast<const float *const *>(scratch + o_sample);
        double *d_psum = reinterpret_cast<double *>(wsb + w_psum);
        float *d_pmax = reinterpret_cast<float *>(wsb + w_pmax);
        tica_sample_sum_kernel<<<UM_SAMPLE_BLOCKS, UM_D, 0, st>>>(d_sample, sample_rows, D, d_psum);
        MSMB_LAUNCH_CHECK();
        tica_sample_max_kernel<<<UM_SAMPLE_BLOCKS, UM_D, 0, st>>>(d_sample, sample_rows, D, d_psum, d_pmax);
        MSMB_LAUNCH_CHECK();
        tica_shift_kernel<<<1, UM_D, 0, st>>>(d_psum, d_pmax, sample_rows, D, d_shift, d_scale);
        MSMB_LAUNCH_CHECK();
    }

    UmmaParams P;
    P.mapsA = reinterpret_cast<const CUtensorMap *>(scratch + o_mapsA);
    P.mapsB = reinterpret_cast<const CUtensorMap *>(scratch + o_mapsB);
    P.tile_prefix = reinterpret_cast<const int *>(scratch + o_prefix);
    P.seq_pairs = reinterpret_cast<const int *>(scratch + o_blocks);
    P.n_seq = n_seq;
    P.n_tiles = (int)tiles;
    P.n_pairs = n_pairs;
    // bf16 MMAs cover 16 frames per accumulate step (tf32: 8), so twice the frames per slab
    // carry the same truncation bias
    // (the fp16 engine's eigenvalue error grows ~3e-6 per 1024 frames of slab, measured: it stays at 1024)
    P.slab_tiles = env_int("MSMB200_UMMA_SLAB_TILES",
                           (bf16 && !f16) ? 2 * UM_SLAB_TILES_DEFAULT : UM_SLAB_TILES_DEFAULT);
    if (P.slab_tiles < 1) P.slab_tiles = 1;
    P.passes = f16 ? 3 : bf16 ? passes - 10 : passes;
    P.collector = env_int("MSMB200_UMMA_COLLECTOR", 1);
    P.flush_red = env_int("MSMB200_UMMA_FLUSH_RED", 2);
    P.partials32 = reinterpret_cast<float *>(wsb + w_part + sizeof(double) * 2 * DD * n_pairs);
    P.fold_every = env_int("MSMB200_UMMA_FOLD_EVERY", 16);
    if (P.fold_every < 1) P.fold_every = 1;
    P.dbg_mode = env_int("MSMB200_UMMA_DBGMODE", 0);
    {
        int hb = env_int("MSMB200_UMMA_HBITS", 11);      // significant bits kept in h (11 = all of fp16)
        if (hb < 4) hb = 4;
        if (hb > 11) hb = 11;
        P.h_add = 1u << (23 - hb);
        P.h_mask = ~((1u << (24 - hb)) - 1u);
    }
    P.shift = d_shift;
    P.scale = d_scale;
    P.overflow = d_flag;
    P.run_if = nullptr;
    P.partials = reinterpret_cast<double *>(wsb + w_part);
    P.sums = reinterpret_cast<double *>(wsb + w_sums);
    P.dbg = nullptr;
    long long *d_dbg = nullptr;
    if (env_int("MSMB200_UMMA_DEBUG", 0)) {
        MSMB_CUDA(cudaMallocAsync(&d_dbg, sizeof(long long) * 16, st));
        MSMB_CUDA(cudaMemsetAsync(d_dbg, 0, sizeof(long long) * 16, st));
        P.dbg = d_dbg;
    }
    V2Params V;
    V.mapsA = P.mapsA;
    V.mapsB = P.mapsB;
    V.tile_prefix = P.tile_prefix;
    V.seq_pairs = P.seq_pairs;
    V.n_seq = n_seq;
    V.n_tiles = (int)tiles;
    V.n_groups = n_groups;
    V.slab_tiles = env_int("MSMB200_UMMA_SLAB_TILES", UM_SLAB_TILES_DEFAULT);
    if (V.slab_tiles < 1) V.slab_tiles = 1;
    V.D = D;
    V.box_blocks = box_blocks;
    V.dbg_mode = P.dbg_mode;
    V.mn = v3 ? v3_stages : 0;
    V.lag = lag;
    V.shift = d_shift;
    V.scale = d_scale;
    V.overflow = d_flag;
    V.lvl1 = reinterpret_cast<float *>(wsb + w_part);
    V.hi = V.lvl1 + v2_per_cta * (size_t)(n_groups * v2_cg);
    V.lo = V.hi + v2_per_cta * (size_t)(n_groups * v2_cg);
    V.sums = P.sums;
    V.dbg = d_dbg;

    const size_t smem = (size_t)UM_STAGES * (UM_RAW_BYTES + UM_STAGE_BYTES) + 1024 /* control block */
                        + (size_t)UM_CONV_WARPS * 2048 /* drain staging */ + 1024 /* alignment */;
    const size_t smem_v2 = (size_t)V2_RAW_STAGES * UM_RAW_BYTES + (size_t)V2_OP_STAGES * V2_STAGE_BYTES
                           + 1536 /* control block */ + 1024 /* alignment */;
    static_assert(sizeof(UmmaSmem) <= 1024, "control block");
    static_assert(sizeof(V2Smem) <= 1536, "control block");
    if (!g_attr_set[dev]) {
        MSMB_CUDA(cudaFuncSetAttribute(tica_umma_kernel<UM_KIND_TF32>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MSMB_CUDA(cudaFuncSetAttribute(tica_umma_kernel<UM_KIND_BF16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MSMB_CUDA(cudaFuncSetAttribute(tica_umma_kernel<UM_KIND_F16>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#define MSMB_V2_ATTR(CGV, MNV)                                                                      \
        MSMB_CUDA(cudaFuncSetAttribute(tica_umma_v2_kernel<CGV, MNV>,                               \
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v2))
        MSMB_V2_ATTR(1, 0); MSMB_V2_ATTR(2, 0); MSMB_V2_ATTR(1, 4); MSMB_V2_ATTR(2, 4);
        MSMB_V2_ATTR(1, 5); MSMB_V2_ATTR(2, 5);
#undef MSMB_V2_ATTR
        g_attr_set[dev] = true;
    }
    if (tiles > 0) {
        if (f16) {
            if (v2) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3((unsigned)(n_groups * v2_cg));
                cfg.blockDim = dim3(V2_THREADS);
                cfg.dynamicSmemBytes = smem_v2;
                cfg.stream = st;
                cudaLaunchAttribute attr[1];
                attr[0].id = cudaLaunchAttributeClusterDimension;
                attr[0].val.clusterDim.x = (unsigned)v2_cg;
                attr[0].val.clusterDim.y = 1;
                attr[0].val.clusterDim.z = 1;
                cfg.attrs = attr;
                cfg.numAttrs = 1;
#define MSMB_V2_LAUNCH(MNV)                                                                         \
                do {                                                                                \
                    if (v2_cg == 2) MSMB_CUDA(cudaLaunchKernelEx(&cfg, tica_umma_v2_kernel<2, MNV>, V)); \
                    else MSMB_CUDA(cudaLaunchKernelEx(&cfg, tica_umma_v2_kernel<1, MNV>, V));       \
                } while (0)
                if (V.mn == 5) MSMB_V2_LAUNCH(5);
                else if (V.mn == 4) MSMB_V2_LAUNCH(4);
                else MSMB_V2_LAUNCH(0);
#undef MSMB_V2_LAUNCH
            } else {
                tica_umma_kernel<UM_KIND_F16><<<2 * n_pairs, UM_THREADS, smem, st>>>(P);
            }
            MSMB_LAUNCH_CHECK();
            if (!v2) {
                // Range rescue, all on the stream (no host round trip): if a scaled value left fp16's
                // range the two launches below wipe the partials and redo the call with the 6xBF16
                // engine (full fp32 exponent range); otherwise both exit at once.
                tica_umma_rescue_clear_kernel<<<2, 256, 0, st>>>(
                    d_flag, reinterpret_cast<double *>(wsb + w_sums), (w_E - w_sums) / sizeof(double));
                MSMB_LAUNCH_CHECK();
                tica_umma_rescue_clear_kernel<<<4 * sm_count(), 256, 0, st>>>(
                    d_flag, P.partials, v1_part_bytes / sizeof(double));
                MSMB_LAUNCH_CHECK();
                UmmaParams Q = P;
                Q.passes = 6;
                Q.slab_tiles = env_int("MSMB200_UMMA_SLAB_TILES", 2 * UM_SLAB_TILES_DEFAULT);
                Q.run_if = d_flag;
                Q.dbg = nullptr;
                tica_umma_kernel<UM_KIND_BF16><<<2 * n_pairs, UM_THREADS, smem, st>>>(Q);
            }
        } else if (bf16) tica_umma_kernel<UM_KIND_BF16><<<2 * n_pairs, UM_THREADS, smem, st>>>(P);
        else tica_umma_kernel<UM_KIND_TF32><<<2 * n_pairs, UM_THREADS, smem, st>>>(P);
        MSMB_LAUNCH_CHECK();
    }
    {
        dim3 grid(n_seq < UM_EDGE_SLOTS ? n_seq : UM_EDGE_SLOTS, D / 16);
        tica_umma_edges_kernel<<<grid, 256, 0, st>>>(
            reinterpret_cast<const EdgeSeq *>(scratch + o_eseq), n_seq, ld, lag, D, d_shift,
            reinterpret_cast<double *>(wsb + w_E), reinterpret_cast<double *>(wsb + w_es));
        MSMB_LAUNCH_CHECK();
    }
    const unsigned fin_blocks = (unsigned)(((size_t)D * D + 255) / 256);
    if (v2) {
        double *d_R = reinterpret_cast<double *>(wsb + w_R);
        const unsigned red_blocks = (unsigned)((v2_cg * v2_per_cta + 255) / 256);
        if (v2_cg == 2) {
            tica_umma_v2_reduce_kernel<2><<<red_blocks, 256, 0, st>>>(V.lvl1, V.hi, V.lo, n_groups, d_flag, d_R);
            MSMB_LAUNCH_CHECK();
            tica_umma_v2_finalize_kernel<2><<<fin_blocks, 256, 0, st>>>(
                d_R, V.sums, n_groups, reinterpret_cast<const double *>(wsb + w_E),
                reinterpret_cast<const double *>(wsb + w_es), d_shift, d_scale, d_flag,
                n_pairs_total, n_obs, (double)n_seq, D, v3 ? 1 : 0, acc);
        } else {
            tica_umma_v2_reduce_kernel<1><<<red_blocks, 256, 0, st>>>(V.lvl1, V.hi, V.lo, n_groups, d_flag, d_R);
            MSMB_LAUNCH_CHECK();
            tica_umma_v2_finalize_kernel<1><<<fin_blocks, 256, 0, st>>>(
                d_R, V.sums, n_groups, reinterpret_cast<const double *>(wsb + w_E),
                reinterpret_cast<const double *>(wsb + w_es), d_shift, d_scale, d_flag,
                n_pairs_total, n_obs, (double)n_seq, D, v3 ? 1 : 0, acc);
        }
        MSMB_LAUNCH_CHECK();
        // Range rescue, all on the stream (no host round trip): when a converter raised the flag the
        // two kernels above left `acc` untouched and the float64 engine (exact reference arithmetic,
        // tica_simt.cu) does the whole call; otherwise its three launches exit at once.
        const int rc = tica_simt_launch(scratch + o_items, n_items, D, ld, MSMB200_F32, lag, simt_n_obs,
                                        simt_n_used, acc, d_flag, st);
        if (rc != MSMB200_OK) return rc;
    } else {
        tica_umma_finalize_kernel<<<fin_blocks, 256, 0, st>>>(
            P.partials, n_pairs, P.sums, reinterpret_cast<const double *>(wsb + w_E),
            reinterpret_cast<const double *>(wsb + w_es), d_shift, f16 ? d_scale : nullptr, d_flag,
            0, n_pairs_total, n_obs, (double)n_seq, D, acc);
        MSMB_LAUNCH_CHECK();
    }
    if (d_dbg) {
        long long h[16];
        MSMB_CUDA(cudaMemcpyAsync(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost, st));
        MSMB_CUDA(cudaStreamSynchronize(st));
        const double nt = (double)(h[3] ? h[3] : 1);
        if (v2)
            fprintf(stderr, "[umma v2 dbg group0 cg=%d] tiles=%lld mma_loop=%lld cyc (%.0f/tile) idle=%.1f%% deferred_polls=%lld "
                    "fast_tiles=%lld | conv thread: wait_raw=%.0f wait_empty=%.0f compute=%.0f fence+sync+arrive=%.0f cyc/tile | "
                    "drain warp: %lld events, wait=%.0f ld=%.0f red=%.0f fold=%.0f cyc/event\n", v2_cg, h[3], h[0],
                    (double)h[0] / nt, 100.0 * h[1] / (h[0] ? h[0] : 1), h[2], h[13],
                    (double)h[4] / nt, (double)h[5] / nt, (double)h[6] / nt, (double)h[7] / nt, h[12],
                    (double)h[8] / (h[12] ? h[12] : 1), (double)h[9] / (h[12] ? h[12] : 1),
                    (double)h[10] / (h[12] ? h[12] : 1), (double)h[11] / (h[12] ? h[12] : 1));
        else
            fprintf(stderr, "[umma dbg pair0] tiles=%lld mma_loop=%lld cyc (%.0f/tile) wait_conv=%.1f%% wait_flush=%.1f%% | "
                    "conv thread: wait_raw=%.0f wait_empty=%.0f compute=%.0f fence+sync+arrive=%.0f cyc/tile | "
                    "flush=%.0f cyc/slab (%lld slabs)\n", h[3], h[0], (double)h[0] / nt,
                    100.0 * h[1] / (h[0] ? h[0] : 1), 100.0 * h[2] / (h[0] ? h[0] : 1),
                    (double)h[4] / nt, (double)h[5] / nt, (double)h[6] / nt,
                    (double)h[7] / nt, (double)h[8] / (h[9] ? h[9] : 1), h[9]);
        MSMB_CUDA(cudaFreeAsync(d_dbg, st));
    }
    MSMB_CUDA(cudaFreeAsync(scratch, st));
    return MSMB200_OK;
}

}  // namespace msmb
