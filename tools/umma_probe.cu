// tools/umma_probe.cu -- hardware probes for the assumptions the tcgen05 K1 kernel
// rests on.  Not part of the library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_probe tools/umma_probe.cu
//   ./umma_probe mma1     # cta_group::1, MN-major no-swizzle "panel" operands, row offsets
//   ./umma_probe tma      # 3-D tensor map (4, rows, D/4) with box (4, R, 32) -> panel layout
//   ./umma_probe mma2     # cta_group::2 pair, remote mbarrier arrive, multicast commit
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

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

// smem matrix descriptor, no swizzle, version 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // version = 1 (Blackwell)
    return d;
}

// instruction descriptor: tf32 x tf32 -> f32, both operands MN-major
__host__ __device__ inline uint32_t make_idesc(int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4;            // c_format = F32
    d |= 2u << 7;            // a_format = TF32
    d |= 2u << 10;           // b_format = TF32
    d |= 1u << 15;           // a_major = MN
    d |= 1u << 16;           // b_major = MN
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

static inline float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 11) - 5); }
static inline float bval(int n, int k) { return (float)(((n * 5 + k * 2) % 13) - 6); }

// ------------------------------------------------------------------------ mma1
constexpr int P1_R = 33;          // panel pitch in rows (odd)
constexpr int P1_M = 128, P1_N = 256;

__global__ void __launch_bounds__(128)
probe_mma1(float *out, int swap_lbo_sbo, int offA, int offB, int ksteps)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sA = reinterpret_cast<float *>(smem);                       // 32 panels x R rows x 4
    float *sB = sA + (P1_M / 4) * P1_R * 4;                            // 64 panels x R rows x 4
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + (P1_N / 4) * P1_R * 4);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < (P1_M / 4) * P1_R * 4; i += 128) {
        int e = i & 3, r = (i >> 2) % P1_R, p = (i >> 2) / P1_R;
        sA[i] = ((int)(p * 4 + e) * 7 + r * 3) % 11 - 5;               // aval(m, row)
    }
    for (int i = tid; i < (P1_N / 4) * P1_R * 4; i += 128) {
        int e = i & 3, r = (i >> 2) % P1_R, p = (i >> 2) / P1_R;
        sB[i] = ((int)(p * 4 + e) * 5 + r * 2) % 13 - 6;               // bval(n, row)
    }
    if (tid == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        const uint32_t pitch = P1_R * 16;
        const uint32_t lbo = swap_lbo_sbo ? pitch : 128, sbo = swap_lbo_sbo ? 128 : pitch;
        const uint32_t idesc = make_idesc(P1_M, P1_N);
        for (int ks = 0; ks < ksteps; ++ks) {
            uint64_t da = make_desc(smem_u32(sA) + (offA + 8 * ks) * 16, lbo, sbo);
            uint64_t db = make_desc(smem_u32(sB) + (offB + 8 * ks) * 16, lbo, sbo);
            uint32_t accum = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     :: "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    // each warp reads its 32 lanes, 256 columns in chunks of 32
    for (int c0 = 0; c0 < P1_N; c0 += 32) {
        uint32_t v[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(size_t)tid * P1_N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

static int run_mma1()
{
    float *d_out;
    CK(cudaMalloc(&d_out, sizeof(float) * P1_M * P1_N));
    size_t smem = ((P1_M / 4) + (P1_N / 4)) * P1_R * 16 + 64;
    CK(cudaFuncSetAttribute(probe_mma1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> h(P1_M * P1_N);
    int ok_any = 0;
    for (int swap = 0; swap < 2; ++swap)
        for (int cfg = 0; cfg < 4; ++cfg) {
            int offA = (cfg == 2) ? 3 : 0, offB = (cfg >= 1) ? 10 : 0, ks = (cfg == 3) ? 2 : 1;
            CK(cudaMemset(d_out, 0, sizeof(float) * P1_M * P1_N));
            probe_mma1<<<1, 128, smem>>>(d_out, swap, offA, offB, ks);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("mma1 swap=%d cfg=%d: CUDA error %s\n", swap, cfg, cudaGetErrorString(e)); return 3; }
            CK(cudaMemcpy(h.data(), d_out, sizeof(float) * P1_M * P1_N, cudaMemcpyDeviceToHost));
            int bad = 0; double maxerr = 0;
            for (int m = 0; m < P1_M; ++m)
                for (int n = 0; n < P1_N; ++n) {
                    double ref = 0;
                    for (int k = 0; k < 8 * ks; ++k) ref += (double)aval(m, k + offA) * bval(n, k + offB);
                    double err = fabs(ref - h[m * P1_N + n]);
                    if (err > 1e-3) ++bad;
                    if (err > maxerr) maxerr = err;
                }
            printf("mma1 swap_lbo_sbo=%d offA=%d offB=%d ksteps=%d : mismatches=%d maxerr=%g  [D00=%g D01=%g D10=%g]\n",
                   swap, offA, offB, ks, bad, maxerr, h[0], h[1], h[P1_N]);
            if (!bad) ++ok_any;
        }
    printf("mma1 configs fully correct: %d / 8\n", ok_any);
    return 0;
}

// ------------------------------------------------------------------------ tma
typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode()
{
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeFn)fn;
}

constexpr int PT_ROWS = 43, PT_PANELS = 32;

__global__ void probe_tma(const __grid_constant__ CUtensorMap tmap, float *out, int row0, int group0)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + PT_PANELS * PT_ROWS * 16);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     :: "r"(smem_u32(bar)), "r"(PT_PANELS * PT_ROWS * 16) : "memory");
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
            :: "r"(smem_u32(tile)), "l"(&tmap), "r"(smem_u32(bar)), "r"(0), "r"(row0), "r"(group0) : "memory");
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < PT_PANELS * PT_ROWS * 4; i += blockDim.x) out[i] = tile[i];
}

static int run_tma()
{
    const int n = 100, D = 256;
    std::vector<float> h((size_t)n * D);
    for (int r = 0; r < n; ++r) for (int c = 0; c < D; ++c) h[(size_t)r * D + c] = r * 1000 + c;
    float *d_x, *d_out;
    CK(cudaMalloc(&d_x, h.size() * 4));
    CK(cudaMemcpy(d_x, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, PT_PANELS * PT_ROWS * 16));
    EncodeFn enc = get_encode();
    CUtensorMap tm;
    cuuint64_t dims[3] = {4, (cuuint64_t)n, (cuuint64_t)(D / 4)};
    cuuint64_t strides[2] = {(cuuint64_t)D * 4, 16};
    cuuint32_t box[3] = {4, PT_ROWS, PT_PANELS};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_x, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("tma encode (4, rows, D/4) strides (ld*4, 16): CUresult=%d\n", (int)r);
    if (r != CUDA_SUCCESS) return 0;
    size_t smem = PT_PANELS * PT_ROWS * 16 + 64;
    CK(cudaFuncSetAttribute(probe_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> o(PT_PANELS * PT_ROWS * 4);
    for (int trial = 0; trial < 2; ++trial) {
        int row0 = trial ? 70 : 5, group0 = trial ? 32 : 0;   // trial 1 runs past the last row: OOB zero fill
        probe_tma<<<1, 128, smem>>>(tm, d_out, row0, group0);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("tma trial %d: CUDA error %s\n", trial, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int p = 0; p < PT_PANELS; ++p) for (int rr = 0; rr < PT_ROWS; ++rr) for (int e2 = 0; e2 < 4; ++e2) {
            int gr = row0 + rr, gc = (group0 + p) * 4 + e2;
            float ref = (gr < n) ? (float)(gr * 1000 + gc) : 0.f;
            if (o[(p * PT_ROWS + rr) * 4 + e2] != ref) ++bad;
        }
        printf("tma trial %d (row0=%d group0=%d): panel-layout mismatches=%d  [first=%g second_row=%g second_panel=%g]\n",
               trial, row0, group0, bad, o[0], o[4], o[PT_ROWS * 4]);
    }
    return 0;
}

// ------------------------------------------------------------------------ mma2 (pair)
constexpr int P2_R = 33;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192)
probe_mma2(float *out, int offB)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sA = reinterpret_cast<float *>(smem);                 // 32 panels (this CTA's 128 features)
    uint64_t *bars = reinterpret_cast<uint64_t *>(sA + 32 * P2_R * 4);
    uint64_t *bar_ready = bars;       // leader: count 2 (one arrive per CTA)
    uint64_t *bar_done = bars + 1;    // both: MMA complete (multicast commit)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));

    // this CTA holds features [128*rank, 128*rank+128): value pattern aval(feature, row)
    for (int i = tid; i < 32 * P2_R * 4; i += blockDim.x) {
        int e = i & 3, r = (i >> 2) % P2_R, p = (i >> 2) / P2_R;
        int f = cta_rank * 128 + p * 4 + e;
        sA[i] = (f * 7 + r * 3) % 11 - 5;
    }
    if (tid == 0) {
        mbar_init(bar_ready, 2);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    // every CTA tells the leader its operands are in place (remote arrive)
    if (tid == 32) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar_ready)), "r"(0));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
    }
    if (cta_rank == 0 && tid == 64) {
        mbar_wait(bar_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t pitch = P2_R * 16;
        const uint32_t idesc = make_idesc(256, 256);
        uint64_t da = make_desc(smem_u32(sA), 128, pitch);
        uint64_t db = make_desc(smem_u32(sA) + offB * 16, 128, pitch);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
            :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(0u));
        asm volatile(
            "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
            :: "r"(smem_u32(bar_done)), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp < 4) {
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j)
                out[(size_t)(cta_rank * 128 + tid) * 256 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}

static int run_mma2()
{
    float *d_out;
    CK(cudaMalloc(&d_out, sizeof(float) * 256 * 256));
    size_t smem = 32 * P2_R * 16 + 64;
    CK(cudaFuncSetAttribute(probe_mma2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> h(256 * 256);
    for (int offB = 0; offB <= 10; offB += 10) {
        CK(cudaMemset(d_out, 0, sizeof(float) * 256 * 256));
        probe_mma2<<<2, 192, smem>>>(d_out, offB);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mma2 offB=%d: CUDA error %s\n", offB, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(h.data(), d_out, sizeof(float) * 256 * 256, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 256; ++m) for (int n = 0; n < 256; ++n) {
            double ref = 0;
            for (int k = 0; k < 8; ++k) ref += (double)aval(m, k) * aval(n, k + offB);
            if (fabs(ref - h[m * 256 + n]) > 1e-3) ++bad;
        }
        printf("mma2 (cta_group::2, M=256 N=256) offB=%d : mismatches=%d [D00=%g D(0,128)=%g D(128,0)=%g D(255,255)=%g]\n",
               offB, bad, h[0], h[128], h[128 * 256], h[255 * 256 + 255]);
    }
    return 0;
}

int main(int argc, char **argv)
{
    const char *which = argc > 1 ? argv[1] : "mma1";
    if (!strcmp(which, "mma1")) return run_mma1();
    if (!strcmp(which, "tma")) return run_tma();
    if (!strcmp(which, "mma2")) return run_mma2();
    printf("unknown probe %s\n", which);
    return 1;
}
