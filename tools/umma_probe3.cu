// tools/umma_probe3.cu -- third-round probes.
//   tma4 : 4-D tensor map (4 feat, 4 rows, D/4 groups, n/4 row-blocks) -> smem [rb][g][i][4]
//   sw128: MN-major SWIZZLE_128B tf32 operands written by TMA (2-D map, 128B swizzle)
//   pair : cta_group::2, K-major no-swizzle operands, M=256 N=256, two K-steps
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
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)layout << 61;
    return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N, int mn_major)
{
    uint32_t d = 0;
    d |= 1u << 4; d |= 2u << 7; d |= 2u << 10;
    if (mn_major) { d |= 1u << 15; d |= 1u << 16; }
    d |= (uint32_t)(N >> 3) << 17; d |= (uint32_t)(M >> 4) << 24;
    return d;
}
#define LD32(v, taddr) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
    : "r"(taddr))

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn get_encode()
{
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    return (EncodeFn)fn;
}
static inline float xval(int row, int col) { return (float)(((row * 7 + col * 3) % 11) - 5); }

// ------------------------------------------------------------------ tma4
constexpr int T4_RB = 6;   // row-blocks per box  (24 frames)
__global__ void probe_tma4(const __grid_constant__ CUtensorMap tmap, float *out, int rb0, int g0)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);
    const int bytes = T4_RB * 32 * 64;
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + bytes);
    if (threadIdx.x == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                     :: "r"(smem_u32(tile)), "l"(&tmap), "r"(smem_u32(bar)), "r"(0), "r"(0), "r"(g0), "r"(rb0) : "memory");
    }
    mbar_wait(bar, 0);
    for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = tile[i];
}
static int run_tma4()
{
    const int n = 50, D = 256, lag = 10;
    std::vector<float> h((size_t)n * D);
    for (int r = 0; r < n; ++r) for (int c = 0; c < D; ++c) h[(size_t)r * D + c] = r * 1000 + c;
    float *d_x, *d_out;
    CK(cudaMalloc(&d_x, h.size() * 4)); CK(cudaMemcpy(d_x, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, T4_RB * 32 * 64));
    EncodeFn enc = get_encode();
    for (int which = 0; which < 2; ++which) {
        const int P = n - lag, Q = P / 4;                 // full row-blocks of pair indices
        CUtensorMap tm;
        cuuint64_t dims[4] = {4, 4, (cuuint64_t)(D / 4), (cuuint64_t)Q};
        cuuint64_t strides[3] = {(cuuint64_t)D * 4, 16, (cuuint64_t)D * 16};
        cuuint32_t box[4] = {4, 4, 32, T4_RB};
        cuuint32_t es[4] = {1, 1, 1, 1};
        void *base = which ? (void *)(d_x + (size_t)lag * D) : (void *)d_x;
        CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, base, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("tma4 encode (base %s): CUresult=%d\n", which ? "+lag rows" : "0", (int)r);
        if (r != CUDA_SUCCESS) continue;
        size_t smem = T4_RB * 32 * 64 + 64;
        CK(cudaFuncSetAttribute(probe_tma4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        std::vector<float> o(T4_RB * 32 * 16);
        for (int trial = 0; trial < 2; ++trial) {
            int rb0 = trial ? 6 : 0, g0 = trial ? 32 : 0;   // trial 1 reaches past Q=10 blocks -> zero fill
            probe_tma4<<<1, 128, smem>>>(tm, d_out, rb0, g0);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("tma4: CUDA error %s\n", cudaGetErrorString(e)); return 3; }
            CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int rb = 0; rb < T4_RB; ++rb) for (int g = 0; g < 32; ++g) for (int i = 0; i < 4; ++i) for (int e2 = 0; e2 < 4; ++e2) {
                int blk = rb0 + rb, row = 4 * blk + i + (which ? lag : 0), col = (g0 + g) * 4 + e2;
                float ref = (blk < Q) ? (float)(row * 1000 + col) : 0.f;
                if (o[((rb * 32 + g) * 4 + i) * 4 + e2] != ref) ++bad;
            }
            printf("tma4 map%d trial %d: [rb][g][i][4] mismatches=%d (o[0]=%g o[4]=%g o[16]=%g o[512]=%g)\n", which, trial, bad, o[0], o[4], o[16], o[512]);
        }
    }
    return 0;
}

// ------------------------------------------------------------------ sw128 (MN-major, TMA-written, cta_group::1)
// X is (rows x 128 feats). A = B = X tile of 8 rows (K=8), M = N = 128. D[m][n] = sum_k X[k][m] X[k][n].
__global__ void __launch_bounds__(128) probe_sw128(const __grid_constant__ CUtensorMap tmap, float *out, int variant)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *tile = reinterpret_cast<float *>(smem);             // [4 blocks][8 rows][128 B]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 4096);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&bars[0])), "r"(4096) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(tile)), "l"(&tmap), "r"(smem_u32(&bars[0])), "r"(0), "r"(0), "r"(0) : "memory");
        mbar_wait(&bars[0], 0);
        // MN-major SW128: ((T,8,m),(8,k)) : ((1,T,LBO),(8T,SBO)); LBO = stride between 32-feature blocks, SBO = 8 rows
        uint32_t lbo = variant ? 1024 : 1024, sbo = 1024;
        (void)lbo;
        uint64_t d = make_desc(smem_u32(tile), 1024, sbo, 2 /*SWIZZLE_128B*/);
        const uint32_t idesc = make_idesc(128, 128, 1);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                     :: "r"(tmem), "l"(d), "l"(d), "r"(idesc), "r"(0u), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(&bars[1])) : "memory");
    }
    mbar_wait(&bars[1], 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        LD32(v, tmem + ((uint32_t)(warp * 32) << 16) + c0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(size_t)tid * 128 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}
static int run_sw128()
{
    const int n = 16, D = 128;
    std::vector<float> h((size_t)n * D);
    for (int r = 0; r < n; ++r) for (int c = 0; c < D; ++c) h[(size_t)r * D + c] = xval(r, c);
    float *d_x, *d_out;
    CK(cudaMalloc(&d_x, h.size() * 4)); CK(cudaMemcpy(d_x, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, 128 * 128 * 4));
    EncodeFn enc = get_encode();
    CUtensorMap tm;
    cuuint64_t dims[3] = {32, (cuuint64_t)n, 4};
    cuuint64_t strides[2] = {(cuuint64_t)D * 4, 128};
    cuuint32_t box[3] = {32, 8, 4};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_x, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("sw128 encode: CUresult=%d\n", (int)r);
    if (r != CUDA_SUCCESS) return 0;
    size_t smem = 4096 + 64;
    std::vector<float> o(128 * 128);
    CK(cudaMemset(d_out, 0xff, 128 * 128 * 4));
    probe_sw128<<<1, 128, smem>>>(tm, d_out, 0);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("sw128: CUDA error %s\n", cudaGetErrorString(e)); return 3; }
    CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 128; ++m) for (int nn = 0; nn < 128; ++nn) {
        double ref = 0; for (int k = 0; k < 8; ++k) ref += (double)xval(k, m) * xval(k, nn);
        if (fabs(ref - o[m * 128 + nn]) > 1e-3) ++bad;
    }
    printf("sw128 MN-major tf32 (TMA swizzle-128B): mismatches=%d/16384 [D00=%g D01=%g D(1,0)=%g D(40,77)=%g]\n", bad, o[0], o[1], o[128], o[40 * 128 + 77]);
    return 0;
}

// ------------------------------------------------------------------ pair (cta_group::2, K-major)
// Each CTA holds 128 features x 16 frames in K-major blocks [rb(4)][m(128)][4 frames]; lo half unused.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192) probe_pair(float *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sA = reinterpret_cast<float *>(smem);                 // 4 rb x 128 m x 4 = 8 KB
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 8192);
    uint64_t *bar_ready = bars, *bar_done = bars + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    for (int i = tid; i < 4 * 128 * 4; i += blockDim.x) {
        int e = i & 3, m = (i >> 2) % 128, rb = (i >> 2) / 128;
        int f = cta_rank * 128 + m, frame = rb * 4 + e;
        sA[i] = (float)(((frame * 7 + f * 3) % 11) - 5);            // xval(frame, feature)
    }
    if (tid == 0) { mbar_init(bar_ready, 2); mbar_init(bar_done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(256));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    if (tid == 32) {
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar_ready)), "r"(0));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(remote) : "memory");
    }
    if (cta_rank == 0 && tid == 64) {
        mbar_wait(bar_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idesc = make_idesc(256, 256, 0);
        for (int ks = 0; ks < 2; ++ks) {
            uint64_t d = make_desc(smem_u32(sA) + ks * 2 * 2048, 2048, 128, 0);
            uint32_t accum = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                         :: "r"(tmem), "l"(d), "l"(d), "r"(idesc), "r"(accum), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     :: "r"(smem_u32(bar_done)), "h"((uint16_t)3) : "memory");
    }
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (warp < 4) {
        for (int c0 = 0; c0 < 256; c0 += 32) {
            uint32_t v[32];
            LD32(v, tmem + ((uint32_t)(warp * 32) << 16) + c0);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 32; ++j) out[(size_t)(cta_rank * 128 + tid) * 256 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(256));
}
static int run_pair()
{
    float *d_out;
    CK(cudaMalloc(&d_out, 256 * 256 * 4));
    CK(cudaMemset(d_out, 0xff, 256 * 256 * 4));
    size_t smem = 8192 + 64;
    probe_pair<<<2, 192, smem>>>(d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pair: CUDA error %s\n", cudaGetErrorString(e)); return 3; }
    std::vector<float> o(256 * 256);
    CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int m = 0; m < 256; ++m) for (int n = 0; n < 256; ++n) {
        double ref = 0; for (int k = 0; k < 16; ++k) ref += (double)xval(k, m) * xval(k, n);
        if (fabs(ref - o[m * 256 + n]) > 1e-3) ++bad;
    }
    printf("pair cta_group::2 K-major M=256 N=256 K=16: mismatches=%d/65536 [D00=%g D(0,128)=%g D(128,0)=%g D(255,255)=%g D(200,3)=%g]\n",
           bad, o[0], o[128], o[128 * 256], o[255 * 256 + 255], o[200 * 256 + 3]);
    return 0;
}

int main(int argc, char **argv)
{
    const char *which = argc > 1 ? argv[1] : "tma4";
    if (!strcmp(which, "tma4")) return run_tma4();
    if (!strcmp(which, "sw128")) return run_sw128();
    if (!strcmp(which, "pair")) return run_pair();
    return 1;
}
