// tools/umma_probe5.cu -- tcgen05 kind::f16 MMA rate by shape: cta_group::2 (M = 256) and cta_group::1
// (M = 128), N = 256 / 128 / 64, K = 16, K-major SWIZZLE_NONE operands static in shared memory.
// Variants per shape: a) one accumulator, b) round robin over the accumulators that fit 512 columns,
// c) like b with .collector::a::fill / use / lastuse on groups of four MMAs that share A.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe5 tools/umma_probe5.cu && ./umma_probe5
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(2); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((2048 >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// K-major SWIZZLE_128B: rows of 128 bytes (64 fp16 along K), 8-row atoms of 1024 bytes
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((16 >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4;                       // f32 accumulate, f16 x f16
    d |= (uint32_t)(N >> 3) << 17; d |= (uint32_t)(M >> 4) << 24;
    return d;
}
#define MMA(CGS, VEC, QUAL, tm, da, db, idesc, acc) asm volatile( \
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" \
    "tcgen05.mma.cta_group::" CGS ".kind::f16" QUAL " [%0], %1, %2, %3, " VEC ", p;\n\t}" \
    :: "r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory")
#define V8 "{%5, %5, %5, %5, %5, %5, %5, %5}"
#define V4 "{%5, %5, %5, %5}"

constexpr int TILE = 8192;   // one 32-frame x 128-feature fp16 operand tile (2 K steps of 16)

template <int CG>
__global__ void __launch_bounds__(128)
rate_kernel(int N, int variant, int n_iter, long long *cycles)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ops = smem;                              // 5 tiles
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 5 * TILE);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t cta_rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    for (int i = tid; i < 5 * TILE / 2; i += blockDim.x) reinterpret_cast<unsigned short *>(ops)[i] = 0x3800;   // 0.5
    if (tid == 0) { mbar_init(&bars[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        if (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    uint32_t elected = 0;
    if (cta_rank == 0 && warp == 1) {
        uint32_t laneid = 0;
        asm volatile("{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\telect.sync %%rx|%%px, %2;\n\t@%%px mov.s32 %1, 1;\n\tmov.s32 %0, %%rx;\n\t}"
                     : "+r"(laneid), "+r"(elected) : "r"(0xFFFFFFFFu));
    }
    if (elected) {
        const uint32_t idesc = make_idesc(128 * CG, N);
        const uint32_t base = smem_u32(ops);
        const int n_acc = (variant == 0 || variant == 3) ? 1 : 512 / N;
        const bool sw = variant == 3;
        // SWIZZLE_128B tiles: 128 rows x 128 B = 16 KB each (A at 0, A2 at 16 KB, B at 32 KB; the three
        // "tiles" of the no-swizzle variants alias the same memory, contents do not matter here)
        const uint64_t dA = sw ? make_desc_sw128(base) : make_desc(base);
        const uint64_t dA2 = sw ? make_desc_sw128(base + 32) : make_desc(base + TILE);
        long long t0 = clock64();
        for (int it = 0; it < n_iter; ++it) {
            // 8 MMAs per iteration: two groups of four that share A
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const uint64_t da = g ? dA2 : dA;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int m = g * 4 + j;
                    const uint32_t d = tmem + (uint32_t)((m % n_acc) * N);
                    const uint64_t db = sw ? make_desc_sw128(base + 16384 + (j & 3) * 32)
                                           : make_desc(base + (2 + (j % 3)) * TILE + (j & 1) * 4096);
                    const uint32_t acc = it ? 1u : 0u;
                    if (CG == 2) {
                        if (variant == 2) {
                            if (j == 0) MMA("2", V8, ".collector::a::fill", d, da, db, idesc, acc);
                            else if (j == 3) MMA("2", V8, ".collector::a::lastuse", d, da, db, idesc, 1u);
                            else MMA("2", V8, ".collector::a::use", d, da, db, idesc, 1u);
                        } else MMA("2", V8, "", d, da, db, idesc, acc);
                    } else {
                        if (variant == 2) {
                            if (j == 0) MMA("1", V4, ".collector::a::fill", d, da, db, idesc, acc);
                            else if (j == 3) MMA("1", V4, ".collector::a::lastuse", d, da, db, idesc, 1u);
                            else MMA("1", V4, ".collector::a::use", d, da, db, idesc, 1u);
                        } else MMA("1", V4, "", d, da, db, idesc, acc);
                    }
                }
            }
        }
        if (CG == 2)
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         :: "r"(smem_u32(&bars[0])), "h"((uint16_t)3) : "memory");
        else
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                         :: "r"(smem_u32(&bars[0])) : "memory");
        mbar_wait(&bars[0], 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    if (!elected) mbar_wait(&bars[0], 0);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (CG == 2) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) {
        if (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
    }
}

template <int CG>
static void run(int N, int variant, int grid, long long *d_cyc)
{
    const int n_iter = 4000;
    const size_t smem = 5 * TILE + 1024;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, rate_kernel<CG>, N, variant, n_iter, d_cyc));
    CK(cudaDeviceSynchronize());
    long long c = 0;
    CK(cudaMemcpy(&c, d_cyc, sizeof(c), cudaMemcpyDeviceToHost));
    const double ideal = 128.0 * N / 256.0;
    printf("cta_group::%d M=%d N=%3d %-28s grid %3d: %.1f cycles/MMA (ideal %.0f)\n", CG, 128 * CG, N,
           variant == 0 ? "one accumulator" : variant == 1 ? "round-robin accumulators"
           : variant == 2 ? "round robin + A collector" : "one accumulator, SWIZZLE_128B",
           grid, (double)c / (n_iter * 8.0), ideal);
}

int main()
{
    long long *d_cyc;
    CK(cudaMalloc(&d_cyc, 8));
    CK(cudaFuncSetAttribute(rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * TILE + 1024));
    CK(cudaFuncSetAttribute(rate_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5 * TILE + 1024));
    const int Ns[3] = {256, 128, 64};
    for (int n = 0; n < 3; ++n)
        for (int v = 0; v < 4; ++v) {
            if (v == 1) continue;
            run<2>(Ns[n], v, 2, d_cyc);
            run<1>(Ns[n], v, 1, d_cyc);
        }
    for (int n = 0; n < 3; ++n) {
        run<2>(Ns[n], 2, 148, d_cyc);
        run<1>(Ns[n], 2, 148, d_cyc);
    }
    return 0;
}
