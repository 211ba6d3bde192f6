// tools/umma_probe4.cu -- tcgen05 tf32 MMA rate microbenchmark on a CTA pair (cta_group::2,
// M=256 N=256 K=8), operands static in smem, no other traffic.  Variants:
//   0: K-major SWIZZLE_NONE (LBO=2048,SBO=128), 6 MMAs/K-step pattern of the tICA kernel
//   1: same with .collector::a::fill/use/lastuse reuse of the A operand
//   2: K-major SWIZZLE_128B operands (rows of 128 B), same 6-MMA pattern
//   3: variant 0 + 256 extra threads hammering shared memory (ld/st) like the converters
// Also checks numerics of the SWIZZLE_128B K-major layout on a small known matrix (mode "check").
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
__host__ __device__ inline uint32_t make_idesc(int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4; d |= 2u << 7; d |= 2u << 10;
    d |= (uint32_t)(N >> 3) << 17; d |= (uint32_t)(M >> 4) << 24;
    return d;
}
#define MMA2(QUAL, tm, da, db, idesc, acc) asm volatile( \
    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t" \
    "tcgen05.mma.cta_group::2.kind::tf32" QUAL " [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" \
    :: "r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory")

constexpr int TILE = 16384;   // one 32-frame x 128-feature operand tile

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384)
rate_kernel(int variant, int n_tiles, long long *cycles, float *sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *ops = smem;                              // 4 tiles: A_hi A_lo B_hi B_lo
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 4 * TILE + 65536);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5;
    uint32_t cta_rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
    for (int i = tid; i < 4 * TILE / 4; i += blockDim.x) reinterpret_cast<float *>(ops)[i] = (float)((i * 7) % 5) - 2.f;
    if (tid == 0) { mbar_init(&bars[0], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    if (cta_rank == 0 && tid == 32) {
        const uint32_t idesc = make_idesc(256, 256);
        const uint32_t base = smem_u32(ops);
        const bool sw = (variant == 2);
        long long t0 = clock64();
        for (int t = 0; t < n_tiles; ++t) {
            for (int ks = 0; ks < 4; ++ks) {
                const uint32_t off = sw ? ks * 32 : ks * 4096;
                const uint32_t lbo = sw ? 16 : 2048, sbo = sw ? 1024 : 128, lay = sw ? 2 : 0;
                const uint64_t dAh = make_desc(base + off, lbo, sbo, lay), dAl = make_desc(base + TILE + off, lbo, sbo, lay);
                const uint64_t dBh = make_desc(base + 2 * TILE + off, lbo, sbo, lay), dBl = make_desc(base + 3 * TILE + off, lbo, sbo, lay);
                const uint32_t acc = (t | ks) ? 1u : 0u;
                if (variant == 1) {
                    MMA2(".collector::a::fill", tmem, dAh, dBh, idesc, acc);
                    MMA2(".collector::a::use", tmem, dAh, dBl, idesc, 1u);
                    MMA2(".collector::a::use", tmem + 256, dAh, dAh, idesc, acc);
                    MMA2(".collector::a::lastuse", tmem + 256, dAh, dAl, idesc, 1u);
                    MMA2(".collector::a::fill", tmem, dAl, dBh, idesc, 1u);
                    MMA2(".collector::a::lastuse", tmem + 256, dAl, dAh, idesc, 1u);
                } else {
                    MMA2("", tmem, dAh, dBh, idesc, acc);
                    MMA2("", tmem, dAh, dBl, idesc, 1u);
                    MMA2("", tmem, dAl, dBh, idesc, 1u);
                    MMA2("", tmem + 256, dAh, dAh, idesc, acc);
                    MMA2("", tmem + 256, dAh, dAl, idesc, 1u);
                    MMA2("", tmem + 256, dAl, dAh, idesc, 1u);
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     :: "r"(smem_u32(&bars[0])), "h"((uint16_t)3) : "memory");
        mbar_wait(&bars[0], 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    } else if (variant == 3 && warp >= 4) {
        // converter-like shared memory traffic on the spare 64 KB until the MMAs are done
        float4 *scratch = reinterpret_cast<float4 *>(smem + 4 * TILE);
        float4 acc4 = make_float4(0, 0, 0, 0);
        volatile uint64_t *done = &bars[0];
        (void)done;
        for (int it = 0; it < n_tiles * 12; ++it) {
            const int j = (tid - 128 + it * 256) & 4095;
            float4 v = scratch[j];
            acc4.x += v.x; acc4.y += v.y; acc4.z += v.z; acc4.w += v.w;
            scratch[(j + 2048) & 4095] = acc4;
            scratch[(j + 1024) & 4095] = v;
        }
        if (acc4.x == 12345.f) sink[0] = acc4.y;
    }
    if (!(cta_rank == 0 && tid == 32)) mbar_wait(&bars[0], 0);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(512));
}

// ---- numerics of SWIZZLE_128B K-major tf32 (cta_group::1, M=128, N=64, K=16 over two K-steps)
#define LD32(v, taddr) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
    : "r"(taddr))
static inline float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 11) - 5); }
static inline float bval(int n, int k) { return (float)(((n * 5 + k * 2) % 13) - 6); }
__global__ void __launch_bounds__(128) check_sw128(float *out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sA = reinterpret_cast<float *>(smem);            // 128 rows x 128 B
    float *sB = reinterpret_cast<float *>(smem + 16384);    // 64 rows x 128 B
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384 + 8192);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 32; i += 128) { int m = i / 32, k = i % 32; int chunk = (k / 4) ^ (m & 7); sA[m * 32 + chunk * 4 + (k & 3)] = ((m * 7 + k * 3) % 11) - 5; }
    for (int i = tid; i < 64 * 32; i += 128) { int n = i / 32, k = i % 32; int chunk = (k / 4) ^ (n & 7); sB[n * 32 + chunk * 4 + (k & 3)] = ((n * 5 + k * 2) % 13) - 6; }
    if (tid == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(64));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = make_idesc(128, 64);
        for (int ks = 0; ks < 2; ++ks) {
            uint64_t da = make_desc(smem_u32(sA) + (ks + 1) * 32, 16, 1024, 2);   // frames 8..23
            uint64_t db = make_desc(smem_u32(sB) + (ks + 1) * 32, 16, 1024, 2);
            uint32_t acc = ks > 0;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                         :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        LD32(v, tmem + ((uint32_t)(warp * 32) << 16) + c0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}

int main()
{
    {   // numerics
        float *d_out; CK(cudaMalloc(&d_out, 128 * 64 * 4));
        size_t smem = 16384 + 8192 + 64;
        check_sw128<<<1, 128, smem>>>(d_out);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("check_sw128: CUDA error %s\n", cudaGetErrorString(e)); return 3; }
        std::vector<float> h(128 * 64);
        CK(cudaMemcpy(h.data(), d_out, h.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
            double ref = 0; for (int k = 8; k < 24; ++k) ref += (double)aval(m, k) * bval(n, k);
            if (fabs(ref - h[m * 64 + n]) > 1e-3) ++bad;
        }
        printf("check K-major SWIZZLE_128B tf32 (start +32B per K-step, SBO=1024): mismatches=%d/8192 [D00=%g D01=%g D10=%g]\n", bad, h[0], h[1], h[64]);
    }
    long long *d_cyc; float *d_sink;
    CK(cudaMalloc(&d_cyc, 8)); CK(cudaMalloc(&d_sink, 4));
    size_t smem = 4 * TILE + 65536 + 64;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int n_tiles = 2000;
    for (int variant = 0; variant < 4; ++variant) {
        for (int rep = 0; rep < 2; ++rep) {
            rate_kernel<<<2, 384, smem>>>(variant, n_tiles, d_cyc, d_sink);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("variant %d: CUDA error %s\n", variant, cudaGetErrorString(e)); return 3; }
        }
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("variant %d: %lld cycles for %d MMAs -> %.1f cycles/MMA (ideal 128)\n", variant, c, n_tiles * 24, (double)c / (n_tiles * 24));
    }
    // all SMs busy: 74 pairs
    for (int variant = 0; variant < 3; variant += 2) {
        rate_kernel<<<148, 384, smem>>>(variant, n_tiles, d_cyc, d_sink);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost));
        printf("variant %d on 74 pairs: %.1f cycles/MMA\n", variant, (double)c / (n_tiles * 24));
    }
    return 0;
}
