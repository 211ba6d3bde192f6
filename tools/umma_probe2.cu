// tools/umma_probe2.cu -- second-round probes: TMEM st/ld round trip, K-major vs
// MN-major no-swizzle tf32 operands.   ./umma_probe2
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo, int version)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)version << 46;
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
static inline float aval(int m, int k) { return (float)(((m * 7 + k * 3) % 11) - 5); }
static inline float bval(int n, int k) { return (float)(((n * 5 + k * 2) % 13) - 6); }

#define LD32(v, taddr) asm volatile( \
    "tcgen05.ld.sync.aligned.32x32b.x32.b32 " \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, " \
    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" \
    : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), \
      "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), \
      "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), \
      "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]) \
    : "r"(taddr))

constexpr int M = 128, N = 64, R = 16;   // R rows of K capacity in MN-major panels

// mode 0: TMEM st/ld round trip (no MMA)
// mode 1: K-major no-swizzle   A[kb][m][4], B[kb][n][4]  (LBO = between k-blocks, SBO = 128)
// mode 2: same, LBO/SBO swapped
// mode 3: MN-major panels, LBO=128, SBO=pitch
// mode 4: MN-major panels, LBO=pitch, SBO=128
// mode 5/6: as 3/4 with descriptor version 0
__global__ void __launch_bounds__(128) probe(float *out, int mode, uint32_t *dbg)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    float *sA = reinterpret_cast<float *>(smem);
    float *sB = sA + 4096;                    // 16 KB each
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 32768);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 8192; i += 128) sA[i] = 0.f;
    __syncthreads();
    if (mode == 1 || mode == 2) {
        for (int i = tid; i < 2 * M * 4; i += 128) { int e = i & 3, m = (i >> 2) % M, kb = (i >> 2) / M; sA[i] = ((m * 7 + (kb * 4 + e) * 3) % 11) - 5; }
        for (int i = tid; i < 2 * N * 4; i += 128) { int e = i & 3, n = (i >> 2) % N, kb = (i >> 2) / N; sB[i] = ((n * 5 + (kb * 4 + e) * 2) % 13) - 6; }
    } else if (mode >= 3) {
        for (int i = tid; i < (M / 4) * R * 4; i += 128) { int e = i & 3, r = (i >> 2) % R, p = (i >> 2) / R; sA[i] = (((p * 4 + e) * 7 + r * 3) % 11) - 5; }
        for (int i = tid; i < (N / 4) * R * 4; i += 128) { int e = i & 3, r = (i >> 2) % R, p = (i >> 2) / R; sB[i] = (((p * 4 + e) * 5 + r * 2) % 13) - 6; }
    }
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
    if (tid == 0) dbg[0] = tmem;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);

    if (mode == 0) {
        // each thread stores 32 columns: value = row*1000 + col
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint((float)(tid * 1000 + c0 + j));
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
                "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
                "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                :: "r"(lane_base + c0),
                   "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                   "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
                   "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
                   "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    } else {
        if (tid == 0) {
            uint32_t lbo, sbo; int ver = (mode >= 5) ? 0 : 1; int mn = mode >= 3;
            uint64_t da, db;
            if (mode == 1) { da = make_desc(smem_u32(sA), M * 16, 128, 1); db = make_desc(smem_u32(sB), N * 16, 128, 1); }
            else if (mode == 2) { da = make_desc(smem_u32(sA), 128, M * 16, 1); db = make_desc(smem_u32(sB), 128, N * 16, 1); }
            else {
                int m2 = (mode - 3) & 1;
                lbo = m2 ? R * 16 : 128; sbo = m2 ? 128 : R * 16;
                da = make_desc(smem_u32(sA), lbo, sbo, ver); db = make_desc(smem_u32(sB), lbo, sbo, ver);
            }
            dbg[1] = (uint32_t)da; dbg[2] = (uint32_t)(da >> 32); dbg[3] = make_idesc(M, N, mn);
            const uint32_t idesc = make_idesc(M, N, mn);
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u), "r"(0u) : "memory");
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
        }
        mbar_wait(bar, 0);
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        LD32(v, lane_base + c0);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 32; ++j) out[(size_t)tid * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(64));
}

int main()
{
    float *d_out; uint32_t *d_dbg;
    CK(cudaMalloc(&d_out, sizeof(float) * M * N));
    CK(cudaMalloc(&d_dbg, 64));
    size_t smem = 32768 + 64;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> h(M * N);
    uint32_t dbg[4];
    for (int mode = 0; mode <= 6; ++mode) {
        CK(cudaMemset(d_out, 0xff, sizeof(float) * M * N));
        CK(cudaMemset(d_dbg, 0, 64));
        probe<<<1, 128, smem>>>(d_out, mode, d_dbg);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("mode %d: CUDA error %s\n", mode, cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(h.data(), d_out, sizeof(float) * M * N, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(dbg, d_dbg, 16, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
            double ref;
            if (mode == 0) ref = m * 1000 + n;
            else { ref = 0; for (int k = 0; k < 8; ++k) ref += (double)aval(m, k) * bval(n, k); }
            if (fabs(ref - h[m * N + n]) > 1e-3) ++bad;
        }
        printf("mode %d: mismatches=%d/%d tmem=0x%x desc_lo=0x%x desc_hi=0x%x idesc=0x%x | row0: %g %g %g %g | row1: %g %g | row 64: %g %g\n",
               mode, bad, M * N, dbg[0], dbg[1], dbg[2], dbg[3], h[0], h[1], h[2], h[3], h[N], h[N + 1], h[64 * N], h[64 * N + 1]);
        double r00 = 0, r01 = 0; for (int k = 0; k < 8; ++k) { r00 += aval(0, k) * bval(0, k); r01 += aval(0, k) * bval(1, k); }
        if (mode == 1) printf("   expected row0: %g %g\n", r00, r01);
    }
    return 0;
}
