// tools/umma_probe7.cu -- can K1 use MN-major fp16 operands (features contiguous, a frame = a 128-byte row,
// SWIZZLE_128B) so that the lagged operand is the SAME converted buffer at a row offset?  
// Not part of the library.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe7 tools/umma_probe7.cu && ./umma_probe7
//
// One CTA, cta_group::1, kind::f16, M = 128 features (2 blocks of 64) x N = 128 features, K = 16 frames per MMA.
// Window in shared memory: [block][row][128 B]; element (feature m, row r) at
//   block (m / 64) * BLK + r * 128 + ((((m % 64) / 8) ^ (r & 7)) << 4) + (m % 8) * 2        (Swizzle<3,4,3>)
// Variants: LBO / SBO roles, base_offset field, row offsets of the two operands (0 / 8 / 10 / odd), K steps.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
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

constexpr int ROWS = 64;                 // rows (frames) of the window
constexpr int BLK = ROWS * 128;          // bytes of one 64-feature block
constexpr int PM = 128, PN = 128;

static inline float aval(int m, int r) { return (float)(((m * 7 + r * 3) % 11) - 5); }

// layout_type: 2 = SWIZZLE_128B (sm_100 encoding, bits 61-63)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t base_off,
                                              uint32_t layout_type)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                       // version = 1 (Blackwell)
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)(layout_type & 7) << 61;
    return d;
}
__host__ __device__ inline uint32_t make_idesc(int M, int N)
{
    uint32_t d = 0;
    d |= 1u << 4;            // c_format = F32; a_format = b_format = F16 (0)
    d |= 1u << 15;           // a_major = MN
    d |= 1u << 16;           // b_major = MN
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__global__ void __launch_bounds__(128)
probe(float *out, int swap_lbo_sbo, int use_base_off, int offA, int offB, int ksteps)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char *win = smem;                                          // 2 blocks
    uint64_t *bar = reinterpret_cast<uint64_t *>(win + 2 * BLK);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < PM * ROWS; i += 128) {
        const int m = i % PM, r = i / PM;
        const int blk = m / 64, c = (m % 64) / 8, e = m % 8;
        __half *p = reinterpret_cast<__half *>(win + blk * BLK + r * 128 + ((c ^ (r & 7)) << 4) + e * 2);
        *p = __float2half((float)(((m * 7 + r * 3) % 11) - 5));
    }
    if (tid == 0) mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     :: "r"(smem_u32(tmem_slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        const uint32_t lbo = swap_lbo_sbo ? 1024 : BLK, sbo = swap_lbo_sbo ? BLK : 1024;
        const uint32_t idesc = make_idesc(PM, PN);
        for (int ks = 0; ks < ksteps; ++ks) {
            const uint32_t sa = smem_u32(win) + (offA + 16 * ks) * 128;
            const uint32_t sb = smem_u32(win) + (offB + 16 * ks) * 128;
            uint64_t da = make_desc(sa, lbo, sbo, use_base_off ? (sa >> 7) & 7 : 0, 2);
            uint64_t db = make_desc(sb, lbo, sbo, use_base_off ? (sb >> 7) & 7 : 0, 2);
            uint32_t accum = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
                :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accum), "r"(0u));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                     :: "r"(smem_u32(bar)) : "memory");
    }
    mbar_wait(bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < PN; c0 += 32) {
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
        for (int j = 0; j < 32; ++j) out[(size_t)tid * PN + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"(128));
}

int main()
{
    float *d_out;
    CK(cudaMalloc(&d_out, sizeof(float) * PM * PN));
    size_t smem = 2 * BLK + 1024 + 64;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<float> h(PM * PN);
    const int offs[6][2] = {{0, 0}, {0, 8}, {0, 10}, {3, 13}, {16, 26}, {5, 5}};
    int ok = 0, total = 0;
    for (int swap = 0; swap < 2; ++swap)
        for (int bo = 0; bo < 2; ++bo)
            for (int oi = 0; oi < 6; ++oi)
                for (int ks = 1; ks <= 2; ++ks) {
                    const int offA = offs[oi][0], offB = offs[oi][1];
                    CK(cudaMemset(d_out, 0, sizeof(float) * PM * PN));
                    probe<<<1, 128, smem>>>(d_out, swap, bo, offA, offB, ks);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("swap=%d bo=%d offA=%d offB=%d ks=%d: CUDA error %s\n", swap, bo, offA, offB, ks, cudaGetErrorString(e)); return 3; }
                    CK(cudaMemcpy(h.data(), d_out, sizeof(float) * PM * PN, cudaMemcpyDeviceToHost));
                    int bad = 0; double maxerr = 0;
                    for (int m = 0; m < PM; ++m)
                        for (int n = 0; n < PN; ++n) {
                            double ref = 0;
                            for (int k = 0; k < 16 * ks; ++k) ref += (double)aval(m, k + offA) * aval(n, k + offB);
                            double err = fabs(ref - h[m * PN + n]);
                            if (err > 1e-3) ++bad;
                            if (err > maxerr) maxerr = err;
                        }
                    printf("swap_lbo_sbo=%d base_off=%d offA=%2d offB=%2d ksteps=%d : mismatches=%5d maxerr=%g  [D00=%g D01=%g D10=%g]\n",
                           swap, bo, offA, offB, ks, bad, maxerr, h[0], h[1], h[PN]);
                    ++total;
                    if (!bad) ++ok;
                }
    printf("configs fully correct: %d / %d\n", ok, total);
    return 0;
}
