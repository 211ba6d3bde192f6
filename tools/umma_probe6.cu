// tools/umma_probe6.cu -- reciprocal throughput of the instructions the K1 converters are made of
// (per SM sub-partition, cycles per warp instruction at 4 resident warps per scheduler).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe6 tools/umma_probe6.cu && ./umma_probe6
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); return 2; } } while (0)

template <int OP>
__global__ void __launch_bounds__(512) tput(long long *cycles, uint32_t *sink, int iters, uint32_t seed)
{
    // 8 independent chains per thread
    uint32_t a[8], b[8];
    uint64_t d[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + 0x3c003c00u + threadIdx.x * 8 + i; b[i] = a[i] ^ 0x01010101u; d[i] = ((uint64_t)a[i] << 32) | b[i]; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a[i]) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(b[i])));
                else if (OP == 1) { float f; asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f) : "r"(a[i])); a[i] = __float_as_uint(f); }
                else if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(d[(i + 1) & 7]));
                else if (OP == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(d[(i + 1) & 7]));
                else if (OP == 4) { float f; asm volatile("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %1; sub.rn.f32.f16 %0, lo, %2;}" : "=f"(f) : "r"(a[i]), "f"(__uint_as_float(b[i]))); a[i] = __float_as_uint(f); }
                else if (OP == 5) asm volatile("max.u16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                else if (OP == 6) asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                else if (OP == 7) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
                else if (OP == 8) asm volatile("fma.rn.f32 %0, %1, %2, %1;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]));
                else if (OP == 9) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(a[i]) : "r"(0x38003800u));
                else if (OP == 10) asm volatile("add.rn.f32 %0, %1, %2;" : "=r"(a[i]) : "r"(a[i]), "r"(b[i]));
            }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
    for (int i = 0; i < 8; ++i) s ^= a[i] ^ (uint32_t)d[i] ^ (uint32_t)(d[i] >> 32);
    if (s == 0x12345) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int OP>
int run(const char *name, long long *d_c, uint32_t *d_s)
{
    const int iters = 2000;
    tput<OP><<<148, 512>>>(d_c, d_s, iters, 1);
    CK(cudaDeviceSynchronize());
    long long c;
    CK(cudaMemcpy(&c, d_c, 8, cudaMemcpyDeviceToHost));
    // 512 threads = 16 warps = 4 per scheduler; each warp issues iters * 32 instructions
    printf("%-34s %.2f cycles per warp instruction per scheduler (4 warps resident)\n", name,
           (double)c / (iters * 32.0 * 4.0));
    return 0;
}

int main()
{
    long long *d_c; uint32_t *d_s;
    CK(cudaMalloc(&d_c, 8)); CK(cudaMalloc(&d_s, 4));
    run<0>("F2FP.F16.F32.PACK_AB (cvt f16x2)", d_c, d_s);
    run<1>("HADD2.F32 (cvt.f32.f16)", d_c, d_s);
    run<2>("FFMA2 (fma.f32x2)", d_c, d_s);
    run<3>("FADD2 (add.f32x2)", d_c, d_s);
    run<4>("FHADD (sub.f32.f16)", d_c, d_s);
    run<5>("VIMNMX.U16x2", d_c, d_s);
    run<6>("LOP3 (xor)", d_c, d_s);
    run<7>("IADD3", d_c, d_s);
    run<8>("FFMA", d_c, d_s);
    run<9>("HMUL2 imm", d_c, d_s);
    run<10>("FADD", d_c, d_s);
    return 0;
}
