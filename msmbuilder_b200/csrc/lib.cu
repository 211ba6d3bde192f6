// lib.cu -- library-level entry points, error plumbing, K1 engine dispatch.
#include "common.cuh"
#include <stdarg.h>
#include <string.h>

namespace msmb {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    return MSMB200_E_CUDA;
}

static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ULL, __ATOMIC_RELAXED); }

int sm_count()
{
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

// engines (tica_simt.cu / tica_umma.cu)
int tica_simt_accumulate(const void *const *seq_ptrs, const int64_t *seq_rows, int n_seq,
                         int D, int64_t ld, int dtype, int lag, double *acc, cudaStream_t st);
bool tica_umma_supported(int D, int64_t ld, int dtype, int lag);
size_t tica_umma_workspace_bytes(int D);
int tica_umma_accumulate(const void *const *seq_ptrs, const int64_t *seq_rows, int n_seq,
                         int D, int64_t ld, int lag, int passes, double *acc, void *workspace,
                         size_t workspace_bytes, cudaStream_t st);

}  // namespace msmb

using namespace msmb;

extern "C" int msmb200_abi_version(void) { return MSMB200_ABI_VERSION; }

extern "C" const char *msmb200_last_error(void) { return g_err; }

extern "C" uint64_t msmb200_launch_count(void)
{
    return __atomic_load_n(&g_launches, __ATOMIC_RELAXED);
}

extern "C" int msmb200_device_info(int device, int *sm, int *cc_major, int *cc_minor,
                                   size_t *total_mem)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        set_error("no CUDA device %d visible (%s)", device,
                  e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
        (void)cudaGetLastError();
        return MSMB200_E_NODEVICE;
    }
    cudaDeviceProp p;
    MSMB_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm) *sm = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return MSMB200_OK;
}

extern "C" size_t msmb200_tica_acc_len(int D)
{
    return 3 * (size_t)D * D + 3 * (size_t)D + 2;
}

extern "C" size_t msmb200_tica_workspace_bytes(int D, int engine)
{
    if (engine == MSMB200_TICA_SIMT_F64) return 256;
    size_t b = tica_umma_workspace_bytes(D);
    return b < 256 ? 256 : b;
}

extern "C" int msmb200_tica_accumulate(const void *const *seq_ptrs, const int64_t *seq_rows,
                                       int n_seq, int D, int64_t ld, int dtype, int lag,
                                       int engine, double *acc, void *workspace,
                                       size_t workspace_bytes, void *stream)
{
    MSMB_REQUIRE(n_seq >= 0 && D > 0 && ld >= D && lag >= 1, "tica_accumulate: bad shape "
                 "n_seq=%d D=%d ld=%lld lag=%d", n_seq, D, (long long)ld, lag);
    MSMB_REQUIRE(acc != nullptr, "tica_accumulate: null accumulator");
    MSMB_REQUIRE(dtype == MSMB200_F32 || dtype == MSMB200_F64, "tica_accumulate: bad dtype");
    if (n_seq == 0) return MSMB200_OK;
    MSMB_REQUIRE(seq_ptrs && seq_rows, "tica_accumulate: null sequence table");
    for (int s = 0; s < n_seq; ++s)
        MSMB_REQUIRE(seq_rows[s] >= 0 && (seq_rows[s] == 0 || seq_ptrs[s]),
                     "tica_accumulate: sequence %d invalid", s);
    cudaStream_t st = (cudaStream_t)stream;
    const bool umma_ok = tica_umma_supported(D, ld, dtype, lag);
    if (engine == MSMB200_TICA_AUTO)   // the 256-wide tensor-core tiles beat the float64 kernel from D = 64 up
        engine = (umma_ok && D >= 64) ? MSMB200_TICA_UMMA_3XF16 : MSMB200_TICA_SIMT_F64;
    if (engine == MSMB200_TICA_SIMT_F64)
        return tica_simt_accumulate(seq_ptrs, seq_rows, n_seq, D, ld, dtype, lag, acc, st);
    if (engine == MSMB200_TICA_UMMA_3XTF32 || engine == MSMB200_TICA_UMMA_TF32 ||
        engine == MSMB200_TICA_UMMA_3XBF16 || engine == MSMB200_TICA_UMMA_6XBF16 ||
        engine == MSMB200_TICA_UMMA_3XF16) {
        if (!umma_ok) {
            set_error("tcgen05 engine does not take D=%d ld=%lld dtype=%d lag=%d", D,
                      (long long)ld, dtype, lag);
            return MSMB200_E_UNSUPPORTED;
        }
        const int mode = engine == MSMB200_TICA_UMMA_3XTF32 ? 3 : engine == MSMB200_TICA_UMMA_TF32 ? 1
                         : engine == MSMB200_TICA_UMMA_3XBF16 ? 13 : engine == MSMB200_TICA_UMMA_3XF16 ? 23 : 16;
        return tica_umma_accumulate(seq_ptrs, seq_rows, n_seq, D, ld, lag, mode, acc, workspace,
                                    workspace_bytes, st);
    }
    set_error("tica_accumulate: unknown engine %d", engine);
    return MSMB200_E_INVALID;
}
