// tica_umma.cu -- K1 on tcgen05 tensor cores (placeholder until the kernel lands).
#include "common.cuh"
namespace msmb {
bool tica_umma_supported(int D, int64_t ld, int dtype, int lag) { (void)D; (void)ld; (void)dtype; (void)lag; return false; }
size_t tica_umma_workspace_bytes(int D) { (void)D; return 0; }
int tica_umma_accumulate(const void *const *, const int64_t *, int, int, int64_t, int, int, double *,
                         void *, size_t, cudaStream_t)
{
    set_error("tcgen05 engine not built");
    return MSMB200_E_UNSUPPORTED;
}
}  // namespace msmb
