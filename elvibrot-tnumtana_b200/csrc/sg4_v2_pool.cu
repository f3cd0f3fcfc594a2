// sg4_v2_pool.cu -- instantiation of the second-generation term kernel (sg4_fast2.cuh) whose 1-D matrices come from the
// de-duplicated pool in shared memory (any set of per-mode matrices, e.g. pyrazine: twelve different frequencies).
#include <cuda_runtime.h>
#include "sg4_fast2.cuh"

namespace evr {

int v2_pool_set_attributes()
{
    if (cudaFuncSetAttribute(sg4_term_kernel_v2<1, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(sg4_term_kernel_v2<1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return fail("evr_sg4: cudaFuncSetAttribute(v2 pool kernel) failed");
    return 0;
}

int v2_pool_launch(int nctas, int nthr, size_t smem, cudaStream_t st,
                   const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
    if (nthr > 512) sg4_term_kernel_v2<1, 768><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    else sg4_term_kernel_v2<1, 512><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    return 0;
}

} // namespace evr
