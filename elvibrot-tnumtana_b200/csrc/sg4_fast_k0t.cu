// sg4_fast_k0t.cu -- ONE instantiation of the separable-KEO term kernel (sg4_fast.cuh): matrices read from global memory (pool > 24 KB),
// cube tiles (512 threads).  One kernel per translation unit keeps the parallel build bounded by the slowest kernel.
#include <cuda_runtime.h>
#include "sg4_fast.cuh"

namespace evr {

int fast_attr_0t()
{
    if (cudaFuncSetAttribute(sg4_term_kernel_fast<0, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return fail("evr_sg4: cudaFuncSetAttribute(fast kernel <0,0,1>) failed");
    return 0;
}

int fast_launch_0t(int nctas, int nthr, size_t smem, cudaStream_t st,
                      const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
    sg4_term_kernel_fast<0, false, true><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    return 0;
}

} // namespace evr
