// sg4_fast_k1r.cu -- ONE instantiation of the separable-KEO term kernel (sg4_fast.cuh): matrices in the shared-memory pool,
// runtime-size tiles.  One kernel per translation unit keeps the parallel build bounded by the slowest kernel.
#include <cuda_runtime.h>
#include "sg4_fast.cuh"

namespace evr {

int fast_attr_1r()
{
    if (cudaFuncSetAttribute(sg4_term_kernel_fast<1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return fail("evr_sg4: cudaFuncSetAttribute(fast kernel <1,1,0>) failed");
    return 0;
}

int fast_launch_1r(int nctas, int nthr, size_t smem, cudaStream_t st,
                      const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
    sg4_term_kernel_fast<1, true, false><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    return 0;
}

} // namespace evr
