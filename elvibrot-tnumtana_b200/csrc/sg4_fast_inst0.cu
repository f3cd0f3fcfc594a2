// sg4_fast_inst0.cu -- instantiations of the separable-KEO term kernel (sg4_fast.cuh) that read the 1-D matrix pool
// from global memory (pools larger than 24 KB); launched through fast_launch() of sg4_fast_inst.cu.
#include <cuda_runtime.h>
#include "sg4_fast.cuh"

namespace evr {

#define EVR_FAST0_VARIANTS(X) X(false, false) X(true, false) X(false, true)

int fast0_set_attributes()
{
#define X(rt, tri) \
    if (cudaFuncSetAttribute(sg4_term_kernel_fast<0, rt, tri>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) \
        return fail("evr_sg4: cudaFuncSetAttribute(fast kernel) failed");
    EVR_FAST0_VARIANTS(X)
#undef X
    return 0;
}

int fast0_launch(bool rt, bool tri, int nctas, int nthr, size_t smem, cudaStream_t st,
                 const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
#define X(r_, t_) \
    if (rt == r_ && tri == t_) { sg4_term_kernel_fast<0, r_, t_><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi); return 0; }
    EVR_FAST0_VARIANTS(X)
#undef X
    return fail("evr_sg4: no such fast-kernel variant");
}

} // namespace evr
