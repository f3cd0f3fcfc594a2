// sg4_fast_inst.cu -- instantiations and launchers of the separable-KEO term kernel (sg4_fast.cuh) whose 1-D
// matrices are read from the shared-memory pool (any set of per-mode matrices); the variants that read the pool from
// global memory live in sg4_fast_inst0.cu (separate translation unit: parallel build).
#include <cuda_runtime.h>
#include "sg4_fast.cuh"

namespace evr {

#define EVR_FAST_VARIANTS(X) X(1, false, false) X(1, true, false) X(1, false, true)

int fast0_set_attributes();
int fast0_launch(bool rt, bool tri, int nctas, int nthr, size_t smem, cudaStream_t st,
                 const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi);

int fast_set_attributes()
{
    if (fast0_set_attributes()) return 1;
#define X(mm, rt, tri) \
    if (cudaFuncSetAttribute(sg4_term_kernel_fast<mm, rt, tri>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) \
        return fail("evr_sg4: cudaFuncSetAttribute(fast kernel) failed");
    EVR_FAST_VARIANTS(X)
#undef X
    return 0;
}

int fast_launch(int mm, bool rt, bool tri, int nctas, int nthr, size_t smem, cudaStream_t st,
                const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
    if (tri) rt = false;
    if (mm == 0) return fast0_launch(rt, tri, nctas, nthr, smem, st, P, C, npsi, psi, Hpsi);
#define X(m_, r_, t_) \
    if (mm == m_ && rt == r_ && tri == t_) { sg4_term_kernel_fast<m_, r_, t_><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi); return 0; }
    EVR_FAST_VARIANTS(X)
#undef X
    return fail("evr_sg4: no such fast-kernel variant");
}

int fast_permute(bool in, const int32_t *perm, long long nb, int nvecs, const double *src, double *dst, cudaStream_t st)
{
    const int thr = 256;
    long long blocks = (nb + thr - 1) / thr;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    if (in) sg4_permute_in<<<(int)blocks, thr, 0, st>>>(perm, nb, nvecs, src, dst);
    else sg4_permute_out<<<(int)blocks, thr, 0, st>>>(perm, nb, nvecs, src, dst);
    return 0;
}

} // namespace evr
