// sg4_fast_inst.cu -- launch dispatch of the separable-KEO term kernel (sg4_fast.cuh) over its instantiations, which live
// in one translation unit each (sg4_fast_k{1,0}{p,r,t}.cu: pool in shared / global memory x templated / runtime-size / cube
// tiles), and the permutation kernels of the block-ordered packed vector.
#include <cuda_runtime.h>
#include "sg4_fast_types.h"

namespace evr {

int fast_attr_1p();
int fast_launch_1p(int, int, size_t, cudaStream_t, const FastPlanDev &, const FastClassDev &, int, const double *, double *);
int fast_attr_1r();
int fast_launch_1r(int, int, size_t, cudaStream_t, const FastPlanDev &, const FastClassDev &, int, const double *, double *);
int fast_attr_1t();
int fast_launch_1t(int, int, size_t, cudaStream_t, const FastPlanDev &, const FastClassDev &, int, const double *, double *);
int fast_attr_0p();
int fast_launch_0p(int, int, size_t, cudaStream_t, const FastPlanDev &, const FastClassDev &, int, const double *, double *);
int fast_attr_0r();
int fast_launch_0r(int, int, size_t, cudaStream_t, const FastPlanDev &, const FastClassDev &, int, const double *, double *);
int fast_attr_0t();
int fast_launch_0t(int, int, size_t, cudaStream_t, const FastPlanDev &, const FastClassDev &, int, const double *, double *);

// packed vectors between the caller's order (RvecB) and the internal block order
static __global__ void sg4_permute_in(const int32_t *__restrict__ perm, const long long nb, const int nvecs,
                               const double *__restrict__ src, double *__restrict__ dst)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nb; i += (long long)gridDim.x * blockDim.x) {
        const int r = __ldg(perm + i);
        for (int v = 0; v < nvecs; ++v) dst[v * nb + i] = __ldg(src + v * nb + r);
    }
}
static __global__ void sg4_permute_out(const int32_t *__restrict__ perm, const long long nb, const int nvecs,
                                const double *__restrict__ src, double *__restrict__ dst)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nb; i += (long long)gridDim.x * blockDim.x) {
        const int r = __ldg(perm + i);
        for (int v = 0; v < nvecs; ++v) dst[v * nb + r] = src[v * nb + i];
    }
}

int fast_set_attributes()
{
    return fast_attr_1p() || fast_attr_1r() || fast_attr_1t() || fast_attr_0p() || fast_attr_0r() || fast_attr_0t();
}

int fast_launch(int mm, bool rt, bool tri, int nctas, int nthr, size_t smem, cudaStream_t st,
                const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
    if (tri) rt = false;
    if (mm) {
        if (tri) return fast_launch_1t(nctas, nthr, smem, st, P, C, npsi, psi, Hpsi);
        if (rt) return fast_launch_1r(nctas, nthr, smem, st, P, C, npsi, psi, Hpsi);
        return fast_launch_1p(nctas, nthr, smem, st, P, C, npsi, psi, Hpsi);
    }
    if (tri) return fast_launch_0t(nctas, nthr, smem, st, P, C, npsi, psi, Hpsi);
    if (rt) return fast_launch_0r(nctas, nthr, smem, st, P, C, npsi, psi, Hpsi);
    return fast_launch_0p(nctas, nthr, smem, st, P, C, npsi, psi, Hpsi);
}

// default (EVR_SG4_PERMUTE=0 restores the kernels above + a separate zero-fill): the zero-fill of the result rides on the
// permute-in kernel, and the read-out gathers through the inverse permutation so that its stores are the coalesced side
static __global__ void sg4_permute_in_zero(const int32_t *__restrict__ perm, const long long nb, const int nvecs,
                                    const double *__restrict__ src, double *__restrict__ dst, double *__restrict__ zero)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nb; i += (long long)gridDim.x * blockDim.x) {
        const int r = __ldg(perm + i);
        for (int v = 0; v < nvecs; ++v) { dst[v * nb + i] = __ldg(src + v * nb + r); zero[v * nb + i] = 0.0; }
    }
}
static __global__ void sg4_permute_out_inv(const int32_t *__restrict__ inv_perm, const long long nb, const int nvecs,
                                    const double *__restrict__ src, double *__restrict__ dst)
{
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < nb; j += (long long)gridDim.x * blockDim.x) {
        const int i = __ldg(inv_perm + j);
        for (int v = 0; v < nvecs; ++v) dst[v * nb + j] = src[v * nb + i];
    }
}
int fast_permute_x(bool in, const int32_t *perm_or_inv, long long nb, int nvecs, const double *src, double *dst, double *zero, cudaStream_t st)
{
    const int thr = 256;
    long long blocks = (nb + thr - 1) / thr;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    if (in) sg4_permute_in_zero<<<(int)blocks, thr, 0, st>>>(perm_or_inv, nb, nvecs, src, dst, zero);
    else sg4_permute_out_inv<<<(int)blocks, thr, 0, st>>>(perm_or_inv, nb, nvecs, src, dst);
    return 0;
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
