// sg4_v2_iso.cu -- constant-matrix ("iso") instantiation of the second-generation term kernel (sg4_fast2.cuh).
// Own translation unit (parallel build); it therefore owns its copy of the __constant__ matrix array, bound per plan
// like the one of sg4_iso.cu.
#include <cuda_runtime.h>
#include <mutex>
#include "sg4_fast2.cuh"

namespace evr {

int v2_iso_set_attributes()
{
    if (cudaFuncSetAttribute(sg4_term_kernel_v2<2, 768>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(sg4_term_kernel_v2<2, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return fail("evr_sg4: cudaFuncSetAttribute(v2 iso kernel) failed");
    return 0;
}

int v2_iso_bind(int device, int id, const double *blocks, cudaStream_t st)
{
    static int bound_id[64] = {0};
    static std::mutex mtx;
    std::lock_guard<std::mutex> lock(mtx);
    int &cur = bound_id[device & 63];
    if (cur == id) return 0;
    // another plan's matrices (or none) are loaded: replace them.  Rare (alternating plans with different bases on one
    // device), so simply drain the device first: kernels of the other plan may still be reading the array; and wait for
    // the copy, so that a launch of this plan on any other stream finds the array populated.
    if (cur != 0 && cudaDeviceSynchronize() != cudaSuccess) return fail("evr_sg4: cudaDeviceSynchronize failed");
    if (cudaMemcpyToSymbolAsync(c_iso, blocks, sizeof(double) * EVR_ISO_LEN, 0, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail("evr_sg4: cudaMemcpyToSymbolAsync(c_iso, v2) failed");
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail("evr_sg4: cudaStreamSynchronize failed");
    cur = id;
    return 0;
}

int v2_iso_launch(int nctas, int nthr, size_t smem, cudaStream_t st,
                  const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
    if (nthr > 512) sg4_term_kernel_v2<2, 768><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    else sg4_term_kernel_v2<2, 512><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    return 0;
}

} // namespace evr
