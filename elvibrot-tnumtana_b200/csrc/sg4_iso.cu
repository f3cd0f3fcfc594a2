// sg4_iso.cu -- the constant-matrix ("iso") instantiation of the separable-KEO term kernel (sg4_fast.cuh).
//
// Selected by the plan when all active modes of equal size n share ONE [B | B^T w | T] block (e.g. the D identical
// Gauss-Hermite modes of the Henon-Heiles inputs, Working_tests/MPI_tests/*D_Davidson_openMP).  The blocks then sit at
// compile-time offsets of a __constant__ array, so that every matrix element is a constant-bank operand of the DFMA
// that uses it: no matrix loads, no matrix registers.  The freed registers pay for larger tiles (3x3x3, 5x5, 3x5, 3x7,
// 3x9 values per thread, 128 registers at <= 512 threads per CTA) and therefore fewer shared-memory sweeps per term.
#include <cuda_runtime.h>
#include <mutex>
#include "sg4_fast.cuh"

namespace evr {

int iso_set_attributes()
{
    if (cudaFuncSetAttribute(sg4_term_kernel_fast<2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(sg4_term_kernel_fast<2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return fail("evr_sg4: cudaFuncSetAttribute(iso kernel) failed");
    return 0;
}

int iso_bind(int device, int id, const double *blocks, cudaStream_t st)
{
    static int bound_id[64] = {0};
    static std::mutex mtx;                      // plans of several host threads (multi-device layer) bind concurrently
    std::lock_guard<std::mutex> lock(mtx);
    int &cur = bound_id[device & 63];
    if (cur == id) return 0;
    // another plan's matrices (or none) are loaded: replace them.  Rare (alternating plans with different bases on one
    // device), so simply drain the device first: kernels of the other plan may still be reading the array.
    if (cur != 0 && cudaDeviceSynchronize() != cudaSuccess) return fail("evr_sg4: cudaDeviceSynchronize failed");
    if (cudaMemcpyToSymbolAsync(c_iso, blocks, sizeof(double) * EVR_ISO_LEN, 0, cudaMemcpyHostToDevice, st) != cudaSuccess)
        return fail("evr_sg4: cudaMemcpyToSymbolAsync(c_iso) failed");
    // the array must be populated before a launch of this plan on ANY stream may find it "bound"
    if (cudaStreamSynchronize(st) != cudaSuccess) return fail("evr_sg4: cudaStreamSynchronize failed");
    cur = id;
    return 0;
}

int iso_launch(bool big_tiles, int nctas, int nthr, size_t smem, cudaStream_t st,
               const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi)
{
    if (big_tiles) sg4_term_kernel_fast<2, false, true><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    else sg4_term_kernel_fast<2, false, false><<<nctas, nthr, smem, st>>>(P, C, npsi, psi, Hpsi);
    return 0;
}

} // namespace evr
