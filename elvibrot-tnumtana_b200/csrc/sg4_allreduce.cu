// sg4_allreduce.cu -- sum of the per-rank partial H|psi> vectors over NVLink peer memory (one node).
//
// Replaces MPI_Reduce_sum_Bcast of the reference's MPI scheme 1 (Action_MPI_S1, sub_Operator/sub_OpPsi_SG4_MPI.f90:
// 535-560) for vectors that live in peer-mapped ("symmetric") buffers: rank r sums slice r of all np buffers in the
// fixed order 0..np-1 and stores the sum into slice r of all np buffers (reduce-scatter and all-gather in ONE pass,
// every element crosses NVLink once in and once out).  All ranks end up with bit-identical vectors.  The caller
// brackets the launch with two cross-rank barriers on the stream (all partial sums complete / all slices written).
#include <algorithm>
#include <cuda_runtime.h>
#include "sg4_internal.h"
#include "../../include/evr_sg4_comm.h"

namespace evr {

struct PeerPtrs { double *p[EVR_SG4_MAX_PEERS]; };

template <int NP>
__global__ void __launch_bounds__(256)
sg4_allreduce_slice_kernel(const PeerPtrs P, const int np_rt, const long long lo2, const long long hi2,
                           const long long tail /* index of a last odd element owned by this rank, or -1 */)
{
    const int np = (NP > 0) ? NP : np_rt;
    for (long long i = lo2 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi2; i += (long long)gridDim.x * blockDim.x) {
        double2 v[(NP > 0) ? NP : EVR_SG4_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < ((NP > 0) ? NP : EVR_SG4_MAX_PEERS); ++r)
            if (r < np) v[r] = __ldcg(reinterpret_cast<const double2 *>(P.p[r]) + i);      // all peer loads in flight together
        double2 s = v[0];
#pragma unroll
        for (int r = 1; r < ((NP > 0) ? NP : EVR_SG4_MAX_PEERS); ++r)
            if (r < np) { s.x += v[r].x; s.y += v[r].y; }
#pragma unroll
        for (int r = 0; r < ((NP > 0) ? NP : EVR_SG4_MAX_PEERS); ++r)
            if (r < np) __stcg(reinterpret_cast<double2 *>(P.p[r]) + i, s);
    }
    if (tail >= 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = __ldcg(P.p[0] + tail);
        for (int r = 1; r < np; ++r) s += __ldcg(P.p[r] + tail);
        for (int r = 0; r < np; ++r) __stcg(P.p[r] + tail, s);
    }
}

} // namespace evr

extern "C" int evr_sg4_allreduce_slices(const void *const *peer_ptrs, int np, int rank, int64_t n, void *cuda_stream)
{
    using namespace evr;
    if (np < 1 || np > EVR_SG4_MAX_PEERS || rank < 0 || rank >= np || n < 0 || !peer_ptrs)
        return fail("evr_sg4_allreduce_slices: bad arguments");
    PeerPtrs P;
    for (int r = 0; r < EVR_SG4_MAX_PEERS; ++r) {
        P.p[r] = (r < np) ? static_cast<double *>(const_cast<void *>(peer_ptrs[r])) : nullptr;
        if (r < np && (!P.p[r] || (reinterpret_cast<uintptr_t>(P.p[r]) & 15)))
            return fail("evr_sg4_allreduce_slices: peer buffers must be non-null and 16-byte aligned");
    }
    if (n == 0) return 0;
    const long long n2 = n / 2;                                  // double2 units
    const long long chunk = (n2 + np - 1) / np;
    const long long lo2 = std::min<long long>(n2, chunk * rank), hi2 = std::min<long long>(n2, lo2 + chunk);
    const long long tail = ((n & 1) && rank == np - 1) ? n - 1 : -1;
    if (hi2 <= lo2 && tail < 0) return 0;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = std::max<long long>(1, (hi2 - lo2 + 255) / 256);
    const int grid = (int)std::min<long long>(want, (long long)sms * 8);
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    switch (np) {
    case 2: sg4_allreduce_slice_kernel<2><<<grid, 256, 0, st>>>(P, np, lo2, hi2, tail); break;
    case 4: sg4_allreduce_slice_kernel<4><<<grid, 256, 0, st>>>(P, np, lo2, hi2, tail); break;
    case 8: sg4_allreduce_slice_kernel<8><<<grid, 256, 0, st>>>(P, np, lo2, hi2, tail); break;
    default: sg4_allreduce_slice_kernel<0><<<grid, 256, 0, st>>>(P, np, lo2, hi2, tail); break;
    }
    if (cudaGetLastError() != cudaSuccess) return fail("evr_sg4_allreduce_slices: kernel launch failed");
    return 0;
}
