// sg4_allreduce.cu -- sum of the per-rank partial H|psi> vectors over NVLink peer memory (one node).
//
// Replaces MPI_Reduce_sum_Bcast of the reference's MPI scheme 1 (Action_MPI_S1, sub_Operator/sub_OpPsi_SG4_MPI.f90:
// 535-560) for vectors that live in peer-mapped ("symmetric") buffers: rank r sums slice r of all np buffers in the
// fixed order 0..np-1 and stores the sum into slice r of all np buffers (reduce-scatter and all-gather in ONE pass,
// every element crosses NVLink once in and once out).  All ranks end up with bit-identical vectors.  The caller
// brackets the launch with two cross-rank barriers on the stream (all partial sums complete / all slices written).
#include <algorithm>
#include <cstdlib>
#include <string>
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

// ---- the same all-reduce with BOTH cross-rank barriers inside the kernel ------------------------------------------
// Every rank owns a flag array of 2 * EVR_SG4_MAX_PEERS 64-bit words in peer-mapped memory: in[r] and out[r] are written
// by rank r only, with the (monotonically increasing) number of the call.
//   entry : one thread per peer stores the call number into in[rank] of that peer (system-scope release after a system
//           fence: this rank's partial sum was produced by earlier kernels of the same stream); every CTA then waits until
//           its own in[0..np) carry the call number (acquire) -- all partial sums are complete;
//   body  : slice `rank` summed over the np buffers, stored into all np buffers;
//   exit  : the last CTA to finish (device-scope counter) signals out[rank] on every peer and waits for its own
//           out[0..np): when the kernel ends, every slice has arrived in this rank's buffer and no peer reads it any more.
// The waits depend only on signals that the peers send at the START of their own kernel (entry) or after their own body
// (exit): no CTA of this grid waits for another CTA of this grid, so residency of the whole grid is not required.
struct FlagPtrs { unsigned long long *p[EVR_SG4_MAX_PEERS]; };

__device__ __forceinline__ void st_release_sys(unsigned long long *a, unsigned long long v)
{ asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a), "l"(v) : "memory"); }
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *a)
{ unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(a) : "memory"); return v; }
// wait until *a >= call; a peer that never arrives (a rank died, mismatched call numbers) ends in a trap after 20 s instead
// of a hung device
__device__ __forceinline__ void wait_flag(const unsigned long long *a, const unsigned long long call)
{
    unsigned long long t0 = 0;
    unsigned int spins = 0;
    while (ld_acquire_sys(a) < call) {
        if ((++spins & 0xFFFu) == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > 20000000000ull) __trap();
        }
    }
}

template <int NP>
__global__ void __launch_bounds__(256)
sg4_allreduce_fused_kernel(const PeerPtrs P, const FlagPtrs F, const int np_rt, const int rank, const unsigned long long call,
                           unsigned int *done_ctas, const long long lo2, const long long hi2, const long long tail)
{
    const int np = (NP > 0) ? NP : np_rt;
    if (blockIdx.x == 0 && threadIdx.x < np) {
        __threadfence_system();
        st_release_sys(F.p[threadIdx.x] + rank, call);
    }
    if (threadIdx.x < np)
        wait_flag(F.p[rank] + threadIdx.x, call);
    __syncthreads();
    for (long long i = lo2 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi2; i += (long long)gridDim.x * blockDim.x) {
        double2 v[(NP > 0) ? NP : EVR_SG4_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < ((NP > 0) ? NP : EVR_SG4_MAX_PEERS); ++r)
            if (r < np) v[r] = __ldcg(reinterpret_cast<const double2 *>(P.p[r]) + i);
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
    // exit barrier
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = (atomicAdd(done_ctas, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *done_ctas = 0;                    // next call (stream-ordered)
    if (threadIdx.x < np) {
        __threadfence_system();
        st_release_sys(F.p[threadIdx.x] + EVR_SG4_MAX_PEERS + rank, call);
        wait_flag(F.p[rank] + EVR_SG4_MAX_PEERS + threadIdx.x, call);
    }
}

// slice `rank` of the local buffer <- sum over the np buffers (reduce-scatter half; fixed order 0..np-1)
template <int NP>
__global__ void __launch_bounds__(256)
sg4_reduce_slice_kernel(const PeerPtrs P, const int np_rt, const int rank, double *__restrict__ dst,
                        const long long lo2, const long long hi2, const long long tail)
{
    const int np = (NP > 0) ? NP : np_rt;
    for (long long i = lo2 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi2; i += (long long)gridDim.x * blockDim.x) {
        double2 v[(NP > 0) ? NP : EVR_SG4_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < ((NP > 0) ? NP : EVR_SG4_MAX_PEERS); ++r)
            if (r < np) v[r] = __ldcg(reinterpret_cast<const double2 *>(P.p[r]) + i);
        double2 s = v[0];
#pragma unroll
        for (int r = 1; r < ((NP > 0) ? NP : EVR_SG4_MAX_PEERS); ++r)
            if (r < np) { s.x += v[r].x; s.y += v[r].y; }
        reinterpret_cast<double2 *>(dst)[i] = s;
    }
    if (tail >= 0 && blockIdx.x == 0 && threadIdx.x == 0) {
        double s = __ldcg(P.p[0] + tail);
        for (int r = 1; r < np; ++r) s += __ldcg(P.p[r] + tail);
        dst[tail] = s;
    }
}

// every slice but the local one <- the owner's copy (all-gather half; psi before the term kernels)
__global__ void __launch_bounds__(256)
sg4_allgather_kernel(const PeerPtrs P, const int np, const int rank, const long long n2, const long long chunk, const long long tail)
{
    double2 *dst = reinterpret_cast<double2 *>(P.p[rank]);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x) {
        const int owner = (int)(i / chunk);
        if (owner != rank) dst[i] = __ldcg(reinterpret_cast<const double2 *>(P.p[owner]) + i);
    }
    if (tail >= 0 && rank != np - 1 && blockIdx.x == 0 && threadIdx.x == 0) P.p[rank][tail] = __ldcg(P.p[np - 1] + tail);
}

} // namespace evr

namespace {
struct Slices { long long n2, chunk, lo2, hi2, tail; };
Slices slices_of(int64_t n, int np, int rank)
{
    Slices S;
    S.n2 = n / 2;                                               // double2 units
    S.chunk = std::max<long long>(1, (S.n2 + np - 1) / np);
    S.lo2 = std::min<long long>(S.n2, S.chunk * rank);
    S.hi2 = std::min<long long>(S.n2, S.lo2 + S.chunk);
    S.tail = ((n & 1) && rank == np - 1) ? n - 1 : -1;          // a last odd element belongs to the last rank
    return S;
}
int peers_of(evr::PeerPtrs &P, const void *const *peer_ptrs, int np, const char *who)
{
    for (int r = 0; r < EVR_SG4_MAX_PEERS; ++r) {
        P.p[r] = (r < np) ? static_cast<double *>(const_cast<void *>(peer_ptrs[r])) : nullptr;
        if (r < np && (!P.p[r] || (reinterpret_cast<uintptr_t>(P.p[r]) & 15)))
            return evr::fail(std::string(who) + ": peer buffers must be non-null and 16-byte aligned");
    }
    return 0;
}
int grid_for(long long work)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // measured at N = 2 (11 MB vector, profiles/r2/allreduce_in_kernel_barriers.txt): 1 / 2 / 4 / 8 / 16 CTAs per SM give
    // 39 / 37 / 38 / 38 / 47 us per all-reduce
    static const int per_sm = getenv("EVR_SG4_AR_CTAS_PER_SM") ? std::max(1, atoi(getenv("EVR_SG4_AR_CTAS_PER_SM"))) : 8;
    return (int)std::min<long long>(std::max<long long>(1, (work + 255) / 256), (long long)sms * per_sm);
}
} // namespace

extern "C" int evr_sg4_slice_bounds(int64_t n, int np, int rank, int64_t *lo, int64_t *hi)
{
    if (np < 1 || rank < 0 || rank >= np || n < 0 || !lo || !hi) return evr::fail("evr_sg4_slice_bounds: bad arguments");
    const Slices S = slices_of(n, np, rank);
    *lo = 2 * S.lo2;
    *hi = (rank == np - 1) ? n : 2 * S.hi2;
    return 0;
}

extern "C" int evr_sg4_allgather_slices(const void *const *peer_ptrs, int np, int rank, int64_t n, void *cuda_stream)
{
    using namespace evr;
    if (np < 1 || np > EVR_SG4_MAX_PEERS || rank < 0 || rank >= np || n < 0 || !peer_ptrs)
        return fail("evr_sg4_allgather_slices: bad arguments");
    PeerPtrs P;
    if (peers_of(P, peer_ptrs, np, "evr_sg4_allgather_slices")) return 1;
    if (n == 0 || np == 1) return 0;
    const Slices S = slices_of(n, np, rank);
    const long long tail = (n & 1) ? n - 1 : -1;
    sg4_allgather_kernel<<<grid_for(S.n2), 256, 0, static_cast<cudaStream_t>(cuda_stream)>>>(P, np, rank, S.n2, S.chunk, tail);
    if (cudaGetLastError() != cudaSuccess) return fail("evr_sg4_allgather_slices: kernel launch failed");
    return 0;
}

static int reduce_launch(const evr::PeerPtrs &P, int np, int rank, double *dst, long long lo2, long long hi2, long long tail, cudaStream_t st)
{
    using namespace evr;
    if (hi2 <= lo2 && tail < 0) return 0;
    const int grid = grid_for(hi2 - lo2);
    switch (np) {
    case 2: sg4_reduce_slice_kernel<2><<<grid, 256, 0, st>>>(P, np, rank, dst, lo2, hi2, tail); break;
    case 4: sg4_reduce_slice_kernel<4><<<grid, 256, 0, st>>>(P, np, rank, dst, lo2, hi2, tail); break;
    case 8: sg4_reduce_slice_kernel<8><<<grid, 256, 0, st>>>(P, np, rank, dst, lo2, hi2, tail); break;
    default: sg4_reduce_slice_kernel<0><<<grid, 256, 0, st>>>(P, np, rank, dst, lo2, hi2, tail); break;
    }
    if (cudaGetLastError() != cudaSuccess) return fail("evr_sg4_reduce: kernel launch failed");
    return 0;
}

extern "C" int evr_sg4_reduce_slice(const void *const *peer_ptrs, int np, int rank, int64_t n, void *cuda_stream)
{
    using namespace evr;
    if (np < 1 || np > EVR_SG4_MAX_PEERS || rank < 0 || rank >= np || n < 0 || !peer_ptrs)
        return fail("evr_sg4_reduce_slice: bad arguments");
    PeerPtrs P;
    if (peers_of(P, peer_ptrs, np, "evr_sg4_reduce_slice")) return 1;
    if (n == 0 || np == 1) return 0;
    const Slices S = slices_of(n, np, rank);
    return reduce_launch(P, np, rank, P.p[rank], S.lo2, S.hi2, S.tail, static_cast<cudaStream_t>(cuda_stream));
}

extern "C" int evr_sg4_reduce_to(const void *const *peer_ptrs, int np, int64_t n, double *dst, void *cuda_stream)
{
    using namespace evr;
    if (np < 1 || np > EVR_SG4_MAX_PEERS || n < 0 || !peer_ptrs || !dst || (reinterpret_cast<uintptr_t>(dst) & 15))
        return fail("evr_sg4_reduce_to: bad arguments");
    PeerPtrs P;
    if (peers_of(P, peer_ptrs, np, "evr_sg4_reduce_to")) return 1;
    if (n == 0) return 0;
    return reduce_launch(P, np, 0, dst, 0, n / 2, (n & 1) ? n - 1 : -1, static_cast<cudaStream_t>(cuda_stream));
}

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

extern "C" int evr_sg4_allreduce_fused(const void *const *peer_ptrs, const void *const *flag_ptrs, int np, int rank, int64_t n,
                                       uint64_t call, void *cuda_stream)
{
    using namespace evr;
    if (np < 1 || np > EVR_SG4_MAX_PEERS || rank < 0 || rank >= np || n < 0 || !peer_ptrs || !flag_ptrs || call == 0)
        return fail("evr_sg4_allreduce_fused: bad arguments");
    PeerPtrs P;
    if (peers_of(P, peer_ptrs, np, "evr_sg4_allreduce_fused")) return 1;
    FlagPtrs F;
    for (int r = 0; r < EVR_SG4_MAX_PEERS; ++r) {
        F.p[r] = (r < np) ? static_cast<unsigned long long *>(const_cast<void *>(flag_ptrs[r])) : nullptr;
        if (r < np && (!F.p[r] || (reinterpret_cast<uintptr_t>(F.p[r]) & 7))) return fail("evr_sg4_allreduce_fused: bad flag pointer");
    }
    // every rank launches (even with an empty slice): the barriers are part of the kernel
    const Slices S = slices_of(n, np, rank);
    unsigned int *done = reinterpret_cast<unsigned int *>(F.p[rank] + 2 * EVR_SG4_MAX_PEERS);   // word 2*MAX_PEERS of the local flag array
    const int grid = grid_for(std::max<long long>(1, S.hi2 - S.lo2));
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    switch (np) {
    case 2: sg4_allreduce_fused_kernel<2><<<grid, 256, 0, st>>>(P, F, np, rank, call, done, S.lo2, S.hi2, S.tail); break;
    case 4: sg4_allreduce_fused_kernel<4><<<grid, 256, 0, st>>>(P, F, np, rank, call, done, S.lo2, S.hi2, S.tail); break;
    case 8: sg4_allreduce_fused_kernel<8><<<grid, 256, 0, st>>>(P, F, np, rank, call, done, S.lo2, S.hi2, S.tail); break;
    default: sg4_allreduce_fused_kernel<0><<<grid, 256, 0, st>>>(P, F, np, rank, call, done, S.lo2, S.hi2, S.tail); break;
    }
    if (cudaGetLastError() != cudaSuccess) return fail("evr_sg4_allreduce_fused: kernel launch failed");
    return 0;
}
