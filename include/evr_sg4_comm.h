/* evr_sg4_comm.h -- C-ABI of the multi-GPU reduction of H|psi> (one node, NVLink peer memory).
 *
 * Replaces the collective of the reference's MPI scheme 1: MPI_Reduce_sum_Bcast over size_RvecB*size_psi doubles in
 * Action_MPI_S1 (Source_ElVibRot/sub_Operator/sub_OpPsi_SG4_MPI.f90:535-560).  Each rank keeps its partial H|psi>
 * (evr_sg4_apply_device over its term range, include/evr_sg4.h) in a buffer that every other rank of the node has
 * mapped (CUDA IPC / VMM "symmetric memory"; the Python host obtains it from torch.distributed._symmetric_memory).
 */
#ifndef EVR_SG4_COMM_H
#define EVR_SG4_COMM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define EVR_SG4_MAX_PEERS 16

/* Rank `rank` of `np` sums slice `rank` of the n-double vectors behind peer_ptrs[0..np-1] (host array of DEVICE
 * pointers, as mapped in THIS process; peer_ptrs[rank] is the local buffer; all 16-byte aligned) in the fixed order
 * 0..np-1 and stores the sum into that slice of all np buffers.  Asynchronous on `cuda_stream`.  When every rank has
 * made this call between two cross-rank barriers, all buffers hold the same full sum, bit-identical on all ranks.
 * Returns 0, or non-zero with a message in evr_sg4_last_error(). */
int evr_sg4_allreduce_slices(const void *const *peer_ptrs, int np, int rank, int64_t n, void *cuda_stream);

/* The same all-reduce with the two cross-rank barriers INSIDE the kernel (no separate barrier launches): flag_ptrs[r] is
 * rank r's flag array -- EVR_SG4_FLAG_WORDS 64-bit words of peer-mapped device memory, zero before the first call, used
 * by nothing else -- and `call` is the number of this collective call on these buffers: 1, 2, 3, ... , the same on every
 * rank.  Every rank must make the call (also with an empty slice).  When the kernel has ended on a rank, its buffer
 * holds the full sum and no peer accesses it any more.  Replaces MPI_Reduce_sum_Bcast of Action_MPI_S1
 * (sub_Operator/sub_OpPsi_SG4_MPI.f90:553-557). */
#define EVR_SG4_FLAG_WORDS (2 * EVR_SG4_MAX_PEERS + 2)
int evr_sg4_allreduce_fused(const void *const *peer_ptrs, const void *const *flag_ptrs, int np, int rank, int64_t n,
                            uint64_t call, void *cuda_stream);

/* The two halves separately, for callers whose input and output live in HOST memory (evr_sg4_apply with
 * evr_sg4_set_devices, bench.py e2e at N > 1): every rank copies only its slice of psi host -> device, the slices are
 * all-gathered over NVLink, and after the term kernels every rank reduces and returns only its slice of H psi, so that N
 * PCIe links carry 1/N of the vector each.  Same slicing as evr_sg4_allreduce_slices: evr_sg4_slice_bounds gives the
 * [lo, hi) range of doubles owned by `rank` (lo is even, i.e. 16-byte aligned; an odd last element goes to the last rank).
 *   allgather_slices : every slice of peer_ptrs[rank] except its own <- the owner's buffer
 *   reduce_slice     : slice `rank` of peer_ptrs[rank] <- sum over the np buffers (fixed order 0..np-1)
 *   reduce_to        : dst[0..n) <- sum over the np buffers (device-resident caller on one device)
 * All asynchronous on `cuda_stream`; the caller orders them against the other ranks' work (events / barriers). */
int evr_sg4_slice_bounds(int64_t n, int np, int rank, int64_t *lo, int64_t *hi);
int evr_sg4_allgather_slices(const void *const *peer_ptrs, int np, int rank, int64_t n, void *cuda_stream);
int evr_sg4_reduce_slice(const void *const *peer_ptrs, int np, int rank, int64_t n, void *cuda_stream);
int evr_sg4_reduce_to(const void *const *peer_ptrs, int np, int64_t n, double *dst, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
