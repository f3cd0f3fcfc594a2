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

#ifdef __cplusplus
}
#endif
#endif
