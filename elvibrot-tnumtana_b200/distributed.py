"""Term-parallel H|psi> over several GPUs: one process per GPU, Smolyak terms split in contiguous
ranges, replicated packed psi, partial results summed with an all-reduce.

This is the reference's MPI "scheme 1" (Action_MPI_S1, sub_Operator/sub_OpPsi_SG4_MPI.f90:454-571):
rank r applies the terms iGs_MPI(1:2,r) (ini_iGs_MPI, sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4_MPI.f90:
639-669) into a zeroed vector, then MPI_Reduce_sum_Bcast (= all-reduce) over size_RvecB*size_psi doubles.
Here the collective is torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Optional, Tuple

from . import lib as _lib


def ini_iGs(nb_SG: int, world_size: int, rank: int) -> Tuple[int, int]:
    """0-based [begin, end) term range of ``rank`` (C-ABI evr_sg4_ini_iGs)."""
    b, e = C.c_int(), C.c_int()
    _lib.check(_lib.lib().evr_sg4_ini_iGs(nb_SG, world_size, rank, C.byref(b), C.byref(e)), "evr_sg4_ini_iGs")
    return b.value, e.value


def balanced_iGs(cost, world_size: int, rank: int) -> Tuple[int, int]:
    """0-based [begin, end) contiguous term range with ~1/world_size of sum(cost) (C-ABI evr_sg4_balanced_iGs)."""
    import numpy as np
    cost = np.ascontiguousarray(cost, dtype=np.int32)
    b, e = C.c_int(), C.c_int()
    _lib.check(_lib.lib().evr_sg4_balanced_iGs(len(cost), cost.ctypes.data, world_size, rank, C.byref(b), C.byref(e)),
               "evr_sg4_balanced_iGs")
    return b.value, e.value


class TermParallelOp:
    """H|psi> with the Smolyak terms of ``para_Op`` restricted to this rank's range + all-reduce.

    ``local_apply(psi_tensor, out_tensor)`` computes this rank's partial sum into ``out_tensor``
    (overwriting it).  By default it is the CUDA path of ``para_Op`` (device pointers, current stream);
    the CPU tests pass their own callable to exercise the partition/all-reduce logic under gloo.
    """

    def __init__(self, para_Op, group=None, local_apply: Optional[Callable] = None):
        import torch.distributed as dist
        self.para_Op = para_Op
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local = local_apply or self._cuda_apply

    def _cuda_apply(self, psi, out):
        import torch
        assert psi.is_cuda and out.is_cuda and psi.dtype == torch.float64 and psi.is_contiguous() and out.is_contiguous()
        npsi = 1 if psi.dim() == 1 else psi.shape[0]
        self.para_Op.apply_device_ptr(npsi, psi.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def apply(self, psi, out):
        """out <- sum over ranks of (H restricted to the rank's terms) psi ; psi replicated on every rank."""
        self._local(psi, out)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out
