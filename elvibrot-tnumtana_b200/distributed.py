"""Term-parallel H|psi> over several GPUs: one process per GPU, Smolyak terms split in contiguous
ranges, replicated packed psi, partial results summed with an all-reduce.

This is the reference's MPI "scheme 1" (Action_MPI_S1, sub_Operator/sub_OpPsi_SG4_MPI.f90:454-571):
rank r applies the terms iGs_MPI(1:2,r) (ini_iGs_MPI, sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4_MPI.f90:
639-669) into a zeroed vector, then MPI_Reduce_sum_Bcast (= all-reduce) over size_RvecB*size_psi doubles.
Here the collective is, on GPUs of one node, the library's own peer-memory kernel (evr_sg4_allreduce_slices,
include/evr_sg4_comm.h: every rank sums one slice of all ranks' buffers over NVLink and stores it back into all of
them) on buffers from ``TermParallelOp.symmetric_empty``; for any other buffer it is torch.distributed's all_reduce
(NCCL on GPUs, gloo in the CPU tests).
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
        self._symm = {}            # data_ptr -> (tensor, symmetric-memory handle, ctypes array of peer pointers)
        self._flags = {}           # data_ptr -> [flag tensor, handle, peer pointers of the flags, calls made]
        self.collective = "torch.distributed.all_reduce"

    def symmetric_empty(self, *shape):
        """float64 CUDA tensor that every rank of the node has mapped (torch.distributed._symmetric_memory); ``apply`` /
        ``all_reduce`` on it use the peer-memory kernel.  Collective call.  Falls back to an ordinary tensor (and the
        NCCL all-reduce) when symmetric memory is unavailable on ANY rank, or with EVR_SG4_ALLREDUCE=nccl."""
        import os
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", torch.cuda.current_device())
        ok, t, hdl = 0, None, None
        want = self.world > 1 and os.environ.get("EVR_SG4_ALLREDUCE", "p2p") != "nccl" and self.world <= 16

        def agree(flag: int) -> int:        # MIN over the ranks: everybody takes the same branch
            if self.world == 1:
                return flag
            f = torch.tensor([flag], dtype=torch.int32, device=dev)
            dist.all_reduce(f, op=dist.ReduceOp.MIN, group=self.group)
            return int(f.item())

        # step 1 (local, cannot hang): import + allocation; agreed on BEFORE the collective rendezvous, so that a rank
        # that fails here does not leave the others waiting inside it
        symm_mem = None
        if want:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                t = symm_mem.empty(*shape, dtype=torch.float64, device=dev)
                ok = 1
            except Exception as e:      # noqa: BLE001 - any failure means "use NCCL", agreed on by all ranks below
                self._symm_error = repr(e)
                ok = 0
        ok = agree(ok) if self.world > 1 else 0
        # step 2 (collective): rendezvous, then agree on its outcome
        if ok:
            try:
                grp = self.group if self.group is not None else dist.group.WORLD
                hdl = symm_mem.rendezvous(t, grp)
                ok = 1 if (hdl.world_size == self.world and hdl.rank == self.rank and t.data_ptr() % 16 == 0
                           and int(hdl.buffer_ptrs[self.rank]) == t.data_ptr()) else 0
            except Exception as e:      # noqa: BLE001
                self._symm_error = repr(e)
                ok = 0
            ok = agree(ok)
        if not ok:
            return torch.empty(*shape, dtype=torch.float64, device=dev)
        ptrs = (C.c_void_p * self.world)(*[int(p) for p in hdl.buffer_ptrs])
        self._symm[t.data_ptr()] = (t, hdl, ptrs)
        self.collective = "evr_sg4_allreduce_slices (NVLink peer memory, reduce-scatter + all-gather in one kernel)"
        # EVR_SG4_ALLREDUCE=fused: the all-reduce with in-kernel barriers (evr_sg4_allreduce_fused; flag words in one more
        # peer-mapped buffer).  Measured equal to the kernel between two symmetric-memory barriers (37-38 us per 11 MB vector
        # on 2 B200, NCCL 46 us; profiles/r2/allreduce_in_kernel_barriers.txt), so the latter stays the default.
        if os.environ.get("EVR_SG4_ALLREDUCE", "p2p") == "fused":
            fok, ft, fh = 0, None, None
            try:
                ft = symm_mem.empty(_lib.FLAG_WORDS, dtype=torch.int64, device=dev)
                ft.zero_()
                fok = 1
            except Exception as e:      # noqa: BLE001
                self._symm_error = repr(e)
            if agree(fok):
                try:
                    fh = symm_mem.rendezvous(ft, self.group if self.group is not None else dist.group.WORLD)
                    fok = 1 if int(fh.buffer_ptrs[self.rank]) == ft.data_ptr() else 0
                except Exception as e:  # noqa: BLE001
                    self._symm_error = repr(e)
                    fok = 0
                torch.cuda.synchronize()
                if agree(fok):          # (also orders every rank's zero-fill before any peer's first signal)
                    fptrs = (C.c_void_p * self.world)(*[int(p) for p in fh.buffer_ptrs])
                    self._flags[t.data_ptr()] = [ft, fh, fptrs, 0]
                    self.collective = "evr_sg4_allreduce_fused (NVLink peer memory; both cross-rank barriers inside the kernel)"
        return t

    def all_reduce(self, out):
        """In-place sum of ``out`` over the ranks (MPI_Reduce_sum_Bcast of Action_MPI_S1)."""
        if self.world == 1:
            return out
        ent = self._symm.get(out.data_ptr()) if out.is_cuda else None
        if ent is None:
            import torch.distributed as dist
            dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
            return out
        import torch
        _, hdl, ptrs = ent
        fl = self._flags.get(out.data_ptr())
        if fl is not None:          # barriers inside the kernel
            fl[3] += 1
            _lib.check(_lib.lib().evr_sg4_allreduce_fused(ptrs, fl[2], self.world, self.rank, out.numel(), fl[3],
                                                          C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                       "evr_sg4_allreduce_fused")
            return out
        hdl.barrier(channel=0)      # every rank's partial sum is complete (stream-ordered, device-side)
        _lib.check(_lib.lib().evr_sg4_allreduce_slices(ptrs, self.world, self.rank, out.numel(),
                                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)),
                   "evr_sg4_allreduce_slices")
        hdl.barrier(channel=1)      # every slice has been written everywhere
        return out

    def apply_host_slices(self, psi_host, out_host, d_psi, d_out):
        """Host-resident caller at N > 1: this rank copies only ITS slice of the (replicated) host psi to the device, the
        slices are all-gathered over NVLink, every rank applies its terms, and after the reduce-scatter this rank returns
        only its slice of H psi to ``out_host`` (the other entries of ``out_host`` are left untouched) -- the N PCIe links
        carry 1/N of the vector each.  ``d_psi`` / ``d_out`` must come from ``symmetric_empty``.  Returns the [lo, hi) slice."""
        import torch
        ep, eo = self._symm.get(d_psi.data_ptr()), self._symm.get(d_out.data_ptr())
        if ep is None or eo is None:
            raise RuntimeError("apply_host_slices needs peer-mapped device buffers (TermParallelOp.symmetric_empty)")
        n = d_psi.numel()
        lo, hi = _lib.slice_bounds(n, self.world, self.rank)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L = _lib.lib()
        d_psi.view(-1)[lo:hi].copy_(psi_host.view(-1)[lo:hi], non_blocking=True)
        ep[1].barrier(channel=0)                                    # every rank's slice is on its device
        _lib.check(L.evr_sg4_allgather_slices(ep[2], self.world, self.rank, n, st), "evr_sg4_allgather_slices")
        self._local(d_psi, d_out)
        eo[1].barrier(channel=0)                                    # every rank's partial sum is complete
        _lib.check(L.evr_sg4_reduce_slice(eo[2], self.world, self.rank, n, st), "evr_sg4_reduce_slice")
        out_host.view(-1)[lo:hi].copy_(d_out.view(-1)[lo:hi], non_blocking=True)
        eo[1].barrier(channel=1)                                    # peers are done reading this rank's buffers
        return lo, hi

    def _cuda_apply(self, psi, out):
        import torch
        assert psi.is_cuda and out.is_cuda and psi.dtype == torch.float64 and psi.is_contiguous() and out.is_contiguous()
        npsi = 1 if psi.dim() == 1 else psi.shape[0]
        self.para_Op.apply_device_ptr(npsi, psi.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)

    def apply(self, psi, out):
        """out <- sum over ranks of (H restricted to the rank's terms) psi ; psi replicated on every rank."""
        self._local(psi, out)
        return self.all_reduce(out)
