"""N>1 host logic on CPU: two gloo ranks, MPI scheme-1 decomposition (term ranges + all-reduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import oracle_apply


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import evr_sg4_b200 as evr
        basis, op_full = evr.workloads.henon_heiles(6, 3)
        lo, hi = evr.distributed.ini_iGs(basis.nb_SG, world, rank)
        op_rank = evr.ParamOp(basis, 1, op_full.OpGrid, iG_range=(lo, hi))

        def local_apply(psi, out):          # stand-in for the CUDA apply: this rank's terms through the oracle
            out.copy_(torch.from_numpy(oracle_apply(op_rank, psi.numpy(), iG_range=(lo, hi))))

        tp = evr.distributed.TermParallelOp(op_rank, local_apply=local_apply)
        psi = torch.from_numpy(np.random.default_rng(5).standard_normal((2, basis.nb)))
        out = torch.empty_like(psi)
        tp.apply(psi, out)
        ref = oracle_apply(op_full, psi.numpy())
        err = float(np.abs(out.numpy() - ref).max() / np.abs(ref).max())
        # host buffers are never peer-mapped: the collective must be torch.distributed's, and all_reduce() alone sums in place
        assert tp.collective == "torch.distributed.all_reduce" and not tp._symm
        ones = torch.full((5,), float(rank + 1), dtype=torch.float64)
        assert torch.equal(tp.all_reduce(ones), torch.full((5,), 3.0, dtype=torch.float64))
        q.put((rank, lo, hi, err))
    finally:
        dist.destroy_process_group()


def test_two_rank_term_partition_allreduce():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == 84      # contiguous ranges covering all terms
    for _, _, _, err in res:
        assert err < 1e-13
