#!/usr/bin/env python
"""Small invocations of every kernel family for compute-sanitizer (profiles/sanitize.sh): fast path (iso and pool
instantiations, batched items, bulk-copy staging, named-barrier thread groups), second-generation kernel, generic kernel,
type_Op=10, nested-SG4 transforms, peer reduction kernels, driver algebra.  Each result is checked against the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import evr_sg4_b200 as evr
from helpers import oracle_apply, random_psi, rel_l2

which = sys.argv[1] if len(sys.argv) > 1 else "all"


def check(name, op, npsi):
    b = op.BasisnD
    psi = random_psi(b.nb * b.nb0, npsi)
    err = rel_l2(op.apply_host(psi), oracle_apply(op, psi))
    print(f"{name}: rel L2 vs oracle {err:.2e}", flush=True)
    assert err < 1e-12


if which in ("all", "fast"):
    check("HH 6-D L=3 (iso fast path)", evr.workloads.henon_heiles(6, 3)[1], 2)
    check("HH 12-D L=3 (iso fast path, batched items)", evr.workloads.henon_heiles(12, 3)[1], 1)
    check("pyrazine 12-D L=2 (pool fast path, nb0=2)", evr.workloads.pyrazine_12d(2)[1], 2)
    check("HH 5-D L=4 LB=2 (dropped functions)", evr.workloads.henon_heiles(5, 4, LB=2)[1], 1)
if which in ("all", "generic"):
    basis = evr.workloads.hm_sg4_basis(3, 4, 5, [10, 1, 1], [10, 2, 2])
    check("HCN shape LB4/LG5 (generic kernel)", evr.workloads.synthetic_curvilinear(basis), 2)
    basis = evr.workloads.hm_sg4_basis(8, 2, 4, 1, 1)
    check("HNO3 shape LB2/LG4 (generic kernel)", evr.workloads.synthetic_curvilinear(basis), 1)
if which in ("all", "nested"):
    basis = evr.workloads.hm_sg4_basis(8, 2, 4, 1, 1)
    tr = evr.SG4Transforms(basis)
    x = np.random.default_rng(0).standard_normal((5, basis.nb))
    g = tr.RvecB_TO_RvecG(x)
    tr.RvecG_TO_RvecB(g); tr.DerivOp_TO_RvecG(g, 1, 3)
    print("nested transforms ran", flush=True)
if which in ("all", "algebra"):
    import torch
    A = torch.randn(7, 5001, dtype=torch.float64).cuda()
    G = evr.algebra.gram(A, A)
    v = torch.randn(5001, dtype=torch.float64).cuda()
    evr.algebra.schmidt(A[:0], v)
    print("algebra ran", float(G[0, 0]), flush=True)
if which in ("all", "late"):
    # kernels added late in round 2: global-buffer class of the generic and nested kernels (terms beyond shared memory),
    # fast-path plan with a generic remainder (mode of 17 points), type_Op=10 with the triangular metric storage, the
    # all-reduce with in-kernel barriers (3 "ranks" on streams of one device)
    import ctypes as C
    import torch
    from helpers import oracle_apply10
    big = evr.workloads.hm_sg4_basis(3, 3, 3, 1, [19, 19, 19], nb0=2)
    check("20^3 x 2 channels (global work buffers, generic kernel)", evr.workloads.synthetic_curvilinear(big), 1)
    tr = evr.SG4Transforms(big)
    x = np.random.default_rng(0).standard_normal((1, big.nb * 2))
    g = tr.RvecB_TO_RvecG(x); tr.RvecG_TO_RvecB(g); tr.DerivOp_TO_RvecG(g, 1, 3)
    print("nested transforms with global work buffers ran", flush=True)
    basis, op = evr.workloads.henon_heiles(4, 8)
    assert op.info(evr.lib.INFO_PATH) == 1 and op.info(evr.lib.INFO_GENERIC_TERMS) > 0
    check("HH 4-D L=8 (fast path + generic remainder)", op, 2)
    b5 = evr.workloads.hm_sg4_basis(5, 2, 4, 1, 1)
    o10 = evr.workloads.synthetic_type10(b5)
    o10 = evr.ParamOp10(b5, np.asfortranarray(0.5 * (o10.GG + o10.GG.transpose(0, 2, 1))), o10.Jac, o10.sqRhoOVERJac, V=o10.V)
    psi = random_psi(b5.nb, 1)
    err = rel_l2(o10.apply_host(psi), oracle_apply10(o10, psi))
    print(f"type_Op=10, triangular metric storage: rel L2 vs oracle {err:.2e}", flush=True)
    assert err < 1e-12
    np_, n = 3, 1001
    bufs = [torch.randn(n, dtype=torch.float64, device="cuda") for _ in range(np_)]
    flags = [torch.zeros(evr.lib.FLAG_WORDS, dtype=torch.int64, device="cuda") for _ in range(np_)]
    streams = [torch.cuda.Stream() for _ in range(np_)]
    want = bufs[0] + bufs[1] + bufs[2]
    torch.cuda.synchronize()
    ptrs = (C.c_void_p * np_)(*[b.data_ptr() for b in bufs]); fptrs = (C.c_void_p * np_)(*[f.data_ptr() for f in flags])
    for r in range(np_):
        evr.lib.check(evr.lib.lib().evr_sg4_allreduce_fused(ptrs, fptrs, np_, r, n, 1, C.c_void_p(streams[r].cuda_stream)), "fused")
    torch.cuda.synchronize()
    assert all(torch.equal(b, want) for b in bufs)
    print("all-reduce with in-kernel barriers ran", flush=True)
