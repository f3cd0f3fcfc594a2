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
