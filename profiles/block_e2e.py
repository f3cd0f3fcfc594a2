"""End-to-end time of a BLOCK of vectors from page-locked host buffers (the sub_TabOpPsi call of a Davidson block):
evr_sg4_apply with the per-vector pipeline (copy-in of vector v+1 | kernels of vector v | copy-out of vector v-1) against
the plain sequence (EVR_SG4_PIPELINE=0: one copy-in, one launch over the block, one copy-out).  One setting per process."""
import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import evr_sg4_b200 as evr
from helpers import oracle_apply, random_psi, rel_l2

L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
npsi = int(sys.argv[2]) if len(sys.argv) > 2 else 8
basis = evr.workloads.hm_sg4_basis(12, L, L, 1, 2)
V = evr.workloads.model_potential_device(basis, 1, [evr.workloads.LAMBDA_HH])
op = evr.ParamOp(basis, 1, evr.workloads.constant_keo_opgrids(12, 1, np.ones(12), V.reshape(-1, 1, 1)))
psi = torch.from_numpy(random_psi(basis.nb, npsi, 7)).pin_memory()
out = torch.empty_like(psi).pin_memory()
x, y = psi.numpy(), out.numpy()
for _ in range(3): op.apply_host(x, out=y)
n = 10
t0 = time.perf_counter()
for _ in range(n): op.apply_host(x, out=y)
ms = (time.perf_counter() - t0) / n * 1e3
ref = oracle_apply(op, x[:1], nthreads=os.cpu_count() or 16)
print(f"HH 12-D L={L} block of {npsi} from pinned host buffers, EVR_SG4_PIPELINE={os.environ.get('EVR_SG4_PIPELINE', '1')}: "
      f"{ms:.3f} ms per block = {ms/npsi:.3f} ms per vector ({npsi/ms*1e3:.0f} H|psi>/s end to end), rel-L2 of vector 0 {rel_l2(y[0], ref[0]):.1e}")
