"""One-off size check beyond the benchmark: HH 12-D SG4 L=8 (125 970 terms, ~1.1e8 grid points, largest term 3^8 = 6561
points) -- parity against the oracle port and device-resident time.  Not part of the test suite (the oracle needs ~1-2 s on
all host cores, the Python set-up a few minutes)."""
import os, sys, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
import evr_sg4_b200 as evr
from helpers import oracle_apply, random_psi, rel_l2

L = int(sys.argv[1]) if len(sys.argv) > 1 else 8
t0 = time.time()
basis = evr.workloads.hm_sg4_basis(12, L, L, 1, 2)
print(f"basis: nb_SG={basis.nb_SG} NQ={basis.nqq} nb={basis.nb} max term {basis.tab_nq_OF_SRep.max()}  ({time.time()-t0:.0f} s)", flush=True)
V = evr.workloads.model_potential_device(basis, 1, [evr.workloads.LAMBDA_HH])
op = evr.ParamOp(basis, 1, evr.workloads.constant_keo_opgrids(12, 1, np.ones(12), V.reshape(-1, 1, 1)))
print(f"plan: path={op.info(evr.lib.INFO_PATH)} iso={op.info(evr.lib.INFO_ISO)}  ({time.time()-t0:.0f} s)", flush=True)
psi = random_psi(basis.nb, 1, 7)
d_psi = torch.from_numpy(psi).cuda(); d_out = torch.empty_like(d_psi)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3): op.apply_device_ptr(1, d_psi.data_ptr(), d_out.data_ptr(), st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 10; e0.record()
for _ in range(n): op.apply_device_ptr(1, d_psi.data_ptr(), d_out.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
alg = op.info(evr.lib.INFO_ALG_BYTES_NPSI1)
nthr = os.cpu_count() or 16
t1 = time.perf_counter(); ref = oracle_apply(op, psi, nthreads=nthr); tc = time.perf_counter() - t1
err = rel_l2(d_out.cpu().numpy(), ref)
print(f"HH 12-D L={L}: {ms:.3f} ms per H|psi> device-resident ({alg/ms/1e6:.0f} GB/s algorithmic), oracle port {tc:.2f} s on {nthr} threads, rel-L2 {err:.2e}")
assert err < 1e-12
