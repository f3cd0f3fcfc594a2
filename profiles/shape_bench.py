import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
import evr_sg4_b200 as evr
from helpers import oracle_apply, random_psi, rel_l2
def bench(name, basis, op, npsi):
    psi = random_psi(basis.nb*basis.nb0, npsi)
    d_psi = torch.from_numpy(psi).cuda(); d_out = torch.empty_like(d_psi)
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(3): op.apply_device_ptr(npsi, d_psi.data_ptr(), d_out.data_ptr(), st)
    torch.cuda.synchronize()
    e0,e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n=20; e0.record()
    for _ in range(n): op.apply_device_ptr(npsi, d_psi.data_ptr(), d_out.data_ptr(), st)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/n
    t0=time.perf_counter(); ref = oracle_apply(op, psi, nthreads=16); tc=time.perf_counter()-t0
    err = rel_l2(d_out.cpu().numpy(), ref)
    alg = op.info(evr.lib.INFO_ALG_BYTES_NPSI1) + (npsi-1)*op.info(evr.lib.INFO_ALG_BYTES_PER_RHS_EXTRA)
    print(f"{name}: nb_SG={basis.nb_SG} NQ={basis.nqq} nb={basis.nb} npsi={npsi} path={op.info(evr.lib.INFO_PATH)} gpu {ms*1e3:.1f} us/apply ({npsi/ms*1e3:.0f} Hpsi/s) cpu16 {tc*1e3:.1f} ms  relL2 {err:.1e}  alg {alg/1e6:.2f} MB -> {alg/ms/1e6:.1f} GB/s")
b = evr.workloads.hm_sg4_basis(3,6,7,[10,1,1],[10,2,2]); bench('HCN-shape curvilinear', b, evr.workloads.synthetic_curvilinear(b), 1)
bench('HCN-shape curvilinear', b, evr.workloads.synthetic_curvilinear(b), 27)
b = evr.workloads.hm_sg4_basis(8,3,5,1,1); bench('HNO3-shape LB3LG5', b, evr.workloads.synthetic_curvilinear(b), 1)
b = evr.workloads.hm_sg4_basis(8,6,7,1,1); bench('HNO3-shape LB6LG7', b, evr.workloads.synthetic_curvilinear(b), 1)
b, op = evr.workloads.pyrazine_12d(1); bench('pyrazine L1 nb0=2', b, op, 2)
b, op = evr.workloads.pyrazine_12d(3); bench('pyrazine L3 nb0=2', b, op, 2)
b, op = evr.workloads.henon_heiles(6,3); bench('HH6D L3', b, op, 28)
b, op = evr.workloads.henon_heiles(21,2); bench('HH21D L2', b, op, 22)
