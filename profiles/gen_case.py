"""One curvilinear (generic-kernel) case for ncu: python profiles/gen_case.py [hno3|hcn] [npsi]"""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch
import evr_sg4_b200 as evr
from helpers import random_psi
case = sys.argv[1] if len(sys.argv) > 1 else "hno3"
npsi = int(sys.argv[2]) if len(sys.argv) > 2 else 1
b = evr.workloads.hm_sg4_basis(8, 6, 7, 1, 1) if case == "hno3" else evr.workloads.hm_sg4_basis(3, 6, 7, [10, 1, 1], [10, 2, 2])
op = evr.workloads.synthetic_curvilinear(b)
psi = random_psi(b.nb * b.nb0, npsi)
d_psi = torch.from_numpy(psi).cuda(); d_out = torch.empty_like(d_psi)
st = torch.cuda.current_stream().cuda_stream
for _ in range(3):
    op.apply_device_ptr(npsi, d_psi.data_ptr(), d_out.data_ptr(), st)
torch.cuda.synchronize()
