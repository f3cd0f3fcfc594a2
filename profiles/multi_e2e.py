#!/usr/bin/env python
"""End-to-end H|psi>/s through the C-ABI host entry (evr_sg4_apply, registered host buffers) for plans that span
1, 2, ... GPUs of ONE process (evr_sg4_set_devices).   usage: python profiles/multi_e2e.py 1 2 4 8 [--L 7]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import evr_sg4_b200 as evr
from helpers import random_psi, rel_l2
L = 7
args = [a for a in sys.argv[1:]]
if "--L" in args:
    L = int(args[args.index("--L") + 1]); del args[args.index("--L"):args.index("--L") + 2]
basis = evr.workloads.hm_sg4_basis(12, L, L, 1, 2)
V = evr.workloads.henon_heiles_potential(basis)
ops = evr.workloads.constant_keo_opgrids(12, 1, np.ones(12), V.reshape(-1, 1, 1))
psi = np.ascontiguousarray(random_psi(basis.nb, 1)); out = np.empty_like(psi)
lib = evr.lib.lib()
evr.lib.check(lib.evr_sg4_host_register(psi.ctypes.data, psi.nbytes)); evr.lib.check(lib.evr_sg4_host_register(out.ctypes.data, out.nbytes))
first = None
for n in [int(a) for a in args]:
    evr.lib.set_devices(n)
    t0 = time.perf_counter()
    op = evr.ParamOp(basis, 1, ops)
    op.plan()
    ts = time.perf_counter() - t0
    for _ in range(3): op.apply_host(psi, out=out)
    K = 30
    t0 = time.perf_counter()
    for _ in range(K): op.apply_host(psi, out=out)
    dt = (time.perf_counter() - t0) / K
    if first is None: first = out.copy()
    print(f"set_devices({n}): e2e {1/dt:.1f} Hpsi/s ({dt*1e3:.3f} ms per call, H2D+D2H {2*psi.nbytes/1e6:.1f} MB, setup {ts:.1f} s), "
          f"rel diff vs first = {rel_l2(out, first):.2e}", flush=True)
    op.close()
evr.lib.set_devices(1)
