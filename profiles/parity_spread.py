#!/usr/bin/env python
"""Run-to-run spread of the parity of the benchmarked configuration (HH-12D L=7): rel L2 of repeated H|psi> against the
oracle and against each other (FP64 atomics: the summation order changes from call to call)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import evr_sg4_b200 as evr
from helpers import oracle_apply, random_psi, rel_l2
L = int(sys.argv[1]) if len(sys.argv) > 1 else 7
basis, op = evr.workloads.henon_heiles(12, L)
psi = random_psi(basis.nb, 1)
ref = oracle_apply(op, psi, nthreads=len(os.sched_getaffinity(0)))
outs = [op.apply_host(psi).copy() for _ in range(8)]
print(os.environ.get("EVR_SG4_ORDER", "interleaved"), os.environ.get("EVR_SG4_DETERMINISTIC", ""),
      "vs oracle:", " ".join(f"{rel_l2(o, ref):.2e}" for o in outs), "| run-to-run:", f"{max(rel_l2(o, outs[0]) for o in outs[1:]):.2e}")
