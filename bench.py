#!/usr/bin/env python
"""bench.py -- H|psi> applications per second on the Smolyak SG4 grid (FP64), BASELINE.json's metric.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--L 7] [--npsi 1]

One "step" = one H|psi> (npsi right-hand sides, default 1) on the synthetic 12-D coupled
Henon-Heiles SG4 configuration (BASELINE.json configs[4]; L=7: 50 388 Smolyak terms, 23.8 M grid
points, packed basis 1 392 065).  N>1 (launched by torchrun, one rank per GPU): Smolyak terms are
partitioned across ranks (MPI scheme 1 of the reference, ini_iGs ranges), every rank applies its terms
to the replicated packed psi, and the partial results are summed with an NCCL all-reduce.

The JSON line carries: value (device-resident, CUDA events), e2e (host buffers through the C-ABI
evr_sg4_apply incl. H2D/D2H), roofline of the term kernel against the measured HBM peak, and the
CPU baseline (oracle port, all host threads) measured in the same run.
`--impl reference` times the CPU implementation only (the Fortran reference cannot be compiled in
this image -- no Fortran compiler -- so the arm is the oracle port, OpenMP over terms like PSG4_omp=1).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "H|psi> applications/sec (FP64, SG4)"
UNIT = "Hpsi/s"


def workload_name(D, L, npsi):
    return f"HenonHeiles-{D}D SG4 LB=LG={L} Hm(nq=nb=1+2L) type_Op=1 (V grid + {D} constant KEO terms), npsi={npsi}"


class ClockSampler:
    """nvidia-smi sampling during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([t.strip() for t in ln.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def host_threads():
    """Host threads for the CPU legs: every core this process may run on.  Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to its workers, which would time the reference arm on ONE core at N > 1."""
    if os.environ.get("EVR_CPU_THREADS"):
        return max(1, int(os.environ["EVR_CPU_THREADS"]))
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_time_hpsi(op, psi, nthreads, budget_s=25.0):
    """Time the oracle port on the full workload when it fits the budget, else on a bounded sample.
    Returns (seconds per full H|psi>, sample text, full oracle result or None)."""
    import numpy as np
    from helpers import oracle_apply
    b = op.BasisnD
    t0 = time.perf_counter()
    ref = oracle_apply(op, psi, nthreads=nthreads)          # also the parity reference of this run
    t_first = time.perf_counter() - t0
    reps = int(min(4, max(0, (budget_s - t_first) // max(t_first, 1e-3))))
    ts = [t_first]
    for _ in range(reps):
        t0 = time.perf_counter()
        oracle_apply(op, psi, nthreads=nthreads)
        ts.append(time.perf_counter() - t0)
    tm = sorted(ts)[len(ts) // 2]
    return tm, f"{len(ts)} full H|psi> (all {b.nb_SG} terms, {b.nqq} grid points), median", ref


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--D", type=int, default=12)
    ap.add_argument("--L", type=int, default=7)
    ap.add_argument("--npsi", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--partition", default="points", choices=["points", "count", "cost"],
                    help="multi-GPU term ranges: equal grid points (default) or equal term counts (reference ini_iGs_MPI)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import numpy as np

    if args.impl == "reference":
        if rank != 0:
            return 0
        import evr_sg4_b200 as evr
        from oracle import sg4_oracle
        from helpers import random_psi
        basis = evr.workloads.hm_sg4_basis(args.D, args.L, args.L, 1, 2)
        V = evr.workloads.henon_heiles_potential(basis)
        op = evr.ParamOp(basis, 1, evr.workloads.constant_keo_opgrids(args.D, 1, np.ones(args.D), V.reshape(-1, 1, 1)))
        psi = random_psi(basis.nb, args.npsi)
        nth = host_threads()
        from helpers import oracle_apply
        # one step = one full H|psi> of the workload (all terms): ~0.2-0.3 s on 16 host threads at L=7, so no sampling
        # (a sample of the first terms under-weights the large terms and biases the CPU arm low)
        per = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            oracle_apply(op, psi, nthreads=nth)
            if i >= args.warmup:
                per.append(time.perf_counter() - t0)
        sec = sum(per) / len(per)
        sample = f"each step = one full H|psi> (all {basis.nb_SG} terms, {basis.nqq} grid points); mean of {len(per)} steps"
        val = 1.0 / sec
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.D, args.L, args.npsi), "nb_SG": basis.nb_SG, "grid_points": basis.nqq,
                           "nb": basis.nb, "npsi": args.npsi},
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": nth, "kind": "port", "sample": sample},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "note": "Fortran reference not buildable here (no Fortran compiler); CPU arm = C/OpenMP port of the "
                        "reference algorithm (oracle/sg4_oracle.c), static schedule over Smolyak terms like PSG4_omp=1"}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import evr_sg4_b200 as evr
    from helpers import random_psi

    t_setup = time.perf_counter()
    basis = evr.workloads.hm_sg4_basis(args.D, args.L, args.L, 1, 2)
    V = evr.workloads.henon_heiles_potential(basis)
    ops = evr.workloads.constant_keo_opgrids(args.D, 1, np.ones(args.D), V.reshape(-1, 1, 1))
    if args.partition == "count":
        lo_, hi_ = evr.distributed.ini_iGs(basis.nb_SG, world, rank)                 # reference ini_iGs_MPI
    elif args.partition == "cost":
        # equal modelled kernel time per rank: a term costs (1 + 0.15 * active modes) per grid point (gather / scatter /
        # staging are per point, the transform passes grow with the number of modes of more than one point)
        nact = (basis.nDind_SmolyakRep_Tab_nDval > 0).sum(axis=1)
        cost = (basis.tab_nq_OF_SRep.astype(np.int64) * (20 + 3 * nact) // 20).astype(np.int32)
        lo_, hi_ = evr.distributed.balanced_iGs(cost, world, rank)
    else:
        lo_, hi_ = evr.distributed.balanced_iGs(basis.tab_nq_OF_SRep, world, rank)   # equal grid points per rank
    op = evr.ParamOp(basis, 1, ops, iG_range=(lo_, hi_), device=local_rank)
    op.plan()
    tp = evr.distributed.TermParallelOp(op)
    t_setup = time.perf_counter() - t_setup

    npsi, nvec = args.npsi, basis.nb * basis.nb0
    psi_h = torch.from_numpy(random_psi(nvec, npsi)).pin_memory()
    out_h = torch.empty_like(psi_h).pin_memory()
    if world > 1:                                   # peer-mapped buffers (NVLink collectives of the library)
        d_psi = tp.symmetric_empty(*psi_h.shape)
        d_psi.copy_(psi_h, non_blocking=True)
        d_out = tp.symmetric_empty(*psi_h.shape)
    else:
        d_psi = psi_h.cuda(non_blocking=True)
        d_out = torch.empty_like(d_psi)
    stream = torch.cuda.current_stream()

    def step():
        tp.apply(d_psi, d_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    allreduce_check = None
    if world > 1:       # the peer-memory reduction against NCCL on the same partial sums (rel. max difference, all ranks)
        op.apply_device_ptr(npsi, d_psi.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        ref = d_out.clone()
        dist.all_reduce(ref, op=dist.ReduceOp.SUM)
        tp.all_reduce(d_out)
        chk = ((d_out - ref).abs().max() / ref.abs().max()).reshape(1)
        dist.all_reduce(chk, op=dist.ReduceOp.MAX)
        allreduce_check = float(chk[0])
        assert allreduce_check < 1e-13, f"peer-memory all-reduce differs from NCCL: {allreduce_check:.3e}"
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = op.info(evr.lib.INFO_LAUNCHES)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    ev0.record(stream)
    for i in range(args.steps):
        kev[i][0].record(stream)
        op.apply_device_ptr(npsi, d_psi.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
        kev[i][1].record(stream)
        if world > 1:
            tp.all_reduce(d_out)
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    k_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    launches = op.info(evr.lib.INFO_LAUNCHES) - launches0
    t = torch.tensor([ms_total, k_ms], dtype=torch.float64, device="cuda")
    k_ms_ranks = [k_ms]
    if world > 1:
        every = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(every, t[1:2].clone())
        k_ms_ranks = [float(e[0]) for e in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, k_ms = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = 1e3 / ms_step

    # ---- end to end: host buffers through evr_sg4_apply (H2D + kernel + D2H [+ all-reduce of host result])
    e2e = None
    if not args.no_e2e:
        xh, yh = psi_h.numpy(), out_h.numpy()
        slice_io = world > 1 and bool(tp._symm) and d_psi.data_ptr() in tp._symm and d_out.data_ptr() in tp._symm

        def e2e_step():
            if world == 1:
                op.apply_host(xh, out=yh)                       # C-ABI evr_sg4_apply: H2D + kernels + D2H
            elif slice_io:                                      # every rank moves only its slice over its own PCIe link
                tp.apply_host_slices(psi_h, out_h, d_psi, d_out)
                torch.cuda.current_stream().synchronize()
            else:                                               # replicated psi H2D, term-parallel apply + all-reduce, D2H
                d_psi.copy_(psi_h, non_blocking=True)
                tp.apply(d_psi, d_out)
                out_h.copy_(d_out, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        te = torch.tensor([(time.perf_counter() - t0) / args.steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # slice-wise I/O: the job as a whole moves the vector once in and once out (1/N per rank and PCIe link); the
        # replicated fallback moves it N times
        nrep = 1 if (world == 1 or slice_io) else world
        e2e = {"value": 1.0 / float(te[0]), "unit": UNIT, "h2d_bytes_per_step": int(npsi * nvec * 8) * nrep,
               "d2h_bytes_per_step": int(npsi * nvec * 8) * nrep,
               "io": "1 GPU: evr_sg4_apply" if world == 1 else
                     ("per rank: H2D of its 1/N slice, NVLink all-gather, terms, NVLink reduce-scatter, D2H of its slice" if slice_io
                      else "per rank: full psi H2D, terms, all-reduce, full H psi D2H")}
        if slice_io:                                            # the slices of the N ranks assemble to the oracle-checked vector
            lo_s, hi_s = evr.lib.slice_bounds(npsi * nvec, world, rank)
            tp.apply(d_psi, d_out)
            torch.cuda.synchronize()
            dev_full = d_out.cpu().view(-1)
            sl_err = float((out_h.view(-1)[lo_s:hi_s] - dev_full[lo_s:hi_s]).abs().max() / dev_full.abs().max())
            te2 = torch.tensor([sl_err], dtype=torch.float64, device="cuda")
            dist.all_reduce(te2, op=dist.ReduceOp.MAX)
            e2e["slice_vs_allreduce_rel_diff"] = float(te2[0])
    allreduce_ms = None
    if world > 1:                                               # the collective alone, for the scaling analysis
        for _ in range(3):
            tp.all_reduce(d_out)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(args.steps):
            tp.all_reduce(d_out)
        a1.record(stream)
        barrier()
        ta = torch.tensor([a0.elapsed_time(a1) / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        allreduce_ms = float(ta[0])
    if rank == 0:
        # nvidia-smi needs 50-200 ms for its first sample and the timed regions above last ~30 ms: keep the same kernels
        # running (untimed, local, no collective) until three samples under load exist (at most 3 s)
        t_wait = time.perf_counter()
        extra = 0
        while len(sampler.rows) < 3 and time.perf_counter() - t_wait < 3.0:
            for _ in range(50):
                op.apply_device_ptr(npsi, d_psi.data_ptr(), d_out.data_ptr(), stream.cuda_stream)
            torch.cuda.synchronize()
            extra += 50
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["sampled"] = f"timed regions + {extra} untimed H|psi> of the same kernels (until 3 samples)"

    # ---- roofline of the term kernel (this rank's launch; algorithmic bytes per SURVEY.md 8d)
    alg1 = op.info(evr.lib.INFO_ALG_BYTES_NPSI1)
    alg = alg1 + (npsi - 1) * op.info(evr.lib.INFO_ALG_BYTES_PER_RHS_EXTRA)
    peak, peak_src = measured_peak()
    achieved = alg / (k_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # ncu figure of the single-GPU launch (profiles/traffic.json); not measured for a term sub-range
                "traffic": (traffic or {}).get("dram_bytes_per_launch") if world == 1 else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
                "kernel": "sg4_term_kernel_fast (permute-in + memset of Hpsi + one launch per term-size class + permute-out, per H|psi>)",
                "flops_per_launch": op.info(evr.lib.INFO_FLOPS_NPSI1) * npsi}

    # ---- parity of THIS run's device result (after the all-reduce when N > 1) against the CPU oracle on the same psi,
    # and the CPU baseline (the same oracle calls, timed; N = 1 only)
    cpu, parity = None, None
    if not args.no_cpu:
        tp.apply(d_psi, d_out)
        torch.cuda.synchronize()
        got = d_out.cpu().numpy()
        if rank == 0:
            from oracle import sg4_oracle
            full = evr.ParamOp(basis, 1, ops) if (lo_, hi_) != (0, basis.nb_SG) else op
            nth = host_threads()
            sec, sample, ref = cpu_time_hpsi(full, psi_h.numpy(), nth, budget_s=25.0 if world == 1 else 0.0)
            from helpers import rel_l2
            errs = [rel_l2(got[i], ref[i]) for i in range(npsi)]
            parity = {"rel_l2": max(errs), "tol": 1e-12, "against": "oracle port (oracle/sg4_oracle.c), same psi, full workload",
                      "n_gpus": world, "ok": bool(max(errs) <= 1e-12)}
            if world == 1:
                cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": nth, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": workload_name(args.D, args.L, npsi), "nb_SG": basis.nb_SG, "grid_points": basis.nqq,
                           "nb": basis.nb, "npsi": npsi, "parallelism": f"terms/{world} ({args.partition}-balanced contiguous ranges) + all-reduce: {tp.collective}" if world > 1 else "1 GPU",
                           "cache": "operator grid + mapping streamed per step (%.0f MB) > L2; no flush needed" % (alg1 / 1e6)
                           if alg1 > 130e6 else "inputs smaller than L2 (L2-warm numbers)",
                           "kernel_path": int(op.info(evr.lib.INFO_PATH)), "iso_flavour": int(op.info(evr.lib.INFO_ISO)), "setup_s": round(t_setup, 2)},
                "e2e": e2e, "allreduce_ms": allreduce_ms, "kernel_ms_per_rank": [round(x, 4) for x in k_ms_ranks], "allreduce_vs_nccl_rel_diff": allreduce_check,
                "gpu_launches": int(launches) + (args.steps if (world > 1 and tp._symm) else 0), "roofline": roofline, "cpu_baseline": cpu, "parity": parity, "clocks": clocks}
        print(json.dumps(line))
        if parity is not None and not parity["ok"]:
            print(f"PARITY FAILURE: rel L2 vs oracle = {parity['rel_l2']:.3e} > 1e-12", file=sys.stderr)
            return 1
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
