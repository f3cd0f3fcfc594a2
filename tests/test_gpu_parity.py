"""Parity of the CUDA path (through the C-ABI) against the CPU oracle.  Tolerance: relative L2
<= 1e-12 per right-hand side (north star), eigenvalues <= 1e-6 cm^-1 = 4.6e-12 au against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import oracle_apply, random_psi, rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-12
CM1_IN_AU = 1.0 / 219474.631363


def _check(op, npsi, seed=12345, tol=TOL):
    b = op.BasisnD
    psi = random_psi(b.nb * b.nb0, npsi, seed)
    ref = oracle_apply(op, psi)
    out = op.apply_host(psi)
    for i in range(npsi):
        assert rel_l2(out[i], ref[i]) < tol, (i, rel_l2(out[i], ref[i]))
    return out, ref


@pytest.mark.parametrize("D,L,npsi", [(6, 3, 1), (6, 3, 5), (12, 2, 2), (12, 3, 1), (12, 4, 3), (21, 2, 2), (3, 5, 1), (1, 4, 2), (2, 0, 1)])
def test_henon_heiles_parity(evr, D, L, npsi):
    basis, op = evr.workloads.henon_heiles(D, L)
    _check(op, npsi)
    assert op.info(evr.lib.INFO_LAUNCHES) >= 1


def test_lb_smaller_than_lg_dropped_functions(evr):
    """LB < LG: mapping entries 0 -> read as 0 on gather, skipped on scatter."""
    basis, op = evr.workloads.henon_heiles(5, 4, LB=2)
    assert basis.count0 > 0
    _check(op, 2)


def test_pyrazine_two_states_complex_psi(evr):
    """nb0 = 2, 2x2 potential matrix on the grid, complex psi handled as two real right-hand sides."""
    basis, op = evr.workloads.pyrazine_12d(1)
    assert (basis.nb_SG, basis.nb, basis.nqq) == (13, 27, 39)
    n = basis.nb * 2
    rng = np.random.default_rng(3)
    c = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    psi, Hpsi = evr.ParamPsi.complex(c), evr.ParamPsi()
    evr.sub_OpPsi(psi, Hpsi, op)
    ref = oracle_apply(op, np.stack([c.real, c.imag]))
    assert Hpsi.cplx
    assert rel_l2(Hpsi.CvecB.real, ref[0]) < TOL and rel_l2(Hpsi.CvecB.imag, ref[1]) < TOL
    basis2, op2 = evr.workloads.pyrazine_12d(2)
    _check(op2, 2)
    # Op_Transfo / TransfoOp branch of sub_TabOpPsi (sub_OpPsi.f90:768-775): (H - E0)(H - E0) psi, complex and real vectors
    op.Op_Transfo, op.E0_Transfo = True, 0.37
    r = rng.standard_normal(n)
    out = []
    evr.sub_TabOpPsi([psi, evr.ParamPsi.real(r)], out, op, TransfoOp=True)
    def twice(v):
        h1 = oracle_apply(op, v[None, :])[0] - 0.37 * v
        return oracle_apply(op, h1[None, :])[0] - 0.37 * h1
    assert out[0].cplx and not out[1].cplx
    assert rel_l2(out[0].CvecB.real, twice(c.real)) < TOL and rel_l2(out[0].CvecB.imag, twice(c.imag)) < TOL
    assert rel_l2(out[1].RvecB, twice(r)) < TOL
    plain = []
    evr.sub_TabOpPsi([evr.ParamPsi.real(r)], plain, op)                 # TransfoOp absent: the plain action
    assert rel_l2(plain[0].RvecB, oracle_apply(op, r[None, :])[0]) < TOL


def test_hcn_shape_curvilinear(evr):
    """HCN_UT shape: 3 modes (10+10L, 1+2L, 1+2L), LB=4/LG=5 guess run and LB=6/LG=7, all 10 term
    grids variable, mixed derivatives d1 x d1 (synthetic N(0,1) operator grids)."""
    for LB, LG, sizes in [(4, 5, (46, 850, 9230)), (6, 7, (85, 2310, 36200))]:
        basis = evr.workloads.hm_sg4_basis(3, LB, LG, [10, 1, 1], [10, 2, 2])
        assert (basis.nb_SG, basis.nb, basis.nqq) == sizes
        op = evr.workloads.synthetic_curvilinear(basis)
        assert op.nb_Term == 10
        _check(op, 2)


def test_hno3_shape_curvilinear(evr):
    """HNO3_UT inner SG4 shape: 8 modes nq=nb=1+L, LB=2/LG=4, nb_Term = 45 (all grids variable)."""
    basis = evr.workloads.hm_sg4_basis(8, 2, 4, 1, 1)
    assert (basis.nb_SG, basis.nb, basis.nqq, basis.count0) == (495, 45, 4845, 1410)
    op = evr.workloads.synthetic_curvilinear(basis)
    assert op.nb_Term == 45
    _check(op, 1)


@pytest.mark.parametrize("dcache", ["0", "2", "4"])
def test_generic_kernel_mixed_derivative_sweeps(evr, dcache, monkeypatch):
    """Generic kernel: mixed derivatives through the cached first derivative (EVR_SG4_DCACHE = number of size
    classes that use it: 0 = always on the fly, 4 = every class) with two channels, terms of every size class,
    and an operator where only some mixed / first-derivative terms are present (others grid_zero / constant)."""
    monkeypatch.setenv("EVR_SG4_DCACHE", dcache)
    basis = evr.workloads.hm_sg4_basis(4, 4, 5, [3, 1, 1, 1], [8, 4, 3, 1], nb0=2)
    assert basis.tab_nq_OF_SRep.max() > 768 and basis.tab_nq_OF_SRep.min() <= 48
    op = evr.workloads.synthetic_curvilinear(basis)
    assert op.info(evr.lib.INFO_PATH) == 0
    _check(op, 2)
    # sparse operator: drop some terms, make others constant
    ops = list(op.OpGrid)
    for it, og in enumerate(ops):
        d = og.derive_termQact
        if d in [(1, 3), (2, 0), (2, 2)]:
            ops[it] = evr.OpGrid(d, grid_zero=True)
        elif d in [(1, 2), (3, 0), (3, 4)]:
            ops[it] = evr.OpGrid(d, grid_cte=True, Mat_cte=np.array([[0.3, 0.1], [-0.2, 0.7]]))
    op2 = evr.ParamOp(basis, 1, ops)
    assert op2.info(evr.lib.INFO_PATH) == 0
    _check(op2, 1)
    part = evr.ParamOp(basis, 1, ops, iG_range=(3, 17))
    psi = random_psi(basis.nb * 2, 1, 9)
    assert rel_l2(part.apply_host(psi), oracle_apply(op2, psi, iG_range=(3, 17))) < TOL


@pytest.mark.parametrize("nb0,B", [(1, 24), (2, 19)])
def test_terms_larger_than_shared_memory(evr, nb0, B, monkeypatch):
    """A Smolyak term whose work buffers exceed the 227 KB of shared memory (25^3 = 15 625 points; 20^3 x 2 channels) runs
    in the generic kernel's global-buffer class -- the reference has no term-size limit (heap RDP arrays,
    sub_module_basis_BtoG_GtoB_SG4.f90:2385-2581).  Curvilinear operator (cached mixed-derivative sweeps: three buffers),
    constant-KEO operator, a term range, a block of vectors, the deterministic mode and a one-CTA scratch budget."""
    basis = evr.workloads.hm_sg4_basis(3, 3, 3, 1, [B, B, B], nb0=nb0)
    assert basis.tab_nq_OF_SRep.max() * nb0 * 16 > 227 * 1024
    op = evr.workloads.synthetic_curvilinear(basis)
    assert op.info(evr.lib.INFO_PATH) == 0
    _check(op, 3)
    psi = random_psi(basis.nb * nb0, 1, 4)
    big = int(np.argmax(basis.tab_nq_OF_SRep))
    part = evr.workloads.synthetic_curvilinear(basis, iG_range=(big, big + 2))
    assert rel_l2(part.apply_host(psi), oracle_apply(op, psi, iG_range=(big, big + 2))) < TOL
    rng = np.random.default_rng(3)
    V = rng.standard_normal((basis.nqq, nb0, nb0))
    hh = evr.ParamOp(basis, 1, evr.workloads.constant_keo_opgrids(3, nb0, np.ones(3), np.asfortranarray(V)))
    _check(hh, 2)
    monkeypatch.setenv("EVR_SG4_SCRATCH_MB", "1")
    monkeypatch.setenv("EVR_SG4_DETERMINISTIC", "1")
    det = evr.workloads.synthetic_curvilinear(basis)
    a = det.apply_host(psi)
    assert rel_l2(a, oracle_apply(op, psi)) < TOL
    assert np.array_equal(a, det.apply_host(psi))


def test_fast_path_plan_with_a_generic_remainder(evr):
    """A constant-KEO plan stays on the fast path when a few of its terms are beyond the register tiles (mode of 17 > 16
    points, as in HH 12-D at L = 8) or beyond the shared memory of an SM: those terms run in the generic kernel on the
    caller's vectors, after the fast part.  Also the scaled action (sub_scaledOpPsi) and a block of vectors."""
    import torch
    basis, op = evr.workloads.henon_heiles(4, 8)                  # modes of 1 .. 17 points
    # (switches that send such a plan to the generic kernel as a whole: the parity checks below still apply)
    mixed = not any(os.environ.get(k) for k in ("EVR_SG4_FORCE_GENERIC", "EVR_SG4_DETERMINISTIC")) and os.environ.get("EVR_SG4_MIXED", "1") != "0"
    rest = op.info(evr.lib.INFO_GENERIC_TERMS)
    if mixed:
        assert op.info(evr.lib.INFO_PATH) == 1
        assert 0 < rest < basis.nb_SG // 4
    else:
        assert op.info(evr.lib.INFO_PATH) == 0 and rest == basis.nb_SG
    out, ref = _check(op, 3)
    psi = random_psi(basis.nb, 3, 12345)
    E0, Esc = 0.7, 1.9
    d_psi = torch.from_numpy(psi).cuda(); d_out = torch.empty_like(d_psi)
    op.apply_device_scaled_ptr(3, d_psi.data_ptr(), d_out.data_ptr(), E0, Esc, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_l2(d_out.cpu().numpy(), (ref - E0 * psi) / Esc) < TOL
    # block-ordered internal vector + remainder
    prev = os.environ.get("EVR_SG4_BLOCK_ORDER")
    os.environ["EVR_SG4_BLOCK_ORDER"] = "1"
    try:
        _, op_b = evr.workloads.henon_heiles(4, 8)
    finally:
        if prev is None:
            del os.environ["EVR_SG4_BLOCK_ORDER"]
        else:
            os.environ["EVR_SG4_BLOCK_ORDER"] = prev
    assert op_b.info(evr.lib.INFO_GENERIC_TERMS) == rest
    assert rel_l2(op_b.apply_host(psi), ref) < TOL
    op_b.apply_device_scaled_ptr(3, d_psi.data_ptr(), d_out.data_ptr(), E0, Esc, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_l2(d_out.cpu().numpy(), (ref - E0 * psi) / Esc) < TOL
    # two channels, pool-based flavour: modes of 1, 9, 17 points
    b2 = evr.workloads.hm_sg4_basis(3, 2, 2, 1, [8, 8, 8], nb0=2)     # sizes 1, 9, 17: the 17s go to the generic kernel
    rng = np.random.default_rng(8)
    V = np.asfortranarray(rng.standard_normal((b2.nqq, 2, 2)))
    op2 = evr.ParamOp(b2, 1, evr.workloads.constant_keo_opgrids(3, 2, np.ones(3), V))
    if mixed:
        assert op2.info(evr.lib.INFO_PATH) == 1 and op2.info(evr.lib.INFO_GENERIC_TERMS) == 3
    _check(op2, 2)


def test_host_block_of_long_vectors_is_pipelined_per_vector(evr):
    """evr_sg4_apply on a block of long vectors (>= 1 MB each) overlaps the copies of the neighbouring vectors with the
    kernels of the current one (three streams); the result must equal the one-vector calls, from pageable and from
    page-locked buffers."""
    import torch
    basis, op = evr.workloads.henon_heiles(12, 6)
    assert basis.nb * 8 >= 1 << 20
    psi = random_psi(basis.nb, 3, 21)
    ref = oracle_apply(op, psi, nthreads=8)
    out = op.apply_host(psi)
    for v in range(3):
        assert rel_l2(out[v], ref[v]) < TOL
        assert rel_l2(op.apply_host(psi[v]), ref[v]) < TOL
    pin_in = torch.from_numpy(psi).pin_memory(); pin_out = torch.empty_like(pin_in).pin_memory()
    op.apply_host(pin_in.numpy(), out=pin_out.numpy())
    assert rel_l2(pin_out.numpy(), ref) < TOL


def test_type_op_0_scalar_operator(evr):
    basis = evr.workloads.hm_sg4_basis(4, 3, 3, 1, 2, nb0=2)
    rng = np.random.default_rng(5)
    g = np.asfortranarray(rng.standard_normal((basis.nqq, 2, 2)))
    op = evr.ParamOp(basis, 0, [evr.OpGrid((0, 0), Grid=g)])
    _check(op, 2)


def test_term_ranges_sum_to_full_action(evr):
    """MPI scheme 1 decomposition: plans over ini_iGs ranges give partial sums that add up."""
    basis, op = evr.workloads.henon_heiles(6, 3)
    psi = random_psi(basis.nb, 2, 7)
    full = op.apply_host(psi)
    L = evr.lib.lib()
    acc = np.zeros_like(full)
    for r in range(3):
        b, e = C.c_int(), C.c_int()
        assert L.evr_sg4_ini_iGs(basis.nb_SG, 3, r, C.byref(b), C.byref(e)) == 0
        part = evr.ParamOp(basis, 1, op.OpGrid, iG_range=(b.value, e.value))
        acc += part.apply_host(psi)
    assert rel_l2(acc, full) < TOL
    assert rel_l2(full, oracle_apply(op, psi)) < TOL


def test_gpu_eigenvalues_match_reference_and_oracle(evr, golden):
    """H matrix built column by column with the CUDA path (block of nb unit vectors, as
    Sub_OpPsi_test does, vib.f90:1507): eigenvalues vs oracle <= 1e-6 cm^-1, vs reference benchmark 2e-7 au."""
    basis, op = evr.workloads.henon_heiles(6, 3)
    I = np.eye(basis.nb)
    Hg = op.apply_host(I).T
    Ho = oracle_apply(op, I).T
    eg = np.sort(np.linalg.eigvals(Hg).real)
    eo = np.sort(np.linalg.eigvals(Ho).real)
    assert np.abs(eg - eo).max() < 1e-6 * CM1_IN_AU
    ref = np.array(golden["kat"]["HH6D_L3"]["levels"])
    assert np.abs(eg[: len(ref)] - ref).max() < 2e-7


def test_error_behaviour_mirrors_reference(evr):
    basis, op = evr.workloads.henon_heiles(3, 2)
    with pytest.raises(evr.EvrStop):                      # STOP: size(Psi) = 0  (sub_OpPsi_SG4.f90:738-743)
        evr.sub_TabOpPsi_FOR_SGtype4([], [], op)
    with pytest.raises(evr.EvrStop):                      # STOP: Psi(1) is complex (:744-749)
        evr.sub_TabOpPsi_FOR_SGtype4([evr.ParamPsi.complex(np.zeros(basis.nb))], [], op)
    L = evr.lib.lib()
    z = np.zeros(basis.nb)
    assert L.evr_sg4_apply(op.plan(), 0, z.ctypes.data, z.ctypes.data) != 0
    assert b"size(Psi)" in L.evr_sg4_last_error()


def test_linearity_and_repeatability(evr):
    basis, op = evr.workloads.henon_heiles(12, 4)
    psi = random_psi(basis.nb, 2, 11)
    a = op.apply_host(psi)
    lin = op.apply_host((2.0 * psi[0] - 3.0 * psi[1])[None, :])[0]
    assert rel_l2(lin, 2.0 * a[0] - 3.0 * a[1]) < 1e-12
    b = op.apply_host(psi)
    assert rel_l2(a, b) < 1e-13       # atomics reorder the sum: reproducible to rounding only (like the reference's OMP ATOMIC)


def test_device_entry_point_with_torch(evr):
    import torch
    basis, op = evr.workloads.henon_heiles(12, 3)
    psi = random_psi(basis.nb, 3, 13)
    d_psi = torch.from_numpy(psi).cuda()
    d_out = torch.full_like(d_psi, 7.0)                   # must be overwritten, not accumulated
    op.apply_device_ptr(3, d_psi.data_ptr(), d_out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rel_l2(d_out.cpu().numpy(), oracle_apply(op, psi)) < TOL


def test_type_op_10_cached_metric(evr):
    """type_Op=10: d_i J G^ij d_j form with G/Jac/sqrt(rho/J) cached per grid point; HCN shape
    (3 modes, 10+10L / 1+2L), a 5-mode nq=nb=1+L shape, and two channels."""
    from helpers import oracle_apply10
    for basis in (evr.workloads.hm_sg4_basis(3, 4, 5, [10, 1, 1], [10, 2, 2]),
                  evr.workloads.hm_sg4_basis(5, 2, 4, 1, 1),
                  evr.workloads.hm_sg4_basis(3, 3, 3, 1, 2, nb0=2)):
        op = evr.workloads.synthetic_type10(basis)
        psi = random_psi(basis.nb * basis.nb0, 3, 21)
        ref = oracle_apply10(op, psi)
        out = op.apply_host(psi)
        for i in range(3):
            assert rel_l2(out[i], ref[i]) < TOL
    nov = evr.workloads.synthetic_type10(evr.workloads.hm_sg4_basis(3, 3, 3, 1, 2), with_V=False)
    psi = random_psi(nov.BasisnD.nb, 1, 22)
    assert rel_l2(nov.apply_host(psi), oracle_apply10(nov, psi)) < TOL
    # terms of more than 2048 grid points (round-1 limit of this kernel): 60 x 7 x 5 = 2100 at l = (2, 3, 2)
    big = evr.workloads.hm_sg4_basis(3, 7, 7, [20, 1, 1], [20, 2, 2])
    assert int(big.tab_nq_OF_SRep.max()) > 2048
    op = evr.workloads.synthetic_type10(big)
    psi = random_psi(big.nb, 2, 23)
    ref = oracle_apply10(op, psi)
    out = op.apply_host(psi)
    for i in range(2):
        assert rel_l2(out[i], ref[i]) < TOL
    # storage of the metric tensor: upper triangle when GG(q,j,i) == GG(q,i,j) exactly (n(n+1)/2 + 2 doubles per point instead
    # of n^2 + 2), full otherwise -- the reference formula takes GG as given (sub_OpPsi_SG4.f90:1598-1618)
    b5 = evr.workloads.hm_sg4_basis(5, 2, 4, 1, 1)
    sym = evr.workloads.synthetic_type10(b5)
    G = 0.5 * (sym.GG + sym.GG.transpose(0, 2, 1))
    sym = evr.ParamOp10(b5, np.asfortranarray(G), sym.Jac, sym.sqRhoOVERJac, V=sym.V)
    nq, n = b5.nqq, 5
    sym_bytes = sym.info(evr.lib.INFO_ALG_BYTES_NPSI1)
    G2 = G.copy(); G2[:, 0, 1] += 0.25                        # not symmetric any more
    asym = evr.ParamOp10(b5, np.asfortranarray(G2), sym.Jac, sym.sqRhoOVERJac, V=sym.V)
    assert asym.info(evr.lib.INFO_ALG_BYTES_NPSI1) - sym_bytes == nq * 8 * (n * n - n * (n + 1) // 2)
    psi = random_psi(b5.nb, 2, 24)
    for op in (sym, asym):
        ref = oracle_apply10(op, psi)
        out = op.apply_host(psi)
        for i in range(2):
            assert rel_l2(out[i], ref[i]) < TOL


def test_gpu_pyrazine_autocorrelation_matches_reference(evr, golden):
    """The H matrix of the 12-D two-state pyrazine model built column by column with the CUDA path (complex
    psi = two real right-hand sides) reproduces the reference's autocorrelation benchmark (1e-8 tolerance there)."""
    from test_oracle_kat import _pyrazine_autocorrelation
    basis, op = evr.workloads.pyrazine_12d(1)
    n = basis.nb * 2
    H = op.apply_host(np.eye(n)).T
    rows = np.array(golden["kat"]["PYR12D_L1_autocor"]["t_re_im_abs"])
    c = _pyrazine_autocorrelation(H, basis.nb, rows[:, 0])
    assert max(np.abs(c.real - rows[:, 1]).max(), np.abs(c.imag - rows[:, 2]).max(), np.abs(np.abs(c) - rows[:, 3]).max()) < 1e-10


@pytest.mark.parametrize("ndev", [1, 2])
def test_cpp_host_mirror(evr, tmp_path, ndev):
    """The C++ mirror of mod_OpPsi (host/evr_oppsi.hpp: param_psi, param_Op, sub_OpPsi, sub_TabOpPsi) driven by a
    compiled program over the C-ABI: block of real vectors + complex wave packets, pyrazine two-state model.
    ndev = 2: the same compiled program calls evr_sg4_set_devices(2) and drives two GPUs through include/evr_sg4.h only."""
    import os
    if ndev > 1:
        import torch
        if torch.cuda.device_count() < ndev:
            pytest.skip("needs two GPUs (gpurun --gpus 2)")
    import struct
    import subprocess
    from helpers import flat_op
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = tmp_path / "host_mirror_main"
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", os.path.join(root, "tests", "cpp", "host_mirror_main.cpp"),
                           "-o", str(exe), "-L" + os.path.join(root, "elvibrot-tnumtana_b200"), "-levr_sg4",
                           "-Wl,-rpath," + os.path.join(root, "elvibrot-tnumtana_b200")])
    basis, op = evr.workloads.pyrazine_12d(2)
    b = basis
    tm, gz, gc, mc, grids = flat_op(op)
    n = b.nb * b.nb0
    npsi, ncplx = 3, 2
    rng = np.random.default_rng(9)
    psi = rng.standard_normal((npsi, n))
    cpsi = rng.standard_normal((ncplx, n)) + 1j * rng.standard_normal((ncplx, n))

    def blk(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        return struct.pack("<q", a.size) + a.tobytes()
    with open(tmp_path / "in.bin", "wb") as f:
        f.write(blk([b.D, b.nb_SG, b.nb0, b.nb, b.LG, op.type_Op, op.nb_Term, npsi, ncplx], np.int64))
        for a, dt in [(b.nDind_SmolyakRep_Tab_nDval, np.int32), (b.WeightSG, np.float64), (b.tab_nq_OF_SRep, np.int32),
                      (b.tab_nb_OF_SRep, np.int32), (b.tab_iB_OF_SRep_TO_iB, np.int32), (b.nq_of, np.int32), (b.nb_of, np.int32),
                      (b.B, np.float64), (b.BTw, np.float64), (b.D1, np.float64), (b.D2, np.float64),
                      (tm, np.int32), (np.array(gz, dtype=np.uint8), np.uint8), (np.array(gc, dtype=np.uint8), np.uint8), (mc, np.float64)]:
            f.write(blk(a, dt))
        for g in grids:
            f.write(blk(np.zeros(0) if g is None else g, np.float64))
        f.write(blk(psi, np.float64))
        f.write(blk(np.stack([cpsi.real, cpsi.imag], axis=-1), np.float64))
    subprocess.check_call([str(exe), str(tmp_path / "in.bin"), str(tmp_path / "out.bin")] + ([str(ndev)] if ndev > 1 else []))
    raw = open(tmp_path / "out.bin", "rb").read()
    stops, count = struct.unpack("<qq", raw[:16])
    assert stops == 2                       # both reference STOP conditions raised
    assert count == npsi + ncplx            # nb_OpPsi bookkeeping
    out = np.frombuffer(raw[16:], dtype=np.float64)
    Hr = out[: npsi * n].reshape(npsi, n)
    Hc = out[npsi * n: npsi * n + 2 * ncplx * n].reshape(ncplx, n, 2)
    T2 = out[npsi * n + 2 * ncplx * n:]
    assert T2.size == n
    h1 = oracle_apply(op, psi[:1])[0] - 0.37 * psi[0]
    assert rel_l2(T2, oracle_apply(op, h1[None, :])[0] - 0.37 * h1) < TOL        # TransfoOp: (H - E0)^2 psi
    ref_r = oracle_apply(op, psi)
    ref_c = oracle_apply(op, np.concatenate([cpsi.real, cpsi.imag]))
    for i in range(npsi):
        assert rel_l2(Hr[i], ref_r[i]) < TOL
    for i in range(ncplx):
        assert rel_l2(Hc[i, :, 0], ref_c[i]) < TOL and rel_l2(Hc[i, :, 1], ref_c[ncplx + i]) < TOL


def test_fast_path_without_potential_and_with_constant_shift(evr):
    """Operator forms the fast path must also accept: pure kinetic energy ((0,0) term grid_zero) and a constant
    (0,0) term (grid_cte, folded into the per-term shift)."""
    basis = evr.workloads.hm_sg4_basis(6, 3, 3, 1, 2)
    ops = evr.workloads.constant_keo_opgrids(6, 1, np.linspace(0.5, 1.5, 6), None)
    # which kernel is selected is only asserted when the selection is not overridden from the environment
    fast = 0 if os.environ.get("EVR_SG4_FORCE_GENERIC") == "1" else 1
    op = evr.ParamOp(basis, 1, ops)
    _check(op, 2)
    assert op.info(evr.lib.INFO_PATH) == fast
    ops2 = evr.workloads.constant_keo_opgrids(6, 1, np.ones(6), None)
    ops2[0] = evr.OpGrid((0, 0), grid_cte=True, Mat_cte=np.array([[0.37]]))
    op2 = evr.ParamOp(basis, 1, ops2)
    _check(op2, 1)
    # first-derivative constant terms (f1) are folded into the same 1-D kinetic matrix
    ops3 = evr.workloads.constant_keo_opgrids(6, 1, np.ones(6), None)
    for it, og in enumerate(ops3):
        if og.derive_termQact[0] > 0 and og.derive_termQact[1] == 0:
            ops3[it] = evr.OpGrid(og.derive_termQact, grid_cte=True, Mat_cte=np.array([[0.1 * og.derive_termQact[0]]]))
    op3 = evr.ParamOp(basis, 1, ops3)
    _check(op3, 1)
    assert op3.info(evr.lib.INFO_PATH) == fast
    # a constant MIXED derivative term does not qualify: generic kernel
    ops4 = evr.workloads.constant_keo_opgrids(6, 1, np.ones(6), None)
    for it, og in enumerate(ops4):
        if og.derive_termQact == (1, 2):
            ops4[it] = evr.OpGrid((1, 2), grid_cte=True, Mat_cte=np.array([[-0.2]]))
    op4 = evr.ParamOp(basis, 1, ops4)
    _check(op4, 1)
    assert op4.info(evr.lib.INFO_PATH) == 0


def test_empty_term_range_and_single_term(evr):
    basis, op = evr.workloads.henon_heiles(4, 2)
    psi = random_psi(basis.nb, 1, 3)
    empty = evr.ParamOp(basis, 1, op.OpGrid, iG_range=(5, 5))
    assert np.abs(empty.apply_host(psi)).max() == 0.0
    one = evr.ParamOp(basis, 1, op.OpGrid, iG_range=(7, 8))
    assert rel_l2(one.apply_host(psi), oracle_apply(op, psi, iG_range=(7, 8))) < TOL


def test_iso_flavour_selection_and_agreement_with_plain_fast_path(evr, monkeypatch):
    """The constant-matrix ("iso") instantiation is selected when all modes of one size share one 1-D basis
    (Henon-Heiles: D identical Hm modes), not when the kinetic constants differ per mode; both flavours of the fast
    path give the same H|psi> (same arithmetic, different tiling).  Sizes 9..15 (L up to 7 in 2-D/3-D) run the
    runtime-size tiles of the iso kernel, even sizes (nq = 2 + L ... ) likewise."""
    if os.environ.get("EVR_SG4_FORCE_GENERIC") == "1" or "EVR_SG4_ISO" in os.environ:
        pytest.skip("kernel selection overridden from the environment")
    for D, L in [(6, 3), (12, 3), (3, 7), (2, 7), (7, 5)]:
        basis, op = evr.workloads.henon_heiles(D, L)
        out, ref = _check(op, 2)
        assert op.info(evr.lib.INFO_PATH) == 1 and op.info(evr.lib.INFO_ISO) == 1
        monkeypatch.setenv("EVR_SG4_ISO", "0")
        plain = evr.ParamOp(basis, 1, op.OpGrid)
        psi = random_psi(basis.nb, 2)
        o2 = plain.apply_host(psi)
        assert plain.info(evr.lib.INFO_ISO) == 0 and plain.info(evr.lib.INFO_PATH) == 1
        monkeypatch.delenv("EVR_SG4_ISO")
        for i in range(2):
            assert rel_l2(out[i], o2[i]) < 1e-13
    # different kinetic constants per mode: one block per (mode, size) -> plain fast path
    basis = evr.workloads.hm_sg4_basis(6, 3, 3, 1, 2)
    op = evr.ParamOp(basis, 1, evr.workloads.constant_keo_opgrids(6, 1, np.linspace(0.5, 1.5, 6), None))
    _check(op, 1)
    assert op.info(evr.lib.INFO_PATH) == 1 and op.info(evr.lib.INFO_ISO) == 0
    # mode sizes other than 3, 5, 7 (here nq = nb = 2 + L: 2, 3, 4, 5): those terms use the pool-based instantiations
    basis = evr.workloads.hm_sg4_basis(5, 3, 3, 2, 1)
    op = evr.ParamOp(basis, 1, evr.workloads.constant_keo_opgrids(5, 1, np.ones(5), None))
    _check(op, 2)
    assert op.info(evr.lib.INFO_ISO) == 0 and op.info(evr.lib.INFO_PATH) == 1


def test_two_iso_plans_with_different_bases_alternate(evr):
    """The iso matrices live in one __constant__ array per device: alternating between plans whose 1-D bases differ
    must re-bind the array (each result still matches the oracle)."""
    b1, op1 = evr.workloads.henon_heiles(4, 3)
    b2 = evr.workloads.hm_sg4_basis(4, 3, 3, 1, 2, scaleQ=1.3)
    op2 = evr.ParamOp(b2, 1, evr.workloads.constant_keo_opgrids(4, 1, np.ones(4), None))
    for _ in range(3):
        _check(op1, 1)
        _check(op2, 1)


def test_device_scaled_entry_point_chebyshev_step(evr, monkeypatch):
    """evr_sg4_apply_device_scaled = H|psi> followed by sub_scaledOpPsi (sub_OpPsi.f90:2823-2866) on the device:
    (H psi - E0 psi)/Esc, checked against the oracle for the plain, the block-ordered (fused into the un-permute
    kernel) and the generic path, plus three steps of the Chebyshev recursion of the propagator
    (phi_{k+1} = 2 Hs phi_k - phi_{k-1}, sub_module_propa_march.f90:4294-4345) with vectors resident on the device."""
    import torch
    E0, Esc = 0.37, 2.5
    st = torch.cuda.current_stream().cuda_stream
    for env in ({}, {"EVR_SG4_BLOCK_ORDER": "1"}, {"EVR_SG4_FORCE_GENERIC": "1"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        basis, op = evr.workloads.henon_heiles(6, 3)
        for k in env:
            monkeypatch.delenv(k)
        psi = random_psi(basis.nb, 2, 17)
        ref = (oracle_apply(op, psi) - E0 * psi) / Esc
        d_psi = torch.from_numpy(psi).cuda()
        d_out = torch.full_like(d_psi, -3.0)
        op.apply_device_scaled_ptr(2, d_psi.data_ptr(), d_out.data_ptr(), E0, Esc, st)
        torch.cuda.synchronize()
        out = d_out.cpu().numpy()
        for i in range(2):
            assert rel_l2(out[i], ref[i]) < TOL
        # Chebyshev recursion on the device vs the same recursion with the oracle on the host
        p0, p1 = d_psi[:1].clone(), d_out[:1].clone()
        h0, h1 = psi[:1].copy(), ref[:1].copy()
        for _ in range(3):
            d_t = torch.empty_like(p1)
            op.apply_device_scaled_ptr(1, p1.data_ptr(), d_t.data_ptr(), E0, Esc, st)
            p0, p1 = p1, 2.0 * d_t - p0
            ht = (oracle_apply(op, h1) - E0 * h1) / Esc
            h0, h1 = h1, 2.0 * ht - h0
        torch.cuda.synchronize()
        assert rel_l2(p1.cpu().numpy()[0], h1[0]) < 1e-11
    with pytest.raises(Exception):
        op.apply_device_scaled_ptr(1, d_psi.data_ptr(), d_out.data_ptr(), 0.0, 0.0, st)


def test_fast_path_with_matrix_pool_in_global_memory(evr):
    """Eight modes with eight different scalings and L = 4: 8 x 5 distinct [B|BTw|T] blocks = 32 KB, more than the
    24 KB shared-memory pool -> the instantiations that read the pool from global memory (sg4_fast_inst0.cu)."""
    basis = evr.workloads.hm_sg4_basis(8, 4, 4, 1, 2, scaleQ=np.linspace(0.8, 1.5, 8))
    op = evr.ParamOp(basis, 1, evr.workloads.constant_keo_opgrids(8, 1, np.linspace(0.9, 1.2, 8), None))
    _check(op, 2)
    if os.environ.get("EVR_SG4_FORCE_GENERIC") != "1":
        assert op.info(evr.lib.INFO_PATH) == 1 and op.info(evr.lib.INFO_ISO) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("np_,n", [(2, 1001), (3, 4096), (8, 12345), (1, 7), (4, 2)])
def test_peer_memory_allreduce_kernel_single_device(evr, np_, n):
    """evr_sg4_allreduce_slices (include/evr_sg4_comm.h) with all "ranks" on one device: after every rank's call all
    buffers hold the sum in the fixed order 0..np-1, bit-identical (odd lengths, slices smaller than a vector unit)."""
    import ctypes as C
    import torch
    torch.manual_seed(n)
    bufs = [torch.randn(n, dtype=torch.float64, device="cuda") for _ in range(np_)]
    want = bufs[0].clone()
    for b in bufs[1:]:
        want = want + b
    ptrs = (C.c_void_p * np_)(*[b.data_ptr() for b in bufs])
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for r in range(np_):
        evr.lib.check(evr.lib.lib().evr_sg4_allreduce_slices(ptrs, np_, r, n, st), "allreduce")
    torch.cuda.synchronize()
    for b in bufs:
        assert torch.equal(b, want)
    assert evr.lib.lib().evr_sg4_allreduce_slices(ptrs, 0, 0, n, st) != 0      # bad arguments fail loudly


@pytest.mark.gpu
@pytest.mark.parametrize("np_,n", [(2, 1001), (4, 4096), (8, 12345), (3, 2)])
def test_allreduce_with_in_kernel_barriers_single_device(evr, np_, n):
    """evr_sg4_allreduce_fused: the two cross-rank barriers are flag exchanges inside the kernel.  All "ranks" on one device,
    one stream each (their small grids are co-resident, as the kernels of the real ranks are on their own devices); three
    consecutive calls on the same buffers (call numbers 1, 2, 3) without any host synchronisation in between."""
    import ctypes as C
    import torch
    torch.manual_seed(n)
    W = evr.lib.FLAG_WORDS
    bufs = [torch.randn(n, dtype=torch.float64, device="cuda") for _ in range(np_)]
    flags = [torch.zeros(W, dtype=torch.int64, device="cuda") for _ in range(np_)]
    streams = [torch.cuda.Stream() for _ in range(np_)]
    torch.cuda.synchronize()
    ptrs = (C.c_void_p * np_)(*[b.data_ptr() for b in bufs])
    fptrs = (C.c_void_p * np_)(*[f.data_ptr() for f in flags])
    want = bufs[0].clone()
    for b in bufs[1:]:
        want = want + b
    for call in (1, 2, 3):
        for r in range(np_):
            evr.lib.check(evr.lib.lib().evr_sg4_allreduce_fused(ptrs, fptrs, np_, r, n, call, C.c_void_p(streams[r].cuda_stream)), "allreduce_fused")
        if call == 1:
            torch.cuda.synchronize()
            for b in bufs:
                assert torch.equal(b, want)
    torch.cuda.synchronize()
    for _ in range(2):                   # two more all-reduces of np identical vectors, summed in the kernel's order
        acc = want.clone()
        for _r in range(np_ - 1):
            acc = acc + want
        want = acc
    for b in bufs:
        assert torch.equal(b, want)
    assert evr.lib.lib().evr_sg4_allreduce_fused(ptrs, fptrs, np_, 0, n, 0, None) != 0      # call number 0 is rejected


# ---- the benchmarked configuration itself (BASELINE.json configs[4] at LB=LG=6 and 7, default plan settings):
# block-ordered packed vector, batched work items of every size class, iso + pool-based instantiations side by side
_BENCH_CACHE = {}


def _bench_case(evr, L):
    if L not in _BENCH_CACHE:
        basis, op = evr.workloads.henon_heiles(12, L)
        psi = random_psi(basis.nb, 1)
        import os as _os
        ref = oracle_apply(op, psi, nthreads=max(1, len(_os.sched_getaffinity(0))))
        _BENCH_CACHE[L] = (basis, op, psi, ref)
    return _BENCH_CACHE[L]


@pytest.mark.parametrize("L", [6, 7])
def test_benchmarked_configuration_parity(evr, L):
    basis, op, psi, ref = _bench_case(evr, L)
    assert basis.nqq == {6: 4195284, 7: 23826372}[L]
    out = op.apply_host(psi)
    if not any(os.environ.get(k) for k in ("EVR_SG4_ISO", "EVR_SG4_FORCE_GENERIC")):      # default kernel selection
        assert op.info(evr.lib.INFO_PATH) == 1 and op.info(evr.lib.INFO_ISO) == 1
    assert rel_l2(out[0], ref[0]) < TOL, rel_l2(out[0], ref[0])


@pytest.mark.parametrize("L,nranks", [(6, 8), (7, 8)])
def test_benchmarked_configuration_term_ranges_of_8_ranks(evr, L, nranks):
    """The N = 8 decomposition of bench.py (points-balanced contiguous term ranges): the partial sums of the eight
    plans add up to the oracle's full H|psi>."""
    basis, op, psi, ref = _bench_case(evr, L)
    acc = np.zeros_like(ref)
    prev = 0
    for r in range(nranks):
        lo, hi = evr.distributed.balanced_iGs(basis.tab_nq_OF_SRep, nranks, r)
        assert lo == prev
        prev = hi
        part = evr.ParamOp(basis, 1, op.OpGrid, iG_range=(lo, hi))
        acc += part.apply_host(psi)
        part.close()
    assert prev == basis.nb_SG
    assert rel_l2(acc[0], ref[0]) < TOL, rel_l2(acc[0], ref[0])


def test_batched_and_unbatched_work_items_agree(evr, monkeypatch):
    """Batches of same-schedule terms (default) vs one term per work item (EVR_SG4_BATCH=0), several right-hand sides."""
    basis, op = evr.workloads.henon_heiles(9, 4)
    psi = random_psi(basis.nb, 3, 11)
    a = op.apply_host(psi)
    monkeypatch.setenv("EVR_SG4_BATCH", "0")
    op2 = evr.ParamOp(basis, 1, op.OpGrid)
    b = op2.apply_host(psi)
    ref = oracle_apply(op, psi)
    for i in range(3):
        assert rel_l2(a[i], ref[i]) < TOL and rel_l2(b[i], ref[i]) < TOL


def _ngpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.skipif(_ngpus() < 2, reason="needs two GPUs with peer access (gpurun --gpus 2)")
@pytest.mark.parametrize("case", ["hh12d_L4", "hcn_generic", "pyrazine_nb0_2"])
def test_set_devices_multi_gpu_plan_through_the_cabi(evr, case):
    """evr_sg4_set_devices(2): one process, the plan spans two GPUs (term ranges of equal grid points), psi / H psi move
    slice-wise, the partial sums are reduced over NVLink peer memory -- all through the C-ABI, no torch.distributed."""
    import torch
    ndev = min(_ngpus(), 4)
    if case == "hh12d_L4":
        basis, op1 = evr.workloads.henon_heiles(12, 4)
        npsi = 3
    elif case == "hcn_generic":
        basis = evr.workloads.hm_sg4_basis(3, 4, 5, [10, 1, 1], [10, 2, 2])
        op1 = evr.workloads.synthetic_curvilinear(basis)
        npsi = 2
    else:
        basis, op1 = evr.workloads.pyrazine_12d(2)
        npsi = 2
    psi = random_psi(basis.nb * basis.nb0, npsi, 21)
    ref = oracle_apply(op1, psi)
    evr.lib.set_devices(ndev)
    try:
        opn = evr.ParamOp(basis, op1.type_Op, op1.OpGrid, mode_of_Qact=op1.mode_of_Qact)
        out = opn.apply_host(psi)
        assert opn.info(evr.lib.INFO_DEVICES) == ndev
        assert opn.info(evr.lib.INFO_NQ_LOCAL) == basis.nqq
        for i in range(npsi):
            assert rel_l2(out[i], ref[i]) < TOL, (i, rel_l2(out[i], ref[i]))
        # device-resident entry: psi / Hpsi on device 0
        x = torch.from_numpy(psi).cuda(0)
        y = torch.empty_like(x)
        opn.apply_device_ptr(npsi, x.data_ptr(), y.data_ptr(), torch.cuda.current_stream(0).cuda_stream)
        torch.cuda.synchronize(0)
        for i in range(npsi):
            assert rel_l2(y[i].cpu().numpy(), ref[i]) < TOL
        # pinned (registered) host buffers: same result, repeated calls
        L = evr.lib.lib()
        xh, yh = np.ascontiguousarray(psi), np.empty_like(psi)
        evr.lib.check(L.evr_sg4_host_register(xh.ctypes.data, xh.nbytes))
        evr.lib.check(L.evr_sg4_host_register(yh.ctypes.data, yh.nbytes))
        for _ in range(3):
            opn.apply_host(xh, out=yh)
        evr.lib.check(L.evr_sg4_host_unregister(xh.ctypes.data))
        evr.lib.check(L.evr_sg4_host_unregister(yh.ctypes.data))
        assert rel_l2(yh, ref) < TOL
        opn.close()
    finally:
        evr.lib.set_devices(1)


def test_peer_kernels_allgather_and_reduce_single_device(evr):
    """evr_sg4_allgather_slices / evr_sg4_reduce_slice / evr_sg4_reduce_to on buffers of ONE device standing in for the peers
    (the kernels only see pointers): slice bounds, odd lengths, bit-exact sums in the fixed order."""
    import torch
    L = evr.lib.lib()
    for np_, n in [(2, 1001), (3, 4096), (8, 12345), (4, 7), (1, 5)]:
        g = torch.Generator().manual_seed(n)
        bufs = [torch.randn(n, dtype=torch.float64, generator=g).cuda() for _ in range(np_)]
        orig = [b.clone() for b in bufs]
        ptrs = (C.c_void_p * np_)(*[b.data_ptr() for b in bufs])
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        bounds = [evr.lib.slice_bounds(n, np_, r) for r in range(np_)]
        assert bounds[0][0] == 0 and bounds[-1][1] == n and all(bounds[r][1] == bounds[r + 1][0] for r in range(np_ - 1))
        assert all(lo % 2 == 0 for lo, _ in bounds)
        # reduce_to: plain sum in the order 0..np-1
        dst = torch.empty(n, dtype=torch.float64, device="cuda")
        evr.lib.check(L.evr_sg4_reduce_to(ptrs, np_, n, dst.data_ptr(), st))
        ref = orig[0].clone()
        for r in range(1, np_):
            ref += orig[r]
        assert torch.equal(dst, ref)
        # reduce_slice by every "rank": slice r of buffer r holds the sum, everything else untouched
        for r in range(np_):
            evr.lib.check(L.evr_sg4_reduce_slice(ptrs, np_, r, n, st))
        torch.cuda.synchronize()
        for r, (lo, hi) in enumerate(bounds):
            if np_ > 1:
                assert torch.equal(bufs[r][lo:hi], ref[lo:hi])
                mask = torch.ones(n, dtype=torch.bool, device="cuda")
                mask[lo:hi] = False
                assert torch.equal(bufs[r][mask], orig[r][mask])
        # allgather: every buffer ends up with the owners' slices
        want = torch.cat([bufs[r][lo:hi] for r, (lo, hi) in enumerate(bounds)]).clone()
        for r in range(np_):
            evr.lib.check(L.evr_sg4_allgather_slices(ptrs, np_, r, n, st))
        torch.cuda.synchronize()
        for r in range(np_):
            assert torch.equal(bufs[r], want)


def test_in_place_call_is_rejected(evr):
    import torch
    basis, op = evr.workloads.henon_heiles(4, 2)
    x = torch.zeros(basis.nb, dtype=torch.float64, device="cuda")
    with pytest.raises(evr.EvrSg4Error, match="overlap"):
        op.apply_device_ptr(1, x.data_ptr(), x.data_ptr())
    h = np.zeros(basis.nb)
    with pytest.raises(evr.EvrSg4Error, match="overlap"):
        op.apply_host(h, out=h)


def test_model_potential_grid_built_on_the_device(evr):
    """evr_sg4_model_grid against the host evaluation of the same closed-form potentials, point by point in
    the reference's grid order (first mode fastest, terms in iG order)."""
    for D, L in [(6, 3), (12, 4), (3, 5)]:
        basis = evr.workloads.hm_sg4_basis(D, L, L, 1, 2)
        V_host = evr.workloads.henon_heiles_potential(basis)
        V_dev = evr.workloads.model_potential_device(basis, 1, [evr.workloads.LAMBDA_HH])
        assert np.abs(V_dev - V_host).max() <= 1e-13 * max(1.0, np.abs(V_host).max())
    basis = evr.workloads.hm_sg4_basis(4, 3, 3, 1, 2)
    k = np.array([1.0, 0.5, 2.0, 0.25])
    Vh = np.concatenate([evr.workloads._term_outer_sum(basis, iG, lambda kk, x: 0.5 * k[kk] * x * x).ravel(order="F")
                         for iG in range(basis.nb_SG)])
    assert np.abs(evr.workloads.model_potential_device(basis, 2, k) - Vh).max() < 1e-13
    with pytest.raises(evr.EvrSg4Error, match="model"):
        evr.workloads.model_potential_device(basis, 7, [1.0])


@pytest.mark.parametrize("case", ["hh12d_L4_fast", "hh12d_L4_block_order", "pyrazine_nb0_2", "hcn_generic"])
def test_deterministic_mode_is_bit_reproducible(evr, monkeypatch, case):
    """EVR_SG4_DETERMINISTIC=1: the scatter stages every weighted entry and a second kernel sums the entries of
    each packed element in a fixed order -- repeated calls are bit-identical and still match the oracle."""
    monkeypatch.setenv("EVR_SG4_DETERMINISTIC", "1")
    if case == "hh12d_L4_block_order":
        monkeypatch.setenv("EVR_SG4_BLOCK_ORDER", "1")
    if case.startswith("hh12d"):
        basis, op = evr.workloads.henon_heiles(12, 4)
        npsi = 2
    elif case == "pyrazine_nb0_2":
        basis, op = evr.workloads.pyrazine_12d(2)
        npsi = 2
    else:
        basis = evr.workloads.hm_sg4_basis(3, 4, 5, [10, 1, 1], [10, 2, 2])
        op = evr.workloads.synthetic_curvilinear(basis)
        npsi = 3
    psi = random_psi(basis.nb * basis.nb0, npsi, 5)
    a = op.apply_host(psi).copy()
    for _ in range(3):
        assert np.array_equal(op.apply_host(psi), a)
    op2 = evr.ParamOp(basis, op.type_Op, op.OpGrid, mode_of_Qact=op.mode_of_Qact)      # a second plan gives the same bits
    assert np.array_equal(op2.apply_host(psi), a)
    ref = oracle_apply(op, psi)
    for i in range(npsi):
        assert rel_l2(a[i], ref[i]) < TOL
