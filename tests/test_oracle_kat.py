"""Known-answer tests that pin the ORACLE's H action to the reference (no GPU needed):
eigenvalues of the Henon-Heiles Hamiltonians shipped in Working_tests/MPI_tests/*_Davidson_openMP/benchmark.

The reference values come from a block-Davidson run converged to conv_ene=1e-4 / conv_resi=5e-4 au
(shell_run:157-165); they agree with the exact eigenvalues of the same matrix to ~1e-7 au.  The
reference's own regression tolerance (1e-8) is run-to-run reproducibility of that Davidson, not accuracy.
"""
import numpy as np
import pytest

from helpers import oracle_apply

TOL_AU = 2e-7


@pytest.mark.parametrize("name,D,L", [("HH6D_L3", 6, 3), ("HH21D_L2", 21, 2)])
def test_oracle_reproduces_reference_eigenvalues(name, D, L, evr, golden):
    basis, op = evr.workloads.henon_heiles(D, L)
    H = oracle_apply(op, np.eye(basis.nb)).T
    assert np.abs(H - H.T).max() < 1e-11
    ev = np.sort(np.linalg.eigvals(H).real)
    ref = np.array(golden["kat"][name]["levels"])
    assert np.abs(ev[: len(ref)] - ref).max() < TOL_AU


def test_oracle_term_ranges_add_up(evr):
    """MPI scheme 1: partial sums over contiguous term ranges add to the full action."""
    basis, op = evr.workloads.henon_heiles(6, 3)
    rng = np.random.default_rng(1)
    psi = rng.standard_normal((2, basis.nb))
    full = oracle_apply(op, psi)
    part = sum(oracle_apply(op, psi, iG_range=r) for r in [(0, 30), (30, 31), (31, basis.nb_SG)])
    assert np.abs(full - part).max() < 1e-12 * np.abs(full).max()


def test_oracle_thread_count_independent(evr):
    basis, op = evr.workloads.henon_heiles(6, 3)
    psi = np.random.default_rng(2).standard_normal((1, basis.nb))
    a = oracle_apply(op, psi, nthreads=1)
    b = oracle_apply(op, psi, nthreads=4)
    assert np.abs(a - b).max() < 1e-12 * np.abs(a).max()


def test_oracle_type10_reduces_to_type1_for_unit_metric(evr):
    """With G = 1, Jac = 1, rho = 1 the type_Op=10 form is -1/2 sum_i d_i d_i + V: identical to the type_Op=1
    action in which dnRGG%d2 is replaced by dnRGG%d1 . dnRGG%d1 (the two oracle code paths cross-check)."""
    from oracle import sg4_oracle as orc
    b = evr.workloads.hm_sg4_basis(3, 3, 3, 1, 2)
    rng = np.random.default_rng(0)
    NQ, n = b.nqq, 3
    V = rng.standard_normal(NQ)
    GG = np.zeros((NQ, n, n), order="F")
    for i in range(n):
        GG[:, i, i] = 1.0
    psi = rng.standard_normal((2, b.nb))
    args = (b.D, b.nb_SG, 1, b.nb, b.LG, b.nDind_SmolyakRep_Tab_nDval, b.WeightSG, b.tab_nq_OF_SRep,
            b.tab_nb_OF_SRep, b.tab_iB_OF_SRep_TO_iB, b.nq_of, b.nb_of, b.B, b.BTw, b.D1)
    h10 = orc.tab_oppsi10(*args, [1, 2, 3], V, GG.ravel(order="F"), np.ones(NQ), np.ones(NQ), psi)
    D1D1 = np.concatenate([(b.tab_basisPrimSG[k][L].D1 @ b.tab_basisPrimSG[k][L].D1).ravel(order="F")
                           for k in range(b.D) for L in range(b.LG + 1)])
    ops = evr.workloads.constant_keo_opgrids(3, 1, np.ones(3), V.reshape(-1, 1, 1))
    tm = np.array(evr.Init_TypeOp(1, 3), dtype=np.int32)
    gz = [o.grid_zero for o in ops]
    gc = [o.grid_cte for o in ops]
    mc = np.array([0.0 if o.Mat_cte is None else o.Mat_cte[0, 0] for o in ops]).reshape(-1, 1)
    grids = [None if (o.grid_zero or o.grid_cte) else np.asfortranarray(o.Grid).ravel(order="F") for o in ops]
    h1 = orc.tab_oppsi(*args, D1D1, 1, tm, gz, gc, mc, grids, psi)
    assert np.abs(h10 - h1).max() < 1e-13 * np.abs(h1).max()


def _pyrazine_autocorrelation(H, nb, times):
    """<psi0|exp(-iHt)|psi0> for psi0 = packed function (1,..,1) on electronic state 2 (the WP0 line
    ' 1 1 1 1 1 1 1 1 1 1 1 1   1 2   1.0 0.' of 12D_propagation_openMP/shell_run)."""
    n = H.shape[0]
    psi0 = np.zeros(n, complex)
    psi0[nb] = 1.0
    w, V = np.linalg.eig(H)
    c0 = np.linalg.solve(V, psi0)
    return np.array([np.vdot(psi0, V @ (np.exp(-1j * w * t) * c0)) for t in times])


def test_oracle_reproduces_reference_pyrazine_autocorrelation(evr, golden):
    """PYR-WP configuration (nb0 = 2, complex wave packet): autocorrelation function of the 12-D pyrazine
    model at L=1 against Working_tests/MPI_tests/12D_propagation_openMP/benchmark (61 samples, 0..6 fs,
    reference tolerance 1e-8).  Exact propagation of the oracle-built H matrix agrees to ~1e-14."""
    basis, op = evr.workloads.pyrazine_12d(1)
    n = basis.nb * 2
    H = oracle_apply(op, np.eye(n)).T
    rows = np.array(golden["kat"]["PYR12D_L1_autocor"]["t_re_im_abs"])
    c = _pyrazine_autocorrelation(H, basis.nb, rows[:, 0])
    assert np.abs(c.real - rows[:, 1]).max() < 1e-10
    assert np.abs(c.imag - rows[:, 2]).max() < 1e-10
    assert np.abs(np.abs(c) - rows[:, 3]).max() < 1e-10
