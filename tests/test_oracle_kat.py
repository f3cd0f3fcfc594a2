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
