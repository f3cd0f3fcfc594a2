"""1-D primitive tables: Gauss-Hermite data vs the reference's Internal_data tables, identities."""
import numpy as np
import pytest


def test_gauss_hermite_matches_reference_tables(evr, golden):
    for nq_s, tab in golden["herm_quadra"].items():
        nq = int(nq_s)
        if nq > 30:
            continue
        x, w = evr.primitives.gauss_hermite(nq)
        assert np.allclose(x, tab["x"], rtol=0, atol=2e-13), nq
        assert np.allclose(w, tab["w"], rtol=2e-12, atol=0), nq


@pytest.mark.parametrize("n", [1, 3, 5, 9, 15])
def test_hm_primitive_identities(evr, n):
    p = evr.primitives.hm_primitive(n, n)
    I = np.eye(n)
    assert np.abs(p.BTw @ p.B - I).max() < 1e-12          # orthonormal basis, exact quadrature
    assert np.abs(p.B @ p.BTw - I).max() < 1e-11          # nq = nb: grid projector is the identity
    d0, d1, d2 = evr.primitives.hermite_functions(p.x, n)
    assert np.abs(p.D1 @ p.B - d1).max() < 1e-10          # dnRGG%d1 B = dB  (check at sub_module_basis.f90:2399-2440)
    assert np.abs(p.D2 @ p.B - d2).max() < 1e-9
    # harmonic oscillator: (-1/2 d2 + x^2/2) phi_l = (l + 1/2) phi_l
    H = p.BTw @ (-0.5 * p.D2 + np.diag(0.5 * p.x ** 2)) @ p.B
    ev = np.sort(np.linalg.eigvals(H).real)
    if n >= 3:
        assert abs(ev[0] - 0.5) < 1e-10


def test_level_sizes_extrapolation(evr):
    """nb(L) beyond LB continues with the last increment (sub_module_Basis_LTO_n.f90:431-440)."""
    r = [evr.Basis_L_TO_n(10, 10, 1), evr.Basis_L_TO_n(1, 2, 1)]
    nq_of, nb_of = evr.level_sizes(2, 6, 7, r)
    assert list(nb_of[0]) == [10, 20, 30, 40, 50, 60, 70, 80]        # HCN golden log
    assert list(nb_of[1]) == [1, 3, 5, 7, 9, 11, 13, 15]
    r2 = [evr.Basis_L_TO_n(1, 1, 2)]
    nq_of, nb_of = evr.level_sizes(1, 2, 4, r2)
    assert list(nq_of[0]) == [1, 2, 5, 10, 17]
    assert list(nb_of[0]) == [1, 2, 5, 8, 11]                        # linear continuation, not the formula
