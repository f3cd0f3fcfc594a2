"""Nested-SG4 entry: whole-vector RvecB <-> RvecG transforms and the grid derivative of an SG4 basis.
CPU: the oracle's restatement of the three routines is tied to its (reference-pinned) H|psi> -- applying an operator
through B->G, derivative + pointwise multiply, G->B reproduces orc_tab_oppsi.  GPU: the CUDA kernels against the oracle on
the HNO3_UT inner-basis shapes (tables pinned by the reference's logs), batched over the outer index (51 basis functions /
81 grid points of the torsion)."""
import numpy as np
import pytest

from helpers import flat_op, oracle_apply, rel_l2
from oracle import sg4_oracle as orc

TOL = 1e-12


def _orc_nested(mode, b, vec, der=(0, 0)):
    return orc.nested(mode, b.D, b.nb_SG, b.nb0, b.nb, b.LG, b.nDind_SmolyakRep_Tab_nDval, b.WeightSG, b.tab_nq_OF_SRep,
                      b.tab_nb_OF_SRep, b.tab_iB_OF_SRep_TO_iB, b.nq_of, b.nb_of, b.B, b.BTw, b.D1, b.D2, vec, der)


def _apply_through_nested(op, psi):
    """H psi = G->B [ sum_iterm F_iterm(Q) d^(i,j) B->G psi ]   (type_Op = 1, nb0 = 1)."""
    b = op.BasisnD
    tm, gz, gc, mc, grids = flat_op(op)
    g = _orc_nested(0, b, psi)
    acc = np.zeros_like(g)
    for it in range(op.nb_Term):
        if gz[it]:
            continue
        d = g if (tm[it, 0] == 0 and tm[it, 1] == 0) else _orc_nested(2, b, g, (int(tm[it, 0]), int(tm[it, 1])))
        F = mc[it][0] if gc[it] else grids[it]
        acc += F * d
    return _orc_nested(1, b, acc)


@pytest.mark.parametrize("case", ["hh4d", "hno3_lb2_lg4", "hcn_lb4_lg5"])
def test_oracle_whole_vector_routines_compose_to_the_oracle_action(case, evr):
    if case == "hh4d":
        basis, op = evr.workloads.henon_heiles(4, 3)
    elif case == "hno3_lb2_lg4":
        basis = evr.workloads.hm_sg4_basis(8, 2, 4, 1, 1)
        op = evr.workloads.synthetic_curvilinear(basis)
    else:
        basis = evr.workloads.hm_sg4_basis(3, 4, 5, [10, 1, 1], [10, 2, 2])
        op = evr.workloads.synthetic_curvilinear(basis)
    psi = np.random.default_rng(2).standard_normal((2, basis.nb))
    assert rel_l2(_apply_through_nested(op, psi), oracle_apply(op, psi)) < 1e-13


def test_oracle_gtob_inverts_btog_on_the_basis_space(evr):
    """Sum_iG W(iG) GtoB_iG BtoG_iG = identity on the packed basis when nq = nb and LB = LG (Smolyak weights sum to one)."""
    basis, _ = evr.workloads.henon_heiles(5, 3)
    x = np.random.default_rng(3).standard_normal((1, basis.nb))
    assert rel_l2(_orc_nested(1, basis, _orc_nested(0, basis, x)), x) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("shape", ["hno3_lb2_lg4_x51", "hno3_lb3_lg5_x81", "hcn_lb6_lg7", "two_channels", "hh12d_L4", "beyond_smem"])
def test_gpu_whole_vector_routines_match_oracle(shape, evr):
    rng = np.random.default_rng(11)
    if shape == "hno3_lb2_lg4_x51":
        basis, nvec = evr.workloads.hm_sg4_basis(8, 2, 4, 1, 1), 51
    elif shape == "hno3_lb3_lg5_x81":
        basis, nvec = evr.workloads.hm_sg4_basis(8, 3, 5, 1, 1), 81
    elif shape == "hcn_lb6_lg7":
        basis, nvec = evr.workloads.hm_sg4_basis(3, 6, 7, [10, 1, 1], [10, 2, 2]), 4
    elif shape == "two_channels":
        basis, nvec = evr.workloads.hm_sg4_basis(4, 3, 3, 1, 2, nb0=2), 3
    elif shape == "beyond_smem":                         # 20^3 points x 2 channels: work buffers in global memory
        basis, nvec = evr.workloads.hm_sg4_basis(3, 3, 3, 1, [19, 19, 19], nb0=2), 2
        assert basis.tab_nq_OF_SRep.max() * 2 * 16 > 227 * 1024
    else:
        basis, nvec = evr.workloads.hm_sg4_basis(12, 4, 4, 1, 2), 2
    tr = evr.SG4Transforms(basis)
    xb = rng.standard_normal((nvec, basis.nb * basis.nb0))
    g_ref = _orc_nested(0, basis, xb)
    g = tr.RvecB_TO_RvecG(xb)
    assert g.shape == (nvec, basis.nqq * basis.nb0)
    assert rel_l2(g, g_ref) < TOL
    xg = rng.standard_normal((nvec, basis.nqq * basis.nb0))
    assert rel_l2(tr.RvecG_TO_RvecB(xg), _orc_nested(1, basis, xg)) < TOL
    D = basis.D
    for der in [(1, 0), (0, D), (2, 2), (1, D), (D, 1), (0, 0)]:
        assert rel_l2(tr.DerivOp_TO_RvecG(xg, *der), _orc_nested(2, basis, xg, der)) < TOL, der
    one = tr.RvecB_TO_RvecG(xb[0])                       # 1-D in, 1-D out
    assert one.ndim == 1 and rel_l2(one, g_ref[0]) < TOL
    tr.close()
