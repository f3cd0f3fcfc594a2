"""Shared test helpers: run the CPU oracle on the same flat inputs the C-ABI receives."""
import numpy as np

from oracle import sg4_oracle as orc


def flat_op(op):
    b = op.BasisnD
    nb0 = b.nb0
    tm = op.term_mode()
    gz = [g.grid_zero for g in op.OpGrid]
    gc = [g.grid_cte for g in op.OpGrid]
    mc = np.zeros((op.nb_Term, nb0 * nb0))
    grids = []
    for it, g in enumerate(op.OpGrid):
        if g.Mat_cte is not None:
            mc[it] = np.asarray(g.Mat_cte, dtype=np.float64).reshape(nb0, nb0).ravel(order="F")
        if g.grid_zero or g.grid_cte or g.Grid is None:
            grids.append(None)
        else:
            grids.append(np.asfortranarray(np.asarray(g.Grid, dtype=np.float64).reshape(b.nqq, nb0, nb0)).ravel(order="F"))
    return tm, gz, gc, mc, grids


def oracle_apply(op, psi, nthreads=4, iG_range=None):
    """H psi through oracle/sg4_oracle.c on the inputs of ParamOp ``op``; psi[npsi, nb*nb0]."""
    b = op.BasisnD
    tm, gz, gc, mc, grids = flat_op(op)
    lo, hi = (0, b.nb_SG) if iG_range is None else iG_range
    return orc.tab_oppsi(b.D, b.nb_SG, b.nb0, b.nb, b.LG, b.nDind_SmolyakRep_Tab_nDval, b.WeightSG,
                         b.tab_nq_OF_SRep, b.tab_nb_OF_SRep, b.tab_iB_OF_SRep_TO_iB, b.nq_of, b.nb_of,
                         b.B, b.BTw, b.D1, b.D2, op.type_Op, tm, gz, gc, mc, grids, psi,
                         nthreads=nthreads, iG_begin=lo, iG_end=hi)


def oracle_apply10(op, psi, nthreads=4, iG_range=None):
    """type_Op=10 through oracle/sg4_oracle.c on the inputs of a ParamOp10."""
    b = op.BasisnD
    lo, hi = (0, b.nb_SG) if iG_range is None else iG_range
    return orc.tab_oppsi10(b.D, b.nb_SG, b.nb0, b.nb, b.LG, b.nDind_SmolyakRep_Tab_nDval, b.WeightSG,
                           b.tab_nq_OF_SRep, b.tab_nb_OF_SRep, b.tab_iB_OF_SRep_TO_iB, b.nq_of, b.nb_of,
                           b.B, b.BTw, b.D1, op.mode_of_Qact,
                           None if op.V is None else op.V.ravel(order="F"), op.GG.ravel(order="F"), op.Jac, op.sqRhoOVERJac,
                           psi, nthreads=nthreads, iG_begin=lo, iG_end=hi)


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def random_psi(nvec, npsi, seed=12345):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((npsi, nvec))
    return x / np.linalg.norm(x, axis=1, keepdims=True)
