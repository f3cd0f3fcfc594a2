"""Set-up of the configurations named in BASELINE.json / SURVEY.md 8(d) (host side).

These functions stand where ElVibRot's input decks + Tnum/PrimOp stand: they produce the
1-D primitives and the operator grids that cross the boundary.  Only closed-form models are
built here (the reference evaluates them through sub_pot/sub_system.f at every grid point,
sub_OpPsi_SG4.f90:2982-3006):

* Henon-Heiles D-dim, unit masses, constant metric (Working_tests/MPI_tests/*_Davidson_*/
  sub_system_HenonHeiles.f:40-47, lambda = 0.111803; ``Hm`` bases nq=nb=1+2L; Gcte=t ->
  Mat_cte = -1/2 G_ii, sub_active/sub_Grid_SG4.f90:114-151)
* pyrazine 12-D two-state vibronic model (UnitTests/PYR-WP_UT_MPI/sub_system_pyrazine.f:30-75,
  calc_f2_f1Q.f90: f2(i,i) = -w_i/2), nb0 = 2
* shape-faithful synthetic curvilinear operators (HCN_UT / HNO3_UT shapes, random grids):
  the physical values need Tnum and cannot be produced without Fortran.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np

from .primitives import hm_primitive
from .sg4 import Basis_L_TO_n, Init_TypeOp, OpGrid, ParamOp, ParamOp10, SG4Basis, level_sizes

LAMBDA_HH = 0.111803


def _au_to_ev_codata2006() -> float:
    """au -> eV exactly as the reference derives it (default version CODATA2006:
    Source_PhysicalConstants/sub_module_constant.f90:260-308, constants :788-798):
    epsi0 = 1/(mhu0 c^2), a0 = (hb/e)^2/me 4 pi epsi0, Eh = e^2/(4 pi epsi0 a0), auTOeV = Eh/e."""
    c, mhu0, h, e, me = 299792458.0, np.pi * 4e-7, 6.62606896e-34, 1.602176487e-19, 9.10938215e-31
    epsi0 = 1.0 / (mhu0 * c * c)
    hb = h / (2.0 * np.pi)
    a0 = (hb / e) ** 2 / me * 4.0 * np.pi * epsi0
    Eh = (e * e / a0) / (4.0 * np.pi * epsi0)
    return Eh / e


EV_TO_AU = 1.0 / _au_to_ev_codata2006()   # 1 / 27.2113838656 (ene_unit='eV' of the pyrazine inputs)


def hm_sg4_basis(D: int, LB: int, LG: int, A, B, nb0: int = 1, Q0=0.0, scaleQ=1.0) -> SG4Basis:
    """SG4 basis of D ``Hm`` modes with nq_k(L) = nb_k(L) = A_k + B_k L."""
    A = np.broadcast_to(A, (D,))
    B = np.broadcast_to(B, (D,))
    Q0 = np.broadcast_to(Q0, (D,))
    sc = np.broadcast_to(scaleQ, (D,))
    rules = [Basis_L_TO_n(int(A[k]), int(B[k]), 1) for k in range(D)]
    nq_of, nb_of = level_sizes(D, LB, LG, rules)
    cache = {}
    prims = []
    for k in range(D):
        row = []
        for L in range(LG + 1):
            key = (int(nq_of[k, L]), int(nb_of[k, L]), float(Q0[k]), float(sc[k]))
            if key not in cache:
                cache[key] = hm_primitive(*key)
            row.append(cache[key])
        prims.append(row)
    return SG4Basis(D, LB, LG, nq_of, nb_of, prims, nb0=nb0)


def _term_outer_sum(basis: SG4Basis, iG: int, f1d, pair=None):
    """sum_k f1d[k](x_k) (+ sum_k pair(x_k, x_{k+1})) on the grid of term iG, first mode fastest."""
    axes = basis.term_grid_axes(iG)
    D = basis.D
    shape = [len(a) for a in axes]
    # build with numpy broadcasting in Fortran order: axis k varies fastest for small k
    V = np.zeros(shape[::-1])           # C array indexed [q_D,...,q_1]  == Fortran (q_1,...,q_D)
    for k in range(D):
        sh = [1] * D
        sh[D - 1 - k] = shape[k]
        V = V + f1d(k, axes[k]).reshape(sh)
        if pair is not None and k + 1 < D and (shape[k] > 1 or shape[k + 1] > 1):
            sh2 = [1] * D
            sh2[D - 1 - k] = shape[k]
            sh2[D - 2 - k] = shape[k + 1]
            V = V + pair(axes[k][None, :], axes[k + 1][:, None]).reshape(sh2)
        elif pair is not None and k + 1 < D:
            V = V + pair(axes[k][0], axes[k + 1][0])
    return V.ravel()


def henon_heiles_potential(basis: SG4Basis, lam: float = LAMBDA_HH) -> np.ndarray:
    """V = 1/2 sum Q_i^2 + lam sum_{i<D} (Q_i^2 Q_{i+1} - Q_{i+1}^3/3) on the whole Smolyak grid."""
    V = np.empty(basis.nqq)
    D = basis.D

    def f1d(k, x):
        v = 0.5 * x * x
        if k >= 1:
            v = v - (lam / 3.0) * x ** 3
        return v

    def pair(xa, xb):
        return lam * xa * xa * xb

    # inactive modes sit at x = 0 where every contribution vanishes -> only active modes matter
    for iG in range(basis.nb_SG):
        l = basis.term_levels(iG)
        act = [k for k in range(D) if basis.nq_of[k, l[k]] > 1]
        sl = basis.term_grid_slice(iG)
        if not act:
            axes = basis.term_grid_axes(iG)
            V[sl] = sum(f1d(k, axes[k])[0] for k in range(D)) + sum(pair(axes[k][0], axes[k + 1][0]) for k in range(D - 1))
            continue
        axes = basis.term_grid_axes(iG)
        if any(len(axes[k]) == 1 and axes[k][0] != 0.0 for k in range(D)):
            V[sl] = _term_outer_sum(basis, iG, f1d, pair)
            continue
        na = len(act)
        shape = [len(axes[k]) for k in act]
        acc = np.zeros(shape[::-1])
        for a, k in enumerate(act):
            sh = [1] * na
            sh[na - 1 - a] = shape[a]
            acc = acc + f1d(k, axes[k]).reshape(sh)
            if a + 1 < na and act[a + 1] == k + 1:
                sh2 = [1] * na
                sh2[na - 1 - a] = shape[a]
                sh2[na - 2 - a] = shape[a + 1]
                acc = acc + pair(axes[k][None, :], axes[k + 1][:, None]).reshape(sh2)
        V[sl] = acc.ravel()
    return V


def constant_keo_opgrids(D: int, nb0: int, Gdiag, V: Optional[np.ndarray]) -> List[OpGrid]:
    """type_Op=1 term list for a constant diagonal metric: (0,0) -> V grid; f2(i,i) constant
    -G_ii/2 on the channel diagonal; every other term grid_zero (sub_Grid_SG4.f90:114-151,
    sub_module_OpGrid.f90:1001-1003)."""
    ops = []
    eye = np.eye(nb0)
    for (i, j) in Init_TypeOp(1, D):
        if (i, j) == (0, 0):
            ops.append(OpGrid((0, 0), grid_zero=V is None, grid_cte=False, Grid=V))
        elif i == j:
            ops.append(OpGrid((i, j), grid_cte=True, Mat_cte=-0.5 * float(Gdiag[i - 1]) * eye))
        else:
            ops.append(OpGrid((i, j), grid_zero=True, grid_cte=True, Mat_cte=np.zeros((nb0, nb0))))
    return ops


def model_potential_device(basis: SG4Basis, model: int, params) -> np.ndarray:
    """Closed-form model potential on the whole Smolyak grid, evaluated on the GPU (C-ABI evr_sg4_model_grid; the
    reference fills this grid point by point during its first H|psi>, sub_OpPsi_SG4.f90:2982-3006)."""
    from . import lib as _lib
    x = np.concatenate([basis.tab_basisPrimSG[k][L].x for k in range(basis.D) for L in range(basis.LG + 1)]).astype(np.float64)
    prm = np.ascontiguousarray(params, dtype=np.float64).ravel()
    V = np.empty(basis.nqq)
    _lib.check(_lib.lib().evr_sg4_model_grid(model, basis.D, basis.nb_SG, basis.LG, basis.nDind_SmolyakRep_Tab_nDval.ctypes.data,
                                             basis.nq_of.ctypes.data, x.ctypes.data, len(prm), prm.ctypes.data, 0, basis.nb_SG,
                                             V.ctypes.data), "evr_sg4_model_grid")
    return V


def henon_heiles(D: int, L: int, LB: Optional[int] = None, iG_range=None, device: int = -1):
    """Henon-Heiles D-dim, SG4 LB=LG=L (LB may differ), Hm nq=nb=1+2L.  Returns (basis, para_H)."""
    LB = L if LB is None else LB
    basis = hm_sg4_basis(D, LB, L, 1, 2)
    V = henon_heiles_potential(basis)
    ops = constant_keo_opgrids(D, 1, np.ones(D), V.reshape(-1, 1, 1))
    return basis, ParamOp(basis, 1, ops, iG_range=iG_range, device=device)


# pyrazine model parameters (eV), sub_system_pyrazine.f:30-44
_PYR_W = np.array([0.09357, 0.0740, 0.1273, 0.1568, 0.1347, 0.3431, 0.1157, 0.3242, 0.3621, 0.2673, 0.3052, 0.0968])
_PYR_K1 = np.array([0.0, -0.0964, 0.0470, 0.1594, 0.0308, 0.0782, 0.0261, 0.0717, 0.0560, 0.0625, 0.0780, 0.0188])
_PYR_K2 = np.array([0.0, 0.1194, 0.2012, 0.0484, -0.0308, -0.0782, -0.0261, -0.0717, -0.0560, -0.0625, -0.0780, -0.0188])
_PYR_DELTA, _PYR_LAMBDA = 0.46165, 0.1825


def pyrazine_12d(L: int = 1, B=(3, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2), iG_range=None, device: int = -1):
    """Pyrazine 12-D, two coupled electronic states (nb0=2), SG4 LB=LG=L, Hm nq=nb=1+B_k L.
    Dimensionless normal coordinates: T = -sum w_i/2 d2/dQ_i2, energies in au."""
    D = 12
    basis = hm_sg4_basis(D, L, L, 1, np.array(B), nb0=2)
    NQ = basis.nqq
    G = np.zeros((NQ, 2, 2), order="F")
    w, k1, k2 = _PYR_W, _PYR_K1, _PYR_K2
    for iG in range(basis.nb_SG):
        sl = basis.term_grid_slice(iG)
        h = _term_outer_sum(basis, iG, lambda k, x: 0.5 * w[k] * x * x)
        l1 = _term_outer_sum(basis, iG, lambda k, x: k1[k] * x)
        l2 = _term_outer_sum(basis, iG, lambda k, x: k2[k] * x)
        q1 = _term_outer_sum(basis, iG, lambda k, x: x if k == 0 else 0.0 * x)
        G[sl, 0, 0] = (h - _PYR_DELTA + l1) * EV_TO_AU
        G[sl, 1, 1] = (h + _PYR_DELTA + l2) * EV_TO_AU
        G[sl, 0, 1] = G[sl, 1, 0] = _PYR_LAMBDA * q1 * EV_TO_AU
    ops = constant_keo_opgrids(D, 2, w * EV_TO_AU, G)
    return basis, ParamOp(basis, 1, ops, iG_range=iG_range, device=device)


def synthetic_curvilinear(basis: SG4Basis, seed: int = 777, iG_range=None, device: int = -1, n_coupled: Optional[int] = None):
    """type_Op=1 operator with ALL (n+1)(n+2)/2 term grids filled with N(0,1) values (shape-faithful
    stand-in for a Tnum curvilinear KEO: HCN_UT n=3 -> 10 grids/pt, HNO3_UT n=8 -> 45+10 grids/pt)."""
    rng = np.random.default_rng(seed)
    D, nb0, NQ = basis.D, basis.nb0, basis.nqq
    ops = []
    for (i, j) in Init_TypeOp(1, D):
        g = np.zeros((NQ, nb0, nb0), order="F")
        if (i, j) == (0, 0):
            g[:] = rng.standard_normal((NQ, nb0, nb0))
        else:
            for c in range(nb0):       # KEO grids are channel-diagonal (sub_OpPsi_SG4.f90:2994-2999)
                g[:, c, c] = rng.standard_normal(NQ)
        ops.append(OpGrid((i, j), Grid=g))
    return ParamOp(basis, 1, ops, iG_range=iG_range, device=device)


def synthetic_type10(basis: SG4Basis, seed: int = 777, with_V: bool = True, iG_range=None, device: int = -1):
    """type_Op=10 operator with a synthetic, smoothly varying positive-definite metric tensor G(Q), Jacobian and
    sqrt(rho/Jac) per grid point (shape-faithful stand-in for Tnum's get_d0GG output)."""
    rng = np.random.default_rng(seed)
    n, nb0, NQ = basis.D, basis.nb0, basis.nqq
    A = 0.3 * rng.standard_normal((NQ, n, n))
    GG = np.einsum("qij,qkj->qik", A, A) + np.eye(n)[None, :, :]          # SPD, symmetric
    Jac = 1.0 + rng.random(NQ)
    sq = 0.5 + rng.random(NQ)
    V = None
    if with_V:
        V = rng.standard_normal((NQ, nb0, nb0))
        V = 0.5 * (V + V.transpose(0, 2, 1))
    return ParamOp10(basis, np.asfortranarray(GG), Jac, sq, V=None if V is None else np.asfortranarray(V),
                     iG_range=iG_range, device=device)
