"""Host-side mirror of the reference interface around the SG4 H|psi> action.

Names, argument meaning and error behaviour follow the Fortran they stand for, so that a
driver written against ElVibRot's ``mod_OpPsi`` reads the same here:

=========================  ==================================================================
here                       reference
=========================  ==================================================================
``Basis_L_TO_n``           TYPE Basis_L_TO_n, sub_Basis/sub_module_Basis_LTO_n.f90:36-60, 267-440
``SG4Basis``               TYPE basis with SparseGrid_type=4 + TYPE param_SGType2
                           (sub_module_basis_set_alloc.f90:204-208, sub_module_param_SGType2.f90:54-102)
``Init_TypeOp``            Source_PrimOperator/sub_module_SimpleOp.f90:256-375 (term numbering)
``OpGrid``/``ParamOp``     TYPE param_OpGrid / param_Op (sub_Operator/sub_module_OpGrid.f90, sub_module_SetOp.f90)
``ParamPsi``               TYPE param_psi (sub_WP/sub_module_psi_set_alloc.f90): RvecB / CvecB, cplx
``sub_TabOpPsi_FOR_SGtype4``  sub_Operator/sub_OpPsi_SG4.f90:678-979
``sub_TabOpPsi``/``sub_OpPsi``  sub_Operator/sub_OpPsi.f90:701-883 / :175-417 (SG4 branch only)
=========================  ==================================================================

All numerical work goes through the C-ABI (``lib.py`` -> libevr_sg4.so -> CUDA kernels).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import lib as _lib
from .primitives import Primitive1D, concat_tables


# --------------------------------------------------------------------------------------
@dataclass
class Basis_L_TO_n:
    """n(L) = A + B*L**expo (L_TO_n_type = 0)."""
    A: int = 1
    B: int = 1
    expo: int = 1

    def table(self, Lmax: int):
        return [self.A + self.B * L ** self.expo for L in range(Lmax + 1)]

    def get_n(self, L: int, Ltab: Optional[int] = None) -> int:
        """Get_n_FROM_Basis_L_TO_n (:431-440): the table is tabulated to ``Ltab`` and extended
        linearly with its last increment beyond (no extension when Ltab == 0)."""
        if Ltab is None or L <= Ltab:
            return self.A + self.B * L ** self.expo
        tab = self.table(Ltab)
        n = tab[Ltab]
        if Ltab > 0:
            n += (L - Ltab) * (tab[Ltab] - tab[Ltab - 1])
        return n


def level_sizes(D, LB, LG, L_TO_nq: Sequence[Basis_L_TO_n], L_TO_nb: Optional[Sequence[Basis_L_TO_n]] = None):
    """nq_k(L), nb_k(L) of tab_basisPrimSG(L,k) (sub_quadra_SparseBasis.f90:1041-1116):
    nq from a table to L; nb from a table to min(L,LB) (+ linear extension)."""
    L_TO_nb = L_TO_nq if L_TO_nb is None else L_TO_nb
    nq_of = np.zeros((D, LG + 1), dtype=np.int32)
    nb_of = np.zeros((D, LG + 1), dtype=np.int32)
    for k in range(D):
        for L in range(LG + 1):
            nq_of[k, L] = L_TO_nq[k].get_n(L, L)
            nb_of[k, L] = L_TO_nb[k].get_n(L, min(L, LB))
    return nq_of, nb_of


def _get(h, what, dtype):
    L = _lib.lib()
    n = L.evr_sg4_tables_size(h, what)
    a = np.empty(int(n), dtype=dtype)
    if n:
        _lib.check(L.evr_sg4_tables_get(h, what, a.ctypes.data), "evr_sg4_tables_get")
    return a


class SG4Basis:
    """The SG4 basis object: per-level 1-D primitives + the integer tables of param_SGType2.

    ``prims[k][L]`` is the Primitive1D of mode k at level L (tab_basisPrimSG(L,k)).
    """

    def __init__(self, D: int, LB: int, LG: int, nq_of, nb_of, prims, nb0: int = 1):
        self.nb_basis = self.D = D
        self.L_SparseBasis = self.LB = LB
        self.L_SparseGrid = self.LG = LG
        self.SparseGrid_type = 4
        self.nb0 = nb0
        self.nq_of = np.ascontiguousarray(nq_of, dtype=np.int32)
        self.nb_of = np.ascontiguousarray(nb_of, dtype=np.int32)
        self.tab_basisPrimSG = prims
        for k in range(D):
            for L in range(LG + 1):
                pr = prims[k][L]
                if pr.nq != self.nq_of[k, L] or pr.nb != self.nb_of[k, L]:
                    raise ValueError(f"primitive (mode {k}, level {L}) has nq,nb={pr.nq},{pr.nb}, "
                                     f"tables say {self.nq_of[k, L]},{self.nb_of[k, L]}")
        L_ = _lib.lib()
        h = C.c_void_p()
        _lib.check(L_.evr_sg4_tables_build(C.byref(h), D, LB, LG, self.nq_of.ctypes.data, self.nb_of.ctypes.data),
                   "evr_sg4_tables_build")
        try:
            self.Lmin = int(L_.evr_sg4_tables_size(h, _lib.TAB_LMIN))
            self.nb_SG = int(L_.evr_sg4_tables_size(h, _lib.TAB_NB_SG))
            self.nb = int(L_.evr_sg4_tables_size(h, _lib.TAB_NB))
            self.Max_Srep = int(L_.evr_sg4_tables_size(h, _lib.TAB_S))
            self.nqq = int(L_.evr_sg4_tables_size(h, _lib.TAB_NQ))
            self.count0 = int(L_.evr_sg4_tables_size(h, _lib.TAB_COUNT0))
            self.nDind_SmolyakRep_Tab_nDval = _get(h, _lib.TAB_TAB_L, np.int32).reshape(self.nb_SG, D)
            self.WeightSG = _get(h, _lib.TAB_WEIGHT, np.float64)
            self.tab_nq_OF_SRep = _get(h, _lib.TAB_TAB_NQ, np.int32)
            self.tab_nb_OF_SRep = _get(h, _lib.TAB_TAB_NB, np.int32)
            self.tab_Sum_nq_OF_SRep = _get(h, _lib.TAB_SUM_NQ, np.int64)
            self.tab_Sum_nb_OF_SRep = _get(h, _lib.TAB_SUM_NB, np.int64)
            self.nDindB_Tab_nDval = _get(h, _lib.TAB_PACKEDB, np.int32).reshape(self.nb, D)
            self.tab_iB_OF_SRep_TO_iB = _get(h, _lib.TAB_MAP, np.int32)
        finally:
            L_.evr_sg4_tables_destroy(C.byref(h))
        self.B, self.BTw, self.D1, self.D2 = concat_tables(prims, D, LG)

    # grid helpers (set-up side; Rec_Qact_SG4_with_Tab_iq, sub_module_basis.f90)
    def term_levels(self, iG: int):
        return self.nDind_SmolyakRep_Tab_nDval[iG]

    def term_grid_axes(self, iG: int):
        """1-D grid points of every mode for term iG (first mode fastest on the term grid)."""
        l = self.term_levels(iG)
        return [self.tab_basisPrimSG[k][int(l[k])].x for k in range(self.D)]

    def term_grid_slice(self, iG: int) -> slice:
        e = int(self.tab_Sum_nq_OF_SRep[iG])
        return slice(e - int(self.tab_nq_OF_SRep[iG]), e)


# --------------------------------------------------------------------------------------
def Init_TypeOp(type_Op: int, nb_Qact: int):
    """derive_termQact(:,iterm) in the reference's order: type 0 -> [(0,0)];
    type 1 -> (0,0); f2 (i,j) i=1..n, j=i..n; f1 (i,0)."""
    if type_Op == 0:
        return [(0, 0)]
    if type_Op == 1:
        terms = [(0, 0)]
        for i in range(1, nb_Qact + 1):
            for j in range(i, nb_Qact + 1):
                terms.append((i, j))
        for i in range(1, nb_Qact + 1):
            terms.append((i, 0))
        return terms
    raise NotImplementedError("type_Op must be 0 or 1 (type_Op=10 is a later row of SURVEY.md 8f)")


@dataclass
class OpGrid:
    """One operator term: OpGrid(iterm) of the reference (grid_zero / grid_cte / Mat_cte / Grid)."""
    derive_termQact: tuple = (0, 0)
    grid_zero: bool = False
    grid_cte: bool = False
    Mat_cte: Optional[np.ndarray] = None   # (nb0,nb0)
    Grid: Optional[np.ndarray] = None      # (NQ,nb0,nb0) Fortran order, whole Smolyak grid


class SG4Transforms:
    """Whole-vector transforms of an SG4 basis (nested-SG4 entry): the SparseGrid_type = 4 branches of the reference's
    RecRvecB_TO_RVecG / RecRVecG_TO_RvecB / DerivOp_TO_RVecG (sub_Basis/sub_module_basis_BtoG_GtoB.f90:831-847, 252-273,
    1394-1416), batched over the rows of the arguments (the outer index when the SG4 basis sits inside a direct product)."""

    def __init__(self, BasisnD: "SG4Basis", device: int = -1):
        self.BasisnD = b = BasisnD
        self._plan = C.c_void_p()
        _lib.check(_lib.lib().evr_sg4_plan_create(
            C.byref(self._plan), device, b.D, b.nb_SG, b.nb0, b.nb, b.LG,
            b.nDind_SmolyakRep_Tab_nDval.ctypes.data, b.WeightSG.ctypes.data,
            b.tab_nq_OF_SRep.ctypes.data, b.tab_nb_OF_SRep.ctypes.data, b.tab_iB_OF_SRep_TO_iB.ctypes.data,
            b.nq_of.ctypes.data, b.nb_of.ctypes.data,
            b.B.ctypes.data, b.BTw.ctypes.data, b.D1.ctypes.data, b.D2.ctypes.data, 0, b.nb_SG), "evr_sg4_plan_create")

    def _rows(self, v, n):
        v = np.ascontiguousarray(v, dtype=np.float64)
        one = v.ndim == 1
        v = v[None, :] if one else v
        if v.shape[1] != n:
            raise ValueError(f"vector has {v.shape[1]} entries, expected {n}")
        return v, one

    def RvecB_TO_RvecG(self, RvecB):
        b = self.BasisnD
        x, one = self._rows(RvecB, b.nb * b.nb0)
        y = np.empty((x.shape[0], b.nqq * b.nb0))
        _lib.check(_lib.lib().evr_sg4_BtoG(self._plan, x.shape[0], x.ctypes.data, y.ctypes.data), "evr_sg4_BtoG")
        return y[0] if one else y

    def RvecG_TO_RvecB(self, RvecG):
        b = self.BasisnD
        x, one = self._rows(RvecG, b.nqq * b.nb0)
        y = np.empty((x.shape[0], b.nb * b.nb0))
        _lib.check(_lib.lib().evr_sg4_GtoB(self._plan, x.shape[0], x.ctypes.data, y.ctypes.data), "evr_sg4_GtoB")
        return y[0] if one else y

    def DerivOp_TO_RvecG(self, RvecG, mode1: int, mode2: int = 0):
        b = self.BasisnD
        x, one = self._rows(RvecG, b.nqq * b.nb0)
        y = np.empty_like(x)
        _lib.check(_lib.lib().evr_sg4_DerivOp_G(self._plan, x.shape[0], x.ctypes.data, y.ctypes.data, int(mode1), int(mode2)),
                   "evr_sg4_DerivOp_G")
        return y[0] if one else y

    def close(self):
        if self._plan:
            _lib.lib().evr_sg4_plan_destroy(C.byref(self._plan))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ParamOp:
    """param_Op restricted to what the SG4 action reads; owns the device plan (a pure cache)."""

    def __init__(self, BasisnD: SG4Basis, type_Op: int, OpGrids: List[OpGrid],
                 mode_of_Qact: Optional[Sequence[int]] = None, iG_range=None, device: int = -1):
        self.BasisnD = BasisnD
        self.type_Op = type_Op
        self.OpGrid = OpGrids
        self.nb_Term = len(OpGrids)
        self.nb_OpPsi = 0
        self.cplx = False
        self.device = device
        D = BasisnD.D
        # Tabder_Qdyn_TO_Qbasis flattened: active coordinate i (1-based) lives in SG4 mode mode_of_Qact[i-1]
        self.mode_of_Qact = list(range(1, D + 1)) if mode_of_Qact is None else list(mode_of_Qact)
        self.iG_range = (0, BasisnD.nb_SG) if iG_range is None else tuple(iG_range)
        # para_ReadOp%Op_Transfo / E0_Transfo (sub_OpPsi.f90:768-775): with TransfoOp the action is (H - E0)^2
        self.Op_Transfo = False
        self.E0_Transfo = 0.0
        self._plan = C.c_void_p()
        self._keep = []

    # -- device cache ---------------------------------------------------------------
    def term_mode(self):
        tm = np.zeros((self.nb_Term, 2), dtype=np.int32)
        for it, og in enumerate(self.OpGrid):
            for s in range(2):
                q = og.derive_termQact[s]
                tm[it, s] = self.mode_of_Qact[q - 1] if q > 0 else 0
        return tm

    def _ensure_plan(self):
        if self._plan:
            return
        b = self.BasisnD
        L = _lib.lib()
        _lib.check(L.evr_sg4_plan_create(
            C.byref(self._plan), self.device, b.D, b.nb_SG, b.nb0, b.nb, b.LG,
            b.nDind_SmolyakRep_Tab_nDval.ctypes.data, b.WeightSG.ctypes.data,
            b.tab_nq_OF_SRep.ctypes.data, b.tab_nb_OF_SRep.ctypes.data, b.tab_iB_OF_SRep_TO_iB.ctypes.data,
            b.nq_of.ctypes.data, b.nb_of.ctypes.data,
            b.B.ctypes.data, b.BTw.ctypes.data, b.D1.ctypes.data, b.D2.ctypes.data,
            int(self.iG_range[0]), int(self.iG_range[1])), "evr_sg4_plan_create")
        nb0 = b.nb0
        tm = self.term_mode()
        gz = np.array([og.grid_zero for og in self.OpGrid], dtype=np.uint8)
        gc = np.array([og.grid_cte for og in self.OpGrid], dtype=np.uint8)
        mc = np.zeros((self.nb_Term, nb0 * nb0))
        ptrs = (C.c_void_p * self.nb_Term)()
        keep = []
        for it, og in enumerate(self.OpGrid):
            if og.Mat_cte is not None:
                mc[it] = np.asarray(og.Mat_cte, dtype=np.float64).reshape(nb0, nb0).ravel(order="F")
            if og.Grid is not None and not (og.grid_zero or og.grid_cte):
                g = np.asarray(og.Grid, dtype=np.float64)
                g = g.reshape(b.nqq, nb0, nb0) if g.ndim != 3 else g
                g = np.asfortranarray(g)
                keep.append(g)
                ptrs[it] = g.ctypes.data
            else:
                ptrs[it] = None
        _lib.check(L.evr_sg4_plan_set_op(self._plan, self.type_Op, self.nb_Term, tm.ctypes.data,
                                         gz.ctypes.data, gc.ctypes.data, mc.ctypes.data, ptrs),
                   "evr_sg4_plan_set_op")

    def plan(self):
        self._ensure_plan()
        return self._plan

    def info(self, what: int) -> int:
        return int(_lib.lib().evr_sg4_plan_info(self.plan(), what))

    def close(self):
        if self._plan:
            _lib.lib().evr_sg4_plan_destroy(C.byref(self._plan))

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- raw entry points -------------------------------------------------------------
    def apply_host(self, psi: np.ndarray, out: Optional[np.ndarray] = None) -> np.ndarray:
        """psi[npsi, nb*nb0] float64 C-contiguous (one RvecB per row) -> H psi (host buffers)."""
        psi = np.ascontiguousarray(psi, dtype=np.float64)
        one = psi.ndim == 1
        if one:
            psi = psi[None, :]
        n = self.BasisnD.nb * self.BasisnD.nb0
        if psi.shape[1] != n:
            raise ValueError(f"psi has {psi.shape[1]} coefficients, basis has {n}")
        if out is None:
            out = np.empty_like(psi)
        elif not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous and out.size == psi.size):
            raise ValueError("out must be a C-contiguous float64 array with as many elements as psi")
        _lib.check(_lib.lib().evr_sg4_apply(self.plan(), psi.shape[0], psi.ctypes.data, out.ctypes.data), "evr_sg4_apply")
        self.nb_OpPsi += psi.shape[0]
        return (out[0] if out.ndim == 2 else out) if one else out

    def apply_device_ptr(self, npsi: int, d_psi: int, d_Hpsi: int, stream: int = 0):
        """Device-resident entry: raw device pointers (e.g. torch ``tensor.data_ptr()``) + CUDA stream."""
        _lib.check(_lib.lib().evr_sg4_apply_device(self.plan(), npsi, d_psi, d_Hpsi, stream), "evr_sg4_apply_device")
        self.nb_OpPsi += npsi

    def apply_device_scaled_ptr(self, npsi: int, d_psi: int, d_Hpsi: int, E0: float, Esc: float, stream: int = 0):
        """Device-resident H|psi> + sub_scaledOpPsi in one call: Hpsi <- (H psi - E0 psi)/Esc (sub_OpPsi.f90:2823-2866)."""
        _lib.check(_lib.lib().evr_sg4_apply_device_scaled(self.plan(), npsi, d_psi, d_Hpsi, float(E0), float(Esc), stream),
                   "evr_sg4_apply_device_scaled")
        self.nb_OpPsi += npsi


class ParamOp10(ParamOp):
    """param_Op with type_Op = 10 and the metric tensor cached per grid point:
    H = -1/2 (Jac sq)^-1 sum_i d_i [ Jac sum_j GG(:,j,i) d_j sq ] + V   (sub_OpPsi_SG4.f90:1548-1650).
    GG(NQ,n,n), Jac(NQ), sqRhoOVERJac(NQ), V(NQ,nb0,nb0) or None: whole Smolyak grid, Fortran order."""

    def __init__(self, BasisnD: SG4Basis, GG, Jac, sqRhoOVERJac, V=None, mode_of_Qact=None, iG_range=None, device: int = -1):
        super().__init__(BasisnD, 10, [], mode_of_Qact=mode_of_Qact, iG_range=iG_range, device=device)
        n = len(self.mode_of_Qact)
        self.nb_act1 = n
        self.GG = np.asfortranarray(np.asarray(GG, dtype=np.float64).reshape(BasisnD.nqq, n, n))
        self.Jac = np.ascontiguousarray(Jac, dtype=np.float64)
        self.sqRhoOVERJac = np.ascontiguousarray(sqRhoOVERJac, dtype=np.float64)
        nb0 = BasisnD.nb0
        self.V = None if V is None else np.asfortranarray(np.asarray(V, dtype=np.float64).reshape(BasisnD.nqq, nb0, nb0))

    def _ensure_plan(self):
        if self._plan:
            return
        b = self.BasisnD
        L = _lib.lib()
        _lib.check(L.evr_sg4_plan_create(
            C.byref(self._plan), self.device, b.D, b.nb_SG, b.nb0, b.nb, b.LG,
            b.nDind_SmolyakRep_Tab_nDval.ctypes.data, b.WeightSG.ctypes.data,
            b.tab_nq_OF_SRep.ctypes.data, b.tab_nb_OF_SRep.ctypes.data, b.tab_iB_OF_SRep_TO_iB.ctypes.data,
            b.nq_of.ctypes.data, b.nb_of.ctypes.data,
            b.B.ctypes.data, b.BTw.ctypes.data, b.D1.ctypes.data, b.D2.ctypes.data,
            int(self.iG_range[0]), int(self.iG_range[1])), "evr_sg4_plan_create")
        am = np.ascontiguousarray(self.mode_of_Qact, dtype=np.int32)
        _lib.check(L.evr_sg4_plan_set_op10(self._plan, len(am), am.ctypes.data,
                                           None if self.V is None else self.V.ctypes.data,
                                           self.GG.ctypes.data, self.Jac.ctypes.data, self.sqRhoOVERJac.ctypes.data),
                   "evr_sg4_plan_set_op10")


@dataclass
class ParamPsi:
    """param_psi: packed basis representation, real (RvecB) or complex (CvecB)."""
    RvecB: Optional[np.ndarray] = None
    CvecB: Optional[np.ndarray] = None
    cplx: bool = False
    symab: int = -1

    @staticmethod
    def real(v):
        return ParamPsi(RvecB=np.array(v, dtype=np.float64), cplx=False)

    @staticmethod
    def complex(v):
        return ParamPsi(CvecB=np.array(v, dtype=np.complex128), cplx=True)


class EvrStop(RuntimeError):
    """The reference STOPs with a message; the mirror raises."""


def sub_TabOpPsi_FOR_SGtype4(Psi: List[ParamPsi], OpPsi: List[ParamPsi], para_Op: ParamOp):
    """OpPsi(:) = H Psi(:) for real psi on the SG4 grid (sub_OpPsi_SG4.f90:678-979)."""
    if len(Psi) == 0:
        raise EvrStop("ERROR in sub_TabOpPsi_FOR_SGtype4: size(Psi) = 0")
    if Psi[0].cplx:
        raise EvrStop("ERROR in sub_TabOpPsi_FOR_SGtype4: Psi(1) is complex")
    x = np.stack([p.RvecB for p in Psi])
    y = para_Op.apply_host(x)
    del OpPsi[:]
    for i, p in enumerate(Psi):
        OpPsi.append(ParamPsi(RvecB=y[i].copy(), cplx=False, symab=p.symab))


def sub_TabOpPsi(TabPsi: List[ParamPsi], TabOpPsi: List[ParamPsi], para_Op: ParamOp, TransfoOp: bool = False):
    """sub_OpPsi.f90:701 -> sub_PrimTabOpPsi :797 (SG4 branch :873-879). Real psi only, like the
    reference's SG4 branch; complex vectors go one by one through sub_OpPsi.
    ``TransfoOp`` with ``para_Op.Op_Transfo`` (:768-775): OpPsi = (H - E0_Transfo)(H - E0_Transfo) Psi, vector by vector --
    two actions, each followed by sub_scaledOpPsi(.., E0_Transfo, ONE)."""
    if TransfoOp and para_Op.Op_Transfo:
        out = []
        for p in TabPsi:
            tmp, o = ParamPsi(), ParamPsi()
            sub_OpPsi(p, tmp, para_Op)
            sub_scaledOpPsi(p, tmp, para_Op.E0_Transfo, 1.0)
            sub_OpPsi(tmp, o, para_Op)
            sub_scaledOpPsi(tmp, o, para_Op.E0_Transfo, 1.0)
            out.append(o)
        TabOpPsi[:] = out
        return
    if any(p.cplx for p in TabPsi):
        out = []
        for p in TabPsi:
            o = ParamPsi()
            sub_OpPsi(p, o, para_Op)
            out.append(o)
        TabOpPsi[:] = out
        return
    sub_TabOpPsi_FOR_SGtype4(TabPsi, TabOpPsi, para_Op)


def sub_OpPsi(Psi: ParamPsi, OpPsi: ParamPsi, para_Op: ParamOp):
    """sub_OpPsi.f90:175 -> sub_PrimOpPsi :271.  Complex psi = two real right-hand sides
    (RCPsi = Psi, :392-407), recombined afterwards."""
    if Psi.cplx:
        RC = [ParamPsi.real(Psi.CvecB.real), ParamPsi.real(Psi.CvecB.imag)]
        RCO: List[ParamPsi] = []
        sub_TabOpPsi_FOR_SGtype4(RC, RCO, para_Op)
        OpPsi.CvecB = RCO[0].RvecB + 1j * RCO[1].RvecB
        OpPsi.RvecB = None
        OpPsi.cplx = True
    else:
        O: List[ParamPsi] = []
        sub_TabOpPsi_FOR_SGtype4([Psi], O, para_Op)
        OpPsi.RvecB = O[0].RvecB
        OpPsi.CvecB = None
        OpPsi.cplx = False
    OpPsi.symab = Psi.symab


def sub_scaledOpPsi(Psi: ParamPsi, OpPsi: ParamPsi, E0: float, Esc: float):
    """OpPsi <- (OpPsi - E0*Psi)/Esc (sub_OpPsi.f90:2823-2866)."""
    if Psi.cplx:
        OpPsi.CvecB = (OpPsi.CvecB - E0 * Psi.CvecB) / Esc
    else:
        OpPsi.RvecB = (OpPsi.RvecB - E0 * Psi.RvecB) / Esc
