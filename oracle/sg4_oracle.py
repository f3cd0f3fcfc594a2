"""ctypes front-end of the CPU oracle (oracle/sg4_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py; never from the product package.
See the header of sg4_oracle.c for what is restated and how it is pinned.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsg4_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc -O3 -fopenmp)."""
    src = os.path.join(_HERE, "sg4_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        ip = C.POINTER(C.c_int)
        L.orc_tables_build.restype = C.c_void_p
        L.orc_tables_build.argtypes = [C.c_int, C.c_int, C.c_int, ip, ip, ip, ip, ip, ip, C.c_int]
        L.orc_tables_free.argtypes = [C.c_void_p]
        for name, rt in [("D", C.c_int), ("Lmin", C.c_int), ("nb_SG", C.c_int), ("nb", C.c_int64),
                         ("S", C.c_int64), ("NQ", C.c_int64), ("count0", C.c_int64)]:
            f = getattr(L, "orc_tables_" + name)
            f.restype = rt
            f.argtypes = [C.c_void_p]
        for name in ["nq_of", "nb_of", "tab_l", "weight", "tab_nq", "tab_nb", "sum_nq", "sum_nb",
                     "packedB", "map", "i_to_l", "i_to_l_off"]:
            f = getattr(L, "orc_tables_" + name)
            f.restype = C.c_void_p
            f.argtypes = [C.c_void_p]
        L.orc_n_of_L.restype = C.c_int
        L.orc_n_of_L.argtypes = [C.c_int] * 5
        L.orc_tab_oppsi.restype = C.c_int
        L.orc_tab_oppsi10.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    ct = {np.int32: C.c_int32, np.int64: C.c_int64, np.float64: C.c_double}[dtype]
    return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ct)), shape=(n,)).copy()


class Tables:
    """All SG4 integer tables + weights, as numpy arrays with the reference's 1-based values."""

    def __init__(self, D, LB, LG, Aq, Bq, expo_q=None, Ab=None, Bb=None, expo_b=None, legacy_LB0=False):
        L = lib()
        Aq = np.ascontiguousarray(np.broadcast_to(Aq, (D,)), dtype=np.int32)
        Bq = np.ascontiguousarray(np.broadcast_to(Bq, (D,)), dtype=np.int32)
        expo_q = np.ascontiguousarray(np.broadcast_to(1 if expo_q is None else expo_q, (D,)), dtype=np.int32)
        # nb parameters default to the nq ones (sub_read_data.f90:738-739)
        Ab = Aq if Ab is None else np.ascontiguousarray(np.broadcast_to(Ab, (D,)), dtype=np.int32)
        Bb = Bq if Bb is None else np.ascontiguousarray(np.broadcast_to(Bb, (D,)), dtype=np.int32)
        expo_b = expo_q if expo_b is None else np.ascontiguousarray(np.broadcast_to(expo_b, (D,)), dtype=np.int32)
        ip = C.POINTER(C.c_int)
        h = L.orc_tables_build(D, LB, LG, Aq.ctypes.data_as(ip), Bq.ctypes.data_as(ip),
                               expo_q.ctypes.data_as(ip), Ab.ctypes.data_as(ip), Bb.ctypes.data_as(ip),
                               expo_b.ctypes.data_as(ip), int(legacy_LB0))
        if not h:
            raise ValueError("orc_tables_build failed")
        try:
            self.D, self.LB, self.LG = D, LB, LG
            self.Lmin = L.orc_tables_Lmin(h)
            self.nb_SG = L.orc_tables_nb_SG(h)
            self.nb = L.orc_tables_nb(h)
            self.S = L.orc_tables_S(h)
            self.NQ = L.orc_tables_NQ(h)
            self.count0 = L.orc_tables_count0(h)
            n = D * (LG + 1)
            self.nq_of = _arr(L.orc_tables_nq_of(h), n, np.int32).reshape(D, LG + 1)
            self.nb_of = _arr(L.orc_tables_nb_of(h), n, np.int32).reshape(D, LG + 1)
            self.tab_l = _arr(L.orc_tables_tab_l(h), self.nb_SG * D, np.int32).reshape(self.nb_SG, D)
            self.weight = _arr(L.orc_tables_weight(h), self.nb_SG, np.float64)
            self.tab_nq = _arr(L.orc_tables_tab_nq(h), self.nb_SG, np.int32)
            self.tab_nb = _arr(L.orc_tables_tab_nb(h), self.nb_SG, np.int32)
            self.sum_nq = _arr(L.orc_tables_sum_nq(h), self.nb_SG, np.int64)
            self.sum_nb = _arr(L.orc_tables_sum_nb(h), self.nb_SG, np.int64)
            self.packedB = _arr(L.orc_tables_packedB(h), self.nb * D, np.int32).reshape(self.nb, D)
            self.map = _arr(L.orc_tables_map(h), self.S, np.int32)
            off = _arr(L.orc_tables_i_to_l_off(h), D + 1, np.int32)
            flat = _arr(L.orc_tables_i_to_l(h), int(off[-1]), np.int32)
            self.i_to_l = [flat[off[k]:off[k + 1]] for k in range(D)]
        finally:
            L.orc_tables_free(h)


def tab_oppsi(D, nb_SG, nb0, nb, LG, tab_l, W, tab_nq, tab_nb, mapping, nq_of, nb_of,
              B, BTw, D1, D2, type_Op, term_mode, grid_zero, grid_cte, Mat_cte, grids,
              psi, nthreads=1, iG_begin=0, iG_end=None, out=None, zero_out=True):
    """H|psi> for psi[npsi, nb*nb0] (C order: one right-hand side per row). Returns Hpsi same shape."""
    L = lib()
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    if psi.ndim == 1:
        psi = psi[None, :]
    npsi = psi.shape[0]
    assert psi.shape[1] == nb * nb0
    nb_Term = len(grid_zero)
    tab_l = np.ascontiguousarray(tab_l, dtype=np.int32)
    W = np.ascontiguousarray(W, dtype=np.float64)
    tab_nq = np.ascontiguousarray(tab_nq, dtype=np.int32)
    tab_nb = np.ascontiguousarray(tab_nb, dtype=np.int32)
    mapping = np.ascontiguousarray(mapping, dtype=np.int32)
    nq_of = np.ascontiguousarray(nq_of, dtype=np.int32)
    nb_of = np.ascontiguousarray(nb_of, dtype=np.int32)
    B, BTw, D1, D2 = (np.ascontiguousarray(a, dtype=np.float64) for a in (B, BTw, D1, D2))
    term_mode = np.ascontiguousarray(term_mode, dtype=np.int32)
    gz = np.ascontiguousarray(grid_zero, dtype=np.uint8)
    gc = np.ascontiguousarray(grid_cte, dtype=np.uint8)
    Mat_cte = np.ascontiguousarray(Mat_cte, dtype=np.float64)
    keep = []
    ptrs = (C.c_void_p * nb_Term)()
    for i, g in enumerate(grids):
        if g is None:
            ptrs[i] = None
        else:
            g = np.ascontiguousarray(g, dtype=np.float64)
            keep.append(g)
            ptrs[i] = g.ctypes.data
    Hpsi = np.empty_like(psi) if out is None else out
    if iG_end is None:
        iG_end = nb_SG
    vp = lambda a: C.c_void_p(a.ctypes.data)
    rc = L.orc_tab_oppsi(C.c_int(D), C.c_int(nb_SG), C.c_int(nb0), C.c_int64(nb), C.c_int(LG),
                         vp(tab_l), vp(W), vp(tab_nq), vp(tab_nb), vp(mapping), vp(nq_of), vp(nb_of),
                         vp(B), vp(BTw), vp(D1), vp(D2), C.c_int(type_Op), C.c_int(nb_Term), vp(term_mode),
                         vp(gz), vp(gc), vp(Mat_cte), ptrs, C.c_int(npsi), vp(psi), vp(Hpsi),
                         C.c_int(nthreads), C.c_int(iG_begin), C.c_int(iG_end), C.c_int(1 if zero_out else 0))
    if rc != 0:
        raise RuntimeError(f"orc_tab_oppsi failed rc={rc}")
    return Hpsi


def tab_oppsi10(D, nb_SG, nb0, nb, LG, tab_l, W, tab_nq, tab_nb, mapping, nq_of, nb_of, B, BTw, D1,
                act_mode, V, GG, Jac, sq, psi, nthreads=1, iG_begin=0, iG_end=None):
    """type_Op=10 action with cached metric tensor (see orc_tab_oppsi10). psi[npsi, nb*nb0]."""
    L = lib()
    psi = np.ascontiguousarray(psi, dtype=np.float64)
    if psi.ndim == 1:
        psi = psi[None, :]
    npsi = psi.shape[0]
    c32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    c64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    tab_l, tab_nq, tab_nb, mapping, nq_of, nb_of, act_mode = map(c32, (tab_l, tab_nq, tab_nb, mapping, nq_of, nb_of, act_mode))
    W, B, BTw, D1, GG, Jac, sq = map(c64, (W, B, BTw, D1, GG, Jac, sq))
    Vc = None if V is None else c64(V)
    Hpsi = np.empty_like(psi)
    if iG_end is None:
        iG_end = nb_SG
    vp = lambda a: C.c_void_p(a.ctypes.data)
    rc = L.orc_tab_oppsi10(C.c_int(D), C.c_int(nb_SG), C.c_int(nb0), C.c_int64(nb), C.c_int(LG),
                           vp(tab_l), vp(W), vp(tab_nq), vp(tab_nb), vp(mapping), vp(nq_of), vp(nb_of),
                           vp(B), vp(BTw), vp(D1), C.c_int(len(act_mode)), vp(act_mode),
                           C.c_void_p(None) if Vc is None else vp(Vc), vp(GG), vp(Jac), vp(sq),
                           C.c_int(npsi), vp(psi), vp(Hpsi), C.c_int(nthreads), C.c_int(iG_begin), C.c_int(iG_end), C.c_int(1))
    if rc != 0:
        raise RuntimeError(f"orc_tab_oppsi10 failed rc={rc}")
    return Hpsi


def nested(mode, D, nb_SG, nb0, nb, LG, tab_l, W, tab_nq, tab_nb, mapping, nq_of, nb_of, B, BTw, D1, D2, vec, der=(0, 0)):
    """Whole-vector SG4 routines (orc_nested): mode 0 RvecB -> RvecG, 1 RvecG -> RvecB, 2 derivative on the grid.
    vec[nvec, len]; returns the transformed vectors."""
    L = lib()
    L.orc_nested.restype = C.c_int
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    if vec.ndim == 1:
        vec = vec[None, :]
    c32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    c64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    tab_l, tab_nq, tab_nb, mapping, nq_of, nb_of = map(c32, (tab_l, tab_nq, tab_nb, mapping, nq_of, nb_of))
    W, B, BTw, D1, D2 = map(c64, (W, B, BTw, D1, D2))
    NQ = int(tab_nq.sum())
    lenB, lenG = nb * nb0, NQ * nb0
    n_in, n_out = (lenB, lenG) if mode == 0 else ((lenG, lenB) if mode == 1 else (lenG, lenG))
    assert vec.shape[1] == n_in
    out = np.empty((vec.shape[0], n_out))
    vp = lambda a: C.c_void_p(a.ctypes.data)
    for i in range(vec.shape[0]):
        rc = L.orc_nested(C.c_int(mode), C.c_int(D), C.c_int(nb_SG), C.c_int(nb0), C.c_int64(nb), C.c_int(LG),
                          vp(tab_l), vp(W), vp(tab_nq), vp(tab_nb), vp(mapping), vp(nq_of), vp(nb_of),
                          vp(B), vp(BTw), vp(D1), vp(D2), C.c_int(der[0]), C.c_int(der[1]),
                          C.c_void_p(vec[i].ctypes.data), C.c_void_p(out[i].ctypes.data))
        if rc != 0:
            raise RuntimeError(f"orc_nested failed rc={rc}")
    return out


def max_threads() -> int:
    return lib().orc_max_threads()
