"""1-D primitive tables that cross the drop-in boundary (host set-up side).

In ElVibRot these tables are built by the Fortran basis builder and handed to the
SG4 action as ``tab_basisPrimSG(L,k)%dnRGB / dnRBGwrho / dnRGG``.
Drivers, tests and bench.py of this repo need the same tables without Fortran, so
this module restates the *definitions*:

* ``Hm`` (harmonic-oscillator) primitive: Gauss-Hermite grid + normalised Hermite
  functions and their analytic derivatives
  (ref: Source_ElVibRot/sub_Basis/sub_quadra_herm.f90:208-410,
        Source_Lib/sub_communf90/sub_math/sub_polyortho.f90:1060-1122);
* ``dnRBGwrho%d0 = B^T diag(w rho)`` (ref: sub_module_basis.f90:2139-2145);
* grid->grid derivative matrices ``dnRGG%d1/d2 = dB . pinv(B)`` with the
  pseudo-inverse through the eigen-decomposition of ``B B^T``
  (ref: sub_dnGB_TO_dnGG, sub_module_basis.f90:2195-2460).

Everything is float64, Fortran (column-major) order.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class Primitive1D:
    """One level of one SG4 mode: the tables the H|psi> kernel consumes."""
    nq: int
    nb: int
    x: np.ndarray      # (nq,)   grid points
    w: np.ndarray      # (nq,)   weights (w*rho)
    B: np.ndarray      # (nq,nb) dnRGB%d0
    BTw: np.ndarray    # (nb,nq) dnRBGwrho%d0
    D1: np.ndarray     # (nq,nq) dnRGG%d1(:,:,1)
    D2: np.ndarray     # (nq,nq) dnRGG%d2(:,:,1,1)


def gauss_hermite(nq: int):
    """Nodes x_i and weights w_i*exp(x_i^2) of the nq-point Gauss-Hermite rule.

    The reference reads them from Internal_data/HermQuadra/herm{nq}.txt (column 2 and
    column 4); tests/test_primitives.py checks this routine against those tables
    (tests/golden/herm_quadra.json) to 1e-13.
    """
    x, w = np.polynomial.hermite.hermgauss(nq)
    # symmetrise: the tabulated rules are exactly antisymmetric with a hard 0 in the middle
    x = 0.5 * (x - x[::-1])
    w = 0.5 * (w + w[::-1])
    return x, w * np.exp(x * x)


def hermite_functions(x: np.ndarray, nb: int):
    """d0,d1,d2[nq,nb] of the normalised Hermite functions phi_l(x), l = 0..nb-1.

    ref: d0d1d2poly_Hermite_exp (sub_polyortho.f90:1060-1122):
      P_l normalised polynomial, P_l' = sqrt(2l) P_{l-1}, P_l'' = 2 (x P_l' - l P_l),
      phi = P e^{-x^2/2}; phi' = (P' - xP) e; phi'' = (P'' - 2xP' + (x^2-1)P) e.
    """
    nq = x.shape[0]
    P = np.zeros((nq, nb))
    P[:, 0] = np.pi ** -0.25
    if nb > 1:
        P[:, 1] = np.sqrt(2.0) * x * P[:, 0]
    for l in range(1, nb - 1):
        P[:, l + 1] = np.sqrt(2.0 / (l + 1)) * x * P[:, l] - np.sqrt(l / (l + 1.0)) * P[:, l - 1]
    dP = np.zeros_like(P)
    d2P = np.zeros_like(P)
    for l in range(1, nb):
        dP[:, l] = np.sqrt(2.0 * l) * P[:, l - 1]
        if l >= 2:
            d2P[:, l] = 2.0 * (x * dP[:, l] - l * P[:, l])
    e = np.exp(-0.5 * x * x)[:, None]
    xx = x[:, None]
    d2 = (d2P - 2.0 * xx * dP + (xx * xx - 1.0) * P) * e
    d1 = (dP - xx * P) * e
    d0 = P * e
    return d0, d1, d2


def grid_derivative_matrices(B, dB, d2B):
    """dnRGG%d1, dnRGG%d2 = dB pinv(B), d2B pinv(B).

    ref: sub_dnGB_TO_dnGG (sub_module_basis.f90:2240-2330): eigen-decomposition of
    B B^T (nq x nq), keep the nb largest eigenvalues, pinv(B) = B^T V diag(1/val) V^T.
    """
    nq, nb = B.shape
    val, vec = np.linalg.eigh(B @ B.T)
    order = np.argsort(val)[::-1][:nb]
    V = vec[:, order]
    inv = (V / val[order][None, :]) @ V.T
    pinv = B.T @ inv
    return dB @ pinv, d2B @ pinv


def hm_primitive(nq: int, nb: int, Q0: float = 0.0, scaleQ: float = 1.0) -> Primitive1D:
    """``Hm`` primitive with nq Gauss-Hermite points and nb functions, Q = Q0 + x/scaleQ.

    Scaling convention of the reference basis builder (sub_scale of the basis,
    sub_module_basis.f90): x -> Q0 + x/scaleQ, w -> w/scaleQ, phi -> sqrt(scaleQ) phi,
    d/dQ -> scaleQ d/dx.
    """
    x, w = gauss_hermite(nq)
    d0, d1, d2 = hermite_functions(x, nb)
    if scaleQ != 1.0 or Q0 != 0.0:
        s = np.sqrt(scaleQ)
        d0, d1, d2 = d0 * s, d1 * s * scaleQ, d2 * s * scaleQ * scaleQ
        w = w / scaleQ
        x = Q0 + x / scaleQ
    D1, D2 = grid_derivative_matrices(d0, d1, d2)
    f = np.asfortranarray
    return Primitive1D(nq=nq, nb=nb, x=x, w=w, B=f(d0), BTw=f(d0.T * w[None, :]), D1=f(D1), D2=f(D2))


def concat_tables(prims, D: int, LG: int):
    """Concatenate per-(mode,level) tables in the canonical C-ABI order
    (mode k = 0..D-1 outer, level L = 0..LG inner; each matrix column-major).
    ``prims[k][L]`` is a Primitive1D.  Returns B, BTw, D1, D2 flat float64 arrays."""
    B = np.concatenate([prims[k][L].B.ravel(order="F") for k in range(D) for L in range(LG + 1)])
    BTw = np.concatenate([prims[k][L].BTw.ravel(order="F") for k in range(D) for L in range(LG + 1)])
    D1 = np.concatenate([prims[k][L].D1.ravel(order="F") for k in range(D) for L in range(LG + 1)])
    D2 = np.concatenate([prims[k][L].D2.ravel(order="F") for k in range(D) for L in range(LG + 1)])
    return B, BTw, D1, D2
