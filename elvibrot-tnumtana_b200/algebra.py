"""Host-side mirror of the driver-side vector algebra of the reference on device-resident vectors
(C-ABI include/evr_sg4_vec.h -> csrc/sg4_algebra.cu).

=========================  ==========================================================================
here                       reference (Source_ElVibRot/sub_propagation/sub_module_Davidson.f90)
=========================  ==========================================================================
``gram``                   ``Overlap_psi1_psi2`` over blocks: H(j,i) = <psi_j|H psi_i> (:1110-1129), S
``lincomb``                Ritz vectors / ``MakeResidual_Davidson`` (:1214), Chebyshev sums (propa_march :4294-4345)
``precond``                NewVec_type = 4 preconditioner (:1440-1455)
``schmidt``                Schmidt orthonormalisation of the new vector (:1503-1518)
=========================  ==========================================================================

Vectors are rows of 2-D float64 CUDA tensors (one RvecB per row); torch only provides the device memory.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib


def _rows(t):
    import torch
    assert t.is_cuda and t.dtype == torch.float64 and t.stride(-1) == 1
    if t.dim() == 1:
        return 1, t.numel(), t.numel()
    assert t.dim() == 2
    return t.shape[0], t.shape[1], t.stride(0)


def _stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def gram(A, B) -> np.ndarray:
    """G[i, j] = <A_i | B_j> (deterministic summation order)."""
    na, n, lda = _rows(A)
    nb, n2, ldb = _rows(B)
    assert n == n2
    G = np.empty((nb, na))                      # column-major [na][nb] seen from C
    _lib.check(_lib.lib().evr_sg4_vec_gram(n, na, A.data_ptr(), lda, nb, B.data_ptr(), ldb, G.ctypes.data, _stream()), "evr_sg4_vec_gram")
    return G.T.copy()


def lincomb(X, Cmat, Y, beta: float = 0.0):
    """Y_k <- beta Y_k + sum_i Cmat[i, k] X_i."""
    nin, n, ldx = _rows(X)
    nout, n2, ldy = _rows(Y)
    Cm = np.asfortranarray(np.asarray(Cmat, dtype=np.float64).reshape(nin, nout))
    assert n == n2
    _lib.check(_lib.lib().evr_sg4_vec_lincomb(n, nin, X.data_ptr(), ldx, nout, Cm.ctypes.data, float(beta), Y.data_ptr(), ldy, _stream()),
               "evr_sg4_vec_lincomb")
    return Y


def scale(x, a: float):
    _lib.check(_lib.lib().evr_sg4_vec_scale(x.numel(), float(a), x.data_ptr(), _stream()), "evr_sg4_vec_scale")
    return x


def precond(g, Ene0, Ene_j: float, conv_resi: float):
    _lib.check(_lib.lib().evr_sg4_vec_precond(g.numel(), g.data_ptr(), Ene0.data_ptr(), float(Ene_j), float(conv_resi), _stream()),
               "evr_sg4_vec_precond")
    return g


def schmidt(Q, v) -> float:
    """Orthonormalise ``v`` against the rows of ``Q`` (twice, like sub_NewVec_Davidson); returns the squared norm before
    the final normalisation."""
    ndim, n, ldq = _rows(Q) if Q is not None and Q.numel() else (0, v.numel(), v.numel())
    nn = C.c_double()
    _lib.check(_lib.lib().evr_sg4_vec_schmidt(n, ndim, Q.data_ptr() if ndim else None, ldq, v.data_ptr(), C.byref(nn), _stream()),
               "evr_sg4_vec_schmidt")
    return nn.value
