"""Driver-side vector algebra on device-resident vectors (include/evr_sg4_vec.h): unit tests against numpy
and a block Davidson in the style of the reference (sub_propagation/sub_module_Davidson.f90) whose vectors never leave the GPU."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_gram_lincomb_precond_schmidt_against_numpy(evr):
    import torch
    rng = np.random.default_rng(1)
    n = 12347
    A, B = rng.standard_normal((37, n)), rng.standard_normal((5, n))
    dA, dB = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    G = evr.algebra.gram(dA, dB)
    assert G.shape == (37, 5) and np.abs(G - A @ B.T).max() < 1e-10
    assert np.array_equal(G, evr.algebra.gram(dA, dB))                      # fixed summation order: bit-reproducible
    Cm = rng.standard_normal((37, 11))
    Y0 = rng.standard_normal((11, n))
    dY = torch.from_numpy(Y0).cuda()
    evr.algebra.lincomb(dA, Cm, dY, beta=0.5)
    assert np.abs(dY.cpu().numpy() - (0.5 * Y0 + Cm.T @ A)).max() < 1e-11
    g, e0 = rng.standard_normal(n), rng.uniform(0.0, 5.0, n)
    e0[:3] = 1.25                                                            # |Di| <= conv_resi branch
    dg = torch.from_numpy(g.copy()).cuda()
    evr.algebra.precond(dg, torch.from_numpy(e0).cuda(), 1.25, 1e-4)
    Di = 1.25 - e0
    ref = g * np.where(np.abs(Di) > 1e-4, 1.0 / np.where(Di == 0, 1.0, Di), 1.0 / (Di + 1e-3))
    assert np.abs(dg.cpu().numpy() - ref).max() < 1e-12 * np.abs(ref).max()
    Q, _ = np.linalg.qr(rng.standard_normal((n, 9)))
    dQ = torch.from_numpy(np.ascontiguousarray(Q.T)).cuda()
    v = rng.standard_normal(n)
    dv = torch.from_numpy(v.copy()).cuda()
    nn = evr.algebra.schmidt(dQ, dv)
    w = dv.cpu().numpy()
    assert 0.9 < nn <= 1.0 + 1e-12 and abs(np.linalg.norm(w) - 1.0) < 1e-13 and np.abs(Q.T @ w).max() < 1e-14
    dep = torch.from_numpy(Q[:, 3].copy()).cuda()                            # a dependent vector is reported by its norm
    assert evr.algebra.schmidt(dQ, dep) < 1e-10


def _davidson(evr, op, Ene0, nb_diago, conv_resi=1e-9, max_it=60, max_dim=400):
    """Block Davidson with the NewVec_type=4 preconditioner (sub_module_Davidson.f90:300-420, 1290-1588): psi, H psi,
    residuals and new vectors are rows of device tensors; only the small Krylov matrix visits the host."""
    import torch
    b = op.BasisnD
    n = b.nb * b.nb0
    psi = torch.zeros((max_dim, n), dtype=torch.float64, device="cuda")
    Hpsi = torch.zeros_like(psi)
    dE0 = torch.from_numpy(Ene0).cuda()
    order = np.argsort(Ene0, kind="stable")[:nb_diago]
    psi[torch.arange(nb_diago), torch.from_numpy(order).cuda()] = 1.0        # guess: the nb_diago lowest zero-order functions
    ndim0, ndim, n_apply = 0, nb_diago, 0
    H = np.zeros((0, 0))
    g = torch.empty(n, dtype=torch.float64, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    for it in range(max_it):
        op.apply_device_ptr(ndim - ndim0, psi[ndim0].data_ptr(), Hpsi[ndim0].data_ptr(), st)     # sub_TabOpPsi on the new block
        n_apply += ndim - ndim0
        Hn = np.zeros((ndim, ndim))
        Hn[:ndim0, :ndim0] = H
        Hn[:, ndim0:] = evr.algebra.gram(psi[:ndim], Hpsi[ndim0:ndim])       # H(j,i) = <psi_j|H psi_i>, new columns
        if ndim0:
            Hn[ndim0:, :ndim0] = evr.algebra.gram(psi[ndim0:ndim], Hpsi[:ndim0])
        H = Hn
        Ene, Vec = np.linalg.eigh(0.5 * (H + H.T))
        ndim0 = ndim
        worst = 0.0
        for j in range(nb_diago):
            # residual g = sum_i Vec(i,j) (H psi_i - Ene_j psi_i)   (MakeResidual_Davidson)
            evr.algebra.lincomb(Hpsi[:ndim0], Vec[:, j:j + 1], g[None, :], beta=0.0)
            evr.algebra.lincomb(psi[:ndim0], -Ene[j] * Vec[:, j:j + 1], g[None, :], beta=1.0)
            res = float(np.sqrt(evr.algebra.gram(g, g)[0, 0]))
            worst = max(worst, res)
            if res < conv_resi or ndim >= max_dim:
                continue
            evr.algebra.precond(g, dE0, Ene[j], 1e-4)
            psi[ndim].copy_(g)
            if evr.algebra.schmidt(psi[:ndim], psi[ndim]) > 1e-10:
                ndim += 1
        if ndim == ndim0:
            break
    return Ene[:nb_diago], worst, n_apply, ndim0


def test_block_davidson_with_device_resident_vectors_reproduces_the_reference_levels(evr, golden):
    """Henon-Heiles 6-D, SG4 L=3 (Working_tests/MPI_tests/6D_Davidson_openMP): the 28 lowest levels from a Davidson run
    whose vectors stay on the GPU, against the dense diagonalisation of the same H and the reference's benchmark file."""
    basis, op = evr.workloads.henon_heiles(6, 3)
    ref = np.array(golden["kat"]["HH6D_L3"]["levels"])
    Ene0 = (basis.nDindB_Tab_nDval - 0.5).sum(axis=1)                        # sum_k (n_k + 1/2), n_k = index - 1
    ene, resid, n_apply, ndim = _davidson(evr, op, Ene0, nb_diago=len(ref))
    Hd = op.apply_host(np.eye(basis.nb)).T
    dense = np.sort(np.linalg.eigvalsh(0.5 * (Hd + Hd.T)))[: len(ref)]
    assert resid < 1e-8, resid
    assert np.abs(ene - dense).max() < 1e-9
    assert np.abs(ene - ref).max() < 2e-7          # same bound as the dense test: the reference run is converged to conv_ene = 1e-4 au (shell_run:157-165)
    assert n_apply < basis.nb                      # far fewer H|psi> than the dense build
