/*
 * evr_sg4_vec.h -- C-ABI of the driver-side vector algebra on device-resident packed vectors.
 *
 * With these calls the drivers of the reference keep psi / H psi on the GPU between two operator actions
 * (evr_sg4_apply_device[_scaled], include/evr_sg4.h) instead of paying one H2D and one D2H copy per H|psi>:
 *   Davidson    (Source_ElVibRot/sub_propagation/sub_module_Davidson.f90)
 *       H(j,i) = <psi_j|H psi_i>, Overlap_psi1_psi2 over blocks       :984, :1110-1129   -> evr_sg4_vec_gram
 *       residual g = sum_i Vec(i,j) (H psi_i - Ene_j psi_i), MakeResidual_Davidson :1214  -> evr_sg4_vec_lincomb (twice)
 *       NewVec_type = 4 preconditioner g(ib) / (Ene_j - Ene0(ib))      :1440-1455         -> evr_sg4_vec_precond
 *       Schmidt orthonormalisation of the new vector (twice)           :1503-1518         -> evr_sg4_vec_schmidt
 *       Ritz vectors psi'_k = sum_i Vec(i,k) psi_i                                        -> evr_sg4_vec_lincomb
 *   Chebyshev   (sub_module_propa_march.f90:4142-4433): w_{k+1} = 2 Hs w_k - w_{k-1}, psi += c_k w_k, norms
 *                                                                                         -> evr_sg4_vec_lincomb, _gram
 *   SIL/Lanczos (:2899-3100): Gram-Schmidt against the Krylov vectors, tridiagonal matrix elements  -> _gram, _lincomb
 * The small dense problems (diagonalisation of the Krylov matrix, Bessel coefficients) stay on the host, as in the
 * reference (LAPACK).  Real vectors only (complex psi = two real vectors, sub_OpPsi.f90:392-407).
 *
 * Conventions: a "block" is nvec vectors of n doubles each, vector k at d_X + k*ld (ld >= n); small matrices live in HOST
 * memory, column-major; every call is ordered on `cuda_stream` (NULL = default stream) and calls that return host results
 * synchronise that stream.  Returns 0, or non-zero with a message in evr_sg4_last_error().  No CPU fallback.
 */
#ifndef EVR_SG4_VEC_H
#define EVR_SG4_VEC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* device storage for callers without a CUDA runtime of their own (Fortran): nvec vectors of n doubles, ld = n */
int evr_sg4_vec_alloc(double **d_X, int64_t n, int nvec);
int evr_sg4_vec_free(double *d_X);
int evr_sg4_vec_upload(double *d_X, const double *X_host, int64_t count, void *cuda_stream);
int evr_sg4_vec_download(double *X_host, const double *d_X, int64_t count, void *cuda_stream);

/* G[i + na*j] = <A_i | B_j>, i < na, j < nb (any sizes).  Deterministic: fixed summation order, bit-reproducible. */
int evr_sg4_vec_gram(int64_t n, int na, const double *d_A, int64_t lda, int nb, const double *d_B, int64_t ldb,
                     double *G_host, void *cuda_stream);

/* Y_k <- beta Y_k + sum_{i < nin} C[i + nin*k] X_i,  k < nout  (X and Y must not overlap). */
int evr_sg4_vec_lincomb(int64_t n, int nin, const double *d_X, int64_t ldx, int nout, const double *C_host, double beta,
                        double *d_Y, int64_t ldy, void *cuda_stream);

/* x <- a x */
int evr_sg4_vec_scale(int64_t n, double a, double *d_x, void *cuda_stream);

/* Davidson preconditioner, NewVec_type = 4: g(ib) <- g(ib) * a(ib), a = 1/Di if |Di| > conv_resi else 1/(Di + 1e-3),
 * Di = Ene_j - d_Ene0[ib]  (sub_module_Davidson.f90:1440-1455, Op_Transfo = F). */
int evr_sg4_vec_precond(int64_t n, double *d_g, const double *d_Ene0, double Ene_j, double conv_resi, void *cuda_stream);

/* Schmidt step of sub_NewVec_Davidson (:1503-1518) for a new vector v against the orthonormal block Q (ndim vectors):
 * v <- v/|v|;  v <- v - sum_i <v|q_i> q_i;  v <- v/|v|;  v <- v - sum_i <v|q_i> q_i;  norm2 = <v|v>;  v <- v/|v|.
 * norm2_host receives the smallest squared norm left by a projection of the unit vector (the reference drops the vector
 * when the last one is below 1e-10; the first one also catches an exactly dependent vector, whose renormalised rounding
 * noise would pass the second test). */
int evr_sg4_vec_schmidt(int64_t n, int ndim, const double *d_Q, int64_t ldq, double *d_v, double *norm2_host, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif
