// sg4_opgrid.cu -- operator-grid construction on the device for closed-form models.
//
// The reference fills the potential grid during its first H|psi>: for every point of every term grid (first index
// fastest, ADD_ONE_TO_nDval_m1) Rec_Qact_SG4_with_Tab_iq builds Qact from the 1-D grid points of the term's levels and
// get_d0MatOp_AT_Qact evaluates the potential (sub_Operator/sub_OpPsi_SG4.f90:2982-3006), one point at a time on the host.
// For models that are closed-form in the coordinates this is one trivially parallel kernel: one CTA per Smolyak term, one
// thread per grid point, the multi-index decoded mode by mode.  Output layout = OpGrid(iterm00)%Grid(1:NQ): terms in iG
// order (offset tab_Sum_nq - nq), first mode fastest -- exactly what evr_sg4_plan_set_op expects.
// Models: 1 = Henon-Heiles  V = 1/2 sum Q_i^2 + lambda sum_{i<D} (Q_i^2 Q_{i+1} - Q_{i+1}^3 / 3)
//             (Working_tests/MPI_tests/6D_Davidson_openMP/sub_system_HenonHeiles.f:40-47; params[0] = lambda)
//         2 = uncoupled harmonic  V = 1/2 sum_i params[i] Q_i^2
#include "../../include/evr_sg4.h"
#include "sg4_internal.h"

#include <cuda_runtime.h>

#include <string>
#include <vector>

using evr::fail;

namespace evr {

__global__ void __launch_bounds__(256)
sg4_model_grid_kernel(const int D, const int LG, const int iG_begin, const int n_terms, const int model,
                      const int32_t *__restrict__ tab_l, const int32_t *__restrict__ nq_of, const long long *__restrict__ xoff,
                      const double *__restrict__ xtab, const long long *__restrict__ goff, const double *__restrict__ params,
                      double *__restrict__ V)
{
    __shared__ int s_n[EVR_MAXD];
    __shared__ long long s_x[EVR_MAXD];
    for (int t = blockIdx.x; t < n_terms; t += gridDim.x) {
        const int iG = iG_begin + t;
        __syncthreads();
        if (threadIdx.x < D) {
            const int i = threadIdx.x * (LG + 1) + tab_l[(long long)iG * D + threadIdx.x];
            s_n[threadIdx.x] = nq_of[i];
            s_x[threadIdx.x] = xoff[i];
        }
        __syncthreads();
        const long long nq = goff[t + 1] - goff[t];
        for (long long q = threadIdx.x; q < nq; q += blockDim.x) {
            long long r = q;
            double v = 0.0, prev = 0.0;
            for (int k = 0; k < D; ++k) {                 // first mode fastest
                const int n = s_n[k];
                const int ik = (int)(r % n);
                r /= n;
                const double x = __ldg(xtab + s_x[k] + ik);
                if (model == 1) {
                    v += 0.5 * x * x;
                    if (k > 0) v += params[0] * (prev * prev * x - x * x * x / 3.0);
                } else {
                    v += 0.5 * params[k] * x * x;
                }
                prev = x;
            }
            V[goff[t] + q] = v;
        }
    }
}

} // namespace evr

extern "C" int evr_sg4_model_grid(int model, int D, int nb_SG, int LG, const int32_t *tab_l, const int32_t *nq_of,
                                  const double *x_tab, int nparam, const double *params, int iG_begin, int iG_end,
                                  double *V_host)
{
    if (!tab_l || !nq_of || !x_tab || !V_host || D < 1 || D > EVR_MAXD || LG < 0 || nb_SG < 1)
        return fail("evr_sg4_model_grid: bad arguments");
    if (iG_begin < 0 || iG_end > nb_SG || iG_begin > iG_end) return fail("evr_sg4_model_grid: bad term range");
    if (model == 1) { if (nparam < 1 || !params) return fail("evr_sg4_model_grid: Henon-Heiles needs params[0] = lambda"); }
    else if (model == 2) { if (nparam < D || !params) return fail("evr_sg4_model_grid: harmonic model needs D force constants"); }
    else return fail("evr_sg4_model_grid: unknown model (1 = Henon-Heiles, 2 = uncoupled harmonic)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail("evr_sg4_model_grid: no CUDA device available (this library has no CPU fallback)");
    const int nT = D * (LG + 1), n_terms = iG_end - iG_begin;
    std::vector<long long> xoff(nT), goff(n_terms + 1, 0);
    long long nx = 0;
    for (int i = 0; i < nT; ++i) { if (nq_of[i] < 1) return fail("evr_sg4_model_grid: nq < 1"); xoff[i] = nx; nx += nq_of[i]; }
    for (int t = 0; t < n_terms; ++t) {
        long long nq = 1;
        for (int k = 0; k < D; ++k) {
            const int l = tab_l[(size_t)(iG_begin + t) * D + k];
            if (l < 0 || l > LG) return fail("evr_sg4_model_grid: level out of range in tab_l");
            nq *= nq_of[k * (LG + 1) + l];
        }
        goff[t + 1] = goff[t] + nq;
    }
    if (n_terms == 0) return 0;
    int32_t *d_l = nullptr, *d_nq = nullptr; long long *d_xoff = nullptr, *d_goff = nullptr; double *d_x = nullptr, *d_p = nullptr, *d_V = nullptr;
    auto cleanup = [&]() { cudaFree(d_l); cudaFree(d_nq); cudaFree(d_xoff); cudaFree(d_goff); cudaFree(d_x); cudaFree(d_p); cudaFree(d_V); };
    bool ok = cudaMalloc((void **)&d_l, (size_t)nb_SG * D * 4) == cudaSuccess && cudaMalloc((void **)&d_nq, nT * 4) == cudaSuccess &&
              cudaMalloc((void **)&d_xoff, nT * 8) == cudaSuccess && cudaMalloc((void **)&d_goff, (size_t)(n_terms + 1) * 8) == cudaSuccess &&
              cudaMalloc((void **)&d_x, (size_t)nx * 8) == cudaSuccess && cudaMalloc((void **)&d_p, (size_t)nparam * 8) == cudaSuccess &&
              cudaMalloc((void **)&d_V, (size_t)goff[n_terms] * 8) == cudaSuccess;
    ok = ok && cudaMemcpy(d_l, tab_l, (size_t)nb_SG * D * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_nq, nq_of, nT * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_xoff, xoff.data(), nT * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_goff, goff.data(), (size_t)(n_terms + 1) * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_x, x_tab, (size_t)nx * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
         cudaMemcpy(d_p, params, (size_t)nparam * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) { cleanup(); return fail("evr_sg4_model_grid: device allocation / upload failed"); }
    const int ctas = std::min(n_terms, 148 * 8);
    evr::sg4_model_grid_kernel<<<ctas, 256>>>(D, LG, iG_begin, n_terms, model, d_l, d_nq, d_xoff, d_x, d_goff, d_p, d_V);
    ok = cudaGetLastError() == cudaSuccess &&
         cudaMemcpy(V_host, d_V, (size_t)goff[n_terms] * 8, cudaMemcpyDeviceToHost) == cudaSuccess;
    cleanup();
    return ok ? 0 : fail("evr_sg4_model_grid: kernel failed");
}
