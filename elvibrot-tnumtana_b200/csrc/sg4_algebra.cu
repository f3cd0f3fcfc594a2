// sg4_algebra.cu -- driver-side vector algebra on device-resident packed vectors (include/evr_sg4_vec.h): the inner
// products, linear combinations, preconditioner and Schmidt step of the reference's Davidson / Chebyshev / SIL drivers
// (sub_propagation/sub_module_Davidson.f90:984-1129, 1214, 1440-1518; sub_module_propa_march.f90:2899-3100, 4142-4433),
// so that psi never leaves the GPU between two H|psi>.
#include "../../include/evr_sg4_vec.h"
#include "sg4_internal.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

using evr::fail;

#define VCUDA(expr)                                                                            \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) return fail(std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

namespace evr {

#define GRAM_T 32          // pairs tile: up to 32 x 32 inner products per launch
#define GRAM_CH 64        // elements staged per round
#define GRAM_PITCH (GRAM_CH + 1)

// stage 1: every CTA accumulates the na x nb inner products over its (grid-strided) chunks of the vectors and writes
// one partial matrix; stage 2 sums the partial matrices in CTA order -> the result does not depend on scheduling
__global__ void __launch_bounds__(256)
vec_gram_partial(const long long n, const int na, const double *__restrict__ A, const long long lda,
                 const int nb, const double *__restrict__ B, const long long ldb, double *__restrict__ part)
{
    __shared__ double sA[GRAM_T * GRAM_PITCH], sB[GRAM_T * GRAM_PITCH];
    const int npairs = na * nb;
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
    for (long long c0 = (long long)blockIdx.x * GRAM_CH; c0 < n; c0 += (long long)gridDim.x * GRAM_CH) {
        const int len = (int)min((long long)GRAM_CH, n - c0);
        for (int i = threadIdx.x; i < na * GRAM_CH; i += blockDim.x) {
            const int r = i / GRAM_CH, k = i - r * GRAM_CH;
            sA[r * GRAM_PITCH + k] = (k < len) ? __ldg(A + r * lda + c0 + k) : 0.0;
        }
        for (int i = threadIdx.x; i < nb * GRAM_CH; i += blockDim.x) {
            const int r = i / GRAM_CH, k = i - r * GRAM_CH;
            sB[r * GRAM_PITCH + k] = (k < len) ? __ldg(B + r * ldb + c0 + k) : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = threadIdx.x + u * 256;
            if (p < npairs) {
                const int i = p % na, j = p / na;
                const double *a = sA + i * GRAM_PITCH, *b = sB + j * GRAM_PITCH;
                double s = acc[u];
#pragma unroll 8
                for (int k = 0; k < GRAM_CH; ++k) s = fma(a[k], b[k], s);
                acc[u] = s;
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        const int p = threadIdx.x + u * 256;
        if (p < npairs) part[(long long)blockIdx.x * npairs + p] = acc[u];
    }
}
__global__ void vec_gram_final(const int nctas, const int npairs, const double *__restrict__ part, double *__restrict__ G)
{
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < nctas; ++c) s += part[(long long)c * npairs + p];
        G[p] = s;
    }
}

#define LC_K 8             // output vectors per pass of the linear-combination kernel
__global__ void __launch_bounds__(256)
vec_lincomb(const long long n, const int nin, const double *__restrict__ X, const long long ldx, const int nout_here,
            const double *__restrict__ C /* device, [nin][LC_K] */, const double beta, double *__restrict__ Y, const long long ldy)
{
    extern __shared__ double sC[];
    for (int i = threadIdx.x; i < nin * LC_K; i += blockDim.x) sC[i] = C[i];
    __syncthreads();
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        double acc[LC_K];
#pragma unroll
        for (int k = 0; k < LC_K; ++k) acc[k] = (k < nout_here && beta != 0.0) ? beta * Y[k * ldy + e] : 0.0;
        for (int i = 0; i < nin; ++i) {
            const double x = __ldg(X + i * ldx + e);
#pragma unroll
            for (int k = 0; k < LC_K; ++k) acc[k] = fma(sC[i * LC_K + k], x, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < LC_K; ++k) if (k < nout_here) Y[k * ldy + e] = acc[k];
    }
}
__global__ void vec_scale(const long long n, const double a, double *__restrict__ x)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) x[e] *= a;
}
__global__ void vec_precond(const long long n, double *__restrict__ g, const double *__restrict__ E0, const double Ej, const double conv)
{
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const double Di = Ej - __ldg(E0 + e);
        g[e] *= (fabs(Di) > conv) ? 1.0 / Di : 1.0 / (Di + 1e-3);
    }
}

static int grid_1d(long long n) { return (int)std::max<long long>(1, std::min<long long>((n + 255) / 256, 148 * 8)); }

// per-device scratch (partial Gram matrices, coefficient blocks); grown on demand, never shrunk
struct Scratch { double *part = nullptr; size_t part_cap = 0; double *coef = nullptr; size_t coef_cap = 0; double *G = nullptr; };
static Scratch g_scratch[64];
static int scratch_for(Scratch **out, size_t part_need, size_t coef_need)
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return fail("evr_sg4_vec: no CUDA device available (this library has no CPU fallback)");
    Scratch &S = g_scratch[dev & 63];
    if (part_need > S.part_cap) {
        cudaFree(S.part); S.part = nullptr; S.part_cap = 0;
        if (cudaMalloc((void **)&S.part, part_need * sizeof(double)) != cudaSuccess) return fail("evr_sg4_vec: cudaMalloc failed");
        S.part_cap = part_need;
    }
    if (coef_need > S.coef_cap) {
        cudaFree(S.coef); S.coef = nullptr; S.coef_cap = 0;
        if (cudaMalloc((void **)&S.coef, coef_need * sizeof(double)) != cudaSuccess) return fail("evr_sg4_vec: cudaMalloc failed");
        S.coef_cap = coef_need;
    }
    if (!S.G && cudaMalloc((void **)&S.G, GRAM_T * GRAM_T * sizeof(double)) != cudaSuccess) return fail("evr_sg4_vec: cudaMalloc failed");
    *out = &S;
    return 0;
}

} // namespace evr

extern "C" int evr_sg4_vec_alloc(double **d_X, int64_t n, int nvec)
{
    if (!d_X || n < 1 || nvec < 1) return fail("evr_sg4_vec_alloc: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail("evr_sg4_vec_alloc: no CUDA device available (this library has no CPU fallback)");
    VCUDA(cudaMalloc((void **)d_X, (size_t)n * nvec * sizeof(double)));
    return 0;
}
extern "C" int evr_sg4_vec_free(double *d_X) { if (d_X) VCUDA(cudaFree(d_X)); return 0; }
extern "C" int evr_sg4_vec_upload(double *d_X, const double *X_host, int64_t count, void *st)
{
    if (!d_X || !X_host || count < 0) return fail("evr_sg4_vec_upload: bad arguments");
    VCUDA(cudaMemcpyAsync(d_X, X_host, (size_t)count * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)st));
    return 0;
}
extern "C" int evr_sg4_vec_download(double *X_host, const double *d_X, int64_t count, void *st)
{
    if (!d_X || !X_host || count < 0) return fail("evr_sg4_vec_download: bad arguments");
    VCUDA(cudaMemcpyAsync(X_host, d_X, (size_t)count * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)st));
    VCUDA(cudaStreamSynchronize((cudaStream_t)st));
    return 0;
}

extern "C" int evr_sg4_vec_gram(int64_t n, int na, const double *d_A, int64_t lda, int nb, const double *d_B, int64_t ldb,
                                double *G_host, void *stream)
{
    using namespace evr;
    if (n < 1 || na < 1 || nb < 1 || !d_A || !d_B || !G_host || lda < n || ldb < n) return fail("evr_sg4_vec_gram: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int nctas = (int)std::max<long long>(1, std::min<long long>((n + GRAM_CH - 1) / GRAM_CH, 148 * 4));
    Scratch *S = nullptr;
    if (scratch_for(&S, (size_t)nctas * GRAM_T * GRAM_T, 0)) return 1;
    std::vector<double> tile(GRAM_T * GRAM_T);
    for (int j0 = 0; j0 < nb; j0 += GRAM_T)
        for (int i0 = 0; i0 < na; i0 += GRAM_T) {
            const int ta = std::min(GRAM_T, na - i0), tb = std::min(GRAM_T, nb - j0);
            vec_gram_partial<<<nctas, 256, 0, st>>>(n, ta, d_A + (size_t)i0 * lda, lda, tb, d_B + (size_t)j0 * ldb, ldb, S->part);
            vec_gram_final<<<4, 256, 0, st>>>(nctas, ta * tb, S->part, S->G);
            VCUDA(cudaGetLastError());
            VCUDA(cudaMemcpyAsync(tile.data(), S->G, (size_t)ta * tb * sizeof(double), cudaMemcpyDeviceToHost, st));
            VCUDA(cudaStreamSynchronize(st));
            for (int j = 0; j < tb; ++j)
                for (int i = 0; i < ta; ++i) G_host[(size_t)(i0 + i) + (size_t)na * (j0 + j)] = tile[i + ta * j];
        }
    return 0;
}

extern "C" int evr_sg4_vec_lincomb(int64_t n, int nin, const double *d_X, int64_t ldx, int nout, const double *C_host, double beta,
                                   double *d_Y, int64_t ldy, void *stream)
{
    using namespace evr;
    if (n < 1 || nin < 0 || nout < 1 || !d_Y || ldy < n || (nin > 0 && (!d_X || !C_host || ldx < n)))
        return fail("evr_sg4_vec_lincomb: bad arguments");
    if (nin > 0) {
        const uintptr_t x0 = (uintptr_t)d_X, x1 = x0 + ((size_t)(nin - 1) * ldx + n) * 8, y0 = (uintptr_t)d_Y, y1 = y0 + ((size_t)(nout - 1) * ldy + n) * 8;
        if (x0 < y1 && y0 < x1) return fail("evr_sg4_vec_lincomb: X and Y overlap");
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int npass = (nout + LC_K - 1) / LC_K;
    Scratch *S = nullptr;
    if (scratch_for(&S, 0, (size_t)std::max(1, nin) * LC_K * npass)) return 1;
    // coefficient blocks [pass][nin][LC_K] (zero padded), one upload
    std::vector<double> blk((size_t)std::max(1, nin) * LC_K * npass, 0.0);
    for (int k = 0; k < nout; ++k)
        for (int i = 0; i < nin; ++i) blk[((size_t)(k / LC_K) * nin + i) * LC_K + (k % LC_K)] = C_host[(size_t)i + (size_t)nin * k];
    VCUDA(cudaMemcpyAsync(S->coef, blk.data(), blk.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    VCUDA(cudaStreamSynchronize(st));                  // blk is a host temporary
    for (int ps = 0; ps < npass; ++ps) {
        const int here = std::min(LC_K, nout - ps * LC_K);
        vec_lincomb<<<grid_1d(n), 256, (size_t)std::max(1, nin) * LC_K * sizeof(double), st>>>(
            n, nin, d_X, ldx, here, S->coef + (size_t)ps * nin * LC_K, beta, d_Y + (size_t)ps * LC_K * ldy, ldy);
    }
    VCUDA(cudaGetLastError());
    return 0;
}

extern "C" int evr_sg4_vec_scale(int64_t n, double a, double *d_x, void *stream)
{
    if (n < 1 || !d_x) return fail("evr_sg4_vec_scale: bad arguments");
    evr::vec_scale<<<evr::grid_1d(n), 256, 0, (cudaStream_t)stream>>>(n, a, d_x);
    VCUDA(cudaGetLastError());
    return 0;
}

extern "C" int evr_sg4_vec_precond(int64_t n, double *d_g, const double *d_Ene0, double Ene_j, double conv_resi, void *stream)
{
    if (n < 1 || !d_g || !d_Ene0) return fail("evr_sg4_vec_precond: bad arguments");
    evr::vec_precond<<<evr::grid_1d(n), 256, 0, (cudaStream_t)stream>>>(n, d_g, d_Ene0, Ene_j, conv_resi);
    VCUDA(cudaGetLastError());
    return 0;
}

extern "C" int evr_sg4_vec_schmidt(int64_t n, int ndim, const double *d_Q, int64_t ldq, double *d_v, double *norm2_host, void *stream)
{
    if (n < 1 || ndim < 0 || !d_v || !norm2_host || (ndim > 0 && (!d_Q || ldq < n))) return fail("evr_sg4_vec_schmidt: bad arguments");
    double nn = 0.0, worst = 1.0;
    std::vector<double> r(std::max(1, ndim));
    for (int round = 0; round < 2; ++round) {
        if (evr_sg4_vec_gram(n, 1, d_v, n, 1, d_v, n, &nn, stream)) return 1;
        if (round == 1) worst = std::min(worst, nn);       // squared norm left by the first projection of the unit vector
        if (!(nn > 0.0)) { *norm2_host = 0.0; return 0; }
        if (evr_sg4_vec_scale(n, 1.0 / std::sqrt(nn), d_v, stream)) return 1;
        if (ndim > 0) {
            if (evr_sg4_vec_gram(n, ndim, d_Q, ldq, 1, d_v, n, r.data(), stream)) return 1;
            for (int i = 0; i < ndim; ++i) r[i] = -r[i];
            if (evr_sg4_vec_lincomb(n, ndim, d_Q, ldq, 1, r.data(), 1.0, d_v, n, stream)) return 1;
        }
    }
    if (evr_sg4_vec_gram(n, 1, d_v, n, 1, d_v, n, &nn, stream)) return 1;
    *norm2_host = std::min(worst, nn);
    if (nn > 0.0 && evr_sg4_vec_scale(n, 1.0 / std::sqrt(nn), d_v, stream)) return 1;
    return 0;
}
