// micro-benchmark (round 2): the one mode product of the reference inputs that is large enough to be compute-bound --
// the Pl0 mode of HCN_UT at L = 7: an 80 x 80 FP64 matrix applied to every pencil of a term grid (out = M x X, X = 80 x ncol,
// ncol = product of the other modes' sizes, batched over terms) -- done three ways:
//   dfma_thread : one output per thread, matrix through the read-only path (what the generic kernel does today)
//   dfma_tile   : register tile of 8 outputs x 4 columns per thread, matrix and X staged in shared memory
//   dmma        : mma.sync.aligned.m8n8k4.row.col.f64 (DMMA), one warp owns an 80 x 8 output block
// Reports time, TFLOP/s and checks the three results against each other.  Run under ncu for the pipe counters
// (smsp__inst_executed_pipe_fp64*, sm__inst_executed_pipe_tensor*, sm__pipe_fp64_cycles_active).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_vs_dfma dmma_vs_dfma.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define N 80
#define NCOL 1024          // columns per CTA-batch (e.g. 6 terms of 13 x 13 columns)
#define NBATCH (148 * 8)

// X, OUT: [batch][col][N] (the mode is the fastest index of the pencil: left = 1)
__global__ void __launch_bounds__(256) k_dfma_thread(const double *__restrict__ M, const double *__restrict__ X, double *__restrict__ OUT)
{
    extern __shared__ double sx[];                       // [NCOL_T][N]
    const int COLS = 32;
    for (int b = blockIdx.x; b < NBATCH; b += gridDim.x)
        for (int c0 = 0; c0 < NCOL; c0 += COLS) {
            __syncthreads();
            for (int i = threadIdx.x; i < COLS * N; i += blockDim.x) sx[i] = X[((size_t)b * NCOL + c0) * N + i];
            __syncthreads();
            for (int o = threadIdx.x; o < COLS * N; o += blockDim.x) {
                const int c = o / N, q = o - c * N;
                double s = 0.0;
                for (int k = 0; k < N; ++k) s = fma(__ldg(M + q + N * k), sx[c * N + k], s);
                OUT[((size_t)b * NCOL + c0) * N + o] = s;
            }
        }
}

// 8 rows x 4 columns per thread; CTA of 160 threads handles 80 rows x 64 columns per round: 10 row tiles x 16 column tiles
__global__ void __launch_bounds__(160) k_dfma_tile(const double *__restrict__ M, const double *__restrict__ X, double *__restrict__ OUT)
{
    extern __shared__ double sm[];
    double *sM = sm;                 // [k][q] : M(q,k) at q + N*k  (column-major as given)
    double *sx = sm + N * N;         // [64][N+1]
    const int COLS = 64, P = N + 1;
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) sM[i] = M[i];
    const int rt = threadIdx.x % 10, ct = threadIdx.x / 10;      // row tile 0..9, column tile 0..15
    for (int b = blockIdx.x; b < NBATCH; b += gridDim.x)
        for (int c0 = 0; c0 < NCOL; c0 += COLS) {
            __syncthreads();
            for (int i = threadIdx.x; i < COLS * N; i += blockDim.x) { const int c = i / N, k = i - c * N; sx[c * P + k] = X[((size_t)b * NCOL + c0) * N + i]; }
            __syncthreads();
            double acc[4][8];
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 8; ++r) acc[c][r] = 0.0;
            for (int k = 0; k < N; ++k) {
                double m[8], x[4];
#pragma unroll
                for (int r = 0; r < 8; ++r) m[r] = sM[rt * 8 + r + N * k];
#pragma unroll
                for (int c = 0; c < 4; ++c) x[c] = sx[(ct * 4 + c) * P + k];
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int r = 0; r < 8; ++r) acc[c][r] = fma(m[r], x[c], acc[c][r]);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c)
#pragma unroll
                for (int r = 0; r < 8; ++r) OUT[((size_t)b * NCOL + c0 + ct * 4 + c) * N + rt * 8 + r] = acc[c][r];
        }
}

// DMMA: C(8x8) += A(8x4) B(4x8); lane l: A row l/4 col l%4 ; B row l%4 col l/4 ; C row l/4 cols 2*(l%4)+{0,1}
__device__ __forceinline__ void dmma(double &c0, double &c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// CTA of 256 threads = 8 warps; a warp owns an 80 x 8 output block (10 C tiles); 64 columns per round
__global__ void __launch_bounds__(256) k_dmma(const double *__restrict__ M, const double *__restrict__ X, double *__restrict__ OUT)
{
    extern __shared__ double sm[];
    double *sM = sm;                 // M(q,k) at q*(N+1) + k  (row-major copy, padded: A fragments read rows)
    double *sx = sm + N * (N + 1);   // [64][N+1]
    const int COLS = 64, P = N + 1;
    for (int i = threadIdx.x; i < N * N; i += blockDim.x) { const int q = i % N, k = i / N; sM[q * P + k] = M[i]; }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ar = lane >> 2, ak = lane & 3;
    for (int b = blockIdx.x; b < NBATCH; b += gridDim.x)
        for (int c0 = 0; c0 < NCOL; c0 += COLS) {
            __syncthreads();
            for (int i = threadIdx.x; i < COLS * N; i += blockDim.x) { const int c = i / N, k = i - c * N; sx[c * P + k] = X[((size_t)b * NCOL + c0) * N + i]; }
            __syncthreads();
            double c[10][2];
#pragma unroll
            for (int t = 0; t < 10; ++t) c[t][0] = c[t][1] = 0.0;
            for (int k0 = 0; k0 < N; k0 += 4) {
                const double bf = sx[(warp * 8 + ar) * P + k0 + ak];          // B(k, col): row k0+ak, col lane/4
#pragma unroll
                for (int t = 0; t < 10; ++t) dmma(c[t][0], c[t][1], sM[(t * 8 + ar) * P + k0 + ak], bf);
            }
#pragma unroll
            for (int t = 0; t < 10; ++t) {
                const size_t col = (size_t)b * NCOL + c0 + warp * 8 + 2 * ak;
                OUT[col * N + t * 8 + ar] = c[t][0];
                OUT[(col + 1) * N + t * 8 + ar] = c[t][1];
            }
        }
}

int main()
{
    const size_t nx = (size_t)NBATCH * NCOL * N;
    std::vector<double> hM(N * N), hX(nx);
    srand(1);
    for (auto &v : hM) v = (rand() / (double)RAND_MAX - 0.5);
    for (auto &v : hX) v = (rand() / (double)RAND_MAX - 0.5);
    double *dM, *dX, *dO[3];
    cudaMalloc(&dM, N * N * 8); cudaMalloc(&dX, nx * 8);
    for (int i = 0; i < 3; ++i) cudaMalloc(&dO[i], nx * 8);
    cudaMemcpy(dM, hM.data(), N * N * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dX, hX.data(), nx * 8, cudaMemcpyHostToDevice);
    const size_t sm0 = 32 * N * 8, sm1 = (N * N + 64 * (N + 1)) * 8, sm2 = (N * (N + 1) + 64 * (N + 1)) * 8;
    cudaFuncSetAttribute(k_dfma_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1);
    cudaFuncSetAttribute(k_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double flop = 2.0 * N * N * (double)NBATCH * NCOL;
    const char *name[3] = {"dfma_thread", "dfma_tile  ", "dmma       "};
    for (int v = 0; v < 3; ++v) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (v == 0) k_dfma_thread<<<148 * 4, 256, sm0>>>(dM, dX, dO[0]);
            if (v == 1) k_dfma_tile<<<148 * 4, 160, sm1>>>(dM, dX, dO[1]);
            if (v == 2) k_dmma<<<148 * 2, 256, sm2>>>(dM, dX, dO[2]);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        printf("%s  %8.3f ms  %6.2f TFLOP/s  %s\n", name[v], best, flop / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
    }
    std::vector<double> r0(1 << 16), r1(1 << 16), r2(1 << 16);
    cudaMemcpy(r0.data(), dO[0], r0.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r1.data(), dO[1], r1.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(r2.data(), dO[2], r2.size() * 8, cudaMemcpyDeviceToHost);
    double d1 = 0, d2 = 0, mx = 0;
    for (size_t i = 0; i < r0.size(); ++i) { d1 = fmax(d1, fabs(r1[i] - r0[i])); d2 = fmax(d2, fabs(r2[i] - r0[i])); mx = fmax(mx, fabs(r0[i])); }
    printf("max |tile - thread| = %.2e, max |dmma - thread| = %.2e (max |value| %.2f)\n", d1, d2, mx);
    return 0;
}
