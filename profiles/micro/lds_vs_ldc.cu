// micro-benchmark: broadcast matrix loads from shared memory (LDS) vs constant bank (LDC, register-indexed)
#include <cstdio>
#include <cuda_runtime.h>
__constant__ double c_pool[4096];
template <int MODE, int N>
__global__ void __launch_bounds__(768, 1) k(const int *offs, double *out, int iters)
{
    extern __shared__ double s_pool[];
    __shared__ int s_off[64];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s_pool[i] = c_pool[i];
    if (threadIdx.x < 64) s_off[threadIdx.x] = offs[threadIdx.x];
    __syncthreads();
    double v[N][N];
    for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) v[j][i] = threadIdx.x * 1e-3 + i + j;
    const int w = threadIdx.x >> 5;
    for (int it = 0; it < iters; ++it) {
        const int off = s_off[(w + it) & 63];      // warp-uniform, changes every iteration (no hoisting)
        double t[N][N];
#pragma unroll
        for (int q = 0; q < N; ++q) {
            double mq[N];
#pragma unroll
            for (int b = 0; b < N; ++b) mq[b] = (MODE == 0) ? s_pool[off + q + N * b] : c_pool[off + q + N * b];
#pragma unroll
            for (int j = 0; j < N; ++j) {
                double s = mq[0] * v[j][0];
#pragma unroll
                for (int b = 1; b < N; ++b) s = fma(mq[b], v[j][b], s);
                t[j][q] = s;
            }
        }
#pragma unroll
        for (int j = 0; j < N; ++j)
#pragma unroll
            for (int q = 0; q < N; ++q) v[j][q] = t[j][q];
    }
    double s = 0;
    for (int j = 0; j < N; ++j) for (int i = 0; i < N; ++i) s += v[j][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE, int N> void run(const char *name, int *d_off, double *d_out)
{
    cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    k<MODE, N><<<148, 768, 4096 * 8>>>(d_off, d_out, 100);
    cudaEventRecord(e0);
    k<MODE, N><<<148, 768, 4096 * 8>>>(d_off, d_out, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double loads = 148.0 * 24 * iters * N * N;   // warp-level load instructions
    double fl = 148.0 * 768 * iters * N * N * N * 2.0;
    printf("%s N=%d: %.3f ms  %.2f warp-loads/clk/SM (1.965GHz)  %.2f TFLOP/s  err=%s\n", name, N, ms,
           loads / 148 / (ms * 1e-3 * 1.965e9), fl / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    double h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 0.3 / (1 + i % 7);
    cudaMemcpyToSymbol(c_pool, h, sizeof(h));
    int ho[64]; for (int i = 0; i < 64; ++i) ho[i] = (i * 37) % 3000;
    int *d_off; double *d_out; cudaMalloc(&d_off, sizeof(ho)); cudaMalloc(&d_out, 148 * 768 * 8);
    cudaMemcpy(d_off, ho, sizeof(ho), cudaMemcpyHostToDevice);
    run<0, 3>("LDS", d_off, d_out); run<1, 3>("LDC", d_off, d_out);
    run<0, 5>("LDS", d_off, d_out); run<1, 5>("LDC", d_off, d_out);
    run<0, 7>("LDS", d_off, d_out); run<1, 7>("LDC", d_off, d_out);
    return 0;
}
