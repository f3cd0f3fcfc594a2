// micro-benchmark (round 2): what do the memory phases of the SG4 term kernel cost in the LSU / L1TEX pipe of one SM?
//   red     : FP64 RED (atomicAdd without return) to an L2-resident vector of 1.39 M doubles
//   ldgsts  : 8-byte cp.async gathers from the same vector into shared memory
//   ldg     : LDG.64 + STS.64 gathers
//   maponly : the index stream alone (4 B per entry, coalesced) -- subtract it from the rows above
// address patterns: random | runs (sorted chunks of 4096 entries made of contiguous runs of 2..64 doubles, like the sorted
// scatter map of a batch in the block-ordered packed vector) | runs-unsorted (same runs, run order random: the gather map)
// | contiguous.   148 CTAs x 768 threads, every lane one entry per round (coalesced index loads).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o l1tex_costs l1tex_costs.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define NB 1392065
#define THREADS 768
#define CTAS 148

__global__ void __launch_bounds__(THREADS, 1) k_red(const int *__restrict__ idx, double *y, int rounds)
{
    const long long base = (long long)blockIdx.x * rounds * THREADS;
    for (int r = 0; r < rounds; ++r) {
        const int m = __ldg(idx + base + (long long)r * THREADS + threadIdx.x);
        atomicAdd(y + m, 1.0);
    }
}
__global__ void __launch_bounds__(THREADS, 1) k_maponly(const int *__restrict__ idx, double *y, int rounds)
{
    const long long base = (long long)blockIdx.x * rounds * THREADS;
    int s = 0;
    for (int r = 0; r < rounds; ++r) s += __ldg(idx + base + (long long)r * THREADS + threadIdx.x);
    if (s == -12345) y[0] = 1.0;
}
__global__ void __launch_bounds__(THREADS, 1) k_ldgsts(const int *__restrict__ idx, const double *__restrict__ x, double *y, int rounds)
{
    extern __shared__ double sm[];
    const long long base = (long long)blockIdx.x * rounds * THREADS;
    for (int r = 0; r < rounds; ++r) {
        const int m = __ldg(idx + base + (long long)r * THREADS + threadIdx.x);
        double *dst = sm + (r & 7) * THREADS + threadIdx.x;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(x + m) : "memory");
        if ((r & 7) == 7) asm volatile("cp.async.commit_group;\ncp.async.wait_group 1;" ::: "memory");
    }
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    if (sm[threadIdx.x] == -1.2345) y[0] = 1.0;
}
__global__ void __launch_bounds__(THREADS, 1) k_ldg(const int *__restrict__ idx, const double *__restrict__ x, double *y, int rounds)
{
    extern __shared__ double sm[];
    const long long base = (long long)blockIdx.x * rounds * THREADS;
    for (int r = 0; r < rounds; r += 4) {
        int m[4]; double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) m[u] = __ldg(idx + base + (long long)(r + u) * THREADS + threadIdx.x);
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = __ldg(x + m[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) sm[((r + u) & 7) * THREADS + threadIdx.x] = v[u];
    }
    if (sm[threadIdx.x] == -1.2345) y[0] = 1.0;
}
// shared-memory sweeps of the transform passes: every thread loads and stores a tile of TILE doubles at stride `stride`
// (tile origin as in sg4_fast.cuh: t + stride*(TILE-1)*(t/stride)); no arithmetic
template <int TILE>
__global__ void __launch_bounds__(THREADS, 1) k_sweep(double *y, int rounds, int stride, int npts)
{
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < npts; i += THREADS) sm[i] = i;
    __syncthreads();
    const int ntiles = npts / TILE;
    double acc = 0.0;
    for (int r = 0; r < rounds; ++r) {
        for (int t = threadIdx.x; t < ntiles; t += THREADS) {
            const int q0 = t + stride * (TILE - 1) * (t / stride);
            double v[TILE];
#pragma unroll
            for (int i = 0; i < TILE; ++i) v[i] = sm[q0 + stride * i];
#pragma unroll
            for (int i = 0; i < TILE; ++i) sm[q0 + stride * i] = v[i] + 1.0;
        }
        __syncthreads();
    }
    if (acc == -1.0) y[0] = sm[0];
}

static std::vector<int> make_pattern(int kind, long long n)
{
    std::vector<int> idx((size_t)n);
    srand(12345);
    auto rnd = [] { return (long long)rand() * 32768 + (rand() & 32767); };
    if (kind == 0) for (long long i = 0; i < n; ++i) idx[i] = (int)(rnd() % NB);
    else if (kind == 3) for (long long i = 0; i < n; ++i) idx[i] = (int)(i % NB);
    else {
        for (long long c = 0; c < n; c += 4096) {
            long long i = c, end = std::min(n, c + 4096);
            std::vector<std::pair<int, int>> runs;
            while (i < end) {
                int len = 2 << (rand() % 6);           // 2..64
                len = (int)std::min<long long>(len, end - i);
                long long st = rnd() % (NB - 64);
                for (int j = 0; j < len; ++j) idx[i + j] = (int)(st + j);
                i += len;
            }
            if (kind == 1) std::sort(idx.begin() + c, idx.begin() + end);
        }
    }
    return idx;
}

int main()
{
    const int rounds = 64;
    const long long n = (long long)CTAS * THREADS * rounds;
    int *d_idx; double *d_x, *d_y;
    cudaMalloc(&d_idx, n * 4); cudaMalloc(&d_x, NB * 8); cudaMalloc(&d_y, NB * 8);
    cudaMemset(d_x, 0, NB * 8); cudaMemset(d_y, 0, NB * 8);
    cudaFuncSetAttribute(k_ldgsts, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * THREADS * 8);
    cudaFuncSetAttribute(k_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * THREADS * 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char *pname[4] = {"random", "runs-sorted", "runs-unsorted", "contiguous"};
    for (int kind = 0; kind < 4; ++kind) {
        std::vector<int> h = make_pattern(kind, n);
        cudaMemcpy(d_idx, h.data(), n * 4, cudaMemcpyHostToDevice);
        for (int which = 0; which < 4; ++which) {
            float best = 1e30f;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0);
                if (which == 0) k_maponly<<<CTAS, THREADS>>>(d_idx, d_y, rounds);
                if (which == 1) k_red<<<CTAS, THREADS>>>(d_idx, d_y, rounds);
                if (which == 2) k_ldgsts<<<CTAS, THREADS, 8 * THREADS * 8>>>(d_idx, d_x, d_y, rounds);
                if (which == 3) k_ldg<<<CTAS, THREADS, 8 * THREADS * 8>>>(d_idx, d_x, d_y, rounds);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0) best = std::min(best, ms);
            }
            const char *wn[4] = {"maponly", "red", "ldgsts", "ldg+sts"};
            printf("%-14s %-8s %8.4f ms  %6.3f cycles/entry/SM (1.965 GHz)  %s\n", pname[kind], wn[which], best,
                   best * 1e-3 * 1.965e9 / ((double)rounds * THREADS), cudaGetErrorString(cudaGetLastError()));
        }
    }
    // shared-memory sweeps: 4536 points (e.g. 8 terms of 567), tiles of 9 / 15 / 3 doubles at strides 1, 3, 9, 15, 45, 135
    const int npts_tab[3] = {4536, 4725, 4374};
    cudaFuncSetAttribute(k_sweep<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5000 * 8);
    cudaFuncSetAttribute(k_sweep<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5000 * 8);
    cudaFuncSetAttribute(k_sweep<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 5000 * 8);
    for (int tile : {9, 15, 3})
        for (int stride : {1, 3, 9, 15, 27, 45, 135, 405}) {
            const int npts = npts_tab[tile == 9 ? 0 : (tile == 15 ? 1 : 2)];
            if ((npts / tile) % 1 != 0 || npts % (stride * tile) != 0) continue;
            const int R = 2000;
            float best = 1e30f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                if (tile == 9) k_sweep<9><<<CTAS, THREADS, 5000 * 8>>>(d_y, R, stride, npts);
                if (tile == 15) k_sweep<15><<<CTAS, THREADS, 5000 * 8>>>(d_y, R, stride, npts);
                if (tile == 3) k_sweep<3><<<CTAS, THREADS, 5000 * 8>>>(d_y, R, stride, npts);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep > 0) best = std::min(best, ms);
            }
            printf("sweep tile=%2d stride=%3d npts=%d: %8.4f ms  %6.3f cycles per point per (load+store) sweep pair per SM  %s\n", tile, stride, npts,
                   best, best * 1e-3 * 1.965e9 / ((double)R * npts), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
