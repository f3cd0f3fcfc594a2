// sg4_fast_types.h -- descriptors shared by the host plan (sg4_plan.cu) and the fast-path kernels
// (sg4_fast.cuh, instantiated in sg4_fast_inst.cu and sg4_iso.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sg4_internal.h"

namespace evr {

#define EVR_MAXG 8          // max groups (<= 16 active modes) per term on the fast path
#define EVR_MAX_FCLASSES 16 // (size class, kernel flavour) pairs = launches per H|psi> on the fast path
#define EVR_FAST_MBAR_BYTES 32 // per thread group: two mbarriers (bulk copies of the map / V slices) + two work-item indices
#define EVR_RT_NMAX 16      // runtime-size single-mode tiles keep up to 16 values in registers

struct FastGroup {
    int stride;             // stride of the first mode of the group (second: stride*n1)
    unsigned magic;         // floor(2^32/stride)+1 : exact t/stride by __umulhi for t*stride < 2^32 (stride > 1)
    unsigned short n1, n2;  // n2 = 0: single mode
    unsigned short tmpl;    // template id (0 = runtime single)
    unsigned short n3;      // > 0: three-mode cube tile (n1 = n2 = n3)
    int mat1, mat2, mat3;   // offsets (doubles) of the [B|BTw|T] blocks of the modes in the matrix pool
};

struct FastTermDev {
    long long map_off, grid_off;
    double weight;          // WeightSG * prod_{1x1 modes} B(0,0) BTw(0,0)
    double vshift;          // sum_{1x1 modes} T(0,0) (+ constant (0,0) term)
    int nq, ngroups;
    long long next_map_off, next_grid_off;   // the term this thread group processes next
    long long next2_map_off;                 // ... and the one after it (software pipeline)
    int next_nq, next2_nq;
    FastGroup g[EVR_MAXG];
};

struct FastClassDev {       // one launch per size class: terms [term_begin, term_begin+n_terms)
    int term_begin, n_terms;
    int gsize;              // threads cooperating on one term: 32, 64 or 128
    int rt;                 // 1: runtime-size tiles (RT instantiation of the kernel)
    int tri;                // 1: terms with three-mode cube tiles (TRI instantiation, 512 threads)
    int cap;                // doubles per psi/acc buffer (max nq*nb0 of the class)
    int cta_threads;        // threads per CTA = groups per CTA * gsize (<= 768)
    int *counter;           // work counter of this launch (zeroed before it); nullptr: static round-robin over the items
};

struct FastPlanDev {
    int nb0, n_terms, has_V;
    int pool_len;           // doubles in the (de-duplicated) matrix pool
    int dbg;                // experiment switch (EVR_SG4_DEBUG): 4 = skip the transform passes
    long long nb, NQ_local;
    const FastTermDev *terms;
    const int32_t *gmap;    // per term (slices padded to 32 entries): packed index of each term-local entry, internal layout order
    const int32_t *map;     // per term: the same indices sorted ascending (-1 = dropped / padding)
    const uint16_t *pos;    // per term: term-local position (internal layout) of each sorted entry
    const double *mats;     // pool of [B|BTw|T] blocks
    const double *V;        // [nb0*nb0][NQ_local] permuted to the internal layout
    // deterministic mode (EVR_SG4_DETERMINISTIC=1): the scatter writes every weighted entry to stage[(rhs*nb0+c)*stage_ld + entry]
    // instead of an FP64 reduction; a second kernel sums the entries of every packed element in a fixed order
    double *stage;
    long long stage_ld;
};


// template ids of the register tiles (host side uses the same table, sg4_plan.cu: fast_template_id)
#define EVR_TMPL_LIST(X) X(1, 3, 1) X(2, 5, 1) X(3, 7, 1) X(4, 3, 3) X(7, 2, 1) X(8, 2, 3) X(9, 4, 1) X(10, 2, 2) \
    X(11, 9, 1) X(12, 11, 1) X(13, 13, 1) X(14, 15, 1) X(15, 6, 1) X(16, 8, 1)
// additional two-mode tiles of the constant-matrix ("iso") instantiation, sg4_iso.cu
#define EVR_TMPL_LIST_ISO(X) X(20, 3, 5) X(21, 3, 7) X(22, 3, 9) X(23, 5, 5)

#define EVR_TMPL_CUBE3 30    // template id of the 3x3x3 cube tile (only in the TRI instantiation of the kernel)
#define EVR_TMPL_CUBE2 31    // 2x2x2

// 512 threads per CTA = 128 registers per thread: the tile passes compile without spills (768 threads / 80 registers:
// 140-870 bytes of spill traffic per thread that misses the ~6 KB of L1 left beside 222 KB of shared memory; measured
// 0.398 ms at 768, 0.375 at 640 / 96 registers, 0.361 at 512, profiles/r2/sweep16_threads_per_cta.txt)
#ifndef EVR_FAST_MAX_THREADS
#define EVR_FAST_MAX_THREADS 512
#endif
#define EVR_FAST_MAX_THREADS_TRI 512   // cube tiles keep 27 values + a 3x3 matrix in registers: 128 registers per thread
#define EVR_ISO_MAX_THREADS 512        // iso kernel: 128 registers per thread (two 27-value tiles in the fused passes)

// ---- launchers (the kernels are instantiated in their own translation units) ---------------------
// mm: 0 = matrix pool read from global memory, 1 = pool resident in shared memory
int fast_set_attributes();
int fast_launch(int mm, bool rt, bool tri, int nctas, int nthr, size_t smem, cudaStream_t st,
                const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi);
int fast_permute(bool in, const int32_t *perm, long long nb, int nvecs, const double *src, double *dst, cudaStream_t st);
int fast_permute_x(bool in, const int32_t *perm_or_inv, long long nb, int nvecs, const double *src, double *dst, double *zero, cudaStream_t st);
// iso flavour: one [B|BTw|T] block per mode size n = 2..EVR_ISO_NMAX at compile-time offsets of a __constant__ array
#define EVR_ISO_NMAX 15
__host__ __device__ constexpr int iso_off(int n) { int o = 0; for (int m = 2; m < n; ++m) o += 3 * m * m; return o; }
#define EVR_ISO_LEN (evr::iso_off(EVR_ISO_NMAX + 1))
int iso_set_attributes();
// makes `blocks` (EVR_ISO_LEN doubles, owned by plan `id`) the content of the device's __constant__ array (stream-ordered)
int iso_bind(int device, int id, const double *blocks, cudaStream_t st);
// big_tiles: the 512-thread instantiation with the 3x3x3 / 5x5 / 3x7 / 3x9 tiles, else the 768-thread one (tiles <= 15 values)
int iso_launch(bool big_tiles, int nctas, int nthr, size_t smem, cudaStream_t st,
               const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi);

// second-generation kernel (sg4_fast2.cuh): nb0 = 1, iso flavour (sg4_v2_iso.cu) or shared-memory pool (sg4_v2_pool.cu)
int v2_iso_set_attributes();
int v2_iso_bind(int device, int id, const double *blocks, cudaStream_t st);
int v2_iso_launch(int nctas, int nthr, size_t smem, cudaStream_t st,
                  const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi);
int v2_pool_set_attributes();
int v2_pool_launch(int nctas, int nthr, size_t smem, cudaStream_t st,
                   const FastPlanDev &P, const FastClassDev &C, int npsi, const double *psi, double *Hpsi);

} // namespace evr
