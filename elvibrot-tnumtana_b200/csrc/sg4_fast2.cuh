// sg4_fast2.cuh -- second generation of the separable-KEO term kernel (sm_100a), single channel (nb0 = 1).
//
// Same mathematics and the same reference routines as sg4_fast.cuh (gather tabPackedBasis_TO_tabR_AT_iG, B->G
// BDP_TO_GDP_OF_SmolyakRep, operator sub_TabOpPsi_OF_ONEDP_FOR_SGtype4 type_Op=1, G->B GDP_TO_BDP_OF_SmolyakRep, weighted
// scatter tabR_AT_iG_TO_tabPackedBasis), but scheduled for the unit that bounds the first-generation kernel: the
// shared-memory / LSU data pipe (ncu, round 2: l1tex__data_pipe_lsu_wavefronts 63-70 % of peak, 31 % of them bank
// conflicts, issue slots 39 %).  Per grid point the old schedule moved 7G-3 doubles through shared memory for G tile
// groups plus ~7 more for staging the map / V / sorted scatter map and the scatter loop.  Here:
//   * the kinetic accumulator rides along the forward (B->G) passes:  pass g reads the psi tile of group g, transforms it
//     (psi' = B_g psi), and updates   acc <- B_g acc + T_g psi'   in the same round trip (the B_g of later groups bring
//     earlier kinetic contributions to the full grid; the 1-D operators of different modes commute).  The last forward
//     pass adds (V + shift) psi' and already applies its own G->B.  Sweeps: 3 + 4 (G-2) + 3 + 2 (G-1) = 6G-4 instead of 7G-3
//     for G >= 2 (G = 3: 14 instead of 18);
//   * V is read from global memory inside the last forward pass (tile elements of the slowest group: consecutive lanes
//     read consecutive doubles) -- no staging buffer, no shared-memory round trip;
//   * the last G->B pass scatters straight from registers: it loads the map entries of its tile (coalesced, L2-resident,
//     the same gather map) and issues the weighted FP64 reductions -- no sorted scatter map, no position array, no
//     scatter loop (they were 20 % of all instructions and 6 of the 32 bytes per point), and the HBM traffic per point
//     drops from 34 to 28 bytes;
//   * the gather reads the map with coalesced 32-bit loads and issues 8-byte LDGSTS with consecutive lanes on consecutive
//     entries (2 shared-memory wavefronts per instruction instead of 11.7, half the L2 sectors).
// Work items are batches of same-schedule terms (sg4_plan.cu); one persistent CTA per SM, thread groups as before.
#pragma once
#include "sg4_fast.cuh"

namespace evr {

enum { P2_FWD = 0, P2_LAST = 1, P2_G2B = 2, P2_FIN = 3 };

struct Pass2 {
    double *psi, *acc;          // shared-memory buffers of this item
    const double *pool;         // matrix pool in shared memory (MS == 1)
    const double *V;            // global: V slice of the item in the internal layout (nullptr: no potential grid)
    const int32_t *gmap;        // global: packed index of every entry of the item (internal layout), -1 = dropped
    double *y;                  // global: result vector of this right-hand side
    double vshift, weight;
    int m1, m2;                 // offsets of the [B|BTw|T] blocks in the pool
    int nq, stride;
    unsigned magic;
    int tid, nthr;
    int nored;                  // experiment switch: skip the reductions
};

// HA: the acc buffer already holds kinetic contributions of earlier groups.  S1: stride-1 group (compile-time addressing).
template <int N1, int N2, int KIND, int MS, bool HA, bool S1>
__device__ __forceinline__ void run_pass2(const Pass2 &A)
{
    constexpr int NN1 = N1 * N1, NN2 = N2 * N2, TILE = N1 * N2;
    const int ntiles = A.nq / TILE;
    const double *__restrict__ B1 = (MS == 2) ? c_iso + IsoOff<N1>::value : A.pool + A.m1, *__restrict__ W1 = B1 + NN1, *__restrict__ T1 = B1 + 2 * NN1;
    const double *__restrict__ B2 = (MS == 2) ? c_iso + IsoOff<(N2 > 1 ? N2 : N1)>::value : A.pool + ((N2 > 1) ? A.m2 : A.m1), *__restrict__ W2 = B2 + NN2, *__restrict__ T2 = B2 + 2 * NN2;
    const int stride = S1 ? 1 : A.stride;
    for (int t = A.tid; t < ntiles; t += A.nthr) {
        const int q0 = tile_origin(t, stride, A.magic, TILE);
        if (KIND == P2_FWD) {
            double v[N2][N1], a[N2][N1];
            tile_load<N1, N2>(v, A.psi + q0, stride);
            tile_xform<N1, N2, MS>(v, B1, B2);
            tile_store<N1, N2>(v, A.psi + q0, stride);
            if (HA) {
                tile_load<N1, N2>(a, A.acc + q0, stride);
                tile_xform<N1, N2, MS>(a, B1, B2);
            } else {
#pragma unroll
                for (int j = 0; j < N2; ++j)
#pragma unroll
                    for (int i = 0; i < N1; ++i) a[j][i] = 0.0;
            }
            tile_keo<N1, N2, MS>(a, v, T1, T2);
            tile_store<N1, N2>(a, A.acc + q0, stride);
        } else if (KIND == P2_LAST) {
            double v[N2][N1], a[N2][N1];
            const bool hv = A.V != nullptr;
            if (hv) {                                   // issued first: the global latency overlaps the B->G arithmetic
#pragma unroll
                for (int j = 0; j < N2; ++j)
#pragma unroll
                    for (int i = 0; i < N1; ++i) a[j][i] = __ldg(A.V + q0 + stride * (i + N1 * j));
            }
            tile_load<N1, N2>(v, A.psi + q0, stride);
            tile_xform<N1, N2, MS>(v, B1, B2);
#pragma unroll
            for (int j = 0; j < N2; ++j)
#pragma unroll
                for (int i = 0; i < N1; ++i) a[j][i] = hv ? (a[j][i] + A.vshift) * v[j][i] : A.vshift * v[j][i];
            tile_keo<N1, N2, MS>(a, v, T1, T2);
            if (HA) {                                   // kinetic contributions of the earlier groups -> full grid
                tile_load<N1, N2>(v, A.acc + q0, stride);
                tile_xform<N1, N2, MS>(v, B1, B2);
#pragma unroll
                for (int j = 0; j < N2; ++j)
#pragma unroll
                    for (int i = 0; i < N1; ++i) a[j][i] += v[j][i];
            }
            tile_xform<N1, N2, MS>(a, W1, W2);          // this group's G->B
            tile_store<N1, N2>(a, A.acc + q0, stride);
        } else if (KIND == P2_G2B) {
            double a[N2][N1];
            tile_load<N1, N2>(a, A.acc + q0, stride);
            tile_xform<N1, N2, MS>(a, W1, W2);
            tile_store<N1, N2>(a, A.acc + q0, stride);
        } else {                                        // P2_FIN: last G->B + weighted scatter-add from registers
            int m[N2][N1];
            double a[N2][N1];
#pragma unroll
            for (int j = 0; j < N2; ++j)
#pragma unroll
                for (int i = 0; i < N1; ++i) m[j][i] = __ldg(A.gmap + q0 + stride * (i + N1 * j));
            tile_load<N1, N2>(a, A.acc + q0, stride);
            tile_xform<N1, N2, MS>(a, W1, W2);
            if (!A.nored) {
#pragma unroll
                for (int j = 0; j < N2; ++j)
#pragma unroll
                    for (int i = 0; i < N1; ++i)
                        if (m[j][i] >= 0) atomicAdd(A.y + m[j][i], A.weight * a[j][i]);
            }
        }
    }
}

template <int KIND, int MS, bool HA, bool S1>
__device__ __forceinline__ void dispatch_pass2(const int tmpl, const Pass2 &A)
{
    if constexpr (MS == 2) {
        switch (tmpl) {
        case 1: run_pass2<3, 1, KIND, MS, HA, S1>(A); break;
        case 2: run_pass2<5, 1, KIND, MS, HA, S1>(A); break;
        case 3: run_pass2<7, 1, KIND, MS, HA, S1>(A); break;
        case 4: run_pass2<3, 3, KIND, MS, HA, S1>(A); break;
        case 11: run_pass2<9, 1, KIND, MS, HA, S1>(A); break;
        case 12: run_pass2<11, 1, KIND, MS, HA, S1>(A); break;
        case 20: run_pass2<3, 5, KIND, MS, HA, S1>(A); break;
        default: break;          // unreachable: the plan routes other tiles to the first-generation kernel
        }
    } else {
        switch (tmpl) {
#define X(id, a, b) case id: run_pass2<a, b, KIND, MS, HA, S1>(A); break;
            EVR_TMPL_LIST(X)
#undef X
        default: break;
        }
    }
}

#define EVR_V2_FUSE_MIN_STRIDE 8   // the scatter is fused into the last G->B pass when consecutive lanes cover runs of >= 8 entries

// dynamic smem:  pool[pool_len] (MS == 1) | per group: psi[cap] | acc[cap] | map[cap] (int32) | FastTermDev[2] | mbarrier[2]
//
// Software pipeline of one thread group over its items (all asynchronous parts overlap the passes of the same group):
//   item i, top    : gathered psi(i) has landed (cp.async wait) -> forward passes
//   after pass 0   : one thread starts the bulk copy (cp.async.bulk, TMA 1-D) of the gather map of item i+1 into the map buffer
//   after LAST(i)  : the psi buffer is dead -> every lane reads its entries of map(i+1) from shared memory and issues the
//                    8-byte LDGSTS gathers of psi(i+1) into it; they fly while the G->B passes and the scatter of item i run
#define EVR_V2_GROUP_BYTES(cap) ((size_t)(cap) * 20 + 2 * sizeof(evr::FastTermDev) + EVR_FAST_MBAR_BYTES)
// MAXT: threads per CTA the instantiation is compiled for (768: 80 registers per thread, 512: 128 registers)
template <int MS, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
sg4_term_kernel_v2(const FastPlanDev P, const FastClassDev Cc, const int npsi,
                   const double *__restrict__ psi, double *__restrict__ Hpsi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int gsize = Cc.gsize;
    const int ngrp = blockDim.x / gsize;
    const int group = threadIdx.x / gsize;
    const int tid = threadIdx.x - group * gsize;
    const int cap = Cc.cap;
    const int pool_doubles = (MS == 1) ? P.pool_len : 0;
    double *s_pool = reinterpret_cast<double *>(smem_raw);
    unsigned char *gbase = smem_raw + (size_t)pool_doubles * sizeof(double) + EVR_V2_GROUP_BYTES(cap) * group;
    double *const s_psi = reinterpret_cast<double *>(gbase);
    double *const s_acc = s_psi + cap;
    int *const s_map = reinterpret_cast<int *>(s_acc + cap);
    FastTermDev *const s_T0 = reinterpret_cast<FastTermDev *>(s_map + cap);
    unsigned long long *const s_bar = reinterpret_cast<unsigned long long *>(s_T0 + 2);

    if (MS == 1) {   // the whole (de-duplicated) 1-D matrix pool lives in shared memory for the kernel's lifetime
        for (int i = threadIdx.x; i < P.pool_len; i += blockDim.x) s_pool[i] = __ldg(P.mats + i);
        __syncthreads();
    }
    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    unsigned ph_map = 0;
    const int step = gridDim.x * ngrp;
    const FastTermDev *terms = P.terms + Cc.term_begin;
    // work item = (batch of terms, right-hand side): item w -> batch w / npsi, RHS w % npsi
    const int n_items = Cc.n_terms * npsi;
    const int w_first = blockIdx.x * ngrp + group;
    auto term_of = [&](const int w) { return (npsi == 1) ? w : (int)((unsigned)w / (unsigned)npsi); };
    auto issue_desc = [&](const int w, const int slot) {
        const double *src = reinterpret_cast<const double *>(terms + term_of(w));
        double *dst = reinterpret_cast<double *>(s_T0 + slot);
        for (int i = tid; i < (int)(sizeof(FastTermDev) / 8); i += gsize) cp_async8(dst + i, src + i);
    };
    auto issue_map = [&](const long long map_off, const int nq) {          // one thread; the map buffer is free
        const unsigned bytes = (unsigned)((nq + 31) & ~31) * 4u;
        fence_proxy_async();
        mbar_expect_tx(s_bar, bytes);
        bulk_g2s(s_map, P.gmap + map_off, bytes, s_bar);
    };
    // gather (tabPackedBasis_TO_tabR_AT_iG) of one item into the psi buffer: consecutive lanes, consecutive entries
    auto issue_gather = [&](const int nq, const double *x) {
        const bool nox = (P.dbg & 16) != 0;
        int e = tid;
        for (; e + 3 * gsize < nq; e += 4 * gsize) {
            int m[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) m[u] = s_map[e + u * gsize];
#pragma unroll
            for (int u = 0; u < 4; ++u) cp_async8_zfill(s_psi + e + u * gsize, x + max(m[u], 0), (m[u] >= 0 && !nox) ? 8 : 0);
        }
        for (; e < nq; e += gsize) {
            const int m = s_map[e];
            cp_async8_zfill(s_psi + e, x + max(m, 0), (m >= 0 && !nox) ? 8 : 0);
        }
    };
    if (w_first < n_items) {   // prologue: descriptor, map and gather of this group's first item
        issue_desc(w_first, 0);
        cp_async_commit_wait_all();
        group_sync(gsize, group);
        if (tid == 0) issue_map(s_T0->map_off, s_T0->nq);
        mbar_wait(s_bar, ph_map); ph_map ^= 1;
        issue_gather(s_T0->nq, psi + (long long)(w_first - term_of(w_first) * npsi) * P.nb);
    }
    int ts = 0;
    for (int w = w_first; w < n_items; w += step, ts ^= 1) {
        const int ip = w - term_of(w) * npsi;
        cp_async_commit_wait_all();            // the gathered psi of this item has landed
        group_sync(gsize, group);              // ... for every lane; and every lane is done with the previous item
        const bool has_next = w + step < n_items;
        long long nx_off = 0; int nx_nq = 0;
        if (has_next) {
            issue_desc(w + step, ts ^ 1);      // descriptor of the next item -> other slot, asynchronously
            if (tid == 0) {                    // what the bulk copy of its map needs (used after the first pass)
                const FastTermDev *Tn = terms + term_of(w + step);
                nx_off = __ldg(&Tn->map_off); nx_nq = __ldg(&Tn->nq);
            }
        }
        const FastTermDev *T = s_T0 + ts;
        const int G = T->ngroups, nq = T->nq;
        const int32_t *gm = P.gmap + T->map_off;
        Pass2 A;
        A.psi = s_psi; A.acc = s_acc; A.pool = s_pool; A.nq = nq; A.tid = tid; A.nthr = gsize;
        A.vshift = T->vshift; A.weight = T->weight;
        A.V = (P.has_V && !(P.dbg & 32)) ? P.V + T->grid_off : nullptr;
        A.gmap = gm;
        A.y = Hpsi + (long long)ip * P.nb;
        A.nored = (P.dbg & 8) ? 1 : 0;
        auto set_group = [&](const int g) {
            const FastGroup &Gr = T->g[g];
            A.stride = Gr.stride; A.magic = Gr.magic; A.m1 = Gr.mat1; A.m2 = Gr.mat2;
        };
        auto start_next_gather = [&]() {       // behind a group barrier that follows the last read of the psi buffer
            if (!has_next) return;
            const FastTermDev *Tn = s_T0 + (ts ^ 1);
            const int wn = w + step;
            mbar_wait(s_bar, ph_map); ph_map ^= 1;
            issue_gather(Tn->nq, psi + (long long)(wn - term_of(wn) * npsi) * P.nb);
        };
        bool scattered = false;
        if (G == 0 || (P.dbg & 4)) {           // a single grid point (all modes 1 x 1)
            if (tid == 0) s_acc[0] = (T->vshift + (A.V ? __ldg(A.V) : 0.0)) * s_psi[0];
            if (tid == 0 && has_next) issue_map(nx_off, nx_nq);
            cp_async_commit_wait_all();        // next descriptor
            group_sync(gsize, group);
            start_next_gather();
        } else {
            // forward passes: B -> G of every group; the kinetic accumulator is carried along
            for (int g = 0; g < G - 1; ++g) {
                set_group(g);
                if (g == 0) dispatch_pass2<P2_FWD, MS, false, true>(T->g[g].tmpl, A);
                else dispatch_pass2<P2_FWD, MS, true, false>(T->g[g].tmpl, A);
                if (g == 0 && tid == 0 && has_next) issue_map(nx_off, nx_nq);
                group_sync(gsize, group);
            }
            set_group(G - 1);                  // last group: + (V + shift) psi, its kinetic part and its own G -> B
            if (G == 1) dispatch_pass2<P2_LAST, MS, false, true>(T->g[0].tmpl, A);
            else dispatch_pass2<P2_LAST, MS, true, false>(T->g[G - 1].tmpl, A);
            if (G == 1 && tid == 0 && has_next) issue_map(nx_off, nx_nq);
            cp_async_commit_wait_all();        // the next item's descriptor has landed
            group_sync(gsize, group);
            start_next_gather();               // psi buffer is dead: the next item's gather overlaps the rest of this item
            // G -> B of the other groups (GDP_TO_BDP_OF_SmolyakRep); the last of them scatters from registers
            const bool fuse = G >= 2 && T->g[G - 2].stride >= EVR_V2_FUSE_MIN_STRIDE;
            for (int g = 0; g <= G - 2; ++g) {
                set_group(g);
                if (fuse && g == G - 2) { dispatch_pass2<P2_FIN, MS, false, false>(T->g[g].tmpl, A); scattered = true; }
                else {
                    if (g == 0) dispatch_pass2<P2_G2B, MS, false, true>(T->g[g].tmpl, A);
                    else dispatch_pass2<P2_G2B, MS, false, false>(T->g[g].tmpl, A);
                    group_sync(gsize, group);
                }
            }
        }
        if (!scattered && !A.nored) {
            // weighted scatter-add (tabR_AT_iG_TO_tabPackedBasis) in layout order: consecutive lanes, consecutive entries
            const double weight = T->weight;
            for (int e = tid; e < nq; e += gsize) {
                const int m = __ldg(gm + e);
                if (m >= 0) atomicAdd(A.y + m, weight * s_acc[e]);
            }
        }
    }
}

} // namespace evr
