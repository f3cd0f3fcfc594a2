// sg4_plan.cu -- host side of the evr_sg4 C-ABI: plan construction (device-resident copy of
// param_SGType2 / tab_basisPrimSG / OpGrid for a term range), launch logic, host<->device staging.
//
// Replaces the body of sub_TabOpPsi_FOR_SGtype4 (sub_Operator/sub_OpPsi_SG4.f90:678-979) and of
// Action_MPI_S1 minus its reduce (sub_OpPsi_SG4_MPI.f90:454-571).  No CPU fallback exists: every
// entry point fails with a message if CUDA is unavailable.
#include "sg4_plan.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <climits>
#include <map>
#include <numeric>
#include <string>
#include <vector>

namespace evr {
static thread_local std::string g_err;
int fail(const std::string &msg) { g_err = msg; return 1; }
}
using evr::fail;

#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));                  \
    } while (0)

extern "C" int evr_sg4_version(void) { return 100; }
extern "C" const char *evr_sg4_last_error(void) { return evr::g_err.c_str(); }

// deterministic mode: entry lists per packed element from a mapping array (value = 0-based packed index, < 0: none)
static int build_entry_lists(evr_sg4_plan *p, const std::vector<int32_t> &map0, int64_t n_entries);

template <class T>
static int upload(T **dptr, const T *h, size_t n)
{
    if (n == 0) n = 1;
    CUDA_TRY(cudaMalloc((void **)dptr, n * sizeof(T)));
    if (h) CUDA_TRY(cudaMemcpy(*dptr, h, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
}

// Size classes of the generic kernel: CTA threads 256 / 128 / 64 / 32 for terms of > 768 / > 192 / > 48 / <= 48
// values.  Classes of larger terms get a third shared-memory buffer for the mixed-derivative sweeps
// (want_cache: the operator has mixed terms) when it fits.  Class 0 ("big") holds the terms whose two work buffers do
// not fit in shared memory at all: their buffers live in global memory (sg4_kernels.cuh, BIG instantiation), so that a
// term may be as large as the reference allows (its RDP arrays are heap-allocated, no limit).
#define EVR_GEN_BIG_BYTES (200 * 1024)
static int gen_class_of(int64_t tsize, int nb0)
{
    if (tsize * nb0 * 2 * (int64_t)sizeof(double) > EVR_GEN_BIG_BYTES) return 0;
    static const int th0 = getenv("EVR_SG4_GTH0") ? atoi(getenv("EVR_SG4_GTH0")) : 768;
    static const int th1 = getenv("EVR_SG4_GTH1") ? atoi(getenv("EVR_SG4_GTH1")) : 192;
    static const int th2 = getenv("EVR_SG4_GTH2") ? atoi(getenv("EVR_SG4_GTH2")) : 48;
    return tsize > th0 ? 1 : (tsize > th1 ? 2 : (tsize > th2 ? 3 : 4));
}
// One set of generic-kernel launches: the size classes of either all terms of the plan (list == nullptr: contiguous
// ranges of the work order) or of the terms listed by work-order index (the remainder of a fast-path plan).
struct GenSet {
    int *n; evr::GenClassDev *cls; int *threads; int *occ; size_t *smem; bool *big; double **scratch; int **d_list;
};
static int gen_configure_set(evr_sg4_plan *p, bool want_cache, int n_opterms, const std::vector<int> *list, GenSet G)
{
    static const int class_threads[5] = {256, 256, 128, 64, 32};
    const int nT = p->D * (p->LG + 1);
    int cache_classes = 5;                                  // every class (measured best: profiles/shape_bench_r1_generic_v2.txt)
    if (getenv("EVR_SG4_DCACHE")) cache_classes = atoi(getenv("EVR_SG4_DCACHE")) + 1;
    *G.n = 0;
    size_t smem_max = 0;
    int w0 = 0;
    const int n_all = list ? (int)list->size() : p->n_terms;
    auto tsize_at = [&](int w) { return p->h_tsize[p->order[list ? (*list)[w] : w]]; };
    const size_t ints = (size_t)EVR_GEN_SMEM_INTS(nT, p->D, n_opterms) * sizeof(int);
    if (ints > 40 * 1024) return fail("evr_sg4: too many operator terms / levels for the per-CTA tables");
    if (*G.scratch) { cudaFree(*G.scratch); *G.scratch = nullptr; }
    if (G.d_list && *G.d_list) { cudaFree(*G.d_list); *G.d_list = nullptr; }
    if (list && n_all > 0 && upload(G.d_list, list->data(), list->size())) return 1;
    for (int c = 0; c < 5; ++c) {
        int w1 = w0;
        int64_t ccap = 1;
        while (w1 < n_all && gen_class_of(tsize_at(w1), p->nb0) == c) { ccap = std::max(ccap, tsize_at(w1) * p->nb0); ++w1; }
        if (w1 == w0) continue;
        const int g = (*G.n)++;
        const bool big = (c == 0);
        bool dc = want_cache && c < cache_classes && (big || (size_t)3 * ccap * sizeof(double) + ints <= 227 * 1024);
        G.cls[g].term_begin = w0; G.cls[g].n_terms = w1 - w0; G.cls[g].cap = (int)ccap; G.cls[g].dcache = dc ? 1 : 0;
        G.cls[g].scratch = nullptr;
        G.cls[g].list = list ? *G.d_list : nullptr;
        G.threads[g] = class_threads[c];
        G.big[g] = big;
        if (big) {
            // global work buffers: as many CTAs as the scratch budget allows (EVR_SG4_SCRATCH_MB, default 4096), at most 4 per SM
            if (ccap >= (int64_t)1 << 30) return fail("evr_sg4: a Smolyak term has more than 2^30 values");
            const size_t per_cta = (size_t)(dc ? 3 : 2) * ccap * sizeof(double);
            const char *e = getenv("EVR_SG4_SCRATCH_MB");
            const size_t budget = (size_t)(e ? std::max(1, atoi(e)) : 4096) << 20;
            int ctas = (int)std::min<size_t>((size_t)p->sm_count * 4, std::max<size_t>(1, budget / per_cta));
            ctas = std::min(ctas, std::max(1, (w1 - w0) * 32));     // never more CTAs than a 32-vector block of these terms can use
            G.smem[g] = ints;
            G.occ[g] = -ctas;                                // negative: absolute CTA count, not per SM
            if (cudaMalloc((void **)G.scratch, per_cta * ctas) != cudaSuccess) {
                cudaGetLastError();
                return fail("evr_sg4: cannot allocate the work buffers of the terms that exceed shared memory (" +
                            std::to_string((per_cta * ctas) >> 20) + " MB)");
            }
            G.cls[g].scratch = *G.scratch;
        } else {
            G.smem[g] = (size_t)(dc ? 3 : 2) * ccap * sizeof(double) + ints;
            smem_max = std::max(smem_max, G.smem[g]);
        }
        w0 = w1;
    }
    if (smem_max > 227 * 1024) return fail("evr_sg4: shared-memory budget exceeded");
    // the attribute belongs to the function, not to the plan: never lower it under another live plan
    static size_t attr_max[64] = {0};
    size_t &amax = attr_max[p->device & 63];
    amax = std::max(amax, smem_max);
    if (cudaFuncSetAttribute(evr::sg4_term_kernel_generic<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)amax) != cudaSuccess)
        return fail("evr_sg4: cudaFuncSetAttribute(smem) failed");
    for (int g = 0; g < *G.n; ++g) {
        if (G.big[g]) continue;
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, evr::sg4_term_kernel_generic<false>, G.threads[g], G.smem[g]) != cudaSuccess || occ < 1)
            return fail("evr_sg4: generic kernel cannot be resident (occupancy 0)");
        G.occ[g] = occ;
    }
    return 0;
}
static int gen_configure(evr_sg4_plan *p, bool want_cache, int n_opterms)
{
    GenSet G{&p->n_gclasses, p->gclass, p->gclass_threads, p->gclass_occ, p->gclass_smem, p->gclass_big, &p->d_gscratch, nullptr};
    if (gen_configure_set(p, want_cache, n_opterms, nullptr, G)) return 1;
    auto ctas_of = [&](int g) { return p->gclass_occ[g] < 0 ? -p->gclass_occ[g] : p->sm_count * p->gclass_occ[g]; };
    p->gen_ctas_max = p->n_gclasses ? ctas_of(0) : p->sm_count;
    p->grid_ctas = p->n_gclasses ? std::max(1, std::min(p->gclass[0].n_terms, p->gen_ctas_max)) : 1;
    return 0;
}
// the terms a fast-path plan cannot take (mode sizes beyond its register tiles, terms beyond its shared-memory budget)
static int gen_configure_rest(evr_sg4_plan *p, const std::vector<int> &work_list)
{
    GenSet G{&p->n_rclasses, p->rclass, p->rclass_threads, p->rclass_occ, p->rclass_smem, p->rclass_big, &p->d_rscratch, &p->d_rlist};
    return gen_configure_set(p, p->pd.n_sweeps > 0, p->n_opterms, &work_list, G);
}

extern "C" int evr_sg4_plan_create(evr_sg4_plan **out, int device,
                                   int D, int nb_SG, int nb0, int64_t nb, int LG,
                                   const int32_t *tab_l, const double *WeightSG,
                                   const int32_t *tab_nq, const int32_t *tab_nb,
                                   const int32_t *tab_iB,
                                   const int32_t *nq_of, const int32_t *nb_of,
                                   const double *B, const double *BTw, const double *D1, const double *D2,
                                   int iG_begin, int iG_end)
{
    // evr_sg4_set_devices(n > 1): a plan on the default device (device < 0) spans the first n devices
    if (device < 0 && evr::multi_devices() > 1)
        return evr::multi_create(out, D, nb_SG, nb0, nb, LG, tab_l, WeightSG, tab_nq, tab_nb, tab_iB, nq_of, nb_of, B, BTw, D1, D2,
                                 iG_begin, iG_end);
    return evr::plan_create_single(out, device, D, nb_SG, nb0, nb, LG, tab_l, WeightSG, tab_nq, tab_nb, tab_iB, nq_of, nb_of,
                                   B, BTw, D1, D2, iG_begin, iG_end);
}

// MPI scheme 1 of the reference keeps only the rank's slice of the mapping table
// (Mapping_table_allocate_MPI, sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4_MPI.f90:62-77: tab_iB_OF_SRep_TO_iB(bounds_MPI(1,id):bounds_MPI(2,id))):
// tab_iB points at global entry tab_iB_first (0-based) and holds tab_iB_len entries, which must cover the plan's term range.
extern "C" int evr_sg4_plan_create_ex(evr_sg4_plan **out, int device,
                                      int D, int nb_SG, int nb0, int64_t nb, int LG,
                                      const int32_t *tab_l, const double *WeightSG,
                                      const int32_t *tab_nq, const int32_t *tab_nb,
                                      const int32_t *tab_iB, int64_t tab_iB_first, int64_t tab_iB_len,
                                      const int32_t *nq_of, const int32_t *nb_of,
                                      const double *B, const double *BTw, const double *D1, const double *D2,
                                      int iG_begin, int iG_end)
{
    if (!tab_iB || !tab_nb || tab_iB_first < 0 || tab_iB_len < 0) return fail("evr_sg4_plan_create_ex: bad mapping-table slice");
    if (nb_SG < 1 || iG_begin < 0 || iG_end > nb_SG || iG_begin > iG_end) return fail("evr_sg4_plan_create_ex: bad term range");
    int64_t first = 0, last = 0;
    for (int iG = 0; iG < iG_end; ++iG) { if (iG < iG_begin) first += tab_nb[iG]; last += tab_nb[iG]; }
    if (first < tab_iB_first || last > tab_iB_first + tab_iB_len)
        return fail("evr_sg4_plan_create_ex: the mapping-table slice does not cover the term range [iG_begin, iG_end)");
    // the single-device builder only reads entries [first, last) of the (virtual) full table
    return evr_sg4_plan_create(out, device, D, nb_SG, nb0, nb, LG, tab_l, WeightSG, tab_nq, tab_nb, tab_iB - tab_iB_first,
                               nq_of, nb_of, B, BTw, D1, D2, iG_begin, iG_end);
}

extern "C" int evr_sg4_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int evr::plan_create_single(evr_sg4_plan **out, int device,
                            int D, int nb_SG, int nb0, int64_t nb, int LG,
                            const int32_t *tab_l, const double *WeightSG,
                            const int32_t *tab_nq, const int32_t *tab_nb,
                            const int32_t *tab_iB,
                            const int32_t *nq_of, const int32_t *nb_of,
                            const double *B, const double *BTw, const double *D1, const double *D2,
                            int iG_begin, int iG_end)
{
    if (!out || !tab_l || !WeightSG || !tab_nq || !tab_nb || !tab_iB || !nq_of || !nb_of || !B || !BTw || !D1 || !D2)
        return fail("evr_sg4_plan_create: null argument");
    if (D < 1 || D > EVR_MAXD) return fail("evr_sg4_plan_create: D must be in [1," + std::to_string(EVR_MAXD) + "]");
    if (nb0 < 1 || nb0 > EVR_MAXCH) return fail("evr_sg4_plan_create: nb0 must be in [1," + std::to_string(EVR_MAXCH) + "]");
    if (LG < 0 || LG > 254) return fail("evr_sg4_plan_create: LG out of range");
    if (nb_SG < 1 || nb < 1) return fail("evr_sg4_plan_create: empty basis");
    if (iG_begin < 0 || iG_end > nb_SG || iG_begin > iG_end) return fail("evr_sg4_plan_create: bad term range");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1)
        return fail("evr_sg4_plan_create: no CUDA device available (this library has no CPU fallback)");
    if (device < 0) CUDA_TRY(cudaGetDevice(&device));
    if (device >= ndev) return fail("evr_sg4_plan_create: device index out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return fail("evr_sg4_plan_create: built for sm_100a (Blackwell) only");

    auto *p = new evr_sg4_plan();
    p->device = device; p->D = D; p->nb_SG = nb_SG; p->nb0 = nb0; p->nb = nb; p->LG = LG;
    p->iG_begin = iG_begin; p->iG_end = iG_end; p->n_terms = iG_end - iG_begin;
    p->sm_count = prop.multiProcessorCount;
    p->deterministic = getenv("EVR_SG4_DETERMINISTIC") && atoi(getenv("EVR_SG4_DETERMINISTIC")) != 0;
    const int nT = D * (LG + 1);
    p->h_nq_of.assign(nq_of, nq_of + nT);
    p->h_nb_of.assign(nb_of, nb_of + nT);
    p->h_tab_l.assign(tab_l, tab_l + (size_t)nb_SG * D);
    p->h_tab_nq.assign(tab_nq, tab_nq + nb_SG);
    p->h_tab_nb.assign(tab_nb, tab_nb + nb_SG);

    // canonical table offsets
    std::vector<int32_t> offB(nT), offG(nT);
    int64_t ob = 0, og = 0;
    for (int i = 0; i < nT; ++i) {
        if (nq_of[i] < 1 || nb_of[i] < 1) { delete p; return fail("evr_sg4_plan_create: nq/nb < 1"); }
        offB[i] = (int32_t)ob; offG[i] = (int32_t)og;
        ob += (int64_t)nq_of[i] * nb_of[i];
        og += (int64_t)nq_of[i] * nq_of[i];
        if (ob >= ((int64_t)1 << 31) || og >= ((int64_t)1 << 31)) { delete p; return fail("evr_sg4_plan_create: 1-D tables too large"); }
    }
    // prefix sums over ALL terms (grid offsets refer to the full Smolyak grid)
    std::vector<int64_t> pre_nq(nb_SG + 1, 0), pre_nb(nb_SG + 1, 0);
    for (int iG = 0; iG < nb_SG; ++iG) {
        int64_t nq = 1, nbT = 1;
        for (int k = 0; k < D; ++k) {
            int l = tab_l[(size_t)iG * D + k];
            if (l < 0 || l > LG) { delete p; return fail("evr_sg4_plan_create: level out of range in tab_l"); }
            nq *= nq_of[k * (LG + 1) + l]; nbT *= nb_of[k * (LG + 1) + l];
        }
        if (nq != tab_nq[iG] || nbT != tab_nb[iG]) { delete p; return fail("evr_sg4_plan_create: tab_nq/nb_OF_SRep inconsistent with tab_l"); }
        pre_nq[iG + 1] = pre_nq[iG] + nq; pre_nb[iG + 1] = pre_nb[iG] + nbT;
    }
    p->NQ_total = pre_nq[nb_SG];
    p->grid_start = pre_nq[iG_begin];
    p->NQ_local = pre_nq[iG_end] - pre_nq[iG_begin];
    p->S_local = pre_nb[iG_end] - pre_nb[iG_begin];
    const int64_t map_start = pre_nb[iG_begin];
    for (int64_t j = 0; j < p->S_local; ++j) {
        int32_t m = tab_iB[map_start + j];
        if (m < 0 || m > nb) { delete p; return fail("evr_sg4_plan_create: mapping entry out of range"); }
    }

    // per-term cost, capacity, work order
    std::vector<double> cost(p->n_terms), wfold(p->n_terms);
    std::vector<int64_t> tsize(p->n_terms);
    int64_t cap = 1, flops = 0;
    for (int t = 0; t < p->n_terms; ++t) {
        const int iG = iG_begin + t;
        int64_t mx = 1, sumn = 0;
        double fold = WeightSG[iG];
        int64_t left = 1, right = tab_nb[iG];
        for (int k = 0; k < D; ++k) {
            int l = tab_l[(size_t)iG * D + k];
            int a = nq_of[k * (LG + 1) + l], b = nb_of[k * (LG + 1) + l];
            mx *= std::max(a, b);
            sumn += a + b;
            if (a == 1 && b == 1) fold *= B[offB[k * (LG + 1) + l]] * BTw[offB[k * (LG + 1) + l]];
            right /= b;
            flops += 2 * 2 * left * a * b * right;       // B->G and G->B (same count)
            left *= a;
        }
        cap = std::max(cap, mx);
        cost[t] = (double)tab_nq[iG] * (double)sumn;
        wfold[t] = fold; tsize[t] = mx;
    }
    p->flops_npsi1 = flops * nb0;
    if (cap * nb0 >= (int64_t)1 << 30) { delete p; return fail("evr_sg4_plan_create: a Smolyak term has more than 2^30 values"); }
    p->cap = (int)(cap * nb0);        // terms beyond the shared-memory budget run in the generic kernel's global-buffer class
    p->h_map.assign(tab_iB + map_start, tab_iB + map_start + p->S_local);
    p->h_map_off.resize(p->n_terms); p->h_grid_off.resize(p->n_terms);
    for (int t = 0; t < p->n_terms; ++t) {
        p->h_map_off[t] = pre_nb[iG_begin + t] - map_start;
        p->h_grid_off[t] = pre_nq[iG_begin + t] - p->grid_start;
    }
    p->h_B.assign(B, B + ob); p->h_BTw.assign(BTw, BTw + ob);
    p->h_D1.assign(D1, D1 + og); p->h_D2.assign(D2, D2 + og);
    p->h_weight.assign(WeightSG, WeightSG + nb_SG);
    p->h_offB = offB; p->h_offG = offG;
    p->h_cost = cost;
    p->order.resize(p->n_terms);
    std::iota(p->order.begin(), p->order.end(), 0);
    // size classes of the generic kernel (CTA threads 256 / 128 / 64 / 32), then cost descending inside a class
    auto gclass_of = [&](int t) { return gen_class_of(tsize[t], nb0); };
    std::stable_sort(p->order.begin(), p->order.end(), [&](int a, int b) {
        const int ca = gclass_of(a), cb = gclass_of(b);
        return ca != cb ? ca < cb : cost[a] > cost[b];
    });

    p->n_big_terms = 0; p->cap_small = 1;
    for (int t = 0; t < p->n_terms; ++t) {
        if (gclass_of(t) == 0) ++p->n_big_terms;
        else p->cap_small = std::max<int64_t>(p->cap_small, tsize[t] * nb0);
    }

    std::vector<evr::TermDev> terms(p->n_terms);
    std::vector<uint8_t> lev((size_t)p->n_terms * D);
    for (int w = 0; w < p->n_terms; ++w) {
        const int t = p->order[w], iG = iG_begin + t;
        evr::TermDev &T = terms[w];
        T.map_off = pre_nb[iG] - map_start;
        T.grid_off = pre_nq[iG] - p->grid_start;
        T.weight = WeightSG[iG];
        T.nbT = tab_nb[iG]; T.nq = tab_nq[iG];
        T.lev_off = w * D; T.pad = 0; T.wfold = wfold[t];
        for (int k = 0; k < D; ++k) lev[(size_t)w * D + k] = (uint8_t)tab_l[(size_t)iG * D + k];
    }
    int rc = 0;
    rc |= upload(&p->d_terms, terms.data(), terms.size());
    rc |= upload(&p->d_lev, lev.data(), lev.size());
    rc |= upload(&p->d_map, tab_iB + map_start, (size_t)p->S_local);
    rc |= upload(&p->d_nq_of, nq_of, (size_t)nT);
    rc |= upload(&p->d_nb_of, nb_of, (size_t)nT);
    rc |= upload(&p->d_offB, offB.data(), (size_t)nT);
    rc |= upload(&p->d_offG, offG.data(), (size_t)nT);
    rc |= upload(&p->d_B, B, (size_t)ob);
    rc |= upload(&p->d_BTw, BTw, (size_t)ob);
    rc |= upload(&p->d_D1, D1, (size_t)og);
    rc |= upload(&p->d_D2, D2, (size_t)og);
    if (rc) { evr_sg4_plan_destroy(&p); return 1; }
    if (cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking) != cudaSuccess) {
        evr_sg4_plan_destroy(&p); return fail("evr_sg4_plan_create: cudaStreamCreate failed");
    }
    if (!getenv("EVR_SG4_SINGLE_STREAM")) {
        bool ok_ev = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) == cudaSuccess;
        for (int c = 1; c < EVR_MAX_FCLASSES && ok_ev; ++c)
            ok_ev = cudaStreamCreateWithFlags(&p->side[c], cudaStreamNonBlocking) == cudaSuccess &&
                    cudaEventCreateWithFlags(&p->ev_join[c], cudaEventDisableTiming) == cudaSuccess;
        if (!ok_ev) { evr_sg4_plan_destroy(&p); return fail("evr_sg4_plan_create: side stream/event creation failed"); }
    }

    p->smem_bytes = (size_t)2 * p->cap * sizeof(double) + (size_t)(4 * nT + 5 * D) * sizeof(int);   // of the largest term (informational)
    p->h_tsize = tsize;
    if (gen_configure(p, false, 0)) { evr_sg4_plan_destroy(&p); return 1; }

    evr::PlanDev &pd = p->pd;
    pd.D = D; pd.LG = LG; pd.nb0 = nb0; pd.n_terms = p->n_terms; pd.nb = nb; pd.NQ_local = p->NQ_local;
    pd.cap = p->cap; pd.terms = p->d_terms; pd.lev = p->d_lev; pd.map = p->d_map;
    pd.nq_of = p->d_nq_of; pd.nb_of = p->d_nb_of; pd.offB = p->d_offB; pd.offG = p->d_offG;
    pd.B = p->d_B; pd.BTw = p->d_BTw; pd.D1 = p->d_D1; pd.D2 = p->d_D2;
    // DMMA mode products (sg4_kernels.cuh: mode_product_dmma) are opt-in: 2.3 x faster on the 80 x 80 mode in isolation, but
    // the HCN_UT terms offer only ~170 columns per matrix and the call is latency-bound: 122 vs 86 us per H|psi>, 228 vs 208 us
    // for a 27-vector block (profiles/r2/dmma_in_generic_kernel.txt)
    pd.use_dmma = (getenv("EVR_SG4_DMMA") && atoi(getenv("EVR_SG4_DMMA")) != 0) ? 1 : 0;
    *out = p;
    return 0;
}

// ------------------------------------------------------------------------------------------------
// Fast path set-up (sg4_fast.cuh): returns 0 and sets p->fast when the operator qualifies,
// returns 0 with p->fast == false when it does not (generic kernel is used), 1 on CUDA errors.
// ------------------------------------------------------------------------------------------------
static int fast_template_id(int n1, int n2, bool iso = false)
{
#define X(id, a, b) if (n1 == a && ((b == 1 && n2 == 0) || (b > 1 && n2 == b))) return id;
    EVR_TMPL_LIST(X)
    if (iso) { EVR_TMPL_LIST_ISO(X) }
#undef X
    return 0;
}
static bool fast_pair_supported(int a, int b)   // a <= b
{
    if (a < 2) return false;
    return fast_template_id(a, b) != 0;
}

static int build_fast_path(evr_sg4_plan *p, int nb_Term, const int32_t *term_mode,
                           const uint8_t *grid_zero, const uint8_t *grid_cte,
                           const double *Mat_cte, const double *const *grids)
{
    p->fast = false;
    p->n_rclasses = 0; p->n_rest_terms = 0;
    if (getenv("EVR_SG4_FORCE_GENERIC")) return 0;
    const int D = p->D, LG = p->LG, nb0 = p->nb0, nT = D * (LG + 1);
    for (int i = 0; i < nT; ++i) if (p->h_nq_of[i] != p->h_nb_of[i]) return 0;
    const int nterm = (p->type_Op == 0) ? 1 : nb_Term;
    std::vector<double> c1(D, 0.0), c2(D, 0.0);
    double c00 = 0.0;
    const double *Vgrid = nullptr;
    for (int it = 0; it < nterm; ++it) {
        if (grid_zero[it]) continue;
        int m1 = (p->type_Op == 0) ? 0 : std::max(0, (int)term_mode[2 * it]);
        int m2 = (p->type_Op == 0) ? 0 : std::max(0, (int)term_mode[2 * it + 1]);
        double c = 0.0;
        if (grid_cte[it]) {
            const double *M = Mat_cte + (size_t)it * nb0 * nb0;
            c = M[0];
            for (int i = 0; i < nb0; ++i)
                for (int j = 0; j < nb0; ++j)
                    if (M[i + nb0 * j] != ((i == j) ? c : 0.0)) return 0;   // must be c * identity
        }
        if (m1 == 0 && m2 == 0) {
            if (grid_cte[it]) c00 += c;
            else { if (Vgrid) return 0; Vgrid = grids[it]; }
        } else {
            if (!grid_cte[it]) return 0;                                     // variable KEO grid -> generic
            if (m1 != 0 && m2 != 0 && m1 != m2) return 0;                    // mixed derivative -> generic
            const int k = (m1 != 0 ? m1 : m2) - 1;
            if (m1 == m2) c2[k] += c; else c1[k] += c;
        }
    }
    // matrix pool [B|BTw|T] per (k,L), identical blocks stored once (e.g. 12 equal Hm modes)
    std::vector<int> moff(nT);
    std::vector<double> pool;
    {
        std::map<std::vector<double>, int> seen;
        for (int k = 0; k < D; ++k)
            for (int L = 0; L <= LG; ++L) {
                const int i = k * (LG + 1) + L, n = p->h_nq_of[i];
                const double *Bm = p->h_B.data() + p->h_offB[i], *Wm = p->h_BTw.data() + p->h_offB[i];
                const double *d1 = p->h_D1.data() + p->h_offG[i], *d2 = p->h_D2.data() + p->h_offG[i];
                std::vector<double> blk;
                blk.reserve((size_t)3 * n * n);
                blk.insert(blk.end(), Bm, Bm + n * n);
                blk.insert(blk.end(), Wm, Wm + n * n);
                for (int e = 0; e < n * n; ++e) blk.push_back(c2[k] * d2[e] + c1[k] * d1[e]);
                auto itf = seen.find(blk);
                if (itf != seen.end()) { moff[i] = itf->second; continue; }
                moff[i] = (int)pool.size();
                seen.emplace(blk, moff[i]);
                pool.insert(pool.end(), blk.begin(), blk.end());
            }
    }
    if (pool.size() & 1) pool.push_back(0.0);                  // keeps the shared-memory buffers behind the pool 16-byte aligned
    const bool pool_in_smem = pool.size() * sizeof(double) <= 24 * 1024;
    // iso flavour (sg4_iso.cu): every mode size n >= 2 has exactly ONE [B|BTw|T] block (all modes of that size share
    // it, e.g. D identical Gauss-Hermite modes) -> the blocks go to a __constant__ array at compile-time offsets
    bool iso = !getenv("EVR_SG4_ISO") || atoi(getenv("EVR_SG4_ISO")) != 0;
    const bool iso_big = iso && getenv("EVR_SG4_ISO") && atoi(getenv("EVR_SG4_ISO")) == 2;   // 512-thread instantiation with the large tiles
    std::vector<double> iso_blocks((size_t)EVR_ISO_LEN, 0.0);
    {
        std::vector<int> off_of_n(EVR_ISO_NMAX + 1, -1);
        bool any = false;
        for (int i = 0; i < nT && iso; ++i) {
            const int n = p->h_nq_of[i];
            if (n < 2) continue;
            if (n > EVR_ISO_NMAX) continue;                 // such a mode never runs in the constant-matrix instantiation
            if (off_of_n[n] < 0) off_of_n[n] = moff[i];
            else if (off_of_n[n] != moff[i]) iso = false;
            any = true;
        }
        if (!any) iso = false;
        if (iso)
            for (int n = 2; n <= EVR_ISO_NMAX; ++n)
                if (off_of_n[n] >= 0) std::copy(pool.begin() + off_of_n[n], pool.begin() + off_of_n[n] + 3 * n * n, iso_blocks.begin() + evr::iso_off(n));
    }
    // per-term schedules + permutation to the internal layout
    // size classes: how many threads cooperate on one term (tiles per pass ~ nq/9 .. nq/21)
    // a term whose active mode sizes all have single-mode templates uses the templated kernel; any other size
    // (<= EVR_RT_NMAX) sends the whole term to the runtime-size instantiation (classes 3..5)
    // iso plans: a term runs in the constant-matrix instantiation when its active mode sizes are 3, 5, 7, 9 or 11; the
    // few other terms of such a plan use the pool-based instantiations like the terms of a non-iso plan
    // term_gen: the terms this path cannot take -- a mode beyond the register tiles (n > EVR_RT_NMAX), more tile groups
    // than a descriptor holds, or psi + acc buffers beyond the shared memory of an SM.  They run in the generic kernel
    // (gen_configure_rest) on the caller's vectors; the plan stays on the fast path for everything else.
    std::vector<char> term_rt(p->n_terms, 0), term_iso(p->n_terms, 0), term_gen(p->n_terms, 0);
    const int64_t fast_cap_max = (int64_t)((227 * 1024 - (pool_in_smem ? pool.size() * sizeof(double) : 0) - 2 * sizeof(evr::FastTermDev) - EVR_FAST_MBAR_BYTES) /
                                           ((nb0 == 1 && getenv("EVR_SG4_V2") && atoi(getenv("EVR_SG4_V2")) != 0) ? 20 : 16)) - 64;
    for (int t = 0; t < p->n_terms; ++t) {
        const int iG = p->iG_begin + t;
        if ((int64_t)p->h_tab_nq[iG] * nb0 > fast_cap_max || p->h_tab_nq[iG] > 65535) { term_gen[t] = 1; continue; }
        int c3 = 0, c79 = 0, cbad = 0, nact = 0;
        for (int k = 0; k < D; ++k) {
            const int n = p->h_nq_of[k * (LG + 1) + p->h_tab_l[(size_t)iG * D + k]];
            if (n < 2) continue;
            ++nact;
            if (n == 3) ++c3; else if (n == 7 || (n == 9 && iso_big)) ++c79; else if (n != 5 && n != 9 && n != 11) ++cbad;
        }
        term_iso[t] = iso && nact > 0 && cbad == 0;
        if (term_iso[t]) continue;
        for (int k = 0; k < D; ++k) {
            const int n = p->h_nq_of[k * (LG + 1) + p->h_tab_l[(size_t)iG * D + k]];
            if (n > 1 && fast_template_id(n, 0) == 0) { term_rt[t] = 1; if (n > EVR_RT_NMAX) term_gen[t] = 1; }
        }
    }
    // ---- tunables (experiments: environment overrides) ---------------------------------------------------------
    auto envi = [](const char *n, int d) { const char *v = getenv(n); return v ? atoi(v) : d; };
    // cube tiles (three equal modes of size 3 or 2 per thread): how many cubes minimise the group count of a term
    const bool use_cubes = iso_big || (!iso && envi("EVR_SG4_CUBES", 0) != 0);
    auto n_cubes = [](int c) { return (c == 3 || c == 5 || c == 6) ? c / 3 : (c >= 7 ? (c - 4) / 3 + 1 : 0); };
    std::vector<char> term_tri(p->n_terms, 0);
    if (use_cubes)
        for (int t = 0; t < p->n_terms; ++t) {
            if (term_rt[t] || (iso && !term_iso[t])) continue;
            const int iG = p->iG_begin + t;
            int c3 = 0, c2 = 0;
            for (int k = 0; k < D; ++k) {
                const int n = p->h_nq_of[k * (LG + 1) + p->h_tab_l[(size_t)iG * D + k]];
                c3 += (n == 3); c2 += (n == 2);
            }
            if (iso_big || n_cubes(c3) > 0 || n_cubes(c2) > 0) term_tri[t] = 1;   // iso with large tiles: every term runs in the cube-capable instantiation
        }
    // flavour = which instantiation of the kernel runs a term:
    //   0 templated tiles, matrices from the pool | 1 runtime-size tiles | 2 cube tiles (512 threads) | 3 iso (768 threads)
    auto flavour_of = [&](int t) { return term_rt[t] ? 1 : (term_tri[t] ? 2 : (term_iso[t] ? 3 : 0)); };
    // ---- internal order of the packed vector: functions reached by exactly the same set of Smolyak terms
    // (= same level vector) are stored contiguously, so that every term reads/updates whole blocks and the
    // gather/scatter of a warp touches a few contiguous runs instead of 32 separate sectors.
    // The membership signature is a sum of per-term 64-bit hashes; no multi-index table is needed.
    // The two permutation kernels cost ~12 us per H|psi>, so small problems keep the caller's order (identity
    // permutation, term entries still sorted by address); EVR_SG4_BLOCK_ORDER=0/1 overrides the size heuristic.
    // Measured at HH-12D L=7 on 1/2/4/8 GPUs (profiles/r2/scale_r2_first.txt): the block order pays down to ~6 M grid points
    // per rank (N <= 4) and costs 17 us of 183 us at N = 8 (3 M points per rank), where the two permutation kernels over the
    // full vector no longer amortise -- hence the threshold on the points of THIS plan's term range.
    std::vector<int32_t> inv_perm((size_t)p->nb), perm((size_t)p->nb);
    bool block_order = p->NQ_local >= 4000000;
    if (getenv("EVR_SG4_BLOCK_ORDER")) block_order = atoi(getenv("EVR_SG4_BLOCK_ORDER")) != 0;
    p->fast_block_order = block_order;
    if (!block_order) {
        std::iota(perm.begin(), perm.end(), 0);
        std::iota(inv_perm.begin(), inv_perm.end(), 0);
    } else {
        std::vector<uint64_t> sig((size_t)p->nb, 0);
        for (int t = 0; t < p->n_terms; ++t) {
            uint64_t h = (uint64_t)(p->iG_begin + t + 1) * 0x9E3779B97F4A7C15ull;
            h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
            const int32_t *m = p->h_map.data() + p->h_map_off[t];
            const int n = p->h_tab_nb[p->iG_begin + t];
            for (int j = 0; j < n; ++j) if (m[j] > 0) sig[m[j] - 1] += h;
        }
        std::iota(perm.begin(), perm.end(), 0);
        std::stable_sort(perm.begin(), perm.end(), [&](int32_t a, int32_t b) { return sig[a] < sig[b]; });
        for (int64_t i = 0; i < p->nb; ++i) inv_perm[perm[i]] = (int32_t)i;
    }
    // ---- per-term schedule: active modes -> register-tile groups, internal (permuted) term layout ---------------
    const int max35 = (nb0 == 1 && envi("EVR_SG4_V2", 0) != 0) ? envi("EVR_SG4_MAX35", 1) : 1000;
    struct TermSched {
        evr::FastTermDev F;                     // nq = points of ONE term; offsets filled per batch below
        std::vector<int> in_n, in_ref;          // per internal mode: size, stride in the reference term layout
    };
    std::vector<TermSched> sched(p->n_terms);
#pragma omp parallel for schedule(dynamic, 64)
    for (int t = 0; t < p->n_terms; ++t) {
        if (term_gen[t]) continue;
        const int iG = p->iG_begin + t;
        evr::FastTermDev &F = sched[t].F;
        std::memset(&F, 0, sizeof(F));
        F.nq = p->h_tab_nq[iG];
        double wgt = p->h_weight[iG], shift = c00;
        // active modes (size > 1), sorted by size
        struct Act { int k, n, refstride; };
        std::vector<Act> act;
        int refstride = 1;
        for (int k = 0; k < D; ++k) {
            const int l = p->h_tab_l[(size_t)iG * D + k], i = k * (LG + 1) + l, n = p->h_nq_of[i];
            if (n == 1) {
                const double *blk = pool.data() + moff[i];
                wgt *= blk[0] * blk[1];
                shift += blk[2];
            } else act.push_back({k, n, refstride});
            refstride *= n;
        }
        std::stable_sort(act.begin(), act.end(), [](const Act &a, const Act &b) { return a.n < b.n; });
        struct Grp { int a1, a2, a3; };        // indices into act; a2 = -1 single; a3 >= 0 cube
        std::vector<Grp> grp;
        std::vector<char> used(act.size(), 0);
        if (term_iso[t]) {
            // tiles of the iso instantiation: 3x3x3, 5x5, 3x5, 3x7, 3x9, 3x3, singles.  Pair every 9 and 7 with a 3,
            // the 5s with each other (a left-over 5 with a 3), then split the remaining 3s into cubes and pairs.
            std::vector<int> i3, i5, i79, other;
            for (size_t a = 0; a < act.size(); ++a) {
                const int n = act[a].n;
                (n == 3 ? i3 : n == 5 ? i5 : (n == 7 || n == 9) ? i79 : other).push_back((int)a);
            }
            for (int a : i79) {
                if (iso_big && !i3.empty()) { grp.push_back({i3.back(), a, -1}); i3.pop_back(); }
                else grp.push_back({a, -1, -1});
            }
            while (iso_big && i5.size() >= 2) { grp.push_back({i5[i5.size() - 2], i5[i5.size() - 1], -1}); i5.pop_back(); i5.pop_back(); }
            // second-generation kernel: the 15-value tile only as the first (stride-1) group -- the later forward passes keep
            // two tiles in registers (psi and the carried kinetic accumulator) and must stay within 80 registers
            int n35 = 0;
            while (!iso_big && !i5.empty() && !i3.empty() && (n35 < max35)) { grp.push_back({i3.back(), i5.back(), -1}); i3.pop_back(); i5.pop_back(); ++n35; }
            while (!iso_big && !i5.empty()) { grp.push_back({i5.back(), -1, -1}); i5.pop_back(); }
            if (!i5.empty()) {
                if (!i3.empty()) { grp.push_back({i3.back(), i5[0], -1}); i3.pop_back(); }
                else grp.push_back({i5[0], -1, -1});
            }
            int r = (int)i3.size();
            int npairs = iso_big ? ((r % 3 == 1 && r >= 4) ? 2 : (r % 3 == 2 ? 1 : 0)) : r / 2;
            int ncube = iso_big ? (r - 2 * npairs) / 3 : 0;
            int at = 0;
            for (int c = 0; c < ncube; ++c, at += 3) grp.push_back({i3[at], i3[at + 1], i3[at + 2]});
            for (int c = 0; c < npairs; ++c, at += 2) grp.push_back({i3[at], i3[at + 1], -1});
            for (; at < r; ++at) grp.push_back({i3[at], -1, -1});
            for (int a : other) grp.push_back({a, -1, -1});
            std::fill(used.begin(), used.end(), 1);
        } else if (term_tri[t]) {                     // cubes of equal size-3 (or size-2) modes first
            for (int sz : {3, 2}) {
                std::vector<int> idxs;
                for (size_t a = 0; a < act.size(); ++a) if (act[a].n == sz) idxs.push_back((int)a);
                int nc = n_cubes((int)idxs.size());
                for (int c = 0; c < nc; ++c) {
                    grp.push_back({idxs[3 * c], idxs[3 * c + 1], idxs[3 * c + 2]});
                    used[idxs[3 * c]] = used[idxs[3 * c + 1]] = used[idxs[3 * c + 2]] = 1;
                }
            }
        }
        {
            std::vector<int> rest;
            for (size_t a = 0; a < act.size(); ++a) if (!used[a]) rest.push_back((int)a);
            int lo = 0, hi = (int)rest.size() - 1;
            while (lo <= hi) {
                if (!term_rt[t] && lo < hi && fast_pair_supported(act[rest[lo]].n, act[rest[hi]].n)) { grp.push_back({rest[lo], rest[hi], -1}); ++lo; --hi; }
                else { grp.push_back({rest[hi], -1, -1}); --hi; }
            }
        }
        auto gsize = [&](const Grp &g) { return act[g.a1].n * (g.a2 >= 0 ? act[g.a2].n : 1) * (g.a3 >= 0 ? act[g.a3].n : 1); };
        std::stable_sort(grp.begin(), grp.end(), [&](const Grp &a, const Grp &b) { return gsize(a) > gsize(b); });
        if ((int)grp.size() > EVR_MAXG) { term_gen[t] = 1; continue; }
        F.ngroups = (int)grp.size();
        F.weight = wgt; F.vshift = shift;
        // internal mode order and strides
        std::vector<int> &in_n = sched[t].in_n, &in_ref = sched[t].in_ref;
        int stride = 1;
        for (size_t g = 0; g < grp.size(); ++g) {
            const Act &A1 = act[grp[g].a1];
            evr::FastGroup &Gd = F.g[g];
            Gd.stride = stride;
            Gd.magic = (stride > 1) ? (unsigned)((((uint64_t)1) << 32) / (uint64_t)stride + 1) : 0u;
            Gd.n1 = (unsigned short)A1.n;
            const int l1 = p->h_tab_l[(size_t)iG * D + A1.k];
            Gd.mat1 = moff[A1.k * (LG + 1) + l1];
            in_n.push_back(A1.n); in_ref.push_back(A1.refstride);
            stride *= A1.n;
            if (grp[g].a2 >= 0) {
                const Act &A2 = act[grp[g].a2];
                Gd.n2 = (unsigned short)A2.n;
                const int l2 = p->h_tab_l[(size_t)iG * D + A2.k];
                Gd.mat2 = moff[A2.k * (LG + 1) + l2];
                in_n.push_back(A2.n); in_ref.push_back(A2.refstride);
                stride *= A2.n;
            } else { Gd.n2 = 0; Gd.mat2 = Gd.mat1; }
            Gd.n3 = 0; Gd.mat3 = Gd.mat1;
            if (grp[g].a3 >= 0) {
                const Act &A3 = act[grp[g].a3];
                Gd.n3 = (unsigned short)A3.n;
                const int l3 = p->h_tab_l[(size_t)iG * D + A3.k];
                Gd.mat3 = moff[A3.k * (LG + 1) + l3];
                in_n.push_back(A3.n); in_ref.push_back(A3.refstride);
                stride *= A3.n;
            }
            Gd.tmpl = term_rt[t] ? 0 : (Gd.n3 == 3 ? EVR_TMPL_CUBE3 : (Gd.n3 == 2 ? EVR_TMPL_CUBE2 : (unsigned short)fast_template_id(Gd.n1, Gd.n2, term_iso[t] != 0)));
            if (!term_rt[t] && Gd.tmpl == 0) term_gen[t] = 1;
        }
    }
    {
        std::vector<int> rest;                               // work-order indices (p->order is sorted by generic size class)
        for (int w = 0; w < p->n_terms; ++w) if (term_gen[p->order[w]]) rest.push_back(w);
        if ((int)rest.size() == p->n_terms) return 0;
        if (!rest.empty() && (p->deterministic || (getenv("EVR_SG4_MIXED") && atoi(getenv("EVR_SG4_MIXED")) == 0))) return 0;   // one kernel family stages the entries
        p->n_rest_terms = (int)rest.size();
        if (!rest.empty() && gen_configure_rest(p, rest)) return 1;
    }
    // ---- batches ("super-terms"): Smolyak terms with the SAME schedule (tile groups, matrices, folded weight and
    // shift) are processed together as one work item.  In the internal layout the term index is simply one more (slowest)
    // dimension, so every pass runs over all tiles of all terms of the batch with full warps, one barrier per pass and one
    // descriptor, and the batch's map / V slices are contiguous.  Every term of a shape qualifies when the modes of equal
    // size share their 1-D basis (iso flavour: the kernel ignores the matrix offsets); otherwise only terms on modes with
    // identical matrices do.  The batch capacity adapts to the problem size so that small grids still fill the GPU.
    // second-generation kernel (sg4_fast2.cuh): single channel, iso tiles or pool-in-shared-memory tiles
    const bool v2_on = nb0 == 1 && envi("EVR_SG4_V2", 0) != 0 && !p->deterministic;   // experimental, off by default
    auto flavour_v2 = [&](int fl) { return v2_on && ((fl == 3 && iso && !iso_big) || (fl == 0 && pool_in_smem)); };
    auto term_is_iso = [&](int t) { const int fl = flavour_of(t); return iso && (fl == 3 || (fl == 2 && iso_big)); };
    const int64_t bcap_max = std::max(1, envi("EVR_SG4_BCAP", v2_on ? 3700 : 2800));                   // doubles per psi/acc buffer
    const int64_t target_items = (int64_t)p->sm_count * std::max(1, envi("EVR_SG4_ITEMS_PER_SM", 16));
    const int64_t bcap = std::max<int64_t>(1, std::min<int64_t>(bcap_max, (p->NQ_local * nb0 + target_items - 1) / target_items));
    struct Batch { std::vector<int> terms; int flavour, szclass; int64_t size; double cost, frac; };
    std::vector<Batch> batches;
    const int64_t th0 = envi("EVR_SG4_TH0", 3000), th1 = envi("EVR_SG4_TH1", 1400), th2 = envi("EVR_SG4_TH2", 600);
    {
        std::map<std::string, std::vector<int>> by_key;
        for (int t = 0; t < p->n_terms; ++t) {
            if (term_gen[t]) continue;
            const evr::FastTermDev &F = sched[t].F;
            std::string k;
            auto put = [&](const void *ptr, size_t n) { k.append(reinterpret_cast<const char *>(ptr), n); };
            const int fl = flavour_of(t);
            put(&fl, sizeof fl); put(&F.nq, sizeof F.nq); put(&F.ngroups, sizeof F.ngroups);
            put(&F.weight, sizeof F.weight); put(&F.vshift, sizeof F.vshift);
            for (int g = 0; g < F.ngroups; ++g) {
                evr::FastGroup Gd = F.g[g];
                if (term_is_iso(t)) Gd.mat1 = Gd.mat2 = Gd.mat3 = 0;
                put(&Gd, sizeof Gd);
            }
            by_key[k].push_back(t);
        }
        for (auto &kv : by_key) {
            const std::vector<int> &tl = kv.second;
            const evr::FastTermDev &F0 = sched[tl[0]].F;
            const int64_t tsz = (int64_t)F0.nq * nb0;
            int64_t T = (F0.ngroups == 0) ? 1 : std::max<int64_t>(1, bcap / tsz);
            if (envi("EVR_SG4_BATCH", 1) == 0) T = 1;
            const int64_t n = (int64_t)tl.size(), nbat = (n + T - 1) / T;
            for (int64_t b = 0; b < nbat; ++b) {                 // near-equal batches (sizes differ by at most one term)
                const int64_t lo = b * n / nbat, hi = (b + 1) * n / nbat;
                Batch Bt;
                Bt.terms.assign(tl.begin() + lo, tl.begin() + hi);
                Bt.flavour = flavour_of(tl[0]);
                Bt.size = tsz * (hi - lo);
                Bt.szclass = Bt.size > th0 ? 0 : (Bt.size > th1 ? 1 : (Bt.size > th2 ? 2 : 3));
                Bt.cost = 0.0;
                Bt.frac = ((double)b + 0.5) / (double)nbat;
                for (int64_t j = lo; j < hi; ++j) Bt.cost += p->h_cost[tl[j]];
                batches.push_back(std::move(Bt));
            }
        }
    }
    // Work order: size class (largest items first), then flavour, then R rounds; round r holds the r-th R-th of the
    // batches of EVERY shape, cost-descending inside the round.  Why rounds: the Smolyak weights alternate in sign with
    // |l| and reach +-C(D-1, k) (462 at D = 12).  Summed shape by shape (plain cost order), the running value of a low
    // packed element climbs to ~1e5 times its final value before the next shape cancels it, and the rounding of those
    // partial sums is what separates two summation orders: 3e-13 ... 7e-13 relative L2 against the CPU restatement of the reference at HH-12D L=7,
    // with 5e-13 run-to-run (FP64 atomics), against a 1e-12 gate.  With all shapes advancing together the running sums stay
    // near (fraction done) x (final value): 1e-13 with a fully interleaved order -- which costs 10 % (0.391 vs 0.356 ms: the
    // cost order is what balances the static round-robin) -- so the compromise is R = 8 rounds (EVR_SG4_ORDER_ROUNDS;
    // 1 = plain cost order).  Measurements: profiles/r2/parity_spread.txt.
    const int rounds = std::max(1, envi("EVR_SG4_ORDER_ROUNDS", 8));
    std::stable_sort(batches.begin(), batches.end(), [rounds](const Batch &a, const Batch &b) {
        if (a.szclass != b.szclass) return a.szclass < b.szclass;
        if (a.flavour != b.flavour) return a.flavour > b.flavour;
        const int ra = (int)(a.frac * rounds), rb = (int)(b.frac * rounds);
        if (ra != rb) return ra < rb;
        return a.cost > b.cost;
    });
    const int n_items = (int)batches.size();
    std::vector<evr::FastTermDev> fterms(n_items);
    // Per-item slices of the fast-path arrays are padded to 32 entries (aligned 128-bit copies, no tail guards):
    //   gmap : packed index of every entry of the batch in the INTERNAL layout (gather: vector loads, linear stores)
    //   fmap : the same indices sorted ascending + fpos = their positions in the batch buffer (scatter: neighbouring
    //          lanes hit neighbouring addresses, so the FP64 reductions of a warp share L2 sectors)
    //   fV   : V in the internal layout
    // padding entries: index -1 (skipped), position 0, V = 0.
    std::vector<int64_t> pad_off(n_items + 1, 0);
    for (int w = 0; w < n_items; ++w) {
        if (batches[w].size / nb0 > 65535) return 0;          // positions are 16-bit
        pad_off[w + 1] = pad_off[w] + ((batches[w].size / nb0 + 31) & ~(int64_t)31);
    }
    const int64_t NQ_pad = std::max<int64_t>(pad_off[n_items], 32);
    std::vector<int32_t> fmap((size_t)NQ_pad, -1), gmap((size_t)NQ_pad, -1);
    std::vector<uint16_t> fpos((size_t)NQ_pad, 0);
    std::vector<double> fV;
    if (Vgrid) fV.assign((size_t)nb0 * nb0 * NQ_pad, 0.0);
#pragma omp parallel for schedule(dynamic, 8)
    for (int w = 0; w < n_items; ++w) {
        const Batch &Bt = batches[w];
        evr::FastTermDev &F = fterms[w];
        F = sched[Bt.terms[0]].F;
        const int nq1 = F.nq, T = (int)Bt.terms.size();
        F.nq = nq1 * T;
        F.map_off = pad_off[w]; F.grid_off = pad_off[w];
        int32_t *mdst = fmap.data() + F.map_off;
        int32_t *gdst = gmap.data() + F.map_off;
        uint16_t *pdst = fpos.data() + F.map_off;
        std::vector<std::pair<int32_t, uint16_t>> ent((size_t)F.nq);
        for (int j = 0; j < T; ++j) {
            const int t = Bt.terms[j];
            const TermSched &S = sched[t];
            // permutation: internal index q' -> reference index q (odometer over internal modes)
            const int nm = (int)S.in_n.size();
            std::vector<int> idx(nm, 0);
            int64_t q = 0;
            const int32_t *msrc = p->h_map.data() + p->h_map_off[t];
            const int64_t ref_grid_off = p->h_grid_off[t];
            for (int qp = j * nq1; qp < (j + 1) * nq1; ++qp) {
                const int32_t m = msrc[q];
                ent[qp] = { m > 0 ? inv_perm[m - 1] : INT32_MAX, (uint16_t)qp };
                gdst[qp] = m > 0 ? inv_perm[m - 1] : -1;
                if (Vgrid)
                    for (int ij = 0; ij < nb0 * nb0; ++ij)
                        fV[(size_t)ij * NQ_pad + F.grid_off + qp] = Vgrid[(size_t)ij * p->NQ_total + p->grid_start + ref_grid_off + q];
                for (int m2 = 0; m2 < nm; ++m2) {
                    q += S.in_ref[m2];
                    if (++idx[m2] < S.in_n[m2]) break;
                    q -= (int64_t)S.in_ref[m2] * S.in_n[m2];
                    idx[m2] = 0;
                }
            }
        }
        if (flavour_v2(Bt.flavour)) continue;             // scatters through the gather map: no sorted map / positions
        std::sort(ent.begin(), ent.end());
        for (int j = 0; j < F.nq; ++j) {
            mdst[j] = (ent[j].first == INT32_MAX) ? -1 : ent[j].first;
            pdst[j] = ent[j].second;
        }
    }
    // launch configuration per (size class, flavour) + "next item" prefetch links
    if (evr::fast_set_attributes()) return 1;
    if (iso && evr::iso_set_attributes()) return 1;
    if (v2_on && (evr::v2_iso_set_attributes() || evr::v2_pool_set_attributes())) return 1;
    p->n_classes = 0;
    {
        const int class_gsize[4] = {envi("EVR_SG4_G0", 256), envi("EVR_SG4_G1", 96), envi("EVR_SG4_G2", 64), envi("EVR_SG4_G3", 32)};
        int w0 = 0;
        while (w0 < n_items) {
            const int szc = batches[w0].szclass, fl = batches[w0].flavour;
            int w1 = w0;
            int64_t cap = 1, nqmax = 1;
            while (w1 < n_items && batches[w1].szclass == szc && batches[w1].flavour == fl) {
                cap = std::max<int64_t>(cap, (int64_t)fterms[w1].nq * nb0);
                nqmax = std::max<int64_t>(nqmax, fterms[w1].nq);
                ++w1;
            }
            if (p->n_classes >= EVR_MAX_FCLASSES) return 0;
            {   // the psi buffer also stages the scatter map (6 bytes per entry of the slice padded to 32 entries)
                // and the bulk copies move whole padded slices (map: 4 B, V: 8 B per entry) into the acc buffer
                const int64_t nq32 = (nqmax + 31) & ~(int64_t)31;
                cap = std::max<int64_t>(cap, std::max<int64_t>((nq32 * 3 + 3) / 4, nq32));
            }
            cap = (cap + 3) & ~(int64_t)3;                      // the gather stores whole quads
            const bool rt = (fl == 1), tri = (fl == 2), iso_class = iso && (fl == 3 || (tri && iso_big));
            const int max_threads = tri ? EVR_FAST_MAX_THREADS_TRI : (flavour_v2(fl) ? envi("EVR_SG4_V2_THREADS", 768) : EVR_FAST_MAX_THREADS);
            const int gsize = std::min(class_gsize[szc], max_threads);
            const size_t per_group = flavour_v2(fl) ? ((size_t)cap * 20 + 2 * sizeof(evr::FastTermDev) + EVR_FAST_MBAR_BYTES)
                                                    : ((size_t)2 * cap * sizeof(double) + 2 * sizeof(evr::FastTermDev) + EVR_FAST_MBAR_BYTES);
            const size_t pool_bytes = (pool_in_smem && !iso_class) ? pool.size() * sizeof(double) : 0;
            // (experiment: a smaller budget leaves part of the unified 256 KB array to the L1 cache, where register spills live)
            const size_t budget = (size_t)std::min(227, std::max(16, envi("EVR_SG4_SMEM_KB", 227))) * 1024;
            if (pool_bytes + per_group > budget) return 0;
            int ngrp = (int)std::min<size_t>((budget - pool_bytes) / per_group, (size_t)(max_threads / gsize));
            if (gsize > 32) ngrp = std::min(ngrp, 15);          // named barriers 1..15
            if (envi("EVR_SG4_GRP_PCT", 100) < 100) ngrp = std::max(1, ngrp * envi("EVR_SG4_GRP_PCT", 100) / 100);   // experiment
            ngrp = std::max(1, std::min(ngrp, (w1 - w0 + p->sm_count - 1) / p->sm_count));
            const size_t smem = pool_bytes + per_group * ngrp;
            const int n = w1 - w0;
            const int ctas = std::max(1, std::min((n + ngrp - 1) / ngrp, p->sm_count));
            const int step = ctas * ngrp;
            for (int w = w0; w < w1; ++w) {
                evr::FastTermDev &F = fterms[w];
                if (w + step < w1) { F.next_map_off = fterms[w + step].map_off; F.next_grid_off = fterms[w + step].grid_off; F.next_nq = fterms[w + step].nq; }
                else { F.next_map_off = 0; F.next_grid_off = 0; F.next_nq = 0; }
                if (w + 2 * step < w1) { F.next2_map_off = fterms[w + 2 * step].map_off; F.next2_nq = fterms[w + 2 * step].nq; }
                else { F.next2_map_off = 0; F.next2_nq = 0; }
            }
            evr::FastClassDev &C = p->fclass[p->n_classes];
            C.term_begin = w0; C.n_terms = n; C.gsize = gsize; C.rt = rt ? 1 : 0; C.tri = tri ? 1 : 0; C.cap = (int)cap; C.cta_threads = ngrp * gsize;
            C.counter = nullptr;
            p->fclass_smem[p->n_classes] = smem; p->fclass_ctas[p->n_classes] = ctas;
            p->fclass_flavour[p->n_classes] = iso_class ? (tri ? 2 : 3) : (rt ? 1 : (tri ? 2 : 0));
            p->fclass_is_iso[p->n_classes] = iso_class;
            p->fclass_v2[p->n_classes] = flavour_v2(fl);
            ++p->n_classes;
            w0 = w1;
        }
    }
    p->n_fitems = n_items;
    if (!p->d_fcounters && cudaMalloc((void **)&p->d_fcounters, EVR_MAX_FCLASSES * sizeof(int)) != cudaSuccess)
        return fail("evr_sg4: cudaMalloc(work counters) failed");
    if (envi("EVR_SG4_DYNAMIC", 0) != 0)     // measured: 0.364 ms dynamic vs 0.355 static at L=7 (profiles/r2/sweep19_dynamic_items.txt): opt-in
        for (int c = 0; c < p->n_classes; ++c) if (!p->fclass_v2[c]) p->fclass[c].counter = p->d_fcounters + c;
    cudaFree(p->d_fterms); cudaFree(p->d_fmap); cudaFree(p->d_fmats); cudaFree(p->d_fV);
    p->d_fterms = nullptr; p->d_fmap = nullptr; p->d_fmats = nullptr; p->d_fV = nullptr;
    if (upload(&p->d_fterms, fterms.data(), fterms.size())) return 1;
    if (upload(&p->d_fmap, fmap.data(), fmap.size())) return 1;
    if (p->deterministic && build_entry_lists(p, fmap, NQ_pad)) return 1;   // entries = sorted scatter map positions
    cudaFree(p->d_gmap); p->d_gmap = nullptr;
    if (upload(&p->d_gmap, gmap.data(), gmap.size())) return 1;
    cudaFree(p->d_fpos); p->d_fpos = nullptr; cudaFree(p->d_perm); p->d_perm = nullptr;
    if (upload(&p->d_fpos, fpos.data(), fpos.size())) return 1;
    if (upload(&p->d_perm, perm.data(), perm.size())) return 1;
    cudaFree(p->d_inv_perm); p->d_inv_perm = nullptr;
    // zero-fill of the result inside the permute-in kernel, read-out gathering through the inverse permutation (coalesced
    // stores): 0.3480 vs 0.3507 ms at L = 7 (profiles/r2/sweep21_phase_isolation_and_latency.txt, sweep 31); EVR_SG4_PERMUTE=0: separate
    // zero-fill, scattering read-out
    if (block_order && !(getenv("EVR_SG4_PERMUTE") && atoi(getenv("EVR_SG4_PERMUTE")) == 0) && upload(&p->d_inv_perm, inv_perm.data(), inv_perm.size())) return 1;
    if (upload(&p->d_fmats, pool.data(), pool.size())) return 1;
    if (Vgrid && upload(&p->d_fV, fV.data(), fV.size())) return 1;
    evr::FastPlanDev &f = p->fpd;
    f.nb0 = nb0; f.n_terms = p->n_fitems; f.has_V = Vgrid ? 1 : 0; f.pool_len = (int)pool.size();
    p->fast_pool_in_smem = pool_in_smem;
    p->fast_iso = iso && std::any_of(term_iso.begin(), term_iso.end(), [](char c) { return c != 0; });
    p->iso_blocks.swap(iso_blocks);
    { static int next_id = 0; p->iso_id = ++next_id; }
    f.dbg = getenv("EVR_SG4_DEBUG") ? atoi(getenv("EVR_SG4_DEBUG")) : 0;
    if (f.dbg & (4 | 8 | 16 | 32)) {            // phase-isolation switches of the measurements in profiles/: results are WRONG
        static bool warned = false;
        if (!warned) fprintf(stderr, "evr_sg4: EVR_SG4_DEBUG=%d disables parts of the H|psi> action (timing experiments only)\n", f.dbg);
        warned = true;
    }
    // the L2 prefetch of the next term's slices stopped paying once the slices stream through LDGSTS (measured: equal
    // times); off unless EVR_SG4_PREFETCH=1
    if (!getenv("EVR_SG4_PREFETCH") || atoi(getenv("EVR_SG4_PREFETCH")) == 0) f.dbg |= 64;
    f.nb = p->nb; f.NQ_local = NQ_pad;      // channel stride of the padded V array
    f.terms = p->d_fterms; f.gmap = p->d_gmap; f.map = p->d_fmap; f.pos = p->d_fpos; f.mats = p->d_fmats; f.V = p->d_fV;
    p->fast = true;
    return 0;
}

static int build_entry_lists(evr_sg4_plan *p, const std::vector<int32_t> &map0, int64_t n_entries)
{
    std::vector<long long> off((size_t)p->nb + 1, 0);
    for (int64_t e = 0; e < n_entries; ++e) if (map0[e] >= 0) ++off[map0[e] + 1];
    for (int64_t i = 0; i < p->nb; ++i) off[i + 1] += off[i];
    std::vector<int32_t> ent((size_t)std::max<long long>(off[p->nb], 1));
    std::vector<long long> at(off.begin(), off.end() - 1);
    for (int64_t e = 0; e < n_entries; ++e) if (map0[e] >= 0) ent[at[map0[e]]++] = (int32_t)e;   // ascending entry order
    cudaFree(p->d_det_off); cudaFree(p->d_det_ent); p->d_det_off = nullptr; p->d_det_ent = nullptr;
    if (upload(&p->d_det_off, off.data(), off.size())) return 1;
    if (upload(&p->d_det_ent, ent.data(), ent.size())) return 1;
    p->stage_ld = std::max<int64_t>(n_entries, 1);
    return 0;
}

extern "C" int evr_sg4_plan_set_op(evr_sg4_plan *p, int type_Op, int nb_Term, const int32_t *term_mode,
                                   const uint8_t *grid_zero, const uint8_t *grid_cte,
                                   const double *Mat_cte, const double *const *grids)
{
    if (!p) return fail("evr_sg4_plan_set_op: null plan");
    if (!p->sub.empty()) return evr::multi_set_op(p, type_Op, nb_Term, term_mode, grid_zero, grid_cte, Mat_cte, grids);
    for (auto &g : p->graphs) cudaGraphExecDestroy(g.exec);      // captured launch sequences belong to the previous operator
    p->graphs.clear();
    if (type_Op != 0 && type_Op != 1)
        return fail("evr_sg4_plan_set_op: type_Op must be 0 or 1 (use evr_sg4_plan_set_op10 for type_Op=10)");
    if (nb_Term < 1 || !grid_zero || !grid_cte) return fail("evr_sg4_plan_set_op: bad term list");
    if (type_Op == 1 && !term_mode) return fail("evr_sg4_plan_set_op: term_mode required for type_Op=1");
    CUDA_TRY(cudaSetDevice(p->device));
    const int nb0 = p->nb0;
    const int nterm = (type_Op == 0) ? 1 : nb_Term;
    std::vector<evr::OpTermDev> ops;
    std::vector<int> var_terms;
    int64_t deriv_flops = 0;
    for (int it = 0; it < nterm; ++it) {
        if (grid_zero[it]) continue;                              // sub_OpPsi_SG4.f90:1511
        evr::OpTermDev O{};
        int m1 = (type_Op == 0) ? 0 : term_mode[2 * it], m2 = (type_Op == 0) ? 0 : term_mode[2 * it + 1];
        if (m1 < 0) m1 = 0;                                         // WHERE (tab_der < 0) tab_der = 0
        if (m2 < 0) m2 = 0;
        if (m1 > p->D || m2 > p->D) return fail("evr_sg4_plan_set_op: term_mode out of range");
        O.m1 = m1 - 1; O.m2 = m2 - 1;
        if (O.m1 >= 0 && O.m2 >= 0 && O.m1 > O.m2) std::swap(O.m1, O.m2);
        if (grid_cte[it]) {
            if (!Mat_cte) return fail("evr_sg4_plan_set_op: Mat_cte required for grid_cte terms");
            O.grid_slot = -1;
            for (int i = 0; i < nb0; ++i)
                for (int j = 0; j < nb0; ++j) O.cte[i + nb0 * j] = Mat_cte[(size_t)it * nb0 * nb0 + i + nb0 * j];
        } else {
            if (!grids || !grids[it]) return fail("evr_sg4_plan_set_op: missing grid for a non-constant term");
            O.grid_slot = (int)var_terms.size();
            var_terms.push_back(it);
        }
        ops.push_back(O);
    }
    // upload the variable grids of this plan's term range: [slot][i + nb0*j][NQ_local]
    if (p->d_grids) { cudaFree(p->d_grids); p->d_grids = nullptr; }
    if (p->d_opterms) { cudaFree(p->d_opterms); p->d_opterms = nullptr; }
    const size_t blk = (size_t)std::max<int64_t>(p->NQ_local, 1);
    const size_t ng = var_terms.size() * nb0 * nb0;
    CUDA_TRY(cudaMalloc((void **)&p->d_grids, std::max<size_t>(ng * blk, 1) * sizeof(double)));
    for (size_t s = 0; s < var_terms.size(); ++s)
        for (int ij = 0; ij < nb0 * nb0; ++ij) {
            const double *src = grids[var_terms[s]] + (size_t)ij * p->NQ_total + p->grid_start;
            CUDA_TRY(cudaMemcpy(p->d_grids + (s * nb0 * nb0 + ij) * blk, src, (size_t)p->NQ_local * sizeof(double), cudaMemcpyHostToDevice));
        }
    {   // on-the-fly terms first, then the mixed-derivative sweeps grouped by their first mode (sg4_kernels.cuh)
        std::vector<char> has_sweep(p->D, 0);
        for (const auto &O : ops) if (O.m1 >= 0 && O.m2 >= 0 && O.m1 != O.m2) has_sweep[O.m1] = 1;
        auto sweep_of = [&](const evr::OpTermDev &O) {
            if (O.m1 >= 0 && O.m2 >= 0) return (O.m1 != O.m2) ? O.m1 : -1;
            const int k = (O.m1 >= 0) ? O.m1 : O.m2;                 // first derivative alone (or none: k = -1)
            return (k >= 0 && has_sweep[k]) ? k : -1;
        };
        std::vector<evr::OpTermDev> sorted;
        for (const auto &O : ops) if (sweep_of(O) < 0) sorted.push_back(O);
        p->pd.n_plain = (int)sorted.size();
        p->pd.n_sweeps = 0;
        for (int a = 0; a < p->D; ++a) {
            if (!has_sweep[a]) continue;
            p->pd.sweep_mode[p->pd.n_sweeps] = a;
            p->pd.sweep_begin[p->pd.n_sweeps] = (int)sorted.size();
            for (auto O : ops) if (sweep_of(O) == a) {
                if (O.m1 != a) std::swap(O.m1, O.m2);               // sweep terms: m1 = cached mode, m2 = other mode or -1
                sorted.push_back(O);
            }
            ++p->pd.n_sweeps;
        }
        p->pd.sweep_begin[p->pd.n_sweeps] = (int)sorted.size();
        ops.swap(sorted);
        if (gen_configure(p, p->pd.n_sweeps > 0, (int)ops.size())) return 1;
    }
    if (upload(&p->d_opterms, ops.data(), ops.size())) return 1;
    p->op10 = false;
    p->type_Op = type_Op; p->n_opterms = (int)ops.size(); p->n_var = (int)var_terms.size();
    p->pd.type_Op = type_Op; p->pd.n_opterms = p->n_opterms; p->pd.n_var = p->n_var;
    p->pd.opterms = p->d_opterms; p->pd.grids = p->d_grids;
    // algorithmic flops of the operator stage
    for (int t = 0; t < p->n_terms; ++t) {
        const int iG = p->iG_begin + t;
        const int64_t nq = p->h_tab_nq[iG];
        auto nqm = [&](int m) { return (int64_t)p->h_nq_of[m * (p->LG + 1) + p->h_tab_l[(size_t)iG * p->D + m]]; };
        for (const auto &O : ops) {
            if (O.m1 >= 0 && O.m2 >= 0 && O.m1 != O.m2) deriv_flops += 2 * nq * (nqm(O.m1) + nqm(O.m2));
            else if (O.m1 >= 0 || O.m2 >= 0) deriv_flops += 2 * nq * nqm(O.m1 >= 0 ? O.m1 : O.m2);
            deriv_flops += 2 * nq * nb0;                           // pointwise multiply-add
        }
    }
    p->flops_npsi1 += deriv_flops * nb0;
    if (build_fast_path(p, nb_Term, term_mode, grid_zero, grid_cte, Mat_cte, grids)) return 1;
    if (p->deterministic && !p->fast) {                      // generic kernel: entries = positions of the mapping slice
        std::vector<int32_t> map0(p->h_map.size());
        for (size_t e = 0; e < map0.size(); ++e) map0[e] = p->h_map[e] - 1;
        if (build_entry_lists(p, map0, (int64_t)map0.size())) return 1;
    }
    p->op_set = true;
    return 0;
}


extern "C" int evr_sg4_plan_set_op10(evr_sg4_plan *p, int n_act, const int32_t *act_mode,
                                     const double *V, const double *GG, const double *Jac, const double *sq)
{
    if (!p) return fail("evr_sg4_plan_set_op10: null plan");
    if (!p->sub.empty()) return evr::multi_set_op10(p, n_act, act_mode, V, GG, Jac, sq);
    for (auto &g : p->graphs) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
    if (n_act < 1 || n_act > EVR_MAXD || !act_mode || !GG || !Jac || !sq) return fail("evr_sg4_plan_set_op10: bad arguments");
    CUDA_TRY(cudaSetDevice(p->device));
    const int nb0 = p->nb0;
    evr::Op10Dev &O = p->o10;
    O.n_act = n_act;
    for (int j = 0; j < n_act; ++j) {
        if (act_mode[j] < 1 || act_mode[j] > p->D) return fail("evr_sg4_plan_set_op10: act_mode out of range");
        O.act_mode[j] = act_mode[j] - 1;
    }
    int nqmax = 1;
    for (int t = 0; t < p->n_terms; ++t) nqmax = std::max(nqmax, (int)p->h_tab_nq[p->iG_begin + t]);
    O.nqmax = nqmax;
    const int nT = p->D * (p->LG + 1);
    p->smem10 = ((size_t)2 * p->cap + (size_t)(n_act + 1) * nqmax) * sizeof(double) + (size_t)(4 * nT + 7 * p->D) * sizeof(int);
    if (p->smem10 > 227 * 1024) return fail("evr_sg4_plan_set_op10: shared-memory budget exceeded for this n_act / term size");
    const size_t blk = (size_t)std::max<int64_t>(p->NQ_local, 1);
    auto up_slice = [&](double **d, const double *h, int ncomp) -> int {
        if (*d) { cudaFree(*d); *d = nullptr; }
        CUDA_TRY(cudaMalloc((void **)d, blk * ncomp * sizeof(double)));
        for (int c = 0; c < ncomp; ++c)
            CUDA_TRY(cudaMemcpy(*d + (size_t)c * blk, h + (size_t)c * p->NQ_total + p->grid_start, (size_t)p->NQ_local * sizeof(double), cudaMemcpyHostToDevice));
        return 0;
    };
    // the metric tensor is symmetric: when GG(q,j,i) == GG(q,i,j) holds exactly on this plan's grid range only the upper
    // triangle is kept on the device (n(n+1)/2 + 2 instead of n^2 + 2 doubles per point)
    bool sym = !(getenv("EVR_SG4_GG_FULL") && atoi(getenv("EVR_SG4_GG_FULL")) != 0);
    for (int i = 0; i < n_act && sym; ++i)
        for (int j = 0; j < i && sym; ++j) {
            const double *a = GG + (size_t)(j + n_act * i) * p->NQ_total + p->grid_start, *b = GG + (size_t)(i + n_act * j) * p->NQ_total + p->grid_start;
            if (std::memcmp(a, b, (size_t)p->NQ_local * sizeof(double)) != 0) sym = false;
        }
    O.sym = sym ? 1 : 0;
    if (sym) {
        if (p->d_GG) { cudaFree(p->d_GG); p->d_GG = nullptr; }
        CUDA_TRY(cudaMalloc((void **)&p->d_GG, blk * (size_t)(n_act * (n_act + 1) / 2) * sizeof(double)));
        for (int i = 0; i < n_act; ++i)
            for (int j = 0; j <= i; ++j)
                CUDA_TRY(cudaMemcpy(p->d_GG + (size_t)(j + i * (i + 1) / 2) * blk, GG + (size_t)(j + n_act * i) * p->NQ_total + p->grid_start,
                                    (size_t)p->NQ_local * sizeof(double), cudaMemcpyHostToDevice));
    } else if (up_slice(&p->d_GG, GG, n_act * n_act)) return 1;
    if (up_slice(&p->d_Jac, Jac, 1)) return 1;
    if (up_slice(&p->d_sq, sq, 1)) return 1;
    O.has_V = V ? 1 : 0;
    if (V) { if (up_slice(&p->d_grids, V, nb0 * nb0)) return 1; }
    O.V = p->d_grids; O.GG = p->d_GG; O.Jac = p->d_Jac; O.sq = p->d_sq;
    {   // the attribute belongs to the function, not to the plan: never lower it under another live plan
        static size_t attr10_max[64] = {0};
        size_t &amax = attr10_max[p->device & 63];
        amax = std::max(amax, p->smem10);
        CUDA_TRY(cudaFuncSetAttribute(evr::sg4_term_kernel_type10, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)amax));
    }
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, evr::sg4_term_kernel_type10, 256, p->smem10));
    if (occ < 1) return fail("evr_sg4_plan_set_op10: kernel cannot be resident");
    p->ctas10_max = p->sm_count * occ;
    p->type_Op = 10; p->pd.type_Op = 10; p->n_var = (V ? 1 : 0);
    p->op10 = true; p->fast = false; p->op_set = true;
    return 0;
}

// sub_scaledOpPsi (sub_OpPsi.f90:2823-2866) on the device: y <- (y - E0 x) / Esc, optionally while un-permuting the
// block-ordered internal result (src[v*nb + i] -> dst[v*nb + perm[i]])
namespace evr {
__global__ void sg4_scale_kernel(const long long n, const double E0, const double Esc,
                                 const double *__restrict__ x, double *__restrict__ y)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        y[i] = (y[i] - E0 * __ldg(x + i)) / Esc;
}
__global__ void sg4_permute_out_scaled(const int32_t *__restrict__ perm, const long long nb, const int nvecs,
                                       const double E0, const double Esc, const double *__restrict__ x_user,
                                       const double *__restrict__ src, double *__restrict__ dst)
{
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < nb; i += (long long)gridDim.x * blockDim.x) {
        const int r = __ldg(perm + i);
        for (int v = 0; v < nvecs; ++v) dst[v * nb + r] = (src[v * nb + i] - E0 * __ldg(x_user + v * nb + r)) / Esc;
    }
}
}

struct ScaleArgs { bool on; double E0, Esc; };

// the launches of one set of generic-kernel size classes (all terms, or the remainder of a fast-path plan), class tails
// overlapping on the side streams
static int launch_generic_set(evr_sg4_plan *p, int ncls, const evr::GenClassDev *cls, const int *threads, const int *occ,
                              const size_t *smem, const bool *big, int npsi, const double *d_psi, double *d_Hpsi, cudaStream_t st)
{
    const bool multi = ncls > 1 && p->ev_fork != nullptr;
    if (multi) CUDA_TRY(cudaEventRecord(p->ev_fork, st));
    for (int c = 0; c < ncls; ++c) {
        cudaStream_t sc = (multi && c > 0) ? p->side[c] : st;
        if (multi && c > 0) CUDA_TRY(cudaStreamWaitEvent(sc, p->ev_fork, 0));
        const long long items = (long long)cls[c].n_terms * npsi;
        const long long cmax = occ[c] < 0 ? -occ[c] : (long long)p->sm_count * occ[c];
        const int ctas = (int)std::max<long long>(1, std::min<long long>(items, cmax));
        if (big[c])
            evr::sg4_term_kernel_generic<true><<<ctas, threads[c], smem[c], sc>>>(p->pd, cls[c], npsi, d_psi, d_Hpsi);
        else
            evr::sg4_term_kernel_generic<false><<<ctas, threads[c], smem[c], sc>>>(p->pd, cls[c], npsi, d_psi, d_Hpsi);
        p->launches += 1;
        if (multi && c > 0) {
            CUDA_TRY(cudaEventRecord(p->ev_join[c], sc));
            CUDA_TRY(cudaStreamWaitEvent(st, p->ev_join[c], 0));
        }
    }
    return 0;
}

static int launch_direct(evr_sg4_plan *p, int npsi, const double *d_psi_user, double *d_Hpsi_user, cudaStream_t st,
                         const ScaleArgs sc)
{
    const size_t bytes = (size_t)npsi * p->nb * p->nb0 * sizeof(double);
    const double *d_psi = d_psi_user;
    double *d_Hpsi = d_Hpsi_user;
    const bool use_int = p->fast && p->fast_block_order && p->n_terms > 0;
    if (use_int) {
        // fast path works on the packed vectors in the internal (block) order
        const int64_t nvecs = (int64_t)npsi * p->nb0;
        if (nvecs * p->nb > p->int_cap) {
            cudaFree(p->d_psi_int); cudaFree(p->d_Hpsi_int); p->d_psi_int = p->d_Hpsi_int = nullptr; p->int_cap = 0;
            CUDA_TRY(cudaMalloc((void **)&p->d_psi_int, bytes));
            CUDA_TRY(cudaMalloc((void **)&p->d_Hpsi_int, bytes));
            p->int_cap = nvecs * p->nb;
        }
        if (p->d_inv_perm && !p->deterministic) evr::fast_permute_x(true, p->d_perm, p->nb, (int)nvecs, d_psi_user, p->d_psi_int, p->d_Hpsi_int, st);
        else evr::fast_permute(true, p->d_perm, p->nb, (int)nvecs, d_psi_user, p->d_psi_int, st);
        p->launches += 1;
        d_psi = p->d_psi_int; d_Hpsi = p->d_Hpsi_int;
    }
    const bool det = p->deterministic && !p->op10 && p->d_det_off != nullptr && p->n_terms > 0;
    if (det) {
        const int64_t vecs = (int64_t)npsi * p->nb0;
        if (vecs > p->stage_vecs) {
            CUDA_TRY(cudaStreamSynchronize(st));
            cudaFree(p->d_stage); p->d_stage = nullptr; p->stage_vecs = 0;
            CUDA_TRY(cudaMalloc((void **)&p->d_stage, (size_t)vecs * p->stage_ld * sizeof(double)));
            p->stage_vecs = vecs;
        }
    }
    p->fpd.stage = det ? p->d_stage : nullptr; p->fpd.stage_ld = p->stage_ld;
    p->pd.stage = det ? p->d_stage : nullptr;  p->pd.stage_ld = p->stage_ld;
    if (!(use_int && p->d_inv_perm && !p->deterministic))
        CUDA_TRY(cudaMemsetAsync(d_Hpsi, 0, bytes, st));             // reference zeroes OpPsi (:765)
    if (p->n_terms > 0) {
        if (p->fast) {
            if ((long long)p->n_fitems * npsi > INT_MAX) return fail("evr_sg4_apply: n_terms * npsi exceeds 2^31 work items");
            bool any_v1_iso = false, any_v2_iso = false;
            for (int c = 0; c < p->n_classes; ++c) if (p->fclass_is_iso[c]) (p->fclass_v2[c] ? any_v2_iso : any_v1_iso) = true;
            if (any_v1_iso && evr::iso_bind(p->device, p->iso_id, p->iso_blocks.data(), st)) return 1;
            if (any_v2_iso && evr::v2_iso_bind(p->device, p->iso_id, p->iso_blocks.data(), st)) return 1;
            if (p->d_fcounters) CUDA_TRY(cudaMemsetAsync(p->d_fcounters, 0, EVR_MAX_FCLASSES * sizeof(int), st));
            const bool multi = p->n_classes > 1 && p->ev_fork != nullptr;
            if (multi) CUDA_TRY(cudaEventRecord(p->ev_fork, st));
            for (int c = 0; c < p->n_classes; ++c) {
                cudaStream_t st_main = st;
                cudaStream_t st = (multi && c > 0) ? p->side[c] : st_main;
                if (multi && c > 0) CUDA_TRY(cudaStreamWaitEvent(st, p->ev_fork, 0));
                const bool ms = p->fast_pool_in_smem, rt = p->fclass[c].rt != 0, tri = p->fclass[c].tri != 0;
                const int nctas = p->fclass_ctas[c], nthr = p->fclass[c].cta_threads;
                const size_t sm = p->fclass_smem[c];
                if (p->fclass_v2[c]) {
                    if (p->fclass_is_iso[c]) { if (evr::v2_iso_launch(nctas, nthr, sm, st, p->fpd, p->fclass[c], npsi, d_psi, d_Hpsi)) return 1; }
                    else if (evr::v2_pool_launch(nctas, nthr, sm, st, p->fpd, p->fclass[c], npsi, d_psi, d_Hpsi)) return 1;
                }
                else if (p->fclass_is_iso[c]) { if (evr::iso_launch(tri, nctas, nthr, sm, st, p->fpd, p->fclass[c], npsi, d_psi, d_Hpsi)) return 1; }
                else if (evr::fast_launch(ms ? 1 : 0, rt, tri, nctas, nthr, sm, st, p->fpd, p->fclass[c], npsi, d_psi, d_Hpsi)) return 1;
                p->launches += 1;
                if (multi && c > 0) {
                    CUDA_TRY(cudaEventRecord(p->ev_join[c], st));
                    CUDA_TRY(cudaStreamWaitEvent(st_main, p->ev_join[c], 0));
                }
            }
        } else if (p->op10) {
            const long long items = (long long)p->n_terms * npsi;
            const int ctas = (int)std::max<long long>(1, std::min<long long>(items, (long long)p->ctas10_max));
            evr::sg4_term_kernel_type10<<<ctas, 256, p->smem10, st>>>(p->pd, p->o10, npsi, d_psi, d_Hpsi);
            p->launches += 1;
        } else {
            if (launch_generic_set(p, p->n_gclasses, p->gclass, p->gclass_threads, p->gclass_occ, p->gclass_smem, p->gclass_big, npsi, d_psi, d_Hpsi, st)) return 1;
        }
        CUDA_TRY(cudaGetLastError());
        if (det) {      // fixed-order sums of the staged entries (all class kernels have joined the stream)
            const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((p->nb + 7) / 8, 148 * 16));   // one warp per packed element
            evr::sg4_collect_kernel<<<blocks, 256, 0, st>>>(p->nb, (int)(npsi * p->nb0), p->d_det_off, p->d_det_ent, p->d_stage, p->stage_ld, d_Hpsi);
            p->launches += 1;
            CUDA_TRY(cudaGetLastError());
        }
    }
    // a fast-path plan may leave a few terms to the generic kernel (gen_configure_rest): they work on the caller's vectors
    // (reference order), after the fast part has been brought back to that order and before the scaling
    const bool has_rest = p->fast && p->n_rclasses > 0 && p->n_terms > 0;
    if (use_int) {
        const int64_t nvecs = (int64_t)npsi * p->nb0;
        if (sc.on && !has_rest) {
            const int thr = 256;
            const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((p->nb + thr - 1) / thr, 148 * 16));
            evr::sg4_permute_out_scaled<<<blocks, thr, 0, st>>>(p->d_perm, p->nb, (int)nvecs, sc.E0, sc.Esc, d_psi_user, p->d_Hpsi_int, d_Hpsi_user);
        } else if (p->d_inv_perm) evr::fast_permute_x(false, p->d_inv_perm, p->nb, (int)nvecs, p->d_Hpsi_int, d_Hpsi_user, nullptr, st);
        else evr::fast_permute(false, p->d_perm, p->nb, (int)nvecs, p->d_Hpsi_int, d_Hpsi_user, st);
        p->launches += 1;
        CUDA_TRY(cudaGetLastError());
    }
    if (has_rest) {
        if (launch_generic_set(p, p->n_rclasses, p->rclass, p->rclass_threads, p->rclass_occ, p->rclass_smem, p->rclass_big, npsi, d_psi_user, d_Hpsi_user, st)) return 1;
        CUDA_TRY(cudaGetLastError());
    }
    if (sc.on && (!use_int || has_rest)) {
        const long long n = (long long)npsi * p->nb * p->nb0;
        const int thr = 256;
        const int blocks = (int)std::max<long long>(1, std::min<long long>((n + thr - 1) / thr, 148 * 16));
        evr::sg4_scale_kernel<<<blocks, thr, 0, st>>>(n, sc.E0, sc.Esc, d_psi_user, d_Hpsi_user);
        p->launches += 1;
        CUDA_TRY(cudaGetLastError());
    }
    return 0;
}

// ---- CUDA graph of one H|psi> ----------------------------------------------------------------------------------------
// The per-call launch sequence (permute-in, memset, one kernel per (size class, flavour) on forked streams, joins,
// permute-out / scaling) is captured once per (npsi, psi, Hpsi, scaling) and replayed with a single cudaGraphLaunch: the
// small reference shapes (HCN_UT, pyrazine, the 6-D KAT) are bound by that sequence, not by the kernels, and at N > 1 it is
// the fixed part of every rank's step.  Measured (profiles/r2/sweep12_cuda_graph.txt): 1-5 % on the small shapes, nothing at
// L = 7, and a Davidson block on the generic kernel (HCN shape, 27 vectors) loses the overlap of its class launches
// (344 vs 216 us) -- so the graph is opt-in: EVR_SG4_GRAPH=1.
static int bind_iso_arrays(evr_sg4_plan *p, cudaStream_t st)
{
    if (!p->fast) return 0;
    bool any_v1_iso = false, any_v2_iso = false;
    for (int c = 0; c < p->n_classes; ++c) if (p->fclass_is_iso[c]) (p->fclass_v2[c] ? any_v2_iso : any_v1_iso) = true;
    if (any_v1_iso && evr::iso_bind(p->device, p->iso_id, p->iso_blocks.data(), st)) return 1;
    if (any_v2_iso && evr::v2_iso_bind(p->device, p->iso_id, p->iso_blocks.data(), st)) return 1;
    return 0;
}

static int launch(evr_sg4_plan *p, int npsi, const double *d_psi_user, double *d_Hpsi_user, cudaStream_t st,
                  const ScaleArgs sc = ScaleArgs{false, 0.0, 1.0})
{
    static const bool graphs_on = getenv("EVR_SG4_GRAPH") && atoi(getenv("EVR_SG4_GRAPH")) != 0;
    // (the deterministic mode allocates its staging vector inside the launch sequence: not captured)
    if (!graphs_on || p->deterministic || p->n_terms == 0 || p->stream == nullptr) return launch_direct(p, npsi, d_psi_user, d_Hpsi_user, st, sc);
    for (auto &g : p->graphs)
        if (g.npsi == npsi && g.psi == d_psi_user && g.Hpsi == d_Hpsi_user && g.scaled == sc.on && g.E0 == sc.E0 && g.Esc == sc.Esc) {
            if (bind_iso_arrays(p, st)) return 1;               // another plan may have re-bound the constant arrays
            CUDA_TRY(cudaGraphLaunch(g.exec, st));
            p->launches += g.kernels;
            g.stamp = ++p->graph_clock;
            return 0;
        }
    // first call with these arguments: allocations and bindings happen outside the capture
    if (p->fast && p->fast_block_order) {
        const int64_t nvecs = (int64_t)npsi * p->nb0;
        if (nvecs * p->nb > p->int_cap) {
            const size_t bytes = (size_t)npsi * p->nb * p->nb0 * sizeof(double);
            for (auto &g : p->graphs) cudaGraphExecDestroy(g.exec);     // they reference the old internal buffers
            p->graphs.clear();
            CUDA_TRY(cudaStreamSynchronize(st));
            cudaFree(p->d_psi_int); cudaFree(p->d_Hpsi_int); p->d_psi_int = p->d_Hpsi_int = nullptr; p->int_cap = 0;
            CUDA_TRY(cudaMalloc((void **)&p->d_psi_int, bytes));
            CUDA_TRY(cudaMalloc((void **)&p->d_Hpsi_int, bytes));
            p->int_cap = nvecs * p->nb;
        }
    }
    if (bind_iso_arrays(p, st)) return 1;
    // capture on the plan's own stream (the caller's may be the legacy default stream, which cannot capture)
    cudaStream_t cs = p->stream;
    const int64_t l0 = p->launches;
    if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        cudaGetLastError();
        return launch_direct(p, npsi, d_psi_user, d_Hpsi_user, st, sc);
    }
    const int rc = launch_direct(p, npsi, d_psi_user, d_Hpsi_user, cs, sc);
    cudaGraph_t graph = nullptr;
    const cudaError_t ec = cudaStreamEndCapture(cs, &graph);
    const int kernels = (int)(p->launches - l0);
    p->launches = l0;
    if (rc || ec != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        if (rc) return 1;
        return launch_direct(p, npsi, d_psi_user, d_Hpsi_user, st, sc);
    }
    evr_sg4_plan::GraphEntry g;
    const cudaError_t ei = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ei != cudaSuccess) { cudaGetLastError(); return launch_direct(p, npsi, d_psi_user, d_Hpsi_user, st, sc); }
    g.npsi = npsi; g.psi = d_psi_user; g.Hpsi = d_Hpsi_user; g.scaled = sc.on; g.E0 = sc.E0; g.Esc = sc.Esc; g.kernels = kernels;
    g.stamp = ++p->graph_clock;
    if (p->graphs.size() >= 8) {                                 // keep the eight most recently used argument sets
        size_t old = 0;
        for (size_t i = 1; i < p->graphs.size(); ++i) if (p->graphs[i].stamp < p->graphs[old].stamp) old = i;
        cudaGraphExecDestroy(p->graphs[old].exec);
        p->graphs.erase(p->graphs.begin() + old);
    }
    p->graphs.push_back(g);
    CUDA_TRY(cudaGraphLaunch(g.exec, st));
    p->launches += kernels;
    return 0;
}

int evr::plan_launch(evr_sg4_plan *p, int npsi, const double *d_psi, double *d_Hpsi, cudaStream_t st)
{
    return launch(p, npsi, d_psi, d_Hpsi, st);
}
int evr::scale_launch(long long n, double E0, double Esc, const double *x, double *y, cudaStream_t st)
{
    const int thr = 256;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((n + thr - 1) / thr, 148 * 16));
    evr::sg4_scale_kernel<<<blocks, thr, 0, st>>>(n, E0, Esc, x, y);
    return cudaGetLastError() == cudaSuccess ? 0 : fail("evr_sg4: scale kernel launch failed");
}
int evr::plan_ensure_staging(evr_sg4_plan *p, int64_t n)
{
    if (n <= p->stage_cap) return 0;
    if (p->d_psi) cudaFree(p->d_psi);
    if (p->d_Hpsi) cudaFree(p->d_Hpsi);
    p->d_psi = p->d_Hpsi = nullptr; p->stage_cap = 0;
    CUDA_TRY(cudaMalloc((void **)&p->d_psi, (size_t)n * sizeof(double)));
    CUDA_TRY(cudaMalloc((void **)&p->d_Hpsi, (size_t)n * sizeof(double)));
    p->stage_cap = n;
    return 0;
}

// [psi, psi+n) and [Hpsi, Hpsi+n) must not overlap: the result vector is zeroed before psi is gathered
static bool ranges_overlap(const double *a, const double *b, int64_t n)
{
    const uintptr_t x = reinterpret_cast<uintptr_t>(a), y = reinterpret_cast<uintptr_t>(b), len = (uintptr_t)n * sizeof(double);
    return x < y + len && y < x + len;
}

extern "C" int evr_sg4_apply_device(evr_sg4_plan *p, int npsi, const double *d_psi, double *d_Hpsi, void *cuda_stream)
{
    if (!p) return fail("evr_sg4_apply_device: null plan");
    if (!p->sub.empty()) {
        if (!d_psi || !d_Hpsi || npsi < 1) return fail("evr_sg4_apply_device: bad arguments");
        return evr::multi_apply_device(p, npsi, d_psi, d_Hpsi, (cudaStream_t)cuda_stream, false, 0.0, 1.0);
    }
    if (!p->op_set) return fail("evr_sg4_apply_device: operator not set (call evr_sg4_plan_set_op)");
    if (npsi < 1) return fail("evr_sg4_apply: size(Psi) = 0");     // reference: STOP (:738-743)
    if (!d_psi || !d_Hpsi) return fail("evr_sg4_apply_device: null buffer");
    if (ranges_overlap(d_psi, d_Hpsi, (int64_t)npsi * p->nb * p->nb0))
        return fail("evr_sg4_apply_device: psi and Hpsi overlap (the action is not in-place)");
    CUDA_TRY(cudaSetDevice(p->device));
    return launch(p, npsi, d_psi, d_Hpsi, (cudaStream_t)cuda_stream);
}

extern "C" int evr_sg4_apply_device_scaled(evr_sg4_plan *p, int npsi, const double *d_psi, double *d_Hpsi,
                                           double E0, double Esc, void *cuda_stream)
{
    if (!p) return fail("evr_sg4_apply_device_scaled: null plan");
    if (!p->sub.empty()) {
        if (!d_psi || !d_Hpsi || npsi < 1 || Esc == 0.0) return fail("evr_sg4_apply_device_scaled: bad arguments");
        return evr::multi_apply_device(p, npsi, d_psi, d_Hpsi, (cudaStream_t)cuda_stream, true, E0, Esc);
    }
    if (!p->op_set) return fail("evr_sg4_apply_device_scaled: operator not set (call evr_sg4_plan_set_op)");
    if (npsi < 1) return fail("evr_sg4_apply: size(Psi) = 0");
    if (!d_psi || !d_Hpsi) return fail("evr_sg4_apply_device_scaled: null buffer");
    if (ranges_overlap(d_psi, d_Hpsi, (int64_t)npsi * p->nb * p->nb0))
        return fail("evr_sg4_apply_device_scaled: psi and Hpsi overlap (the action is not in-place)");
    if (Esc == 0.0) return fail("evr_sg4_apply_device_scaled: Esc = 0");
    CUDA_TRY(cudaSetDevice(p->device));
    return launch(p, npsi, d_psi, d_Hpsi, (cudaStream_t)cuda_stream, ScaleArgs{true, E0, Esc});
}

extern "C" int evr_sg4_apply(evr_sg4_plan *p, int npsi, const double *psi, double *Hpsi)
{
    if (!p) return fail("evr_sg4_apply: null plan");
    if (!p->sub.empty()) {
        if (!psi || !Hpsi) return fail("evr_sg4_apply: null buffer");
        if (npsi < 1) return fail("evr_sg4_apply: size(Psi) = 0");
        return evr::multi_apply_host(p, npsi, psi, Hpsi);
    }
    if (!p->op_set) return fail("evr_sg4_apply: operator not set (call evr_sg4_plan_set_op)");
    if (npsi < 1) return fail("evr_sg4_apply: size(Psi) = 0");
    if (!psi || !Hpsi) return fail("evr_sg4_apply: null buffer");
    CUDA_TRY(cudaSetDevice(p->device));
    const int64_t n = (int64_t)npsi * p->nb * p->nb0;
    if (ranges_overlap(psi, Hpsi, n)) return fail("evr_sg4_apply: psi and Hpsi overlap (the action is not in-place)");
    if (evr::plan_ensure_staging(p, n)) return 1;
    // Blocks of long vectors (sub_TabOpPsi with a Davidson block): vector v+1 travels to the device and vector v-1 back to
    // the host while vector v is in the kernels -- three streams, full-duplex PCIe.  Page-locked caller buffers
    // (evr_sg4_host_register) are needed for the copies to overlap; pageable ones still give the right result.
    const int64_t nv = p->nb * p->nb0;
    static const bool no_pipe = getenv("EVR_SG4_PIPELINE") && atoi(getenv("EVR_SG4_PIPELINE")) == 0;
    if (npsi >= 2 && nv * (int64_t)sizeof(double) >= ((int64_t)1 << 20) && !no_pipe) {
        if (!p->s_in) {
            if (cudaStreamCreateWithFlags(&p->s_in, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithFlags(&p->s_out, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_pin, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_pout, cudaEventDisableTiming) != cudaSuccess)
                return fail("evr_sg4_apply: stream/event creation failed");
        }
        for (int v = 0; v < npsi; ++v) {
            CUDA_TRY(cudaMemcpyAsync(p->d_psi + v * nv, psi + v * nv, (size_t)nv * sizeof(double), cudaMemcpyHostToDevice, p->s_in));
            CUDA_TRY(cudaEventRecord(p->ev_pin, p->s_in));
            CUDA_TRY(cudaStreamWaitEvent(p->stream, p->ev_pin, 0));
            if (launch(p, 1, p->d_psi + v * nv, p->d_Hpsi + v * nv, p->stream)) return 1;
            CUDA_TRY(cudaEventRecord(p->ev_pout, p->stream));
            CUDA_TRY(cudaStreamWaitEvent(p->s_out, p->ev_pout, 0));
            CUDA_TRY(cudaMemcpyAsync(Hpsi + v * nv, p->d_Hpsi + v * nv, (size_t)nv * sizeof(double), cudaMemcpyDeviceToHost, p->s_out));
        }
        CUDA_TRY(cudaStreamSynchronize(p->s_out));
        CUDA_TRY(cudaStreamSynchronize(p->stream));
        return 0;
    }
    CUDA_TRY(cudaMemcpyAsync(p->d_psi, psi, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, p->stream));
    if (launch(p, npsi, p->d_psi, p->d_Hpsi, p->stream)) return 1;
    CUDA_TRY(cudaMemcpyAsync(Hpsi, p->d_Hpsi, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, p->stream));
    CUDA_TRY(cudaStreamSynchronize(p->stream));
    return 0;
}

extern "C" int64_t evr_sg4_plan_info(const evr_sg4_plan *p, int what)
{
    if (!p) return -1;
    if (!p->sub.empty()) return evr::multi_info(p, what);
    const int64_t nb0 = p->nb0;
    switch (what) {
    case EVR_INFO_LAUNCHES: return p->launches;
    case EVR_INFO_ALG_BYTES_NPSI1:
        if (p->op10) return p->S_local * nb0 * 8 * 2 + p->S_local * 4 * 2 + p->NQ_local * 8 * (nb0 * nb0 * p->n_var + (p->o10.sym ? p->o10.n_act * (p->o10.n_act + 1) / 2 : p->o10.n_act * p->o10.n_act) + 2) + p->nb * nb0 * 8 * 2;
        return p->S_local * nb0 * 8 * 2 + p->S_local * 4 * 2 + p->NQ_local * nb0 * nb0 * 8 * p->n_var + p->nb * nb0 * 8 * 2;
    case EVR_INFO_ALG_BYTES_PER_RHS_EXTRA:
        return p->S_local * nb0 * 8 * 2 + p->nb * nb0 * 8 * 2;
    case EVR_INFO_NQ_LOCAL: return p->NQ_local;
    case EVR_INFO_S_LOCAL: return p->S_local;
    case EVR_INFO_SMEM_BYTES: return (int64_t)(p->fast ? p->fclass_smem[0] : p->smem_bytes);
    case EVR_INFO_GRID_CTAS: return p->fast ? p->fclass_ctas[0] : p->grid_ctas;
    case EVR_INFO_PATH: return p->fast ? 1 : 0;
    case EVR_INFO_FLOPS_NPSI1: return p->flops_npsi1;
    case EVR_INFO_ISO: return (p->fast && p->fast_iso) ? 1 : 0;
    case EVR_INFO_GENERIC_TERMS: return p->fast ? p->n_rest_terms : p->n_terms;
    default: return -1;
    }
}

extern "C" int evr_sg4_plan_destroy(evr_sg4_plan **pp)
{
    if (!pp || !*pp) return 0;
    evr_sg4_plan *p = *pp;
    if (!p->sub.empty()) { evr::multi_destroy(p); delete p; *pp = nullptr; return 0; }
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    for (auto &g : p->graphs) cudaGraphExecDestroy(g.exec);
    p->graphs.clear();
    cudaFree(p->d_terms); cudaFree(p->d_lev); cudaFree(p->d_map); cudaFree(p->d_nq_of); cudaFree(p->d_nb_of);
    cudaFree(p->d_offB); cudaFree(p->d_offG); cudaFree(p->d_B); cudaFree(p->d_BTw); cudaFree(p->d_D1); cudaFree(p->d_D2);
    cudaFree(p->d_opterms); cudaFree(p->d_grids); cudaFree(p->d_psi); cudaFree(p->d_Hpsi);
    cudaFree(p->d_fterms); cudaFree(p->d_fmap); cudaFree(p->d_fmats); cudaFree(p->d_fV);
    cudaFree(p->d_fpos); cudaFree(p->d_gmap); cudaFree(p->d_perm); cudaFree(p->d_psi_int); cudaFree(p->d_Hpsi_int);
    cudaFree(p->d_GG); cudaFree(p->d_Jac); cudaFree(p->d_sq);
    cudaFree(p->d_det_off); cudaFree(p->d_det_ent); cudaFree(p->d_stage); cudaFree(p->d_fcounters); cudaFree(p->d_gscratch); cudaFree(p->d_nscratch); cudaFree(p->d_rscratch); cudaFree(p->d_rlist); cudaFree(p->d_inv_perm);
    if (p->s_in) { cudaStreamDestroy(p->s_in); cudaStreamDestroy(p->s_out); cudaEventDestroy(p->ev_pin); cudaEventDestroy(p->ev_pout); }
    if (p->stream) cudaStreamDestroy(p->stream);
    for (int c = 0; c < EVR_MAX_FCLASSES; ++c) { if (p->side[c]) cudaStreamDestroy(p->side[c]); if (p->ev_join[c]) cudaEventDestroy(p->ev_join[c]); }
    if (p->ev_fork) cudaEventDestroy(p->ev_fork);
    delete p;
    *pp = nullptr;
    return 0;
}
