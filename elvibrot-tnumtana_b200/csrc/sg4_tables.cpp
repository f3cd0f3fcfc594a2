// sg4_tables.cpp -- host-side builder of the Smolyak type-4 index/term tables.
//
// Produces tables bit-identical to the reference's, but not by the
// reference's procedure: the reference enumerates multi-indices one at a time with
// ADD_ONE_TO_nDindex and locates each term-local basis function in the packed basis
// by an estimate + linear search (calc_nDI).  Here both enumerations are direct
// recursions and the packed position is a closed-form lexicographic RANK obtained
// from suffix-count tables, so the mapping table costs O(S * D * nb_k) integer adds
// and is embarrassingly parallel over terms.
//
// Reference semantics reproduced (file:line in the reference tree):
//   nDindB            type 5 : all (i_1..i_D), 1<=i_k<=nb_k(LG), sum_k l_k(i_k) <= LB,
//                     LAST index fastest   (sub_module_nDindex.f90:971-1082, 2396-2462)
//   l_k(i)            level at which 1-D function i first appears = Tab_L of the
//                     level-LG primitive  (sub_module_basis.f90:699-706)
//   nDind_SmolyakRep  type -5: all l, l_k>=0, Lmin<=sum l<=LG, FIRST index fastest
//                     (sub_module_nDindex.f90:1083-1185, 2463-2517)
//   WeightSG          (-1)^dL C(D-1,dL), dL = LG - sum l (sub_module_param_SGType2.f90:784-806)
//   tab_nq/nb_OF_SRep, inclusive sums      (sub_quadra_SparseBasis.f90:1315-1345)
//   tab_iB_OF_SRep_TO_iB: term-local basis index (FIRST index fastest) -> packed
//                     position, 0 when sum_k l_k(i_k) > LB
//                     (sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4.f90:625-949)
#include "../../include/evr_sg4.h"
#include "sg4_internal.h"

#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

struct evr_sg4_tables {
    int D = 0, LB = 0, LG = 0, Lmin = 0;
    int64_t nb_SG = 0, nb = 0, S = 0, NQ = 0, count0 = 0;
    std::vector<int32_t> nq_of, nb_of;      // [k*(LG+1)+L]
    std::vector<std::vector<int>> lev;      // lev[k][i-1] = level of 1-D function i of mode k
    std::vector<int32_t> tab_l;             // [iG*D+k]
    std::vector<double>  weight;
    std::vector<int32_t> tab_nq, tab_nb;
    std::vector<int64_t> sum_nq, sum_nb;
    std::vector<int32_t> packedB;           // [iB*D+k]
    std::vector<int32_t> map;               // [S]
};

namespace {

// enumerate the term table: first index fastest  <=>  last index is the outermost loop
void enum_terms(int D, int Lmin, int LG, std::vector<int32_t> &out)
{
    std::vector<int> l(D, 0);
    // odometer over (l_D outermost ... l_1 innermost) with pruning on the running sum
    // Implemented iteratively: generate all compositions with sum<=LG, keep sum>=Lmin.
    // Recursion depth D, order: for l_D = 0..LG: for l_{D-1} = 0..LG-l_D: ... l_1 fastest.
    struct Rec {
        int D, Lmin, LG; std::vector<int> &l; std::vector<int32_t> &out;
        void go(int k, int budget) {           // choose l[k], k from D-1 down to 0
            if (k < 0) {
                int s = LG - budget;
                if (s >= Lmin) for (int j = 0; j < D; ++j) out.push_back(l[j]);
                return;
            }
            for (int v = 0; v <= budget; ++v) { l[k] = v; go(k - 1, budget - v); }
            l[k] = 0;
        }
    } r{D, Lmin, LG, l, out};
    r.go(D - 1, LG);
}

// enumerate the packed basis: last index fastest <=> first index is the outermost loop
void enum_packed(int D, int LB, const std::vector<std::vector<int>> &lev, std::vector<int32_t> &out)
{
    std::vector<int> idx(D, 1);
    struct Rec {
        int D; const std::vector<std::vector<int>> &lev; std::vector<int> &idx; std::vector<int32_t> &out;
        void go(int k, int budget) {
            if (k == D) { for (int j = 0; j < D; ++j) out.push_back(idx[j]); return; }
            const std::vector<int> &lv = lev[k];
            for (size_t i = 0; i < lv.size(); ++i) {
                if (lv[i] > budget) break;          // levels are non-decreasing in i
                idx[k] = (int)i + 1;
                go(k + 1, budget - lv[i]);
            }
        }
    } r{D, lev, idx, out};
    r.go(0, LB);
}

double binom(int n, int k)
{
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    return (double)(int64_t)(r + 0.5);
}

} // namespace

extern "C" int evr_sg4_tables_build(evr_sg4_tables **out, int D, int LB, int LG,
                                    const int32_t *nq_of, const int32_t *nb_of)
{
    if (!out || !nq_of || !nb_of) return evr::fail("evr_sg4_tables_build: null argument");
    if (D < 1 || D > EVR_MAXD) return evr::fail("evr_sg4_tables_build: D out of range [1," + std::to_string(EVR_MAXD) + "]");
    if (LG < 0 || LB < 0 || LG > 127) return evr::fail("evr_sg4_tables_build: LB/LG out of range");
    auto *t = new evr_sg4_tables();
    t->D = D; t->LB = LB; t->LG = LG;
    t->Lmin = std::max(0, LG - D + 1);
    t->nq_of.assign(nq_of, nq_of + (size_t)D * (LG + 1));
    t->nb_of.assign(nb_of, nb_of + (size_t)D * (LG + 1));
    for (int k = 0; k < D; ++k)
        for (int L = 0; L <= LG; ++L) {
            int q = t->nq_of[k * (LG + 1) + L], b = t->nb_of[k * (LG + 1) + L];
            if (q < 1 || b < 1) { delete t; return evr::fail("evr_sg4_tables_build: nq/nb must be >= 1"); }
            if (L > 0 && b < t->nb_of[k * (LG + 1) + L - 1]) { delete t; return evr::fail("evr_sg4_tables_build: nb_k(L) must be non-decreasing in L"); }
        }
    // level of each 1-D function: smallest L with nb_k(L) >= i
    t->lev.resize(D);
    for (int k = 0; k < D; ++k) {
        int nbmax = t->nb_of[k * (LG + 1) + LG];
        t->lev[k].assign(nbmax, LG);
        for (int L = LG; L >= 0; --L) {
            int n = std::min(nbmax, (int)t->nb_of[k * (LG + 1) + L]);
            for (int i = 0; i < n; ++i) t->lev[k][i] = L;
        }
    }
    enum_packed(D, LB, t->lev, t->packedB);
    t->nb = (int64_t)(t->packedB.size() / D);
    enum_terms(D, t->Lmin, LG, t->tab_l);
    t->nb_SG = (int64_t)(t->tab_l.size() / D);
    if (t->nb >= (int64_t)1 << 31) { delete t; return evr::fail("evr_sg4_tables_build: packed basis exceeds int32"); }

    t->weight.resize(t->nb_SG);
    t->tab_nq.resize(t->nb_SG); t->tab_nb.resize(t->nb_SG);
    t->sum_nq.resize(t->nb_SG); t->sum_nb.resize(t->nb_SG);
    int64_t nqq = 0, nbb = 0;
    for (int64_t iG = 0; iG < t->nb_SG; ++iG) {
        int s = 0; int64_t nq = 1, nbT = 1;
        for (int k = 0; k < D; ++k) {
            int l = t->tab_l[iG * D + k];
            s += l;
            nq *= t->nq_of[k * (LG + 1) + l];
            nbT *= t->nb_of[k * (LG + 1) + l];
        }
        int dL = LG - s;
        t->weight[iG] = (dL < 0 || dL > D - 1) ? 0.0 : ((dL % 2 == 0) ? 1.0 : -1.0) * binom(D - 1, dL);
        if (nq >= (int64_t)1 << 31 || nbT >= (int64_t)1 << 31) { delete t; return evr::fail("evr_sg4_tables_build: term size exceeds int32"); }
        nqq += nq; nbb += nbT;
        t->tab_nq[iG] = (int32_t)nq; t->tab_nb[iG] = (int32_t)nbT;
        t->sum_nq[iG] = nqq; t->sum_nb[iG] = nbb;
    }
    t->NQ = nqq; t->S = nbb;

    // suffix counts: cnt[k][r] = number of (i_{k+1}..i_D) (0-based modes k..D-1) with level sum <= r
    std::vector<std::vector<int64_t>> cnt(D + 1, std::vector<int64_t>(LB + 1, 0));
    for (int r = 0; r <= LB; ++r) cnt[D][r] = 1;
    for (int k = D - 1; k >= 0; --k)
        for (int r = 0; r <= LB; ++r) {
            int64_t c = 0;
            for (int lv : t->lev[k]) { if (lv > r) break; c += cnt[k + 1][r - lv]; }
            cnt[k][r] = c;
        }
    t->map.assign((size_t)t->S, 0);
    int64_t count0 = 0;
#pragma omp parallel for schedule(dynamic, 32) reduction(+:count0)
    for (int64_t iG = 0; iG < t->nb_SG; ++iG) {
        int nbk[EVR_MAXD], ib[EVR_MAXD];
        for (int k = 0; k < D; ++k) { nbk[k] = t->nb_of[k * (LG + 1) + t->tab_l[iG * D + k]]; ib[k] = 1; }
        int64_t base = t->sum_nb[iG] - t->tab_nb[iG];
        for (int64_t j = 0; j < t->tab_nb[iG]; ++j) {
            // rank of ib(:) in the packed (last-index-fastest, level-constrained) order
            int budget = LB; int64_t rank = 1; bool in = true;
            for (int k = 0; k < D; ++k) {
                const std::vector<int> &lv = t->lev[k];
                int lk = lv[ib[k] - 1];
                if (lk > budget) { in = false; break; }
                for (int i = 0; i < ib[k] - 1; ++i) rank += cnt[k + 1][budget - lv[i]];
                budget -= lk;
            }
            if (in) t->map[base + j] = (int32_t)rank; else ++count0;
            // next term-local index, first index fastest
            for (int k = 0; k < D; ++k) { if (++ib[k] <= nbk[k]) break; ib[k] = 1; }
        }
    }
    t->count0 = count0;
    *out = t;
    return 0;
}

extern "C" int evr_sg4_tables_destroy(evr_sg4_tables **t)
{
    if (t && *t) { delete *t; *t = nullptr; }
    return 0;
}

extern "C" int64_t evr_sg4_tables_size(const evr_sg4_tables *t, int what)
{
    if (!t) return -1;
    switch (what) {
    case EVR_TAB_NB_SG:   return t->nb_SG;
    case EVR_TAB_NB:      return t->nb;
    case EVR_TAB_S:       return t->S;
    case EVR_TAB_NQ:      return t->NQ;
    case EVR_TAB_COUNT0:  return t->count0;
    case EVR_TAB_LMIN:    return t->Lmin;
    case EVR_TAB_TAB_L:   return t->nb_SG * t->D;
    case EVR_TAB_WEIGHT:
    case EVR_TAB_TAB_NQ:
    case EVR_TAB_TAB_NB:
    case EVR_TAB_SUM_NQ:
    case EVR_TAB_SUM_NB:  return t->nb_SG;
    case EVR_TAB_PACKEDB: return t->nb * t->D;
    case EVR_TAB_MAP:     return t->S;
    default:              return -1;
    }
}

extern "C" int evr_sg4_tables_get(const evr_sg4_tables *t, int what, void *dst)
{
    if (!t || !dst) return evr::fail("evr_sg4_tables_get: null argument");
    auto cp = [&](const void *src, size_t bytes) { if (bytes) std::memcpy(dst, src, bytes); return 0; };
    switch (what) {
    case EVR_TAB_TAB_L:   return cp(t->tab_l.data(),   t->tab_l.size() * sizeof(int32_t));
    case EVR_TAB_WEIGHT:  return cp(t->weight.data(),  t->weight.size() * sizeof(double));
    case EVR_TAB_TAB_NQ:  return cp(t->tab_nq.data(),  t->tab_nq.size() * sizeof(int32_t));
    case EVR_TAB_TAB_NB:  return cp(t->tab_nb.data(),  t->tab_nb.size() * sizeof(int32_t));
    case EVR_TAB_SUM_NQ:  return cp(t->sum_nq.data(),  t->sum_nq.size() * sizeof(int64_t));
    case EVR_TAB_SUM_NB:  return cp(t->sum_nb.data(),  t->sum_nb.size() * sizeof(int64_t));
    case EVR_TAB_PACKEDB: return cp(t->packedB.data(), t->packedB.size() * sizeof(int32_t));
    case EVR_TAB_MAP:     return cp(t->map.data(),     t->map.size() * sizeof(int32_t));
    default:              return evr::fail("evr_sg4_tables_get: unknown table id");
    }
}

// ini_iGs_MPI (sub_module_basis_BtoG_GtoB_SG4_MPI.f90:639-669): rank r owns
// q = nb_SG/np terms, the first (nb_SG mod np) ranks one more; contiguous ranges.
extern "C" int evr_sg4_ini_iGs(int nb_SG, int np, int rank, int *iG_begin, int *iG_end)
{
    if (np < 1 || rank < 0 || rank >= np || nb_SG < 0 || !iG_begin || !iG_end)
        return evr::fail("evr_sg4_ini_iGs: bad arguments");
    int q = nb_SG / np, rem = nb_SG % np;
    int b = rank * q + std::min(rank, rem);
    int e = b + q + (rank < rem ? 1 : 0);
    *iG_begin = b; *iG_end = e;
    return 0;
}

// contiguous ranges with (nearly) equal cumulative cost: boundary r is the first term at which the
// inclusive prefix sum reaches r/np of the total.
extern "C" int evr_sg4_balanced_iGs(int nb_SG, const int32_t *cost, int np, int rank, int *iG_begin, int *iG_end)
{
    if (np < 1 || rank < 0 || rank >= np || nb_SG < 0 || !cost || !iG_begin || !iG_end)
        return evr::fail("evr_sg4_balanced_iGs: bad arguments");
    int64_t total = 0;
    for (int i = 0; i < nb_SG; ++i) { if (cost[i] < 0) return evr::fail("evr_sg4_balanced_iGs: negative cost"); total += cost[i]; }
    auto boundary = [&](int r) {
        if (r <= 0) return 0;
        if (r >= np) return nb_SG;
        const double target = (double)total * r / np;
        int64_t acc = 0;
        for (int i = 0; i < nb_SG; ++i) { acc += cost[i]; if ((double)acc >= target) return i + 1; }
        return nb_SG;
    };
    *iG_begin = boundary(rank);
    *iG_end = boundary(rank + 1);
    if (*iG_end < *iG_begin) *iG_end = *iG_begin;
    return 0;
}
