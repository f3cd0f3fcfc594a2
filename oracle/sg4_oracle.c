/*
 * sg4_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference algorithm for H|psi> on the Smolyak
 * type-4 sparse grid (ElVibRot-TnumTana).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this file's
 * shared object.  The product path (elvibrot-tnumtana_b200/csrc) never links,
 * calls or falls back to anything in here.
 *
 * Parity status: the reference (Fortran 2003) cannot be compiled in this
 * container (no Fortran compiler), so this restatement is pinned by the
 * reference's own golden outputs instead:
 *   - integer tables (term list, weights, sizes, "count 0") against
 *     UnitTests/HNO3_UT/RES_old/res_HNO3_RPH_LB*-LG*.gz and
 *     UnitTests/HCN_UT/RES_old/res_RPH_AutoContract_Davidson_SG4.gz
 *   - the H action end-to-end against the eigenvalue known-answer files
 *     Working_tests/MPI_tests/{6D,21D}_Davidson_openMP/benchmark (1e-8 au)
 * (fixtures extracted to tests/golden/ by tests/golden/make_golden.py).
 *
 * Every function cites the reference file:line it follows (paths relative to
 * the reference root).  Nothing here is copied; the Fortran is re-expressed
 * with flat arrays.
 *
 * Conventions: all multi-indices and table entries keep the reference's
 * 1-based values; C arrays themselves are 0-based.  Column-major matrices.
 */
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MAXD 64

/* ------------------------------------------------------------------------
 * Basis_L_TO_n : n(L) = A + B*L**expo  (L_TO_n_type = 0)
 * ref: Source_ElVibRot/sub_Basis/sub_module_Basis_LTO_n.f90:307-311 (table),
 *      :431-440 (Get_n_FROM_Basis_L_TO_n: linear extrapolation past the table
 *      end with the last increment; NO extrapolation when the table has the
 *      single entry L=0).
 * ---------------------------------------------------------------------- */
static int ipow(int b, int e) { int r = 1; while (e-- > 0) r *= b; return r; }

int orc_n_of_L(int A, int B, int expo, int Ltab_max, int L)
{
    /* Ltab_max = ubound of tab_L_TO_n (the table holds L = 0..Ltab_max) */
    if (L <= Ltab_max) return A + B * ipow(L, expo);
    int nu = A + B * ipow(Ltab_max, expo);
    if (Ltab_max > 0) {
        int num1 = A + B * ipow(Ltab_max - 1, expo);
        nu += (L - Ltab_max) * (nu - num1);
    }
    return nu;
}

/* ------------------------------------------------------------------------
 * The SG4 table set.
 * ref: RecSparseGrid_ForDP_type4, Source_ElVibRot/sub_Basis/sub_quadra_SparseBasis.f90:924-1435
 *      param_SGType2,              Source_ElVibRot/sub_Basis/sub_module_param_SGType2.f90:54-102
 * ---------------------------------------------------------------------- */
typedef struct {
    int D, LB, LG, Lmin;
    int nb_SG;            /* number of Smolyak terms                     */
    int64_t nb;           /* packed basis size (nDindB%Max_nDI)          */
    int64_t S, NQ;        /* sum of term basis sizes / grid sizes        */
    int *nq_of, *nb_of;   /* [k*(LG+1)+L]  (Fortran tab(0:LG,1:D))       */
    int *i_to_l_off;      /* [D+1] offsets into i_to_l                   */
    int *i_to_l;          /* concatenated tab_i_TO_l(k)%vec(1:nb_k(LG))  */
    int *tab_l;           /* [iG*D+k] = Tab_nDval(k,iG) of nDind_SmolyakRep */
    double *weight;       /* WeightSG(iG)                                */
    int *tab_nq, *tab_nb; /* per term                                    */
    int64_t *sum_nq, *sum_nb; /* inclusive prefix sums                   */
    int *packedB;         /* [iB*D+k] = nDindB%Tab_nDval(k,iB)           */
    int32_t *map;         /* tab_iB_OF_SRep_TO_iB(1:S), 0 = dropped      */
    int64_t count0;
} orc_tables;

/* level sum of a packed-basis multi-index through tab_i_TO_l;
 * ref: calc_LL1L2_OF_nDindex_type5, sub_module_nDindex.f90:2518-2560
 * (index beyond the table contributes Lmax+1). */
static int L_of_ib(const orc_tables *t, const int *ib, int Lmax)
{
    int L = 0;
    for (int k = 0; k < t->D; ++k) {
        int n = t->i_to_l_off[k + 1] - t->i_to_l_off[k];
        if (ib[k] > n) L += Lmax + 1;
        else L += t->i_to_l[t->i_to_l_off[k] + ib[k] - 1];
    }
    return L;
}

/* ADD_ONE_TO_nDindex_type5p: next multi-index, LAST index fastest, keeping
 * sum of levels <= Lmax.  ref: sub_module_nDindex.f90:2396-2462.
 * returns 1 if a new in-list value was produced, 0 at the end. */
static int add_one_5p(const orc_tables *t, int *v, const int *vend, int Lmax)
{
    const int D = t->D;
    for (;;) {
        v[D - 1] += 1;
        int L = L_of_ib(t, v, Lmax);
        int inlist = (L <= Lmax);
        for (int k = 0; k < D && inlist; ++k) if (v[k] > vend[k]) inlist = 0;
        if (v[D - 1] > vend[D - 1] || L > Lmax) {
            for (int i = D - 1; i >= 1; --i) {
                v[i] = 1;
                v[i - 1] += 1;
                L = L_of_ib(t, v, Lmax);
                inlist = (L <= Lmax);
                for (int k = 0; k < D && inlist; ++k) if (v[k] > vend[k]) inlist = 0;
                if (inlist) break;
            }
        }
        if (v[0] > vend[0] || L > Lmax) return 0;
        if (inlist) return 1;
    }
}

/* ADD_ONE_TO_nDindex_type5m: next term multi-index l(:), FIRST index fastest,
 * Lmin <= sum(l) <= Lmax, values start at 0.  ref: sub_module_nDindex.f90:2463-2517. */
static int add_one_5m(int D, int *v, int vend, int Lmin, int Lmax)
{
    for (;;) {
        v[0] += 1;
        int L = 0; for (int k = 0; k < D; ++k) L += v[k];
        int inrange = (L <= Lmax && L >= Lmin);
        for (int k = 0; k < D && inrange; ++k) if (v[k] > vend) inrange = 0;
        if (v[0] > vend || L > Lmax) {
            for (int i = 0; i < D - 1; ++i) {
                v[i] = 0;
                v[i + 1] += 1;
                L = 0; for (int k = 0; k < D; ++k) L += v[k];
                inrange = (L <= Lmax && L >= Lmin);
                for (int k = 0; k < D && inrange; ++k) if (v[k] > vend) inrange = 0;
                if (inrange || L < Lmin) break;
            }
        }
        if (v[D - 1] > vend || L > Lmax) return 0;
        if (inrange) return 1;
    }
}

static double binomial(int n, int k)
{
    double r = 1.0;
    for (int i = 1; i <= k; ++i) r = r * (double)(n - k + i) / (double)i;
    return floor(r + 0.5);
}

void orc_tables_free(orc_tables *t)
{
    if (!t) return;
    free(t->nq_of); free(t->nb_of); free(t->i_to_l_off); free(t->i_to_l);
    free(t->tab_l); free(t->weight); free(t->tab_nq); free(t->tab_nb);
    free(t->sum_nq); free(t->sum_nb); free(t->packedB); free(t->map);
    free(t);
}

/*
 * Build every SG4 table from the per-mode level rules.
 *   nq_k(L) = Aq[k] + Bq[k]*L**expo_q[k]                (table to L,   :1062)
 *   nb_k(L) = get_n(table to min(L,LB), L)              (:1063-1064 + extrapolation)
 * legacy_LB0 != 0 reproduces ElVibRot <= 181.3 where nb(L) was not capped
 * when LB = 0 (the stored HNO3 LB0-LG3 log comes from that version).
 */
orc_tables *orc_tables_build(int D, int LB, int LG,
                             const int *Aq, const int *Bq, const int *expo_q,
                             const int *Ab, const int *Bb, const int *expo_b,
                             int legacy_LB0)
{
    if (D < 1 || D > ORC_MAXD || LG < 0 || LB < 0) return NULL;
    orc_tables *t = (orc_tables *)calloc(1, sizeof(orc_tables));
    t->D = D; t->LB = LB; t->LG = LG;
    /* ref :993-995 */
    t->Lmin = LG - D + 1; if (t->Lmin < 0) t->Lmin = 0;

    t->nq_of = (int *)malloc(sizeof(int) * D * (LG + 1));
    t->nb_of = (int *)malloc(sizeof(int) * D * (LG + 1));
    for (int k = 0; k < D; ++k)
        for (int L = 0; L <= LG; ++L) {
            t->nq_of[k * (LG + 1) + L] = orc_n_of_L(Aq[k], Bq[k], expo_q[k], L, L);
            int LB_L = (L < LB) ? L : LB;
            if (legacy_LB0 && LB == 0) LB_L = L;
            t->nb_of[k * (LG + 1) + L] = orc_n_of_L(Ab[k], Bb[k], expo_b[k], LB_L, L);
        }

    /* tab_i_TO_l(k)%vec = Tab_L of the level-LG primitive:
     * ref sub_module_basis.f90:699-706 (loop L = L_SparseBasis..0, Tab_L(1:nb(L)) = L)
     * where nb(L) is evaluated on the LEVEL-LG primitive's own L_TO_nb table
     * (tabulated to min(LG,LB)), i.e. Get_nb_FROM_l_OF_PrimBasis(L, prim(LG)). */
    t->i_to_l_off = (int *)malloc(sizeof(int) * (D + 1));
    t->i_to_l_off[0] = 0;
    for (int k = 0; k < D; ++k)
        t->i_to_l_off[k + 1] = t->i_to_l_off[k] + t->nb_of[k * (LG + 1) + LG];
    t->i_to_l = (int *)malloc(sizeof(int) * (t->i_to_l_off[D] > 0 ? t->i_to_l_off[D] : 1));
    for (int k = 0; k < D; ++k) {
        int *vec = t->i_to_l + t->i_to_l_off[k];
        int nbmax = t->nb_of[k * (LG + 1) + LG];
        for (int i = 0; i < nbmax; ++i) vec[i] = -1;
        int LBtab = (LG < LB) ? LG : LB;
        if (legacy_LB0 && LB == 0) LBtab = LG;
        for (int L = LG; L >= 0; --L) {
            int nbL = orc_n_of_L(Ab[k], Bb[k], expo_b[k], LBtab, L);
            if (nbL > nbmax) nbL = nbmax;
            for (int i = 0; i < nbL; ++i) vec[i] = L;
        }
    }

    /* ---- packed basis nDindB: type 5, Lmax = LB, last index fastest (:1168-1184) */
    int vend[ORC_MAXD], v[ORC_MAXD];
    for (int k = 0; k < D; ++k) vend[k] = t->nb_of[k * (LG + 1) + LG]; /* nDsize = nb(LG), nDinit = 1 */
    int64_t cap = 1024, n = 0;
    t->packedB = (int *)malloc(sizeof(int) * cap * D);
    for (int k = 0; k < D; ++k) v[k] = 1;
    v[D - 1] -= 1;
    while (add_one_5p(t, v, vend, LB)) {
        if (n == cap) { cap *= 2; t->packedB = (int *)realloc(t->packedB, sizeof(int) * cap * D); }
        memcpy(t->packedB + n * D, v, sizeof(int) * D);
        ++n;
    }
    t->nb = n;

    /* ---- Smolyak term table: type -5, Lmin..LG, first index fastest (:1239-1262) */
    cap = 1024; n = 0;
    t->tab_l = (int *)malloc(sizeof(int) * cap * D);
    for (int k = 0; k < D; ++k) v[k] = 0;
    v[0] -= 1;
    while (add_one_5m(D, v, LG, t->Lmin, LG)) {
        if (n == cap) { cap *= 2; t->tab_l = (int *)realloc(t->tab_l, sizeof(int) * cap * D); }
        memcpy(t->tab_l + n * D, v, sizeof(int) * D);
        ++n;
    }
    t->nb_SG = (int)n;

    /* ---- Smolyak weights, ref sub_module_param_SGType2.f90:784-806 */
    t->weight = (double *)malloc(sizeof(double) * t->nb_SG);
    for (int iG = 0; iG < t->nb_SG; ++iG) {
        int s = 0; for (int k = 0; k < D; ++k) s += t->tab_l[iG * D + k];
        int dL = LG - s;
        if (dL < 0 || dL > D - 1) t->weight[iG] = 0.0;
        else t->weight[iG] = ((dL % 2 == 0) ? 1.0 : -1.0) * binomial(D - 1, dL);
    }

    /* ---- per-term sizes and inclusive prefix sums, ref :1315-1345 */
    t->tab_nq = (int *)malloc(sizeof(int) * t->nb_SG);
    t->tab_nb = (int *)malloc(sizeof(int) * t->nb_SG);
    t->sum_nq = (int64_t *)malloc(sizeof(int64_t) * t->nb_SG);
    t->sum_nb = (int64_t *)malloc(sizeof(int64_t) * t->nb_SG);
    int64_t nqq = 0, nbb = 0;
    for (int iG = 0; iG < t->nb_SG; ++iG) {
        int64_t nq = 1, nb = 1;
        for (int k = 0; k < D; ++k) {
            int l = t->tab_l[iG * D + k];
            nq *= t->nq_of[k * (LG + 1) + l];
            nb *= t->nb_of[k * (LG + 1) + l];
        }
        nqq += nq; nbb += nb;
        t->tab_nq[iG] = (int)nq; t->tab_nb[iG] = (int)nb;
        t->sum_nq[iG] = nqq; t->sum_nb[iG] = nbb;
    }
    t->NQ = nqq; t->S = nbb;

    /* ---- mapping table, ref Set_tables_FOR_SmolyakRepBasis_TO_tabPackedBasis,
     * Source_ElVibRot/sub_Basis/sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4.f90:625-949.
     * Step 1 (:690-714): MaxnD_with_id_and_L(id,l) = number of packed functions
     * sharing one value of index id (whose level is l), counted inside the first
     * "block" of the enumeration. */
    const int nLB = LB + 1;
    int64_t *MaxnD = (int64_t *)calloc((size_t)D * nLB, sizeof(int64_t));
    {
        int64_t max_NBB = t->nb;
        for (int id = 0; id < D; ++id) {
            int ib = 1; int64_t iVal = 1;
            for (int64_t iBB = 0; iBB < max_NBB; ++iBB) {
                const int *nd = t->packedB + iBB * D;
                if (nd[id] != ib) { ib = nd[id]; iVal = 1; }
                int l = t->i_to_l[t->i_to_l_off[id] + ib - 1];
                MaxnD[id * nLB + l] = iVal;
                iVal += 1;
            }
            max_NBB = MaxnD[id * nLB + 0];
        }
    }
    t->map = (int32_t *)calloc((size_t)(t->S > 0 ? t->S : 1), sizeof(int32_t));
    int64_t count0 = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+:count0)
    for (int iG = 0; iG < t->nb_SG; ++iG) {
        int tab_nb[ORC_MAXD], tab_ib[ORC_MAXD];
        for (int k = 0; k < D; ++k) tab_nb[k] = t->nb_of[k * (LG + 1) + t->tab_l[iG * D + k]];
        int64_t iBSRep = t->sum_nb[iG] - t->tab_nb[iG];
        for (int k = 0; k < D; ++k) tab_ib[k] = 1;
        tab_ib[0] = 0;
        for (int iBDP = 0; iBDP < t->tab_nb[iG]; ++iBDP, ++iBSRep) {
            /* ADD_ONE_TO_nDval_m1 (first index fastest), sub_module_nDindex.f90:2613-2629 */
            tab_ib[0] += 1;
            for (int i = 0; i < D - 1; ++i) {
                if (tab_ib[i] > tab_nb[i]) { tab_ib[i] = 1; tab_ib[i + 1] += 1; } else break;
            }
            int LL = L_of_ib(t, tab_ib, LB);
            if (LL > LB) { count0 += 1; continue; }           /* stays 0 (:812) */
            /* first estimate of nDI (:815-826) */
            int64_t nDI = 1; LL = 0;
            for (int k = 0; k < D; ++k) {
                const int *vec = t->i_to_l + t->i_to_l_off[k];
                for (int iib = 1; iib <= tab_ib[k] - 1; ++iib) {
                    int l = LL + vec[iib - 1];
                    if (l <= LB) nDI += MaxnD[k * nLB + l];
                }
                LL += vec[tab_ib[k] - 1];
            }
            /* calc_nDI (packed): bidirectional search from the estimate,
             * sub_module_nDindex.f90:2674-2720 */
            if (nDI < 1 || nDI > t->nb) nDI = 1;
            int64_t ibp = nDI, ibm = nDI, found = -1;
            if (memcmp(t->packedB + (nDI - 1) * D, tab_ib, sizeof(int) * D) == 0) found = nDI;
            while (found < 0) {
                if (ibp < t->nb) {
                    ++ibp;
                    if (memcmp(t->packedB + (ibp - 1) * D, tab_ib, sizeof(int) * D) == 0) { found = ibp; break; }
                }
                if (ibm > 1) {
                    --ibm;
                    if (memcmp(t->packedB + (ibm - 1) * D, tab_ib, sizeof(int) * D) == 0) { found = ibm; break; }
                }
                if (ibm == 1 && ibp == t->nb) break;
            }
            if (found > 0) t->map[iBSRep] = (int32_t)found;
            else count0 += 1;
        }
    }
    free(MaxnD);
    t->count0 = count0;
    return t;
}

/* accessors for ctypes */
int     orc_tables_D(const orc_tables *t)      { return t->D; }
int     orc_tables_Lmin(const orc_tables *t)   { return t->Lmin; }
int     orc_tables_nb_SG(const orc_tables *t)  { return t->nb_SG; }
int64_t orc_tables_nb(const orc_tables *t)     { return t->nb; }
int64_t orc_tables_S(const orc_tables *t)      { return t->S; }
int64_t orc_tables_NQ(const orc_tables *t)     { return t->NQ; }
int64_t orc_tables_count0(const orc_tables *t) { return t->count0; }
const int    *orc_tables_nq_of(const orc_tables *t)   { return t->nq_of; }
const int    *orc_tables_nb_of(const orc_tables *t)   { return t->nb_of; }
const int    *orc_tables_tab_l(const orc_tables *t)   { return t->tab_l; }
const double *orc_tables_weight(const orc_tables *t)  { return t->weight; }
const int    *orc_tables_tab_nq(const orc_tables *t)  { return t->tab_nq; }
const int    *orc_tables_tab_nb(const orc_tables *t)  { return t->tab_nb; }
const int64_t*orc_tables_sum_nq(const orc_tables *t)  { return t->sum_nq; }
const int64_t*orc_tables_sum_nb(const orc_tables *t)  { return t->sum_nb; }
const int    *orc_tables_packedB(const orc_tables *t) { return t->packedB; }
const int32_t*orc_tables_map(const orc_tables *t)     { return t->map; }
const int    *orc_tables_i_to_l(const orc_tables *t)  { return t->i_to_l; }
const int    *orc_tables_i_to_l_off(const orc_tables *t) { return t->i_to_l_off; }

/* ------------------------------------------------------------------------
 * One mode product  out = (I x ... x M x ... x I) in
 *   in [a + left*(b + n_in *c)],  out[a + left*(q + n_out*c)],  M(n_out,n_in) col-major.
 * ref: the matmul on rank-1 slices RTemp(iq,:,ib) in BDP_TO_GDP_OF_SmolyakRep
 *      (sub_module_basis_BtoG_GtoB_SG4.f90:2457-2459), GDP_TO_BDP (:2549-2555)
 *      and DerivOp_TO_RDP_OF_SmolaykRep (:2764-2768).
 * ---------------------------------------------------------------------- */
static void mode_apply(const double *M, int n_out, int n_in,
                       const double *in, double *out, int64_t left, int64_t right)
{
    for (int64_t c = 0; c < right; ++c)
        for (int64_t a = 0; a < left; ++a) {
            const double *x = in + a + left * (int64_t)n_in * c;
            double *y = out + a + left * (int64_t)n_out * c;
            for (int q = 0; q < n_out; ++q) {
                double s = 0.0;
                for (int b = 0; b < n_in; ++b) s += M[q + (int64_t)n_out * b] * x[left * b];
                y[left * q] = s;
            }
        }
}

/* canonical offsets of the concatenated 1-D tables: for k = 0..D-1, L = 0..LG
 * (L fastest):  B(nq,nb), BTw(nb,nq), D1(nq,nq), D2(nq,nq). */
static void table_offsets(int D, int LG, const int *nq_of, const int *nb_of,
                          int64_t *offB, int64_t *offG)
{
    int64_t ob = 0, og = 0;
    for (int k = 0; k < D; ++k)
        for (int L = 0; L <= LG; ++L) {
            int i = k * (LG + 1) + L;
            offB[i] = ob; offG[i] = og;
            ob += (int64_t)nq_of[i] * nb_of[i];
            og += (int64_t)nq_of[i] * nq_of[i];
        }
}

/*
 * H|psi> on the SG4 grid for npsi real right-hand sides, type_Op = 0 or 1.
 * ref: sub_TabOpPsi_FOR_SGtype4          sub_OpPsi_SG4.f90:678-979   (term loop, OMP static)
 *      tabPackedBasis_TO_tabR_AT_iG      ...BtoG_GtoB_SG4.f90:1176-1246 (gather)
 *      sub_TabOpPsi_OF_ONEDP_FOR_SGtype4 sub_OpPsi_SG4.f90:1354-1546 (per-term operator)
 *      tabR_AT_iG_TO_tabPackedBasis      ...BtoG_GtoB_SG4.f90:1250-1289 (weighted atomic scatter)
 *
 * term_mode[2*iterm+{0,1}]: 1-based SG4 mode owning each derivative index of
 *   derive_termQdyn(:,iterm) (0 = no derivative); equal modes -> dnRGG%d2,
 *   one mode -> d1, two different modes -> d1 then d1 (Get_MatdnRGG,
 *   sub_module_basis_set_alloc.f90:1966-2020).
 * grids[iterm]: NULL for grid_zero/grid_cte terms, else Grid(NQ,nb0,nb0) column-major
 *   over the whole Smolyak grid (term offset = tab_Sum_nq - nq).
 * Mat_cte[iterm*nb0*nb0 + j + nb0*i] = Mat_cte(j,i) of term iterm.
 * psi / Hpsi: [ipsi*nb*nb0 + (ib0)*nb + iB].
 * Terms iG in [iG_begin, iG_end) (0-based) are applied; Hpsi is zeroed first
 * when zero_out != 0 (reference zeroes OpPsi, :765).
 */
int orc_tab_oppsi(int D, int nb_SG, int nb0, int64_t nb, int LG,
                  const int *tab_l, const double *W,
                  const int *tab_nq, const int *tab_nb, const int32_t *map,
                  const int *nq_of, const int *nb_of,
                  const double *Bm, const double *BTw, const double *D1, const double *D2,
                  int type_Op, int nb_Term, const int *term_mode,
                  const unsigned char *grid_zero, const unsigned char *grid_cte,
                  const double *Mat_cte, const double *const *grids,
                  int npsi, const double *psi, double *Hpsi,
                  int nthreads, int iG_begin, int iG_end, int zero_out)
{
    if (npsi < 1) return 1;                      /* ref :738-743 STOP size(Psi)=0 */
    if (type_Op != 0 && type_Op != 1) return 2;
    const int nT = D * (LG + 1);
    int64_t *offB = (int64_t *)malloc(sizeof(int64_t) * nT);
    int64_t *offG = (int64_t *)malloc(sizeof(int64_t) * nT);
    table_offsets(D, LG, nq_of, nb_of, offB, offG);
    int64_t *sum_nq = (int64_t *)malloc(sizeof(int64_t) * (nb_SG + 1));
    int64_t *sum_nb = (int64_t *)malloc(sizeof(int64_t) * (nb_SG + 1));
    sum_nq[0] = sum_nb[0] = 0;
    int64_t maxn = 1;   /* largest intermediate of any term: prod_k max(nq_k, nb_k) */
    for (int iG = 0; iG < nb_SG; ++iG) {
        sum_nq[iG + 1] = sum_nq[iG] + tab_nq[iG];
        sum_nb[iG + 1] = sum_nb[iG] + tab_nb[iG];
        int64_t m = 1;
        for (int k = 0; k < D; ++k) {
            int l = tab_l[iG * D + k];
            int a = nq_of[k * (LG + 1) + l], b = nb_of[k * (LG + 1) + l];
            m *= (a > b) ? a : b;
        }
        if (m > maxn) maxn = m;
    }
    const int64_t NQ = sum_nq[nb_SG];
    const int64_t nvec = nb * nb0;
    if (zero_out) memset(Hpsi, 0, sizeof(double) * (size_t)nvec * npsi);
    if (nthreads < 1) nthreads = 1;

#pragma omp parallel num_threads(nthreads)
    {
        double *X   = (double *)malloc(sizeof(double) * (size_t)maxn);          /* ping  */
        double *Y   = (double *)malloc(sizeof(double) * (size_t)maxn);          /* pong  */
        double *Pg  = (double *)malloc(sizeof(double) * (size_t)maxn * nb0);    /* psi on grid (nq,nb0)   */
        double *Pch = (double *)malloc(sizeof(double) * (size_t)maxn * nb0);    /* Psi_ch (nq,nb0)        */
        double *Op  = (double *)malloc(sizeof(double) * (size_t)maxn * nb0);    /* OpPsi (nq,nb0)         */
        int tnq[ORC_MAXD], tnb[ORC_MAXD], lk[ORC_MAXD];
#pragma omp for schedule(static)
        for (int iG = iG_begin; iG < iG_end; ++iG) {
            const int nq = tab_nq[iG], nbT = tab_nb[iG];
            for (int k = 0; k < D; ++k) {
                lk[k] = tab_l[iG * D + k];
                tnq[k] = nq_of[k * (LG + 1) + lk[k]];
                tnb[k] = nb_of[k * (LG + 1) + lk[k]];
            }
            const int32_t *mp = map + sum_nb[iG];
            for (int ip = 0; ip < npsi; ++ip) {
                const double *x = psi + (int64_t)ip * nvec;
                double *y = Hpsi + (int64_t)ip * nvec;
                for (int ib0 = 0; ib0 < nb0; ++ib0) {
                    /* gather (:1217-1235) */
                    for (int j = 0; j < nbT; ++j) {
                        int32_t m = mp[j];
                        X[j] = (m > 0 && m <= nb) ? x[(int64_t)ib0 * nb + m - 1] : 0.0;
                    }
                    /* B -> G, mode 1 first (:2434-2477) */
                    double *a = X, *b = Y;
                    int64_t left = 1, right = nbT;
                    for (int k = 0; k < D; ++k) {
                        right /= tnb[k];
                        mode_apply(Bm + offB[k * (LG + 1) + lk[k]], tnq[k], tnb[k], a, b, left, right);
                        left *= tnq[k];
                        double *tmp = a; a = b; b = tmp;
                    }
                    memcpy(Pg + (int64_t)ib0 * nq, a, sizeof(double) * nq);
                }
                memset(Op, 0, sizeof(double) * (size_t)nq * nb0);
                const int nterm = (type_Op == 0) ? 1 : nb_Term;
                for (int it = 0; it < nterm; ++it) {
                    if (grid_zero && grid_zero[it]) continue;                 /* :1511 */
                    const int m1 = (type_Op == 0) ? 0 : term_mode[2 * it];
                    const int m2 = (type_Op == 0) ? 0 : term_mode[2 * it + 1];
                    /* Psi_ch = copy of psi-grid, then derivative(s) (:1513-1517) */
                    for (int jb0 = 0; jb0 < nb0; ++jb0) {
                        double *dst = Pch + (int64_t)jb0 * nq;
                        memcpy(dst, Pg + (int64_t)jb0 * nq, sizeof(double) * nq);
                        if (m1 == 0 && m2 == 0) continue;
                        int64_t left = 1, right = nq;
                        for (int k = 0; k < D; ++k) {            /* loop over ibasis (:2747-2787) */
                            right /= tnq[k];
                            const int own1 = (m1 == k + 1), own2 = (m2 == k + 1);
                            if (own1 || own2) {
                                const double *M = (own1 && own2) ? D2 + offG[k * (LG + 1) + lk[k]]
                                                                 : D1 + offG[k * (LG + 1) + lk[k]];
                                mode_apply(M, tnq[k], tnq[k], dst, X, left, right);
                                memcpy(dst, X, sizeof(double) * nq);
                            }
                            left *= tnq[k];
                        }
                    }
                    /* Op(:,i) += GridOp(:,i,j,iterm) * Psi_ch(:,j)  (:1521-1525) */
                    const int cte = (grid_cte && grid_cte[it]);
                    for (int ib0 = 0; ib0 < nb0; ++ib0)
                        for (int jb0 = 0; jb0 < nb0; ++jb0) {
                            double *o = Op + (int64_t)ib0 * nq;
                            const double *p = Pch + (int64_t)jb0 * nq;
                            if (cte) {
                                const double c = Mat_cte[(int64_t)it * nb0 * nb0 + ib0 + nb0 * jb0];
                                for (int q = 0; q < nq; ++q) o[q] += c * p[q];
                            } else {
                                const double *g = grids[it] + sum_nq[iG] + NQ * (ib0 + (int64_t)nb0 * jb0);
                                for (int q = 0; q < nq; ++q) o[q] += g[q] * p[q];
                            }
                        }
                }
                /* G -> B (:2534-2571) and weighted scatter (:1271-1287) */
                for (int ib0 = 0; ib0 < nb0; ++ib0) {
                    memcpy(X, Op + (int64_t)ib0 * nq, sizeof(double) * nq);
                    double *a = X, *b = Y;
                    int64_t left = 1, right = nq;
                    for (int k = 0; k < D; ++k) {
                        right /= tnq[k];
                        mode_apply(BTw + offB[k * (LG + 1) + lk[k]], tnb[k], tnq[k], a, b, left, right);
                        left *= tnb[k];
                        double *tmp = a; a = b; b = tmp;
                    }
                    const double w = W[iG];
                    for (int j = 0; j < nbT; ++j) {
                        int32_t m = mp[j];
                        if (m > 0 && m <= nb) {
                            double val = w * a[j];
#pragma omp atomic
                            y[(int64_t)ib0 * nb + m - 1] += val;
                        }
                    }
                }
            }
        }
        free(X); free(Y); free(Pg); free(Pch); free(Op);
    }
    free(offB); free(offG); free(sum_nq); free(sum_nb);
    return 0;
}

/*
 * type_Op = 10:  H = -1/2 J^-1 rho^-1/2 d_i [ J G^ij d_j ( rho^1/2 ... ) ] + V   with the metric tensor,
 * Jacobian and sqrt(rho/J) CACHED per grid point (the reference recomputes G with Tnum at
 * every call, get_OpGrid_type10_OF_ONEDP_FOR_SG4, sub_OpPsi_SG4.f90:2717-2728).
 * ref: sub_TabOpPsi_OF_ONEDP_FOR_SGtype4, CASE (10), sub_OpPsi_SG4.f90:1548-1650:
 *   VPsi(:,i) = sum_j V(:,i,j) psi(:,j)
 *   per channel: phi = psi*sqRhoOVERJac ; phi_j = d/dQ_j phi ; chi_i = Jac * sum_j GG(:,j,i) phi_j ;
 *                Op = -1/2 sum_i d/dQ_i chi_i / (Jac*sqRhoOVERJac) + VPsi
 * act_mode[j] = 1-based SG4 mode owning active coordinate j.  V may be NULL (no potential).
 * GG[iq + NQ*(j + n*i)] = GGiq(iq,j,i) on the whole Smolyak grid; Jac, sq: [NQ].
 */
int orc_tab_oppsi10(int D, int nb_SG, int nb0, int64_t nb, int LG,
                    const int *tab_l, const double *W,
                    const int *tab_nq, const int *tab_nb, const int32_t *map,
                    const int *nq_of, const int *nb_of,
                    const double *Bm, const double *BTw, const double *D1,
                    int n_act, const int *act_mode,
                    const double *V, const double *GG, const double *Jac, const double *sq,
                    int npsi, const double *psi, double *Hpsi,
                    int nthreads, int iG_begin, int iG_end, int zero_out)
{
    if (npsi < 1) return 1;
    if (n_act < 1 || n_act > ORC_MAXD) return 2;
    const int nT = D * (LG + 1);
    int64_t *offB = (int64_t *)malloc(sizeof(int64_t) * nT);
    int64_t *offG = (int64_t *)malloc(sizeof(int64_t) * nT);
    table_offsets(D, LG, nq_of, nb_of, offB, offG);
    int64_t *sum_nq = (int64_t *)malloc(sizeof(int64_t) * (nb_SG + 1));
    int64_t *sum_nb = (int64_t *)malloc(sizeof(int64_t) * (nb_SG + 1));
    sum_nq[0] = sum_nb[0] = 0;
    int64_t maxn = 1;
    for (int iG = 0; iG < nb_SG; ++iG) {
        sum_nq[iG + 1] = sum_nq[iG] + tab_nq[iG];
        sum_nb[iG + 1] = sum_nb[iG] + tab_nb[iG];
        int64_t m = 1;
        for (int k = 0; k < D; ++k) {
            int l = tab_l[iG * D + k];
            int a = nq_of[k * (LG + 1) + l], b = nb_of[k * (LG + 1) + l];
            m *= (a > b) ? a : b;
        }
        if (m > maxn) maxn = m;
    }
    const int64_t NQ = sum_nq[nb_SG];
    const int64_t nvec = nb * nb0;
    if (zero_out) memset(Hpsi, 0, sizeof(double) * (size_t)nvec * npsi);
    if (nthreads < 1) nthreads = 1;

#pragma omp parallel num_threads(nthreads)
    {
        double *X    = (double *)malloc(sizeof(double) * (size_t)maxn);
        double *Y    = (double *)malloc(sizeof(double) * (size_t)maxn);
        double *Pg   = (double *)malloc(sizeof(double) * (size_t)maxn * nb0);
        double *VPsi = (double *)malloc(sizeof(double) * (size_t)maxn * nb0);
        double *Rj   = (double *)malloc(sizeof(double) * (size_t)maxn * n_act);
        double *Ri   = (double *)malloc(sizeof(double) * (size_t)maxn);
        double *Op   = (double *)malloc(sizeof(double) * (size_t)maxn);
        int tnq[ORC_MAXD], tnb[ORC_MAXD], lk[ORC_MAXD];
#pragma omp for schedule(static)
        for (int iG = iG_begin; iG < iG_end; ++iG) {
            const int nq = tab_nq[iG], nbT = tab_nb[iG];
            for (int k = 0; k < D; ++k) {
                lk[k] = tab_l[iG * D + k];
                tnq[k] = nq_of[k * (LG + 1) + lk[k]];
                tnb[k] = nb_of[k * (LG + 1) + lk[k]];
            }
            const int32_t *mp = map + sum_nb[iG];
            const int64_t g0 = sum_nq[iG];
            for (int ip = 0; ip < npsi; ++ip) {
                const double *x = psi + (int64_t)ip * nvec;
                double *y = Hpsi + (int64_t)ip * nvec;
                for (int ib0 = 0; ib0 < nb0; ++ib0) {           /* gather + B -> G */
                    for (int j = 0; j < nbT; ++j) {
                        int32_t m = mp[j];
                        X[j] = (m > 0 && m <= nb) ? x[(int64_t)ib0 * nb + m - 1] : 0.0;
                    }
                    double *a = X, *b = Y;
                    int64_t left = 1, right = nbT;
                    for (int k = 0; k < D; ++k) {
                        right /= tnb[k];
                        mode_apply(Bm + offB[k * (LG + 1) + lk[k]], tnq[k], tnb[k], a, b, left, right);
                        left *= tnq[k];
                        double *tmp = a; a = b; b = tmp;
                    }
                    memcpy(Pg + (int64_t)ib0 * nq, a, sizeof(double) * nq);
                }
                /* VPsi (:1578-1588) */
                memset(VPsi, 0, sizeof(double) * (size_t)nq * nb0);
                if (V)
                    for (int ib0 = 0; ib0 < nb0; ++ib0)
                        for (int jb0 = 0; jb0 < nb0; ++jb0) {
                            const double *g = V + g0 + NQ * (ib0 + (int64_t)nb0 * jb0);
                            for (int q = 0; q < nq; ++q) VPsi[(int64_t)ib0 * nq + q] += g[q] * Pg[(int64_t)jb0 * nq + q];
                        }
                for (int ib0 = 0; ib0 < nb0; ++ib0) {
                    double *P = Pg + (int64_t)ib0 * nq;
                    for (int q = 0; q < nq; ++q) P[q] *= sq[g0 + q];                    /* :1595 */
                    for (int j = 0; j < n_act; ++j) {                                    /* :1602-1609 */
                        double *dst = Rj + (int64_t)j * nq;
                        memcpy(dst, P, sizeof(double) * nq);
                        int64_t left = 1, right = nq;
                        for (int k = 0; k < D; ++k) {
                            right /= tnq[k];
                            if (act_mode[j] == k + 1) {
                                mode_apply(D1 + offG[k * (LG + 1) + lk[k]], tnq[k], tnq[k], dst, X, left, right);
                                memcpy(dst, X, sizeof(double) * nq);
                            }
                            left *= tnq[k];
                        }
                    }
                    memset(Op, 0, sizeof(double) * nq);
                    for (int i = 0; i < n_act; ++i) {                                    /* :1613-1630 */
                        memset(Ri, 0, sizeof(double) * nq);
                        for (int j = 0; j < n_act; ++j) {
                            const double *g = GG + g0 + NQ * (j + (int64_t)n_act * i);
                            const double *r = Rj + (int64_t)j * nq;
                            for (int q = 0; q < nq; ++q) Ri[q] += g[q] * r[q];
                        }
                        for (int q = 0; q < nq; ++q) Ri[q] *= Jac[g0 + q];
                        int64_t left = 1, right = nq;
                        for (int k = 0; k < D; ++k) {
                            right /= tnq[k];
                            if (act_mode[i] == k + 1) {
                                mode_apply(D1 + offG[k * (LG + 1) + lk[k]], tnq[k], tnq[k], Ri, X, left, right);
                                memcpy(Ri, X, sizeof(double) * nq);
                            }
                            left *= tnq[k];
                        }
                        for (int q = 0; q < nq; ++q) Op[q] += Ri[q];
                    }
                    for (int q = 0; q < nq; ++q)                                          /* :1634 */
                        Op[q] = -0.5 * Op[q] / (Jac[g0 + q] * sq[g0 + q]) + VPsi[(int64_t)ib0 * nq + q];
                    /* G -> B and weighted scatter */
                    memcpy(X, Op, sizeof(double) * nq);
                    double *a = X, *b = Y;
                    int64_t left = 1, right = nq;
                    for (int k = 0; k < D; ++k) {
                        right /= tnq[k];
                        mode_apply(BTw + offB[k * (LG + 1) + lk[k]], tnb[k], tnq[k], a, b, left, right);
                        left *= tnb[k];
                        double *tmp = a; a = b; b = tmp;
                    }
                    const double w = W[iG];
                    for (int j = 0; j < nbT; ++j) {
                        int32_t m = mp[j];
                        if (m > 0 && m <= nb) {
                            double val = w * a[j];
#pragma omp atomic
                            y[(int64_t)ib0 * nb + m - 1] += val;
                        }
                    }
                }
            }
        }
        free(X); free(Y); free(Pg); free(VPsi); free(Rj); free(Ri); free(Op);
    }
    free(offB); free(offG); free(sum_nq); free(sum_nb);
    return 0;
}

/* ------------------------------------------------------------------------
 * Whole-vector routines of the nested-SG4 entry, one vector at a time, sequential, exactly in the
 * reference's order of operations.
 *   mode 0  RvecB -> RvecG :  tabPackedBasis_TO_SmolyakRepBasis  (...BtoG_GtoB_SG4.f90:1032-1105)
 *                             BSmolyakRep_TO3_GSmolyakRep          (:2307-2383, per term BDP_TO_GDP_OF_SmolyakRep :2385-2496)
 *                             SmolyakRep2_TO_tabR1bis              (:1336-1379)
 *   mode 1  RvecG -> RvecB :  tabR2bis_TO_SmolyakRep1              (:1452-1496)
 *                             GSmolyakRep_TO3_BSmolyakRep          (:2153-2214, per term GDP_TO_BDP_OF_SmolyakRep :2497-2581)
 *                             SmolyakRepBasis_TO_tabPackedBasis    (:951-1028): tabR = 0, terms with |W| < 1e-6 skipped
 *   mode 2  RvecG -> RvecG :  DerivOp_TO3_GSmolyakRep              (:2583-2634, per term DerivOp_TO_RDP_OF_SmolaykRep :2690-2795)
 * der1/der2: 1-based SG4 mode owning each index of tab_der (0 = none).
 * RvecB[ib0*nb + iB], RvecG[ib0*NQ + q] (q over all terms in iG order, first mode fastest inside a term).
 * ---------------------------------------------------------------------- */
int orc_nested(int mode, int D, int nb_SG, int nb0, int64_t nb, int LG,
               const int *tab_l, const double *W, const int *tab_nq, const int *tab_nb, const int32_t *map,
               const int *nq_of, const int *nb_of,
               const double *Bm, const double *BTw, const double *D1, const double *D2,
               int der1, int der2, const double *in, double *out)
{
    const int nT = D * (LG + 1);
    int64_t *offB = (int64_t *)malloc(sizeof(int64_t) * nT);
    int64_t *offG = (int64_t *)malloc(sizeof(int64_t) * nT);
    table_offsets(D, LG, nq_of, nb_of, offB, offG);
    int64_t NQ = 0, maxn = 1;
    for (int iG = 0; iG < nb_SG; ++iG) {
        NQ += tab_nq[iG];
        int64_t m = 1;
        for (int k = 0; k < D; ++k) {
            int l = tab_l[iG * D + k];
            int a = nq_of[k * (LG + 1) + l], b = nb_of[k * (LG + 1) + l];
            m *= (a > b) ? a : b;
        }
        if (m * nb0 > maxn) maxn = m * nb0;
    }
    double *X = (double *)malloc(sizeof(double) * (size_t)maxn), *Y = (double *)malloc(sizeof(double) * (size_t)maxn);
    if (mode == 1) memset(out, 0, sizeof(double) * (size_t)nb * nb0);                /* tabR(:) = ZERO, :978 */
    if (mode == 2 && out != in) memcpy(out, in, sizeof(double) * (size_t)NQ * nb0);
    int64_t off_b = 0, off_q = 0;
    int tnq[ORC_MAXD], tnb[ORC_MAXD];
    for (int iG = 0; iG < nb_SG; ++iG) {
        const int nq = tab_nq[iG], nbT = tab_nb[iG];
        int64_t ob[ORC_MAXD], og[ORC_MAXD];
        for (int k = 0; k < D; ++k) {
            const int i = k * (LG + 1) + tab_l[iG * D + k];
            tnq[k] = nq_of[i]; tnb[k] = nb_of[i]; ob[k] = offB[i]; og[k] = offG[i];
        }
        const int32_t *mp = map + off_b;
        double *cur = X, *oth = Y;
        if (mode == 0) {
            for (int c = 0; c < nb0; ++c)
                for (int j = 0; j < nbT; ++j) cur[c * nbT + j] = (mp[j] > 0 && mp[j] <= nb) ? in[(int64_t)c * nb + mp[j] - 1] : 0.0;
            int64_t left = 1, right = (int64_t)nbT * nb0;
            for (int k = 0; k < D; ++k) {                    /* every mode, also the 1 x 1 ones (:2441-2470) */
                right /= tnb[k];
                mode_apply(Bm + ob[k], tnq[k], tnb[k], cur, oth, left, right);
                double *t = cur; cur = oth; oth = t;
                left *= tnq[k];
            }
            for (int c = 0; c < nb0; ++c)
                for (int q = 0; q < nq; ++q) out[(int64_t)c * NQ + off_q + q] = cur[c * nq + q];
        } else if (mode == 1) {
            if (fabs(W[iG]) >= 1e-6) {                       /* :1004 */
                for (int c = 0; c < nb0; ++c)
                    for (int q = 0; q < nq; ++q) cur[c * nq + q] = in[(int64_t)c * NQ + off_q + q];
                int64_t left = 1, right = (int64_t)nq * nb0;
                for (int k = 0; k < D; ++k) {
                    right /= tnq[k];
                    mode_apply(BTw + ob[k], tnb[k], tnq[k], cur, oth, left, right);
                    double *t = cur; cur = oth; oth = t;
                    left *= tnb[k];
                }
                for (int c = 0; c < nb0; ++c)
                    for (int j = 0; j < nbT; ++j)
                        if (mp[j] > 0 && mp[j] <= nb) out[(int64_t)c * nb + mp[j] - 1] += W[iG] * cur[c * nbT + j];
            }
        } else {
            for (int c = 0; c < nb0; ++c)
                for (int q = 0; q < nq; ++q) cur[c * nq + q] = out[(int64_t)c * NQ + off_q + q];
            int64_t left = 1;
            for (int k = 0; k < D; ++k) {                    /* loop over the modes as the reference does (:2745-2785) */
                const int hit1 = (der1 == k + 1), hit2 = (der2 == k + 1);
                if (hit1 || hit2) {
                    const double *M = (hit1 && hit2) ? D2 + og[k] : D1 + og[k];
                    const int64_t right = ((int64_t)nq / (left * tnq[k])) * nb0;
                    mode_apply(M, tnq[k], tnq[k], cur, oth, left, right);
                    double *t = cur; cur = oth; oth = t;
                }
                left *= tnq[k];
            }
            for (int c = 0; c < nb0; ++c)
                for (int q = 0; q < nq; ++q) out[(int64_t)c * NQ + off_q + q] = cur[c * nq + q];
        }
        off_b += nbT; off_q += nq;
    }
    free(X); free(Y); free(offB); free(offG);
    return 0;
}

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
