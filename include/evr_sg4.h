/*
 * evr_sg4.h -- C-ABI of the B200-native H|psi> (OpPsi) action on the Smolyak
 * type-4 sparse grid of ElVibRot-TnumTana.
 *
 * This is the drop-in boundary: the body of the reference's
 *     SUBROUTINE sub_TabOpPsi_FOR_SGtype4(Psi,OpPsi,para_Op)
 *         Source_ElVibRot/sub_Operator/sub_OpPsi_SG4.f90:678-979
 * (and, in MPI builds, Action_MPI_S1, sub_OpPsi_SG4_MPI.f90:454-571) is replaced
 * by  plan_create (first call) + apply (every call).  The ISO_C_BINDING shim a
 * maintainer adds on the Fortran side is shown in INTEGRATION.md and shipped as
 * elvibrot-tnumtana_b200/fortran/evr_sg4_shim.f90.
 *
 * Conventions (same as the reference's own C bindings, TnumTana_Lib.f90:655-740):
 * flat contiguous arrays with explicit sizes, column-major matrices, table
 * ENTRIES keep their Fortran 1-based values (a mapping entry 0 means "dropped").
 * Every function returns 0 on success, non-zero on error; evr_sg4_last_error()
 * gives the message (the reference STOPs with a message; the shim does the same).
 * Not re-entrant per plan: one caller at a time, like para_Op (intent(inout)).
 * There is NO CPU fallback: if no CUDA device is usable the calls fail.
 */
#ifndef EVR_SG4_H
#define EVR_SG4_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct evr_sg4_plan   evr_sg4_plan;
typedef struct evr_sg4_tables evr_sg4_tables;

int         evr_sg4_version(void);
const char *evr_sg4_last_error(void);

/* ---------------------------------------------------------------------------
 * SG4 index / term tables (host, integer, bit-exact with the reference).
 * Replaces the table part of RecSparseGrid_ForDP_type4
 *   (Source_ElVibRot/sub_Basis/sub_quadra_SparseBasis.f90:1130-1362):
 *   nDindB (type 5, Source_Lib/sub_nDindex/sub_module_nDindex.f90:971-1082),
 *   nDind_SmolyakRep (type -5, :1083-1185), WeightSG
 *   (sub_module_param_SGType2.f90:784-806), tab_nq/nb_OF_SRep (+sums,
 *   sub_quadra_SparseBasis.f90:1315-1345) and tab_iB_OF_SRep_TO_iB
 *   (Set_tables_FOR_SmolyakRepBasis_TO_tabPackedBasis,
 *   sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4.f90:625-949).
 * Inputs: nq_of/nb_of[k*(LG+1)+L] = nq_k(L), nb_k(L) of tab_basisPrimSG(L,k).
 * ------------------------------------------------------------------------- */
int evr_sg4_tables_build(evr_sg4_tables **out, int D, int LB, int LG,
                         const int32_t *nq_of, const int32_t *nb_of);
int evr_sg4_tables_destroy(evr_sg4_tables **t);

enum {                           /* 'what' for evr_sg4_tables_size / _get             */
    EVR_TAB_NB_SG   = 0,         /* scalar: number of Smolyak terms                   */
    EVR_TAB_NB      = 1,         /* scalar: packed basis size (nDindB%Max_nDI)        */
    EVR_TAB_S       = 2,         /* scalar: sum_iG nb(iG)  (Max_Srep)                 */
    EVR_TAB_NQ      = 3,         /* scalar: sum_iG nq(iG)                             */
    EVR_TAB_COUNT0  = 4,         /* scalar: number of zero entries of the map         */
    EVR_TAB_LMIN    = 5,         /* scalar: Lmin = max(0, LG-D+1)                     */
    EVR_TAB_TAB_L   = 10,        /* int32 [nb_SG][D]   Tab_nDval(:,iG) of the terms   */
    EVR_TAB_WEIGHT  = 11,        /* double[nb_SG]      WeightSG                       */
    EVR_TAB_TAB_NQ  = 12,        /* int32 [nb_SG]      tab_nq_OF_SRep                 */
    EVR_TAB_TAB_NB  = 13,        /* int32 [nb_SG]      tab_nb_OF_SRep                 */
    EVR_TAB_SUM_NQ  = 14,        /* int64 [nb_SG]      tab_Sum_nq_OF_SRep (inclusive) */
    EVR_TAB_SUM_NB  = 15,        /* int64 [nb_SG]      tab_Sum_nb_OF_SRep (inclusive) */
    EVR_TAB_PACKEDB = 16,        /* int32 [nb][D]      nDindB%Tab_nDval(:,iB)         */
    EVR_TAB_MAP     = 17         /* int32 [S]          tab_iB_OF_SRep_TO_iB           */
};
int64_t evr_sg4_tables_size(const evr_sg4_tables *t, int what);   /* scalar value or element count */
int     evr_sg4_tables_get(const evr_sg4_tables *t, int what, void *dst); /* copy an array out */

/* Term range of one rank: ini_iGs_MPI,
 * sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4_MPI.f90:639-669 (0-based, end exclusive). */
int evr_sg4_ini_iGs(int nb_SG, int np, int rank, int *iG_begin, int *iG_end);

/* Same contiguous decomposition but balanced by work instead of by term count: the ranges hold
 * (nearly) equal sums of cost[iG] (e.g. tab_nq_OF_SRep), the way the reference balances its OpenMP
 * thread ranges by grid points (Set_nDval_init_FOR_SG4 version 1, sub_module_param_SGType2.f90:538-650)
 * and its MPI ranges by timing (auto_iGs_MPI, sub_OpPsi_SG4_MPI.f90:2689-2815). */
int evr_sg4_balanced_iGs(int nb_SG, const int32_t *cost, int np, int rank, int *iG_begin, int *iG_end);

/* ---------------------------------------------------------------------------
 * Several GPUs from ONE process (no MPI): after evr_sg4_set_devices(n), every plan created with device < 0 spans
 * devices 0..n-1 of the node.  The plan's term range is split in n contiguous sub-ranges of equal grid points (the
 * decomposition of the reference's MPI scheme 1, Action_MPI_S1, sub_Operator/sub_OpPsi_SG4_MPI.f90:454-571, with
 * ini_iGs_MPI / auto_iGs_MPI ranges) and the partial results are summed over NVLink peer memory instead of
 * MPI_Reduce_sum_Bcast (:535-560).  evr_sg4_apply then copies psi slice-wise (device d: slice d, all-gathered over NVLink)
 * and returns H psi slice-wise, so the n PCIe links work in parallel; evr_sg4_apply_device[_scaled] expects psi / Hpsi on
 * device 0.  Needs peer access between all pairs of the n devices; fails otherwise.  n = 1 restores single-device plans.
 * evr_sg4_host_register page-locks a caller-owned host buffer (e.g. the shim's packed psi / Hpsi arrays) so that those
 * copies are asynchronous; unregister before freeing it.
 * ------------------------------------------------------------------------- */
int evr_sg4_set_devices(int ndev);
int evr_sg4_get_devices(void);
int evr_sg4_host_register(void *ptr, int64_t bytes);
int evr_sg4_host_unregister(void *ptr);

/* ---------------------------------------------------------------------------
 * Plan = device-resident copy of everything sub_TabOpPsi_FOR_SGtype4 reads from
 * para_Op%BasisnD (param_SGType2, WeightSG, tab_basisPrimSG), for the terms
 * [iG_begin, iG_end) of this process (all terms: 0, nb_SG).
 *   tab_l[iG*D+k]           nDind_SmolyakRep%Tab_nDval(k,iG)
 *   tab_iB[S]               tab_iB_OF_SRep_TO_iB (full table; the slice is taken here)
 *   nq_of/nb_of[k*(LG+1)+L]
 *   B, BTw, D1, D2          concatenation over k = 0..D-1 (outer), L = 0..LG (inner) of
 *                           dnRGB%d0(nq,nb), dnRBGwrho%d0(nb,nq), dnRGG%d1(nq,nq,1),
 *                           dnRGG%d2(nq,nq,1,1), each column-major.
 * device < 0 : use the current CUDA device.
 * ------------------------------------------------------------------------- */
int evr_sg4_plan_create(evr_sg4_plan **plan, int device,
                        int D, int nb_SG, int nb0, int64_t nb, int LG,
                        const int32_t *tab_l, const double *WeightSG,
                        const int32_t *tab_nq_OF_SRep, const int32_t *tab_nb_OF_SRep,
                        const int32_t *tab_iB,
                        const int32_t *nq_of, const int32_t *nb_of,
                        const double *B, const double *BTw, const double *D1, const double *D2,
                        int iG_begin, int iG_end);

/* Same, for a caller that only holds a slice of the mapping table: with MPI scheme 1 the reference allocates
 * tab_iB_OF_SRep_TO_iB(bounds_MPI(1,id):bounds_MPI(2,id)) on rank id (Mapping_table_allocate_MPI,
 * sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4_MPI.f90:62-77).  tab_iB[0] is global entry tab_iB_first (0-based) of the
 * full table, tab_iB_len entries are valid; they must cover the entries of the terms [iG_begin, iG_end). */
int evr_sg4_plan_create_ex(evr_sg4_plan **plan, int device,
                           int D, int nb_SG, int nb0, int64_t nb, int LG,
                           const int32_t *tab_l, const double *WeightSG,
                           const int32_t *tab_nq_OF_SRep, const int32_t *tab_nb_OF_SRep,
                           const int32_t *tab_iB, int64_t tab_iB_first, int64_t tab_iB_len,
                           const int32_t *nq_of, const int32_t *nb_of,
                           const double *B, const double *BTw, const double *D1, const double *D2,
                           int iG_begin, int iG_end);
int evr_sg4_device_count(void);      /* CUDA devices visible to this process (0 if none): rank -> device maps of MPI callers */

/* Operator description = para_Op%{type_Op, nb_Term, derive_termQdyn, OpGrid(:)}
 * (sub_OpPsi_SG4.f90:1447-1546; term numbering Init_TypeOp,
 * Source_PrimOperator/sub_module_SimpleOp.f90:256-375).
 *   type_Op   0 (scalar operator, one term) or 1 (H = sum_iterm F_iterm(Q) d^(i,j)).
 *   term_mode[2*iterm+{0,1}]  1-based SG4 mode owning each index of
 *             derive_termQdyn(:,iterm) through Tabder_Qdyn_TO_Qbasis; 0 = none.
 *   grid_zero/grid_cte[iterm] OpGrid(iterm)%grid_zero / %grid_cte.
 *   Mat_cte[iterm*nb0*nb0 + j + nb0*i] = OpGrid(iterm)%Mat_cte(j,i).
 *   grids[iterm]  NULL for zero/cte terms, else OpGrid(iterm)%Grid(1:NQ,1:nb0,1:nb0)
 *             (column-major, the whole Smolyak grid; term offset tab_Sum_nq-nq).
 * The grids are copied to the device once (they are immutable after the first
 * H|psi>, Save_MemGrid_done). */
int evr_sg4_plan_set_op(evr_sg4_plan *plan, int type_Op, int nb_Term,
                        const int32_t *term_mode,
                        const uint8_t *grid_zero, const uint8_t *grid_cte,
                        const double *Mat_cte, const double *const *grids);

/* Operator-grid construction on the device for closed-form models: the potential on the Smolyak grid of
 * the terms [iG_begin, iG_end), in the layout of OpGrid(iterm00)%Grid(:) (terms in iG order, first mode fastest), i.e.
 * what the reference computes point by point during its first H|psi> (Rec_Qact_SG4_with_Tab_iq + get_d0MatOp_AT_Qact,
 * sub_OpPsi_SG4.f90:2982-3006).  x_tab = concatenation over k (outer), L (inner) of the nq_k(L) grid points of
 * tab_basisPrimSG(L,k).  model 1: Henon-Heiles, params[0] = lambda (sub_system_HenonHeiles.f:40-47);
 * model 2: 1/2 sum_i params[i] Q_i^2.  V_host receives sum_iG nq(iG) doubles; pass it to evr_sg4_plan_set_op. */
int evr_sg4_model_grid(int model, int D, int nb_SG, int LG, const int32_t *tab_l, const int32_t *nq_of,
                       const double *x_tab, int nparam, const double *params, int iG_begin, int iG_end,
                       double *V_host);

/* type_Op = 10 with the metric tensor cached per grid point (next row 8f-1 of SURVEY.md):
 *   H psi = -1/2 (Jac sq)^-1 sum_i d_i [ Jac sum_j GG(:,j,i) d_j (sq psi) ] + V psi
 * (sub_OpPsi_SG4.f90:1548-1650).  The reference recomputes GG/Jac/rho with Tnum at every grid point of every
 * call (get_OpGrid_type10_OF_ONEDP_FOR_SG4, :2717-2728); here they are plan data, uploaded once.
 *   act_mode[j]  1-based SG4 mode owning active coordinate j (liste_QactTOQdyn + Tabder_Qdyn_TO_Qbasis)
 *   V            Grid(1:NQ,1:nb0,1:nb0) of the (0,0) term or NULL
 *   GG           GGiq(1:NQ,1:n_act,1:n_act), Jac(1:NQ), sqRhoOVERJac(1:NQ), whole Smolyak grid, column-major.
 * The device keeps only the upper triangle of GG (n_act(n_act+1)/2 + 2 doubles per point) when GG(q,j,i) == GG(q,i,j)
 * holds exactly on the plan's grid range, the full tensor otherwise. */
int evr_sg4_plan_set_op10(evr_sg4_plan *plan, int n_act, const int32_t *act_mode,
                          const double *V, const double *GG, const double *Jac, const double *sqRhoOVERJac);

/* H|psi> for npsi real right-hand sides (complex psi = 2 real RHS, sub_OpPsi.f90:392-407).
 * psi/Hpsi[ipsi*nb*nb0 + ib0*nb + iB]  (= Psi(ipsi)%RvecB).  Hpsi is overwritten
 * (the reference zeroes it, sub_OpPsi_SG4.f90:765); with a term sub-range it holds
 * this rank's partial sum (to be summed over ranks: MPI_Reduce_sum_Bcast / NCCL).
 * evr_sg4_apply on a block (npsi >= 2) of long vectors (>= 1 MB each) moves vector v+1 to the device and vector v-1 back
 * while vector v is in the kernels; page-lock the buffers (evr_sg4_host_register) for the copies to overlap.
 * There is no limit on the size of a Smolyak term: terms beyond the shared memory of an SM work in global buffers. */
int evr_sg4_apply(evr_sg4_plan *plan, int npsi, const double *psi, double *Hpsi);          /* host buffers   */
int evr_sg4_apply_device(evr_sg4_plan *plan, int npsi, const double *d_psi, double *d_Hpsi,
                         void *cuda_stream);

/* Device-resident H|psi> followed by the Chebyshev / SIL scaling of sub_scaledOpPsi
 * (Source_ElVibRot/sub_Operator/sub_OpPsi.f90:2823-2866):   Hpsi <- (H psi - E0 psi) / Esc
 * in one call: the propagators apply this after every H|psi> (sub_module_propa_march.f90:4294-4345), so with the
 * vectors resident on the device no host round trip remains in the recursion.  On the fast path
 * with the block-ordered internal vector the scaling is fused into the kernel that writes the result back in
 * the caller's order; otherwise it is one extra element-wise kernel.  Esc must not be 0. */
int evr_sg4_apply_device_scaled(evr_sg4_plan *plan, int npsi, const double *d_psi, double *d_Hpsi,
                                double E0, double Esc, void *cuda_stream);                                               /* device buffers */

/* ---------------------------------------------------------------------------
 * Whole-vector transforms of the SG4 basis (nested-SG4 entry): what the reference's recursive
 * RecRvecB_TO_RVecG / RecRVecG_TO_RvecB / DerivOp_TO_RVecG do for SparseGrid_type = 4 when the SG4 basis is a sub-basis of
 * an outer direct product (sub_Basis/sub_module_basis_BtoG_GtoB.f90:831-847, 252-273, 1394-1416), batched over the outer
 * index.  They only need the plan (no operator).  Layouts:  RvecB[iv*nb*nb0 + ib0*nb + iB]  (packed basis, as psi);
 * RvecG[iv*NQ*nb0 + ib0*NQ + q], q over the Smolyak grid of the plan's term range, terms in iG order, first mode
 * fastest inside a term (SmolyakRep2_TO_tabR1bis, sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4.f90:1336-1379).
 *   BtoG     : tabPackedBasis_TO_SmolyakRepBasis (:1032) + BSmolyakRep_TO[3]_GSmolyakRep (:2216, :2307)
 *   GtoB     : GSmolyakRep_TO[3]_BSmolyakRep (:2101, :2153) + SmolyakRepBasis_TO_tabPackedBasis (:951): weighted sum over the
 *              terms (terms with |WeightSG| < 1e-6 skipped, :1004); RvecB is overwritten
 *   DerivOp_G: DerivOp_TO3_GSmolyakRep (:2583); mode1/mode2 = 1-based SG4 mode owning each index of tab_der
 *              (Tabder_Qdyn_TO_Qbasis), 0 = none: (k,k) second derivative, (k,l) d/dQ_k d/dQ_l, (k,0) first derivative,
 *              (0,0) unchanged.  The _device variant works in place.
 * ------------------------------------------------------------------------- */
int evr_sg4_BtoG(evr_sg4_plan *plan, int nvec, const double *RvecB, double *RvecG);
int evr_sg4_GtoB(evr_sg4_plan *plan, int nvec, const double *RvecG, double *RvecB);
int evr_sg4_DerivOp_G(evr_sg4_plan *plan, int nvec, const double *RvecG_in, double *RvecG_out, int mode1, int mode2);
int evr_sg4_BtoG_device(evr_sg4_plan *plan, int nvec, const double *d_RvecB, double *d_RvecG, void *cuda_stream);
int evr_sg4_GtoB_device(evr_sg4_plan *plan, int nvec, const double *d_RvecG, double *d_RvecB, void *cuda_stream);
int evr_sg4_DerivOp_G_device(evr_sg4_plan *plan, int nvec, double *d_RvecG, int mode1, int mode2, void *cuda_stream);

enum {                           /* 'what' for evr_sg4_plan_info */
    EVR_INFO_LAUNCHES        = 0,   /* kernels launched by this plan so far            */
    EVR_INFO_ALG_BYTES_NPSI1 = 1,   /* algorithmic bytes of one H|psi>, npsi = 1       */
    EVR_INFO_ALG_BYTES_PER_RHS_EXTRA = 2, /* additional bytes per extra RHS            */
    EVR_INFO_NQ_LOCAL        = 3,   /* grid points in this plan's term range           */
    EVR_INFO_S_LOCAL         = 4,   /* term-local basis coefficients in the range      */
    EVR_INFO_SMEM_BYTES      = 5,   /* dynamic shared memory per CTA                   */
    EVR_INFO_GRID_CTAS       = 6,   /* CTAs per launch                                 */
    EVR_INFO_PATH            = 7,   /* 0 = generic term kernel, 1 = constant-KEO fast path */
    EVR_INFO_FLOPS_NPSI1     = 8,   /* algorithmic flops of one H|psi>      */
    EVR_INFO_ISO             = 9,   /* 1 = fast path runs its constant-matrix instantiation (all modes of a size share one 1-D basis) */
    EVR_INFO_DEVICES         = 10,  /* number of devices the plan spans (evr_sg4_set_devices) */
    EVR_INFO_GENERIC_TERMS   = 11   /* Smolyak terms of the plan that run in the generic kernel (all of them when PATH = 0) */
};
int64_t evr_sg4_plan_info(const evr_sg4_plan *plan, int what);
int     evr_sg4_plan_destroy(evr_sg4_plan **plan);

#ifdef __cplusplus
}
#endif
#endif /* EVR_SG4_H */
