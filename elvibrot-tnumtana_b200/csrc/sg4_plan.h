// sg4_plan.h -- the plan object behind the opaque evr_sg4_plan handle (shared by sg4_plan.cu and sg4_multi.cu).
#pragma once
#include "../../include/evr_sg4.h"
#include "sg4_internal.h"
#include "sg4_kernels.cuh"
#include "sg4_fast_types.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

struct evr_sg4_plan {
    int device = 0;
    int D = 0, nb_SG = 0, nb0 = 1, LG = 0;
    int64_t nb = 0;
    int iG_begin = 0, iG_end = 0, n_terms = 0;
    int64_t S_local = 0, NQ_local = 0, NQ_total = 0;
    int64_t grid_start = 0;                 // first grid point of the range in the full Smolyak grid
    int cap = 0;                            // doubles per smem buffer (incl. nb0)
    int sm_count = 0;
    bool op_set = false;
    int type_Op = 1, n_opterms = 0, n_var = 0;
    int64_t launches = 0;
    int64_t flops_npsi1 = 0;
    // host copies needed later
    std::vector<int32_t> h_tab_l, h_nq_of, h_nb_of, h_tab_nq, h_tab_nb;
    std::vector<int> order;                 // work order -> local term index
    std::vector<int32_t> h_map;             // mapping slice of the range (reference order)
    std::vector<int64_t> h_map_off, h_grid_off;   // per local term (reference order)
    std::vector<double> h_B, h_BTw, h_D1, h_D2, h_weight;
    std::vector<int32_t> h_offB, h_offG;
    // fast path (sg4_fast.cuh)
    bool fast = false;
    evr::FastTermDev *d_fterms = nullptr;
    int32_t *d_fmap = nullptr;           // per term: internal packed index (sorted ascending), -1 = dropped
    int32_t *d_gmap = nullptr;           // per term: internal packed index in term-local (internal layout) order
    uint16_t *d_fpos = nullptr;          // per term: term-local position of each sorted entry
    int32_t *d_perm = nullptr;           // internal packed order -> reference packed index (0-based)
    int32_t *d_inv_perm = nullptr;       // the inverse (read-out with coalesced stores; nullptr with EVR_SG4_PERMUTE=0)
    double *d_psi_int = nullptr, *d_Hpsi_int = nullptr;   // packed vectors in the internal (block) order
    int64_t int_cap = 0;
    double *d_fmats = nullptr, *d_fV = nullptr;
    evr::FastPlanDev fpd{};
    std::vector<double> h_cost;             // per local term
    std::vector<int64_t> h_tsize;           // per local term: prod max(nq_k, nb_k)
    int n_classes = 0;
    bool fast_pool_in_smem = false;
    bool fast_block_order = false;
    bool fast_iso = false;                  // constant-matrix instantiation (sg4_iso.cu)
    std::vector<double> iso_blocks;         // its [B|BTw|T] blocks, bound to the __constant__ array before each launch
    int iso_id = 0;
    int n_fitems = 0;                       // work items (batches of same-schedule terms) of the fast path
    evr::FastClassDev fclass[EVR_MAX_FCLASSES];
    size_t fclass_smem[EVR_MAX_FCLASSES] = {0};
    int fclass_ctas[EVR_MAX_FCLASSES] = {0};
    bool fclass_is_iso[EVR_MAX_FCLASSES] = {false};
    bool fclass_v2[EVR_MAX_FCLASSES] = {false};   // second-generation kernel (sg4_fast2.cuh)
    int fclass_flavour[EVR_MAX_FCLASSES] = {0};        // 0 templated, 1 runtime-size, 2 cube tiles (plain) / iso with large tiles, 3 iso
    // device
    evr::TermDev *d_terms = nullptr;
    uint8_t *d_lev = nullptr;
    int32_t *d_map = nullptr, *d_nq_of = nullptr, *d_nb_of = nullptr, *d_offB = nullptr, *d_offG = nullptr;
    double *d_B = nullptr, *d_BTw = nullptr, *d_D1 = nullptr, *d_D2 = nullptr;
    evr::OpTermDev *d_opterms = nullptr;
    double *d_grids = nullptr;
    double *d_psi = nullptr, *d_Hpsi = nullptr;   // staging for the host-buffer entry point
    // type_Op = 10
    bool op10 = false;
    evr::Op10Dev o10{};
    double *d_GG = nullptr, *d_Jac = nullptr, *d_sq = nullptr;
    size_t smem10 = 0;
    int ctas10_max = 0;
    int64_t stage_cap = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side[EVR_MAX_FCLASSES] = {nullptr};   // class kernels overlap their tails
    cudaEvent_t ev_fork = nullptr, ev_join[EVR_MAX_FCLASSES] = {nullptr};
    cudaStream_t s_in = nullptr, s_out = nullptr;      // host-buffer blocks: copy-in / copy-out streams (evr_sg4_apply)
    cudaEvent_t ev_pin = nullptr, ev_pout = nullptr;
    size_t smem_bytes = 0;
    int grid_ctas = 0, gen_ctas_max = 0;
    // generic kernel: one launch per term-size class (CTA of 256/128/64/32 threads)
    int n_gclasses = 0;
    evr::GenClassDev gclass[5];              // [0] may be the class of terms beyond shared memory (global work buffers)
    int gclass_threads[5] = {0}, gclass_occ[5] = {0};
    size_t gclass_smem[5] = {0};
    bool gclass_big[5] = {false};
    // fast-path plans: the terms left to the generic kernel (work-order indices in d_rlist), same class scheme
    int n_rclasses = 0, n_rest_terms = 0;
    evr::GenClassDev rclass[5];
    int rclass_threads[5] = {0}, rclass_occ[5] = {0};
    size_t rclass_smem[5] = {0};
    bool rclass_big[5] = {false};
    double *d_rscratch = nullptr;
    int *d_rlist = nullptr;
    double *d_gscratch = nullptr, *d_nscratch = nullptr;   // global work buffers of the generic / nested kernels (terms beyond shared memory)
    int n_big_terms = 0;                     // such terms come first in work order
    int64_t cap_small = 1;                   // largest term*nb0 among the others
    evr::PlanDev pd{};
    int *d_fcounters = nullptr;              // one work counter per fast-path launch (dynamic item scheduling)
    // deterministic mode (EVR_SG4_DETERMINISTIC=1 when the plan is created): staged scatter + ordered collection
    bool deterministic = false;
    long long *d_det_off = nullptr;          // [nb+1] entry-list offsets per packed element (internal order on the fast path)
    int32_t *d_det_ent = nullptr;            // entry positions, grouped by packed element, ascending
    double *d_stage = nullptr;
    int64_t stage_ld = 0, stage_vecs = 0;
    // CUDA graphs of the per-call launch sequence, keyed by the call's arguments (sg4_plan.cu: launch)
    struct GraphEntry { int npsi; const double *psi; double *Hpsi; bool scaled; double E0, Esc; cudaGraphExec_t exec; int kernels; uint64_t stamp; };
    std::vector<GraphEntry> graphs;
    uint64_t graph_clock = 0;
    // multi-device parent (evr_sg4_set_devices, sg4_multi.cu): one sub-plan per device over a share of the term range;
    // a parent owns no device data of its own
    std::vector<evr_sg4_plan *> sub;
    std::vector<cudaEvent_t> ev_in, ev_done;
    int64_t multi_launches = 0;
};

namespace evr {
// single-device entry points (sg4_plan.cu), used by the multi-device layer
int plan_create_single(evr_sg4_plan **out, int device, int D, int nb_SG, int nb0, int64_t nb, int LG,
                       const int32_t *tab_l, const double *WeightSG, const int32_t *tab_nq, const int32_t *tab_nb,
                       const int32_t *tab_iB, const int32_t *nq_of, const int32_t *nb_of,
                       const double *B, const double *BTw, const double *D1, const double *D2, int iG_begin, int iG_end);
int plan_launch(evr_sg4_plan *p, int npsi, const double *d_psi, double *d_Hpsi, cudaStream_t st);
int plan_ensure_staging(evr_sg4_plan *p, int64_t n);
int scale_launch(long long n, double E0, double Esc, const double *x, double *y, cudaStream_t st);
// multi-device layer (sg4_multi.cu)
int multi_devices();
int multi_create(evr_sg4_plan **out, int D, int nb_SG, int nb0, int64_t nb, int LG,
                 const int32_t *tab_l, const double *WeightSG, const int32_t *tab_nq, const int32_t *tab_nb,
                 const int32_t *tab_iB, const int32_t *nq_of, const int32_t *nb_of,
                 const double *B, const double *BTw, const double *D1, const double *D2, int iG_begin, int iG_end);
int multi_set_op(evr_sg4_plan *p, int type_Op, int nb_Term, const int32_t *term_mode, const uint8_t *grid_zero,
                 const uint8_t *grid_cte, const double *Mat_cte, const double *const *grids);
int multi_set_op10(evr_sg4_plan *p, int n_act, const int32_t *act_mode, const double *V, const double *GG,
                   const double *Jac, const double *sq);
int multi_apply_host(evr_sg4_plan *p, int npsi, const double *psi, double *Hpsi);
int multi_apply_device(evr_sg4_plan *p, int npsi, const double *d_psi, double *d_Hpsi, cudaStream_t st, bool scaled, double E0, double Esc);
int64_t multi_info(const evr_sg4_plan *p, int what);
int multi_destroy(evr_sg4_plan *p);
}
