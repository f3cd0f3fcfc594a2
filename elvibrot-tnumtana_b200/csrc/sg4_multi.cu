// sg4_multi.cu -- one process, several GPUs: evr_sg4_set_devices(n) makes every plan created on the default device span the
// first n devices of the node.  The Smolyak terms of the plan's range are split in n contiguous sub-ranges with equal
// numbers of grid points (the reference's MPI scheme 1 decomposition, ini_iGs_MPI / auto_iGs_MPI,
// sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4_MPI.f90:639-669, sub_OpPsi_SG4_MPI.f90:2689-2815), every device applies its
// terms to the (replicated) packed psi, and the partial results are summed over NVLink peer memory -- the reference's
// MPI_Reduce_sum_Bcast (Action_MPI_S1, sub_OpPsi_SG4_MPI.f90:535-560) without MPI, reachable from a plain Fortran / C caller.
//
// Host entry (evr_sg4_apply) moves data slice-wise so that the n PCIe links work in parallel:
//   H2D: device d receives slice d of psi;  all-gather of the slices over NVLink (one kernel per device pulls its peers'
//   slices);  term kernels;  reduce-scatter over NVLink (device d sums slice d of all partial results, fixed order 0..n-1);
//   D2H: device d returns slice d of H psi.
// One host thread per device issues that device's work (OpenMP), cross-device ordering by CUDA events.
#include "sg4_plan.h"
#include "../../include/evr_sg4_comm.h"

#include <omp.h>

#include <algorithm>
#include <string>
#include <vector>

using evr::fail;

namespace {
int g_ndev = 1;

#define MCUDA(expr, rc, msg)                                                                   \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess && (rc) == 0) { (rc) = 1; (msg) = std::string(#expr) + ": " + cudaGetErrorString(e__); } \
    } while (0)

// run body(d) on one host thread per device; collects the first error (the per-thread error strings of the
// single-device layer are thread-local, so they are copied out here)
template <class F>
int for_each_device(evr_sg4_plan *p, F body)
{
    const int n = (int)p->sub.size();
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
#pragma omp parallel for num_threads(n) schedule(static, 1)
    for (int d = 0; d < n; ++d) {
        if (cudaSetDevice(d) != cudaSuccess) { rc[d] = 1; msg[d] = "cudaSetDevice failed"; continue; }   // sub-plan d lives on device d
        rc[d] = body(d, msg[d]);
        if (rc[d] && msg[d].empty()) msg[d] = evr_sg4_last_error();
    }
    for (int d = 0; d < n; ++d) if (rc[d]) return fail("device " + std::to_string(d) + ": " + msg[d]);
    return 0;
}
} // namespace

extern "C" int evr_sg4_set_devices(int ndev)
{
    int have = 0;
    if (cudaGetDeviceCount(&have) != cudaSuccess || have < 1)
        return fail("evr_sg4_set_devices: no CUDA device available (this library has no CPU fallback)");
    if (ndev < 1 || ndev > have || ndev > EVR_SG4_MAX_PEERS)
        return fail("evr_sg4_set_devices: ndev must be in [1, " + std::to_string(std::min(have, EVR_SG4_MAX_PEERS)) + "]");
    int cur = 0;
    cudaGetDevice(&cur);
    for (int a = 0; a < ndev && ndev > 1; ++a) {
        if (cudaSetDevice(a) != cudaSuccess) return fail("evr_sg4_set_devices: cudaSetDevice failed");
        for (int b = 0; b < ndev; ++b) {
            if (a == b) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, a, b) != cudaSuccess || !can) {
                cudaSetDevice(cur);
                return fail("evr_sg4_set_devices: device " + std::to_string(a) + " cannot access device " + std::to_string(b) + " (no NVLink / P2P path)");
            }
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (e != cudaSuccess) { cudaSetDevice(cur); return fail(std::string("evr_sg4_set_devices: cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e)); }
        }
    }
    cudaSetDevice(cur);
    g_ndev = ndev;
    return 0;
}
extern "C" int evr_sg4_get_devices(void) { return g_ndev; }

// page-locks a caller-owned host buffer so that the slice-wise copies of evr_sg4_apply are truly asynchronous (the shim
// registers its packed x / y buffers once); a registered range must be unregistered before it is freed
extern "C" int evr_sg4_host_register(void *ptr, int64_t bytes)
{
    if (!ptr || bytes <= 0) return fail("evr_sg4_host_register: bad arguments");
    cudaError_t e = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return fail(std::string("evr_sg4_host_register: ") + cudaGetErrorString(e));
    return 0;
}
extern "C" int evr_sg4_host_unregister(void *ptr)
{
    if (!ptr) return 0;
    cudaError_t e = cudaHostUnregister(ptr);
    if (e == cudaErrorHostMemoryNotRegistered) { cudaGetLastError(); return 0; }
    if (e != cudaSuccess) return fail(std::string("evr_sg4_host_unregister: ") + cudaGetErrorString(e));
    return 0;
}

int evr::multi_devices() { return g_ndev; }

int evr::multi_create(evr_sg4_plan **out, int D, int nb_SG, int nb0, int64_t nb, int LG,
                      const int32_t *tab_l, const double *WeightSG, const int32_t *tab_nq, const int32_t *tab_nb,
                      const int32_t *tab_iB, const int32_t *nq_of, const int32_t *nb_of,
                      const double *B, const double *BTw, const double *D1, const double *D2, int iG_begin, int iG_end)
{
    if (!out || !tab_nq) return fail("evr_sg4_plan_create: null argument");
    if (iG_begin < 0 || iG_end > nb_SG || iG_begin > iG_end) return fail("evr_sg4_plan_create: bad term range");
    const int n = g_ndev;
    auto *p = new evr_sg4_plan();
    p->device = 0; p->D = D; p->nb_SG = nb_SG; p->nb0 = nb0; p->nb = nb; p->LG = LG;
    p->iG_begin = iG_begin; p->iG_end = iG_end; p->n_terms = iG_end - iG_begin;
    p->sub.assign(n, nullptr);
    p->ev_in.assign(n, nullptr); p->ev_done.assign(n, nullptr);
    // contiguous sub-ranges with (nearly) equal numbers of grid points
    std::vector<int> lo(n), hi(n);
    for (int d = 0; d < n; ++d) {
        int b = 0, e = 0;
        if (evr_sg4_balanced_iGs(iG_end - iG_begin, tab_nq + iG_begin, n, d, &b, &e)) { delete p; return 1; }
        lo[d] = iG_begin + b; hi[d] = iG_begin + e;
    }
    const int rc = for_each_device(p, [&](int d, std::string &) {
        return evr::plan_create_single(&p->sub[d], d, D, nb_SG, nb0, nb, LG, tab_l, WeightSG, tab_nq, tab_nb, tab_iB, nq_of, nb_of,
                                       B, BTw, D1, D2, lo[d], hi[d]);
    });
    if (!rc) {
        int cur = 0;
        cudaGetDevice(&cur);
        for (int d = 0; d < n; ++d) {
            cudaSetDevice(d);
            if (cudaEventCreateWithFlags(&p->ev_in[d], cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&p->ev_done[d], cudaEventDisableTiming) != cudaSuccess) {
                cudaSetDevice(cur); evr::multi_destroy(p); delete p;
                return fail("evr_sg4_plan_create: cudaEventCreate failed");
            }
        }
        cudaSetDevice(cur);
    }
    if (rc) { const std::string m = evr_sg4_last_error(); evr::multi_destroy(p); delete p; return fail(m); }
    *out = p;
    return 0;
}

int evr::multi_set_op(evr_sg4_plan *p, int type_Op, int nb_Term, const int32_t *term_mode, const uint8_t *grid_zero,
                      const uint8_t *grid_cte, const double *Mat_cte, const double *const *grids)
{
    const int rc = for_each_device(p, [&](int d, std::string &) {
        return evr_sg4_plan_set_op(p->sub[d], type_Op, nb_Term, term_mode, grid_zero, grid_cte, Mat_cte, grids);
    });
    if (!rc) p->op_set = true;
    return rc;
}

int evr::multi_set_op10(evr_sg4_plan *p, int n_act, const int32_t *act_mode, const double *V, const double *GG,
                        const double *Jac, const double *sq)
{
    const int rc = for_each_device(p, [&](int d, std::string &) {
        return evr_sg4_plan_set_op10(p->sub[d], n_act, act_mode, V, GG, Jac, sq);
    });
    if (!rc) p->op_set = true;
    return rc;
}

int evr::multi_apply_host(evr_sg4_plan *p, int npsi, const double *psi, double *Hpsi)
{
    if (!p->op_set) return fail("evr_sg4_apply: operator not set (call evr_sg4_plan_set_op)");
    const int n = (int)p->sub.size();
    const int64_t len = (int64_t)npsi * p->nb * p->nb0;
    {
        const uintptr_t x = reinterpret_cast<uintptr_t>(psi), y = reinterpret_cast<uintptr_t>(Hpsi), bytes = (uintptr_t)len * 8;
        if (x < y + bytes && y < x + bytes) return fail("evr_sg4_apply: psi and Hpsi overlap (the action is not in-place)");
    }
    std::vector<int> rc(n, 0);
    std::vector<std::string> msg(n);
    std::vector<const void *> psi_ptrs(n), out_ptrs(n);
#pragma omp parallel num_threads(n)
    {
        const int d = omp_get_thread_num();
        evr_sg4_plan *s = p->sub[d];
        int &r = rc[d];
        std::string &m = msg[d];
        MCUDA(cudaSetDevice(s->device), r, m);
        if (!r && evr::plan_ensure_staging(s, len)) { r = 1; m = evr_sg4_last_error(); }
        psi_ptrs[d] = s->d_psi; out_ptrs[d] = s->d_Hpsi;
        int64_t lo = 0, hi = 0;
        evr_sg4_slice_bounds(len, n, d, &lo, &hi);
        // (A) this device's slice of psi, host -> device
        if (!r && hi > lo) MCUDA(cudaMemcpyAsync(s->d_psi + lo, psi + lo, (size_t)(hi - lo) * 8, cudaMemcpyHostToDevice, s->stream), r, m);
        if (!r) MCUDA(cudaEventRecord(p->ev_in[d], s->stream), r, m);
#pragma omp barrier
        // (B) all-gather of the other slices over NVLink, then this device's terms
        for (int e = 0; e < n && !r; ++e) if (e != d) MCUDA(cudaStreamWaitEvent(s->stream, p->ev_in[e], 0), r, m);
        if (!r && evr_sg4_allgather_slices(psi_ptrs.data(), n, d, len, s->stream)) { r = 1; m = evr_sg4_last_error(); }
        if (!r && evr::plan_launch(s, npsi, s->d_psi, s->d_Hpsi, s->stream)) { r = 1; m = evr_sg4_last_error(); }
        if (!r) MCUDA(cudaEventRecord(p->ev_done[d], s->stream), r, m);
#pragma omp barrier
        // (C) reduce-scatter: slice d of all partial results is summed here (fixed order), then returned to the host
        for (int e = 0; e < n && !r; ++e) if (e != d) MCUDA(cudaStreamWaitEvent(s->stream, p->ev_done[e], 0), r, m);
        if (!r && evr_sg4_reduce_slice(out_ptrs.data(), n, d, len, s->stream)) { r = 1; m = evr_sg4_last_error(); }
        if (!r && hi > lo) MCUDA(cudaMemcpyAsync(Hpsi + lo, s->d_Hpsi + lo, (size_t)(hi - lo) * 8, cudaMemcpyDeviceToHost, s->stream), r, m);
#pragma omp barrier
        // (D) nobody may start the next call while a peer still reads this device's buffers
        cudaError_t e2 = cudaStreamSynchronize(s->stream);
        if (e2 != cudaSuccess && !r) { r = 1; m = std::string("cudaStreamSynchronize: ") + cudaGetErrorString(e2); }
    }
    p->multi_launches += 2 * n;
    for (int d = 0; d < n; ++d) if (rc[d]) return fail("device " + std::to_string(p->sub[d]->device) + ": " + msg[d]);
    return 0;
}

// device-resident caller: psi / Hpsi live on the first device, `st` is a stream of that device.  The other devices pull psi
// over NVLink, all apply their terms, and the first device sums the partial results into Hpsi (and applies the optional
// sub_scaledOpPsi epilogue).
int evr::multi_apply_device(evr_sg4_plan *p, int npsi, const double *d_psi, double *d_Hpsi, cudaStream_t st, bool scaled, double E0, double Esc)
{
    if (!p->op_set) return fail("evr_sg4_apply_device: operator not set (call evr_sg4_plan_set_op)");
    const int n = (int)p->sub.size();
    const int64_t len = (int64_t)npsi * p->nb * p->nb0;
    int cur = 0;
    cudaGetDevice(&cur);
    int r = 0;
    std::string m;
    evr_sg4_plan *s0 = p->sub[0];
    MCUDA(cudaSetDevice(s0->device), r, m);
    if (!r && evr::plan_ensure_staging(s0, len)) { r = 1; m = evr_sg4_last_error(); }
    if (!r) MCUDA(cudaEventRecord(p->ev_in[0], st), r, m);                  // psi is ready on the caller's stream
    std::vector<const void *> out_ptrs(n);
    out_ptrs[0] = s0->d_Hpsi;
    for (int d = 1; d < n && !r; ++d) {
        evr_sg4_plan *s = p->sub[d];
        MCUDA(cudaSetDevice(s->device), r, m);
        if (!r && evr::plan_ensure_staging(s, len)) { r = 1; m = evr_sg4_last_error(); }
        out_ptrs[d] = s->d_Hpsi;
        if (!r) MCUDA(cudaStreamWaitEvent(s->stream, p->ev_in[0], 0), r, m);
        if (!r) MCUDA(cudaMemcpyPeerAsync(s->d_psi, s->device, d_psi, s0->device, (size_t)len * 8, s->stream), r, m);
        if (!r && evr::plan_launch(s, npsi, s->d_psi, s->d_Hpsi, s->stream)) { r = 1; m = evr_sg4_last_error(); }
        if (!r) MCUDA(cudaEventRecord(p->ev_done[d], s->stream), r, m);
    }
    if (!r) MCUDA(cudaSetDevice(s0->device), r, m);
    if (!r && evr::plan_launch(s0, npsi, d_psi, s0->d_Hpsi, st)) { r = 1; m = evr_sg4_last_error(); }
    for (int d = 1; d < n && !r; ++d) MCUDA(cudaStreamWaitEvent(st, p->ev_done[d], 0), r, m);
    if (!r && evr_sg4_reduce_to(out_ptrs.data(), n, len, d_Hpsi, st)) { r = 1; m = evr_sg4_last_error(); }
    if (!r && scaled && evr::scale_launch(len, E0, Esc, d_psi, d_Hpsi, st)) { r = 1; m = evr_sg4_last_error(); }
    // the next call may overwrite the peers' psi copies only after this one's kernels are done: the peers' streams
    // order that by themselves (same stream); the caller's stream orders the reads of the partial results
    if (!r) MCUDA(cudaEventRecord(p->ev_done[0], st), r, m);
    for (int d = 1; d < n && !r; ++d) {
        MCUDA(cudaSetDevice(p->sub[d]->device), r, m);
        MCUDA(cudaStreamWaitEvent(p->sub[d]->stream, p->ev_done[0], 0), r, m);
    }
    cudaSetDevice(cur);
    p->multi_launches += 1 + (scaled ? 1 : 0);
    return r ? fail(m) : 0;
}

int64_t evr::multi_info(const evr_sg4_plan *p, int what)
{
    int64_t sum = 0;
    switch (what) {
    case EVR_INFO_LAUNCHES:
        for (auto *s : p->sub) sum += evr_sg4_plan_info(s, what);
        return sum + p->multi_launches;
    case EVR_INFO_NQ_LOCAL: case EVR_INFO_S_LOCAL: case EVR_INFO_FLOPS_NPSI1: case EVR_INFO_GENERIC_TERMS:
        for (auto *s : p->sub) sum += evr_sg4_plan_info(s, what);
        return sum;
    case EVR_INFO_ALG_BYTES_NPSI1: case EVR_INFO_ALG_BYTES_PER_RHS_EXTRA: {
        // the zero-fill / read-out term of the packed vector is counted once, like on one device
        const int64_t vec = p->nb * p->nb0 * 8 * 2;
        for (auto *s : p->sub) sum += evr_sg4_plan_info(s, what) - vec;
        return sum + vec;
    }
    case EVR_INFO_DEVICES: return (int64_t)p->sub.size();
    default: return p->sub.empty() ? -1 : evr_sg4_plan_info(p->sub[0], what);
    }
}

int evr::multi_destroy(evr_sg4_plan *p)
{
    int cur = 0;
    cudaGetDevice(&cur);
    for (size_t d = 0; d < p->sub.size(); ++d) {
        if (p->sub[d]) { cudaSetDevice(p->sub[d]->device); cudaDeviceSynchronize(); }
    }
    for (size_t d = 0; d < p->sub.size(); ++d) {
        if (p->sub[d]) evr_sg4_plan_destroy(&p->sub[d]);
        if (d < p->ev_in.size() && p->ev_in[d]) cudaEventDestroy(p->ev_in[d]);
        if (d < p->ev_done.size() && p->ev_done[d]) cudaEventDestroy(p->ev_done[d]);
    }
    p->sub.clear();
    cudaSetDevice(cur);
    return 0;
}
