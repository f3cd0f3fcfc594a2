// sg4_nested.cu -- whole-vector transforms of an SG4 basis between the packed basis representation and the Smolyak
// grid: the entry points the reference uses when an SG4 basis is a sub-basis of an outer direct product
// (HNO3_UT: SG4 8-D inside a Fourier torsion basis, 51 basis / 81 grid slices), where the generic recursive routines call
//   RecRvecB_TO_RVecG  -> tabPackedBasis_TO_SmolyakRepBasis + BSmolyakRep_TO[3]_GSmolyakRep + SmolyakRep2_TO_tabR1bis
//                         (sub_Basis/sub_module_basis_BtoG_GtoB.f90:831-847; sub_Basis_SG4/...SG4.f90:1032-1105, 2216-2383, 1336-1379)
//   RecRVecG_TO_RvecB  -> tabR2bis_TO_SmolyakRep1 + GSmolyakRep_TO[3]_BSmolyakRep + SmolyakRepBasis_TO_tabPackedBasis
//                         (...BtoG_GtoB.f90:252-273; ...SG4.f90:1452-1496, 2101-2214, 951-1028)
//   DerivOp_TO_RVecG   -> tabR2bis_TO_SmolyakRep1 + DerivOp_TO3_GSmolyakRep + SmolyakRep2_TO_tabR1bis
//                         (...BtoG_GtoB.f90:1394-1416; ...SG4.f90:2583-2634)
// once per outer index.  Here each of the three is ONE launch over all (Smolyak term, vector) pairs of a batch of
// vectors -- the outer index is the batch.  The grid vector keeps the reference layout RVecG((ib0-1)*NQ + q), q running
// over the terms in iG order (first mode fastest inside a term); no intermediate SmolyakRep containers exist.
#include "sg4_plan.h"

#include <algorithm>
#include <cstdlib>
#include <string>

using evr::fail;

namespace evr {

enum { NESTED_BTOG = 0, NESTED_GTOB = 1, NESTED_DERIV = 2 };

// dynamic smem: bufA[cap] | bufB[cap] | ints: nq_of, nb_of, offB, offG (4*D*(LG+1)) | per-term ints 5*D | magic 3*(D+1)
// BIG = true: terms [term_begin, term_begin + n_terms) do not fit in shared memory; the two buffers of a CTA live in its
// slice of `scratch` (global memory) and the index divisions are exact for any size (sg4_kernels.cuh: mdivT)
template <bool BIG>
static __global__ void __launch_bounds__(256, 2)
sg4_nested_kernel(const PlanDev P, const int mode, const int nvec, const double *__restrict__ in, double *__restrict__ out,
                  const int der1, const int der2, const int term_begin, const int n_terms, const int cap, double *scratch)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *bufA = BIG ? scratch + (size_t)blockIdx.x * 2 * cap : reinterpret_cast<double *>(smem_raw);
    double *bufB = bufA + cap;
    int *s_nq_of = BIG ? reinterpret_cast<int *>(smem_raw) : reinterpret_cast<int *>(bufB + cap);
    const int nT = P.D * (P.LG + 1);
    int *s_nb_of = s_nq_of + nT;
    int *s_offB  = s_nb_of + nT;
    int *s_offG  = s_offB + nT;
    int *s_tnq   = s_offG + nT;
    int *s_tnb   = s_tnq + P.D;
    int *s_oB    = s_tnb + P.D;
    int *s_oG    = s_oB + P.D;
    int *s_str   = s_oG + P.D;
    unsigned *s_mgq = reinterpret_cast<unsigned *>(s_str + P.D);
    unsigned *s_mgb = s_mgq + P.D + 1;
    unsigned *s_mgn = s_mgb + P.D + 1;
    for (int i = threadIdx.x; i < nT; i += blockDim.x) {
        s_nq_of[i] = P.nq_of[i]; s_nb_of[i] = P.nb_of[i];
        s_offB[i] = P.offB[i];   s_offG[i] = P.offG[i];
    }
    __syncthreads();
    const int D = P.D, nb0 = P.nb0;
    const long long lenB = P.nb * nb0, lenG = P.NQ_local * nb0;
    const long long n_items = (long long)n_terms * nvec;
    for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int it = (int)(w / nvec);
        const int iv = (int)(w - (long long)it * nvec);
        const TermDev T = P.terms[term_begin + it];
        const uint8_t *lev = P.lev + T.lev_off;
        __syncthreads();
        for (int k = threadIdx.x; k <= D; k += blockDim.x) {
            int strq = 1, strb = 1;
            for (int j = 0; j < k; ++j) { strq *= s_nq_of[j * (P.LG + 1) + lev[j]]; strb *= s_nb_of[j * (P.LG + 1) + lev[j]]; }
            s_mgq[k] = magic_ofT<BIG>(strq); s_mgb[k] = magic_ofT<BIG>(strb);
            if (k < D) {
                const int i = k * (P.LG + 1) + lev[k];
                s_tnq[k] = s_nq_of[i]; s_tnb[k] = s_nb_of[i];
                s_oB[k] = s_offB[i];   s_oG[k] = s_offG[i];
                s_str[k] = strq;       s_mgn[k] = magic_ofT<BIG>(s_nq_of[i]);
            }
        }
        __syncthreads();
        const int nq = T.nq, nbT = T.nbT;
        const int32_t *mp = P.map + T.map_off;
        double *cur = bufA, *oth = bufB;
        if (mode == NESTED_BTOG) {
            // tabPackedBasis_TO_SmolyakRepBasis: V(iB) = tabR(map) (dropped functions stay 0)
            const double *x = in + (long long)iv * lenB;
            for (int j = threadIdx.x; j < nbT; j += blockDim.x) {
                const int m = mp[j];
                for (int c = 0; c < nb0; ++c) cur[c * nbT + j] = (m > 0) ? __ldg(x + (long long)c * P.nb + (m - 1)) : 0.0;
            }
            __syncthreads();
            // BDP_TO_GDP_OF_SmolyakRep, mode 1 first; the 1 x 1 modes are one scalar factor
            double fold = 1.0;
            int left = 1, right = nbT * nb0;
            for (int k = 0; k < D; ++k) {
                const int nbk = s_tnb[k], nqk = s_tnq[k];
                right /= nbk;
                if (nbk == 1 && nqk == 1) { fold *= __ldg(P.B + s_oB[k]); continue; }
                mode_product_any<BIG>(P.B + s_oB[k], nqk, nbk, cur, oth, left, right, s_mgq[k], s_mgq[k + 1], P.use_dmma);
                double *t = cur; cur = oth; oth = t;
                left *= nqk;
                __syncthreads();
            }
            // SmolyakRep2_TO_tabR1bis: RVecG((ib0-1)*NQ + offset(iG) + q)
            double *y = out + (long long)iv * lenG + T.grid_off;
            for (int o = threadIdx.x; o < nq * nb0; o += blockDim.x) {
                const int c = mdivT<BIG>(o, s_mgq[D]), q = o - c * nq;
                y[(long long)c * P.NQ_local + q] = fold * cur[o];
            }
        } else if (mode == NESTED_GTOB) {
            // SmolyakRepBasis_TO_tabPackedBasis skips the terms with |WeightSG| < 1e-6 (...SG4.f90:1004)
            if (fabs(T.weight) < 1e-6) continue;
            const double *x = in + (long long)iv * lenG + T.grid_off;
            for (int o = threadIdx.x; o < nq * nb0; o += blockDim.x) {
                const int c = mdivT<BIG>(o, s_mgq[D]), q = o - c * nq;
                cur[o] = __ldg(x + (long long)c * P.NQ_local + q);
            }
            __syncthreads();
            double fold = T.weight;
            int left = 1, right = nq * nb0;
            for (int k = 0; k < D; ++k) {                    // GDP_TO_BDP_OF_SmolyakRep
                const int nbk = s_tnb[k], nqk = s_tnq[k];
                right /= nqk;
                if (nbk == 1 && nqk == 1) { fold *= __ldg(P.BTw + s_oB[k]); continue; }
                mode_product_any<BIG>(P.BTw + s_oB[k], nbk, nqk, cur, oth, left, right, s_mgb[k], s_mgb[k + 1], P.use_dmma);
                double *t = cur; cur = oth; oth = t;
                left *= nbk;
                __syncthreads();
            }
            double *y = out + (long long)iv * lenB;
            for (int j = threadIdx.x; j < nbT; j += blockDim.x) {
                const int m = mp[j];
                if (m > 0)
                    for (int c = 0; c < nb0; ++c) atomicAdd(y + (long long)c * P.nb + (m - 1), fold * cur[c * nbT + j]);
            }
        } else {
            // DerivOp_TO_RDP_OF_SmolaykRep on the term grid, in place in the grid vector: d2 when both indices belong to
            // one mode, d1 d1 for two modes, d1 for one index (...SG4.f90:2690-2795)
            double *y = out + (long long)iv * lenG + T.grid_off;
            for (int o = threadIdx.x; o < nq * nb0; o += blockDim.x) {
                const int c = mdivT<BIG>(o, s_mgq[D]), q = o - c * nq;
                cur[o] = y[(long long)c * P.NQ_local + q];
            }
            __syncthreads();
            for (int pass = 0; pass < 2; ++pass) {
                int k; const double *M;
                if (der1 >= 0 && der2 >= 0 && der1 == der2) { if (pass) break; k = der1; M = P.D2 + s_oG[k]; }
                else { k = pass ? der2 : der1; if (k < 0) continue; M = P.D1 + s_oG[k]; }
                const int n = s_tnq[k];
                int left = s_str[k], right = (nq / (left * n)) * nb0;
                mode_product_any<BIG>(M, n, n, cur, oth, left, right, s_mgq[k], s_mgq[k + 1], P.use_dmma);
                double *t = cur; cur = oth; oth = t;
                __syncthreads();
            }
            for (int o = threadIdx.x; o < nq * nb0; o += blockDim.x) {
                const int c = mdivT<BIG>(o, s_mgq[D]), q = o - c * nq;
                y[(long long)c * P.NQ_local + q] = cur[o];
            }
        }
    }
}

} // namespace evr

static int nested_launch(evr_sg4_plan *p, int mode, int nvec, const double *d_in, double *d_out, int der1, int der2, cudaStream_t st)
{
    if (p->n_terms == 0) return 0;
    const int nT = p->D * (p->LG + 1);
    const size_t ints = (size_t)(4 * nT + 5 * p->D + 3 * (p->D + 1)) * sizeof(int);
    // the terms beyond the shared-memory budget come first in work order (sg4_plan.cu: gen_class_of): global work buffers
    const int n_big = p->n_big_terms, n_small = p->n_terms - n_big;
    if (n_big > 0) {
        const size_t per_cta = (size_t)2 * p->cap * sizeof(double);
        const char *e = getenv("EVR_SG4_SCRATCH_MB");
        const size_t budget = (size_t)(e ? std::max(1, atoi(e)) : 4096) << 20;
        const int cmax = (int)std::min<size_t>((size_t)p->sm_count * 2, std::max<size_t>(1, budget / per_cta));
        if (!p->d_nscratch) {
            if (cudaMalloc((void **)&p->d_nscratch, per_cta * cmax) != cudaSuccess) { cudaGetLastError(); return fail("evr_sg4 nested: cannot allocate the work buffers of the large terms"); }
        }
        const long long items = (long long)n_big * nvec;
        const int ctas = (int)std::max<long long>(1, std::min<long long>(items, cmax));
        evr::sg4_nested_kernel<true><<<ctas, 256, ints, st>>>(p->pd, mode, nvec, d_in, d_out, der1, der2, 0, n_big, p->cap, p->d_nscratch);
        if (cudaGetLastError() != cudaSuccess) return fail("evr_sg4 nested: kernel launch failed");
        p->launches += 1;
    }
    if (n_small == 0) return 0;
    const size_t smem = (size_t)2 * p->cap_small * sizeof(double) + ints;
    if (smem > 227 * 1024) return fail("evr_sg4 nested: shared-memory budget exceeded");
    static size_t attr_max[64] = {0};
    size_t &amax = attr_max[p->device & 63];
    if (smem > amax) {
        if (cudaFuncSetAttribute(evr::sg4_nested_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return fail("evr_sg4 nested: cudaFuncSetAttribute(smem) failed");
        amax = smem;
    }
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, evr::sg4_nested_kernel<false>, 256, smem) != cudaSuccess || occ < 1)
        return fail("evr_sg4 nested: kernel cannot be resident");
    const long long items = (long long)n_small * nvec;
    const int ctas = (int)std::max<long long>(1, std::min<long long>(items, (long long)p->sm_count * occ));
    evr::sg4_nested_kernel<false><<<ctas, 256, smem, st>>>(p->pd, mode, nvec, d_in, d_out, der1, der2, n_big, n_small, p->cap_small, nullptr);
    if (cudaGetLastError() != cudaSuccess) return fail("evr_sg4 nested: kernel launch failed");
    p->launches += 1;
    return 0;
}

static int nested_check(evr_sg4_plan *p, int nvec, const void *a, const void *b, const char *who)
{
    if (!p) return fail(std::string(who) + ": null plan");
    if (!p->sub.empty()) return fail(std::string(who) + ": not available on a multi-device plan (create the plan with evr_sg4_set_devices(1))");
    if (nvec < 1 || !a || !b) return fail(std::string(who) + ": bad arguments");
    if (cudaSetDevice(p->device) != cudaSuccess) return fail(std::string(who) + ": cudaSetDevice failed");
    return 0;
}

extern "C" int evr_sg4_BtoG_device(evr_sg4_plan *p, int nvec, const double *d_RvecB, double *d_RvecG, void *stream)
{
    if (nested_check(p, nvec, d_RvecB, d_RvecG, "evr_sg4_BtoG_device")) return 1;
    return nested_launch(p, evr::NESTED_BTOG, nvec, d_RvecB, d_RvecG, -1, -1, (cudaStream_t)stream);
}
extern "C" int evr_sg4_GtoB_device(evr_sg4_plan *p, int nvec, const double *d_RvecG, double *d_RvecB, void *stream)
{
    if (nested_check(p, nvec, d_RvecG, d_RvecB, "evr_sg4_GtoB_device")) return 1;
    if (cudaMemsetAsync(d_RvecB, 0, (size_t)nvec * p->nb * p->nb0 * sizeof(double), (cudaStream_t)stream) != cudaSuccess)   // tabR(:) = ZERO (:978)
        return fail("evr_sg4_GtoB_device: cudaMemsetAsync failed");
    return nested_launch(p, evr::NESTED_GTOB, nvec, d_RvecG, d_RvecB, -1, -1, (cudaStream_t)stream);
}
extern "C" int evr_sg4_DerivOp_G_device(evr_sg4_plan *p, int nvec, double *d_RvecG, int mode1, int mode2, void *stream)
{
    if (nested_check(p, nvec, d_RvecG, d_RvecG, "evr_sg4_DerivOp_G_device")) return 1;
    if (mode1 < 0 || mode2 < 0 || mode1 > p->D || mode2 > p->D) return fail("evr_sg4_DerivOp_G_device: mode out of range");
    int a = mode1 - 1, b = mode2 - 1;                        // 1-based SG4 mode owning each derivative index, 0 = none
    if (a < 0 && b < 0) return 0;                            // (0,0): RvecG unchanged (...BtoG_GtoB.f90:1396-1397)
    if (a < 0) { a = b; b = -1; }
    return nested_launch(p, evr::NESTED_DERIV, nvec, d_RvecG, d_RvecG, a, b, (cudaStream_t)stream);
}

// host-buffer variants (staging on the plan's stream)
static int nested_host(evr_sg4_plan *p, int mode, int nvec, const double *in, double *out, int m1, int m2, const char *who)
{
    if (nested_check(p, nvec, in, out, who)) return 1;
    const size_t lenB = (size_t)p->nb * p->nb0, lenG = (size_t)p->NQ_local * p->nb0;
    const size_t n_in = nvec * (mode == evr::NESTED_BTOG ? lenB : lenG), n_out = nvec * (mode == evr::NESTED_GTOB ? lenB : lenG);
    double *d_in = nullptr, *d_out = nullptr;
    if (cudaMalloc((void **)&d_in, std::max<size_t>(n_in, 1) * 8) != cudaSuccess) return fail(std::string(who) + ": cudaMalloc failed");
    if (mode != evr::NESTED_DERIV && cudaMalloc((void **)&d_out, std::max<size_t>(n_out, 1) * 8) != cudaSuccess) { cudaFree(d_in); return fail(std::string(who) + ": cudaMalloc failed"); }
    int rc = 0;
    if (cudaMemcpyAsync(d_in, in, n_in * 8, cudaMemcpyHostToDevice, p->stream) != cudaSuccess) rc = fail(std::string(who) + ": H2D copy failed");
    if (!rc && mode == evr::NESTED_BTOG) rc = evr_sg4_BtoG_device(p, nvec, d_in, d_out, p->stream);
    if (!rc && mode == evr::NESTED_GTOB) rc = evr_sg4_GtoB_device(p, nvec, d_in, d_out, p->stream);
    if (!rc && mode == evr::NESTED_DERIV) rc = evr_sg4_DerivOp_G_device(p, nvec, d_in, m1, m2, p->stream);
    if (!rc && cudaMemcpyAsync(out, mode == evr::NESTED_DERIV ? d_in : d_out, n_out * 8, cudaMemcpyDeviceToHost, p->stream) != cudaSuccess)
        rc = fail(std::string(who) + ": D2H copy failed");
    if (cudaStreamSynchronize(p->stream) != cudaSuccess && !rc) rc = fail(std::string(who) + ": kernel failed");
    cudaFree(d_in); cudaFree(d_out);
    return rc;
}
extern "C" int evr_sg4_BtoG(evr_sg4_plan *p, int nvec, const double *RvecB, double *RvecG)
{ return nested_host(p, evr::NESTED_BTOG, nvec, RvecB, RvecG, 0, 0, "evr_sg4_BtoG"); }
extern "C" int evr_sg4_GtoB(evr_sg4_plan *p, int nvec, const double *RvecG, double *RvecB)
{ return nested_host(p, evr::NESTED_GTOB, nvec, RvecG, RvecB, 0, 0, "evr_sg4_GtoB"); }
extern "C" int evr_sg4_DerivOp_G(evr_sg4_plan *p, int nvec, const double *RvecG_in, double *RvecG_out, int mode1, int mode2)
{ return nested_host(p, evr::NESTED_DERIV, nvec, RvecG_in, RvecG_out, mode1, mode2, "evr_sg4_DerivOp_G"); }
