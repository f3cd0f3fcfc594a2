// sg4_kernels.cuh -- sm_100a kernels of the SG4 H|psi> action.
//
// One CTA processes one Smolyak term at a time (persistent grid, terms sorted by
// cost): gather -> mode-by-mode B->G -> sum_iterm F_iterm(Q) d^(i,j) on the term grid
// -> mode-by-mode G->B -> weighted scatter-add.  The whole term lives in shared memory
// between the gather and the scatter; HBM sees only the packed psi / Hpsi vectors, the
// int32 mapping slice and the operator grids (SURVEY.md 8d "algorithmic bytes").
//
// Reference routines fused here (reference tree, file:line):
//   tabPackedBasis_TO_tabR_AT_iG        sub_Basis_SG4/sub_module_basis_BtoG_GtoB_SG4.f90:1176-1246
//   BDP_TO_GDP_OF_SmolyakRep            ...:2385-2496
//   sub_TabOpPsi_OF_ONEDP_FOR_SGtype4   sub_Operator/sub_OpPsi_SG4.f90:1354-1546 (type_Op 0 and 1)
//   DerivOp_TO_RDP_OF_SmolaykRep        ...BtoG_GtoB_SG4.f90:2690-2795
//   GDP_TO_BDP_OF_SmolyakRep            ...:2497-2581
//   tabR_AT_iG_TO_tabPackedBasis        ...:1250-1289
#pragma once
#include <cstdint>
#include "sg4_internal.h"

namespace evr {

// ---- device-side plan -------------------------------------------------------------
struct TermDev {                 // one per Smolyak term of this plan's range (work order)
    long long map_off;           // start of the term's slice in d_map
    long long grid_off;          // start of the term's slice in every operator grid
    double    weight;            // WeightSG(iG)
    int       nbT, nq;           // prod nb_k, prod nq_k
    int       lev_off;           // start of the term's D levels in d_lev
    int       pad;
    double    wfold;             // weight * prod over the 1x1 modes of B*BTw (generic kernel skips those modes)
};

struct GenClassDev {             // one launch of the generic kernel: terms of similar size, one CTA size
    int term_begin, n_terms;     // range in work order
    int cap;                     // doubles per work buffer (max term size of the class * nb0)
    int dcache;                  // 1: third buffer holds the first derivative of the current sweep mode
    double *scratch;             // class of terms too large for shared memory: [CTAs][2 or 3][cap] in global memory, else nullptr
    const int *list;             // nullptr: terms term_begin .. term_begin + n_terms of the work order; else work-order indices
                                 // list[term_begin ..] (the terms a fast-path plan leaves to this kernel)
};

struct OpTermDev {               // one per live (not grid_zero) operator term
    int m1, m2;                  // 0-based SG4 modes of the derivative, -1 = none (m1<=m2 if both)
    int grid_slot;               // >=0: index of the variable grid; -1: constant (Mat_cte)
    int pad;
    double cte[EVR_MAXCH * EVR_MAXCH];   // cte[i + nb0*j] = Mat_cte(i,j)
};

struct PlanDev {
    int D, LG, nb0, n_terms;     // n_terms = terms in this plan's range
    long long nb;                // packed basis size per channel
    long long NQ_local;          // grid points of the range (= stride between grid blocks)
    int n_opterms, n_var;
    int cap;                     // doubles per shared-memory buffer
    int type_Op;
    const TermDev   *terms;      // [n_terms], work order
    const uint8_t   *lev;        // [n_terms*D]
    const int32_t   *map;        // [S_local]
    const int32_t   *nq_of;      // [D*(LG+1)]
    const int32_t   *nb_of;
    const int32_t   *offB;       // [D*(LG+1)] offsets into B/BTw pools
    const int32_t   *offG;       // [D*(LG+1)] offsets into D1/D2 pools
    const double    *B, *BTw, *D1, *D2;
    const OpTermDev *opterms;    // [n_opterms]: n_plain on-the-fly terms, then the terms of each sweep
    const double    *grids;      // [n_var][nb0*nb0][NQ_local]  (slot-major; (i + nb0*j)-major; point)
    // mixed-derivative sweeps: sweep s caches d/dQ_a psi (a = sweep_mode[s]) of the whole term in shared memory
    // and serves opterms [sweep_begin[s], sweep_begin[s+1]): d_a d_b (m1 = a, m2 = b) and d_a alone (m2 = -1)
    int n_plain, n_sweeps;
    int sweep_mode[EVR_MAXD];
    int sweep_begin[EVR_MAXD + 1];
    // deterministic mode: see FastPlanDev (entry = position in the plan's mapping slice)
    double *stage;
    long long stage_ld;
    int use_dmma;                // FP64 tensor-core mode products for the large modes (EVR_SG4_DMMA=0 disables)
};

// ---- division by a per-term constant: q / d = umulhi(q, magic(d)) --------------------------------
// magic(d) = floor(2^32 / d) + 1, exact while q * d < 2^32 (term sizes are < 2^15 values); d = 1 is encoded as 0
__device__ __forceinline__ unsigned magic_of(const int d) { return d <= 1 ? 0u : 0xFFFFFFFFu / (unsigned)d + 1u; }
__device__ __forceinline__ int mdiv(const int q, const unsigned mg) { return mg ? (int)__umulhi((unsigned)q, mg) : q; }
// BIG = true (terms that do not fit in shared memory, work buffers in global memory): the per-term constant is the
// divisor itself and the division is the hardware sequence -- no bound on q * d
template <bool BIG> __device__ __forceinline__ unsigned magic_ofT(const int d) { return BIG ? (unsigned)(d <= 1 ? 0 : d) : magic_of(d); }
template <bool BIG> __device__ __forceinline__ int mdivT(const int q, const unsigned mg)
{
    if (BIG) return mg ? (int)((unsigned)q / mg) : q;
    return mdiv(q, mg);
}

// ---- sum_b M[b*ms] * x[b*xs]: one accumulator.  Four independent partial sums (to break the chain of n dependent DFMAs of
// the large modes) were measured and are SLOWER: 268 vs 216 us for the 27-vector HCN block, 223 vs 190 us for HNO3 LB6/LG7
// (profiles/r2/sweep20_generic_dot4.txt) -- the extra registers spill at the kernel's 64-register bound and the loop is
// bound by its two memory instructions per multiply-add, not by the DFMA latency.
__device__ __forceinline__ double dot_strided(const double *__restrict__ M, const int ms, const double *x, const int xs, const int n)
{
    double s = 0.0;
    for (int b = 0; b < n; ++b) s = fma(__ldg(M + b * ms), x[b * xs], s);
    return s;
}

// ---- generic one-mode product through shared memory ----------------------------------
//   out[a + left*(q + n_out*c)] = sum_b M[q + n_out*b] * in[a + left*(b + n_in*c)]
//   a < left, c < right (right already includes the channel count), M in global (L1-resident).
template <bool BIG = false>
__device__ __forceinline__ void mode_product(const double *__restrict__ M, int n_out, int n_in,
                                             const double *in, double *out, int left, int right,
                                             const unsigned mg_left, const unsigned mg_lo)
{
    const int total = left * n_out * right;
    const int lo = left * n_out;
    for (int o = threadIdx.x; o < total; o += blockDim.x) {
        const int c = mdivT<BIG>(o, mg_lo);
        const int r = o - c * lo;
        const int q = mdivT<BIG>(r, mg_left);
        const int a = r - q * left;
        out[o] = dot_strided(M + q, n_out, in + a + left * n_in * c, left, n_in);
    }
}

// ---- the same mode product on the FP64 tensor cores (DMMA, mma.sync.aligned.m8n8k4.f64) --------------------------
// Used for modes whose matrices are large enough to be compute-bound (n_out, n_in >= EVR_DMMA_MIN: the Pl0 mode of
// HCN_UT, 20 ... 80 points): one fragment load feeds 256 multiply-adds, where the thread-per-output product above issues
// one matrix load and one shared-memory load per multiply-add.  Measured on the 80 x 80 matrix in isolation
// (profiles/micro/dmma_vs_dfma.cu, profiles/r2/dmma_vs_dfma_pl0_mode.txt): 10.9 TFLOP/s against 4.7 (this file's scalar
// product) and 5.6 (register tiles), 30 M tensor-pipe instructions instead of 242 M FP64-pipe instructions.  For the
// 3 x 3 ... 15 x 15 matrices of every other mode an 8 x 8 x 4 fragment would be 14-47 % full: they stay on DFMA.
// A warp owns 8 columns (a, c) and all row tiles of the output (<= EVR_DMMA_MAXROWT x 8 rows); edges are zero-padded.
// Fragment layout (PTX ISA, m8n8k4 .f64): lane l holds A(row l/4, k l%4), B(k l%4, col l/4), C(row l/4, cols 2*(l%4)+{0,1}).
#define EVR_DMMA_MIN 16
#define EVR_DMMA_MAXROWT 32
#define EVR_DMMA_ROWCHUNK 5
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, const double a, const double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <bool BIG = false>
__device__ __forceinline__ void mode_product_dmma(const double *__restrict__ M, const int n_out, const int n_in,
                                                  const double *in, double *out, const int left, const int right,
                                                  const unsigned mg_left)
{
    const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int ar = lane >> 2, ak = lane & 3;
    const int ncols = left * right, ncolt = (ncols + 7) >> 3, nrowt = (n_out + 7) >> 3;
    for (int ct = warp; ct < ncolt; ct += nwarps) {
        // B operand: column ct*8 + ar  ->  (a, c), base address of its pencil
        const int colB = ct * 8 + ar;
        const bool okB = colB < ncols;
        const int cB = okB ? mdivT<BIG>(colB, mg_left) : 0, aB = colB - cB * left;
        const double *xB = in + aB + left * n_in * cB;
        // row tiles in chunks of EVR_DMMA_ROWCHUNK (10 accumulator registers pairs per lane: the generic kernel is compiled for
        // 64 registers); every chunk re-reads the B fragments of its columns from shared memory
        for (int t0 = 0; t0 < nrowt; t0 += EVR_DMMA_ROWCHUNK) {
            double acc[EVR_DMMA_ROWCHUNK][2];
#pragma unroll
            for (int t = 0; t < EVR_DMMA_ROWCHUNK; ++t) acc[t][0] = acc[t][1] = 0.0;
            for (int k0 = 0; k0 < n_in; k0 += 4) {
                const int k = k0 + ak;
                const double bf = (okB && k < n_in) ? xB[left * k] : 0.0;
#pragma unroll
                for (int t = 0; t < EVR_DMMA_ROWCHUNK; ++t)
                    if (t0 + t < nrowt) {
                        const int row = (t0 + t) * 8 + ar;
                        const double af = (row < n_out && k < n_in) ? __ldg(M + row + n_out * k) : 0.0;
                        dmma_m8n8k4(acc[t][0], acc[t][1], af, bf);
                    }
            }
            // C: rows (t0+t)*8 + ar, columns ct*8 + 2*ak + {0, 1}
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = ct * 8 + 2 * ak + u;
                if (col < ncols) {
                    const int c = mdivT<BIG>(col, mg_left), a = col - c * left;
                    double *y = out + a + left * n_out * c;
#pragma unroll
                    for (int t = 0; t < EVR_DMMA_ROWCHUNK; ++t) {
                        const int row = (t0 + t) * 8 + ar;
                        if (t0 + t < nrowt && row < n_out) y[left * row] = acc[t][u];
                    }
                }
            }
        }
    }
}
// dispatcher: DMMA for the large modes (whole warps only), DFMA otherwise
template <bool BIG = false>
__device__ __forceinline__ void mode_product_any(const double *__restrict__ M, int n_out, int n_in,
                                                 const double *in, double *out, int left, int right,
                                                 const unsigned mg_left, const unsigned mg_lo, const int use_dmma)
{
    if (use_dmma && n_out >= EVR_DMMA_MIN && n_in >= EVR_DMMA_MIN && n_out <= 8 * EVR_DMMA_MAXROWT && (blockDim.x & 31) == 0)
        mode_product_dmma<BIG>(M, n_out, n_in, in, out, left, right, mg_left);
    else
        mode_product<BIG>(M, n_out, n_in, in, out, left, right, mg_left, mg_lo);
}

// ---- generic term kernel (any type_Op 0/1 term list, any mode sizes that fit) ----------
// Launched once per size class (GenClassDev) with a CTA of 32/64/128/256 threads, so that the many small
// terms of a curvilinear configuration (HNO3_UT: 38 points per term on average) do not idle a wide CTA.
// dynamic smem: [2 or 3 * cap doubles][ints: nq_of,nb_of,offB,offG (4*D*(LG+1))][per-term ints 5*D + 3*(D+1)]
//               [3 ints per operator term]
#define EVR_GEN_SMEM_INTS(nT, D, nop) (4 * (nT) + 5 * (D) + 3 * ((D) + 1) + 3 * (nop))
// BIG = true: the class of terms that do not fit in shared memory (no such limit in the reference): the two or three work
// buffers of a CTA live in its slice of Cc.scratch (global memory, L2-resident for all but enormous terms) and the index
// divisions are exact for any term size; everything else is the same code.
#ifndef EVR_GEN_MINBLOCKS
#define EVR_GEN_MINBLOCKS 4
#endif
template <bool BIG>
static __global__ void __launch_bounds__(256, EVR_GEN_MINBLOCKS)
sg4_term_kernel_generic(const PlanDev P, const GenClassDev Cc, const int npsi,
                        const double *__restrict__ psi, double *__restrict__ Hpsi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nbuf = Cc.dcache ? 3 : 2;
    double *bufA = BIG ? Cc.scratch + (size_t)blockIdx.x * nbuf * Cc.cap : reinterpret_cast<double *>(smem_raw);
    double *bufB = bufA + Cc.cap;
    double *bufC = bufB + Cc.cap;                  // only with Cc.dcache
    int *s_nq_of = BIG ? reinterpret_cast<int *>(smem_raw) : reinterpret_cast<int *>(bufA + (size_t)nbuf * Cc.cap);
    const int nT = P.D * (P.LG + 1);
    int *s_nb_of = s_nq_of + nT;
    int *s_offB  = s_nb_of + nT;
    int *s_offG  = s_offB + nT;
    int *s_tnq   = s_offG + nT;        // per-term: nq_k
    int *s_tnb   = s_tnq + P.D;        //           nb_k
    int *s_oB    = s_tnb + P.D;        //           table offsets for (k,l_k)
    int *s_oG    = s_oB + P.D;
    int *s_str   = s_oG + P.D;         //           grid stride of mode k (first mode fastest)
    unsigned *s_mgq = reinterpret_cast<unsigned *>(s_str + P.D);   // magic of prod_{j<k} nq_j, k = 0..D
    unsigned *s_mgb = s_mgq + P.D + 1;                             // magic of prod_{j<k} nb_j, k = 0..D
    unsigned *s_mgn = s_mgb + P.D + 1;                             // magic of nq_k
    int *s_op = reinterpret_cast<int *>(s_mgn + P.D + 1);          // per operator term: m1, m2, grid_slot

    for (int i = threadIdx.x; i < nT; i += blockDim.x) {
        s_nq_of[i] = P.nq_of[i]; s_nb_of[i] = P.nb_of[i];
        s_offB[i] = P.offB[i];   s_offG[i] = P.offG[i];
    }
    for (int t = threadIdx.x; t < P.n_opterms; t += blockDim.x) {
        s_op[3 * t] = P.opterms[t].m1; s_op[3 * t + 1] = P.opterms[t].m2; s_op[3 * t + 2] = P.opterms[t].grid_slot;
    }
    __syncthreads();

    const int D = P.D, nb0 = P.nb0;
    const long long nvec = P.nb * nb0;

    // work item = (term, right-hand side): blocks of RHS (Davidson) fill the GPU even when a configuration has
    // few Smolyak terms (HCN_UT: 85 terms x 27 vectors)
    const long long n_items = (long long)Cc.n_terms * npsi;
    for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int it = (int)(w / npsi);
        const int ip_only = (int)(w - (long long)it * npsi);
        const TermDev T = P.terms[Cc.list ? Cc.list[Cc.term_begin + it] : Cc.term_begin + it];
        const uint8_t *lev = P.lev + T.lev_off;
        __syncthreads();               // previous term fully done before the per-term tables change
        for (int k = threadIdx.x; k <= D; k += blockDim.x) {
            int strq = 1, strb = 1;    // grid / basis stride of mode k (first mode fastest)
            for (int j = 0; j < k; ++j) { strq *= s_nq_of[j * (P.LG + 1) + lev[j]]; strb *= s_nb_of[j * (P.LG + 1) + lev[j]]; }
            s_mgq[k] = magic_ofT<BIG>(strq); s_mgb[k] = magic_ofT<BIG>(strb);
            if (k < D) {
                const int i = k * (P.LG + 1) + lev[k];
                s_tnq[k] = s_nq_of[i]; s_tnb[k] = s_nb_of[i];
                s_oB[k] = s_offB[i];   s_oG[k] = s_offG[i];
                s_str[k] = strq;       s_mgn[k] = magic_ofT<BIG>(s_nq_of[i]);
            }
        }
        // pull the term's slices of the operator grids into L2 while the gather and the B->G passes run
        if (P.n_var > 0) {
            const int lines = (T.nq * 8 + 127) / 128 + 1;
            const int nslice = P.n_var * nb0 * nb0;
            const char *g0 = reinterpret_cast<const char *>(P.grids + T.grid_off);
            for (int i = threadIdx.x; i < nslice * lines; i += blockDim.x) {
                const int sl = i / lines, ln = i - sl * lines;
                const int byte = min(ln * 128, T.nq * 8 - 8);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(g0 + (long long)sl * P.NQ_local * 8 + byte));
            }
        }
        __syncthreads();
        const int nq = T.nq, nbT = T.nbT;
        const int32_t *mp = P.map + T.map_off;

        for (int ip = ip_only; ip <= ip_only; ++ip) {
            const double *x = psi + (long long)ip * nvec;
            double *y = Hpsi + (long long)ip * nvec;
            // ---- gather
            for (int j = threadIdx.x; j < nbT; j += blockDim.x) {
                const int m = mp[j];
                for (int c = 0; c < nb0; ++c)
                    bufA[c * nbT + j] = (m > 0) ? __ldg(x + (long long)c * P.nb + (m - 1)) : 0.0;
            }
            __syncthreads();
            double *cur = bufA, *oth = bufB;
            // ---- B -> G, mode 1 first
            {
                int left = 1, right = nbT * nb0;
                for (int k = 0; k < D; ++k) {
                    const int nbk = s_tnb[k], nqk = s_tnq[k];
                    right /= nbk;
                    if (nbk == 1 && nqk == 1) continue;        // scalar folded into T.wfold by the plan
                    mode_product_any<BIG>(P.B + s_oB[k], nqk, nbk, cur, oth, left, right, s_mgq[k], s_mgq[k + 1], P.use_dmma);
                    double *t = cur; cur = oth; oth = t;
                    left *= nqk;
                    __syncthreads();
                }
            }
            // ---- operator on the term grid: oth(q,i) = sum_iterm sum_j F(q,i,j) [d psi](q,j)
            auto add_term = [&](const int t, const double (&d)[EVR_MAXCH], double (&acc)[EVR_MAXCH], const int q) {
                const int slot = s_op[3 * t + 2];
                if (slot < 0) {
                    const double *cte = P.opterms[t].cte;
#pragma unroll
                    for (int i = 0; i < EVR_MAXCH; ++i) if (i < nb0)
#pragma unroll
                        for (int j = 0; j < EVR_MAXCH; ++j) if (j < nb0)
                            acc[i] = fma(__ldg(cte + i + nb0 * j), d[j], acc[i]);
                } else {
                    const double *g = P.grids + ((long long)slot * nb0 * nb0) * P.NQ_local + T.grid_off + q;
#pragma unroll
                    for (int i = 0; i < EVR_MAXCH; ++i) if (i < nb0)
#pragma unroll
                        for (int j = 0; j < EVR_MAXCH; ++j) if (j < nb0)
                            acc[i] = fma(__ldg(g + (long long)(i + nb0 * j) * P.NQ_local), d[j], acc[i]);
                }
            };
            const int n_otf = Cc.dcache ? P.n_plain : P.n_opterms;    // terms whose derivative is formed on the fly
            for (int q = threadIdx.x; q < nq; q += blockDim.x) {
                double acc[EVR_MAXCH];
#pragma unroll
                for (int i = 0; i < EVR_MAXCH; ++i) acc[i] = 0.0;
                for (int t = 0; t < n_otf; ++t) {
                    double d[EVR_MAXCH];
                    const int m1 = s_op[3 * t], m2 = s_op[3 * t + 1];
                    if (m1 < 0 && m2 < 0) {
#pragma unroll
                        for (int j = 0; j < EVR_MAXCH; ++j) if (j < nb0) d[j] = cur[j * nq + q];
                    } else if (m1 < 0 || m2 < 0 || m1 == m2) {
                        const int k = (m1 >= 0) ? m1 : m2;
                        const double *M = ((m1 == m2) ? P.D2 : P.D1) + s_oG[k];
                        const int n = s_tnq[k], st = s_str[k];
                        const int qd = mdivT<BIG>(q, s_mgq[k]);
                        const int qk = qd - mdivT<BIG>(qd, s_mgn[k]) * n;
                        const int base = q - qk * st;
#pragma unroll
                        for (int j = 0; j < EVR_MAXCH; ++j) if (j < nb0) d[j] = dot_strided(M + qk, n, cur + j * nq + base, st, n);
                    } else {
                        const double *Ma = P.D1 + s_oG[m1], *Mb = P.D1 + s_oG[m2];
                        const int na = s_tnq[m1], sa = s_str[m1], nbb = s_tnq[m2], sb = s_str[m2];
                        const int qda = mdivT<BIG>(q, s_mgq[m1]), qdb = mdivT<BIG>(q, s_mgq[m2]);
                        const int qa = qda - mdivT<BIG>(qda, s_mgn[m1]) * na, qb = qdb - mdivT<BIG>(qdb, s_mgn[m2]) * nbb;
                        const int base = q - qa * sa - qb * sb;
#pragma unroll
                        for (int j = 0; j < EVR_MAXCH; ++j) if (j < nb0) {
                            double s = 0.0;
                            for (int b2 = 0; b2 < nbb; ++b2) {
                                const double s1 = dot_strided(Ma + qa, na, cur + j * nq + base + b2 * sb, sa, na);
                                s = fma(__ldg(Mb + qb + nbb * b2), s1, s);
                            }
                            d[j] = s;
                        }
                    }
                    add_term(t, d, acc, q);
                }
#pragma unroll
                for (int i = 0; i < EVR_MAXCH; ++i) if (i < nb0) oth[i * nq + q] = acc[i];
            }
            __syncthreads();
            if (Cc.dcache) {
                // mixed derivatives d_a d_b: the first derivative along a is formed once per sweep for the whole
                // term (n_a multiply-adds per point) and every term of the sweep needs n_b more, instead of
                // n_a*n_b per point and term on the fly
                for (int sw = 0; sw < P.n_sweeps; ++sw) {
                    const int a = P.sweep_mode[sw];
                    const int na = s_tnq[a], sa = s_str[a];
                    const double *Ma = P.D1 + s_oG[a];
                    for (int o = threadIdx.x; o < nq * nb0; o += blockDim.x) {
                        const int j = mdivT<BIG>(o, s_mgq[D]), q = o - j * nq;
                        const int qda = mdivT<BIG>(q, s_mgq[a]);
                        const int qa = qda - mdivT<BIG>(qda, s_mgn[a]) * na;
                        bufC[o] = dot_strided(Ma + qa, na, cur + j * nq + (q - qa * sa), sa, na);
                    }
                    __syncthreads();
                    const int t0 = P.sweep_begin[sw], t1 = P.sweep_begin[sw + 1];
                    for (int q = threadIdx.x; q < nq; q += blockDim.x) {
                        double acc[EVR_MAXCH];
#pragma unroll
                        for (int i = 0; i < EVR_MAXCH; ++i) acc[i] = (i < nb0) ? oth[i * nq + q] : 0.0;
                        for (int t = t0; t < t1; ++t) {
                            double d[EVR_MAXCH];
                            const int m2 = s_op[3 * t + 1];
                            if (m2 < 0) {
#pragma unroll
                                for (int j = 0; j < EVR_MAXCH; ++j) if (j < nb0) d[j] = bufC[j * nq + q];
                            } else {
                                const double *Mb = P.D1 + s_oG[m2];
                                const int nbb = s_tnq[m2], sb = s_str[m2];
                                const int qdb = mdivT<BIG>(q, s_mgq[m2]);
                                const int qb = qdb - mdivT<BIG>(qdb, s_mgn[m2]) * nbb;
                                const int base = q - qb * sb;
#pragma unroll
                                for (int j = 0; j < EVR_MAXCH; ++j) if (j < nb0) d[j] = dot_strided(Mb + qb, nbb, bufC + j * nq + base, sb, nbb);
                            }
                            add_term(t, d, acc, q);
                        }
#pragma unroll
                        for (int i = 0; i < EVR_MAXCH; ++i) if (i < nb0) oth[i * nq + q] = acc[i];
                    }
                    __syncthreads();
                }
            }
            { double *t = cur; cur = oth; oth = t; }
            // ---- G -> B
            {
                int left = 1, right = nq * nb0;
                for (int k = 0; k < D; ++k) {
                    const int nbk = s_tnb[k], nqk = s_tnq[k];
                    right /= nqk;
                    if (nbk == 1 && nqk == 1) continue;
                    mode_product_any<BIG>(P.BTw + s_oB[k], nbk, nqk, cur, oth, left, right, s_mgb[k], s_mgb[k + 1], P.use_dmma);
                    double *t = cur; cur = oth; oth = t;
                    left *= nbk;
                    __syncthreads();
                }
            }
            // ---- weighted scatter-add
            if (P.stage) {
                double *sg = P.stage + (long long)ip * nb0 * P.stage_ld + T.map_off;
                for (int j = threadIdx.x; j < nbT; j += blockDim.x)
                    if (mp[j] > 0)
                        for (int c = 0; c < nb0; ++c) sg[(long long)c * P.stage_ld + j] = T.wfold * cur[c * nbT + j];
            } else
            for (int j = threadIdx.x; j < nbT; j += blockDim.x) {
                const int m = mp[j];
                if (m > 0)
                    for (int c = 0; c < nb0; ++c)
                        atomicAdd(y + (long long)c * P.nb + (m - 1), T.wfold * cur[c * nbT + j]);
            }
            __syncthreads();
        }
    }
}

// deterministic mode: out[v*nb + i] = sum of the staged entries of packed element i in a fixed order.  One warp per
// element (the lists are very uneven: the lowest basis functions belong to every Smolyak term): lane L adds the entries
// k = L, L+32, ... in order, then a fixed shuffle tree combines the 32 partial sums.
static __global__ void __launch_bounds__(256)
sg4_collect_kernel(const long long nb, const int nvec, const long long *__restrict__ off,
                   const int32_t *__restrict__ ent, const double *__restrict__ stage, const long long stage_ld,
                   double *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long i = warp0; i < nb; i += nwarps) {
        const long long k0 = off[i], k1 = off[i + 1];
        for (int v = 0; v < nvec; ++v) {
            const double *sg = stage + (long long)v * stage_ld;
            double s = 0.0;
            for (long long k = k0 + lane; k < k1; k += 32) s += sg[ent[k]];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, d);
            if (lane == 0) out[(long long)v * nb + i] = s;
        }
    }
}

} // namespace evr

namespace evr {

// ---- type_Op = 10 with the metric tensor cached on the device --------------------------
// H psi = -1/2 (Jac sq)^-1 sum_i d_i [ Jac sum_j G^{ji} d_j (sq psi) ] + V psi,  sq = sqrt(rho/Jac)
// ref: sub_TabOpPsi_OF_ONEDP_FOR_SGtype4 CASE (10), sub_Operator/sub_OpPsi_SG4.f90:1548-1650; the reference
// recomputes G(Q) with Tnum at every grid point of every call (:2717-2728), here G, Jac and sq are plan data.
struct Op10Dev {
    int n_act;                   // number of active coordinates
    int act_mode[EVR_MAXD];      // 0-based SG4 mode owning active coordinate j
    int has_V;
    int nqmax;                   // largest term grid of the plan's range
    const double *V;             // [nb0*nb0][NQ_local]
    const double *GG;            // [n_act*n_act][NQ_local]   GG[(j + n*i)*NQ + q] = GGiq(q,j,i); sym: [n(n+1)/2][NQ_local],
                                 // component (min(i,j) + max(i,j)(max(i,j)+1)/2) -- the metric tensor is symmetric
    const double *Jac, *sq;      // [NQ_local]
    int sym;                     // 1: GG stored as its upper triangle (plan set-up found GG(q,j,i) == GG(q,i,j) everywhere)
};

// dynamic smem: bufA[cap] | bufB[cap] | R[n_act][nqmax] | chi[nqmax] | ints (as the generic kernel)
static __global__ void __launch_bounds__(256)
sg4_term_kernel_type10(const PlanDev P, const Op10Dev O, const int npsi,
                       const double *__restrict__ psi, double *__restrict__ Hpsi)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *bufA = reinterpret_cast<double *>(smem_raw);
    double *bufB = bufA + P.cap;
    double *sR   = bufB + P.cap;
    double *chi  = sR + (size_t)O.n_act * O.nqmax;
    int *s_nq_of = reinterpret_cast<int *>(chi + O.nqmax);
    const int nT = P.D * (P.LG + 1);
    int *s_nb_of = s_nq_of + nT;
    int *s_offB  = s_nb_of + nT;
    int *s_offG  = s_offB + nT;
    int *s_tnq   = s_offG + nT;
    int *s_tnb   = s_tnq + P.D;
    int *s_oB    = s_tnb + P.D;
    int *s_oG    = s_oB + P.D;
    int *s_str   = s_oG + P.D;
    unsigned *s_mgs = reinterpret_cast<unsigned *>(s_str + P.D);   // magic of the grid stride of mode k
    unsigned *s_mgn = s_mgs + P.D;                                 // magic of nq_k

    for (int i = threadIdx.x; i < nT; i += blockDim.x) {
        s_nq_of[i] = P.nq_of[i]; s_nb_of[i] = P.nb_of[i];
        s_offB[i] = P.offB[i];   s_offG[i] = P.offG[i];
    }
    __syncthreads();
    const int D = P.D, nb0 = P.nb0, n = O.n_act;
    const long long nvec = P.nb * nb0;
    const long long n_items = (long long)P.n_terms * npsi;

    for (long long w = blockIdx.x; w < n_items; w += gridDim.x) {
        const int it = (int)(w / npsi);
        const int ip = (int)(w - (long long)it * npsi);
        const TermDev T = P.terms[it];
        const uint8_t *lev = P.lev + T.lev_off;
        __syncthreads();
        if (threadIdx.x == 0) {
            int str = 1;
            for (int k = 0; k < D; ++k) {
                const int i = k * (P.LG + 1) + lev[k];
                s_tnq[k] = s_nq_of[i]; s_tnb[k] = s_nb_of[i];
                s_oB[k] = s_offB[i];   s_oG[k] = s_offG[i];
                s_str[k] = str; s_mgs[k] = magic_of(str); s_mgn[k] = magic_of(s_nq_of[i]);
                str *= s_nq_of[i];
            }
        }
        __syncthreads();
        const int nq = T.nq, nbT = T.nbT;
        const int32_t *mp = P.map + T.map_off;
        const double *x = psi + (long long)ip * nvec;
        double *y = Hpsi + (long long)ip * nvec;
        for (int j = threadIdx.x; j < nbT; j += blockDim.x) {
            const int m = mp[j];
            for (int c = 0; c < nb0; ++c)
                bufA[c * nbT + j] = (m > 0) ? __ldg(x + (long long)c * P.nb + (m - 1)) : 0.0;
        }
        __syncthreads();
        double *cur = bufA, *oth = bufB;
        {   // B -> G
            int left = 1, right = nbT * nb0;
            for (int k = 0; k < D; ++k) {
                const int nbk = s_tnb[k], nqk = s_tnq[k];
                right /= nbk;
                if (nbk == 1 && nqk == 1) {
                    const double s = __ldg(P.B + s_oB[k]);
                    const int total = left * right;
                    for (int o = threadIdx.x; o < total; o += blockDim.x) cur[o] *= s;
                } else {
                    mode_product(P.B + s_oB[k], nqk, nbk, cur, oth, left, right, magic_of(left), magic_of(left * nqk));
                    double *t = cur; cur = oth; oth = t;
                }
                left *= nqk;
                __syncthreads();
            }
        }
        const double *Jq = O.Jac + T.grid_off, *Sq = O.sq + T.grid_off;
        // derivative of a grid array along the mode owning active coordinate a, at point q
        auto deriv = [&](const double *arr, int a, int q) {
            const int k = O.act_mode[a];
            const double *M = P.D1 + s_oG[k];
            const int nk = s_tnq[k], st = s_str[k];
            const int qd = mdiv(q, s_mgs[k]);                        // no integer division per point
            const int qk = qd - mdiv(qd, s_mgn[k]) * nk;
            const int base = q - qk * st;
            return dot_strided(M + qk, nk, arr + base, st, nk);
        };
        for (int c = 0; c < nb0; ++c) {
            // phi = psi_c * sq  -> chi buffer
            for (int q = threadIdx.x; q < nq; q += blockDim.x) chi[q] = cur[c * nq + q] * __ldg(Sq + q);
            __syncthreads();
            for (int a = 0; a < n; ++a)
                for (int q = threadIdx.x; q < nq; q += blockDim.x) sR[(size_t)a * O.nqmax + q] = deriv(chi, a, q);
            __syncthreads();
            // sum_i d_i [Jac sum_j G^{ji} d_j phi] accumulates in the second buffer (every thread owns its points), so the term
            // size is bounded by shared memory only
            for (int q = threadIdx.x; q < nq; q += blockDim.x) oth[c * nq + q] = 0.0;
            for (int i = 0; i < n; ++i) {
                for (int q = threadIdx.x; q < nq; q += blockDim.x) {
                    double s = 0.0;
                    for (int j = 0; j < n; ++j) {
                        const int comp = O.sym ? (j <= i ? j + i * (i + 1) / 2 : i + j * (j + 1) / 2) : j + n * i;
                        s = fma(__ldg(O.GG + (long long)comp * P.NQ_local + T.grid_off + q), sR[(size_t)j * O.nqmax + q], s);
                    }
                    chi[q] = s * __ldg(Jq + q);
                }
                __syncthreads();
                for (int q = threadIdx.x; q < nq; q += blockDim.x) oth[c * nq + q] += deriv(chi, i, q);
                __syncthreads();
            }
            {
                for (int q = threadIdx.x; q < nq; q += blockDim.x) {
                    double r = -0.5 * oth[c * nq + q] / (__ldg(Jq + q) * __ldg(Sq + q));
                    if (O.has_V)
                        for (int j = 0; j < nb0; ++j)
                            r = fma(__ldg(O.V + (long long)(c + nb0 * j) * P.NQ_local + T.grid_off + q), cur[j * nq + q], r);
                    oth[c * nq + q] = r;
                }
            }
        }
        __syncthreads();
        { double *t = cur; cur = oth; oth = t; }
        {   // G -> B
            int left = 1, right = nq * nb0;
            for (int k = 0; k < D; ++k) {
                const int nbk = s_tnb[k], nqk = s_tnq[k];
                right /= nqk;
                if (nbk == 1 && nqk == 1) {
                    const double s = __ldg(P.BTw + s_oB[k]);
                    const int total = left * right;
                    for (int o = threadIdx.x; o < total; o += blockDim.x) cur[o] *= s;
                } else {
                    mode_product(P.BTw + s_oB[k], nbk, nqk, cur, oth, left, right, magic_of(left), magic_of(left * nbk));
                    double *t = cur; cur = oth; oth = t;
                }
                left *= nbk;
                __syncthreads();
            }
        }
        for (int j = threadIdx.x; j < nbT; j += blockDim.x) {
            const int m = mp[j];
            if (m > 0)
                for (int c = 0; c < nb0; ++c)
                    atomicAdd(y + (long long)c * P.nb + (m - 1), T.weight * cur[c * nbT + j]);
        }
    }
}

} // namespace evr
