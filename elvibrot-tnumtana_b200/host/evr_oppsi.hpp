// evr_oppsi.hpp -- C++ host-side mirror of the reference's operator-action interface (mod_OpPsi) over the
// C-ABI of include/evr_sg4.h.  Same names, argument meaning and error behaviour as the Fortran it stands for:
//
//   param_psi                      TYPE param_psi   Source_ElVibRot/sub_WP/sub_module_psi_set_alloc.f90 (RvecB / CvecB, cplx, symab)
//   param_Op                       TYPE param_Op    Source_ElVibRot/sub_Operator/sub_module_SetOp.f90 (only what the SG4 action reads)
//   sub_TabOpPsi_FOR_SGtype4       sub_Operator/sub_OpPsi_SG4.f90:678-979
//   sub_OpPsi / sub_TabOpPsi       sub_Operator/sub_OpPsi.f90:175-417 / :701-883 (SparseGrid_type == 4 branch)
//   sub_scaledOpPsi                sub_Operator/sub_OpPsi.f90:2823-2866
//
// The reference STOPs with a message on error; this mirror throws evr::Stop carrying the same message.
// Header-only; link with -levr_sg4.  No CPU fallback exists behind these calls.
#pragma once
#include <complex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/evr_sg4.h"

namespace evr {

struct Stop : std::runtime_error { using std::runtime_error::runtime_error; };

struct param_psi {
    std::vector<double> RvecB;                    // packed basis representation, real
    std::vector<std::complex<double>> CvecB;      // ... complex
    bool cplx = false;
    int symab = -1;
};

struct param_Op {
    evr_sg4_plan *plan = nullptr;                 // device-resident cache of para_Op%BasisnD / OpGrid (first call)
    long long nb = 0;                             // packed basis size per channel
    int nb0 = 1;                                  // channels (nb_bie)
    int symab = -1;
    bool cplx = false;
    long long nb_OpPsi = 0;                       // counter bumped by sub_OpPsi (sub_OpPsi.f90:258)
    bool Op_Transfo = false;                      // para_ReadOp%Op_Transfo / E0_Transfo (sub_OpPsi.f90:768-775)
    double E0_Transfo = 0.0;
};

inline void check(int rc, const char *where)
{
    if (rc != 0) throw Stop(std::string(" ERROR in ") + where + ": " + evr_sg4_last_error());
}

// OpPsi(:) = Op Psi(:), real psi only (the routine refuses complex input, sub_OpPsi_SG4.f90:744-749)
inline void sub_TabOpPsi_FOR_SGtype4(const std::vector<param_psi> &Psi, std::vector<param_psi> &OpPsi, param_Op &para_Op)
{
    if (Psi.empty()) throw Stop(" ERROR in sub_TabOpPsi_FOR_SGtype4: size(Psi) = 0");
    if (Psi[0].cplx) throw Stop(" ERROR in sub_TabOpPsi_FOR_SGtype4: Psi(1) is complex");
    const size_t n = (size_t)para_Op.nb * para_Op.nb0;
    std::vector<double> x(n * Psi.size()), y(n * Psi.size());
    for (size_t i = 0; i < Psi.size(); ++i) {
        if (Psi[i].RvecB.size() != n) throw Stop(" ERROR in sub_TabOpPsi_FOR_SGtype4: wrong size of RvecB");
        std::copy(Psi[i].RvecB.begin(), Psi[i].RvecB.end(), x.begin() + i * n);
    }
    check(evr_sg4_apply(para_Op.plan, (int)Psi.size(), x.data(), y.data()), "sub_TabOpPsi_FOR_SGtype4");
    OpPsi.assign(Psi.size(), param_psi());
    for (size_t i = 0; i < Psi.size(); ++i) {
        OpPsi[i].RvecB.assign(y.begin() + i * n, y.begin() + (i + 1) * n);
        OpPsi[i].cplx = false;
        OpPsi[i].symab = Psi[i].symab;               // Calc_symab1_EOR_symab2 with a totally symmetric H
    }
}

// complex psi = two real right-hand sides (RCPsi = Psi, sub_OpPsi.f90:392-407)
inline void sub_OpPsi(const param_psi &Psi, param_psi &OpPsi, param_Op &para_Op)
{
    para_Op.nb_OpPsi += 1;
    if (Psi.cplx) {
        std::vector<param_psi> RC(2), RCO;
        RC[0].RvecB.resize(Psi.CvecB.size()); RC[1].RvecB.resize(Psi.CvecB.size());
        for (size_t i = 0; i < Psi.CvecB.size(); ++i) { RC[0].RvecB[i] = Psi.CvecB[i].real(); RC[1].RvecB[i] = Psi.CvecB[i].imag(); }
        sub_TabOpPsi_FOR_SGtype4(RC, RCO, para_Op);
        OpPsi.CvecB.resize(Psi.CvecB.size());
        for (size_t i = 0; i < Psi.CvecB.size(); ++i) OpPsi.CvecB[i] = {RCO[0].RvecB[i], RCO[1].RvecB[i]};
        OpPsi.RvecB.clear(); OpPsi.cplx = true;
    } else {
        std::vector<param_psi> in(1, Psi), out;
        sub_TabOpPsi_FOR_SGtype4(in, out, para_Op);
        OpPsi.RvecB = std::move(out[0].RvecB); OpPsi.CvecB.clear(); OpPsi.cplx = false;
    }
    OpPsi.symab = Psi.symab;
}

// OpPsi <- (OpPsi - E0 Psi)/Esc
inline void sub_scaledOpPsi(const param_psi &Psi, param_psi &OpPsi, double E0, double Esc)
{
    if (Psi.cplx) for (size_t i = 0; i < Psi.CvecB.size(); ++i) OpPsi.CvecB[i] = (OpPsi.CvecB[i] - E0 * Psi.CvecB[i]) / Esc;
    else for (size_t i = 0; i < Psi.RvecB.size(); ++i) OpPsi.RvecB[i] = (OpPsi.RvecB[i] - E0 * Psi.RvecB[i]) / Esc;
}

// TransfoOp with para_Op%para_ReadOp%Op_Transfo: OpPsi = (H - E0_Transfo)(H - E0_Transfo) Psi, vector by vector (:768-775)
inline void sub_TabOpPsi(const std::vector<param_psi> &TabPsi, std::vector<param_psi> &TabOpPsi, param_Op &para_Op,
                         bool TransfoOp = false)
{
    if (TransfoOp && para_Op.Op_Transfo) {
        TabOpPsi.assign(TabPsi.size(), param_psi());
        for (size_t i = 0; i < TabPsi.size(); ++i) {
            param_psi tmp;
            sub_OpPsi(TabPsi[i], tmp, para_Op);
            sub_scaledOpPsi(TabPsi[i], tmp, para_Op.E0_Transfo, 1.0);
            sub_OpPsi(tmp, TabOpPsi[i], para_Op);
            sub_scaledOpPsi(tmp, TabOpPsi[i], para_Op.E0_Transfo, 1.0);
        }
        return;
    }
    bool any_cplx = false;
    for (const auto &p : TabPsi) any_cplx = any_cplx || p.cplx;
    if (!any_cplx) { para_Op.nb_OpPsi += (long long)TabPsi.size(); sub_TabOpPsi_FOR_SGtype4(TabPsi, TabOpPsi, para_Op); return; }
    TabOpPsi.assign(TabPsi.size(), param_psi());
    for (size_t i = 0; i < TabPsi.size(); ++i) sub_OpPsi(TabPsi[i], TabOpPsi[i], para_Op);
}

} // namespace evr
