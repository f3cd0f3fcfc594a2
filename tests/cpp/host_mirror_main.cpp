// Test driver of the C++ host mirror (elvibrot-tnumtana_b200/host/evr_oppsi.hpp).
// Reads a problem dumped by tests/test_gpu_parity.py::test_cpp_host_mirror (flat little-endian arrays, the
// exact arguments of evr_sg4_plan_create / evr_sg4_plan_set_op), applies H through sub_OpPsi / sub_TabOpPsi
// and writes the results back.  usage: host_mirror_main <in.bin> <out.bin> [ndev]
// ndev > 1: evr_sg4_set_devices(ndev) first -- the same program then drives ndev GPUs (term ranges per device, NVLink
// reduction, slice-wise host copies) through include/evr_sg4.h only.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../elvibrot-tnumtana_b200/host/evr_oppsi.hpp"

template <class T> static std::vector<T> rd(FILE *f) {
    long long n = 0;
    if (fread(&n, sizeof(n), 1, f) != 1) { fprintf(stderr, "short read\n"); exit(2); }
    std::vector<T> v((size_t)n);
    if (n && fread(v.data(), sizeof(T), (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
    return v;
}

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    auto hdr = rd<long long>(f);           // D, nb_SG, nb0, nb, LG, type_Op, nb_Term, npsi_real, ncplx
    const int D = (int)hdr[0], nb_SG = (int)hdr[1], nb0 = (int)hdr[2], LG = (int)hdr[4], type_Op = (int)hdr[5], nb_Term = (int)hdr[6];
    const long long nb = hdr[3];
    const int npsi = (int)hdr[7], ncplx = (int)hdr[8];
    auto tab_l = rd<int32_t>(f); auto W = rd<double>(f); auto tnq = rd<int32_t>(f); auto tnb = rd<int32_t>(f);
    auto map = rd<int32_t>(f); auto nq_of = rd<int32_t>(f); auto nb_of = rd<int32_t>(f);
    auto B = rd<double>(f); auto BTw = rd<double>(f); auto D1 = rd<double>(f); auto D2 = rd<double>(f);
    auto term_mode = rd<int32_t>(f); auto gz = rd<uint8_t>(f); auto gc = rd<uint8_t>(f); auto mc = rd<double>(f);
    std::vector<std::vector<double>> grids(nb_Term);
    std::vector<const double *> gp(nb_Term, nullptr);
    for (int it = 0; it < nb_Term; ++it) { grids[it] = rd<double>(f); if (!grids[it].empty()) gp[it] = grids[it].data(); }
    auto psi = rd<double>(f);               // npsi real vectors
    auto cpsi = rd<double>(f);              // ncplx complex vectors (re,im interleaved)
    fclose(f);

    evr::param_Op H;
    H.nb = nb; H.nb0 = nb0;
    const int ndev = argc > 3 ? atoi(argv[3]) : 1;
    try {
        if (ndev > 1) evr::check(evr_sg4_set_devices(ndev), "set_devices");
        evr::check(evr_sg4_plan_create(&H.plan, -1, D, nb_SG, nb0, nb, LG, tab_l.data(), W.data(), tnq.data(), tnb.data(), map.data(),
                                       nq_of.data(), nb_of.data(), B.data(), BTw.data(), D1.data(), D2.data(), 0, nb_SG), "plan_create");
        evr::check(evr_sg4_plan_set_op(H.plan, type_Op, nb_Term, term_mode.data(), gz.data(), gc.data(), mc.data(), gp.data()), "plan_set_op");
        const size_t n = (size_t)nb * nb0;
        std::vector<evr::param_psi> Tab(npsi), TabH;
        for (int i = 0; i < npsi; ++i) Tab[i].RvecB.assign(psi.begin() + i * n, psi.begin() + (i + 1) * n);
        if (npsi) evr::sub_TabOpPsi(Tab, TabH, H);                       // block of real vectors (Davidson)
        std::vector<evr::param_psi> CH(ncplx);
        for (int i = 0; i < ncplx; ++i) {                                 // complex wave packets (propagation)
            evr::param_psi P; P.cplx = true; P.CvecB.resize(n);
            for (size_t k = 0; k < n; ++k) P.CvecB[k] = {cpsi[2 * (i * n + k)], cpsi[2 * (i * n + k) + 1]};
            evr::sub_OpPsi(P, CH[i], H);
        }
        // error behaviour: empty table and complex input to the SG4 routine must "STOP"
        int stops = 0;
        try { std::vector<evr::param_psi> e, o; evr::sub_TabOpPsi_FOR_SGtype4(e, o, H); } catch (const evr::Stop &) { ++stops; }
        try { std::vector<evr::param_psi> e(1), o; e[0].cplx = true; evr::sub_TabOpPsi_FOR_SGtype4(e, o, H); } catch (const evr::Stop &) { ++stops; }
        const long long n_OpPsi = H.nb_OpPsi;
        // Op_Transfo / TransfoOp branch of sub_TabOpPsi (sub_OpPsi.f90:768-775): (H - E0)(H - E0) of the first real vector
        std::vector<evr::param_psi> T2;
        if (npsi) {
            H.Op_Transfo = true; H.E0_Transfo = 0.37;
            std::vector<evr::param_psi> one(1, Tab[0]);
            evr::sub_TabOpPsi(one, T2, H, true);
        }
        FILE *g = fopen(argv[2], "wb");
        long long cnt = stops; fwrite(&cnt, sizeof(cnt), 1, g);
        cnt = n_OpPsi; fwrite(&cnt, sizeof(cnt), 1, g);
        if (ndev > 1 && evr_sg4_plan_info(H.plan, EVR_INFO_DEVICES) != ndev) { fprintf(stderr, "plan does not span %d devices\n", ndev); return 4; }
        for (int i = 0; i < npsi; ++i) fwrite(TabH[i].RvecB.data(), sizeof(double), n, g);
        for (int i = 0; i < ncplx; ++i) fwrite(CH[i].CvecB.data(), sizeof(double), 2 * n, g);
        if (npsi) fwrite(T2[0].RvecB.data(), sizeof(double), n, g);
        fclose(g);
        evr_sg4_plan_destroy(&H.plan);
    } catch (const evr::Stop &e) {
        fprintf(stderr, "%s\n", e.what());
        return 3;
    }
    return 0;
}
