// tests/emu/pencil_emu.cc -- TEST INFRASTRUCTURE ONLY.
//
// Executes the per-thread phase functions of the pencil stage kernel (warpii_b200/csrc/dgsem_pencil_stage.cuh) on the host,
// thread by thread, one phase after the other: a block barrier separates the phases on the device, so running every thread
// of a block through phase k before any thread enters phase k+1 is one legal schedule of the kernel.  The CPU-only test
// tier uses it to check the kernel's indexing (pencil ownership, halo ends, partial patches, ghost traces, the
// troubled-cell correction, the stage-update modes) against the oracle before a GPU is spent on it.  It is compiled by a
// plain host compiler with wgpu_portable.cuh's shims (MUFU seeds replaced by single-precision ones), so its results agree
// with the device's to round-off, not bit for bit.  Nothing under warpii_b200/ builds, links or loads this file; the
// product has no CPU path.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../warpii_b200/csrc/dgsem_maxwell_thread.cuh"
#include "../../warpii_b200/csrc/dgsem_pencil_stage.cuh"
#include "../../warpii_b200/host/reference_element.hpp"

using namespace wgpu;

namespace {

template <int DIM, int NP>
double run(const StageParams& P) {
    using G = PGeo<DIM, NP>;
    const int64_t n = P.elem_end - P.elem_begin;
    const int64_t blocks = (n + G::E - 1) / G::E;
    std::vector<double> smem(G::SMEM_DOUBLES + 2);
    std::vector<PencilHalo> halo(G::THREADS);
    double vmax = 0.0;
    bool nan_seen = false;
    double* sm = smem.data();
    if ((reinterpret_cast<uintptr_t>(sm) & 15) != 0) sm++;   // 16-byte alignment of the double2 planes
    for (int64_t b = 0; b < blocks; b++) {
        const int64_t e0 = P.elem_begin + b * G::E;
        std::fill(smem.begin(), smem.end(), std::nan(""));   // a read of something no phase wrote poisons the result
        for (int sp = 0; sp < P.nsp; sp++) {
            for (int t = 0; t < G::THREADS; t++) pencil_phase0<DIM, NP>(P, sm, t, e0, sp, halo[t]);
#if WGPU_PENCIL_RTDIR
            for (int ph = 1; ph <= DIM; ph++)
                for (int t = 0; t < G::THREADS; t++) {
                    const double v = pencil_phase_flux_rt<DIM, NP>(P, sm, t, e0, sp, P.dt, ph < DIM ? ph : 0, halo[t]);
                    if (v != v) nan_seen = true;
                    vmax = std::max(vmax, v);
                }
#else
            for (int t = 0; t < G::THREADS; t++) pencil_phase_mid<DIM, NP, 1>(P, sm, t, e0, sp, halo[t]);
            if (DIM == 3)
                for (int t = 0; t < G::THREADS; t++) pencil_phase_mid<DIM, NP, DIM - 1>(P, sm, t, e0, sp, halo[t]);
            for (int t = 0; t < G::THREADS; t++) {
                const double v = pencil_phase_final<DIM, NP>(P, sm, t, e0, sp, P.dt, halo[t]);
                if (v != v) nan_seen = true;
                vmax = std::max(vmax, v);
            }
#endif
        }
        if (P.mx_on && P.nc >= 5 * P.nsp + 8) {
            std::vector<FieldHalo> fh(G::THREADS);
            for (int t = 0; t < G::THREADS; t++) field_phase0<DIM, NP>(P, sm, t, e0, fh[t]);
#if WGPU_PENCIL_RTDIR
            for (int ph = 1; ph <= DIM; ph++)
                for (int t = 0; t < G::THREADS; t++) {
                    const double v = field_phase_flux_rt<DIM, NP>(P, sm, t, e0, P.dt, ph < DIM ? ph : 0, fh[t]);
                    if (v != v) nan_seen = true;
                    vmax = std::max(vmax, v);
                }
#else
            for (int t = 0; t < G::THREADS; t++) field_phase_mid<DIM, NP, 1>(P, sm, t, e0, fh[t]);
            if (DIM == 3)
                for (int t = 0; t < G::THREADS; t++) field_phase_mid<DIM, NP, DIM - 1>(P, sm, t, e0, fh[t]);
            for (int t = 0; t < G::THREADS; t++) {
                const double v = field_phase_final<DIM, NP>(P, sm, t, e0, P.dt, fh[t]);
                if (v != v) nan_seen = true;
                vmax = std::max(vmax, v);
            }
#endif
        } else {
            for (int t = 0; t < G::THREADS; t++) pencil_phase_fields<DIM, NP>(P, t, e0, P.dt);
        }
    }
    return nan_seen ? std::nan("") : vmax;
}

}  // namespace

// The stand-alone field kernel (warpii_b200/csrc/dgsem_maxwell_thread.cuh, launched by dgsem_maxwell_kernel.cu): every
// thread of a block through maxwell_pre, then - the block barrier - every thread through maxwell_post.
template <int DIM, int NP>
double run_maxwell(const StageParams& P, const MaxwellParams& M) {
    using GEO = MGeo<DIM, NP>;
    const int64_t n = P.elem_end - P.elem_begin;
    const int64_t blocks = (n + GEO::G - 1) / GEO::G;
    std::vector<double2> smem(GEO::SMEM_DOUBLE2);
    std::vector<MaxwellCarry<DIM>> carry(GEO::THREADS);
    double vmax = 0.0;
    bool nan_seen = false;
    for (int64_t b = 0; b < blocks; b++) {
        for (auto& v : smem) v = make_double2(std::nan(""), std::nan(""));
        for (int t = 0; t < GEO::THREADS; t++) maxwell_pre<DIM, NP>(P, smem.data(), t, b, carry[t]);
        for (int t = 0; t < GEO::THREADS; t++) {
            const double v = maxwell_post<DIM, NP>(P, M, smem.data(), P.dt, carry[t]);
            if (v != v) nan_seen = true;
            vmax = std::max(vmax, v);
        }
    }
    return nan_seen ? std::nan("") : vmax;
}

int g_field_kernel = 0;

extern "C" {

// 1: the next emu_pencil_stage calls run the stand-alone field kernel (which updates the 8 field components only) instead
// of the pencil stage kernel; 0: back to the pencil kernel
void emu_select_field_kernel(int on) { g_field_kernel = on; }

int emu_pencil_available(int dim, int np) { return (dim == 2 || dim == 3) && np >= 2 && np <= 6; }
int emu_pencil_patch_elems(int dim, int np) { return pencil_elems(dim, np); }

// One launch of the stage kernel over elements [elem_begin, elem_end).  Arguments mirror StageParams / stage_params()
// of warpii_gpu.cu; h[d] are the element sizes.  Returns 0, or 1 for an unsupported (dim, np).
int emu_pencil_stage(int dim, int np, int64_t n_elems, int64_t elem_begin, int64_t elem_end, int nc, int nsp, const double* u,
                     double* dst, const int32_t* nbr, const double* ghost, const double* bres, double* alpha_out, double* vmax_out,
                     int mode, double gamma, double dt, double a, double beta, const double* h, const double* sol_in, double* dst2,
                     int src_on, double epsilon0, double chi, const double* qm, int mx_on, double light_speed, double mx_chi,
                     double mx_gamma) {
    if (!emu_pencil_available(dim, np)) return 1;
    StageParams P;
    std::memset(&P, 0, sizeof P);
    warpii_b200::ReferenceElement re(np - 1);
    for (int i = 0; i < np * np; i++) { P.T.D[i] = re.D[i]; P.T.V[i] = re.V[i]; }
    for (int i = 0; i < np; i++) P.T.w[i] = re.w[i];
    double inv_h[3] = {1, 1, 1};
    for (int d = 0; d < dim; d++) {
        inv_h[d] = 1.0 / h[d];
        P.inv_h[d] = inv_h[d];
        P.inv_hw[d] = 1.0 / (h[d] * re.w[0]);
    }
    {   // fluid_flux_es_dgsem_operator.h:487-502 (as warpii_gpu_create does)
        double ev[3] = {1, 1, 1};
        for (int it = 0; it < 5; it++) {
            double nrm = 0;
            for (int d = 0; d < dim; d++) { ev[d] = inv_h[d] * (inv_h[d] * ev[d]); nrm = std::fmax(nrm, std::fabs(ev[d])); }
            for (int d = 0; d < dim; d++) ev[d] /= nrm;
        }
        double num = 0, den = 0;
        for (int d = 0; d < dim; d++) { const double jv = inv_h[d] * ev[d]; num += jv * jv; den += ev[d] * ev[d]; }
        P.max_eig = std::sqrt(num / den);
    }
    P.ind_T = 0.5 * std::pow(10.0, -1.8 * std::pow((double)np, 0.25));
    P.ind_sT = 9.21024 / P.ind_T;
    P.u = u; P.dst = dst; P.nbr = nbr; P.ghost = ghost; P.bres = bres; P.alpha_out = alpha_out;
    unsigned long long vslot = 0;
    P.vmax = vmax_out ? &vslot : nullptr;
    P.elem_begin = elem_begin; P.elem_end = elem_end; P.n_elems = n_elems;
    P.nc = nc; P.nsp = nsp; P.ncf = 5 * nsp; P.fields_skip = 0; P.mode = mode;
    P.sol_in = sol_in; P.dst2 = dst2;
    P.gamma = gamma; P.dt = dt; P.a = a; P.beta = beta;
    P.hig = 0.5 / (gamma - 1.0);
    P.src_on = src_on; P.inv_eps0 = src_on ? 1.0 / epsilon0 : 1.0; P.chi = chi; P.qm = qm;
    if (mx_on == 2) {   // the field system is evolved by the stand-alone field kernel: the fluid kernel skips the field components
        P.fields_skip = 1; P.ncf = nc;
    } else if (mx_on) {   // as stage_params() of warpii_gpu.cu
        const double big = std::max(1.0, std::max(mx_chi, mx_gamma));
        P.mx_on = 1; P.fields_skip = 1; P.ncf = nc;
        P.mx_c2 = light_speed * light_speed; P.mx_chi = mx_chi; P.mx_gam = mx_gamma; P.mx_lam = light_speed * big;
        P.mx_floor = P.max_eig * P.mx_lam; P.mx_omega_factor = 5.0 / (double)(np * np);
    }
    double vmax = 0.0;
    if (g_field_kernel) {
        if (!mx_on) return 1;
        // as make_params() of dgsem_maxwell_kernel.cu; the fluid kernels leave the field components alone
        MaxwellParams M;
        M.c2 = P.mx_c2; M.chi = P.mx_chi; M.gam = P.mx_gam; M.lam = P.mx_lam; M.inv_eps0 = P.inv_eps0;
        M.speed_floor = P.mx_floor; M.omega_factor = P.mx_omega_factor; M.sources_on = src_on ? 1 : 0;
        P.mx_on = 0;
#define CALLM(D_, N_) vmax = run_maxwell<D_, N_>(P, M)
        if (dim == 2) {
            switch (np) { case 2: CALLM(2, 2); break; case 3: CALLM(2, 3); break; case 4: CALLM(2, 4); break; case 5: CALLM(2, 5); break; case 6: CALLM(2, 6); break; }
        } else {
            switch (np) { case 2: CALLM(3, 2); break; case 3: CALLM(3, 3); break; case 4: CALLM(3, 4); break; case 5: CALLM(3, 5); break; case 6: CALLM(3, 6); break; }
        }
#undef CALLM
        if (vmax_out) *vmax_out = vmax;
        return 0;
    }
#define CALL(D_, N_) vmax = run<D_, N_>(P)
    if (dim == 2) {
        switch (np) { case 2: CALL(2, 2); break; case 3: CALL(2, 3); break; case 4: CALL(2, 4); break; case 5: CALL(2, 5); break; case 6: CALL(2, 6); break; }
    } else {
        switch (np) { case 2: CALL(3, 2); break; case 3: CALL(3, 3); break; case 4: CALL(3, 4); break; case 5: CALL(3, 5); break; case 6: CALL(3, 6); break; }
    }
#undef CALL
    if (vmax_out) *vmax_out = vmax;
    return 0;
}

}  // extern "C"
