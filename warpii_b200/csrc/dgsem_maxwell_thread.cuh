// Per-thread halves of the stand-alone field kernel (dgsem_maxwell_kernel.cu): what a thread does before and after the
// block barrier.  Perfectly hyperbolic Maxwell (PHM) fluxes for the 8 field components [Ex,Ey,Ez,Bx,By,Bz,phi,psi]
// (five_moment.h:123-138), north_star kernel 4 / BASELINE config 5.  The reference only allocates these components; the
// system evolved here is the one SURVEY.md 8(c) names (parity unpinned upstream):
//   dF/dt + sum_d d f_d(F)/dx_d = S,   f_d(E) = -c^2 (e_d x B) + chi c^2 phi e_d,  f_d(B) = e_d x E + gamma psi e_d,
//   f_d(phi) = chi E_d,  f_d(psi) = gamma c^2 B_d,   S = (-J/eps0, 0, chi rho_c/eps0, 0)
// collocated DGSEM on the fluid's Gauss-Lobatto nodes, Rusanov flux with lambda = c max(1, chi, gamma).
//
// Plain __device__ code over (params, shared memory, tid): wgpu_portable.cuh lets tests/emu/ compile the same source for
// the host and run it thread by thread (test infrastructure; the product has no CPU path).
#pragma once
#include "dgsem_kernels.cuh"

namespace wgpu {

struct MaxwellParams {
    double c2, chi, gam, lam;     // c^2, cleaning speeds (units of c), Rusanov speed c max(1, chi, gamma)
    double inv_eps0;
    double speed_floor;           // max_eig * lam: the field system's constant share of the transport speed
    double omega_factor;          // 5 / Np^2: omega dt <= 0.1 expressed as a speed (dt = 0.5 / (vmax Np^2))
    int32_t sources_on;
};

// f_d(F) for one direction, all 8 components
__device__ __forceinline__ void phm_flux(const int d, const MaxwellParams& M, const double F[8], double f[8]) {
    const int i1 = (d + 1) % 3, i2 = (d + 2) % 3;
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = 0.0;
    // (e_d x B)_{i1} = -B_{i2}, (e_d x B)_{i2} = B_{i1}
    f[i1] = M.c2 * F[3 + i2];
    f[i2] = -M.c2 * F[3 + i1];
    f[3 + i1] = -F[i2];
    f[3 + i2] = F[i1];
    f[d] = M.chi * M.c2 * F[6];
    f[3 + d] = M.gam * F[7];
    f[6] = M.chi * F[d];
    f[7] = M.gam * M.c2 * F[3 + d];
}

template <int DIM, int NP>
struct MGeo {
    static constexpr int NN = ipow_c(NP, DIM), NF = ipow_c(NP, DIM - 1);
    static constexpr int G = (128 / NN) > 0 ? (128 / NN) : 1;   // elements per block
    static constexpr int THREADS = G * NN;
    static constexpr int SMEM_DOUBLE2 = G * 4 * NN;             // component pairs (Ex,Ey) (Ez,Bx) (By,Bz) (phi,psi) per node
};

// what a thread keeps across the block barrier
template <int DIM>
struct MaxwellCarry {
    int64_t e;                // its element
    int le, j;                // element within the block, node within the element
    bool active;              // e < elem_end
    const double* src[DIM];   // where the outside trace of direction d lives (the node itself: no jump)
    int stride[DIM];
    bool jump[DIM];           // the node lies on a face of direction d that has a neighbour
};

// Before the barrier: every DRAM access of the thread is issued here, before any arithmetic: its 8 field values (parked in
// shared memory for the pencil sums) as loads; the old destination (second stage), the species' densities and momenta for
// the current and the updated densities for the plasma frequency as L2 prefetches (the values are loaded in maxwell_post
// where they are used, from L2, and hold no registers across the barrier and the flux sums).  ncu, round 2: the first
// version, with four to six dependent load rounds, spent 5-10 stall samples per issued instruction on long_scoreboard.
template <int DIM, int NP>
__device__ __forceinline__ void maxwell_pre(const StageParams& P, double2* sF, const int tid, const int64_t block, MaxwellCarry<DIM>& c) {
    using GEO = MGeo<DIM, NP>;
    constexpr int NN = GEO::NN, NF = GEO::NF, G = GEO::G;
    c.le = tid / NN;
    c.j = tid - c.le * NN;
    c.e = P.elem_begin + block * G + c.le;
    c.active = c.e < P.elem_end;
    if (!c.active) return;
    const int j = c.j;
    const int64_t e = c.e;
    const int nf0 = 5 * P.nsp;   // first field component
    const int i0 = j % NP, i1 = (DIM > 1) ? (j / NP) % NP : 0, i2 = (DIM > 2) ? j / (NP * NP) : 0;
    const int idx[3] = {i0, i1, i2};
    const size_t own = ((size_t)e * P.nc + nf0) * NN + j;
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        const int jd = idx[d];
        const bool on_face = (jd == 0 || jd == NP - 1);
        const int side = (jd == 0) ? 0 : 1;   // (Np >= 2: a node is on at most one face per direction)
        const int v = on_face ? P.nbr[(size_t)e * (2 * DIM) + 2 * d + side] : -1;
        const int t = face_node_index<DIM, NP>(d, i0, i1, i2);
        c.jump[d] = v >= 0;
        c.src[d] = P.u + own;
        c.stride[d] = NN;
        if (v >= P.n_elems) {
            c.src[d] = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + nf0) * NF + t;
            c.stride[d] = NF;
        } else if (v >= 0) {
            c.src[d] = P.u + ((size_t)v * P.nc + nf0) * NN + node_of_face_node<DIM, NP>(d, 1 - side, t);
        }
    }
    double F[8];
#pragma unroll
    for (int k = 0; k < 8; k++) F[k] = P.u[own + (size_t)k * NN];
#if !WGPU_HOST_EMU
    if (P.mode == 2 || (P.mode == 0 && P.beta != 0.0)) {
        const double* const op = (P.mode == 2 ? P.sol_in : P.dst) + own;
#pragma unroll
        for (int k = 0; k < 8; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(op + (size_t)k * NN));
    }
    if (P.src_on) {
        const bool want_speed = P.vmax && P.mode == 0;
        for (int sp = 0; sp < P.nsp; sp++) {
            const size_t so = ((size_t)e * P.nc + 5 * sp) * NN + j;
#pragma unroll
            for (int k = 0; k < 4; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.u + so + (size_t)k * NN));
            if (want_speed) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.dst + so));
        }
    }
#endif
#pragma unroll
    for (int k = 0; k < 4; k++) sF[(c.le * 4 + k) * NN + j] = make_double2(F[2 * k], F[2 * k + 1]);
}

// After the barrier: volume and face terms from shared memory and the neighbours' traces (L2 hits, loaded at the head of each
// direction and in flight during its volume term), sources, stage update, store; returns the thread's share of the
// transport speed in the stage that fuses the CFL reduction (0 otherwise).
template <int DIM, int NP>
__device__ __forceinline__ double maxwell_post(const StageParams& P, const MaxwellParams& M, const double2* sF, const double dt,
                                               const MaxwellCarry<DIM>& c) {
    using GEO = MGeo<DIM, NP>;
    constexpr int NN = GEO::NN;
    if (!c.active) return 0.0;
    const int j = c.j, le = c.le;
    const int64_t e = c.e;
    const int nf0 = 5 * P.nsp;
    const bool want_speed = P.vmax && P.mode == 0;
    const int i0 = j % NP, i1 = (DIM > 1) ? (j / NP) % NP : 0, i2 = (DIM > 2) ? j / (NP * NP) : 0;
    const int idx[3] = {i0, i1, i2};
    double rate[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < DIM; d++) {
        const int st = stride_of(NP, d), jd = idx[d];
        double Fod[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (c.jump[d]) {
#pragma unroll
            for (int k = 0; k < 8; k++) Fod[k] = c.src[d][(size_t)k * c.stride[d]];
        }
        // volume: -(1/h_d) sum_l D[j_d][l] f_d(F_l) = -(1/h_d) f_d(sum_l D[j_d][l] F_l): the flux is linear
        double dF_[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int l = 0; l < NP; l++) {
            const int q = j + (l - jd) * st;
            const double w = P.T.D[jd * NP + l];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double2 v = sF[(le * 4 + k) * NN + q];
                dF_[2 * k] = fma(w, v.x, dF_[2 * k]);
                dF_[2 * k + 1] = fma(w, v.y, dF_[2 * k + 1]);
            }
        }
        double acc[8];
        phm_flux(d, M, dF_, acc);
#pragma unroll
        for (int k = 0; k < 8; k++) rate[k] = fma(-P.inv_h[d], acc[k], rate[k]);
        // faces: (f(F_m).n - f*) / (h_d w_0) = (lambda dF - sgn f_d(dF)) / (2 h_d w_0), dF = F_p - F_m; a domain boundary
        // has outside state = inside state, i.e. no jump
        if (c.jump[d]) {
            double dF[8], fn[8];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const double2 own = sF[(le * 4 + k) * NN + j];
                dF[2 * k] = Fod[2 * k] - own.x;
                dF[2 * k + 1] = Fod[2 * k + 1] - own.y;
            }
            phm_flux(d, M, dF, fn);
            const double cf = 0.5 * P.inv_hw[d], cl = cf * M.lam, cs = (jd == 0) ? cf : -cf;
#pragma unroll
            for (int k = 0; k < 8; k++) rate[k] = fma(cs, fn[k], fma(cl, dF[k], rate[k]));
        }
    }
    // sources: -J/eps0 on E, chi rho_c/eps0 on phi (the same sums, in the same order, as the fluid kernels' field phase)
    double wp2 = 0.0, qmax = 0.0;
    if (P.src_on) {
        double Jx = 0.0, Jy = 0.0, Jz = 0.0, rc = 0.0;
        for (int sp = 0; sp < P.nsp; sp++) {
            const size_t so = ((size_t)e * P.nc + 5 * sp) * NN + j;
            const double qm = P.qm[sp];
            rc += qm * P.u[so];
            Jx += qm * P.u[so + NN];
            Jy += qm * P.u[so + 2 * (size_t)NN];
            Jz += qm * P.u[so + 3 * (size_t)NN];
            if (want_speed) {
                // plasma frequency of the UPDATED state (the fluid kernel of this range has already written dst)
                wp2 += qm * qm * P.dst[so] * P.inv_eps0;
                qmax = fmax(qmax, fabs(qm));
            }
        }
        rate[0] += -Jx * P.inv_eps0;
        rate[1] += -Jy * P.inv_eps0;
        rate[2] += -Jz * P.inv_eps0;
        rate[6] += P.chi * rc * P.inv_eps0;
    }
    double Fn[8], F[8], old[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double2 own = sF[(le * 4 + k) * NN + j];
        F[2 * k] = own.x;
        F[2 * k + 1] = own.y;
    }
    if (P.mode == 2 || (P.mode == 0 && P.beta != 0.0)) {
        const double* const op = (P.mode == 2 ? P.sol_in : P.dst) + ((size_t)e * P.nc + nf0) * NN + j;
#pragma unroll
        for (int k = 0; k < 8; k++) old[k] = op[(size_t)k * NN];
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const size_t off = ((size_t)e * P.nc + nf0 + k) * NN + j;
        double v;
        if (P.mode == 1) v = rate[k];
        else if (P.mode == 2) {
            v = fma(P.a, rate[k], old[k]);
            if (P.beta != 0.0) P.dst2[off] = fma(P.beta, rate[k], old[k]);
        }
        else if (P.beta == 0.0) v = P.a * (F[k] + dt * rate[k]);
        else v = P.beta * old[k] + P.a * (F[k] + dt * rate[k]);
        P.dst[off] = v;
        Fn[k] = v;
    }
    double vmax_local = 0.0;
    if (want_speed) {
        vmax_local = M.speed_floor;
        if (P.src_on) {
            const double b2 = Fn[3] * Fn[3] + Fn[4] * Fn[4] + Fn[5] * Fn[5];
            const double omega = fmax(sqrt(wp2), qmax * sqrt(b2));
            vmax_local = nan_max(vmax_local, M.omega_factor * omega);
        }
    }
    return vmax_local;
}

}  // namespace wgpu
