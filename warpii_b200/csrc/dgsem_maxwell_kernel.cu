// Perfectly hyperbolic Maxwell (PHM) fluxes for the 8 field components [Ex,Ey,Ez,Bx,By,Bz,phi,psi] (five_moment.h:123-138),
// north_star kernel 4 / BASELINE config 5.  The reference only allocates these components; the system evolved here is the
// one SURVEY.md 8(c) names (DESIGN.md section 7 states the discretisation; parity unpinned upstream, GPU-vs-CPU-restatement here):
//   dF/dt + sum_d d f_d(F)/dx_d = S,   f_d(E) = -c^2 (e_d x B) + chi c^2 phi e_d,  f_d(B) = e_d x E + gamma psi e_d,
//   f_d(phi) = chi E_d,  f_d(psi) = gamma c^2 B_d,   S = (-J/eps0, 0, chi rho_c/eps0, 0)
// collocated DGSEM on the fluid's Gauss-Lobatto nodes, Rusanov flux with lambda = c max(1, chi, gamma).
//
// One launch updates the field components of an element range with the same stage formula as the fluid kernels (modes
// 0/1/2 of StageParams) and, in the stage that fuses the CFL reduction, adds the field system's share to the transport
// speed: c max(1,chi,gamma) through the metric factor, and the plasma / cyclotron frequency bound.  It runs right after
// the fluid stage kernel of the same range on the same stream (the fluid kernels then skip the field components).
// HBM-bound: 8 components read + written per node plus the species' rho, m for the current (L2 hits: the stage kernel
// has just read them).  One thread per node, the element's fields staged in shared memory for the pencil sums.
#include "dgsem_common.cuh"
#include "dgsem_physics.cuh"

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

#ifndef WGPU_MAXWELL_BLOCKS
#define WGPU_MAXWELL_BLOCKS 6
#endif
#ifndef WGPU_MAXWELL_LATE_OLD
#define WGPU_MAXWELL_LATE_OLD 1   // the old destination is only PREFETCHED (to L2) before the barrier and loaded at the update: 16 registers less (N3D stage 3.82 -> 3.75 ms)
#endif
#ifndef WGPU_MAXWELL_LATE_J
#define WGPU_MAXWELL_LATE_J 1   // the species' values for the current are only prefetched (L2) before the barrier and loaded after the fluxes (N3D stage 3.74 -> 3.64 ms)
#endif
#ifndef WGPU_MAXWELL_PREFETCH_FACES
#define WGPU_MAXWELL_PREFETCH_FACES 0   // 1: the neighbours' traces are loaded before the barrier too (48 more registers)
#endif

template <int DIM, int NP>
struct MGeo {
    static constexpr int NN = ipow_c(NP, DIM), NF = ipow_c(NP, DIM - 1);
    static constexpr int G = (128 / NN) > 0 ? (128 / NN) : 1;   // elements per block
    static constexpr int THREADS = G * NN;
};

// Every DRAM access of the thread is issued before the block barrier and before any arithmetic: its 8 field values, the
// old destination (second stage) and the species' densities and momenta for the current (both as L2 prefetches: the values
// are loaded where they are needed, from L2, and hold no registers across the barrier and the flux sums), the updated densities for the plasma
// frequency, and the neighbours' traces for the (at most DIM) faces the node lies on -- through one unconditional load per
// component and direction whose address falls back to the node itself where there is no face or no neighbour.  One
// DRAM round trip per thread instead of four to six dependent ones (ncu, round 2: the serial version spent 5-10 stall
// samples per issued instruction on long_scoreboard and reached 2.6-3.1 TB/s).
template <int DIM, int NP>
__global__ void __launch_bounds__(MGeo<DIM, NP>::THREADS, (MGeo<DIM, NP>::THREADS <= 128) ? WGPU_MAXWELL_BLOCKS : 2)
    maxwell_kernel(const StageParams P, const MaxwellParams M) {
    using GEO = MGeo<DIM, NP>;
    constexpr int NN = GEO::NN, NF = GEO::NF, G = GEO::G;
    __shared__ double2 sF[G][4][NN];   // component pairs (Ex,Ey) (Ez,Bx) (By,Bz) (phi,psi): 16-byte accesses
    __shared__ double sRed[32];
    const int skip = P.skip_dev ? *P.skip_dev : 0;
    const double dt = P.dt_dev ? *P.dt_dev : P.dt;
    if (skip) return;
    const int tid = threadIdx.x, le = tid / NN, j = tid - le * NN;
    const int64_t e = P.elem_begin + (int64_t)blockIdx.x * G + le;
    const bool active = e < P.elem_end;
    const int nf0 = 5 * P.nsp;   // first field component
    const bool want_speed = P.vmax && P.mode == 0;
    const int i0 = j % NP, i1 = (DIM > 1) ? (j / NP) % NP : 0, i2 = (DIM > 2) ? j / (NP * NP) : 0;
    const int idx[3] = {i0, i1, i2};
    double old[8];
#if WGPU_MAXWELL_PREFETCH_FACES
    double Fo[DIM][8];
#endif
    const double* src[DIM];   // where the outside trace of direction d lives (the node itself: no jump)
    int stride[DIM];
    double Jx = 0.0, Jy = 0.0, Jz = 0.0, rc = 0.0, wp2 = 0.0, qmax = 0.0;
    bool jump[DIM];   // the node lies on a face of direction d that has a neighbour
#pragma unroll
    for (int k = 0; k < 8; k++) old[k] = 0.0;
    if (active) {
        const size_t own = ((size_t)e * P.nc + nf0) * NN + j;
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int jd = idx[d];
            const bool on_face = (jd == 0 || jd == NP - 1);
            const int side = (jd == 0) ? 0 : 1;   // (Np >= 2: a node is on at most one face per direction)
            const int v = on_face ? P.nbr[(size_t)e * (2 * DIM) + 2 * d + side] : -1;
            const int t = face_node_index<DIM, NP>(d, i0, i1, i2);
            jump[d] = v >= 0;
            src[d] = P.u + own;
            stride[d] = NN;
            if (v >= P.n_elems) {
                src[d] = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + nf0) * NF + t;
                stride[d] = NF;
            } else if (v >= 0) {
                src[d] = P.u + ((size_t)v * P.nc + nf0) * NN + node_of_face_node<DIM, NP>(d, 1 - side, t);
            }
        }
        double F[8];   // parked in shared memory; re-read from there where needed (16 registers less across the kernel)
#pragma unroll
        for (int k = 0; k < 8; k++) F[k] = P.u[own + (size_t)k * NN];
#if WGPU_MAXWELL_PREFETCH_FACES
#pragma unroll
        for (int d = 0; d < DIM; d++)
#pragma unroll
            for (int k = 0; k < 8; k++) Fo[d][k] = src[d][(size_t)k * stride[d]];
#endif
#if WGPU_MAXWELL_LATE_OLD
        if (P.mode == 2 || (P.mode == 0 && P.beta != 0.0)) {
            const double* const op = (P.mode == 2 ? P.sol_in : P.dst) + own;
#pragma unroll
            for (int k = 0; k < 8; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(op + (size_t)k * NN));
        }
#else
        if (P.mode == 2) {
#pragma unroll
            for (int k = 0; k < 8; k++) old[k] = P.sol_in[own + (size_t)k * NN];
        } else if (P.mode == 0 && P.beta != 0.0) {
#pragma unroll
            for (int k = 0; k < 8; k++) old[k] = P.dst[own + (size_t)k * NN];
        }
#endif
#if WGPU_MAXWELL_LATE_J
        if (P.src_on) {
            for (int sp = 0; sp < P.nsp; sp++) {
                const size_t so = ((size_t)e * P.nc + 5 * sp) * NN + j;
#pragma unroll
                for (int k = 0; k < 4; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.u + so + (size_t)k * NN));
                if (want_speed) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.dst + so));
            }
        }
#else
        if (P.src_on) {
            // the same sums, in the same order, as the fluid kernels' field phase
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
        }
#endif
#pragma unroll
        for (int k = 0; k < 4; k++) sF[le][k][j] = make_double2(F[2 * k], F[2 * k + 1]);
    }
    __syncthreads();
    double vmax_local = 0.0;
    if (active) {
        double rate[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int st = stride_of(NP, d), jd = idx[d];
#if !WGPU_MAXWELL_PREFETCH_FACES
            // the outside trace (L2 hits: the neighbour's block has just been streamed), in flight during the volume term
            double Fod[8];
            if (jump[d]) {
#pragma unroll
                for (int k = 0; k < 8; k++) Fod[k] = src[d][(size_t)k * stride[d]];
            }
#else
            const double* const Fod = Fo[d];
#endif
            // volume: -(1/h_d) sum_l D[j_d][l] f_d(F_l) = -(1/h_d) f_d(sum_l D[j_d][l] F_l): the flux is linear
            double dF_[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int l = 0; l < NP; l++) {
                const int q = j + (l - jd) * st;
                const double w = P.T.D[jd * NP + l];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double2 v = sF[le][k][q];
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
            if (jump[d]) {
                double dF[8], fn[8];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const double2 own = sF[le][k][j];
                    dF[2 * k] = Fod[2 * k] - own.x;
                    dF[2 * k + 1] = Fod[2 * k + 1] - own.y;
                }
                phm_flux(d, M, dF, fn);
                const double cf = 0.5 * P.inv_hw[d], cl = cf * M.lam, cs = (jd == 0) ? cf : -cf;
#pragma unroll
                for (int k = 0; k < 8; k++) rate[k] = fma(cs, fn[k], fma(cl, dF[k], rate[k]));
            }
        }
#if WGPU_MAXWELL_LATE_J
        if (P.src_on) {
            // the same sums, in the same order, as the fluid kernels' field phase
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
        }
#endif
        // sources: -J/eps0 on E, chi rho_c/eps0 on phi
        if (P.src_on) {
            rate[0] += -Jx * P.inv_eps0;
            rate[1] += -Jy * P.inv_eps0;
            rate[2] += -Jz * P.inv_eps0;
            rate[6] += P.chi * rc * P.inv_eps0;
        }
        double Fn[8], F[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const double2 own = sF[le][k][j];
            F[2 * k] = own.x;
            F[2 * k + 1] = own.y;
        }
#if WGPU_MAXWELL_LATE_OLD
        if (P.mode == 2 || (P.mode == 0 && P.beta != 0.0)) {
            const double* const op = (P.mode == 2 ? P.sol_in : P.dst) + ((size_t)e * P.nc + nf0) * NN + j;
#pragma unroll
            for (int k = 0; k < 8; k++) old[k] = op[(size_t)k * NN];
        }
#endif
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
        if (want_speed) {
            vmax_local = M.speed_floor;
            if (P.src_on) {
                const double b2 = Fn[3] * Fn[3] + Fn[4] * Fn[4] + Fn[5] * Fn[5];
                const double omega = fmax(sqrt(wp2), qmax * sqrt(b2));
                vmax_local = nan_max(vmax_local, M.omega_factor * omega);
            }
        }
    }
    if (want_speed) {
        const double m = block_max(vmax_local, sRed);
        if (tid == 0) atomicMax(P.vmax, (unsigned long long)__double_as_longlong(m));
    }
}

// stand-alone: the field system's share of the transport speed of a vector (recommend_dt after an upload)
template <int DIM, int NP>
__global__ void maxwell_cfl_kernel(const double* __restrict__ u, int64_t n_elems, int nc, int nsp, const double* __restrict__ qmv,
                                   const MaxwellParams M, unsigned long long* vmax) {
    constexpr int NN = ipow_c(NP, DIM);
    __shared__ double sRed[32];
    double m = M.speed_floor;
    if (M.sources_on) {
        const int64_t total = n_elems * NN;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
            const int j = (int)(i % NN);
            const int64_t e = i / NN;
            double wp2 = 0.0, qmax = 0.0;
            for (int sp = 0; sp < nsp; sp++) {
                const double qm = qmv[sp];
                wp2 += qm * qm * u[((size_t)e * nc + 5 * sp) * NN + j] * M.inv_eps0;
                qmax = fmax(qmax, fabs(qm));
            }
            const size_t fo = ((size_t)e * nc + 5 * nsp) * NN + j;
            const double bx = u[fo + 3 * (size_t)NN], by = u[fo + 4 * (size_t)NN], bz = u[fo + 5 * (size_t)NN];
            m = nan_max(m, M.omega_factor * fmax(sqrt(wp2), qmax * sqrt(bx * bx + by * by + bz * bz)));
        }
    }
    m = block_max(m, sRed);
    if (threadIdx.x == 0) atomicMax(vmax, (unsigned long long)__double_as_longlong(m));
}

static MaxwellParams make_params(int Np, double light_speed, double chi, double gamma, double inv_eps0, double max_eig, bool sources_on) {
    MaxwellParams M;
    M.c2 = light_speed * light_speed;
    M.chi = chi;
    M.gam = gamma;
    double big = 1.0;
    if (chi > big) big = chi;
    if (gamma > big) big = gamma;
    M.lam = light_speed * big;
    M.inv_eps0 = inv_eps0;
    M.speed_floor = max_eig * M.lam;
    M.omega_factor = 5.0 / (double)(Np * Np);
    M.sources_on = sources_on ? 1 : 0;
    return M;
}

void launch_maxwell(int dim, int Np, const StageParams& P, double light_speed, double chi, double gamma, cudaStream_t s) {
    const int64_t n = P.elem_end - P.elem_begin;
    if (n <= 0) return;
    const MaxwellParams M = make_params(Np, light_speed, chi, gamma, P.inv_eps0, P.max_eig, P.src_on != 0);
#define CALL(D_, N_)                                                                              \
    {                                                                                             \
        using GEO = MGeo<D_, N_>;                                                                 \
        const int64_t blocks = (n + GEO::G - 1) / GEO::G;                                         \
        maxwell_kernel<D_, N_><<<(unsigned)blocks, GEO::THREADS, 0, s>>>(P, M);                   \
    }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

void launch_maxwell_cfl(int dim, int Np, const double* u, int64_t n_elems, int nc, int nsp, const double* qm, double light_speed,
                        double chi, double gamma, double inv_eps0, double max_eig, bool sources_on, unsigned long long* vmax,
                        cudaStream_t s) {
    const MaxwellParams M = make_params(Np, light_speed, chi, gamma, inv_eps0, max_eig, sources_on);
    int64_t blocks = sources_on ? (n_elems * ipow_c(Np, dim) + 255) / 256 : 1;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
#define CALL(D_, N_) maxwell_cfl_kernel<D_, N_><<<(unsigned)blocks, 256, 0, s>>>(u, n_elems, nc, nsp, qm, M, vmax)
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

}  // namespace wgpu
