// The stand-alone field kernel: perfectly hyperbolic Maxwell (PHM) fluxes for the 8 field components
// [Ex,Ey,Ez,Bx,By,Bz,phi,psi] (five_moment.h:123-138), north_star kernel 4 / BASELINE config 5; system, discretisation and the
// two halves of a thread's work in dgsem_maxwell_thread.cuh.
//
// One launch updates the field components of an element range with the same stage formula as the fluid kernels (modes
// 0/1/2 of StageParams) and, in the stage that fuses the CFL reduction, adds the field system's share to the transport
// speed: c max(1,chi,gamma) through the metric factor, and the plasma / cyclotron frequency bound.  It runs right after
// the fluid stage kernel of the same range on the same stream (the fluid kernels then skip the field components).
#include "dgsem_common.cuh"
#include "dgsem_maxwell_thread.cuh"

namespace wgpu {

#ifndef WGPU_MAXWELL_BLOCKS
#define WGPU_MAXWELL_BLOCKS 6   // 80 registers x 128 threads; measured on N3D (stage total): 5 blocks 3.756, 6: 3.754, 7: 3.859 ms
#endif

// One thread per node, 128-thread blocks; the two halves of a thread's work are in dgsem_maxwell_thread.cuh.
template <int DIM, int NP>
__global__ void __launch_bounds__(MGeo<DIM, NP>::THREADS, (MGeo<DIM, NP>::THREADS <= 128) ? WGPU_MAXWELL_BLOCKS : 2)
    maxwell_kernel(const StageParams P, const MaxwellParams M) {
    using GEO = MGeo<DIM, NP>;
    __shared__ double2 sF[GEO::SMEM_DOUBLE2];
    __shared__ double sRed[32];
    const int skip = P.skip_dev ? *P.skip_dev : 0;
    const double dt = P.dt_dev ? *P.dt_dev : P.dt;
    if (skip) return;
    MaxwellCarry<DIM> c;
    maxwell_pre<DIM, NP>(P, sF, threadIdx.x, blockIdx.x, c);
    __syncthreads();
    const double vmax_local = maxwell_post<DIM, NP>(P, M, sF, dt, c);
    if (P.vmax && P.mode == 0) {
        const double m = block_max(vmax_local, sRed);
        if (threadIdx.x == 0) atomicMax(P.vmax, (unsigned long long)__double_as_longlong(m));
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
