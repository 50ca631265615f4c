// Kernel parameter blocks and launch prototypes of the ES-DGSEM stage (host <-> device contract inside the
// library; nothing here is part of the C ABI).
#pragma once
#include <stdint.h>

#include "wgpu_portable.cuh"

namespace wgpu {

constexpr int kMaxNp = 7;   // fe_degree <= 6 (five_moment.h:116)

__host__ __device__ constexpr int ipow_c(int b, int e) { return e == 0 ? 1 : b * ipow_c(b, e - 1); }

// small index helpers shared by the kernels (and, through wgpu_portable.cuh, by the host emulation of tests/emu/)
__device__ __forceinline__ int stride_of(const int NP, const int d) { return d == 0 ? 1 : (d == 1 ? NP : NP * NP); }

// tangential (face-node) index of node (i0,i1,i2) on a face normal to d: remaining dims in increasing order
template <int DIM, int NP>
__device__ __forceinline__ int face_node_index(const int d, const int i0, const int i1, const int i2) {
    if (DIM == 1) return 0;
    if (DIM == 2) return d == 0 ? i1 : i0;
    return d == 0 ? (i1 + NP * i2) : (d == 1 ? (i0 + NP * i2) : (i0 + NP * i1));
}
// inverse: element-local node index of face node t on face (d, side)
template <int DIM, int NP>
__device__ __forceinline__ int node_of_face_node(const int d, const int side, const int t) {
    const int e = side ? NP - 1 : 0;
    if (DIM == 1) return e;
    if (DIM == 2) return d == 0 ? (e + NP * t) : (t + NP * e);
    const int t0 = t % NP, t1 = t / NP;
    return d == 0 ? (e + NP * (t0 + NP * t1)) : (d == 1 ? (t0 + NP * (e + NP * t1)) : (t0 + NP * (t1 + NP * e)));
}

// max that keeps a NaN once it has seen one (an unphysical state must surface as a NaN time step)
__device__ __forceinline__ double nan_max(const double a, const double b) { return (b > a || b != b) ? b : a; }


// Elements per thread block ("patch"): the largest power of two that keeps a block at <= 128 nodes (= threads); a
// neighbour inside the patch is read from shared memory instead of being gathered and re-derived.  Measured on C2
// (profiles/README.md): 64 -> 0.413 ms, 128 -> 0.416 ms, 256 -> 0.441 ms, 512 -> 0.486 ms per stage.
#ifndef WGPU_BLOCK_NODES
#define WGPU_BLOCK_NODES 128
#endif
__host__ __device__ constexpr int elems_per_block(int dim, int np) {
    int g = 1;
    while (2 * g * ipow_c(np, dim) <= WGPU_BLOCK_NODES) g *= 2;
    return g;
}

// 1-D reference-element tables (Gauss-Lobatto on [0,1]); filled by reference_element.hpp on the host.
struct ElemTables {
    double D[kMaxNp * kMaxNp];    // D[j*Np+l] = l_l'(x_j)
    double w[kMaxNp];             // GLL weights
    double V[kMaxNp * kMaxNp];    // Legendre analysis, V[k*Np+q] = (k+1/2) w_q sqrt2 P_k(2x_q-1)
};

struct StageParams {
    const double* u;              // input state  [elem][comp][node]
    double* dst;                  // output state (same layout); dst != u
    const int32_t* nbr;           // [n_elems][2*dim] face-pair table (warpii_gpu.h)
    const double* ghost;          // [n_ghost][5*nsp][nF] received face traces
    const double* bres;           // [n_bfaces][nsp][5][nF] boundary-face rate contributions (boundary kernel)
    double* alpha_out;            // optional [n_elems][nsp]
    unsigned long long* vmax;     // optional: max transport speed of dst (bits of a non-negative double)
    int64_t elem_begin, elem_end; // element range of this launch
    int64_t n_elems;
    int32_t nc, nsp;
    int32_t ncf;                  // components per ghost face trace: 5*nsp, or nc when the field system is evolved (Maxwell)
    int32_t fields_skip;          // != 0: the field components are updated by maxwell_kernel, not by the stage kernel
    int32_t mode;                 // 0: dst = beta*dst + a*(u + dt*rate); 1: dst = rate;
                                  // 2: dst = sol_in + a*rate, dst2 = sol_in + beta*rate (low-storage RK stage, rk.h:53-71)
    const double* sol_in;         // mode 2 only (may alias dst: every thread touches its own node only)
    double* dst2;                 // mode 2 only; differs from u and dst
    double gamma, dt, a, beta;
    const double* dt_dev;         // if non-null the time step is read from device memory (device-resident time loop)
    const int* skip_dev;          // if non-null and *skip_dev != 0 the launch is a no-op (time loop already finished)
    double hig;                   // 1 / (2 (gamma - 1))
    double inv_h[3];              // 1/h_d
    double inv_hw[3];             // 1/(h_d * w_0): face lifting factor (face JxW / cell JxW on a Cartesian cell)
    double max_eig;               // sqrt(lambda_max(J^-T J^-1)), :487-502 of the reference operator
    double ind_T, ind_sT;         // shock-indicator threshold T(Np) and gain s/T (persson_peraire_shock_indicator.h:110-112)
    // two-fluid source terms (north_star kernel 4; not in the reference, off by default): Lorentz force and E.J work on
    // the species, -J/eps0 on E and chi rho_c/eps0 on phi, fused into the stage update
    int32_t src_on;
    double inv_eps0, chi;
    const double* qm;             // [nsp] charge / mass
    // perfectly hyperbolic Maxwell fluxes for the field components, fused into the pencil stage kernel (mx_on; the other
    // stage kernels leave the fields to maxwell_kernel, see fields_skip): c^2, cleaning speeds chi / gamma (units of c),
    // Rusanov speed c max(1, chi, gamma), the field system's constant share of the transport speed (max_eig * lam) and the
    // factor 5 / Np^2 that turns the plasma / cyclotron frequency bound omega dt <= 0.1 into an equivalent speed
    int32_t mx_on;
    double mx_c2, mx_chi, mx_gam, mx_lam, mx_floor, mx_omega_factor;
    // pencil kernel: the patch this many elements ahead is the one the block that FOLLOWS this block on its SM will work on
    // (resident blocks x elements per patch, set by launch_pencil_stage; 0: off); its state is prefetched to L2
    int64_t lookahead;
    ElemTables T;
};

struct BoundaryParams {
    const double* u;
    double* bres;                 // [n_bfaces][nsp][5][nF]
    double* bflux;                // [n_bfaces][nsp][5]  integrated numerical flux per face (for bif)
    const int32_t* bf_elem;
    const int32_t* bf_side;
    const int32_t* bf_id;
    const int32_t* bc_kind;       // [nsp][n_boundaries]
    const double* inflow;         // [nsp][n_boundaries][5]
    const double* inflow_table;   // nullable: [nsp][n_bfaces][NG][5], state at every boundary quadrature point
    int64_t n_bfaces;
    int32_t nc, nsp, n_boundaries;
    double gamma;
    double inv_h[3], h[3];
    double w[kMaxNp];             // GLL weights
    int32_t Ng;
    double wg[kMaxNp + 1];        // Gauss(p+2) weights
    double Ig[(kMaxNp + 1) * kMaxNp];   // Ig[q*Np+i] = l_i(xg_q)
};

#if !WGPU_HOST_EMU
// General geometry (curved / non-rectangular elements, arbitrary conforming connectivity): per-node and per-face metric
// tables derived on the host from warpii_gpu_geometry (warpii_gpu_set_geometry), read by stage_kernel_general and the
// *_general auxiliary kernels.  K = dim*dim below.
struct GeneralParams {
    const double* gnode;      // [n_elems][K+2][NN]: plane d*dim+r = component r of Ja^d = Jdet * column d of J^-T
                              //   (jacobian_utils.h:32-40), plane K = 1/Jdet, plane K+1 = power-iteration value of :487-502
    const double* gsub;       // [n_elems][K][NN]: plane d*dim+r = component r of the subcell-face normal to the RIGHT of the
                              //   node along d, n_i = Ja^d_0 + sum_{k<=i} sum_m Q(k,m) Ja^d_m (subcell_finite_volume_flux.h:101-106,
                              //   :140-145); read only where alpha > 0
    const double* gface;      // [n_elems][2*dim][dim+1][NF]: unit outward normal, then face Jacobian / (Jdet * w_0) at the node
    const int2* nbr2;         // [n_elems][2*dim]: (face_neighbor entry, neighbour's local face + 8 * (tangential order reversed))
    const double* jdet;       // [n_elems][NN]: Jdet at the nodes (global integrals)
    const double* bgeo;       // [n_bfaces][NG][dim+1]: unit outward normal and surface Jacobian at the boundary Gauss points
    const double* bmass;      // [n_bfaces][NF]: 1 / (Jdet * tensor GLL weight) of the face nodes
};

// Device-resident clock of warpii_gpu_advance_to: the inner loop of advance() (timestepper.cc:34-42) without a host
// round trip per step.
struct DevClock {
    double t, dt, t_stop, fixed_dt;
    long long steps, max_steps;
    int done, error, pending, np;
};
// finalize != 0: only book the pending step (t += dt, steps++).  Otherwise also decide whether to continue and set the
// next dt = min(fixed_dt or 0.5/(vmax np^2), t_stop - t), then clear *vmax for the fused reduction of the coming step.
void launch_clock(DevClock* clock, unsigned long long* vmax, int finalize, cudaStream_t s);

// one thread spins 40 us and reports (SM cycles / wall ns): the SM clock in MHz at that point of the stream
void launch_sm_clock_probe(double* out_mhz, cudaStream_t s);

int stage_smem_bytes(int dim, int Np);
int prepare_kernels(int dim, int Np);   // opt in to the dynamic shared memory the stage kernel needs; 0 on success
void launch_stage(int dim, int Np, const StageParams& P, cudaStream_t s);
void launch_boundary(int dim, int Np, const BoundaryParams& P, cudaStream_t s);
// bif_rate[n_boundaries*5] = sum over faces/species of bflux (fixed order); then the stage update of the
// boundary-integrated fluxes (fluid_flux_es_dgsem_operator.h:207-212)
void launch_bif_update(const double* bflux, const int32_t* bf_id, int64_t n_bfaces, int nsp, int n_boundaries,
                       double* bif_dst, const double* bif_u, double dt, const double* dt_dev, const int* skip_dev, double a,
                       double beta, int mode, cudaStream_t s);
void launch_cfl(int dim, int Np, const double* u, int64_t n_elems, int nc, int nsp, double gamma,
                const double* inv_h, double max_eig, unsigned long long* vmax, cudaStream_t s);
// deterministic two-pass reduction: out[5] = sum_e sum_j u * (Jdet * w_j); partial must hold 5*n_blocks doubles
int integral_blocks(int64_t n_elems);
void launch_integral(int dim, int Np, const double* u, int64_t n_elems, int nc, int species, double Jdet,
                     const double* w, double* partial, double* out, cudaStream_t s);
// pencil-per-thread stage kernel (dgsem_pencil_kernel.cu, dgsem_pencil_stage.cuh): Cartesian boxes, 2-D / 3-D, Np = 3..5
bool pencil_available(int dim, int Np);
int pencil_patch_elems(int dim, int Np);   // elements per block (= patch of the element numbering); 0 if not available
int pencil_smem_bytes(int dim, int Np);
int prepare_pencil_kernels(int dim, int Np);
void launch_pencil_stage(int dim, int Np, const StageParams& P, cudaStream_t s);
// general-geometry counterparts (dgsem_general_kernel.cu)
int stage_general_smem_bytes(int dim, int Np);
int prepare_general_kernels(int dim, int Np);
void launch_stage_general(int dim, int Np, const StageParams& P, const GeneralParams& GP, cudaStream_t s);
void launch_boundary_general(int dim, int Np, const BoundaryParams& P, const GeneralParams& GP, cudaStream_t s);
void launch_cfl_general(int dim, int Np, const double* u, int64_t n_elems, int nc, int nsp, double gamma,
                        const GeneralParams& GP, unsigned long long* vmax, cudaStream_t s);
void launch_integral_general(int dim, int Np, const double* u, int64_t n_elems, int nc, int species, const GeneralParams& GP,
                             const double* w, double* partial, double* out, cudaStream_t s);
// halo: sendbuf[i][ncf][nF] = trace of the first ncf components on (send_elem[i], send_side[i])
void launch_pack(int dim, int Np, const double* u, const int32_t* send_elem, const int32_t* send_side,
                 int64_t n_send, int nc, int ncf, double* sendbuf, cudaStream_t s);
// perfectly hyperbolic Maxwell fluxes for the field components (dgsem_maxwell_kernel.cu)
void launch_maxwell(int dim, int Np, const StageParams& P, double light_speed, double chi, double gamma, cudaStream_t s);
void launch_maxwell_cfl(int dim, int Np, const double* u, int64_t n_elems, int nc, int nsp, const double* qm, double light_speed,
                        double chi, double gamma, double inv_eps0, double max_eig, bool sources_on, unsigned long long* vmax,
                        cudaStream_t s);

#endif  // !WGPU_HOST_EMU

}  // namespace wgpu
