// Auxiliary kernels of the ES-DGSEM path for sm_100a: boundary faces, boundary-integrated-flux update, stand-alone
// CFL reduction, global integrals, halo packing.  The hot kernel is in dgsem_stage_kernel.cu.
#include "dgsem_common.cuh"
#include "dgsem_physics.cuh"

namespace wgpu {

// --------------------------------------------------------------------------------------------------------
// Boundary faces (fluid_flux_es_dgsem_operator.h:344-440): Gauss(p+2) face quadrature, ghost state by
// boundary kind, Lax-Friedrichs flux; one thread per (boundary face, species).  Emits the rate contribution
// per face node (already divided by the cell's diagonal mass) and the integrated numerical flux.
// --------------------------------------------------------------------------------------------------------
// One warp per (boundary face, species): lanes take the Gauss points (ghost state, Lax-Friedrichs flux), park the
// weighted values in shared memory, then lanes take the face nodes and integrate against the nodal basis in Gauss-point
// order (fixed order => deterministic).  The first version ran one THREAD per face and cost 0.7 ms per stage on the
// 1024^2 Kelvin-Helmholtz box, more than half of the fused stage kernel itself.
constexpr int kBoundaryWarps = 4;
template <int DIM, int NP>
__global__ void __launch_bounds__(32 * kBoundaryWarps) boundary_kernel(const BoundaryParams P) {
    constexpr int NN = ipow_c(NP, DIM), NF = ipow_c(NP, DIM - 1), NG1 = NP + 1, NG = ipow_c(NG1, DIM - 1);
    __shared__ double sv[kBoundaryWarps][NG][10];   // [0..4] (f(u_m).n - f*) wq, [5..9] f* area wq
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gid = (int64_t)blockIdx.x * kBoundaryWarps + warp;
    if (gid >= P.n_bfaces * P.nsp) return;   // whole warp
    const int64_t bf = gid / P.nsp;
    const int sp = (int)(gid - bf * P.nsp);
    const int e = P.bf_elem[bf], f = P.bf_side[bf], bid = P.bf_id[bf];
    const int d = f / 2, side = f % 2;
    const double sgn = side ? 1.0 : -1.0;
    const int kind = P.bc_kind[sp * P.n_boundaries + bid];

    double area = 1.0;
    for (int a = 0; a < DIM; a++) if (a != d) area *= P.h[a];

    for (int g = lane; g < NG; g += 32) {
        const int g0 = g % NG1, g1 = (g / NG1) % NG1;
        double wm[5] = {0, 0, 0, 0, 0};
        for (int t = 0; t < NF; t++) {
            const int t0 = t % NP, t1 = (t / NP) % NP;
            double phi = 1.0;
            if (DIM >= 2) phi *= P.Ig[g0 * NP + t0];
            if (DIM >= 3) phi *= P.Ig[g1 * NP + t1];
            // node of face node t: tangential dims in increasing order
            int id[3] = {0, 0, 0};
            {
                int tt[2] = {t0, t1}, k = 0;
                for (int a = 0; a < DIM; a++) { if (a == d) continue; id[a] = tt[k++]; }
                id[d] = side ? NP - 1 : 0;
            }
            const int node = id[0] + NP * (id[1] + NP * id[2]);
            const size_t off = ((size_t)e * P.nc + 5 * sp) * NN + node;
            for (int c = 0; c < 5; c++) wm[c] += phi * P.u[off + (size_t)c * NN];
        }
        double wp[5];
        if (kind == 2 || kind == 3) {   // inflow: prescribed conserved state; subsonic outflow: only its energy is used
            if (P.inflow_table) {
                const double* tab = P.inflow_table + (((size_t)sp * P.n_bfaces + bf) * NG + g) * 5;
                for (int c = 0; c < 5; c++) wp[c] = tab[c];
            } else {
                for (int c = 0; c < 5; c++) wp[c] = P.inflow[((size_t)sp * P.n_boundaries + bid) * 5 + c];
            }
            if (kind == 3) {   // fluid_flux_es_dgsem_operator.h:385-390: w_p = w_m with the total energy replaced
                for (int c = 0; c < 4; c++) wp[c] = wm[c];
            }
        } else if (kind == 1) {   // (supersonic) outflow
            for (int c = 0; c < 5; c++) wp[c] = wm[c];
        } else {   // wall: reflect the normal momentum over the first dim components, zero the rest
            const double rho_u_dot_n = wm[1 + d] * sgn;
            wp[0] = wm[0];
            for (int a = 0; a < 3; a++) wp[a + 1] = (a < DIM) ? wm[a + 1] : 0.0;
            wp[1 + d] = wm[1 + d] - 2.0 * rho_u_dot_n * sgn;
            wp[4] = wm[4];
        }
        double Fs[5], Fm[5];
        lf_flux_d<DIM>(d, sgn, wm, wp, P.gamma, Fs, Fm);
        double wq = 1.0;
        if (DIM >= 2) wq *= P.wg[g0];
        if (DIM >= 3) wq *= P.wg[g1];
        for (int c = 0; c < 5; c++) {
            sv[warp][g][c] = (Fm[c] - Fs[c]) * wq;
            sv[warp][g][5 + c] = Fs[c] * (area * wq);
        }
    }
    __syncwarp();
    // divide by the cell mass at the face node: Jdet * w_end * wF_t  (area / Jdet = 1/h_d)
    for (int t = lane; t < NF; t += 32) {
        const int t0 = t % NP, t1 = (t / NP) % NP;
        double acc[5] = {0, 0, 0, 0, 0};
        for (int g = 0; g < NG; g++) {
            const int g0 = g % NG1, g1 = (g / NG1) % NG1;
            double phi = 1.0;
            if (DIM >= 2) phi *= P.Ig[g0 * NP + t0];
            if (DIM >= 3) phi *= P.Ig[g1 * NP + t1];
            for (int c = 0; c < 5; c++) acc[c] += phi * sv[warp][g][c];
        }
        double wF = 1.0;
        if (DIM >= 2) wF *= P.w[t0];
        if (DIM >= 3) wF *= P.w[t1];
        const double cf = P.inv_h[d] / (P.w[0] * wF);
        for (int c = 0; c < 5; c++) P.bres[((size_t)(bf * P.nsp + sp) * 5 + c) * NF + t] = acc[c] * cf;
    }
    if (lane < 5) {
        double bsum = 0.0;
        for (int g = 0; g < NG; g++) bsum += sv[warp][g][5 + lane];
        P.bflux[(size_t)(bf * P.nsp + sp) * 5 + lane] = bsum;
    }
}

// one block per boundary id; threads stride over the faces, then a fixed-order sum of the 256 partials
// => deterministic boundary-integrated fluxes
constexpr int kBifThreads = 256;
__global__ void __launch_bounds__(kBifThreads) bif_update_kernel(const double* bflux, const int32_t* bf_id, int64_t n_bfaces, int nsp,
                                  int n_boundaries, double* bif_dst, const double* bif_u, double dt_host,
                                  const double* dt_dev, const int* skip_dev, double a, double beta, int mode) {
    if (skip_dev && *skip_dev) return;
    __shared__ double part[kBifThreads][5];
    const double dt = dt_dev ? *dt_dev : dt_host;
    const int bid = blockIdx.x, tid = threadIdx.x;
    double acc[5] = {0, 0, 0, 0, 0};
    for (int64_t bf = tid; bf < n_bfaces; bf += kBifThreads) {
        if (bf_id[bf] != bid) continue;
        for (int sp = 0; sp < nsp; sp++)
            for (int c = 0; c < 5; c++) acc[c] += bflux[(size_t)(bf * nsp + sp) * 5 + c];
    }
    for (int c = 0; c < 5; c++) part[tid][c] = acc[c];
    __syncthreads();
    if (tid >= 5) return;
    double rate = 0.0;
    for (int k = 0; k < kBifThreads; k++) rate += part[k][tid];
    const int i = bid * 5 + tid;
    if (mode == 1) { bif_dst[i] = rate; return; }
    double v = beta * bif_dst[i] + (a * dt) * rate;
    v = v + a * bif_u[i];
    bif_dst[i] = v;
}

template <int DIM, int NP>
__global__ void cfl_kernel(const double* __restrict__ u, int64_t n_elems, int nc, int nsp, double gamma,
                           double ih0, double ih1, double ih2, double max_eig, unsigned long long* vmax) {
    constexpr int NN = ipow_c(NP, DIM);
    __shared__ double sRed[32];
    const int64_t total = n_elems * nsp * NN;
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(i % NN);
        const int64_t es = i / NN;
        const int sp = (int)(es % nsp);
        const int64_t e = es / nsp;
        const size_t off = ((size_t)e * nc + 5 * sp) * NN + j;
        const double q0 = u[off], q1 = u[off + NN], q2 = u[off + 2 * (size_t)NN], q3 = u[off + 3 * (size_t)NN],
                     q4 = u[off + 4 * (size_t)NN];
        const double inv = 1.0 / q0;
        const double sm = q1 * q1 + q2 * q2 + q3 * q3;
        const double pr = (gamma - 1.0) * (q4 - sm * (0.5 * inv));
        double conv = fabs(q1 * inv) * ih0;
        if (DIM > 1) conv = fmax(conv, fabs(q2 * inv) * ih1);
        if (DIM > 2) conv = fmax(conv, fabs(q3 * inv) * ih2);
        m = nan_max(m, max_eig * sqrt(gamma * pr * inv) + conv);   // NaN (negative pressure) is kept, see nan_max
    }
    m = block_max(m, sRed);
    if (threadIdx.x == 0) atomicMax(vmax, (unsigned long long)__double_as_longlong(m));
}

template <int DIM, int NP>
__global__ void integral_partial_kernel(const double* __restrict__ u, int64_t n_elems, int nc, int species,
                                        double Jdet, const double* __restrict__ w1, double* partial) {
    // each block owns a contiguous element range; thread 0..4 = component; fixed order within the block
    constexpr int NN = ipow_c(NP, DIM);
    const int c = threadIdx.x;
    if (c >= 5) return;
    const int64_t per = (n_elems + gridDim.x - 1) / gridDim.x;
    const int64_t e0 = blockIdx.x * per, e1 = (e0 + per < n_elems) ? e0 + per : n_elems;
    double s = 0.0;
    for (int64_t e = e0; e < e1; e++) {
        const double* ue = u + ((size_t)e * nc + 5 * species + c) * NN;
        double cell = 0.0;
        for (int j = 0; j < NN; j++) {
            double wj = w1[j % NP];
            if (DIM > 1) wj *= w1[(j / NP) % NP];
            if (DIM > 2) wj *= w1[j / (NP * NP)];
            cell += ue[j] * (Jdet * wj);
        }
        s += cell;
    }
    partial[(size_t)blockIdx.x * 5 + c] = s;
}
__global__ void integral_final_kernel(const double* partial, int n_blocks, double* out) {
    const int c = threadIdx.x;
    if (c >= 5) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; b++) s += partial[(size_t)b * 5 + c];
    out[c] = s;
}

template <int DIM, int NP>
__global__ void pack_kernel(const double* __restrict__ u, const int32_t* __restrict__ send_elem,
                            const int32_t* __restrict__ send_side, int64_t n_send, int nc, int ncf,
                            double* __restrict__ sendbuf) {
    constexpr int NN = ipow_c(NP, DIM), NF = ipow_c(NP, DIM - 1);
    const int64_t total = n_send * ncf * NF;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(i % NF);
        const int c = (int)((i / NF) % ncf);
        const int64_t s = i / ((int64_t)NF * ncf);
        const int e = send_elem[s], f = send_side[s];
        const int d = f / 2, side = f % 2;
        int id[3] = {0, 0, 0};
        {
            int tt[2] = {t % NP, (t / NP) % NP}, k = 0;
            for (int a = 0; a < DIM; a++) { if (a == d) continue; id[a] = tt[k++]; }
            id[d] = side ? NP - 1 : 0;
        }
        const int node = id[0] + NP * (id[1] + NP * id[2]);
        sendbuf[i] = u[((size_t)e * nc + c) * NN + node];
    }
}

void launch_boundary(int dim, int Np, const BoundaryParams& P, cudaStream_t s) {
    const int64_t n = P.n_bfaces * P.nsp;
    if (n <= 0) return;
#define CALL(D_, N_) { boundary_kernel<D_, N_><<<(unsigned)((n + kBoundaryWarps - 1) / kBoundaryWarps), 32 * kBoundaryWarps, 0, s>>>(P); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

void launch_bif_update(const double* bflux, const int32_t* bf_id, int64_t n_bfaces, int nsp, int n_boundaries,
                       double* bif_dst, const double* bif_u, double dt, const double* dt_dev, const int* skip_dev, double a,
                       double beta, int mode, cudaStream_t s) {
    if (n_boundaries <= 0) return;
    bif_update_kernel<<<n_boundaries, kBifThreads, 0, s>>>(bflux, bf_id, n_bfaces, nsp, n_boundaries, bif_dst, bif_u, dt, dt_dev,
                                                           skip_dev, a, beta, mode);
}

void launch_cfl(int dim, int Np, const double* u, int64_t n_elems, int nc, int nsp, double gamma,
                const double* inv_h, double max_eig, unsigned long long* vmax, cudaStream_t s) {
    if (n_elems <= 0) return;
#define CALL(D_, N_) { cfl_kernel<D_, N_><<<148 * 8, 256, 0, s>>>(u, n_elems, nc, nsp, gamma, inv_h[0], inv_h[1], inv_h[2], max_eig, vmax); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

int integral_blocks(int64_t n_elems) { return (int)(n_elems < 1184 ? (n_elems > 0 ? n_elems : 1) : 1184); }

void launch_integral(int dim, int Np, const double* u, int64_t n_elems, int nc, int species, double Jdet,
                     const double* w, double* partial, double* out, cudaStream_t s) {
    const int nb = integral_blocks(n_elems);
#define CALL(D_, N_) { integral_partial_kernel<D_, N_><<<nb, 32, 0, s>>>(u, n_elems, nc, species, Jdet, w, partial); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
    integral_final_kernel<<<1, 32, 0, s>>>(partial, nb, out);
}

void launch_pack(int dim, int Np, const double* u, const int32_t* send_elem, const int32_t* send_side,
                 int64_t n_send, int nc, int ncf, double* sendbuf, cudaStream_t s) {
    if (n_send <= 0) return;
#define CALL(D_, N_) { pack_kernel<D_, N_><<<148 * 4, 256, 0, s>>>(u, send_elem, send_side, n_send, nc, ncf, sendbuf); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

}  // namespace wgpu

// ---- device-resident clock (inner loop of advance(), timestepper.cc:34-42) ------------------------------------
namespace wgpu {
__global__ void clock_kernel(DevClock* ck, unsigned long long* vmax, int finalize) {
    if (ck->pending) {   // book the step that has just been computed
        ck->t += ck->dt;
        ck->steps += 1;
        ck->pending = 0;
    }
    if (ck->done) return;
    if (!(ck->t < ck->t_stop - 1e-12) || (ck->max_steps > 0 && ck->steps >= ck->max_steps)) {
        ck->done = 1;
        ck->dt = 0.0;
        return;
    }
    if (finalize) return;   // end of a batch: the next batch starts with a full call
    double dt = ck->fixed_dt;
    if (!(dt > 0.0)) {
        const double v = __longlong_as_double((long long)*vmax);
        dt = 0.5 / (v * ck->np * ck->np);   // fluid_flux_es_dgsem_operator.h:446-447
    }
    if (!(dt > 0.0) || isinf(dt) || isnan(dt)) {   // unphysical state: stop, the host reports it
        ck->error = 1;
        ck->done = 1;
        ck->dt = dt;
        return;
    }
    ck->dt = fmin(dt, ck->t_stop - ck->t);
    ck->pending = 1;
    *vmax = 0ull;   // the second stage of the coming step reduces the new maximum into it
}
void launch_clock(DevClock* clock, unsigned long long* vmax, int finalize, cudaStream_t s) {
    clock_kernel<<<1, 1, 0, s>>>(clock, vmax, finalize);
}
}  // namespace wgpu

// ---- SM clock measured ON the device (no NVML call inside a timed region: those stall NCCL runs for milliseconds) ----
namespace wgpu {
__global__ void sm_clock_probe_kernel(double* out_mhz) {
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    const long long c0 = clock64();
    do {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    } while (t1 - t0 < 40000ull);   // 40 us window
    const long long c1 = clock64();
    *out_mhz = (double)(c1 - c0) / (double)(t1 - t0) * 1e3;
}
void launch_sm_clock_probe(double* out_mhz, cudaStream_t s) { sm_clock_probe_kernel<<<1, 1, 0, s>>>(out_mhz); }
}  // namespace wgpu

// ---- FP64 peak of the CUDA cores, measured: 8 independent fused-multiply-add chains per thread, all SMs full ----------
// (the denominator of the FP64 figure bench.py reports next to the HBM roofline; SURVEY.md 8(d) asks for a measured value)
namespace wgpu {
constexpr int kFmaChains = 8, kFmaIters = 4096;
__global__ void __launch_bounds__(256) fp64_fma_chain_kernel(double* sink, double seed) {
    double a[kFmaChains];
#pragma unroll
    for (int k = 0; k < kFmaChains; k++) a[k] = seed + 1e-3 * (threadIdx.x + k);
    const double m = 1.0 - 1e-9, c = 1e-9;
#pragma unroll 4
    for (int it = 0; it < kFmaIters; it++) {
#pragma unroll
        for (int k = 0; k < kFmaChains; k++) a[k] = fma(a[k], m, c);
    }
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < kFmaChains; k++) sum += a[k];
    if (sum == 123.456) sink[0] = sum;   // never true: keeps the chains alive
}
// returns fused multiply-adds per launch (thread level)
double launch_fp64_fma_chain(int blocks, double* sink, cudaStream_t s) {
    fp64_fma_chain_kernel<<<blocks, 256, 0, s>>>(sink, 0.5);
    return (double)blocks * 256.0 * kFmaChains * kFmaIters;
}
}  // namespace wgpu

// ---- FP64 rate as a function of resident warps and independent chains per thread (tuning diagnostic) ------------------
// One block of `threads` threads per SM and `ILP` independent FMA chains per thread: the rate this reaches relative to the peak
// above tells how many warps x chains the stage kernel needs in flight to keep the FP64 pipe busy (profiles/README.md).
namespace wgpu {
template <int ILP>
__global__ void fp64_ilp_kernel(double* sink, double seed) {
    double a[ILP];
#pragma unroll
    for (int k = 0; k < ILP; k++) a[k] = seed + 1e-3 * (threadIdx.x + k);
    const double m = 1.0 - 1e-9, c = 1e-9;
#pragma unroll 4
    for (int it = 0; it < kFmaIters; it++) {
#pragma unroll
        for (int k = 0; k < ILP; k++) a[k] = fma(a[k], m, c);
    }
    double sum = 0.0;
#pragma unroll
    for (int k = 0; k < ILP; k++) sum += a[k];
    if (sum == 123.456) sink[0] = sum;
}
double launch_fp64_ilp(int blocks, int threads, int ilp, double* sink, cudaStream_t s) {
    switch (ilp) {
        case 1: fp64_ilp_kernel<1><<<blocks, threads, 0, s>>>(sink, 0.5); break;
        case 2: fp64_ilp_kernel<2><<<blocks, threads, 0, s>>>(sink, 0.5); break;
        case 3: fp64_ilp_kernel<3><<<blocks, threads, 0, s>>>(sink, 0.5); break;
        case 4: fp64_ilp_kernel<4><<<blocks, threads, 0, s>>>(sink, 0.5); break;
        case 6: fp64_ilp_kernel<6><<<blocks, threads, 0, s>>>(sink, 0.5); break;
        default: fp64_ilp_kernel<8><<<blocks, threads, 0, s>>>(sink, 0.5); ilp = 8; break;
    }
    return (double)blocks * threads * ilp * kFmaIters;
}
}  // namespace wgpu

// ---- point physics on the device, for known-answer tests (the reference's euler_test.cc goldens) -----------------
namespace wgpu {
__global__ void point_flux_kernel(int n, const double* qa, const double* qb, int d, double gamma, double* ec,
                                  double* es, double* prim) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* a5 = qa + 5 * i;
    const double* b5 = qb + 5 * i;
    const Prim a = make_prim(a5[0], a5[1], a5[2], a5[3], a5[4], gamma);
    const Prim b = make_prim(b5[0], b5[1], b5[2], b5[3], b5[4], gamma);
    const double hig = 0.5 / (gamma - 1.0);
    double F[5], Dv[5], ibl;
    ec_flux_d(d, a, b, hig, F, ibl);
    es_dissipation(a, b, ibl, hig, Dv);
    for (int c = 0; c < 5; c++) { ec[5 * i + c] = F[c]; es[5 * i + c] = F[c] - Dv[c]; }
    double* p = prim + 12 * i;
    p[0] = a.rho; p[1] = a.u0; p[2] = a.u1; p[3] = a.u2; p[4] = a.beta; p[5] = a.lrho; p[6] = a.lbeta; p[7] = a.q2;
    p[8] = a.p; p[9] = a.H; p[10] = a.lam; p[11] = a.ib;
}
void launch_point_flux(int n, const double* qa, const double* qb, int d, double gamma, double* ec, double* es,
                       double* prim, cudaStream_t s) {
    if (n > 0) point_flux_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, qa, qb, d, gamma, ec, es, prim);
}
}  // namespace wgpu

// ---- self-check of div_rn_fast against the IEEE division, bit for bit ------------------------------------------
namespace wgpu {
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long& x) {
    unsigned long long z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
// mode 0: random significands, exponents in [-60, 60], random signs; mode 1: f / (2 + f) with f = m - 1,
// m in [sqrt(2)/2, sqrt(2)) (the division of det_log); mode 2: ratios close to 1 (b within a few thousand ulp of a);
// mode 3: numerators with few significant bits (exact and nearly exact quotients, ties).
__global__ void division_check_kernel(long long n, unsigned long long seed, int mode, unsigned long long* mismatches,
                                      double* first_bad) {
    const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long bad = 0;
    for (long long i = i0; i < n; i += stride) {
        unsigned long long st = seed + 0x632be59bd9b4e019ull * (unsigned long long)(i + 1);
        const unsigned long long r0 = splitmix64(st), r1 = splitmix64(st), r2 = splitmix64(st);
        double a, b;
        if (mode == 1) {
            const double u = (double)(r0 >> 11) * 0x1.0p-53;
            const double m = 0.70710678118654752 + u * (1.41421356237309505 - 0.70710678118654752);
            a = m - 1.0;
            b = 2.0 + a;
        } else if (mode == 2) {
            a = __longlong_as_double((long long)((r0 & 0x000fffffffffffffull) | 0x3ff0000000000000ull));
            b = __longlong_as_double(__double_as_longlong(a) + (long long)(r1 % 8192) - 4096);
            a = ldexp(a, (int)(r2 % 41) - 20);
        } else {
            const int ea = (int)(r2 % 121) - 60, eb = (int)((r2 >> 16) % 121) - 60;
            unsigned long long ma = r0 & 0x000fffffffffffffull;
            if (mode == 3) ma &= 0x000ff00000000000ull << (r2 >> 40) % 8;
            a = ldexp(__longlong_as_double((long long)(ma | 0x3ff0000000000000ull)), ea);
            b = ldexp(__longlong_as_double((long long)((r1 & 0x000fffffffffffffull) | 0x3ff0000000000000ull)), eb);
            if (r2 >> 63) a = -a;
            if ((r2 >> 62) & 1) b = -b;
        }
        const double want = __ddiv_rn(a, b);
        const double got = div_rn_fast(a, b);
        if (__double_as_longlong(want) != __double_as_longlong(got)) {
            if (atomicAdd(mismatches, 1ull) == 0ull) { first_bad[0] = a; first_bad[1] = b; first_bad[2] = got; first_bad[3] = want; }
            bad++;
        }
    }
    (void)bad;
}
void launch_division_check(long long n, unsigned long long seed, int mode, unsigned long long* mismatches, double* first_bad,
                           cudaStream_t s) {
    division_check_kernel<<<148 * 8, 256, 0, s>>>(n, seed, mode, mismatches, first_bad);
}
}  // namespace wgpu

