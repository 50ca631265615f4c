// ES-DGSEM stage kernels for sm_100a (FP64 on the CUDA cores).
//
// stage_kernel fuses, for every element it owns, what the reference does in separate global sweeps
// (fluid_flux_es_dgsem_operator.h:127-214): gather, shock indicator (persson_peraire_shock_indicator.h),
// split-form volume term (split_form_volume_flux.h), subcell-FV blend (subcell_finite_volume_flux.h),
// both sides' face lifting (:301-342) in gather form, inverse diagonal mass, the RK stage update and
// (optionally) the CFL reduction of the updated state (:450-514).  One HBM read of u, one write of dst.
#include "dgsem_kernels.cuh"
#include "dgsem_physics.cuh"

#include <cstdio>

namespace wgpu {

template <int DIM, int NP>
struct Geo {
    static constexpr int NN = ipow_c(NP, DIM);          // nodes per element
    static constexpr int NF = ipow_c(NP, DIM - 1);      // nodes per face
    static constexpr int G = elems_per_block(DIM, NP);  // elements per block
    static constexpr int NODES = NN * G;                // nodes (= threads) per block
    static constexpr int NFACE = 2 * DIM;
    static constexpr int NSLOT = G * NFACE * NF;        // face-node result slots per block
    static constexpr int THREADS = NODES;
    // dynamic shared memory, in doubles
    static constexpr int OFF_D = 0;
    static constexpr int OFF_V = OFF_D + NP * NP;
    static constexpr int OFF_W = OFF_V + NP * NP;
    static constexpr int OFF_P = OFF_W + 8;                 // [kPrim][NODES]
    static constexpr int OFF_A = OFF_P + kPrim * NODES;     // indicator scratch
    static constexpr int OFF_B = OFF_A + NODES;
    static constexpr int OFF_FACE = OFF_B + NODES;          // [5][NSLOT]
    static constexpr int OFF_G = OFF_FACE + 5 * NSLOT;      // [5][NODES] subcell interface fluxes
    static constexpr int OFF_ALPHA = OFF_G + 5 * NODES;     // [G]
    static constexpr int OFF_RED = OFF_ALPHA + G;           // [32]
    static constexpr int SMEM_DOUBLES = OFF_RED + 32;
};

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

__device__ __forceinline__ double block_max(double v, double* s_red) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double r = s_red[0];
    for (int i = 1; i < nwarps; i++) r = fmax(r, s_red[i]);
    return r;
}

// persson_peraire_shock_indicator.h:96-122 given the two modal energies; T and s/T are host constants
__device__ __forceinline__ double blending_from_energies(const double g0, const double g1, const double T, const double sT) {
    const double n0 = sqrt(g0), n1 = sqrt(g1);
    double total_m1 = 0.0, top_m1 = 0.0;
    if (n0 > 1e-10) total_m1 += n0 * n0;
    if (n1 > 1e-10) { const double e = n1 * n1; top_m1 += e; total_m1 += e; }
    const double E = fmax(0.0, top_m1 / total_m1);
    double alpha = 1.0 / (1.0 + exp(-sT * (E - T)));
    if (alpha < 1e-3) alpha = 0.0;
    else if (alpha > 0.5) alpha = 0.5;
    return alpha;
}

__device__ __forceinline__ Prim load_prim(const double* sP, const int nodes, const int n) {
    Prim o;
    o.rho = sP[0 * nodes + n];  o.u0 = sP[1 * nodes + n];   o.u1 = sP[2 * nodes + n];    o.u2 = sP[3 * nodes + n];
    o.beta = sP[4 * nodes + n]; o.lrho = sP[5 * nodes + n]; o.lbeta = sP[6 * nodes + n]; o.p = sP[7 * nodes + n];
    o.H = sP[8 * nodes + n];    o.lam = sP[9 * nodes + n];  o.ib = sP[10 * nodes + n];
    return o;
}

template <int DIM, int NP>
__global__ void __launch_bounds__(Geo<DIM, NP>::THREADS) stage_kernel(const StageParams P) {
    using GEO = Geo<DIM, NP>;
    constexpr int NN = GEO::NN, NF = GEO::NF, G = GEO::G, NODES = GEO::NODES, NFACE = GEO::NFACE, NSLOT = GEO::NSLOT;

    extern __shared__ double smem[];
    double* const sD = smem + GEO::OFF_D;
    double* const sV = smem + GEO::OFF_V;
    double* const sW = smem + GEO::OFF_W;
    double* const sP = smem + GEO::OFF_P;
    double* const sA = smem + GEO::OFF_A;
    double* const sB = smem + GEO::OFF_B;
    double* const sFace = smem + GEO::OFF_FACE;
    double* const sG = smem + GEO::OFF_G;
    double* const sAlpha = smem + GEO::OFF_ALPHA;
    double* const sRed = smem + GEO::OFF_RED;

    const int tid = threadIdx.x;
    const int le = tid / NN;
    const int j = tid - le * NN;
    const int i0 = j % NP, i1 = (DIM > 1) ? (j / NP) % NP : 0, i2 = (DIM > 2) ? j / (NP * NP) : 0;
    const int idx[3] = {i0, i1, i2};
    const int64_t e0 = P.elem_begin + (int64_t)blockIdx.x * G;
    const int64_t e = e0 + le;
    const bool active = e < P.elem_end;
    const int64_t bl = P.block_begin + blockIdx.x;
    const int n_face_tasks = P.face_count[bl] * NF;
    const int32_t* const flist = P.face_list + bl * (G * NFACE);

    for (int i = tid; i < NP * NP; i += NODES) { sD[i] = P.T.D[i]; sV[i] = P.T.V[i]; }
    for (int i = tid; i < NP; i += NODES) sW[i] = P.T.w[i];

    const double gamma = P.gamma, gm1 = P.gamma - 1.0;
    const double hig = 0.5 / gm1;   // 1 / (2 (gamma - 1))
    const int nc = P.nc;
    double vmax_local = 0.0;

    for (int sp = 0; sp < P.nsp; sp++) {
        __syncthreads();   // smem reuse across species (and table load)
        // ---- node phase: load, primitives, logs, wave speed ----------------------------------------
        const size_t off = ((size_t)e * nc + 5 * sp) * NN + j;
        double q[5] = {1.0, 0.0, 0.0, 0.0, 1.0};
        if (active) {
#pragma unroll
            for (int c = 0; c < 5; c++) q[c] = P.u[off + (size_t)c * NN];
        }
        const Prim me = make_prim(q[0], q[1], q[2], q[3], q[4], gamma);
        sP[0 * NODES + tid] = me.rho;  sP[1 * NODES + tid] = me.u0;   sP[2 * NODES + tid] = me.u1;    sP[3 * NODES + tid] = me.u2;
        sP[4 * NODES + tid] = me.beta; sP[5 * NODES + tid] = me.lrho; sP[6 * NODES + tid] = me.lbeta; sP[7 * NODES + tid] = me.p;
        sP[8 * NODES + tid] = me.H;    sP[9 * NODES + tid] = me.lam;  sP[10 * NODES + tid] = me.ib;
        sA[tid] = me.p * me.rho;   // indicator variable, fluid_flux_es_dgsem_operator.h:286-290
        __syncthreads();

        // ---- face phase: compact list of the block's faces; internal ones serve both elements ----------
        for (int k = tid; k < n_face_tasks; k += NODES) {
            const int fi = k / NF, t = k - fi * NF;
            const int32_t desc = flist[fi];
            const int fle = desc & 255, f = (desc >> 8) & 15, kind = (desc >> 12) & 3, nle = (desc >> 16) & 255;
            const int d = f >> 1, side = f & 1;
            const int slot = (fle * NFACE + f) * NF + t;
            if (kind == kFaceBoundary) {
                const int v = P.nbr[(size_t)(e0 + fle) * NFACE + f];
                const size_t offb = (((size_t)(-1 - v) * P.nsp + sp) * 5) * NF + t;
#pragma unroll
                for (int c = 0; c < 5; c++) sFace[c * NSLOT + slot] = P.bres[offb + (size_t)c * NF];
                continue;
            }
            const Prim a = load_prim(sP, NODES, fle * NN + node_of_face_node<DIM, NP>(d, side, t));
            Prim b;
            if (kind == kFaceInternal) {
                b = load_prim(sP, NODES, nle * NN + node_of_face_node<DIM, NP>(d, 1 - side, t));
            } else {
                const int v = P.nbr[(size_t)(e0 + fle) * NFACE + f];
                double qn[5];
                if (kind == kFaceElem) {
                    const size_t offn = ((size_t)v * nc + 5 * sp) * NN + node_of_face_node<DIM, NP>(d, 1 - side, t);
#pragma unroll
                    for (int c = 0; c < 5; c++) qn[c] = P.u[offn + (size_t)c * NN];
                } else {
                    const size_t offg = ((size_t)(v - P.n_elems) * (5 * P.nsp) + 5 * sp) * NF + t;
#pragma unroll
                    for (int c = 0; c < 5; c++) qn[c] = P.ghost[offg + (size_t)c * NF];
                }
                b = make_prim(qn[0], qn[1], qn[2], qn[3], qn[4], gamma);
            }
            const double sgn = side ? 1.0 : -1.0;
            const double cf = P.inv_hw[d];
            double Fs[5], Fm[5];
            es_flux_d<DIM>(d, a, b, sgn, hig, Fs);
            phys_flux_d<DIM>(d, a, Fm);
#pragma unroll
            for (int c = 0; c < 5; c++) sFace[c * NSLOT + slot] = cf * (sgn * Fm[c] - Fs[c]);
            if (kind == kFaceInternal) {
                // the neighbour's side of the same face: n' = -n, f*(b,a,n') = -f*(a,b,n) exactly
                double Fn[5];
                phys_flux_d<DIM>(d, b, Fn);
                const int slot2 = (nle * NFACE + (f ^ 1)) * NF + t;
#pragma unroll
                for (int c = 0; c < 5; c++) sFace[c * NSLOT + slot2] = cf * (Fs[c] - sgn * Fn[c]);
            }
        }

        // ---- shock indicator: sum-factorised Legendre analysis of p*rho ------------------------------
        {
            double* src = sA;
            double* dstb = sB;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                const int st = stride_of(NP, d);
                const int base = tid - idx[d] * st;
                double acc = 0.0;
#pragma unroll
                for (int m = 0; m < NP; m++) acc += sV[idx[d] * NP + m] * src[base + m * st];
                dstb[tid] = acc;
                __syncthreads();
                double* t = src; src = dstb; dstb = t;
            }
            // src holds the modal coefficients c_k at k = (i0,i1,i2)
            const double ck = src[tid];
            // two-level fixed-order sums: along i0, then over the remaining NP^(DIM-1) partials
            dstb[tid] = ck * ck;
            __syncthreads();
            if (i0 == 0) {
                double g0 = 0.0, g1 = 0.0;
                const bool row_shell = (i1 == NP - 1) || (i2 == NP - 1);
#pragma unroll
                for (int m = 0; m < NP; m++) {
                    const double v = dstb[tid + m];
                    if (row_shell || m == NP - 1) g1 += v; else g0 += v;
                }
                src[tid] = g0;
                src[tid + 1] = g1;   // NP >= 2
            }
            __syncthreads();
            if (j == 0) {
                double g0 = 0.0, g1 = 0.0;
                for (int r = 0; r < NF; r++) { g0 += src[tid + r * NP]; g1 += src[tid + r * NP + 1]; }
                const double al = blending_from_energies(g0, g1, P.ind_T, P.ind_sT);
                sAlpha[le] = al;
                if (active && P.alpha_out) P.alpha_out[(size_t)e * P.nsp + sp] = al;
            }
            __syncthreads();
        }
        const double alpha = sAlpha[le];

        double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};

        // ---- split-form volume term: rate = (1-alpha) * sum_d (-2/h_d) sum_l D[j_d][l] F#_d(u_j,u_l) ----
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int st = stride_of(NP, d);
            const int jd = idx[d];
            const int base = tid - jd * st;
            double acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int l = 0; l < NP; l++) {
                const double djl = sD[jd * NP + l];
                double F[5];
                if (l == jd) {
                    if (djl == 0.0) continue;    // interior diagonal of the GLL derivative matrix vanishes
                    phys_flux_d<DIM>(d, me, F);   // F#(u,u) = f(u)
                } else {
                    const Prim o = load_prim(sP, NODES, base + l * st);
                    ec_flux_d<DIM>(d, me, o, hig, F);
                }
#pragma unroll
                for (int c = 0; c < 5; c++) acc[c] += djl * F[c];
            }
            const double s = -2.0 * P.inv_h[d];
#pragma unroll
            for (int c = 0; c < 5; c++) r[c] += s * acc[c];
        }
        {
            const double oma = 1.0 - alpha;
#pragma unroll
            for (int c = 0; c < 5; c++) r[c] *= oma;
        }

        // ---- subcell finite-volume blend (only where the indicator fired) -----------------------------
        if (__syncthreads_or(active && alpha > 0.0)) {
            for (int d = 0; d < DIM; d++) {
                const int st = stride_of(NP, d);
                const int jd = idx[d];
                const bool on = active && alpha > 0.0;
                if (on && jd < NP - 1) {
                    const Prim o = load_prim(sP, NODES, tid + st);
                    double F[5];
                    es_flux_d<DIM>(d, me, o, 1.0, hig, F);
#pragma unroll
                    for (int c = 0; c < 5; c++) sG[c * NODES + tid] = F[c];
                }
                __syncthreads();
                if (on) {
                    double Fp[5];
                    phys_flux_d<DIM>(d, me, Fp);
                    const double cf = alpha * P.inv_h[d] / sW[jd];
#pragma unroll
                    for (int c = 0; c < 5; c++) {
                        const double left = (jd == 0) ? Fp[c] : sG[c * NODES + tid - st];
                        const double right = (jd == NP - 1) ? Fp[c] : sG[c * NODES + tid];
                        r[c] += cf * (left - right);
                    }
                }
                __syncthreads();
            }
        }

        // ---- collect this node's face contributions (written by the face phase; barriers above order them) --
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int t = face_node_index<DIM, NP>(d, i0, i1, i2);
            if (idx[d] == 0) {
                const int slot = (le * NFACE + 2 * d) * NF + t;
#pragma unroll
                for (int c = 0; c < 5; c++) r[c] += sFace[c * NSLOT + slot];
            }
            if (idx[d] == NP - 1) {
                const int slot = (le * NFACE + 2 * d + 1) * NF + t;
#pragma unroll
                for (int c = 0; c < 5; c++) r[c] += sFace[c * NSLOT + slot];
            }
        }

        // ---- inverse mass is folded into the factors above; stage update --------------------------------
        if (active) {
            double qn[5];
            if (P.mode == 1) {
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = r[c];
            } else if (P.beta == 0.0) {
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = P.a * (q[c] + P.dt * r[c]);
            } else {
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = P.beta * P.dst[off + (size_t)c * NN] + P.a * (q[c] + P.dt * r[c]);
            }
#pragma unroll
            for (int c = 0; c < 5; c++) P.dst[off + (size_t)c * NN] = qn[c];

            if (P.vmax && P.mode == 0) {
                // compute_cell_transport_speed (:450-514) of the updated state
                const double inv = 1.0 / qn[0];
                const double sm = qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3];
                const double pr = gm1 * (qn[4] - sm * (0.5 * inv));
                double conv = fabs(qn[1] * inv) * P.inv_h[0];
                if (DIM > 1) conv = fmax(conv, fabs(qn[2] * inv) * P.inv_h[1]);
                if (DIM > 2) conv = fmax(conv, fabs(qn[3] * inv) * P.inv_h[2]);
                const double cs = sqrt(gamma * pr * inv);
                vmax_local = fmax(vmax_local, P.max_eig * cs + conv);
            }
        }
    }

    // ---- field components are carried through unchanged by this operator -------------
    if (active && nc > 5 * P.nsp) {
        for (int c = 5 * P.nsp; c < nc; c++) {
            const size_t off = ((size_t)e * nc + c) * NN + j;
            double v;
            if (P.mode == 1) v = 0.0;
            else if (P.beta == 0.0) v = P.a * P.u[off];
            else v = P.beta * P.dst[off] + P.a * P.u[off];
            P.dst[off] = v;
        }
    }

    if (P.vmax && P.mode == 0) {
        const double m = block_max(vmax_local, sRed);
        if (tid == 0) atomicMax(P.vmax, (unsigned long long)__double_as_longlong(m));
    }
}

// --------------------------------------------------------------------------------------------------------
// Boundary faces (fluid_flux_es_dgsem_operator.h:344-440): Gauss(p+2) face quadrature, ghost state by
// boundary kind, Lax-Friedrichs flux; one thread per (boundary face, species).  Emits the rate contribution
// per face node (already divided by the cell's diagonal mass) and the integrated numerical flux.
// --------------------------------------------------------------------------------------------------------
template <int DIM, int NP>
__global__ void boundary_kernel(const BoundaryParams P) {
    constexpr int NN = ipow_c(NP, DIM), NF = ipow_c(NP, DIM - 1), NG1 = NP + 1, NG = ipow_c(NG1, DIM - 1);
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P.n_bfaces * P.nsp) return;
    const int64_t bf = gid / P.nsp;
    const int sp = (int)(gid - bf * P.nsp);
    const int e = P.bf_elem[bf], f = P.bf_side[bf], bid = P.bf_id[bf];
    const int d = f / 2, side = f % 2;
    const double sgn = side ? 1.0 : -1.0;
    const int kind = P.bc_kind[sp * P.n_boundaries + bid];
    const int st = stride_of(NP, d);

    double area = 1.0;
    for (int a = 0; a < DIM; a++) if (a != d) area *= P.h[a];

    double acc[5][NF];
#pragma unroll
    for (int c = 0; c < 5; c++)
        for (int t = 0; t < NF; t++) acc[c][t] = 0.0;
    double bsum[5] = {0, 0, 0, 0, 0};

    for (int g = 0; g < NG; g++) {
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
        if (kind == 2) {   // inflow: prescribed conserved state
            for (int c = 0; c < 5; c++) wp[c] = P.inflow[((size_t)sp * P.n_boundaries + bid) * 5 + c];
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
        for (int c = 0; c < 5; c++) bsum[c] += Fs[c] * (area * wq);
        for (int t = 0; t < NF; t++) {
            const int t0 = t % NP, t1 = (t / NP) % NP;
            double phi = 1.0;
            if (DIM >= 2) phi *= P.Ig[g0 * NP + t0];
            if (DIM >= 3) phi *= P.Ig[g1 * NP + t1];
            for (int c = 0; c < 5; c++) acc[c][t] += phi * ((Fm[c] - Fs[c]) * wq);
        }
    }
    // divide by the cell mass at the face node: Jdet * w_end * wF_t  (area / Jdet = 1/h_d)
    for (int t = 0; t < NF; t++) {
        const int t0 = t % NP, t1 = (t / NP) % NP;
        double wF = 1.0;
        if (DIM >= 2) wF *= P.w[t0];
        if (DIM >= 3) wF *= P.w[t1];
        const double cf = P.inv_h[d] / (P.w[0] * wF);
        for (int c = 0; c < 5; c++) P.bres[((size_t)(bf * P.nsp + sp) * 5 + c) * NF + t] = acc[c][t] * cf;
    }
    for (int c = 0; c < 5; c++) P.bflux[(size_t)(bf * P.nsp + sp) * 5 + c] = bsum[c];
    (void)st;
}

// one block; fixed summation order => deterministic boundary-integrated fluxes
__global__ void bif_update_kernel(const double* bflux, const int32_t* bf_id, int64_t n_bfaces, int nsp,
                                  int n_boundaries, double* bif_dst, const double* bif_u, double dt, double a,
                                  double beta, int mode) {
    const int i = threadIdx.x;   // i = bid*5 + c
    if (i >= n_boundaries * 5) return;
    const int bid = i / 5, c = i % 5;
    double rate = 0.0;
    for (int64_t bf = 0; bf < n_bfaces; bf++) {
        if (bf_id[bf] != bid) continue;
        for (int sp = 0; sp < nsp; sp++) rate += bflux[(size_t)(bf * nsp + sp) * 5 + c];
    }
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
        m = fmax(m, max_eig * sqrt(gamma * pr * inv) + conv);
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
                            const int32_t* __restrict__ send_side, int64_t n_send, int nc, int nsp,
                            double* __restrict__ sendbuf) {
    constexpr int NN = ipow_c(NP, DIM), NF = ipow_c(NP, DIM - 1);
    const int ncf = 5 * nsp;
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

// --------------------------------------------------------------------------------------------------------
// host-side dispatch over (dim, Np)
// --------------------------------------------------------------------------------------------------------
#define WGPU_DISPATCH(dim, Np, CALL)                                                     \
    do {                                                                                 \
        if ((dim) == 1) {                                                                \
            switch (Np) { case 2: { CALL(1, 2); } break; case 3: { CALL(1, 3); } break; case 4: { CALL(1, 4); } break; \
                          case 5: { CALL(1, 5); } break; case 6: { CALL(1, 6); } break; case 7: { CALL(1, 7); } break; } \
        } else if ((dim) == 2) {                                                         \
            switch (Np) { case 2: { CALL(2, 2); } break; case 3: { CALL(2, 3); } break; case 4: { CALL(2, 4); } break; \
                          case 5: { CALL(2, 5); } break; case 6: { CALL(2, 6); } break; case 7: { CALL(2, 7); } break; } \
        } else {                                                                         \
            switch (Np) { case 2: { CALL(3, 2); } break; case 3: { CALL(3, 3); } break; case 4: { CALL(3, 4); } break; \
                          case 5: { CALL(3, 5); } break; case 6: { CALL(3, 6); } break; case 7: { CALL(3, 7); } break; } \
        }                                                                                \
    } while (0)

int stage_smem_bytes(int dim, int Np) {
    int bytes = 0;
#define CALL(D_, N_) { bytes = Geo<D_, N_>::SMEM_DOUBLES * (int)sizeof(double); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
    return bytes;
}

int prepare_kernels(int dim, int Np) {
    cudaError_t err = cudaSuccess;
#define CALL(D_, N_)                                                                                         \
    {                                                                                                        \
        err = cudaFuncSetAttribute(stage_kernel<D_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                   Geo<D_, N_>::SMEM_DOUBLES * (int)sizeof(double));                         \
    }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
    return err == cudaSuccess ? 0 : 1;
}

void launch_stage(int dim, int Np, const StageParams& P, cudaStream_t s) {
    const int64_t n = P.elem_end - P.elem_begin;
    if (n <= 0) return;
#define CALL(D_, N_)                                                                                         \
    {                                                                                                        \
        using GEO = Geo<D_, N_>;                                                                             \
        const int64_t blocks = (n + GEO::G - 1) / GEO::G;                                                    \
        stage_kernel<D_, N_><<<(unsigned)blocks, GEO::THREADS, GEO::SMEM_DOUBLES * sizeof(double), s>>>(P);  \
    }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

void launch_boundary(int dim, int Np, const BoundaryParams& P, cudaStream_t s) {
    const int64_t n = P.n_bfaces * P.nsp;
    if (n <= 0) return;
#define CALL(D_, N_) { boundary_kernel<D_, N_><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(P); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

void launch_bif_update(const double* bflux, const int32_t* bf_id, int64_t n_bfaces, int nsp, int n_boundaries,
                       double* bif_dst, const double* bif_u, double dt, double a, double beta, int mode,
                       cudaStream_t s) {
    if (n_boundaries <= 0) return;
    bif_update_kernel<<<1, ((n_boundaries * 5 + 31) / 32) * 32, 0, s>>>(bflux, bf_id, n_bfaces, nsp, n_boundaries,
                                                                          bif_dst, bif_u, dt, a, beta, mode);
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
                 int64_t n_send, int nc, int nsp, double* sendbuf, cudaStream_t s) {
    if (n_send <= 0) return;
#define CALL(D_, N_) { pack_kernel<D_, N_><<<148 * 4, 256, 0, s>>>(u, send_elem, send_side, n_send, nc, nsp, sendbuf); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

}  // namespace wgpu
