// The fused ES-DGSEM stage kernel on GENERAL geometry for sm_100a: curved or merely non-rectangular quadrilaterals /
// hexahedra (the reference's MappingQ(fe_degree) on an arbitrary conforming triangulation), with
// arbitrary face pairing (any local face of the neighbour, reversed tangential order in 2D).
//
// Same decomposition as dgsem_stage_kernel.cu (one block = a patch of G elements, one thread per node, primitives once
// per node in shared memory, every unordered node pair once, gather-form faces, one HBM read of u and one write of dst).
// What changes with the geometry, following the reference expression by expression:
//   volume   R_j = -(1-alpha)/J_j sum_d sum_l D[j_d][l] F#(u_j,u_l) . (Ja^d_j + Ja^d_l)     split_form_volume_flux.h:68-98
//            (the pair flux is contracted with the SUM of the two contravariant vectors, still symmetric in (j,l));
//   subcell  interface normals n_i = Ja^d_0 + sum_{k<=i} sum_m Q(k,m) Ja^d_m (precomputed on the host, read only where
//            alpha > 0), flux f*(u_i,u_i+1, n/|n|) |n|                                        subcell_finite_volume_flux.h:75-158
//   faces    unit normal and surface Jacobian per face node from the host tables (one normal per face, mirrored bit for
//            bit to the other side => conservative to round-off), neighbour trace through (face, orientation) codes
//                                                                                             fluid_flux_es_dgsem_operator.h:301-342
//   dt       |J^-T u|_inf with the full matrix and the per-node power-iteration value          :476-502
// The Cartesian kernel keeps its own file: there the metric terms are compile-time structure (one flux direction per
// pair, constants folded on the host), which is worth ~35 % of its run time.
#include "dgsem_common.cuh"
#include "dgsem_stage_common.cuh"

namespace wgpu {

// f(u) . n, n not necessarily of unit length.  Odd in n bit for bit.
template <int DIM>
__device__ __forceinline__ void phys_flux_n(const double* n, const Prim& P, double F[5]) {
    double un = P.u0 * n[0];
    if (DIM > 1) un = fma(P.u1, n[1], un);
    if (DIM > 2) un = fma(P.u2, n[2], un);
    const double m = P.rho * un;
    F[0] = m;
    F[1] = fma(m, P.u0, P.p * n[0]);
    F[2] = (DIM > 1) ? fma(m, P.u1, P.p * n[1]) : m * P.u1;
    F[3] = (DIM > 2) ? fma(m, P.u2, P.p * n[2]) : m * P.u2;
    F[4] = un * P.H;
}

// Entropy-conserving two-point flux contracted with n: sum_d F#_d n_d (euler.h:186-228).  Symmetric in (a, b) and odd
// in n bit for bit; 1/beta_ln is returned for the dissipation term.
template <int DIM>
__device__ __forceinline__ void ec_flux_n(const double* n, const Prim& a, const Prim& b, const double half_inv_gm1,
                                          double F[5], double& inv_beta_ln) {
    const double s_rho = a.rho + b.rho, s_beta = a.beta + b.beta;
    const double n_rho = dmax(1e6 * fabs(b.rho - a.rho), s_rho);
    const double d_rho = dmax(1e6 * fabs(b.lrho - a.lrho), 2.0);
    const double n_beta = dmax(1e6 * fabs(b.beta - a.beta), s_beta);
    const double d_beta = dmax(1e6 * fabs(b.lbeta - a.lbeta), 2.0);
    const double r1 = rcp_pos(n_rho * d_rho);
    const double rho_ln = (n_rho * n_rho) * r1;
    const double inv_rho_ln = (d_rho * d_rho) * r1;
    const double r2 = rcp_pos(n_beta * s_beta);
    const double ibl = (d_beta * s_beta) * r2;
    inv_beta_ln = ibl;
    const double p_hat = ((0.5 * s_rho) * n_beta) * r2;
    const double U0 = a.u0 + b.u0, U1 = a.u1 + b.u1, U2 = a.u2 + b.u2;
    const double SU = U0 * U0 + U1 * U1 + U2 * U2;
    const double h_hat = ibl * half_inv_gm1 + p_hat * inv_rho_ln + 0.25 * (SU - (a.q2 + b.q2));
    const double h0 = 0.5 * U0, h1 = 0.5 * U1, h2 = 0.5 * U2;
    double hn = h0 * n[0];
    if (DIM > 1) hn = fma(h1, n[1], hn);
    if (DIM > 2) hn = fma(h2, n[2], hn);
    const double m = rho_ln * hn;
    F[0] = m;
    F[1] = fma(m, h0, p_hat * n[0]);
    F[2] = (DIM > 1) ? fma(m, h1, p_hat * n[1]) : m * h1;
    F[3] = (DIM > 2) ? fma(m, h2, p_hat * n[2]) : m * h2;
    F[4] = m * h_hat;
}

#ifndef WGPU_GENERAL_RESIDENT
#define WGPU_GENERAL_RESIDENT 0     // 0: 512 threads per SM in 1D/2D (4 blocks of 128 at 128 registers, no spills), 384 in 3D
#endif                              // measured on a 512^2 p=3 curved mesh: 384 -> 0.477 ms, 512 -> 0.393 ms, 640 -> 0.430 ms per stage

template <int DIM, int NP>
struct GeoG {
    using GEO = Geo<DIM, NP>;
    static constexpr int K = DIM * DIM;
    static constexpr int OFF_JA = GEO::SMEM_DOUBLES;                // [K][NODES] contravariant vectors of the block's nodes
    static constexpr int OFF_GF = OFF_JA + K * GEO::NODES;          // [DIM+1][NSLOT] face normals and lifting factors
    static constexpr int SMEM_DOUBLES = OFF_GF + (DIM + 1) * GEO::NSLOT;
    static constexpr int RESIDENT = WGPU_GENERAL_RESIDENT > 0 ? WGPU_GENERAL_RESIDENT : (DIM <= 2 ? 512 : 384);
    static constexpr int BLOCKS_BY_THREADS = (RESIDENT / GEO::THREADS) > 0 ? (RESIDENT / GEO::THREADS) : 1;
    static constexpr int MIN_BLOCKS = BLOCKS_BY_THREADS > 4 ? 4 : BLOCKS_BY_THREADS;   // never below 128 registers per thread
};

template <int DIM, int NP>
__global__ void __launch_bounds__(Geo<DIM, NP>::THREADS, GeoG<DIM, NP>::MIN_BLOCKS)
    stage_kernel_general(const StageParams P, const GeneralParams GP) {
    using GEO = Geo<DIM, NP>;
    constexpr int NN = GEO::NN, NF = GEO::NF, G = GEO::G, NODES = GEO::NODES, NFACE = GEO::NFACE, NSLOT = GEO::NSLOT;
    constexpr int NFULL = GEO::NFULL, HALF = GEO::HALF, NCL = GEO::NCL, PLANE = GEO::PLANE, GROUP = GEO::GROUP;
    constexpr int K = DIM * DIM;

    extern __shared__ __align__(16) double smem[];
    double* const sD = smem + GEO::OFF_D;
    double* const sV = smem + GEO::OFF_V;
    double* const sW = smem + GEO::OFF_W;
    double* const sP = smem + GEO::OFF_P;
    double* const sPair = smem + GEO::OFF_PAIR;
    double* const sA = smem + GEO::OFF_A;
    double* const sB = sA + NODES;
    double* const sFace = smem + GEO::OFF_FACE;
    double* const sAlpha = smem + GEO::OFF_ALPHA;
    double* const sRed = smem + GEO::OFF_RED;
    double* const sJa = smem + GeoG<DIM, NP>::OFF_JA;
    double* const sGF = smem + GeoG<DIM, NP>::OFF_GF;

    const int skip = P.skip_dev ? *P.skip_dev : 0;
    const double dt = P.dt_dev ? *P.dt_dev : P.dt;

    const int tid = threadIdx.x;
    const int le = tid / NN;
    const int j = tid - le * NN;
    const int i0 = j % NP, i1 = (DIM > 1) ? (j / NP) % NP : 0, i2 = (DIM > 2) ? j / (NP * NP) : 0;
    const int idx[3] = {i0, i1, i2};
    const int64_t e0 = P.elem_begin + (int64_t)blockIdx.x * G;
    const int64_t e = e0 + le;
    const bool active = e < P.elem_end;
    const int64_t e_hi = (e0 + G < P.elem_end) ? e0 + G : P.elem_end;

    const double gamma = P.gamma, gm1 = P.gamma - 1.0;
    const double hig = P.hig;
    const int nc = P.nc;
    double vmax_local = 0.0;

    // this node's metric terms: Ja[d][r] = component r of Ja^d, 1/Jdet (the same for every species)
    // (the loads are issued here; their first use - the shared-memory copy - comes after the state loads and the face
    // prefetches of the first species have been issued, so that the block pays one memory latency at its start)
    double Ja[DIM][DIM], invJ = 1.0, eig = 0.0;
#pragma unroll
    for (int d = 0; d < DIM; d++)
#pragma unroll
        for (int r = 0; r < DIM; r++) Ja[d][r] = (d == r) ? 1.0 : 0.0;
    if (active) {
        const double* g = GP.gnode + ((size_t)e * (K + 2)) * NN + j;
#pragma unroll
        for (int d = 0; d < DIM; d++)
#pragma unroll
            for (int r = 0; r < DIM; r++) Ja[d][r] = g[(size_t)(d * DIM + r) * NN];
        invJ = g[(size_t)K * NN];
        if (P.vmax && P.mode == 0) eig = g[(size_t)(K + 1) * NN];
    }
    // the face tables of this thread's face tasks travel to shared memory while the node phase computes (they are the
    // same for every species; the first species' cp.async group carries them)
    if (active) {
        for (int ft = j; ft < NFACE * NF; ft += NN) {
            const int f = ft / NF, t = ft - f * NF;
            const double* gf = GP.gface + ((size_t)(e * NFACE + f) * (DIM + 1)) * NF + t;
            double* const dstg = sGF + (le * NFACE + f) * NF + t;
#pragma unroll
            for (int r = 0; r <= DIM; r++) cp_async8(dstg + r * NSLOT, gf + (size_t)r * NF);
        }
    }

    for (int sp = 0; sp < P.nsp; sp++) {
        if (sp > 0) __syncthreads();
        // ---- node phase ----------------------------------------------------------------------------------------
        const size_t off = ((size_t)e * nc + 5 * sp) * NN + j;
        double q[5] = {1.0, 0.0, 0.0, 0.0, 1.0};
        if (active) {
#pragma unroll
            for (int c = 0; c < 5; c++) q[c] = P.u[off + (size_t)c * NN];
            for (int ft = j; ft < NFACE * NF; ft += NN) {
                const int f = ft / NF, t = ft - f * NF;
                const int2 vc = GP.nbr2[(size_t)e * NFACE + f];   // (neighbour, its face + orientation): one load
                const int v = vc.x;
                if (v >= e0 && v < e_hi) continue;
                double* const rec = sFace + (le * NFACE + f) * NF + t;
                const double* src;
                size_t stride;
                if (v < 0) {
                    src = P.bres + (((size_t)(-1 - v) * P.nsp + sp) * 5) * NF + t;
                    stride = NF;
                } else {
                    const int code = vc.y;
                    const int nfa = code & 7, tp = (code >> 3) ? NF - 1 - t : t;
                    if (v < P.n_elems) {
                        src = P.u + ((size_t)v * nc + 5 * sp) * NN + node_of_face_node<DIM, NP>(nfa >> 1, nfa & 1, tp);
                        stride = NN;
                    } else {
                        src = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + 5 * sp) * NF + tp;
                        stride = NF;
                    }
                }
#pragma unroll
                for (int c = 0; c < 5; c++) cp_async8(rec + c * NSLOT, src + (size_t)c * stride);
            }
            cp_async_commit();
        }
        if (sp == 0) {
            if (skip) {
                cp_async_wait_all();
                return;
            }
#pragma unroll
            for (int d = 0; d < DIM; d++)
#pragma unroll
                for (int r = 0; r < DIM; r++) sJa[(d * DIM + r) * NODES + tid] = Ja[d][r];
            for (int i = tid; i < NP * NP; i += NODES) { sD[i] = P.T.D[i]; sV[i] = P.T.V[i]; }
            for (int i = tid; i < NP; i += NODES) sW[i] = P.T.w[i];
        }
        Prim mine;
        {
            const Prim me = make_prim(q[0], q[1], q[2], q[3], q[4], gamma);
            store_prim<NP>(sP, tid, me);
            sA[tid] = me.p * me.rho;
            mine = me;
        }
        __syncthreads();

        // ---- pair rounds, full classes: F#(u_j,u_l) . (Ja^d_j + Ja^d_l), once per unordered pair -------------------
        {
            double* const slot = sPair + (le * DIM * NCL) * NN + j;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                const int st = stride_of(NP, d);
#pragma unroll
                for (int c = 1; c <= NFULL; c++) {
                    int l = idx[d] + c;
                    if (l >= NP) l -= NP;
                    const int pt = tid + (l - idx[d]) * st;
                    const Prim other = load_prim_ec<NP>(sP, pt);
                    double n[DIM];
#pragma unroll
                    for (int r = 0; r < DIM; r++) n[r] = Ja[d][r] + sJa[(d * DIM + r) * NODES + pt];
                    double F[5], ibl;
                    ec_flux_n<DIM>(n, mine, other, hig, F, ibl);
                    double* const o = slot + (d * NCL + (c - 1)) * NN;
#pragma unroll
                    for (int qq = 0; qq < 5; qq++) o[qq * PLANE] = F[qq];
                }
            }
        }

        // ---- shock indicator (identical to the Cartesian kernel: the modal analysis lives on the reference element) --
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
                group_sync<GROUP, NODES>(tid);
                double* t = src; src = dstb; dstb = t;
            }
            const double ck = src[tid];
            dstb[tid] = ck * ck;
            group_sync<GROUP, NODES>(tid);
            if (i0 == 0) {
                double g0 = 0.0, g1 = 0.0;
                const bool row_shell = (i1 == NP - 1) || (i2 == NP - 1);
#pragma unroll
                for (int m = 0; m < NP; m++) {
                    const double v = dstb[tid + m];
                    if (row_shell || m == NP - 1) g1 += v; else g0 += v;
                }
                src[tid] = g0;
                src[tid + 1] = g1;
            }
            group_sync<GROUP, NODES>(tid);
            if (j == 0) {
                double g0 = 0.0, g1 = 0.0;
                for (int r = 0; r < NF; r++) { g0 += src[tid + r * NP]; g1 += src[tid + r * NP + 1]; }
                const double al = blending_from_energies(g0, g1, P.ind_T, P.ind_sT);
                sAlpha[le] = al;
                if (active && P.alpha_out) P.alpha_out[(size_t)e * P.nsp + sp] = al;
            }
            group_sync<GROUP, NODES>(tid);
        }

        // ---- half class of an even NP ---------------------------------------------------------------------------------
        if (HALF) {
#pragma unroll
            for (int r = 0; r < GEO::HROUNDS; r++) {
                const int p = r * NN + j;
                if (GEO::HTASKS % NN == 0 || p < GEO::HTASKS) {
                    const bool hi = j >= NN / 2;
                    const int qq = hi ? j - NN / 2 : j;
                    const int d = 2 * r + (hi ? 1 : 0);
                    constexpr int DMAX = DIM - 1;
                    const int n0 = hi ? half_class_node<DIM, NP>((2 * r + 1 < DIM) ? 2 * r + 1 : DMAX, qq)
                                      : half_class_node<DIM, NP>((2 * r < DIM) ? 2 * r : DMAX, qq);
                    const int st = stride_of(NP, d);
                    const int na = le * NN + n0, nb = na + (NP / 2) * st;
                    const Prim a = load_prim_ec<NP>(sP, na);
                    const Prim b = load_prim_ec<NP>(sP, nb);
                    double n[DIM];
#pragma unroll
                    for (int rr = 0; rr < DIM; rr++) n[rr] = sJa[(d * DIM + rr) * NODES + na] + sJa[(d * DIM + rr) * NODES + nb];
                    double F[5], ibl;
                    ec_flux_n<DIM>(n, a, b, hig, F, ibl);
                    double* const o = sPair + ((le * DIM + d) * NCL + NFULL) * NN + n0;
#pragma unroll
                    for (int c = 0; c < 5; c++) o[c * PLANE] = F[c];
                }
            }
        }
        // ---- faces: the element's own side of every face ------------------------------------------------------------
        cp_async_wait_all();
        for (int ft = j; ft < NFACE * NF; ft += NN) {
            const int f = ft / NF, t = ft - f * NF;
            const int d = f >> 1, side = f & 1;
            double* const rec = sFace + (le * NFACE + f) * NF + t;
            const int2 vc = active ? GP.nbr2[(size_t)e * NFACE + f] : make_int2((int)e0, f ^ 1);
            const int v = vc.x;
            if (v < 0) continue;
            const int code = vc.y;
            const int nfa = code & 7, tp = (code >> 3) ? NF - 1 - t : t;
            double n[DIM], lift = 0.0;
#pragma unroll
            for (int r = 0; r < DIM; r++) n[r] = (r == d) ? 1.0 : 0.0;
            if (active) {
                const double* gf = sGF + (le * NFACE + f) * NF + t;
#pragma unroll
                for (int r = 0; r < DIM; r++) n[r] = gf[r * NSLOT];
                lift = gf[DIM * NSLOT];
            }
            const Prim a = load_prim<NP>(sP, le * NN + node_of_face_node<DIM, NP>(d, side, t));
            Prim b;
            if (v >= e0 && v < e_hi) {
                b = load_prim<NP>(sP, (int)(v - e0) * NN + node_of_face_node<DIM, NP>(nfa >> 1, nfa & 1, tp));
            } else {
                double qn[5];
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = rec[c * NSLOT];
                b = make_prim(qn[0], qn[1], qn[2], qn[3], qn[4], gamma);
            }
            double Fe[5], Dv[5], Fm[5], ibl;
            ec_flux_n<DIM>(n, a, b, hig, Fe, ibl);
            es_dissipation(a, b, ibl, hig, Dv);
            phys_flux_n<DIM>(n, a, Fm);
            // (f(u_m).n - f*(u_m,u_p,n)) * face JxW / cell JxW,  f* = F#.n - D   (fluid_flux_es_dgsem_operator.h:318-333)
#pragma unroll
            for (int c = 0; c < 5; c++) rec[c * NSLOT] = lift * ((Fm[c] - Fe[c]) + Dv[c]);
        }
        double dw[DIM][NP - 1];
        int poff[DIM][NP - 1];
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int jd = idx[d];
            const int st = stride_of(NP, d);
#pragma unroll
            for (int k = 1; k < NP; k++) {
                int l = jd + k;
                if (l >= NP) l -= NP;
                dw[d][k - 1] = sD[jd * NP + l];
                const int c = (k < NP - k) ? k : NP - k;
                const bool own_first = (2 * k < NP) || (2 * k == NP && jd < NP / 2);
                poff[d][k - 1] = ((le * DIM + d) * NCL + (c - 1)) * NN + (own_first ? j : j + (l - jd) * st);
            }
        }
        group_sync<GROUP, NODES>(tid);

        // ---- node phase 2 -----------------------------------------------------------------------------------------------
        const double alpha = sAlpha[le];
        const Prim me = load_prim_phys<NP>(sP, tid);
        double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int jd = idx[d];
            double acc[5];
            {   // l = j: F#(u,u) . (2 Ja^d_j) = 2 f(u) . Ja^d_j; its weight D[j][j] vanishes except at the two end nodes
                const double djj = 2.0 * sD[jd * NP + jd];
                double Fp[5];
                phys_flux_n<DIM>(Ja[d], me, Fp);
#pragma unroll
                for (int c = 0; c < 5; c++) acc[c] = djj * Fp[c];
            }
#pragma unroll
            for (int k = 0; k < NP - 1; k++) {
#pragma unroll
                for (int c = 0; c < 5; c++) acc[c] = fma(dw[d][k], sPair[c * PLANE + poff[d][k]], acc[c]);
            }
#pragma unroll
            for (int c = 0; c < 5; c++) r[c] = fma(-invJ, acc[c], r[c]);
        }
        if (alpha > 0.0) {
            const Prim mef = load_prim<NP>(sP, tid);
            const double oma = 1.0 - alpha;
#pragma unroll
            for (int c = 0; c < 5; c++) r[c] *= oma;
            const double* const gs = GP.gsub + ((size_t)e * K) * NN + j;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                const int jd = idx[d];
                const int st = stride_of(NP, d);
                double Ln[DIM], Rn[DIM];
#pragma unroll
                for (int rr = 0; rr < DIM; rr++) {
                    Rn[rr] = active ? gs[(ptrdiff_t)(d * DIM + rr) * NN] : Ja[d][rr];
                    Ln[rr] = (active && jd > 0) ? gs[(ptrdiff_t)(d * DIM + rr) * NN - st] : Ja[d][rr];
                }
                double left[5], right[5];
                if (jd > 0) {
                    const Prim o = load_prim<NP>(sP, tid - st);
                    double Fd[5], Dv[5], ibl, l2 = 0.0;
                    ec_flux_n<DIM>(Ln, o, mef, hig, Fd, ibl);
                    es_dissipation(o, mef, ibl, hig, Dv);
#pragma unroll
                    for (int rr = 0; rr < DIM; rr++) l2 = fma(Ln[rr], Ln[rr], l2);
                    const double len = sqrt_pos(l2);
#pragma unroll
                    for (int c = 0; c < 5; c++) left[c] = fma(-len, Dv[c], Fd[c]);
                } else {
                    phys_flux_n<DIM>(Ln, mef, left);
                }
                if (jd < NP - 1) {
                    const Prim o = load_prim<NP>(sP, tid + st);
                    double Fd[5], Dv[5], ibl, l2 = 0.0;
                    ec_flux_n<DIM>(Rn, mef, o, hig, Fd, ibl);
                    es_dissipation(mef, o, ibl, hig, Dv);
#pragma unroll
                    for (int rr = 0; rr < DIM; rr++) l2 = fma(Rn[rr], Rn[rr], l2);
                    const double len = sqrt_pos(l2);
#pragma unroll
                    for (int c = 0; c < 5; c++) right[c] = fma(-len, Dv[c], Fd[c]);
                } else {
                    phys_flux_n<DIM>(Rn, mef, right);
                }
                const double cf = alpha * invJ / sW[jd];
#pragma unroll
                for (int c = 0; c < 5; c++) r[c] += cf * (left[c] - right[c]);
            }
        }
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int t = face_node_index<DIM, NP>(d, i0, i1, i2);
            if (idx[d] == 0 || idx[d] == NP - 1) {
                const int f = 2 * d + (idx[d] == 0 ? 0 : 1);
                const double* const rec = sFace + (le * NFACE + f) * NF + t;
#pragma unroll
                for (int c = 0; c < 5; c++) r[c] += rec[c * NSLOT];
            }
        }

        if (P.src_on && active) {
            const size_t fo = ((size_t)e * nc + 5 * P.nsp) * NN + j;
            const double Ex = P.u[fo], Ey = P.u[fo + NN], Ez = P.u[fo + 2 * (size_t)NN];
            const double Bx = P.u[fo + 3 * (size_t)NN], By = P.u[fo + 4 * (size_t)NN], Bz = P.u[fo + 5 * (size_t)NN];
            const double qm = P.qm[sp];
            r[1] += qm * (q[0] * Ex + (q[2] * Bz - q[3] * By));
            r[2] += qm * (q[0] * Ey + (q[3] * Bx - q[1] * Bz));
            r[3] += qm * (q[0] * Ez + (q[1] * By - q[2] * Bx));
            r[4] += qm * (q[1] * Ex + q[2] * Ey + q[3] * Ez);
        }

        if (active) {
            double qn[5];
            if (P.mode == 1) {
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = r[c];
            } else if (P.mode == 2) {   // low-storage RK stage (rk.h:53-71; tutorial-67.cc:880-899): with s = old solution,
                                        // solution = s + b_i dt k, next r = s + a_i dt k (not written when a_i = 0)
#pragma unroll
                for (int c = 0; c < 5; c++) {
                    const double s0 = P.sol_in[off + (size_t)c * NN];
                    qn[c] = fma(P.a, r[c], s0);
                    if (P.beta != 0.0) P.dst2[off + (size_t)c * NN] = fma(P.beta, r[c], s0);
                }
            } else if (P.beta == 0.0) {
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = P.a * (q[c] + dt * r[c]);
            } else {
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = P.beta * P.dst[off + (size_t)c * NN] + P.a * (q[c] + dt * r[c]);
            }
#pragma unroll
            for (int c = 0; c < 5; c++) P.dst[off + (size_t)c * NN] = qn[c];

            if (P.vmax && P.mode == 0) {
                // compute_cell_transport_speed (:450-514): J^-T[r][c] = Ja^c[r] / Jdet
                const double inv = rcp_pos(qn[0]);
                const double sm = qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3];
                const double pr = gm1 * (qn[4] - sm * (0.5 * inv));
                const double vel[3] = {qn[1] * inv, qn[2] * inv, qn[3] * inv};
                double conv = 0.0;
#pragma unroll
                for (int rr = 0; rr < DIM; rr++) {
                    double s = 0.0;
#pragma unroll
                    for (int c = 0; c < DIM; c++) s = fma(Ja[c][rr], vel[c], s);
                    conv = fmax(conv, fabs(s * invJ));
                }
                const double c2 = gamma * pr * inv;
                const double cs = (c2 > 0.0) ? sqrt_pos(c2) : sqrt(c2 - 1.0);
                const double speed = eig * cs + conv;
                vmax_local = (speed > vmax_local || speed != speed) ? speed : vmax_local;
            }
        }
    }

    if (active && nc > 5 * P.nsp) {
        double S[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
        if (P.src_on) {
            double Jx = 0.0, Jy = 0.0, Jz = 0.0, rc = 0.0;
            for (int sp = 0; sp < P.nsp; sp++) {
                const size_t so = ((size_t)e * nc + 5 * sp) * NN + j;
                const double qm = P.qm[sp];
                rc += qm * P.u[so];
                Jx += qm * P.u[so + NN];
                Jy += qm * P.u[so + 2 * (size_t)NN];
                Jz += qm * P.u[so + 3 * (size_t)NN];
            }
            S[0] = -Jx * P.inv_eps0;
            S[1] = -Jy * P.inv_eps0;
            S[2] = -Jz * P.inv_eps0;
            S[6] = P.chi * rc * P.inv_eps0;
        }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (5 * P.nsp + k >= nc) break;
            const size_t off = ((size_t)e * nc + 5 * P.nsp + k) * NN + j;
            const double rate = S[k];
            double v;
            if (P.mode == 1) v = rate;
            else if (P.mode == 2) {
                const double s0 = P.sol_in[off];
                v = fma(P.a, rate, s0);
                if (P.beta != 0.0) P.dst2[off] = fma(P.beta, rate, s0);
            }
            else if (P.beta == 0.0) v = P.a * (P.u[off] + dt * rate);
            else v = P.beta * P.dst[off] + P.a * (P.u[off] + dt * rate);
            P.dst[off] = v;
        }
    }

    if (P.vmax && P.mode == 0) {
        const double m = block_max(vmax_local, sRed);
        if (tid == 0) atomicMax(P.vmax, (unsigned long long)__double_as_longlong(m));
    }
}

// --------------------------------------------------------------------------------------------------------
// Boundary faces on general geometry (fluid_flux_es_dgsem_operator.h:344-440): as boundary_kernel, with the unit
// normal and surface Jacobian of every Gauss(p+2) point from the host tables.
// --------------------------------------------------------------------------------------------------------
template <int DIM>
__device__ __forceinline__ void lf_flux_n(const double* n, const double* qi, const double* qo, const double gamma,
                                          double F[5], double Fin[5]) {
    const Prim a = make_prim(qi[0], qi[1], qi[2], qi[3], qi[4], gamma);
    const Prim b = make_prim(qo[0], qo[1], qo[2], qo[3], qo[4], gamma);
    double Fout[5];
    phys_flux_n<DIM>(n, a, Fin);
    phys_flux_n<DIM>(n, b, Fout);
    double nsq_in = a.u0 * a.u0, nsq_out = b.u0 * b.u0;   // dim-component speed (euler.h:73-80)
    if (DIM > 1) { nsq_in += a.u1 * a.u1; nsq_out += b.u1 * b.u1; }
    if (DIM > 2) { nsq_in += a.u2 * a.u2; nsq_out += b.u2 * b.u2; }
    const double lambda = 0.5 * sqrt(fmax(nsq_out + gamma * b.p / b.rho, nsq_in + gamma * a.p / a.rho));
#pragma unroll
    for (int c = 0; c < 5; c++) F[c] = 0.5 * (Fin[c] + Fout[c]) + 0.5 * lambda * (qi[c] - qo[c]);
}

constexpr int kBoundaryWarpsG = 4;
template <int DIM, int NP>
__global__ void __launch_bounds__(32 * kBoundaryWarpsG) boundary_kernel_general(const BoundaryParams P, const GeneralParams GP) {
    constexpr int NN = ipow_c(NP, DIM), NF = ipow_c(NP, DIM - 1), NG1 = NP + 1, NG = ipow_c(NG1, DIM - 1);
    __shared__ double sv[kBoundaryWarpsG][NG][10];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t gid = (int64_t)blockIdx.x * kBoundaryWarpsG + warp;
    if (gid >= P.n_bfaces * P.nsp) return;
    const int64_t bf = gid / P.nsp;
    const int sp = (int)(gid - bf * P.nsp);
    const int e = P.bf_elem[bf], f = P.bf_side[bf], bid = P.bf_id[bf];
    const int d = f / 2, side = f % 2;
    const int kind = P.bc_kind[sp * P.n_boundaries + bid];

    for (int g = lane; g < NG; g += 32) {
        const int g0 = g % NG1, g1 = (g / NG1) % NG1;
        double wm[5] = {0, 0, 0, 0, 0};
        for (int t = 0; t < NF; t++) {
            const int t0 = t % NP, t1 = (t / NP) % NP;
            double phi = 1.0;
            if (DIM >= 2) phi *= P.Ig[g0 * NP + t0];
            if (DIM >= 3) phi *= P.Ig[g1 * NP + t1];
            const int node = node_of_face_node<DIM, NP>(d, side, t);
            const size_t off = ((size_t)e * P.nc + 5 * sp) * NN + node;
            for (int c = 0; c < 5; c++) wm[c] += phi * P.u[off + (size_t)c * NN];
        }
        const double* bg = GP.bgeo + ((size_t)bf * NG + g) * (DIM + 1);
        double n[DIM];
        for (int a = 0; a < DIM; a++) n[a] = bg[a];
        const double sJ = bg[DIM];
        double wp[5];
        if (kind == 2 || kind == 3) {
            if (P.inflow_table) {
                const double* tab = P.inflow_table + (((size_t)sp * P.n_bfaces + bf) * NG + g) * 5;
                for (int c = 0; c < 5; c++) wp[c] = tab[c];
            } else {
                for (int c = 0; c < 5; c++) wp[c] = P.inflow[((size_t)sp * P.n_boundaries + bid) * 5 + c];
            }
            if (kind == 3) {   // subsonic outflow (:385-390): w_p = w_m with the total energy replaced
                for (int c = 0; c < 4; c++) wp[c] = wm[c];
            }
        } else if (kind == 1) {
            for (int c = 0; c < 5; c++) wp[c] = wm[c];
        } else {   // wall (:394-405): reflect the normal momentum over the first dim components, zero the rest
            double rho_u_dot_n = wm[1] * n[0];
            for (int a = 1; a < DIM; a++) rho_u_dot_n += wm[1 + a] * n[a];
            wp[0] = wm[0];
            for (int a = 0; a < 3; a++) wp[a + 1] = (a < DIM) ? wm[a + 1] - 2.0 * rho_u_dot_n * n[a < DIM ? a : 0] : 0.0;
            wp[4] = wm[4];
        }
        double Fs[5], Fm[5];
        lf_flux_n<DIM>(n, wm, wp, P.gamma, Fs, Fm);
        double wq = sJ;
        if (DIM >= 2) wq *= P.wg[g0];
        if (DIM >= 3) wq *= P.wg[g1];
        for (int c = 0; c < 5; c++) {
            sv[warp][g][c] = (Fm[c] - Fs[c]) * wq;
            sv[warp][g][5 + c] = Fs[c] * wq;
        }
    }
    __syncwarp();
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
        const double cf = GP.bmass[(size_t)bf * NF + t];
        for (int c = 0; c < 5; c++) P.bres[((size_t)(bf * P.nsp + sp) * 5 + c) * NF + t] = acc[c] * cf;
    }
    if (lane < 5) {
        double bsum = 0.0;
        for (int g = 0; g < NG; g++) bsum += sv[warp][g][5 + lane];
        P.bflux[(size_t)(bf * P.nsp + sp) * 5 + lane] = bsum;
    }
}

template <int DIM, int NP>
__global__ void cfl_kernel_general(const double* __restrict__ u, int64_t n_elems, int nc, int nsp, double gamma,
                                   const double* __restrict__ gnode, unsigned long long* vmax) {
    constexpr int NN = ipow_c(NP, DIM), K = DIM * DIM;
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
        const double vel[3] = {q1 * inv, q2 * inv, q3 * inv};
        const double* g = gnode + ((size_t)e * (K + 2)) * NN + j;
        const double invJ = g[(size_t)K * NN], eig = g[(size_t)(K + 1) * NN];
        double conv = 0.0;
        for (int r = 0; r < DIM; r++) {
            double s = 0.0;
            for (int c = 0; c < DIM; c++) s = fma(g[(size_t)(c * DIM + r) * NN], vel[c], s);
            conv = fmax(conv, fabs(s * invJ));
        }
        m = nan_max(m, eig * sqrt(gamma * pr * inv) + conv);
    }
    m = block_max(m, sRed);
    if (threadIdx.x == 0) atomicMax(vmax, (unsigned long long)__double_as_longlong(m));
}

template <int DIM, int NP>
__global__ void integral_partial_general(const double* __restrict__ u, int64_t n_elems, int nc, int species,
                                         const double* __restrict__ jdet, const double* __restrict__ w1, double* partial) {
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
            cell += ue[j] * (jdet[(size_t)e * NN + j] * wj);
        }
        s += cell;
    }
    partial[(size_t)blockIdx.x * 5 + c] = s;
}
__global__ void integral_final_general(const double* partial, int n_blocks, double* out) {
    const int c = threadIdx.x;
    if (c >= 5) return;
    double s = 0.0;
    for (int b = 0; b < n_blocks; b++) s += partial[(size_t)b * 5 + c];
    out[c] = s;
}

int stage_general_smem_bytes(int dim, int Np) {
    int bytes = 0;
#define CALL(D_, N_) { bytes = GeoG<D_, N_>::SMEM_DOUBLES * (int)sizeof(double); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
    return bytes;
}

int prepare_general_kernels(int dim, int Np) {
    cudaError_t err = cudaSuccess;
#define CALL(D_, N_)                                                                                                  \
    {                                                                                                                 \
        err = cudaFuncSetAttribute(stage_kernel_general<D_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                   GeoG<D_, N_>::SMEM_DOUBLES * (int)sizeof(double));                                 \
        if (err == cudaSuccess)                                                                                       \
            err = cudaFuncSetAttribute(stage_kernel_general<D_, N_>, cudaFuncAttributePreferredSharedMemoryCarveout,  \
                                       cudaSharedmemCarveoutMaxShared);                                               \
    }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
    return err == cudaSuccess ? 0 : 1;
}

void launch_stage_general(int dim, int Np, const StageParams& P, const GeneralParams& GP, cudaStream_t s) {
    const int64_t n = P.elem_end - P.elem_begin;
    if (n <= 0) return;
#define CALL(D_, N_)                                                                                               \
    {                                                                                                              \
        using GEO = Geo<D_, N_>;                                                                                   \
        const int64_t blocks = (n + GEO::G - 1) / GEO::G;                                                          \
        stage_kernel_general<D_, N_>                                                                               \
            <<<(unsigned)blocks, GEO::THREADS, GeoG<D_, N_>::SMEM_DOUBLES * sizeof(double), s>>>(P, GP);           \
    }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

void launch_boundary_general(int dim, int Np, const BoundaryParams& P, const GeneralParams& GP, cudaStream_t s) {
    const int64_t n = P.n_bfaces * P.nsp;
    if (n <= 0) return;
#define CALL(D_, N_) { boundary_kernel_general<D_, N_><<<(unsigned)((n + kBoundaryWarpsG - 1) / kBoundaryWarpsG), 32 * kBoundaryWarpsG, 0, s>>>(P, GP); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

void launch_cfl_general(int dim, int Np, const double* u, int64_t n_elems, int nc, int nsp, double gamma,
                        const GeneralParams& GP, unsigned long long* vmax, cudaStream_t s) {
    if (n_elems <= 0) return;
#define CALL(D_, N_) { cfl_kernel_general<D_, N_><<<148 * 8, 256, 0, s>>>(u, n_elems, nc, nsp, gamma, GP.gnode, vmax); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

void launch_integral_general(int dim, int Np, const double* u, int64_t n_elems, int nc, int species, const GeneralParams& GP,
                             const double* w, double* partial, double* out, cudaStream_t s) {
    const int nb = integral_blocks(n_elems);
#define CALL(D_, N_) { integral_partial_general<D_, N_><<<nb, 32, 0, s>>>(u, n_elems, nc, species, GP.jdet, w, partial); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
    integral_final_general<<<1, 32, 0, s>>>(partial, nb, out);
}

}  // namespace wgpu
