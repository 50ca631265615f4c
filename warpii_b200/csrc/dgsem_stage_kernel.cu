// The fused ES-DGSEM stage kernel for sm_100a (FP64 on the CUDA cores).
//
// One launch does, for every element, what the reference does in separate global sweeps
// (fluid_flux_es_dgsem_operator.h:127-214): gather, shock indicator (persson_peraire_shock_indicator.h),
// split-form volume term (split_form_volume_flux.h), subcell-FV blend (subcell_finite_volume_flux.h), both
// sides' face lifting (:301-342), inverse diagonal mass, the RK stage update and (optionally) the CFL
// reduction of the updated state (:450-514): one HBM read of u, one write of dst.
//
// Work decomposition inside a block (a patch of G elements, NODES = G * Np^dim threads, one thread per node):
//   node phase   load the 5 conserved values (coalesced), form the node's primitives, logarithms and wave speed
//                once, park them in shared memory (node-major, 16-byte vector access).  ONE block barrier: from here
//                on a thread only reads the block's primitive table and writes its own element's scratch, so the
//                rest is synchronised per GROUP (the warp(s) that hold whole elements: __syncwarp for Np^dim | 32, a
//                named barrier for Np^dim a multiple of 32, the block barrier otherwise) and warps never wait for
//                the slowest warp of the block;
//   indicator    sum-factorised Legendre analysis of p*rho, alpha per element;
//   task phase   (a) every UNORDERED node pair of a pencil once (the two-point flux is symmetric), grouped by cyclic
//                distance: each thread evaluates the pairs (own node, own node + c) and parks the flux in its own slot
//                of a component-plane table, so stores and the later gathers are bank-conflict free and only the
//                partner's record is loaded; (b) the element's face nodes; a neighbour inside the patch is read from
//                the shared primitive table, others were prefetched from L2/HBM (or the NCCL ghost buffer) by cp.async;
//   node phase 2 gather D-weighted pair fluxes, FV differences (only where alpha > 0) and face terms, update,
//                store, CFL.
#include "dgsem_common.cuh"
#include "dgsem_stage_common.cuh"

namespace wgpu {

template <int DIM, int NP>
using GeoC = Geo<DIM, NP, shuffle_pairs(DIM, NP)>;

template <int DIM, int NP>
__global__ void __launch_bounds__(GeoC<DIM, NP>::THREADS, GeoC<DIM, NP>::MIN_BLOCKS) stage_kernel(const StageParams P) {
    using GEO = GeoC<DIM, NP>;
    constexpr bool SHUF = shuffle_pairs(DIM, NP);
    constexpr int NN = GEO::NN, NF = GEO::NF, G = GEO::G, NODES = GEO::NODES, NFACE = GEO::NFACE, NSLOT = GEO::NSLOT;
    constexpr int NFULL = GEO::NFULL, HALF = GEO::HALF, NCL = GEO::NCL, PLANE = GEO::PLANE, GROUP = GEO::GROUP;

    extern __shared__ __align__(16) double smem[];
    double* const sD = smem + GEO::OFF_D;
    double* const sV = smem + GEO::OFF_V;
    double* const sW = smem + GEO::OFF_W;
    double* const sP = smem + GEO::OFF_P;
    double* const sPair = smem + GEO::OFF_PAIR;
    double* const sA = smem + GEO::OFF_A;   // indicator scratch; only the owning group touches its nodes' slots
    double* const sB = sA + NODES;
    double* const sFace = smem + GEO::OFF_FACE;
    double* const sAlpha = smem + GEO::OFF_ALPHA;
    double* const sRed = smem + GEO::OFF_RED;

    // device-resident time loop: "finished" flag and dt live in global memory.  The loads are issued here, the flag is
    // tested only after this thread's state and neighbour-table loads are in flight, so that the block pays one
    // memory latency at its start instead of two.
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
    const int64_t e_hi = (e0 + G < P.elem_end) ? e0 + G : P.elem_end;   // elements [e0, e_hi) have primitives in shared memory

    const double gamma = P.gamma, gm1 = P.gamma - 1.0;
    const double hig = P.hig;   // 1 / (2 (gamma - 1)), formed on the host
    const int nc = P.nc;
    double vmax_local = 0.0;

    for (int sp = 0; sp < P.nsp; sp++) {
        if (sp > 0) __syncthreads();   // shared-memory reuse across species
        // ---- node phase: load, primitives, logs, wave speed ----------------------------------------------
        const size_t off = ((size_t)e * nc + 5 * sp) * NN + j;
        double q[5] = {1.0, 0.0, 0.0, 0.0, 1.0};
        if (active) {
#pragma unroll
            for (int c = 0; c < 5; c++) q[c] = P.u[off + (size_t)c * NN];
            // Prefetch what this thread's face tasks will need from outside the patch (neighbour traces, ghost traces,
            // boundary contributions) straight into the task's own result record: the copies are in flight during the
            // node, indicator and pair phases instead of stalling the face phase.
            for (int ft = j; ft < NFACE * NF; ft += NN) {
                const int f = ft / NF, t = ft - f * NF;
                const int v = P.nbr[(size_t)e * NFACE + f];
                if (v >= e0 && v < e_hi) continue;
                double* const rec = sFace + (le * NFACE + f) * NF + t;   // component c at rec[c * NSLOT]
                const double* src;
                size_t stride;
                if (v < 0) {
                    src = P.bres + (((size_t)(-1 - v) * P.nsp + sp) * 5) * NF + t;
                    stride = NF;
                } else if (v < P.n_elems) {
                    src = P.u + ((size_t)v * nc + 5 * sp) * NN + node_of_face_node<DIM, NP>(f >> 1, 1 - (f & 1), t);
                    stride = NN;
                } else {
                    src = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + 5 * sp) * NF + t;
                    stride = NF;
                }
#pragma unroll
                for (int c = 0; c < 5; c++) cp_async8(rec + c * NSLOT, src + (size_t)c * stride);
            }
            cp_async_commit();
        }
        if (sp == 0) {
            if (skip) {   // uniform: the device-resident time loop has finished
                cp_async_wait_all();
                return;
            }
            // small tables, filled while the loads above are in flight (visible after the barrier below)
            for (int i = tid; i < NP * NP; i += NODES) { sD[i] = P.T.D[i]; sV[i] = P.T.V[i]; }
            for (int i = tid; i < NP; i += NODES) sW[i] = P.T.w[i];
        }
        Prim mine;   // this node's record stays in registers for the pair rounds that follow the barrier
        {
            const Prim me = make_prim(q[0], q[1], q[2], q[3], q[4], gamma);
            store_prim<NP>(sP, tid, me);
            sA[tid] = me.p * me.rho;   // indicator variable, fluid_flux_es_dgsem_operator.h:286-290
            mine = me;
        }
        __syncthreads();

        // ---- task phase (a), full classes: unordered node pairs of the pencils, symmetric two-point flux, once ------
        // Full classes: this thread's own node is the first endpoint, so only the partner's record is loaded; the five
        // components go to this node's slot of the class (component planes: consecutive lanes, consecutive words).
        // SHUF: accS[d][.] collects sum_l D[j_d][l] F#(u_j,u_l) in the order of the gather below (l = j+1, j+2, j+3 mod NP)
        double accS[SHUF ? DIM : 1][5];
        if (SHUF) {
            const int lane = tid & 31;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                const int st = stride_of(NP, d);
                const int jd = idx[d];
                const int lp = (jd + 1) & (NP - 1), lm = (jd + NP - 1) & (NP - 1), lh = jd ^ (NP / 2);
                // own pair (j, j+1): computed here, also needed by the node j+1, which fetches it from this lane
                const Prim other = load_prim_ec<NP>(sP, tid + (lp - jd) * st);
                double F[5], ibl;
                ec_flux_d(d, mine, other, hig, F, ibl);
                // half pair (j, j+2): both endpoints evaluate it (symmetric bit for bit; a warp instruction costs the same
                // whether 16 or 32 lanes take part)
                const Prim oh = load_prim_ec<NP>(sP, tid + (lh - jd) * st);
                double Fh[5];
                ec_flux_d(d, mine, oh, hig, Fh, ibl);
                const double wp = sD[jd * NP + lp], wh = sD[jd * NP + lh], wm = sD[jd * NP + lm], djj = sD[jd * NP + jd];
                const int src = lane + (lm - jd) * st;   // the lane of node j-1 (cyclic) of this pencil
                double Fp[5];
                phys_flux_d(d, mine, Fp);   // F#(u,u) = f(u), weight D[j][j]
#pragma unroll
                for (int q = 0; q < 5; q++) {
                    const double Fm = __shfl_sync(0xffffffffu, F[q], src);
                    accS[d][q] = fma(wm, Fm, fma(wh, Fh[q], fma(wp, F[q], djj * Fp[q])));   // same order as the table gather
                }
            }
        } else {
            double* const slot = sPair + (le * DIM * NCL) * NN + j;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                const int st = stride_of(NP, d);
#pragma unroll
                for (int c = 1; c <= NFULL; c++) {
                    int l = idx[d] + c;
                    if (l >= NP) l -= NP;
                    const Prim other = load_prim_ec<NP>(sP, tid + (l - idx[d]) * st);
                    double F[5], ibl;
                    ec_flux_d(d, mine, other, hig, F, ibl);
                    double* const o = slot + (d * NCL + (c - 1)) * NN;
#pragma unroll
                    for (int q = 0; q < 5; q++) o[q * PLANE] = F[q];
                }
            }
        }

        // ---- shock indicator: sum-factorised Legendre analysis of p*rho (group-local from here on) ------------
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
            // src holds the modal coefficients c_k at k = (i0,i1,i2); fixed-order sums: along i0, then the rest
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
                src[tid + 1] = g1;   // NP >= 2
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

        // ---- task phase ---------------------------------------------------------------------------------------
        // (a) unordered node pairs of the pencils of this thread's element: symmetric two-point flux, once
        // Half class of an even NP (cyclic distance NP/2): one pair per node in the lower half of its pencil, DIM * NN/2
        // tasks per element spread over all threads; the direction is a run-time value (branch-free in ec_flux_d).
        if (HALF && !SHUF) {
#pragma unroll
            for (int r = 0; r < GEO::HROUNDS; r++) {
                const int p = r * NN + j;
                if (GEO::HTASKS % NN == 0 || p < GEO::HTASKS) {
                    // NN/2 tasks per direction: round r covers direction 2r (threads j < NN/2) and 2r+1 (the others), so
                    // both candidate decodes have compile-time extents
                    const bool hi = j >= NN / 2;
                    const int q = hi ? j - NN / 2 : j;
                    const int d = 2 * r + (hi ? 1 : 0);
                    constexpr int DMAX = DIM - 1;
                    const int n = hi ? half_class_node<DIM, NP>((2 * r + 1 < DIM) ? 2 * r + 1 : DMAX, q)
                                     : half_class_node<DIM, NP>((2 * r < DIM) ? 2 * r : DMAX, q);
                    const int st = stride_of(NP, d);
                    const Prim a = load_prim_ec<NP>(sP, le * NN + n);
                    const Prim b = load_prim_ec<NP>(sP, le * NN + n + (NP / 2) * st);
                    double F[5], ibl;
                    ec_flux_d(d, a, b, hig, F, ibl);
                    double* const o = sPair + ((le * DIM + d) * NCL + NFULL) * NN + n;
#pragma unroll
                    for (int c = 0; c < 5; c++) o[c * PLANE] = F[c];
                }
            }
        }
        // (b) the element's face nodes: its own side of every face (gather form, no atomics)
        cp_async_wait_all();   // this thread's prefetched records (each task reads only what its own thread fetched)
        for (int ft = j; ft < NFACE * NF; ft += NN) {
            const int f = ft / NF, t = ft - f * NF;
            const int d = f >> 1, side = f & 1;
            double* const rec = sFace + (le * NFACE + f) * NF + t;
            const int v = active ? P.nbr[(size_t)e * NFACE + f] : (int)e0;
            if (v < 0) continue;   // domain boundary: the rate contribution prepared by boundary_kernel is already in rec
            const Prim a = load_prim<NP>(sP, le * NN + node_of_face_node<DIM, NP>(d, side, t));
            const int nn = node_of_face_node<DIM, NP>(d, 1 - side, t);
            Prim b;
            if (v >= e0 && v < e_hi) {   // neighbour inside the patch: its primitives are in shared memory
                b = load_prim<NP>(sP, (int)(v - e0) * NN + nn);
            } else {   // owned element outside the patch or ghost trace: conserved values prefetched into rec
                double qn[5];
#pragma unroll
                for (int c = 0; c < 5; c++) qn[c] = rec[c * NSLOT];
                b = make_prim(qn[0], qn[1], qn[2], qn[3], qn[4], gamma);
            }
            const double sgn = side ? 1.0 : -1.0;
            const double cf = P.inv_hw[d];
            double Fe[5], Dv[5], Fm[5], ibl;
            ec_flux_d(d, a, b, hig, Fe, ibl);
            es_dissipation(a, b, ibl, hig, Dv);
            phys_flux_d(d, a, Fm);
            // (f(u_m).n - f*) / (h_d w_0) with f* = sgn F# - D   (fluid_flux_es_dgsem_operator.h:318-333)
#pragma unroll
            for (int c = 0; c < 5; c++) rec[c * NSLOT] = cf * (sgn * (Fm[c] - Fe[c]) + Dv[c]);
        }
        // The differentiation weights and pair-record offsets of this node's rows depend only on its position: fetch them
        // before the barrier so that the flux records can be read back to back after it.  Row l runs over the other
        // NP-1 nodes of the pencil in rotated order, l = (j_d + k) mod NP (no wasted slot, no divergence).
        double dw[DIM][NP - 1];
        int poff[DIM][NP - 1];
#pragma unroll
        for (int d = 0; d < (SHUF ? 0 : DIM); d++) {
            const int jd = idx[d];
            const int st = stride_of(NP, d);
#pragma unroll
            for (int k = 1; k < NP; k++) {
                int l = jd + k;
                if (l >= NP) l -= NP;
                dw[d][k - 1] = sD[jd * NP + l];
                // pair (jd, l) has cyclic distance c = min(k, NP - k); its slot belongs to its first endpoint
                const int c = (k < NP - k) ? k : NP - k;
                const bool own_first = (2 * k < NP) || (2 * k == NP && jd < NP / 2);
                poff[d][k - 1] = ((le * DIM + d) * NCL + (c - 1)) * NN + (own_first ? j : j + (l - jd) * st);
            }
        }
        group_sync<GROUP, NODES>(tid);

        // ---- node phase 2: assemble the rate of this node ----------------------------------------------------
        const double alpha = sAlpha[le];
        const Prim me = load_prim_phys<NP>(sP, tid);
        double r[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
        // split-form volume term: (1-alpha) * sum_d (-2/h_d) sum_l D[j_d][l] F#_d(u_j,u_l)   (split_form_volume_flux.h:68-98)
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int jd = idx[d];
            double acc[5];
            if (SHUF) {
#pragma unroll
                for (int c = 0; c < 5; c++) acc[c] = accS[d][c];
            } else {
                {   // F#(u,u) = f(u); its weight D[j][j] vanishes except at the two end nodes
                    const double djj = sD[jd * NP + jd];
                    double Fp[5];
                    phys_flux_d(d, me, Fp);
#pragma unroll
                    for (int c = 0; c < 5; c++) acc[c] = djj * Fp[c];
                }
#pragma unroll
                for (int k = 0; k < NP - 1; k++) {
#pragma unroll
                    for (int c = 0; c < 5; c++) acc[c] = fma(dw[d][k], sPair[c * PLANE + poff[d][k]], acc[c]);
                }
            }
            const double s = -2.0 * P.inv_h[d];
#pragma unroll
            for (int c = 0; c < 5; c++) r[c] = fma(s, acc[c], r[c]);
        }
        if (alpha > 0.0) {
            const Prim me = load_prim<NP>(sP, tid);   // the dissipation needs the whole record
            const double oma = 1.0 - alpha;
#pragma unroll
            for (int c = 0; c < 5; c++) r[c] *= oma;
            // subcell finite-volume blend (subcell_finite_volume_flux.h:75-158).  Interface flux between subcells
            // i and i+1 = F#(u_i,u_{i+1}) - D(u_i,u_{i+1}); F# is the pair flux already in shared memory (adjacent
            // pairs are ids 0..NP-2), the dissipation is formed here because few elements need it.
#pragma unroll
            for (int d = 0; d < DIM; d++) {
                const int jd = idx[d];
                const int st = stride_of(NP, d);
                // adjacent pairs (a, a+1) have cyclic distance 1: class 0, slot of node a
                const double* const cls0 = sPair + ((le * DIM + d) * NCL) * NN + j;
                double Fp[5], left[5], right[5];
                phys_flux_d(d, me, Fp);
#pragma unroll
                for (int c = 0; c < 5; c++) { left[c] = Fp[c]; right[c] = Fp[c]; }
                if (jd > 0) {
                    const Prim o = load_prim<NP>(sP, tid - st);
                    double Fd[5], Dv[5], ibl;
                    ec_flux_d(d, o, me, hig, Fd, ibl);   // only for 1/beta_ln; the flux itself comes from the table
                    es_dissipation(o, me, ibl, hig, Dv);
#pragma unroll
                    for (int c = 0; c < 5; c++) left[c] = (SHUF ? Fd[c] : cls0[c * PLANE - st]) - Dv[c];
                }
                if (jd < NP - 1) {
                    const Prim o = load_prim<NP>(sP, tid + st);
                    double Fd[5], Dv[5], ibl;
                    ec_flux_d(d, me, o, hig, Fd, ibl);
                    es_dissipation(me, o, ibl, hig, Dv);
#pragma unroll
                    for (int c = 0; c < 5; c++) right[c] = (SHUF ? Fd[c] : cls0[c * PLANE]) - Dv[c];
                }
                const double cf = alpha * P.inv_h[d] / sW[jd];
#pragma unroll
                for (int c = 0; c < 5; c++) r[c] += cf * (left[c] - right[c]);
            }
        }
        // faces
#pragma unroll
        for (int d = 0; d < DIM; d++) {
            const int t = face_node_index<DIM, NP>(d, i0, i1, i2);
            if (idx[d] == 0 || idx[d] == NP - 1) {   // NP >= 2: a node is on at most one face per direction
                const int f = 2 * d + (idx[d] == 0 ? 0 : 1);
                const double* const rec = sFace + (le * NFACE + f) * NF + t;
#pragma unroll
                for (int c = 0; c < 5; c++) r[c] += rec[c * NSLOT];
            }
        }

        // ---- two-fluid sources on this species: (q/m)(rho E + m x B) and (q/m) m.E  (uniform branch) ----------
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

        // ---- inverse mass is folded into the factors above; stage update ------------------------------------
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
                // compute_cell_transport_speed (:450-514) of the updated state
                const double inv = rcp_pos(qn[0]);
                const double sm = qn[1] * qn[1] + qn[2] * qn[2] + qn[3] * qn[3];
                const double pr = gm1 * (qn[4] - sm * (0.5 * inv));
                double conv = fabs(qn[1] * inv) * P.inv_h[0];
                if (DIM > 1) conv = fmax(conv, fabs(qn[2] * inv) * P.inv_h[1]);
                if (DIM > 2) conv = fmax(conv, fabs(qn[3] * inv) * P.inv_h[2]);
                const double c2 = gamma * pr * inv;
                // a non-positive or NaN c^2 (unphysical state) must reach the host as a NaN speed, not be clamped
                const double cs = (c2 > 0.0) ? sqrt_pos(c2) : sqrt(c2 - 1.0);
                const double speed = P.max_eig * cs + conv;
                vmax_local = (speed > vmax_local || speed != speed) ? speed : vmax_local;
            }
        }
    }

    // ---- field components: carried through unchanged by the reference's operator; with the two-fluid
    // sources switched on, E gets -J/eps0 and phi gets chi rho_c/eps0 (J, rho_c summed over the species of this node)
    if (active && nc > 5 * P.nsp && !P.fields_skip) {
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
        for (int k = 0; k < 8; k++) {   // fields_enabled means exactly these 8 components (five_moment.h:123-138)
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

int stage_smem_bytes(int dim, int Np) {
    int bytes = 0;
#define CALL(D_, N_) { bytes = GeoC<D_, N_>::SMEM_DOUBLES * (int)sizeof(double); }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
    return bytes;
}

int prepare_kernels(int dim, int Np) {
    cudaError_t err = cudaSuccess;
#define CALL(D_, N_)                                                                                         \
    {                                                                                                        \
        err = cudaFuncSetAttribute(stage_kernel<D_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                   GeoC<D_, N_>::SMEM_DOUBLES * (int)sizeof(double));                         \
        if (err == cudaSuccess)                                                                              \
            err = cudaFuncSetAttribute(stage_kernel<D_, N_>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                       cudaSharedmemCarveoutMaxShared);                                      \
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
        using GEO = GeoC<D_, N_>;                                                                            \
        const int64_t blocks = (n + GEO::G - 1) / GEO::G;                                                    \
        stage_kernel<D_, N_><<<(unsigned)blocks, GEO::THREADS, GEO::SMEM_DOUBLES * sizeof(double), s>>>(P);  \
    }
    WGPU_DISPATCH(dim, Np, CALL);
#undef CALL
}

}  // namespace wgpu
