// The fused ES-DGSEM stage, "pencil" decomposition (round 2), per-thread phase functions.
//
// What is computed is the same forward-Euler stage as stage_kernel (dgsem_stage_kernel.cu; reference
// fluid_flux_es_dgsem_operator.h:127-214 with split_form_volume_flux.h:61-99, subcell_finite_volume_flux.h:68-159,
// persson_peraire_shock_indicator.h:44-123, :301-342 faces, :216-240 inverse mass, :450-514 transport speed).  How:
//
//   * one THREAD owns one PENCIL (the Np nodes of an element along one direction) instead of one node.  All Np(Np-1)/2
//     unordered node pairs of the pencil are evaluated from registers (the Chandrashekar flux is symmetric, so each pair
//     feeds both of its nodes), the two face nodes at the ends of the pencil are added by the same thread: no pair-flux
//     table, no face records, no task lists, every direction is a compile-time constant (no selects), index arithmetic is
//     paid once per Np nodes, and a thread has Np(Np-1)/2 + 2 independent flux evaluations to interleave;
//   * a block works on a patch of E elements (2x2x2 for 3-D degree 3) in DIM+1 phases separated by block barriers:
//       P0 (x-pencil owner)  load u, node primitives once -> shared records; first indicator transform
//       Py (y-pencil owner)  pair + end fluxes along y -> shared accumulator (first writer); second indicator transform
//       Pz (z-pencil owner)  pair + end fluxes along z -> accumulator += ; third transform, modal energies
//       Px (x-pencil owner)  pair + end fluxes along x, + accumulator, blend, sources, stage update, store, CFL
//     The last phase owns Np consecutive nodes per thread: 16-byte vector loads/stores of u and dst;
//   * a pencil end whose neighbour element lies in the patch reads the neighbour's record from shared memory; otherwise
//     the neighbour's 5 conserved values are fetched (other patch / NCCL ghost trace) and turned into a record by the
//     thread itself (in a full 2x2x2 patch every thread has exactly one such end per direction: no divergence);
//   * the diagonal term D_jj f(u_j) of the volume sum, the f(u_m).n part of the face term and the pencil-end terms
//     alpha f(u_0)/(w_0 h) of the subcell scheme cancel analytically for every alpha ((1-alpha) + alpha - 1 = 0 with
//     D_00 = -1/(2 w_0), the SBP property): none of them is evaluated.  The reference forms them separately and leaves
//     their round-off (one ulp of f/(h w_0), the size of every other summand's rounding);
//   * DRAM latency is taken off the threads' paths with L2 prefetch instructions (no registers held): what lies beyond the
//     pencil ends is LOADED one phase ahead; E and B of the Lorentz force, the old destination, the next species' state and
//     the state of the patch that the next block of this SM will start with are PREFETCHED a phase or a block ahead
//     (WGPU_PENCIL_PREFETCH, StageParams::lookahead);
//   * elements whose blending factor is positive (rare: the troubled cells) take a per-node correction
//     -alpha vol + alpha fv recomputed from the shared records; everybody else never touches the subcell scheme.
//
// Shared memory: records as 5 planes of double2 [(rho,u0) (u1,u2) (beta,ln rho) (ln beta,|u|^2) (|u|+c, 1/beta)],
// accumulator as 3 planes of double2 [(r0,r1) (r2,r3) (r4, indicator scratch)], node index XOR-swizzled for Np = 4 so
// that the 16-byte accesses of x-, y- and z-pencil owners are all bank-conflict free.
//
// The functions are __device__ code; wgpu_portable.cuh lets tests/emu/ compile the same source for the host and run it
// thread by thread, phase by phase (test infrastructure; the product has no CPU path).
#pragma once
#include "dgsem_kernels.cuh"
#include "dgsem_physics.cuh"

#ifndef WGPU_PENCIL_THREADS
#define WGPU_PENCIL_THREADS 128
#endif
#ifndef WGPU_PENCIL_TMA
#define WGPU_PENCIL_TMA 0        // 1: P0 stages the patch's state block in shared memory with cp.async.bulk + mbarrier (A/B only)
#endif
#ifndef WGPU_PENCIL_PREFETCH
#define WGPU_PENCIL_PREFETCH 6   // L2 prefetches ahead of use (0: off; 1: u again to L1 + old dst before the x pair fluxes; 2: old dst; 3: + E, B of the Lorentz force; 4: + the next species' state; 5: + the state of the patch the next block of this SM will work on; 6: + phi, psi and the old field values for the fused field phases)
#endif
#ifndef WGPU_PENCIL_P0_ROLL
#define WGPU_PENCIL_P0_ROLL 0    // 1: P0 forms the node records two at a time in a rolled loop (half the code, ILP 2 instead of NP)
#endif
#ifndef WGPU_PENCIL_RTDIR
#define WGPU_PENCIL_RTDIR 0      // 1: one flux-phase body for all directions (run-time direction, pencil_phase_flux_rt); measured slower
#endif

namespace wgpu {

__host__ __device__ constexpr int pencil_elems(int dim, int np) {
    int e = 1;
    while (2 * e * ipow_c(np, dim - 1) <= WGPU_PENCIL_THREADS) e *= 2;
    return e;
}

// compile-time loop: f(IC<B>{}), ..., f(IC<E-1>{}) (the pair loops must be unrolled whatever the unroller's budget says:
// a rolled iteration would index the register arrays dynamically, i.e. put them in local memory)
template <int I>
struct IC { static constexpr int value = I; };
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(IC<B>{});
        static_for<B + 1, E>(f);
    }
}

template <int DIM, int NP>
struct PGeo {
    static constexpr int NN = ipow_c(NP, DIM);
    static constexpr int NPEN = ipow_c(NP, DIM - 1);      // pencils per element and direction = nodes per face
    static constexpr int E = pencil_elems(DIM, NP);       // elements per block
    static constexpr int USED = E * NPEN;                 // threads that own a pencil
    static constexpr int THREADS = ((USED + 31) / 32) * 32;
    static constexpr int NODES = E * NN;
    static constexpr int NFACE = 2 * DIM;
    // dynamic shared memory, in doubles
    static constexpr int OFF_REC = 0;                     // double2 [5][NODES]
    static constexpr int OFF_BUF = 10 * NODES;            // double2 [3][NODES]
    static constexpr int OFF_EN = 16 * NODES;             // double2 [THREADS] modal-energy partials
    static constexpr int OFF_NBR = OFF_EN + 2 * THREADS;  // int [E][NFACE]
    static constexpr int OFF_RED = OFF_NBR + (E * NFACE + 1) / 2 + ((E * NFACE + 1) / 2) % 2;
    static constexpr int SMEM_DOUBLES = OFF_RED + 32;
};

// node index -> slot in a shared plane.  Np = 4: slot = n ^ b1 ^ 6 c0 with n = i0 + 4 i1 + 16 (i2 or element) ..., which
// makes the eight 16-byte accesses of a quarter-warp distinct modulo 8 for x-owners (n = 4 pe + m), y-owners
// (n = i0 + 4 m + 16 i2) and z-owners (n = pe + 16 m) alike (tests/test_pencil_emu_cpu.py::test_swizzled_planes_are_bijective_and_bank_conflict_free enumerates them).
template <int NP>
__device__ __forceinline__ constexpr int pslot(const int n) {
    return (NP == 4) ? (n ^ ((n >> 3) & 1) ^ (((n >> 4) & 1) * 6)) : n;
}

// element-local node m of pencil pe along D
template <int DIM, int NP, int D>
__device__ __forceinline__ constexpr int pencil_node(const int pe, const int m) {
    if (DIM == 2) return D == 0 ? pe * NP + m : pe + NP * m;
    return D == 0 ? pe * NP + m : (D == 1 ? (pe % NP) + NP * m + NP * NP * (pe / NP) : pe + NP * NP * m);
}

__device__ __forceinline__ void put_rec(double2* sRec, const int nodes, const int slot, const Prim& r) {
    sRec[slot] = make_double2(r.rho, r.u0);
    sRec[nodes + slot] = make_double2(r.u1, r.u2);
    sRec[2 * nodes + slot] = make_double2(r.beta, r.lrho);
    sRec[3 * nodes + slot] = make_double2(r.lbeta, r.q2);
    sRec[4 * nodes + slot] = make_double2(r.lam, r.ib);
}
// FULL: also the two fields only the dissipation needs
template <bool FULL>
__device__ __forceinline__ Prim get_rec(const double2* sRec, const int nodes, const int slot) {
    const double2 a = sRec[slot], b = sRec[nodes + slot], c = sRec[2 * nodes + slot], d = sRec[3 * nodes + slot];
    Prim o;
    o.rho = a.x; o.u0 = a.y; o.u1 = b.x; o.u2 = b.y; o.beta = c.x; o.lrho = c.y; o.lbeta = d.x; o.q2 = d.y;
    o.p = 0.0; o.H = 0.0; o.lam = 0.0; o.ib = 0.0;
    if (FULL) {
        const double2 e = sRec[4 * nodes + slot];
        o.lam = e.x; o.ib = e.y;
    }
    return o;
}

// NP consecutive doubles owned by one thread: one 32-byte access for Np = 4 (LDG/STG.256, sm_100: the thread's data is
// exactly one sector, so a warp instruction moves 1 KB of whole sectors), 16-byte accesses for other even Np
template <int NP>
__device__ __forceinline__ void load_run(const double* p, double (&v)[NP]) {
#if !WGPU_HOST_EMU
    if constexpr (NP == 4) {
        asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
        return;
    } else if constexpr (NP % 2 == 0) {
#pragma unroll
        for (int m = 0; m + 1 < NP; m += 2) {
            const double2 t = *reinterpret_cast<const double2*>(p + m);
            v[m] = t.x;
            v[m + 1] = t.y;
        }
        return;
    }
#endif
#pragma unroll
    for (int m = 0; m < NP; m++) v[m] = p[m];
}
template <int NP>
__device__ __forceinline__ void store_run(double* p, const double (&v)[NP]) {
#if !WGPU_HOST_EMU
    if constexpr (NP == 4) {
        asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
        return;
    } else if constexpr (NP % 2 == 0) {
#pragma unroll
        for (int m = 0; m + 1 < NP; m += 2) *reinterpret_cast<double2*>(p + m) = make_double2(v[m], v[m + 1]);
        return;
    }
#endif
#pragma unroll
    for (int m = 0; m < NP; m++) p[m] = v[m];
}

// What lies beyond the two ends of a pencil, fetched one phase ahead (the loads fly while the previous phase finishes and
// the block barrier is crossed) and consumed first thing in the phase that owns the pencil
struct PencilHalo {
    double q[2][5];   // kind 1: conserved values of the neighbour's face node; kind 2: boundary_kernel's contribution
    int kind[2];      // 0: record in shared memory, 1: conserved values (other patch / ghost trace), 2: domain boundary
    int nslot[2];     // kind 0: slot of the neighbour's end node
};

template <int DIM, int NP, int D>
__device__ __forceinline__ void pencil_halo_fetch(const StageParams& P, const int v0, const int v1, const int pe, const int64_t e0,
                                                  const int64_t e_hi, const int sp, PencilHalo& h) {
    using G = PGeo<DIM, NP>;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int v = side ? v1 : v0;
        h.nslot[side] = 0;
        if (v >= e0 && v < e_hi) {
            h.kind[side] = 0;
            h.nslot[side] = pslot<NP>((int)(v - e0) * G::NN + pencil_node<DIM, NP, D>(pe, side ? 0 : NP - 1));
        } else {
            const double* src;
            size_t stride;
            if (v < 0) {
                h.kind[side] = 2;
                src = P.bres + (((size_t)(-1 - v) * P.nsp + sp) * 5) * G::NPEN + pe;
                stride = G::NPEN;
            } else if (v < P.n_elems) {
                h.kind[side] = 1;
                src = P.u + ((size_t)v * P.nc + 5 * sp) * G::NN + pencil_node<DIM, NP, D>(pe, side ? 0 : NP - 1);
                stride = G::NN;
            } else {
                h.kind[side] = 1;
                src = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + 5 * sp) * G::NPEN + pe;
                stride = G::NPEN;
            }
#pragma unroll
            for (int c = 0; c < 5; c++) h.q[side][c] = src[(size_t)c * stride];
        }
    }
}

#if WGPU_PENCIL_TMA && !WGPU_HOST_EMU
// The north_star sketch ("stages each element's block in shared memory via TMA"), as an A/B variant: one thread arms an
// mbarrier with the byte count and issues one 1-D bulk copy (cp.async.bulk, the TMA engine) per element of the patch -- the
// species' 5 x Np^dim doubles of an element are contiguous -- into the (still unused) record area; every thread waits on the
// barrier's phase, reads its Np x 5 values back from shared memory, and a block barrier frees the area for the records.
// Measured against the default (each thread loads its own 32 bytes per component straight into registers): profiles/README.md.
// bulk copies move multiples of 16 bytes between 16-byte-aligned addresses: the element block of a species qualifies for Np = 4
template <int DIM, int NP>
__device__ __forceinline__ constexpr bool pencil_tma_ok() { return (PGeo<DIM, NP>::NN * sizeof(double)) % 16 == 0; }
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@!p bra WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
#endif

// make_prim for the rare second outside end of a pencil (partial patches, patches of one element): out of line, so that the
// hot code of every flux phase carries one inlined copy of the ~200-instruction routine instead of two
__device__ __noinline__ void make_prim_cold(const double q0, const double q1, const double q2, const double q3, const double q4,
                                            const double gamma, Prim* out) {
    *out = make_prim(q0, q1, q2, q3, q4, gamma);
}

// ---------------------------------------------------------------------------------------------------------------------
// P0: the thread that owns x-pencil (le, pe): node primitives, first indicator transform, neighbour table, and the
// fetch of what lies beyond the ends of the y-pencil this thread owns in the next phase
// ---------------------------------------------------------------------------------------------------------------------
template <int DIM, int NP>
__device__ __forceinline__ void pencil_phase0(const StageParams& P, double* smem, const int tid, const int64_t e0, const int sp,
                                              PencilHalo& next) {
    using G = PGeo<DIM, NP>;
    double2* const sRec = reinterpret_cast<double2*>(smem + G::OFF_REC);
    double2* const sBuf = reinterpret_cast<double2*>(smem + G::OFF_BUF);
    int* const sNbr = reinterpret_cast<int*>(smem + G::OFF_NBR);
#if WGPU_PENCIL_TMA && !WGPU_HOST_EMU
    // staged through shared memory by the bulk-copy engine (see above); the staging area is the record area, so every
    // thread of the block passes the barrier below before the first record is written
    double tma_q[5][NP];
    if (pencil_tma_ok<DIM, NP>()) {
        const int le_ = tid / G::NPEN, pe_ = tid - le_ * G::NPEN;
        mbar_wait(smem + G::OFF_RED + 30, (unsigned)(sp & 1));
        if (tid < G::USED && e0 + le_ < P.elem_end) {
            const double* mine = smem + G::OFF_REC + (size_t)le_ * 5 * G::NN + pe_ * NP;
#pragma unroll
            for (int c = 0; c < 5; c++)
#pragma unroll
                for (int m = 0; m < NP; m++) tma_q[c][m] = mine[c * G::NN + m];
        }
        __syncthreads();
    }
#endif
    if (tid >= G::USED) return;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    const int64_t e_hi = (e0 + G::E < P.elem_end) ? e0 + G::E : P.elem_end;
    if (e >= P.elem_end) return;   // (its records are never read: the in-patch test uses e_hi)
#if WGPU_PENCIL_PREFETCH >= 5 && !WGPU_HOST_EMU
    // the block that follows this one on its SM starts by loading the state of the patch P.lookahead elements ahead: ask for
    // those lines now (L2), a whole block lifetime early
    if (sp == 0 && P.lookahead > 0 && e + P.lookahead < P.elem_end) {
        const double* const nx = P.u + ((size_t)(e + P.lookahead) * P.nc) * G::NN + pe * NP;
#pragma unroll
        for (int c = 0; c < 5; c++) asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)c * G::NN));
    }
#endif
    // neighbour ids of the y-faces first: the loads they address (ends of next phase's pencil) then overlap with the state's
    const int v0 = P.nbr[(size_t)e * G::NFACE + 2], v1 = P.nbr[(size_t)e * G::NFACE + 3];
    const double* src = P.u + ((size_t)e * P.nc + 5 * sp) * G::NN + pe * NP;
    double v[NP];
#if WGPU_PENCIL_P0_ROLL
    pencil_halo_fetch<DIM, NP, 1>(P, v0, v1, pe, e0, e_hi, sp, next);
    if (sp == 0)
        for (int f = pe; f < G::NFACE; f += G::NPEN) sNbr[le * G::NFACE + f] = P.nbr[(size_t)e * G::NFACE + f];
    if (NP % 2 == 0) {
        // two nodes per iteration of a REAL loop: half the code of four inlined make_prim, two independent chains in flight
        double va[NP / 2 > 0 ? NP / 2 : 1], vb[NP / 2 > 0 ? NP / 2 : 1];
#pragma unroll 1
        for (int it = 0; it < NP / 2; it++) {
            double qa[5], qb[5];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                const double2 t = *reinterpret_cast<const double2*>(src + (size_t)c * G::NN + 2 * it);
                qa[c] = t.x;
                qb[c] = t.y;
            }
            const Prim ra = make_prim(qa[0], qa[1], qa[2], qa[3], qa[4], P.gamma);
            const Prim rb = make_prim(qb[0], qb[1], qb[2], qb[3], qb[4], P.gamma);
            put_rec(sRec, G::NODES, pslot<NP>(le * G::NN + pe * NP + 2 * it), ra);
            put_rec(sRec, G::NODES, pslot<NP>(le * G::NN + pe * NP + 2 * it + 1), rb);
            // (the indicator variable goes through the scratch slot of its own node; transformed below)
            sBuf[2 * G::NODES + pslot<NP>(le * G::NN + pe * NP + 2 * it)].y = ra.p * ra.rho;
            sBuf[2 * G::NODES + pslot<NP>(le * G::NN + pe * NP + 2 * it + 1)].y = rb.p * rb.rho;
        }
#pragma unroll
        for (int m = 0; m < NP; m++) v[m] = sBuf[2 * G::NODES + pslot<NP>(le * G::NN + pe * NP + m)].y;
    } else
#endif
    {
    double q[5][NP];
#if WGPU_PENCIL_TMA && !WGPU_HOST_EMU
    if (pencil_tma_ok<DIM, NP>()) {
#pragma unroll
        for (int c = 0; c < 5; c++)
#pragma unroll
            for (int m = 0; m < NP; m++) q[c][m] = tma_q[c][m];
    } else
#endif
    {
#pragma unroll
        for (int c = 0; c < 5; c++) load_run<NP>(src + (size_t)c * G::NN, q[c]);
    }
#if !WGPU_PENCIL_P0_ROLL
    pencil_halo_fetch<DIM, NP, 1>(P, v0, v1, pe, e0, e_hi, sp, next);
    if (sp == 0)
        for (int f = pe; f < G::NFACE; f += G::NPEN) sNbr[le * G::NFACE + f] = P.nbr[(size_t)e * G::NFACE + f];
#endif
#pragma unroll
    for (int m = 0; m < NP; m++) {
        const Prim r = make_prim(q[0][m], q[1][m], q[2][m], q[3][m], q[4][m], P.gamma);
        put_rec(sRec, G::NODES, pslot<NP>(le * G::NN + pe * NP + m), r);
        v[m] = r.p * r.rho;   // indicator variable, fluid_flux_es_dgsem_operator.h:286-290
    }
    }
    // Legendre analysis along x (persson_peraire_shock_indicator.h:56 via FESeries::Legendre, sum-factorised)
#pragma unroll
    for (int k = 0; k < NP; k++) {
        double ck = 0.0;
#pragma unroll
        for (int m = 0; m < NP; m++) ck = fma(P.T.V[k * NP + m], v[m], ck);
        sBuf[2 * G::NODES + pslot<NP>(le * G::NN + pe * NP + k)].y = ck;
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Flux sums of one pencil along D: acc[m][c] = sum_{l != m} (-2 D[m][l] / h_D) F#_D(u_m, u_l)  +  the face terms at the
// two ends (split_form_volume_flux.h:68-98 without the cancelling diagonal; fluid_flux_es_dgsem_operator.h:318-333
// without the cancelling f(u_m).n).  The ends come first (their data was fetched a phase ahead); the pairs are ordered
// (0,1) (0,2) ... (0,Np-1) (1,2) ... so that node m is complete - done(m, acc[m]) - and its record dead as early as possible.
// ---------------------------------------------------------------------------------------------------------------------
template <int DIM, int NP, int D, class Done>
__device__ __forceinline__ void pencil_sums(const StageParams& P, const double2* sRec, const PencilHalo& h, const int le, const int pe,
                                            Done&& done) {
    using G = PGeo<DIM, NP>;
    const double hig = P.hig, gamma = P.gamma;
    Prim r[NP];
#pragma unroll
    for (int m = 0; m < NP; m++) {
        const int slot = pslot<NP>(le * G::NN + pencil_node<DIM, NP, D>(pe, m));
        r[m] = (m == 0 || m == NP - 1) ? get_rec<true>(sRec, G::NODES, slot) : get_rec<false>(sRec, G::NODES, slot);
    }
    double acc[NP][5];
#pragma unroll
    for (int m = 0; m < NP; m++)
#pragma unroll
        for (int c = 0; c < 5; c++) acc[m][c] = 0.0;
    // ---- the two ends: own side of the face, gather form (f*(a,b,n) = -f*(b,a,-n) bit for bit) -----------------------
    // In a full patch every thread has exactly ONE end whose neighbour lies outside the patch, but WHICH one differs between
    // the lanes of a warp (the left elements of the patch look left, the right ones right).  So the outside end's record is
    // formed once, side-agnostic (no divergence: one make_prim per warp instead of two half-empty ones), and each side then
    // picks its neighbour record with selects; the in-patch read always happens (slot 0 when unused) to stay branch-free.
    const double cf = P.inv_hw[D];
    const int os = (h.kind[0] == 1) ? 0 : 1;
    Prim bo;
    {
        double qo[5];
#pragma unroll
        for (int c = 0; c < 5; c++) qo[c] = os ? h.q[1][c] : h.q[0][c];
        if (h.kind[0] == 1 || h.kind[1] == 1) bo = make_prim(qo[0], qo[1], qo[2], qo[3], qo[4], gamma);
        else bo = r[0];
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int m = side ? NP - 1 : 0;
        Prim b = get_rec<true>(sRec, G::NODES, h.nslot[side]);
        if (h.kind[side] == 1) {
            if (side == os) b = bo;
            else {   // both ends outside
                Prim cold;
                make_prim_cold(h.q[side][0], h.q[side][1], h.q[side][2], h.q[side][3], h.q[side][4], gamma, &cold);
                b = cold;
            }
        }
        // (a domain-boundary end evaluates the flux against a harmless record and discards it: the common path stays one
        // basic block with the pair loop below, which is what lets the scheduler interleave the two ends and the pairs)
        double Fe[5], Dv[5], ibl;
        ec_flux_d(D, r[m], b, hig, Fe, ibl);
        es_dissipation(r[m], b, ibl, hig, Dv);
        // (f(u_m).n - f*) / (h w_0) with f* = sgn F# - D and the f(u_m).n part dropped (see the file header)
        double ct[5];
#pragma unroll
        for (int c = 0; c < 5; c++) ct[c] = cf * (side ? (Dv[c] - Fe[c]) : (Dv[c] + Fe[c]));
        if (h.kind[side] == 2) {
            // domain boundary: boundary_kernel's contribution (f(u_m).n - f*_LF, Gauss(p+2) quadrature) plus the diagonal
            // volume term, which only cancels against a NODAL f(u_m).n
            Prim a = r[m];
            a.p = 0.5 * a.rho * a.ib;
            a.H = a.p * (2.0 * hig) + 0.5 * a.rho * a.q2 + a.p;
            double Fp[5];
            phys_flux_d(D, a, Fp);
            const double sg = side ? -cf : cf;
#pragma unroll
            for (int c = 0; c < 5; c++) ct[c] = h.q[side][c] + sg * Fp[c];
        }
#pragma unroll
        for (int c = 0; c < 5; c++) acc[m][c] += ct[c];
    }
    const double s = -2.0 * P.inv_h[D];
    static_for<0, NP>([&](auto J) {
        constexpr int j = decltype(J)::value;
        static_for<j + 1, NP>([&](auto L) {
            constexpr int l = decltype(L)::value;
            double F[5], ibl;
            ec_flux_d(D, r[j], r[l], hig, F, ibl);
            const double wj = s * P.T.D[j * NP + l], wl = s * P.T.D[l * NP + j];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                acc[j][c] = fma(wj, F[c], acc[j][c]);
                acc[l][c] = fma(wl, F[c], acc[l][c]);
            }
        });
        done(j, acc[j]);
    });
}

// Legendre analysis of the indicator scratch along one direction, in place in registers
template <int NP>
__device__ __forceinline__ void legendre_1d(const StageParams& P, double (&v)[NP]) {
    double o[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
        double ck = 0.0;
#pragma unroll
        for (int m = 0; m < NP; m++) ck = fma(P.T.V[k * NP + m], v[m], ck);
        o[k] = ck;
    }
#pragma unroll
    for (int k = 0; k < NP; k++) v[k] = o[k];
}

// modal energies of this pencil's coefficients (last transform direction LD): shell = some mode index equals NP-1
template <int DIM, int NP, int LD>
__device__ __forceinline__ double2 pencil_energies(const int pe, const double (&c)[NP]) {
    bool shell_pencil;   // a tangential index of this pencil is already NP-1
    if (DIM == 2) shell_pencil = (pe == NP - 1);
    else shell_pencil = (pe % NP == NP - 1) || (pe / NP == NP - 1);
    double g0 = 0.0, g1 = 0.0;
#pragma unroll
    for (int k = 0; k < NP; k++) {
        const double sq = c[k] * c[k];
        if (shell_pencil || k == NP - 1) g1 += sq; else g0 += sq;
    }
    return make_double2(g0, g1);
}

// ---------------------------------------------------------------------------------------------------------------------
// Py / Pz: pencil owner along D (D >= 1).  FIRST: the accumulator is written, not added to.  LAST_IND: this is the last
// indicator transform (the modal energies of the pencil go to sEn).  `h` holds this pencil's ends on entry and the ends of
// the pencil the thread owns in the NEXT phase (direction D+1, or x for the final phase) on return.
// ---------------------------------------------------------------------------------------------------------------------
template <int DIM, int NP, int D>
__device__ __forceinline__ void pencil_phase_mid(const StageParams& P, double* smem, const int tid, const int64_t e0, const int sp,
                                                 PencilHalo& h) {
    using G = PGeo<DIM, NP>;
    constexpr bool FIRST = (D == 1), LAST_IND = (D == DIM - 1);
    constexpr int NEXT = (D + 1 < DIM) ? D + 1 : 0;
    const double2* const sRec = reinterpret_cast<const double2*>(smem + G::OFF_REC);
    double2* const sBuf = reinterpret_cast<double2*>(smem + G::OFF_BUF);
    double2* const sEn = reinterpret_cast<double2*>(smem + G::OFF_EN);
    const int* const sNbr = reinterpret_cast<const int*>(smem + G::OFF_NBR);
    if (tid >= G::USED) return;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    const int64_t e_hi = (e0 + G::E < P.elem_end) ? e0 + G::E : P.elem_end;
    if (e >= P.elem_end) {
        if (LAST_IND) sEn[tid] = make_double2(0.0, 0.0);
        return;
    }
    // indicator scratch of the pencil's nodes: transform along D now (cheap), written back with the sums
    double ind[NP];
#pragma unroll
    for (int m = 0; m < NP; m++) ind[m] = sBuf[2 * G::NODES + pslot<NP>(le * G::NN + pencil_node<DIM, NP, D>(pe, m))].y;
    legendre_1d<NP>(P, ind);
    if (LAST_IND) sEn[tid] = pencil_energies<DIM, NP, D>(pe, ind);

    pencil_sums<DIM, NP, D>(P, sRec, h, le, pe, [&](const int m, const double (&a)[5]) {
        const int slot = pslot<NP>(le * G::NN + pencil_node<DIM, NP, D>(pe, m));
        if (FIRST) {
            sBuf[slot] = make_double2(a[0], a[1]);
            sBuf[G::NODES + slot] = make_double2(a[2], a[3]);
            sBuf[2 * G::NODES + slot] = make_double2(a[4], ind[m]);
        } else {
            const double2 x = sBuf[slot], y = sBuf[G::NODES + slot], z = sBuf[2 * G::NODES + slot];
            sBuf[slot] = make_double2(a[0] + x.x, a[1] + x.y);
            sBuf[G::NODES + slot] = make_double2(a[2] + y.x, a[3] + y.y);
            sBuf[2 * G::NODES + slot] = make_double2(a[4] + z.x, ind[m]);
        }
    });
    pencil_halo_fetch<DIM, NP, NEXT>(P, sNbr[le * G::NFACE + 2 * NEXT], sNbr[le * G::NFACE + 2 * NEXT + 1], pe, e0, e_hi, sp, h);
}

// Troubled elements only (alpha > 0): -alpha * (volume sums) + alpha * (interior subcell-interface fluxes) of node j,
// recomputed from the shared records (subcell_finite_volume_flux.h:75-158; the pencil-end terms cancel, see header)
template <int DIM, int NP>
__device__ __forceinline__ void pencil_fv_correction(const StageParams& P, const double2* sRec, const int le, const int j, const double alpha,
                                                     double (&corr)[5]) {
    using G = PGeo<DIM, NP>;
    const double hig = P.hig;
    const Prim me = get_rec<true>(sRec, G::NODES, pslot<NP>(le * G::NN + j));
#pragma unroll
    for (int c = 0; c < 5; c++) corr[c] = 0.0;
    int rem = j;
#pragma unroll 1
    for (int d = 0; d < DIM; d++) {
        const int jd = rem % NP;
        rem /= NP;
        const int st = (d == 0) ? 1 : (d == 1 ? NP : NP * NP);
        const double s = -2.0 * P.inv_h[d];
        double vol[5] = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 1
        for (int l = 0; l < NP; l++) {
            if (l == jd) continue;
            const Prim o = get_rec<true>(sRec, G::NODES, pslot<NP>(le * G::NN + j + (l - jd) * st));
            double F[5], ibl;
            ec_flux_d(d, me, o, hig, F, ibl);
            const double w = s * P.T.D[jd * NP + l];
#pragma unroll
            for (int c = 0; c < 5; c++) vol[c] = fma(w, F[c], vol[c]);
            if (l == jd - 1 || l == jd + 1) {
                // interface between subcells l and jd: F# - D(left, right); it enters node jd with +/- alpha / (h w_jd)
                double Dv[5];
                if (l < jd) es_dissipation(o, me, ibl, hig, Dv); else es_dissipation(me, o, ibl, hig, Dv);
                const double cfv = ((l < jd) ? alpha : -alpha) * P.inv_h[d] / P.T.w[jd];
#pragma unroll
                for (int c = 0; c < 5; c++) corr[c] = fma(cfv, F[c] - Dv[c], corr[c]);
            }
        }
#pragma unroll
        for (int c = 0; c < 5; c++) corr[c] = fma(-alpha, vol[c], corr[c]);
    }
}

// Final-phase epilogue of a species: sources, stage update, store, transport speed of the values written.  rate[c][m] holds
// the complete rate of node m of the x-pencil (e, pe) on entry and the new state on return.
template <int DIM, int NP>
__device__ __forceinline__ double pencil_finish(const StageParams& P, const int64_t e, const int pe, const int sp, const double dt,
                                                double (&rate)[5][NP]) {
    using G = PGeo<DIM, NP>;
    const size_t off = ((size_t)e * P.nc + 5 * sp) * G::NN + pe * NP;   // of node 0 of the pencil, component 0
    // two-fluid sources on this species: (q/m)(rho E + m x B) and (q/m) m.E.  (u is read again here and below: it was
    // read by this very thread in P0, an L1/L2 hit, not HBM traffic.)
    if (P.src_on) {
        const double* const fp = P.u + ((size_t)e * P.nc + 5 * P.nsp) * G::NN + pe * NP;
        const double qm = P.qm[sp];
        double F[6][NP], m0[NP], m1[NP], m2[NP], m3[NP];
#pragma unroll
        for (int k = 0; k < 6; k++) load_run<NP>(fp + (size_t)k * G::NN, F[k]);
        load_run<NP>(P.u + off, m0);
        load_run<NP>(P.u + off + (size_t)G::NN, m1);
        load_run<NP>(P.u + off + 2 * (size_t)G::NN, m2);
        load_run<NP>(P.u + off + 3 * (size_t)G::NN, m3);
#pragma unroll
        for (int m = 0; m < NP; m++) {
            rate[1][m] += qm * (m0[m] * F[0][m] + (m2[m] * F[5][m] - m3[m] * F[4][m]));
            rate[2][m] += qm * (m0[m] * F[1][m] + (m3[m] * F[3][m] - m1[m] * F[5][m]));
            rate[3][m] += qm * (m0[m] * F[2][m] + (m1[m] * F[4][m] - m2[m] * F[3][m]));
            rate[4][m] += qm * (m1[m] * F[0][m] + m2[m] * F[1][m] + m3[m] * F[2][m]);
        }
    }

    // stage update (inverse mass is folded into the factors above), :190-212 of the reference operator; component by
    // component with the next component's runs of u / old values already in flight.  rate[][] becomes the new state.
    const bool need_old = (P.mode == 2) || (P.mode == 0 && P.beta != 0.0);
    const double* const oldp = (P.mode == 2) ? P.sol_in : P.dst;
    double qn[NP] = {}, on[NP] = {};
    if (P.mode != 1) load_run<NP>(P.u + off, qn);
    if (need_old) load_run<NP>(oldp + off, on);
#pragma unroll
    for (int c = 0; c < 5; c++) {
        double q[NP], old[NP], out2[NP];
#pragma unroll
        for (int m = 0; m < NP; m++) { q[m] = qn[m]; old[m] = on[m]; }
        if (c < 4) {
            if (P.mode != 1) load_run<NP>(P.u + off + (size_t)(c + 1) * G::NN, qn);
            if (need_old) load_run<NP>(oldp + off + (size_t)(c + 1) * G::NN, on);
        }
#pragma unroll
        for (int m = 0; m < NP; m++) {
            const double r = rate[c][m];
            double v;
            if (P.mode == 1) v = r;
            else if (P.mode == 2) { v = fma(P.a, r, old[m]); out2[m] = fma(P.beta, r, old[m]); }
            else if (P.beta == 0.0) v = P.a * (q[m] + dt * r);
            else v = P.beta * old[m] + P.a * (q[m] + dt * r);
            rate[c][m] = v;
        }
        store_run<NP>(P.dst + off + (size_t)c * G::NN, rate[c]);
        if (P.mode == 2 && P.beta != 0.0) store_run<NP>(P.dst2 + off + (size_t)c * G::NN, out2);
    }

    double vmax_local = 0.0;
    if (P.vmax && P.mode == 0) {
        // compute_cell_transport_speed (:450-514) of the updated state
        const double gm1 = P.gamma - 1.0;
#pragma unroll
        for (int m = 0; m < NP; m++) {
            const double inv = rcp_pos(rate[0][m]);
            const double sm = rate[1][m] * rate[1][m] + rate[2][m] * rate[2][m] + rate[3][m] * rate[3][m];
            const double pr = gm1 * (rate[4][m] - sm * (0.5 * inv));
            double conv = fabs(rate[1][m] * inv) * P.inv_h[0];
            if (DIM > 1) conv = fmax(conv, fabs(rate[2][m] * inv) * P.inv_h[1]);
            if (DIM > 2) conv = fmax(conv, fabs(rate[3][m] * inv) * P.inv_h[2]);
            const double c2 = P.gamma * pr * inv;
            // a non-positive or NaN c^2 (unphysical state) must reach the host as a NaN speed, not be clamped
            const double cs = (c2 > 0.0) ? sqrt_pos(c2) : __hiloint2double(0x7ff80000, 0);
            const double speed = P.max_eig * cs + conv;
            vmax_local = (speed > vmax_local || speed != speed) ? speed : vmax_local;
        }
    }
    return vmax_local;
}

// ---------------------------------------------------------------------------------------------------------------------
// Px (final): x-pencil owner.  Returns this thread's maximum transport speed of the values it wrote (0 if none).
// ---------------------------------------------------------------------------------------------------------------------
template <int DIM, int NP>
__device__ __forceinline__ double pencil_phase_final(const StageParams& P, double* smem, const int tid, const int64_t e0, const int sp,
                                                     const double dt, const PencilHalo& h) {
    using G = PGeo<DIM, NP>;
    const double2* const sRec = reinterpret_cast<const double2*>(smem + G::OFF_REC);
    double2* const sBuf = reinterpret_cast<double2*>(smem + G::OFF_BUF);
    const double2* const sEn = reinterpret_cast<const double2*>(smem + G::OFF_EN);
    if (tid >= G::USED) return 0.0;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    if (e >= P.elem_end) return 0.0;

    // blending factor of the element: fixed-order sum of the pencils' modal energies (persson_peraire...:96-122)
    double g0 = 0.0, g1 = 0.0;
    for (int k = 0; k < G::NPEN; k++) {
        const double2 en = sEn[le * G::NPEN + k];
        g0 += en.x;
        g1 += en.y;
    }
    const double alpha = blending_from_energies(g0, g1, P.ind_T, P.ind_sT);
    if (pe == 0 && P.alpha_out) P.alpha_out[(size_t)e * P.nsp + sp] = alpha;
    if (alpha > 0.0) {
        // troubled element (rare): the correction of this thread's nodes goes into their accumulator slots before the
        // regular path reads them (a run-time loop over the nodes: nothing of the regular path is live yet)
#pragma unroll 1
        for (int m = 0; m < NP; m++) {
            double corr[5];
            pencil_fv_correction<DIM, NP>(P, sRec, le, pe * NP + m, alpha, corr);
            const int slot = pslot<NP>(le * G::NN + pe * NP + m);
            double2 a = sBuf[slot], b = sBuf[G::NODES + slot], c = sBuf[2 * G::NODES + slot];
            a.x += corr[0]; a.y += corr[1]; b.x += corr[2]; b.y += corr[3]; c.x += corr[4];
            sBuf[slot] = a; sBuf[G::NODES + slot] = b; sBuf[2 * G::NODES + slot] = c;
        }
    }

#if WGPU_PENCIL_PREFETCH && !WGPU_HOST_EMU
    {
        // the epilogue (pencil_finish) reads u again and, in a second stage, the old destination: ask for the lines now, the
        // x pair fluxes below hide the latency (no registers held: prefetch instructions)
        const size_t off = ((size_t)e * P.nc + 5 * sp) * G::NN + pe * NP;
        const bool need_old = (P.mode == 2) || (P.mode == 0 && P.beta != 0.0);
        const double* const oldp = (P.mode == 2) ? P.sol_in : P.dst;
#pragma unroll
        for (int c = 0; c < 5; c++) {
#if WGPU_PENCIL_PREFETCH == 1
            if (P.mode != 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(P.u + off + (size_t)c * G::NN));
#endif
            if (need_old) asm volatile("prefetch.global.L2 [%0];" ::"l"(oldp + off + (size_t)c * G::NN));
        }
#if WGPU_PENCIL_PREFETCH >= 4
        // (4: also the next species' state, which its P0 loads right after the barrier that ends this phase)
        if (sp + 1 < P.nsp) {
#pragma unroll
            for (int c = 0; c < 5; c++) asm volatile("prefetch.global.L2 [%0];" ::"l"(P.u + off + (size_t)(5 + c) * G::NN));
        }
#endif
#if WGPU_PENCIL_PREFETCH >= 3
        // (3: also E and B for the Lorentz force: with the field system in its own kernel nobody has touched them yet)
        if (P.src_on && sp == 0) {
            const double* const fp = P.u + ((size_t)e * P.nc + 5 * P.nsp) * G::NN + pe * NP;
#pragma unroll
            for (int k = 0; k < 6; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(fp + (size_t)k * G::NN));
        }
#if WGPU_PENCIL_PREFETCH >= 6
        // (6: the fused field phases that follow the last species: phi, psi and, in a second stage, the old field values)
        if (P.mx_on && sp == P.nsp - 1 && P.nc >= 5 * P.nsp + 8) {
            const size_t fo = ((size_t)e * P.nc + 5 * P.nsp) * G::NN + pe * NP;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.u + fo + (size_t)6 * G::NN));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(P.u + fo + (size_t)7 * G::NN));
            if (need_old) {
#pragma unroll
                for (int k = 0; k < 8; k++) asm volatile("prefetch.global.L2 [%0];" ::"l"(oldp + fo + (size_t)k * G::NN));
            }
        }
#endif
#endif
    }
#endif
    double rate[5][NP];   // [component][node of the pencil]: the layout of the vector loads / stores below
    pencil_sums<DIM, NP, 0>(P, sRec, h, le, pe, [&](const int m, const double (&a)[5]) {
        const int slot = pslot<NP>(le * G::NN + pe * NP + m);
        const double2 x = sBuf[slot], y = sBuf[G::NODES + slot], z = sBuf[2 * G::NODES + slot];
        rate[0][m] = a[0] + x.x; rate[1][m] = a[1] + x.y; rate[2][m] = a[2] + y.x; rate[3][m] = a[3] + y.y; rate[4][m] = a[4] + z.x;
    });

    return pencil_finish<DIM, NP>(P, e, pe, sp, dt, rate);
}

// ---------------------------------------------------------------------------------------------------------------------
// ONE code body for the flux phases of all directions (WGPU_PENCIL_RTDIR, the default).  With a direction-templated body the
// kernel is ~9000 instructions of straight-line code per species pass (146 KB), the 12 resident warps of an SM sit at
// different places in it and 11-18 % of the stall samples were instruction-cache misses (profiles/README.md).  Here the
// direction is a run-time value that only enters (a) the node addressing and (b) a ROTATION of the velocity / momentum
// components when records are read and rates are written: the pencil's records are relabelled so that the pencil direction
// is component 0, the fluxes are evaluated for "direction 0" (compile-time constant inside the flux code, no selects
// there), and the five rates are relabelled back when they leave the registers.  Bit for bit the same arithmetic as the
// templated body (the relabelling is exact), a third of the code.
// ---------------------------------------------------------------------------------------------------------------------
template <int DIM, int NP>
__device__ __forceinline__ int pencil_base_rt(const int d, const int pe) {
    if (DIM == 2) return d == 0 ? pe * NP : pe;
    return d == 0 ? pe * NP : (d == 1 ? (pe % NP) + NP * NP * (pe / NP) : pe);
}
template <int DIM, int NP>
__device__ __forceinline__ int pencil_stride_rt(const int d) { return d == 0 ? 1 : (d == 1 ? NP : NP * NP); }

// relabel (x, y, z) components so that direction d becomes component 0 (cyclic): (a0,a1,a2) -> (a_d, a_{d+1}, a_{d+2})
__device__ __forceinline__ void rot3(const int d, double& a0, double& a1, double& a2) {
    const double x = a0, y = a1, z = a2;
    a0 = (d == 0) ? x : (d == 1 ? y : z);
    a1 = (d == 0) ? y : (d == 1 ? z : x);
    a2 = (d == 0) ? z : (d == 1 ? x : y);
}
// ... and back: (b0,b1,b2) in the rotated frame -> (x, y, z)
__device__ __forceinline__ void unrot3(const int d, double& b0, double& b1, double& b2) {
    const double p = b0, q = b1, r = b2;
    b0 = (d == 0) ? p : (d == 1 ? r : q);
    b1 = (d == 0) ? q : (d == 1 ? p : r);
    b2 = (d == 0) ? r : (d == 1 ? q : p);
}

template <int DIM, int NP>
__device__ __forceinline__ void pencil_halo_fetch_rt(const StageParams& P, const int d, const int v0, const int v1, const int pe,
                                                     const int64_t e0, const int64_t e_hi, const int sp, PencilHalo& h) {
    using G = PGeo<DIM, NP>;
    const int base = pencil_base_rt<DIM, NP>(d, pe), st = pencil_stride_rt<DIM, NP>(d);
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int v = side ? v1 : v0;
        const int nn = base + (side ? 0 : NP - 1) * st;   // the neighbour's end node: its opposite face, same tangential index
        h.nslot[side] = 0;
        if (v >= e0 && v < e_hi) {
            h.kind[side] = 0;
            h.nslot[side] = pslot<NP>((int)(v - e0) * G::NN + nn);
        } else {
            const double* src;
            size_t stride;
            if (v < 0) {
                h.kind[side] = 2;
                src = P.bres + (((size_t)(-1 - v) * P.nsp + sp) * 5) * G::NPEN + pe;
                stride = G::NPEN;
            } else if (v < P.n_elems) {
                h.kind[side] = 1;
                src = P.u + ((size_t)v * P.nc + 5 * sp) * G::NN + nn;
                stride = G::NN;
            } else {
                h.kind[side] = 1;
                src = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + 5 * sp) * G::NPEN + pe;
                stride = G::NPEN;
            }
#pragma unroll
            for (int c = 0; c < 5; c++) h.q[side][c] = src[(size_t)c * stride];
        }
    }
}

// Flux phase of direction d (run-time): d = 1 .. DIM-1 deposit their sums in the shared accumulator (d = 1 writes, the
// others add) and carry the indicator transform along; d = 0 is the final phase (x-pencil owner: blend, sources, stage
// update, store, transport speed).  On return h holds the ends of the pencil this thread owns in the next phase.
// Returns the thread's maximum transport speed (final phase with the fused CFL reduction, else 0).
template <int DIM, int NP, bool WITH_FINAL = true>
__device__ __forceinline__ double pencil_phase_flux_rt(const StageParams& P, double* smem, const int tid, const int64_t e0, const int sp,
                                                       const double dt, const int d, PencilHalo& h) {
    using G = PGeo<DIM, NP>;
    const double2* const sRec = reinterpret_cast<const double2*>(smem + G::OFF_REC);
    double2* const sBuf = reinterpret_cast<double2*>(smem + G::OFF_BUF);
    double2* const sEn = reinterpret_cast<double2*>(smem + G::OFF_EN);
    const int* const sNbr = reinterpret_cast<const int*>(smem + G::OFF_NBR);
    if (tid >= G::USED) return 0.0;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    const int64_t e_hi = (e0 + G::E < P.elem_end) ? e0 + G::E : P.elem_end;
    const bool final_phase = WITH_FINAL && (d == 0), first = (d == 1), last_ind = (d == DIM - 1);
    if (e >= P.elem_end) {
        if (last_ind) sEn[tid] = make_double2(0.0, 0.0);
        return 0.0;
    }
    const int base = le * G::NN + pencil_base_rt<DIM, NP>(d, pe), st = pencil_stride_rt<DIM, NP>(d);
    const double hig = P.hig, gamma = P.gamma;

    double ind[NP];
    if (!final_phase) {
        // indicator scratch of the pencil's nodes: transform along d now (cheap), written back with the sums
#pragma unroll
        for (int m = 0; m < NP; m++) ind[m] = sBuf[2 * G::NODES + pslot<NP>(base + m * st)].y;
        legendre_1d<NP>(P, ind);
        if (last_ind) {
            bool shell_pencil;   // a tangential mode index of this pencil is already NP-1 (pencil_energies)
            if (DIM == 2) shell_pencil = (pe == NP - 1);
            else shell_pencil = (pe % NP == NP - 1) || (pe / NP == NP - 1);
            double g0 = 0.0, g1 = 0.0;
#pragma unroll
            for (int k = 0; k < NP; k++) {
                const double sq = ind[k] * ind[k];
                if (shell_pencil || k == NP - 1) g1 += sq; else g0 += sq;
            }
            sEn[tid] = make_double2(g0, g1);
        }
    } else {
        // blending factor of the element: fixed-order sum of the pencils' modal energies (persson_peraire...:96-122)
        double g0 = 0.0, g1 = 0.0;
        for (int k = 0; k < G::NPEN; k++) {
            const double2 en = sEn[le * G::NPEN + k];
            g0 += en.x;
            g1 += en.y;
        }
        const double alpha = blending_from_energies(g0, g1, P.ind_T, P.ind_sT);
        if (pe == 0 && P.alpha_out) P.alpha_out[(size_t)e * P.nsp + sp] = alpha;
        if (alpha > 0.0) {
            // troubled element (rare): the correction of this thread's nodes goes into their accumulator slots before the
            // regular path reads them
#pragma unroll 1
            for (int m = 0; m < NP; m++) {
                double corr[5];
                pencil_fv_correction<DIM, NP>(P, sRec, le, pe * NP + m, alpha, corr);
                const int slot = pslot<NP>(le * G::NN + pe * NP + m);
                double2 a = sBuf[slot], b = sBuf[G::NODES + slot], c = sBuf[2 * G::NODES + slot];
                a.x += corr[0]; a.y += corr[1]; b.x += corr[2]; b.y += corr[3]; c.x += corr[4];
                sBuf[slot] = a; sBuf[G::NODES + slot] = b; sBuf[2 * G::NODES + slot] = c;
            }
        }
    }

    // ---- the pencil's records, relabelled so that the pencil direction is component 0 ---------------------------------
    Prim r[NP];
#pragma unroll
    for (int m = 0; m < NP; m++) {
        const int slot = pslot<NP>(base + m * st);
        r[m] = (m == 0 || m == NP - 1) ? get_rec<true>(sRec, G::NODES, slot) : get_rec<false>(sRec, G::NODES, slot);
        rot3(d, r[m].u0, r[m].u1, r[m].u2);
    }
    double acc[NP][5];
#pragma unroll
    for (int m = 0; m < NP; m++)
#pragma unroll
        for (int c = 0; c < 5; c++) acc[m][c] = 0.0;
    // ---- the two ends (see pencil_sums) --------------------------------------------------------------------------------
    const double cf = P.inv_hw[d];
    const int os = (h.kind[0] == 1) ? 0 : 1;
    Prim bo;
    {
        double qo[5];
#pragma unroll
        for (int c = 0; c < 5; c++) qo[c] = os ? h.q[1][c] : h.q[0][c];
        if (h.kind[0] == 1 || h.kind[1] == 1) bo = make_prim(qo[0], qo[1], qo[2], qo[3], qo[4], gamma);
        else bo = r[0];
        if (h.kind[0] == 1 || h.kind[1] == 1) rot3(d, bo.u0, bo.u1, bo.u2);
    }
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int m = side ? NP - 1 : 0;
        Prim b = get_rec<true>(sRec, G::NODES, h.nslot[side]);
        rot3(d, b.u0, b.u1, b.u2);
        if (h.kind[side] == 1) {
            if (side == os) b = bo;
            else {
                Prim cold;   // both ends outside
                make_prim_cold(h.q[side][0], h.q[side][1], h.q[side][2], h.q[side][3], h.q[side][4], gamma, &cold);
                b = cold;
                rot3(d, b.u0, b.u1, b.u2);
            }
        }
        double Fe[5], Dv[5], ibl;
        ec_flux_d(0, r[m], b, hig, Fe, ibl);
        es_dissipation(r[m], b, ibl, hig, Dv);
        double ct[5];
#pragma unroll
        for (int c = 0; c < 5; c++) ct[c] = cf * (side ? (Dv[c] - Fe[c]) : (Dv[c] + Fe[c]));
        if (h.kind[side] == 2) {
            // domain boundary: boundary_kernel's contribution (in x, y, z labels: relabel) plus the diagonal volume term
            Prim a = r[m];
            a.p = 0.5 * a.rho * a.ib;
            a.H = a.p * (2.0 * hig) + 0.5 * a.rho * a.q2 + a.p;
            double Fp[5], bq[5];
            phys_flux_d(0, a, Fp);
#pragma unroll
            for (int c = 0; c < 5; c++) bq[c] = h.q[side][c];
            rot3(d, bq[1], bq[2], bq[3]);
            const double sg = side ? -cf : cf;
#pragma unroll
            for (int c = 0; c < 5; c++) ct[c] = bq[c] + sg * Fp[c];
        }
#pragma unroll
        for (int c = 0; c < 5; c++) acc[m][c] += ct[c];
    }

    // ---- what the final phase needs besides the sums --------------------------------------------------------------------
    double rate[5][NP];   // final phase: [component][node], the layout of the vector loads / stores

    // ---- the pairs; node m is complete after its row --------------------------------------------------------------------
    const double s = -2.0 * P.inv_h[d];
    static_for<0, NP>([&](auto J) {
        constexpr int j = decltype(J)::value;
        static_for<j + 1, NP>([&](auto L) {
            constexpr int l = decltype(L)::value;
            double F[5], ibl;
            ec_flux_d(0, r[j], r[l], hig, F, ibl);
            const double wj = s * P.T.D[j * NP + l], wl = s * P.T.D[l * NP + j];
#pragma unroll
            for (int c = 0; c < 5; c++) {
                acc[j][c] = fma(wj, F[c], acc[j][c]);
                acc[l][c] = fma(wl, F[c], acc[l][c]);
            }
        });
        // node j is complete: back to x, y, z labels, then into the accumulator (or out of it, in the final phase)
        double a1 = acc[j][1], a2 = acc[j][2], a3 = acc[j][3];
        unrot3(d, a1, a2, a3);
        const int slot = pslot<NP>(base + j * st);
        if (final_phase) {
            const double2 x = sBuf[slot], y = sBuf[G::NODES + slot], z = sBuf[2 * G::NODES + slot];
            rate[0][j] = acc[j][0] + x.x; rate[1][j] = a1 + x.y; rate[2][j] = a2 + y.x; rate[3][j] = a3 + y.y; rate[4][j] = acc[j][4] + z.x;
        } else if (first) {
            sBuf[slot] = make_double2(acc[j][0], a1);
            sBuf[G::NODES + slot] = make_double2(a2, a3);
            sBuf[2 * G::NODES + slot] = make_double2(acc[j][4], ind[j]);
        } else {
            const double2 x = sBuf[slot], y = sBuf[G::NODES + slot], z = sBuf[2 * G::NODES + slot];
            sBuf[slot] = make_double2(acc[j][0] + x.x, a1 + x.y);
            sBuf[G::NODES + slot] = make_double2(a2 + y.x, a3 + y.y);
            sBuf[2 * G::NODES + slot] = make_double2(acc[j][4] + z.x, ind[j]);
        }
    });

    if (!final_phase) {
        const int nd = (d + 1 < DIM) ? d + 1 : 0;
        pencil_halo_fetch_rt<DIM, NP>(P, nd, sNbr[le * G::NFACE + 2 * nd], sNbr[le * G::NFACE + 2 * nd + 1], pe, e0, e_hi, sp, h);
        return 0.0;
    }
    if (WITH_FINAL) return pencil_finish<DIM, NP>(P, e, pe, sp, dt, rate);
    return 0.0;
}

// ---------------------------------------------------------------------------------------------------------------------
// The field system fused into the same kernel (mx_on): perfectly hyperbolic Maxwell fluxes for [Ex,Ey,Ez,Bx,By,Bz,phi,psi]
// (north_star kernel 4; new physics, see include/warpii_gpu.h::warpii_gpu_set_maxwell), the same pencil phases as a species:
//   F0 (x owner) fields -> shared planes   Fy, Fz  -(1/h) sum_l D[m][l] f_d(F_l) + Rusanov face terms at the two ends -> accumulator
//   Fx (x owner) x sums + accumulator + sources, stage update, store, the field system's share of the transport speed.
// Shared memory: the record planes 0-3 hold the fields [(Ex,Ey) (Ez,Bx) (By,Bz) (phi,psi)], the accumulator planes plus record
// plane 4 hold the 8 partial rates.  The flux is linear and sparse: ~150 FMAs per node against ~750 FP64 instructions per
// node and species of the fluid, and no extra pass over HBM.
// ---------------------------------------------------------------------------------------------------------------------
struct FieldHalo {
    double q[2][8];
    int kind[2];      // 0: in the patch (nslot), 1: values in q, 2: domain boundary (zero gradient)
    int nslot[2];
};

template <int D>
__device__ __forceinline__ void phm_flux_d(const StageParams& P, const double (&F)[8], double (&f)[8]) {
    constexpr int i1 = (D + 1) % 3, i2 = (D + 2) % 3;
#pragma unroll
    for (int i = 0; i < 8; i++) f[i] = 0.0;
    f[i1] = P.mx_c2 * F[3 + i2];            // -c^2 (e_d x B)
    f[i2] = -P.mx_c2 * F[3 + i1];
    f[3 + i1] = -F[i2];                     // e_d x E
    f[3 + i2] = F[i1];
    f[D] = P.mx_chi * P.mx_c2 * F[6];
    f[3 + D] = P.mx_gam * F[7];
    f[6] = P.mx_chi * F[D];
    f[7] = P.mx_gam * P.mx_c2 * F[3 + D];
}

template <int DIM, int NP, int D>
__device__ __forceinline__ void field_halo_fetch(const StageParams& P, const int v0, const int v1, const int pe, const int64_t e0,
                                                 const int64_t e_hi, FieldHalo& h) {
    using G = PGeo<DIM, NP>;
    const int nf0 = 5 * P.nsp;
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int v = side ? v1 : v0;
        h.nslot[side] = 0;
        if (v >= e0 && v < e_hi) {
            h.kind[side] = 0;
            h.nslot[side] = pslot<NP>((int)(v - e0) * G::NN + pencil_node<DIM, NP, D>(pe, side ? 0 : NP - 1));
        } else if (v < 0) {
            h.kind[side] = 2;
        } else {
            h.kind[side] = 1;
            const double* src;
            size_t stride;
            if (v < P.n_elems) {
                src = P.u + ((size_t)v * P.nc + nf0) * G::NN + pencil_node<DIM, NP, D>(pe, side ? 0 : NP - 1);
                stride = G::NN;
            } else {
                src = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + nf0) * G::NPEN + pe;
                stride = G::NPEN;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) h.q[side][k] = src[(size_t)k * stride];
        }
    }
}

__device__ __forceinline__ void get_fields(const double2* sFld, const int nodes, const int slot, double (&F)[8]) {
    const double2 a = sFld[slot], b = sFld[nodes + slot], c = sFld[2 * nodes + slot], d = sFld[3 * nodes + slot];
    F[0] = a.x; F[1] = a.y; F[2] = b.x; F[3] = b.y; F[4] = c.x; F[5] = c.y; F[6] = d.x; F[7] = d.y;
}

template <int DIM, int NP, int D, class Done>
__device__ __forceinline__ void field_sums(const StageParams& P, const double2* sFld, const FieldHalo& h, const int le, const int pe,
                                           Done&& done) {
    using G = PGeo<DIM, NP>;
    // Only the NP x 8 partial rates stay in registers: the nodes' fields are read from shared memory one node at a time,
    // turned into the flux and scattered to all nodes of the pencil with the column of D.
    double acc[NP][8];
#pragma unroll
    for (int m = 0; m < NP; m++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[m][k] = 0.0;
    const double cf = 0.5 * P.inv_hw[D];
    static_for<0, NP>([&](auto L_) {
        constexpr int l = decltype(L_)::value;
        double F[8], f[8];
        get_fields(sFld, G::NODES, pslot<NP>(le * G::NN + pencil_node<DIM, NP, D>(pe, l)), F);
        if (l == 0 || l == NP - 1) {
            // an end: (f(F_m).n - f*) / (h w_0) = (lambda dF - sgn f_d(dF)) / (2 h w_0), dF = F_p - F_m
            constexpr int side = (l == 0) ? 0 : 1;
            double Fn[8], dF[8], fn[8];
            get_fields(sFld, G::NODES, h.nslot[side], Fn);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const double other = (h.kind[side] == 0) ? Fn[k] : ((h.kind[side] == 1) ? h.q[side][k] : F[k]);
                dF[k] = other - F[k];
            }
            phm_flux_d<D>(P, dF, fn);
#pragma unroll
            for (int k = 0; k < 8; k++) acc[l][k] += cf * (P.mx_lam * dF[k] - (side ? fn[k] : -fn[k]));
        }
        // volume: -(1/h) D[m][l] f_d(F_l) to every node m of the pencil
        phm_flux_d<D>(P, F, f);
#pragma unroll
        for (int m = 0; m < NP; m++) {
            const double w = -P.inv_h[D] * P.T.D[m * NP + l];
#pragma unroll
            for (int k = 0; k < 8; k++) acc[m][k] = fma(w, f[k], acc[m][k]);
        }
    });
#pragma unroll
    for (int m = 0; m < NP; m++) done(m, acc[m]);
}

template <int DIM, int NP>
__device__ __forceinline__ void field_phase0(const StageParams& P, double* smem, const int tid, const int64_t e0, FieldHalo& next) {
    using G = PGeo<DIM, NP>;
    double2* const sFld = reinterpret_cast<double2*>(smem + G::OFF_REC);
    const int* const sNbr = reinterpret_cast<const int*>(smem + G::OFF_NBR);
    if (tid >= G::USED) return;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    const int64_t e_hi = (e0 + G::E < P.elem_end) ? e0 + G::E : P.elem_end;
    if (e >= P.elem_end) return;
    double F[8][NP];
    const double* src = P.u + ((size_t)e * P.nc + 5 * P.nsp) * G::NN + pe * NP;
#pragma unroll
    for (int k = 0; k < 8; k++) load_run<NP>(src + (size_t)k * G::NN, F[k]);
    field_halo_fetch<DIM, NP, 1>(P, sNbr[le * G::NFACE + 2], sNbr[le * G::NFACE + 3], pe, e0, e_hi, next);
#pragma unroll
    for (int m = 0; m < NP; m++) {
        const int slot = pslot<NP>(le * G::NN + pe * NP + m);
        sFld[slot] = make_double2(F[0][m], F[1][m]);
        sFld[G::NODES + slot] = make_double2(F[2][m], F[3][m]);
        sFld[2 * G::NODES + slot] = make_double2(F[4][m], F[5][m]);
        sFld[3 * G::NODES + slot] = make_double2(F[6][m], F[7][m]);
    }
}

template <int DIM, int NP, int D>
__device__ __forceinline__ void field_phase_mid(const StageParams& P, double* smem, const int tid, const int64_t e0, FieldHalo& h) {
    using G = PGeo<DIM, NP>;
    constexpr bool FIRST = (D == 1);
    constexpr int NEXT = (D + 1 < DIM) ? D + 1 : 0;
    double2* const sFld = reinterpret_cast<double2*>(smem + G::OFF_REC);   // planes 0-3 fields (read), plane 4 rates 6,7
    double2* const sBuf = reinterpret_cast<double2*>(smem + G::OFF_BUF);
    const int* const sNbr = reinterpret_cast<const int*>(smem + G::OFF_NBR);
    if (tid >= G::USED) return;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    const int64_t e_hi = (e0 + G::E < P.elem_end) ? e0 + G::E : P.elem_end;
    if (e >= P.elem_end) return;
    field_sums<DIM, NP, D>(P, sFld, h, le, pe, [&](const int m, const double (&a)[8]) {
        const int slot = pslot<NP>(le * G::NN + pencil_node<DIM, NP, D>(pe, m));
        if (FIRST) {
            sBuf[slot] = make_double2(a[0], a[1]);
            sBuf[G::NODES + slot] = make_double2(a[2], a[3]);
            sBuf[2 * G::NODES + slot] = make_double2(a[4], a[5]);
            sFld[4 * G::NODES + slot] = make_double2(a[6], a[7]);
        } else {
            const double2 x = sBuf[slot], y = sBuf[G::NODES + slot], z = sBuf[2 * G::NODES + slot], w = sFld[4 * G::NODES + slot];
            sBuf[slot] = make_double2(a[0] + x.x, a[1] + x.y);
            sBuf[G::NODES + slot] = make_double2(a[2] + y.x, a[3] + y.y);
            sBuf[2 * G::NODES + slot] = make_double2(a[4] + z.x, a[5] + z.y);
            sFld[4 * G::NODES + slot] = make_double2(a[6] + w.x, a[7] + w.y);
        }
    });
    field_halo_fetch<DIM, NP, NEXT>(P, sNbr[le * G::NFACE + 2 * NEXT], sNbr[le * G::NFACE + 2 * NEXT + 1], pe, e0, e_hi, h);
}

// Final-phase epilogue of the field system: sources, stage update, store, the field system's share of the transport speed.
template <int DIM, int NP>
__device__ __forceinline__ double field_finish(const StageParams& P, double* smem, const int64_t e, const int le, const int pe,
                                               const double dt, double (&rate)[8][NP]) {
    using G = PGeo<DIM, NP>;
    const double2* const sFld = reinterpret_cast<const double2*>(smem + G::OFF_REC);
    const size_t base = (size_t)e * P.nc * G::NN + pe * NP;   // node 0 of the pencil, component 0
    // sources: -J/eps0 on E, chi rho_c/eps0 on phi (species in order, as everywhere)
    if (P.src_on) {
        double J[3][NP], rc[NP];
#pragma unroll
        for (int m = 0; m < NP; m++) { J[0][m] = 0.0; J[1][m] = 0.0; J[2][m] = 0.0; rc[m] = 0.0; }
        for (int sp = 0; sp < P.nsp; sp++) {
            const double qm = P.qm[sp];
            double r[NP], mx[NP], my[NP], mz[NP];
            load_run<NP>(P.u + base + (size_t)(5 * sp) * G::NN, r);
            load_run<NP>(P.u + base + (size_t)(5 * sp + 1) * G::NN, mx);
            load_run<NP>(P.u + base + (size_t)(5 * sp + 2) * G::NN, my);
            load_run<NP>(P.u + base + (size_t)(5 * sp + 3) * G::NN, mz);
#pragma unroll
            for (int m = 0; m < NP; m++) { rc[m] += qm * r[m]; J[0][m] += qm * mx[m]; J[1][m] += qm * my[m]; J[2][m] += qm * mz[m]; }
        }
#pragma unroll
        for (int m = 0; m < NP; m++) {
            rate[0][m] += -J[0][m] * P.inv_eps0;
            rate[1][m] += -J[1][m] * P.inv_eps0;
            rate[2][m] += -J[2][m] * P.inv_eps0;
            rate[6][m] += P.chi * rc[m] * P.inv_eps0;
        }
    }
    const bool need_old = (P.mode == 2) || (P.mode == 0 && P.beta != 0.0);
    const double* const oldp = (P.mode == 2) ? P.sol_in : P.dst;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        const size_t off = base + (size_t)(5 * P.nsp + k) * G::NN;
        double F[NP], old[NP] = {}, out2[NP];
        const int slotbase = le * G::NN + pe * NP;
#pragma unroll
        for (int m = 0; m < NP; m++) {
            const double2 v = sFld[(k / 2) * G::NODES + pslot<NP>(slotbase + m)];
            F[m] = (k & 1) ? v.y : v.x;
        }
        if (need_old) load_run<NP>(oldp + off, old);
#pragma unroll
        for (int m = 0; m < NP; m++) {
            const double r = rate[k][m];
            double v;
            if (P.mode == 1) v = r;
            else if (P.mode == 2) { v = fma(P.a, r, old[m]); out2[m] = fma(P.beta, r, old[m]); }
            else if (P.beta == 0.0) v = P.a * (F[m] + dt * r);
            else v = P.beta * old[m] + P.a * (F[m] + dt * r);
            rate[k][m] = v;
        }
        store_run<NP>(P.dst + off, rate[k]);
        if (P.mode == 2 && P.beta != 0.0) store_run<NP>(P.dst2 + off, out2);
    }
    double vmax_local = 0.0;
    if (P.vmax && P.mode == 0) {
        vmax_local = P.mx_floor;
        if (P.src_on) {
            // plasma and cyclotron frequency of the UPDATED state (this thread wrote the species' part of dst itself)
            double wp2[NP], qmax = 0.0;
#pragma unroll
            for (int m = 0; m < NP; m++) wp2[m] = 0.0;
            for (int sp = 0; sp < P.nsp; sp++) {
                const double qm = P.qm[sp];
                double r[NP];
                load_run<NP>(P.dst + base + (size_t)(5 * sp) * G::NN, r);
#pragma unroll
                for (int m = 0; m < NP; m++) wp2[m] += qm * qm * r[m] * P.inv_eps0;
                qmax = fmax(qmax, fabs(qm));
            }
#pragma unroll
            for (int m = 0; m < NP; m++) {
                const double b2 = rate[3][m] * rate[3][m] + rate[4][m] * rate[4][m] + rate[5][m] * rate[5][m];
                const double omega = fmax(sqrt(wp2[m]), qmax * sqrt(b2));
                const double sp_ = P.mx_omega_factor * omega;
                vmax_local = (sp_ > vmax_local || sp_ != sp_) ? sp_ : vmax_local;
            }
        }
    }
    return vmax_local;
}

// Fx: returns the field system's share of the transport speed at this thread's nodes (0 unless the CFL reduction is fused)
template <int DIM, int NP>
__device__ __forceinline__ double field_phase_final(const StageParams& P, double* smem, const int tid, const int64_t e0, const double dt,
                                                    const FieldHalo& h) {
    using G = PGeo<DIM, NP>;
    const double2* const sFld = reinterpret_cast<const double2*>(smem + G::OFF_REC);
    const double2* const sBuf = reinterpret_cast<const double2*>(smem + G::OFF_BUF);
    if (tid >= G::USED) return 0.0;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    if (e >= P.elem_end) return 0.0;
    double rate[8][NP];
    field_sums<DIM, NP, 0>(P, sFld, h, le, pe, [&](const int m, const double (&a)[8]) {
        const int slot = pslot<NP>(le * G::NN + pe * NP + m);
        const double2 x = sBuf[slot], y = sBuf[G::NODES + slot], z = sBuf[2 * G::NODES + slot], w = sFld[4 * G::NODES + slot];
        rate[0][m] = a[0] + x.x; rate[1][m] = a[1] + x.y; rate[2][m] = a[2] + y.x; rate[3][m] = a[3] + y.y;
        rate[4][m] = a[4] + z.x; rate[5][m] = a[5] + z.y; rate[6][m] = a[6] + w.x; rate[7][m] = a[7] + w.y;
    });
    return field_finish<DIM, NP>(P, smem, e, le, pe, dt, rate);
}

// The field phases with a run-time direction (one code body, see pencil_phase_flux_rt): E and B are relabelled so that the
// pencil direction is component 0, the flux is evaluated for direction 0, the rates are relabelled back.
template <int DIM, int NP>
__device__ __forceinline__ void field_halo_fetch_rt(const StageParams& P, const int d, const int v0, const int v1, const int pe,
                                                    const int64_t e0, const int64_t e_hi, FieldHalo& h) {
    using G = PGeo<DIM, NP>;
    const int nf0 = 5 * P.nsp;
    const int base = pencil_base_rt<DIM, NP>(d, pe), st = pencil_stride_rt<DIM, NP>(d);
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int v = side ? v1 : v0;
        const int nn = base + (side ? 0 : NP - 1) * st;
        h.nslot[side] = 0;
        if (v >= e0 && v < e_hi) {
            h.kind[side] = 0;
            h.nslot[side] = pslot<NP>((int)(v - e0) * G::NN + nn);
        } else if (v < 0) {
            h.kind[side] = 2;
        } else {
            h.kind[side] = 1;
            const double* src;
            size_t stride;
            if (v < P.n_elems) {
                src = P.u + ((size_t)v * P.nc + nf0) * G::NN + nn;
                stride = G::NN;
            } else {
                src = P.ghost + ((size_t)(v - P.n_elems) * P.ncf + nf0) * G::NPEN + pe;
                stride = G::NPEN;
            }
#pragma unroll
            for (int k = 0; k < 8; k++) h.q[side][k] = src[(size_t)k * stride];
        }
    }
}

__device__ __forceinline__ void rot_fields(const int d, double (&F)[8]) {
    rot3(d, F[0], F[1], F[2]);
    rot3(d, F[3], F[4], F[5]);
}

// direction d = 1 .. DIM-1: sums into the accumulator; d = 0: final phase.  Returns the transport-speed share (final phase).
template <int DIM, int NP>
__device__ __forceinline__ double field_phase_flux_rt(const StageParams& P, double* smem, const int tid, const int64_t e0, const double dt,
                                                      const int d, FieldHalo& h) {
    using G = PGeo<DIM, NP>;
    double2* const sFld = reinterpret_cast<double2*>(smem + G::OFF_REC);   // planes 0-3 fields (read), plane 4 rates 6,7
    double2* const sBuf = reinterpret_cast<double2*>(smem + G::OFF_BUF);
    const int* const sNbr = reinterpret_cast<const int*>(smem + G::OFF_NBR);
    if (tid >= G::USED) return 0.0;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    const int64_t e_hi = (e0 + G::E < P.elem_end) ? e0 + G::E : P.elem_end;
    if (e >= P.elem_end) return 0.0;
    const bool final_phase = (d == 0), first = (d == 1);
    const int base = le * G::NN + pencil_base_rt<DIM, NP>(d, pe), st = pencil_stride_rt<DIM, NP>(d);
    double acc[NP][8];
#pragma unroll
    for (int m = 0; m < NP; m++)
#pragma unroll
        for (int k = 0; k < 8; k++) acc[m][k] = 0.0;
    const double cf = 0.5 * P.inv_hw[d], ihd = -P.inv_h[d];
    // the two ends first (their outside data was fetched a phase ahead and now leaves the registers)
#pragma unroll
    for (int side = 0; side < 2; side++) {
        const int l = side ? NP - 1 : 0;
        double F[8], Fn[8], dF[8], fn[8];
        get_fields(sFld, G::NODES, pslot<NP>(base + l * st), F);
        get_fields(sFld, G::NODES, h.nslot[side], Fn);
#pragma unroll
        for (int k = 0; k < 8; k++) dF[k] = (h.kind[side] == 2) ? 0.0 : ((h.kind[side] == 0) ? Fn[k] : h.q[side][k]) - F[k];
        rot_fields(d, dF);
        phm_flux_d<0>(P, dF, fn);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[l][k] = cf * (P.mx_lam * dF[k] - (side ? fn[k] : -fn[k]));
    }
    static_for<0, NP>([&](auto L_) {
        constexpr int l = decltype(L_)::value;
        double F[8], f[8];
        get_fields(sFld, G::NODES, pslot<NP>(base + l * st), F);
        rot_fields(d, F);
        phm_flux_d<0>(P, F, f);
#pragma unroll
        for (int m = 0; m < NP; m++) {
            const double w = ihd * P.T.D[m * NP + l];
#pragma unroll
            for (int k = 0; k < 8; k++) acc[m][k] = fma(w, f[k], acc[m][k]);
        }
    });
    double rate[8][NP];
#pragma unroll
    for (int m = 0; m < NP; m++) {
        unrot3(d, acc[m][0], acc[m][1], acc[m][2]);
        unrot3(d, acc[m][3], acc[m][4], acc[m][5]);
        const int slot = pslot<NP>(base + m * st);
        if (first) {
            sBuf[slot] = make_double2(acc[m][0], acc[m][1]);
            sBuf[G::NODES + slot] = make_double2(acc[m][2], acc[m][3]);
            sBuf[2 * G::NODES + slot] = make_double2(acc[m][4], acc[m][5]);
            sFld[4 * G::NODES + slot] = make_double2(acc[m][6], acc[m][7]);
        } else {
            const double2 x = sBuf[slot], y = sBuf[G::NODES + slot], z = sBuf[2 * G::NODES + slot], w = sFld[4 * G::NODES + slot];
            if (final_phase) {
                rate[0][m] = acc[m][0] + x.x; rate[1][m] = acc[m][1] + x.y; rate[2][m] = acc[m][2] + y.x; rate[3][m] = acc[m][3] + y.y;
                rate[4][m] = acc[m][4] + z.x; rate[5][m] = acc[m][5] + z.y; rate[6][m] = acc[m][6] + w.x; rate[7][m] = acc[m][7] + w.y;
            } else {
                sBuf[slot] = make_double2(acc[m][0] + x.x, acc[m][1] + x.y);
                sBuf[G::NODES + slot] = make_double2(acc[m][2] + y.x, acc[m][3] + y.y);
                sBuf[2 * G::NODES + slot] = make_double2(acc[m][4] + z.x, acc[m][5] + z.y);
                sFld[4 * G::NODES + slot] = make_double2(acc[m][6] + w.x, acc[m][7] + w.y);
            }
        }
    }
    if (!final_phase) {
        const int nd = (d + 1 < DIM) ? d + 1 : 0;
        field_halo_fetch_rt<DIM, NP>(P, nd, sNbr[le * G::NFACE + 2 * nd], sNbr[le * G::NFACE + 2 * nd + 1], pe, e0, e_hi, h);
        return 0.0;
    }
    return field_finish<DIM, NP>(P, smem, e, le, pe, dt, rate);
}

// ---------------------------------------------------------------------------------------------------------------------
// Field components (after the species): carried through unchanged by the reference's operator; with the
// two-fluid sources on, E gets -J/eps0 and phi gets chi rho_c/eps0.  x-pencil owner, Np consecutive nodes.
// ---------------------------------------------------------------------------------------------------------------------
template <int DIM, int NP>
__device__ __forceinline__ void pencil_phase_fields(const StageParams& P, const int tid, const int64_t e0, const double dt) {
    using G = PGeo<DIM, NP>;
    if (tid >= G::USED) return;
    const int le = tid / G::NPEN, pe = tid - le * G::NPEN;
    const int64_t e = e0 + le;
    if (e >= P.elem_end || P.nc <= 5 * P.nsp || P.fields_skip) return;
    const size_t base = (size_t)e * P.nc * G::NN + pe * NP;   // node 0 of the pencil, component 0
    double J[3][NP], rc[NP];
#pragma unroll
    for (int m = 0; m < NP; m++) { J[0][m] = 0.0; J[1][m] = 0.0; J[2][m] = 0.0; rc[m] = 0.0; }
    if (P.src_on) {
        for (int sp = 0; sp < P.nsp; sp++) {
            const double qm = P.qm[sp];
            double r[NP], mx[NP], my[NP], mz[NP];
            load_run<NP>(P.u + base + (size_t)(5 * sp) * G::NN, r);
            load_run<NP>(P.u + base + (size_t)(5 * sp + 1) * G::NN, mx);
            load_run<NP>(P.u + base + (size_t)(5 * sp + 2) * G::NN, my);
            load_run<NP>(P.u + base + (size_t)(5 * sp + 3) * G::NN, mz);
#pragma unroll
            for (int m = 0; m < NP; m++) { rc[m] += qm * r[m]; J[0][m] += qm * mx[m]; J[1][m] += qm * my[m]; J[2][m] += qm * mz[m]; }
        }
    }
    const bool need_old = (P.mode == 2) || (P.mode == 0 && P.beta != 0.0);
    const double* const oldp = (P.mode == 2) ? P.sol_in : P.dst;
#pragma unroll
    for (int k = 0; k < 8; k++) {   // fields_enabled means exactly these 8 components (five_moment.h:123-138)
        if (5 * P.nsp + k >= P.nc) break;
        const size_t off = base + (size_t)(5 * P.nsp + k) * G::NN;
        double f[NP], old[NP], out[NP], out2[NP];
        if (P.mode != 1) load_run<NP>(P.u + off, f);
        if (need_old) load_run<NP>(oldp + off, old);
#pragma unroll
        for (int m = 0; m < NP; m++) {
            double rate = 0.0;
            if (P.src_on) rate = (k < 3) ? -J[k < 3 ? k : 0][m] * P.inv_eps0 : (k == 6 ? P.chi * rc[m] * P.inv_eps0 : 0.0);
            double v;
            if (P.mode == 1) v = rate;
            else if (P.mode == 2) { v = fma(P.a, rate, old[m]); out2[m] = fma(P.beta, rate, old[m]); }
            else if (P.beta == 0.0) v = P.a * (f[m] + dt * rate);
            else v = P.beta * old[m] + P.a * (f[m] + dt * rate);
            out[m] = v;
        }
        store_run<NP>(P.dst + off, out);
        if (P.mode == 2 && P.beta != 0.0) store_run<NP>(P.dst2 + off, out2);
    }
}

}  // namespace wgpu
