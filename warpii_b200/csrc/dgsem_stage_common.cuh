// Pieces shared by the stage kernels (Cartesian: dgsem_stage_kernel.cu, general geometry: dgsem_general_kernel.cu):
// block geometry and shared-memory layout, primitive records, the blending function, cp.async and group barriers.
#pragma once
#include "dgsem_common.cuh"
#include "dgsem_physics.cuh"

#ifndef WGPU_RESIDENT_THREADS
#define WGPU_RESIDENT_THREADS 0     // 0: 640 threads per SM in 1D/2D (5 blocks of 128, 96 registers), 512 in 3D (4 blocks, 128 registers)
#endif

namespace wgpu {

constexpr int kPS = 14;   // doubles per node in the shared primitive table (12 used; 14 keeps 16-byte loads conflict-free)

// Offset (in doubles) of node n's record in the shared primitive table.  Records are 7 sixteen-byte units apart, so the
// 8 lanes of a quarter-warp reading 8 consecutive nodes (or two x-rows of another y/z position, which is what the pair
// and face phases do) hit 8 different bank groups.  With an even NP two xy-planes are a multiple of 8 units apart; one
// unit of padding per plane separates them (3D: 2x fewer wavefronts on the pair loads, measured and modelled).
template <int NP>
__device__ __forceinline__ constexpr int prim_off(const int n) {
    return (NP % 2 == 0) ? n * kPS + 2 * (n / (NP * NP)) : n * kPS;
}
template <int NP>
constexpr int prim_table_doubles(const int nodes) {
    return (NP % 2 == 0) ? nodes * kPS + 2 * (nodes / (NP * NP) + 1) : nodes * kPS;
}

// SHUF: the pair fluxes of 2D p=3 elements are exchanged with warp shuffles instead of a shared-memory table (an element
// is 16 consecutive lanes, so both endpoints of every pencil pair sit in one warp): 160 B of shared memory per thread less.
// Measured on C2 (stage kernel): table 0.310 ms; shuffles at 5 blocks/SM (96 registers) 0.301 ms, 6 blocks (80 registers)
// 0.308, 7 blocks (72) 0.301, 8 blocks (64) 0.329: the freed shared memory would allow 10 blocks, but the extra warps pay
// for themselves in spills, so the gain is the table traffic alone.
#ifndef WGPU_SHUFFLE_PAIRS
#define WGPU_SHUFFLE_PAIRS 1
#endif
#ifndef WGPU_SHUFFLE_RESIDENT
#define WGPU_SHUFFLE_RESIDENT 640
#endif
__host__ __device__ constexpr bool shuffle_pairs(int dim, int np) { return WGPU_SHUFFLE_PAIRS && dim == 2 && np == 4; }

template <int DIM, int NP, bool SHUF = false>
struct Geo {
    static constexpr int NN = ipow_c(NP, DIM);          // nodes per element
    static constexpr int NF = ipow_c(NP, DIM - 1);      // nodes per face = pencils per direction
    static constexpr int G = elems_per_block(DIM, NP);  // elements per block
    static constexpr int NODES = NN * G;                // nodes (= threads) per block
    static constexpr int NFACE = 2 * DIM;
    static constexpr int NSLOT = G * NFACE * NF;        // face-node result slots per block
    // Node pairs of a pencil are grouped by cyclic distance c = 1..NP/2: class c holds the pairs (j, (j+c) mod NP).  A full
    // class has one pair per node (its "first" endpoint j), the half class c = NP/2 of an even NP one per node with
    // j < NP/2.  One pair-flux slot per (direction, class, first-endpoint node).
    static constexpr int NFULL = (NP - 1) / 2;          // full classes per direction
    static constexpr int HALF = (NP % 2 == 0) ? 1 : 0;  // is there a half class
    static constexpr int NCL = NFULL + HALF;            // classes per direction
    static constexpr int HTASKS = HALF * DIM * (NN / 2);            // half-class pair tasks per element
    static constexpr int HROUNDS = (HTASKS + NN - 1) / NN;          // ... per thread (last round may be partial)
    static constexpr int PLANE = G * DIM * NCL * NN;    // pair-flux slots per block = doubles per component plane
    static constexpr int THREADS = NODES;
    // threads that synchronise among themselves after the node phase: whole warps holding whole elements
    static constexpr int GROUP = (32 % NN == 0 && NODES % 32 == 0) ? 32 : ((NN % 32 == 0 && NODES / NN <= 15) ? NN : NODES);
    // measured: 2D p=3 gains 6 % from the fifth block per SM despite 56 bytes of spills; 3D p=3 loses 9 %
    static constexpr int RESIDENT = SHUF ? WGPU_SHUFFLE_RESIDENT : (WGPU_RESIDENT_THREADS > 0 ? WGPU_RESIDENT_THREADS : (DIM <= 2 ? 640 : 512));
    static constexpr int MIN_BLOCKS = (RESIDENT / THREADS) > 0 ? (RESIDENT / THREADS) : 1;
    // dynamic shared memory, in doubles (every offset even => 16-byte aligned)
    static constexpr int even(int x) { return (x + 1) & ~1; }
    static constexpr int OFF_D = 0;
    static constexpr int OFF_V = even(OFF_D + NP * NP);
    static constexpr int OFF_W = even(OFF_V + NP * NP);
    static constexpr int OFF_P = OFF_W + 8;                                 // primitive records of the block's nodes
    static constexpr int OFF_A = even(OFF_P + prim_table_doubles<NP>(NODES));                      // [2][NODES] indicator scratch
    static constexpr int OFF_PAIR = OFF_A + 2 * NODES;                      // [5][PLANE] pair fluxes, component planes
    static constexpr int OFF_FACE = even(OFF_PAIR + (SHUF ? 0 : 5 * PLANE)); // [5][NSLOT] face-node records, component planes
    static constexpr int OFF_ALPHA = even(OFF_FACE + 5 * NSLOT);            // [G]
    static constexpr int OFF_RED = even(OFF_ALPHA + G);                     // [32]
    static constexpr int SMEM_DOUBLES = OFF_RED + 32;
};

// node record: [rho u0 | u1 u2 | beta lrho | lbeta q2 | p H | lam ib | - -]
template <int NP>
__device__ __forceinline__ void store_prim(double* sP, const int n, const Prim& P) {
    double2* r = reinterpret_cast<double2*>(sP + prim_off<NP>(n));
    r[0] = make_double2(P.rho, P.u0);
    r[1] = make_double2(P.u1, P.u2);
    r[2] = make_double2(P.beta, P.lrho);
    r[3] = make_double2(P.lbeta, P.q2);
    r[4] = make_double2(P.p, P.H);
    r[5] = make_double2(P.lam, P.ib);
}
// the 8 fields the entropy-conserving flux needs
template <int NP>
__device__ __forceinline__ Prim load_prim_ec(const double* sP, const int n) {
    const double2* r = reinterpret_cast<const double2*>(sP + prim_off<NP>(n));
    const double2 a = r[0], b = r[1], c = r[2], d = r[3];
    Prim o;
    o.rho = a.x; o.u0 = a.y; o.u1 = b.x; o.u2 = b.y; o.beta = c.x; o.lrho = c.y; o.lbeta = d.x; o.q2 = d.y;
    o.p = 0.0; o.H = 0.0; o.lam = 0.0; o.ib = 0.0;
    return o;
}
// the 6 fields the physical flux needs
template <int NP>
__device__ __forceinline__ Prim load_prim_phys(const double* sP, const int n) {
    const double2* r = reinterpret_cast<const double2*>(sP + prim_off<NP>(n));
    const double2 a = r[0], b = r[1], e = r[4];
    Prim o;
    o.rho = a.x; o.u0 = a.y; o.u1 = b.x; o.u2 = b.y; o.p = e.x; o.H = e.y;
    o.beta = 0.0; o.lrho = 0.0; o.lbeta = 0.0; o.q2 = 0.0; o.lam = 0.0; o.ib = 0.0;
    return o;
}
template <int NP>
__device__ __forceinline__ Prim load_prim(const double* sP, const int n) {
    const double2* r = reinterpret_cast<const double2*>(sP + prim_off<NP>(n));
    const double2 a = r[0], b = r[1], c = r[2], d = r[3], e = r[4], f = r[5];
    Prim o;
    o.rho = a.x; o.u0 = a.y; o.u1 = b.x; o.u2 = b.y; o.beta = c.x; o.lrho = c.y; o.lbeta = d.x; o.q2 = d.y;
    o.p = e.x; o.H = e.y; o.lam = f.x; o.ib = f.y;
    return o;
}
// 8-byte asynchronous global -> shared copy (LDGSTS): the data of a later phase travels while this thread computes
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int GROUP, int NODES>
__device__ __forceinline__ void group_sync(const int tid) {
    if (GROUP == 32) __syncwarp();
    else if (GROUP == NODES) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + tid / GROUP), "n"(GROUP) : "memory");
}

// Half-class pair tasks of an even NP: task q (0 <= q < NN/2) of direction d is the node whose coordinate along d is below
// NP/2, numbered with d's extent halved.  Called with d known at compile time (the extents are then literals: no integer
// division at run time, which is what the generic decode cost: ~50 instructions per node).
template <int DIM, int NP>
__device__ __forceinline__ int half_class_node(const int d, int q) {
    int n = 0, mul = 1;
#pragma unroll
    for (int a = 0; a < DIM; a++) {
        const int ext = (a == d) ? NP / 2 : NP;
        n += (q % ext) * mul;
        q /= ext;
        mul *= NP;
    }
    return n;
}

// first node of pencil pe (= tangential index) in direction d
template <int DIM, int NP>
__device__ __forceinline__ int pencil_first_node(const int d, const int pe) {
    if (DIM == 1) return 0;
    if (DIM == 2) return d == 0 ? NP * pe : pe;
    const int t0 = pe % NP, t1 = pe / NP;
    return d == 0 ? NP * (t0 + NP * t1) : (d == 1 ? (t0 + NP * NP * t1) : (t0 + NP * t1));
}

}  // namespace wgpu
