// The pencil-decomposed stage kernel for sm_100a: the __global__ wrapper around the per-thread phase functions of
// dgsem_pencil_stage.cuh (design and reference citations there).  Used for Cartesian boxes in 2-D and 3-D with
// 3 <= Np <= 5 nodes per direction; every other (dim, Np) keeps the node-per-thread kernel of dgsem_stage_kernel.cu.
#include "dgsem_common.cuh"
#include "dgsem_pencil_stage.cuh"

namespace wgpu {

// resident blocks per SM the register allocation is asked to allow (shared memory allows 3 blocks of the 512-node patches)
#ifndef WGPU_PENCIL_MIN_BLOCKS
#define WGPU_PENCIL_MIN_BLOCKS 3
#endif

// WGPU_PENCIL_SUB > 1 (tuning variant): one CTA carries SUB patches side by side ("sub-blocks" of THREADS threads, each with
// its own shared-memory image), so that the 4 SUB warps of a CTA walk the same stretch of the (80 KB, straight-line) code at
// about the same time and share its instruction-cache lines.  WGPU_PENCIL_SUB_HARD = 1 separates the phases with the CTA
// barrier (lock step), 0 with one named barrier per sub-block (they start together and drift).
#ifndef WGPU_PENCIL_LOOKAHEAD_MULT
#define WGPU_PENCIL_LOOKAHEAD_MULT 1   // how many block generations ahead the state prefetch looks (measured: 1 and 2 tie)
#endif
#ifndef WGPU_PENCIL_SUB
#define WGPU_PENCIL_SUB 1
#endif
#ifndef WGPU_PENCIL_SUB_HARD
#define WGPU_PENCIL_SUB_HARD 0
#endif
#if WGPU_PENCIL_SUB == 1 || WGPU_PENCIL_SUB_HARD
#define PBAR() __syncthreads()
#else
#define PBAR() asm volatile("bar.sync %0, %1;" ::"r"(sub + 1), "n"(G::THREADS) : "memory")
#endif
constexpr int pencil_cta_blocks() { return WGPU_PENCIL_MIN_BLOCKS / WGPU_PENCIL_SUB > 0 ? WGPU_PENCIL_MIN_BLOCKS / WGPU_PENCIL_SUB : 1; }

template <int DIM, int NP>
__global__ void __launch_bounds__(PGeo<DIM, NP>::THREADS * WGPU_PENCIL_SUB, pencil_cta_blocks()) pencil_stage_kernel(const StageParams P) {
    using G = PGeo<DIM, NP>;
    static_assert(G::SMEM_DOUBLES % 2 == 0, "sub-block images stay 16-byte aligned");
    extern __shared__ __align__(16) double smem_cta[];
    // device-resident time loop: "finished" flag and dt live in global memory
    const int skip = P.skip_dev ? *P.skip_dev : 0;
    const double dt = P.dt_dev ? *P.dt_dev : P.dt;
    if (skip) return;   // uniform
#if WGPU_PENCIL_SUB == 1
    const int tid = threadIdx.x;
    const int64_t e0 = P.elem_begin + (int64_t)blockIdx.x * G::E;
    double* const smem = smem_cta;
#else
    const int sub = threadIdx.x / G::THREADS;
    const int tid = threadIdx.x - sub * G::THREADS;
    const int64_t e0 = P.elem_begin + ((int64_t)blockIdx.x * WGPU_PENCIL_SUB + sub) * G::E;
    double* const smem = smem_cta + (size_t)sub * G::SMEM_DOUBLES;
#if !WGPU_PENCIL_SUB_HARD
    if (e0 >= P.elem_end) return;   // a whole sub-block without a patch (its named barrier has no other participant)
#endif
#endif
    double vmax_local = 0.0;
    PencilHalo halo;
    for (int sp = 0; sp < P.nsp; sp++) {
        if (sp > 0) PBAR();   // shared-memory reuse across species
#if WGPU_PENCIL_TMA
        if (pencil_tma_ok<DIM, NP>() && tid == 0) {
            // arm the barrier and let the bulk-copy engine fetch the patch's state block of this species (A/B variant)
            void* const bar = smem + G::OFF_RED + 30;
            if (sp == 0) mbar_init(bar, 1);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses of the area come first
            const int64_t e_hi = (e0 + G::E < P.elem_end) ? e0 + G::E : P.elem_end;
            const unsigned per_elem = 5u * G::NN * (unsigned)sizeof(double);
            mbar_expect_tx(bar, per_elem * (unsigned)(e_hi - e0));
            for (int64_t e = e0; e < e_hi; e++)
                bulk_g2s(smem + G::OFF_REC + (size_t)(e - e0) * 5 * G::NN, P.u + ((size_t)e * P.nc + 5 * sp) * G::NN, per_elem, bar);
        }
        if (pencil_tma_ok<DIM, NP>() && sp == 0) PBAR();   // the barrier object is initialised before anybody waits on it
#endif
        pencil_phase0<DIM, NP>(P, smem, tid, e0, sp, halo);
#if WGPU_PENCIL_RTDIR == 2
        // the mid phases (y, z) through one copy of the flux code, the final x phase through its own
#pragma unroll 1
        for (int ph = 1; ph < DIM; ph++) {
            PBAR();
            pencil_phase_flux_rt<DIM, NP, false>(P, smem, tid, e0, sp, dt, ph, halo);
        }
        PBAR();
        vmax_local = nan_max(vmax_local, pencil_phase_final<DIM, NP>(P, smem, tid, e0, sp, dt, halo));
#elif WGPU_PENCIL_RTDIR
        // directions 1, .., DIM-1, 0 through ONE copy of the flux code (a real loop: the body must not be replicated)
#pragma unroll 1
        for (int ph = 1; ph <= DIM; ph++) {
            PBAR();
            vmax_local = nan_max(vmax_local, pencil_phase_flux_rt<DIM, NP>(P, smem, tid, e0, sp, dt, ph < DIM ? ph : 0, halo));
        }
#else
        PBAR();
        pencil_phase_mid<DIM, NP, 1>(P, smem, tid, e0, sp, halo);
        PBAR();
        if (DIM == 3) {
            pencil_phase_mid<DIM, NP, DIM - 1>(P, smem, tid, e0, sp, halo);
            PBAR();
        }
        vmax_local = nan_max(vmax_local, pencil_phase_final<DIM, NP>(P, smem, tid, e0, sp, dt, halo));
#endif
    }
    if (P.mx_on && P.nc >= 5 * P.nsp + 8) {
        // the field system, same phases (uniform branch: kernel parameter)
        FieldHalo fh;
        PBAR();   // the last species' final phase still reads the record planes
        field_phase0<DIM, NP>(P, smem, tid, e0, fh);
#if WGPU_PENCIL_RTDIR
#pragma unroll 1
        for (int ph = 1; ph <= DIM; ph++) {
            PBAR();
            vmax_local = nan_max(vmax_local, field_phase_flux_rt<DIM, NP>(P, smem, tid, e0, dt, ph < DIM ? ph : 0, fh));
        }
#else
        PBAR();
        field_phase_mid<DIM, NP, 1>(P, smem, tid, e0, fh);
        PBAR();
        if (DIM == 3) {
            field_phase_mid<DIM, NP, DIM - 1>(P, smem, tid, e0, fh);
            PBAR();
        }
        vmax_local = nan_max(vmax_local, field_phase_final<DIM, NP>(P, smem, tid, e0, dt, fh));
#endif
    } else {
        pencil_phase_fields<DIM, NP>(P, tid, e0, dt);
    }
    if (P.vmax && P.mode == 0) {
#if WGPU_PENCIL_SUB == 1
        const double m = block_max(vmax_local, smem + G::OFF_RED);
#else
        double m = vmax_local;   // per sub-block
        for (int o = 16; o > 0; o >>= 1) m = nan_max(m, __shfl_xor_sync(0xffffffffu, m, o));
        PBAR();
        if ((tid & 31) == 0) smem[G::OFF_RED + (tid >> 5)] = m;
        PBAR();
        m = smem[G::OFF_RED];
        for (int i = 1; i < G::THREADS / 32; i++) m = nan_max(m, smem[G::OFF_RED + i]);
#endif
        if (tid == 0) atomicMax(P.vmax, (unsigned long long)__double_as_longlong(m));
    }
}

#define WGPU_PENCIL_DISPATCH(dim, Np, CALL)                                                                      \
    do {                                                                                                         \
        if ((dim) == 2) {                                                                                        \
            switch (Np) { case 3: { CALL(2, 3); } break; case 4: { CALL(2, 4); } break; case 5: { CALL(2, 5); } break; } \
        } else if ((dim) == 3) {                                                                                 \
            switch (Np) { case 3: { CALL(3, 3); } break; case 4: { CALL(3, 4); } break; case 5: { CALL(3, 5); } break; } \
        }                                                                                                        \
    } while (0)

bool pencil_available(int dim, int Np) { return (dim == 2 || dim == 3) && Np >= 3 && Np <= 5; }

int pencil_patch_elems(int dim, int Np) { return pencil_available(dim, Np) ? pencil_elems(dim, Np) : 0; }

int pencil_smem_bytes(int dim, int Np) {
    int bytes = 0;
#define CALL(D_, N_) { bytes = PGeo<D_, N_>::SMEM_DOUBLES * (int)sizeof(double); }
    WGPU_PENCIL_DISPATCH(dim, Np, CALL);
#undef CALL
    return bytes;
}

int prepare_pencil_kernels(int dim, int Np) {
    cudaError_t err = cudaSuccess;
#define CALL(D_, N_)                                                                                              \
    {                                                                                                             \
        err = cudaFuncSetAttribute(pencil_stage_kernel<D_, N_>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                   PGeo<D_, N_>::SMEM_DOUBLES * WGPU_PENCIL_SUB * (int)sizeof(double));            \
        if (err == cudaSuccess)                                                                                   \
            err = cudaFuncSetAttribute(pencil_stage_kernel<D_, N_>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                       cudaSharedmemCarveoutMaxShared);                                           \
    }
    WGPU_PENCIL_DISPATCH(dim, Np, CALL);
#undef CALL
    return err == cudaSuccess ? 0 : 1;
}

void launch_pencil_stage(int dim, int Np, const StageParams& P, cudaStream_t s) {
    const int64_t n = P.elem_end - P.elem_begin;
    if (n <= 0) return;
#define CALL(D_, N_)                                                                                              \
    {                                                                                                             \
        using G = PGeo<D_, N_>;                                                                                   \
        const int64_t per_cta = (int64_t)G::E * WGPU_PENCIL_SUB;                                                  \
        const int64_t blocks = (n + per_cta - 1) / per_cta;                                                       \
        static int resident = -1;   /* blocks of this instantiation the device holds at once */                  \
        if (resident < 0) {                                                                                       \
            int per_sm = 0, dev = 0, sms = 0;                                                                     \
            cudaGetDevice(&dev);                                                                                  \
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);                                    \
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, pencil_stage_kernel<D_, N_>,                   \
                                                          G::THREADS * WGPU_PENCIL_SUB,                            \
                                                          G::SMEM_DOUBLES * WGPU_PENCIL_SUB * sizeof(double));     \
            resident = per_sm * sms;                                                                              \
        }                                                                                                         \
        StageParams Q = P;                                                                                        \
        Q.lookahead = (blocks > resident) ? (int64_t)resident * per_cta * WGPU_PENCIL_LOOKAHEAD_MULT : 0;         \
        pencil_stage_kernel<D_, N_><<<(unsigned)blocks, G::THREADS * WGPU_PENCIL_SUB,                             \
                                      G::SMEM_DOUBLES * WGPU_PENCIL_SUB * sizeof(double), s>>>(Q);                \
    }
    WGPU_PENCIL_DISPATCH(dim, Np, CALL);
#undef CALL
}

}  // namespace wgpu
