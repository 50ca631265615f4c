// Small device helpers and the (dim, Np) dispatch shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dgsem_kernels.cuh"

namespace wgpu {

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

__device__ __forceinline__ double block_max(double v, double* s_red) {
    for (int o = 16; o > 0; o >>= 1) v = nan_max(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwarps = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double r = s_red[0];
    for (int i = 1; i < nwarps; i++) r = nan_max(r, s_red[i]);
    return r;
}

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

}  // namespace wgpu
