// Small device helpers and the (dim, Np) dispatch shared by the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "dgsem_kernels.cuh"

namespace wgpu {

// (stride_of, face_node_index, node_of_face_node and nan_max live in dgsem_kernels.cuh: host-portable)

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
