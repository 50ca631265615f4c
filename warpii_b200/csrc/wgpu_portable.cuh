// Lets the device-side arithmetic headers (det_log.cuh, dgsem_physics.cuh, dgsem_pencil_stage.cuh) also compile with a
// plain host compiler.  The PRODUCT never does that: libwarpii_b200.so is built by nvcc for sm_100a only and has no CPU
// path.  The host build exists for tests/emu/ (test infrastructure): it executes the stage kernel's per-thread phase
// functions thread by thread, barrier by barrier, so that the kernel's indexing and arithmetic can be checked against
// the oracle in the CPU-only test tier (-m "not gpu") before a GPU is spent on it.
#pragma once

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define WGPU_HOST_EMU 0
#else
#define WGPU_HOST_EMU 1
#include <cmath>
#include <cstdint>
#include <cstring>
#ifndef __device__
#define __device__
#endif
#ifndef __host__
#define __host__
#endif
#ifndef __noinline__
#define __noinline__
#endif
#ifndef __forceinline__
#define __forceinline__ inline
#endif
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
// round-to-nearest single operations: what plain C++ arithmetic is when the file is compiled with -ffp-contract=off
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline int __double2hiint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u >> 32); }
static inline int __double2loint(double x) { uint64_t u; std::memcpy(&u, &x, 8); return (int)(u & 0xffffffffu); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t u = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; std::memcpy(&x, &u, 8); return x;
}
using std::fma;
using std::fabs;
using std::sqrt;
using std::exp;
using std::log;
using std::fmax;
#endif

namespace wgpu {

// seeds of the Newton iterations (MUFU.RCP64H / MUFU.RSQ64H on the device: ~9 good bits)
__device__ __forceinline__ double rcp_seed(const double x) {
#if !WGPU_HOST_EMU
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return (double)(1.0f / (float)x);
#endif
}
__device__ __forceinline__ double rsqrt_seed(const double x) {
#if !WGPU_HOST_EMU
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    return y;
#else
    return (double)(1.0f / std::sqrt((float)x));
#endif
}

}  // namespace wgpu
