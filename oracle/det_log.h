// oracle/det_log.h -- deterministic natural logarithm for the CPU oracle (TEST INFRASTRUCTURE).
//
// Why it exists.  The reference's ln_avg (src/five_moment/euler.h:118-125) evaluates
// (b - a) / (ln b - ln a); for nearly equal a, b (adjacent nodes on a fine mesh) this amplifies a last-bit
// difference of ln by |ln a| / |ln b - ln a|, up to ~5e5 just before the 1e6 switch to the arithmetic mean.
// The reference takes ln from the platform's libm, and libms (glibc's FMA / SSE2 variants, CUDA's log, ...) differ
// in the last bit for ~0.5 % of arguments, so the reference's own RHS is only reproducible to ~1e-10 across
// platforms on BASELINE-sized meshes (tests/test_oracle_golden.py::test_libm_sensitivity measures it).
// To compare the CUDA path with the CPU restatement at the 1e-12 level both sides therefore use THIS logarithm:
// a fixed sequence of IEEE-754 binary64 operations (explicit fused multiply-adds, no other contraction, one correctly
// rounded division), hence
// bit-identical on x86-64 and on sm_100a (warpii_b200/csrc/det_log.cuh is the same sequence).
//
// Algorithm: the classical argument reduction x = 2^k m, m in [sqrt(2)/2, sqrt(2)), s = f/(2+f) with f = m - 1, and
// the degree-14 minimax polynomial in s of Sun's fdlibm e_log.c (coefficients Lg1..Lg7 as published there); error
// < 1 ulp.  Arguments that are not positive normal numbers go to libm (never the case for rho, beta of a valid state).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace detlog {

// (a build that contracts multiply-adds elsewhere keeps this function as written: oracle/Makefile, liboracle_fma.so)
#ifndef DETLOG_ATTR
#define DETLOG_ATTR
#endif

inline DETLOG_ATTR double det_log(double x) {
    uint64_t ix;
    std::memcpy(&ix, &x, sizeof ix);
    uint32_t hx = (uint32_t)(ix >> 32);
    if (hx < 0x00100000u || hx >= 0x7ff00000u) return std::log(x);   // zero, subnormal, negative, inf, nan
    static const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                        Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
                        Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
                        Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
                        Lg7 = 1.479819860511658591e-01;
    // reduce x into [sqrt(2)/2, sqrt(2))
    hx += 0x3ff00000u - 0x3fe6a09eu;
    const int k = (int)(hx >> 20) - 0x3ff;
    hx = (hx & 0x000fffffu) + 0x3fe6a09eu;
    ix = ((uint64_t)hx << 32) | (ix & 0xffffffffu);
    double m;
    std::memcpy(&m, &ix, sizeof m);
    const double f = m - 1.0;
    const double hfsq = (0.5 * f) * f;
    const double s = f / (2.0 + f);
    const double z = s * s;
    const double w = z * z;
    // std::fma is the IEEE fused multiply-add (one rounding): vfmadd on x86-64-v3, FMA on sm_100a
    const double t1 = w * std::fma(w, std::fma(w, Lg6, Lg4), Lg2);
    const double t2 = z * std::fma(w, std::fma(w, std::fma(w, Lg7, Lg5), Lg3), Lg1);
    const double R = t2 + t1;
    const double dk = (double)k;
    double r = std::fma(s, hfsq + R, dk * ln2_lo);
    r = r - hfsq;
    r = r + f;
    return std::fma(dk, ln2_hi, r);
}

}  // namespace detlog
