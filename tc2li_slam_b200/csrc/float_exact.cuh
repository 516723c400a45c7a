// float_exact.cuh — the floating-point pieces of the path, evaluated so that they reproduce the CPU results bit for bit.
//
//  * fast_atan2_deg : cv::fastAtan2's scalar polynomial (used by IC_Angle, ORBextractor.cc:76) in the exact
//    float operation order, no FMA (explicit __f*_rn intrinsics; IEEE division).
//  * sinf_cosf_glibc: what glibc >= 2.28 computes for sinf()/cosf() (computeOrbDescriptor calls them through
//    `using namespace std`, ORBextractor.cc:86-88): the float argument is widened to double, reduced by a multiple of pi/2
//    with a 2^24-prescaled 2/pi, and a degree-7 (sine) or degree-8 (cosine) polynomial is evaluated in double and rounded
//    once to float.  Those values are NOT the correctly rounded sine/cosine in ~1 % of arguments, so CUDA's sinf/cosf
//    (and a correctly rounded evaluation) would flip descriptor bits; this routine follows the published algorithm
//    (ARM optimized-routines sincosf, as adopted by glibc sysdeps/ieee754/flt-32/s_sincosf.h) with its coefficients.
//    Valid for the path's arguments: 0 <= x < 120.
//    tests/test_float_exact.py checks it against the host libm over a dense sweep of [0, 2*pi].
#pragma once
#include <stdint.h>

#ifdef __CUDA_ARCH__
#define FX_MUL(a, b) __dmul_rn((a), (b))
#define FX_ADD(a, b) __dadd_rn((a), (b))
#define FX_FMUL(a, b) __fmul_rn((a), (b))
#define FX_FADD(a, b) __fadd_rn((a), (b))
#define FX_FSUB(a, b) __fsub_rn((a), (b))
#define FX_FDIV(a, b) __fdiv_rn((a), (b))
#else
#define FX_MUL(a, b) ((a) * (b))
#define FX_ADD(a, b) ((a) + (b))
#define FX_FMUL(a, b) ((a) * (b))
#define FX_FADD(a, b) ((a) + (b))
#define FX_FSUB(a, b) ((a) - (b))
#define FX_FDIV(a, b) ((a) / (b))
#endif

#ifndef __CUDACC__
#ifndef __host__
#define __host__
#define __device__
#endif
#endif

namespace orbx {

__host__ __device__ inline float fast_atan2_deg(float y, float x) {
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;  // constant-folded in float, one rounding each (as the reference build does)
    const float eps = (float)2.2204460492503131e-16;  // (float)DBL_EPSILON
    const float ax = x < 0 ? -x : x, ay = y < 0 ? -y : y;
    float a, c, c2;
    if (ax >= ay) {
        c = FX_FDIV(ay, FX_FADD(ax, eps));
        c2 = FX_FMUL(c, c);
        a = FX_FMUL(FX_FADD(FX_FMUL(FX_FADD(FX_FMUL(FX_FADD(FX_FMUL(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = FX_FDIV(ax, FX_FADD(ay, eps));
        c2 = FX_FMUL(c, c);
        a = FX_FSUB(90.f, FX_FMUL(FX_FADD(FX_FMUL(FX_FADD(FX_FMUL(FX_FADD(FX_FMUL(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = FX_FSUB(180.f, a);
    if (y < 0) a = FX_FSUB(360.f, a);
    return a;
}

namespace detail {
// polynomial coefficients of the published sincosf tables (cosine c0..c4, sine s1..s3)
__host__ __device__ inline double sc_poly(double x, double x2, int n, bool neg) {
    const double c0 = neg ? -0x1p0 : 0x1p0, c1 = neg ? 0x1.ffffffd0c621cp-2 : -0x1.ffffffd0c621cp-2,
                 c2 = neg ? -0x1.55553e1068f19p-5 : 0x1.55553e1068f19p-5, c3 = neg ? 0x1.6c087e89a359dp-10 : -0x1.6c087e89a359dp-10,
                 c4 = neg ? -0x1.99343027bf8c3p-16 : 0x1.99343027bf8c3p-16;
    const double s1 = -0x1.555545995a603p-3, s2 = 0x1.1107605230bc4p-7, s3 = -0x1.994eb3774cf24p-13;
    if ((n & 1) == 0) {
        const double x3 = FX_MUL(x, x2);
        const double t1 = FX_ADD(s2, FX_MUL(x2, s3));
        const double x7 = FX_MUL(x3, x2);
        const double s = FX_ADD(x, FX_MUL(x3, s1));
        return FX_ADD(s, FX_MUL(x7, t1));
    } else {
        const double x4 = FX_MUL(x2, x2);
        const double t2 = FX_ADD(c3, FX_MUL(x2, c4));
        const double t1 = FX_ADD(c0, FX_MUL(x2, c1));
        const double x6 = FX_MUL(x4, x2);
        const double c = FX_ADD(t1, FX_MUL(x4, c2));
        return FX_ADD(c, FX_MUL(x6, t2));
    }
}
__host__ __device__ inline uint32_t abstop12(float v) {
    union { float f; uint32_t u; } q;
    q.f = v;
    return (q.u >> 20) & 0x7ff;
}
}  // namespace detail

// sinf(y) and cosf(y) as glibc computes them, 0 <= y < 120
__host__ __device__ inline void sinf_cosf_glibc(float y, float* sinOut, float* cosOut) {
    double x = (double)y;
    if (detail::abstop12(y) < detail::abstop12(0x1.921FB6p-1f)) {  // |y| < pi/4 on the top 12 bits
        const double x2 = FX_MUL(x, x);
        if (detail::abstop12(y) < detail::abstop12(0x1p-12f)) {
            *sinOut = y;
            *cosOut = 1.0f;
            return;
        }
        *sinOut = (float)detail::sc_poly(x, x2, 0, false);
        *cosOut = (float)detail::sc_poly(x, x2, 1, false);
        return;
    }
    const double hpi_inv = 0x1.45F306DC9C883p+23, hpi = 0x1.921FB54442D18p0;
    const double r = FX_MUL(x, hpi_inv);
    const int n = ((int32_t)r + 0x800000) >> 24;
    x = FX_ADD(x, -FX_MUL((double)n, hpi));
    const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;  // sign table {1,-1,-1,1}
    const bool neg = (n & 2) != 0;
    const double xs = FX_MUL(x, sgn), x2 = FX_MUL(x, x);
    *sinOut = (float)detail::sc_poly(xs, x2, n, neg);
    *cosOut = (float)detail::sc_poly(xs, x2, n ^ 1, neg);
}

}  // namespace orbx
