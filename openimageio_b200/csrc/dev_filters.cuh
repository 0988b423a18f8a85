// dev_filters.cuh -- device-side evaluation of Filter2D::operator()(x, y).
//
// The resize kernels never evaluate a filter (their weights are host tables);
// warp does, per tap, because every destination pixel lands on its own
// sub-pixel position.  Same profiles, same expression order as
// recon_filter.cpp (= src/libutil/filter.cpp), with explicit round-to-nearest
// multiplies and adds so nvcc cannot contract them; sinf / cosf are the
// device's (<= 2 ulp from glibc's).
#pragma once

#include "kernels.h"

namespace b2r {

#define B2R_MUL(a, b) __fmul_rn((a), (b))
#define B2R_ADD(a, b) __fadd_rn((a), (b))
#define B2R_SUB(a, b) __fsub_rn((a), (b))

// fast_exp2 / fast_exp (fmath.h): polynomial, not expf
__device__ __forceinline__ float
dev_exp_poly(float xin)
{
    float x = B2R_MUL(xin, float(1.0 / 0.693147180559945309417));
    if (!(-126.0f <= x))
        x = -126.0f;
    if (x > 126.0f)
        x = 126.0f;
    int m = int(x);
    x     = B2R_SUB(x, float(m));
    x     = B2R_SUB(1.0f, B2R_SUB(1.0f, x));
    float r = 1.33336498402e-3f;
    r       = B2R_ADD(B2R_MUL(x, r), 9.810352697968e-3f);
    r       = B2R_ADD(B2R_MUL(x, r), 5.551834031939e-2f);
    r       = B2R_ADD(B2R_MUL(x, r), 0.2401793301105f);
    r       = B2R_ADD(B2R_MUL(x, r), 0.693144857883f);
    r       = B2R_ADD(B2R_MUL(x, r), 1.0f);
    return __uint_as_float(__float_as_uint(r) + (unsigned(m) << 23));
}

__device__ __forceinline__ float
dev_profile(int profile, float x, float param)
{
    const float pi = 3.14159265358979323846f;
    switch (profile) {
    case 0: return 1.0f;
    case 1: x = fabsf(x); return (x < 1.0f) ? B2R_SUB(1.0f, x) : 0.0f;
    case 2:
        x = fabsf(x);
        return (x < 1.0f) ? dev_exp_poly(B2R_MUL(param, B2R_MUL(x, x))) : 0.0f;
    case 3: {
        x        = fabsf(x);
        float x2 = B2R_MUL(x, x), x3 = B2R_MUL(x, x2);
        if (x >= 2.0f)
            return 0.0f;
        return (x < 1.0f)
                   ? B2R_ADD(B2R_SUB(B2R_MUL(3.0f, x3), B2R_MUL(5.0f, x2)), 2.0f)
                   : B2R_ADD(B2R_SUB(B2R_ADD(-x3, B2R_MUL(5.0f, x2)), B2R_MUL(8.0f, x)), 4.0f);
    }
    case 4: {
        if (x < -1.0f || x > 1.0f)
            return 0.0f;
        x        = B2R_MUL(B2R_ADD(x, 1.0f), 0.5f);
        float c1 = cosf(B2R_MUL(B2R_MUL(2.f, pi), x));
        float c2 = B2R_SUB(B2R_MUL(B2R_MUL(2.0f, c1), c1), 1.0f);
        float c3 = B2R_MUL(c1, B2R_SUB(B2R_MUL(2.0f, c2), 1.0f));
        return B2R_ADD(B2R_ADD(B2R_ADD(0.35875f, B2R_MUL(-0.48829f, c1)), B2R_MUL(0.14128f, c2)),
                       B2R_MUL(-0.01168f, c3));
    }
    case 5: {
        x = fabsf(x);
        if (x > param)
            return 0.0f;
        return (x < 0.0001f) ? 1.0f : __fdiv_rn(sinf(B2R_MUL(pi, x)), B2R_MUL(pi, x));
    }
    case 6: {
        x = fabsf(x);
        if (x > 3.0f)
            return 0.0f;
        if (x < 0.0001f)
            return 1.0f;
        float s1 = sinf(B2R_MUL(B2R_MUL(x, 1.0f / 3.0f), pi));
        float s3 = B2R_MUL(B2R_ADD(B2R_MUL(B2R_MUL(-4.0f, s1), s1), 3.0f), s1);
        return B2R_MUL(B2R_MUL(__fdiv_rn(3.0f, B2R_MUL(B2R_MUL(x, x), B2R_MUL(pi, pi))), s1), s3);
    }
    case 7: {
        x = fabsf(x);
        if (x > 1.0f)
            return 0.0f;
        x                 = B2R_MUL(x, 2.0f);
        float x2          = B2R_MUL(x, x);
        const float B     = 1.0f / 3.0f;
        const float C     = 1.0f / 3.0f;
        const float SIXTH = 1.0f / 6.0f;
        // the coefficient expressions are compile-time constants in the
        // reference too (float arithmetic on literals)
        const float a3 = -B - 6.0f * C, a2 = 6.0f * B + 30.0f * C, a1 = -12.0f * B - 48.0f * C,
                    a0 = 8.0f * B + 24.0f * C;
        const float b3 = 12.0f - 9.0f * B - 6.0f * C, b2 = -18.0f + 12.0f * B + 6.0f * C,
                    b0 = 6.0f - 2.0f * B;
        if (x >= 1.0f)
            return B2R_MUL(B2R_ADD(B2R_ADD(B2R_ADD(B2R_MUL(B2R_MUL(a3, x), x2), B2R_MUL(a2, x2)),
                                           B2R_MUL(a1, x)),
                                   a0),
                           SIXTH);
        return B2R_MUL(B2R_ADD(B2R_ADD(B2R_MUL(B2R_MUL(b3, x), x2), B2R_MUL(b2, x2)), b0), SIXTH);
    }
    case 8: {
        x = fabsf(x);
        if (x <= 1.0f) {
            float t = B2R_SUB(1.0f, x);
            return B2R_ADD(B2R_MUL(B2R_MUL(0.5f, t), B2R_ADD(B2R_MUL(t, B2R_SUB(1.0f, t)), 1.0f)),
                           1.0f / 6.0f);
        }
        if (x < 2.0f) {
            float t = B2R_SUB(2.0f, x);
            return __fdiv_rn(B2R_MUL(B2R_MUL(t, t), t), 6.0f);
        }
        return 0.0f;
    }
    default: {  // 9: cubic family, param = a
        const float a = param;
        x             = fabsf(x);
        if (x > 1.0f)
            return 0.0f;
        x = B2R_MUL(x, 2.0f);
        if (x >= 1.0f)
            return B2R_MUL(a, B2R_SUB(B2R_MUL(x, B2R_ADD(B2R_MUL(x, B2R_SUB(x, 5.0f)), 8.0f)),
                                      4.0f));
        return B2R_ADD(B2R_MUL(B2R_MUL(x, x),
                               B2R_SUB(B2R_MUL(B2R_ADD(a, 2.0f), x), B2R_ADD(a, 3.0f))),
                       1.0f);
    }
    }
}

struct DevFilter {
    int profile, scale;
    float param, w, h;
};

// One axis of a separable filter with everything that does not depend on the
// argument computed once (the argument scale is an IEEE division).
struct DevAxis {
    int profile, mode;  // mode 0: box edge, 1: scaled argument, 2: sinc (param = radius)
    float k, param;
};

__device__ __forceinline__ DevAxis
dev_make_axis(const DevFilter& f, float ext)
{
    DevAxis a;
    a.profile = f.profile;
    a.param   = f.param;
    a.mode    = 1;
    a.k       = 0.0f;
    switch (f.scale) {
    case 0: a.mode = 0, a.k = B2R_MUL(ext, 0.5f); break;
    case 1: a.k = __fdiv_rn(2.0f, ext); break;
    case 2: a.k = __fdiv_rn(4.0f, ext); break;
    case 3:
    case 6: a.k = __fdiv_rn(6.0f, ext); break;
    case 4: a.mode = 2, a.param = __fdiv_rn(ext, 2.0f); break;
    default: a.mode = 3; break;
    }
    return a;
}

__device__ __forceinline__ float
dev_axis_eval(const DevAxis& a, float x)
{
    if (a.mode == 1)
        return dev_profile(a.profile, B2R_MUL(x, a.k), a.param);
    if (a.mode == 0)
        return fabsf(x) <= a.k ? 1.0f : 0.0f;
    if (a.mode == 2)
        return dev_profile(a.profile, x, a.param);
    return 0.0f;
}

__device__ __forceinline__ float
dev_axis_value(const DevFilter& f, float ext, float x)
{
    return dev_axis_eval(dev_make_axis(f, ext), x);
}

// Filter2D::operator()(x, y)
__device__ __forceinline__ float
dev_filter2d(const DevFilter& f, float x, float y)
{
    if (f.scale == 5) {  // disk
        x = __fdiv_rn(x, B2R_MUL(f.w, 0.5f));
        y = __fdiv_rn(y, B2R_MUL(f.h, 0.5f));
        return (B2R_ADD(B2R_MUL(x, x), B2R_MUL(y, y)) < 1.0f) ? 1.0f : 0.0f;
    }
    if (f.scale == 6) {  // radial lanczos3
        x = B2R_MUL(x, __fdiv_rn(6.0f, f.w));
        y = B2R_MUL(y, __fdiv_rn(6.0f, f.h));
        return dev_profile(6, __fsqrt_rn(B2R_ADD(B2R_MUL(x, x), B2R_MUL(y, y))), 0.0f);
    }
    return B2R_MUL(dev_axis_value(f, f.w, x), dev_axis_value(f, f.h, y));
}

}  // namespace b2r
