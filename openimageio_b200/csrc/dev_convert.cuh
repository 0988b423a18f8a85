// dev_convert.cuh -- device-side pixel type conversion shared by the kernels.
#pragma once

#include "kernels.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace b2r {

// ------------------------------------------------------------------------
// type conversion: convert_type<S,float> / convert_type<float,D>
// (src/include/OpenImageIO/fmath.h:753-764, 1009-1031).  Explicit _rn
// intrinsics keep nvcc from contracting v*max+0.5 into one FMA, which would
// round differently from the reference's multiply-then-add.
// ------------------------------------------------------------------------
__device__ __forceinline__ float
load_elem(const unsigned char* p, int bt)
{
    switch (bt) {
    case BT_UINT8: return __fmul_rn(float(*p), 1.0f / 255.0f);
    case BT_UINT16:
        return __fmul_rn(float(*(const unsigned short*)p), 1.0f / 65535.0f);
    case BT_HALF: return __half2float(*(const __half*)p);
    default: return *(const float*)p;
    }
}

__device__ __forceinline__ unsigned int
quant_int(float v, float mx)
{
    float s = __fmul_rn(v, mx);
    s       = __fadd_rn(s, (s < 0.0f) ? -0.5f : 0.5f);
    if (!(0.0f <= s))  // also catches NaN -> 0
        s = 0.0f;
    if (s > mx)
        s = mx;
    return (unsigned int)s;  // truncation
}

__device__ __forceinline__ void
store_elem(unsigned char* p, int bt, float v)
{
    switch (bt) {
    case BT_UINT8: *p = (unsigned char)quant_int(v, 255.0f); break;
    case BT_UINT16: *(unsigned short*)p = (unsigned short)quant_int(v, 65535.0f); break;
    case BT_HALF: *(__half*)p = __float2half_rn(v); break;
    default: *(float*)p = v; break;
    }
}

// convert_type<float,uint>(v) (fmath.h:753-764, 1009-1031) for v*mx computed
// with a separate multiply: s = v*mx; s += 0.5 (s >= 0) ; clamp ; truncate.
// For s < 0 the reference adds -0.5 and clamps to 0: s + 0.5 is below 0.5
// there and truncates (or saturates) to 0 as well, so one FADD serves both
// signs.  A float -> u8/u16 cvt clamps to the destination range by itself (and
// sends NaN to 0), which is the reference's clamp: no min/max instructions.
__device__ __forceinline__ unsigned
quant_u8(float v)
{
    float s = __fadd_rn(__fmul_rn(v, 255.0f), 0.5f);
    unsigned u;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(u) : "f"(s));
    return u;
}
__device__ __forceinline__ unsigned
quant_u16(float v)
{
    float s = __fadd_rn(__fmul_rn(v, 65535.0f), 0.5f);
    unsigned u;
    asm("cvt.rzi.u16.f32 %0, %1;" : "=r"(u) : "f"(s));
    return u;
}

// The same quantiser for two values with the multiply and the add contracted
// into one packed FMA (expand kernel: the quantiser was a third of its
// instructions).  One rounding instead of two; see the call site.
__device__ __forceinline__ float2
quant2_fma(float a, float b, float mx)
{
    return __ffma2_rn(make_float2(a, b), make_float2(mx, mx), make_float2(0.5f, 0.5f));
}
__device__ __forceinline__ unsigned
cvt_u8(float s)
{
    unsigned u;
    asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(u) : "f"(s));
    return u;
}
__device__ __forceinline__ unsigned
cvt_u16(float s)
{
    unsigned u;
    asm("cvt.rzi.u16.f32 %0, %1;" : "=r"(u) : "f"(s));
    return u;
}

// rangecompress / rangeexpand of one value (imagebufalgo_pixelmath.cpp:560-735)
__device__ __forceinline__ float
dev_rangecompress(float x)
{
    const float x1 = 0.18f, a = -0.54576885700225830078f;
    const float b = 0.18351669609546661377f, c = 284.3577880859375f;
    float absx = fabsf(x);
    if (absx <= x1)
        return x;
    float t = __fadd_rn(__fmul_rn(c, absx), 1.0f);
    return copysignf(__fadd_rn(a, __fmul_rn(b, logf(fabsf(t)))), x);
}

__device__ __forceinline__ float
dev_rangeexpand(float y)
{
    const float x1 = 0.18f, a = -0.54576885700225830078f;
    const float b = 0.18351669609546661377f, c = 284.3577880859375f;
    float absy = fabsf(y);
    if (absy <= x1)
        return y;
    float xi = expf(__fdiv_rn(__fsub_rn(absy, a), b));
    float x  = __fdiv_rn(__fsub_rn(xi, 1.0f), c);
    if (x < x1)
        x = __fdiv_rn(__fsub_rn(-xi, 1.0f), c);
    return copysignf(x, y);
}

__device__ __forceinline__ bool
nonfinite(float v)
{
    return (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u;
}

}  // namespace b2r
