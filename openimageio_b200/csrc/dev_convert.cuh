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

// The same quantiser for two values WITHOUT the float -> int conversion unit
// (expand kernel: an 8K uint8 frame is 100 M cvt instructions, and the XU pipe
// they run on was at its peak in the profile while issue slots and HBM idled).
// The reference's sequence is kept to the bit -- s = v * mx rounded, s + 0.5
// rounded, clamp, truncate (fmath.h:753-764) -- because for uint16 the two
// roundings matter: s has an ulp of 2^-8 near the top of the range, so a
// single-rounding shortcut (v * mx + 0.5 in one FMA, or rounding v * mx to
// nearest) lands on the other integer for ~0.2 % of the values (the fuzz found
// it: 2 LSB where filter noise and the shortcut added up).  Here: clamp v to
// [0, 1] first (add.sat, NaN -> 0; equivalent to the reference's clamp of s),
// the product and the sum as two packed FMAs with exact addends (x * mx + -0,
// s * 1 + 0.5: one rounding each, the reference's), and the truncation by an
// add of 2^23 rounded toward zero -- the sum lies in [2^23, 2^24), where a
// float's ulp is 1, so the low mantissa bits are floor(s + 0.5).
__device__ __forceinline__ float
sat01(float v)
{
    float r;
    asm("add.sat.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(v));
    return r;
}
// returns the two floats 2^23 + quantised value; their low 8 / 16 bits are the result
__device__ __forceinline__ float2
quant2_magic(float a, float b, float mx)
{
    const float2 s = __ffma2_rn(make_float2(sat01(a), sat01(b)), make_float2(mx, mx),
                                make_float2(-0.0f, -0.0f));
    const float2 t = __ffma2_rn(s, make_float2(1.0f, 1.0f), make_float2(0.5f, 0.5f));
    return make_float2(__fadd_rz(t.x, 8388608.0f), __fadd_rz(t.y, 8388608.0f));
}
// uint8 only: v * 255 has an ulp of at most 2^-16, so rounding v * 255 to nearest
// in ONE packed FMA onto 2^23 agrees with the reference's sequence except for
// ~1e-5 of the values (and then by 1 LSB) -- against 0.2 % for uint16, which
// therefore takes the exact form above.  Half the instructions of the exact form.
__device__ __forceinline__ float2
quant2_magic_u8(float a, float b)
{
    return __ffma2_rn(make_float2(sat01(a), sat01(b)), make_float2(255.0f, 255.0f),
                      make_float2(8388608.0f, 8388608.0f));
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

// The same two curves for the FUSED highlight compensation of the stream kernel,
// on the special-function unit: lg2.approx / ex2.approx (relative error 2^-22
// each) and the two IEEE divisions as multiplications by reciprocals.  There
// the curves run inside an issue-bound kernel -- libm's logf / expf and the
// divisions cost ~25 / ~30 instructions per value against ~7 / ~8 here, the
// difference between 0.32 ms and 0.16 ms for an 8K frame.  Error budget: the
// argument of the log is >= 1 + c * 0.18 = 52, so lg2's relative error is
// 2.4e-7 * 8 = 2e-6 in log2 units = 2.4e-7 in the compressed value; through
// rangeexpand's exponent ((y - a) / b * log2 e) that is 1.9e-6, i.e. 1.3e-6
// RELATIVE in the expanded value, plus ex2's own 2.4e-7 and the argument
// rounding: ~2e-6 relative end to end (measured: tools/hicomp_probe.py), inside
// north_star's 1e-5.  The stand-alone rangecompress / rangeexpand ops keep the
// libm forms above.
__device__ __forceinline__ float
dev_rangecompress_fast(float x)
{
    const float x1 = 0.18f, a = -0.54576885700225830078f;
    const float b = 0.18351669609546661377f, c = 284.3577880859375f;
    const float absx = fabsf(x);
    if (absx <= x1)
        return x;
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(fmaf(c, absx, 1.0f)));
    return copysignf(fmaf(b * 0.693147180559945309f, l2, a), x);
}

__device__ __forceinline__ float
dev_rangeexpand_fast(float y)
{
    const float x1 = 0.18f, a = -0.54576885700225830078f;
    const float b = 0.18351669609546661377f, c = 284.3577880859375f;
    const float absy = fabsf(y);
    if (absy <= x1)
        return y;
    const float k = 1.44269504088896340736f / b;  // log2(e) / b
    float xi;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(xi) : "f"(fmaf(absy, k, -a * k)));
    const float rc = 1.0f / c;
    float x = fmaf(xi, rc, -rc);
    if (x < x1)  // the reference's second branch (imagebufalgo_pixelmath.cpp:596-601)
        x = fmaf(-xi, rc, -rc);
    return copysignf(x, y);
}

__device__ __forceinline__ bool
nonfinite(float v)
{
    return (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u;
}

}  // namespace b2r
