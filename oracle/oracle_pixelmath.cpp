// oracle_pixelmath.cpp -- CPU restatement of ImageBufAlgo::rangecompress /
// rangeexpand (src/libOpenImageIO/imagebufalgo_pixelmath.cpp:560-735).
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned against
// testsuite/oiiotool/ref/{rangecompress,rangeexpand}{,-luma}.tif
// (tests/golden/pixelmath_ref.npz, tests/test_cpu.py).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace {

// pixelmath.cpp:560-575
inline float
rangecompress(float x)
{
    const float x1 = 0.18f, a = -0.54576885700225830078f;
    const float b = 0.18351669609546661377f, c = 284.3577880859375f;
    float absx = fabsf(x);
    if (absx <= x1)
        return x;
    return copysignf(a + b * logf(fabsf(c * absx + 1.0f)), x);
}

// pixelmath.cpp:580-603
inline float
rangeexpand(float y)
{
    const float x1 = 0.18f, a = -0.54576885700225830078f;
    const float b = 0.18351669609546661377f, c = 284.3577880859375f;
    float absy = fabsf(y);
    if (absy <= x1)
        return y;
    float xIntermediate = expf((absy - a) / b);
    float x             = (xIntermediate - 1.0f) / c;
    if (x < x1)
        x = (-xIntermediate - 1.0f) / c;
    return copysignf(x, y);
}

int
esize(int bt)
{
    return bt == ORC_UINT8 ? 1 : (bt == ORC_FLOAT ? 4 : 2);
}

// rangecompress_ / rangeexpand_ (:607-735); both branches (in place or not)
// give the same pixels, the copy of the alpha / z channels aside.
int
range_op(bool expand, const orc_image* dst, const orc_image* src, int useluma,
         int alpha_channel, int z_channel, orc_roi roi, int chbegin, int chend)
{
    if (!dst || !src || !dst->pixels || !src->pixels)
        return 1;
    chend = std::min(chend, std::min(dst->nchannels, src->nchannels));
    if (chend - chbegin < 3 || (alpha_channel >= chbegin && alpha_channel < chbegin + 3)
        || (z_channel >= chbegin && z_channel < chbegin + 3))
        useluma = 0;
    const bool in_place = dst->pixels == src->pixels;
    const int ses = esize(src->basetype), des = esize(dst->basetype);
    for (int y = roi.ybegin; y < roi.yend; ++y) {
        for (int x = roi.xbegin; x < roi.xend; ++x) {
            if (x < dst->x || x >= dst->x + dst->width || y < dst->y
                || y >= dst->y + dst->height)
                continue;
            unsigned char* dp = (unsigned char*)dst->pixels + (y - dst->y) * dst->ystride
                                + (x - dst->x) * dst->xstride;
            const bool inside = x >= src->x && x < src->x + src->width && y >= src->y
                                && y < src->y + src->height;
            const unsigned char* sp = (const unsigned char*)src->pixels
                                      + (y - src->y) * src->ystride
                                      + (x - src->x) * src->xstride;
            auto a = [&](int c) {
                return inside ? orc_load_as_float(sp + c * ses, src->basetype) : 0.0f;
            };
            float scale = 0.0f;
            if (useluma) {
                float luma = 0.21264f * a(chbegin) + 0.71517f * a(chbegin + 1)
                             + 0.07219f * a(chbegin + 2);
                scale = luma > 0.0f
                            ? (expand ? rangeexpand(luma) : rangecompress(luma)) / luma
                            : 0.0f;
            }
            for (int c = chbegin; c < chend; ++c) {
                float v = a(c);
                if (c == alpha_channel || c == z_channel) {
                    if (in_place)
                        continue;
                } else if (useluma) {
                    v = v * scale;
                } else {
                    v = expand ? rangeexpand(v) : rangecompress(v);
                }
                orc_store_from_float(dp + c * des, dst->basetype, v);
            }
        }
    }
    return 0;
}

}  // namespace

extern "C" int
orc_rangecompress(const orc_image* dst, const orc_image* src, int useluma,
                  int alpha_channel, int z_channel, orc_roi roi, int chbegin, int chend)
{
    return range_op(false, dst, src, useluma, alpha_channel, z_channel, roi, chbegin, chend);
}

extern "C" int
orc_rangeexpand(const orc_image* dst, const orc_image* src, int useluma, int alpha_channel,
                int z_channel, orc_roi roi, int chbegin, int chend)
{
    return range_op(true, dst, src, useluma, alpha_channel, z_channel, roi, chbegin, chend);
}

// ---------------------------------------------------------------------------
// ImageBufAlgo::fixNonFinite (imagebufalgo_pixelmath.cpp:1424-1610) and maketx's
// check_nan_block (maketexture.cpp:340-360, = mode NONE's count).
// mode: 0 none (count only), 1 black, 2 box3, 100 error (count only).
// Float and half images; the other types cannot hold non-finite values and are
// only copied.  Raster order, one band: box3 repairs in place, so a pixel
// repaired earlier in the scan is a finite neighbour of the next one (the
// reference's result depends on its thread bands the same way).
// ---------------------------------------------------------------------------
namespace {
inline float
round_to_half(float v)
{
    unsigned short h;
    orc_store_from_float(&h, ORC_HALF, v);
    return orc_load_as_float(&h, ORC_HALF);
}
}  // namespace

extern "C" int
orc_fix_nonfinite(const orc_image* dst, const orc_image* src, int mode, orc_roi roi, int chbegin,
                  int chend, int* pixels_fixed)
{
    if (!dst || !src || dst->basetype != src->basetype || dst->nchannels != src->nchannels)
        return 1;
    const int es = esize(dst->basetype), nc = dst->nchannels;
    chend = std::min(chend, nc);
    auto dpx = [&](int x, int y) {
        return (unsigned char*)dst->pixels + int64_t(y - dst->y) * dst->ystride
               + int64_t(x - dst->x) * dst->xstride;
    };
    // copy(dst, src, roi)
    if (dst->pixels != src->pixels)
        for (int y = roi.ybegin; y < roi.yend; ++y)
            for (int x = roi.xbegin; x < roi.xend; ++x) {
                if (x < dst->x || x >= dst->x + dst->width || y < dst->y
                    || y >= dst->y + dst->height)
                    continue;
                const bool in = x >= src->x && x < src->x + src->width && y >= src->y
                                && y < src->y + src->height;
                unsigned char* d = dpx(x, y);
                if (in)
                    memcpy(d,
                           (const unsigned char*)src->pixels + int64_t(y - src->y) * src->ystride
                               + int64_t(x - src->x) * src->xstride,
                           (size_t)nc * es);
                else
                    memset(d, 0, (size_t)nc * es);
            }
    int count = 0;
    const bool is_half = dst->basetype == ORC_HALF;
    if (dst->basetype == ORC_FLOAT || is_half)
        for (int y = std::max(roi.ybegin, dst->y); y < std::min(roi.yend, dst->y + dst->height);
             ++y)
            for (int x = std::max(roi.xbegin, dst->x);
                 x < std::min(roi.xend, dst->x + dst->width); ++x) {
                unsigned char* d = dpx(x, y);
                bool fixed       = false;
                for (int c = chbegin; c < chend; ++c) {
                    const float value = orc_load_as_float(d + c * es, dst->basetype);
                    if (std::isfinite(value))
                        continue;
                    fixed = true;
                    if (mode == 1) {
                        orc_store_from_float(d + c * es, dst->basetype, 0.0f);
                    } else if (mode == 2) {
                        int numvals = 0;
                        float sum   = 0.0f;
                        for (int j = std::max(y - 1, dst->y);
                             j < std::min(y + 2, dst->y + dst->height); ++j)
                            for (int i = std::max(x - 1, dst->x);
                                 i < std::min(x + 2, dst->x + dst->width); ++i) {
                                const float v = orc_load_as_float(dpx(i, j) + c * es,
                                                                  dst->basetype);
                                if (std::isfinite(v)) {
                                    sum = is_half ? round_to_half(sum + v) : sum + v;
                                    ++numvals;
                                }
                            }
                        orc_store_from_float(d + c * es, dst->basetype,
                                             numvals ? sum / numvals : 0.0f);
                    } else {
                        break;  // counting modes: one per pixel
                    }
                }
                if (fixed)
                    ++count;
            }
    if (pixels_fixed)
        *pixels_fixed = count;
    return 0;
}
