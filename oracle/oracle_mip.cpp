// oracle_mip.cpp -- CPU restatement of maketx's per-level box downsample.
// TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Follows /root/reference/src/libOpenImageIO/maketexture.cpp:
//   interppixel_NDC      :141-198  (non-envlatl branch)
//   resize_block_        :205-241
//   halve_scanline       :245-254
//   resize_block_2pass   :260-300
//   resize_block         :304-334  (which of the two runs)
// The level-size rule of write_mipmap (:848-857) and the filter rule of
// setup_filter (:38-66, identical to get_resize_filter with filterwidth 0)
// are exercised from Python by calling orc_resize per level.
// Compile with -ffp-contract=off.
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <vector>

namespace {

inline float
floorfrac(float x, int* xint)
{
    float f = std::floor(x);
    *xint   = int(f);
    return x - f;
}
inline int
clampi(int a, int lo, int hi)
{
    return a < lo ? lo : (a > hi ? hi : a);
}

const float*
src_px(const orc_image* im, int x, int y, bool wrap_black)
{
    bool inside = x >= im->x && x < im->x + im->width && y >= im->y
                  && y < im->y + im->height;
    if (!inside) {
        if (wrap_black)
            return nullptr;
        x      = clampi(x, im->full_x, im->full_x + im->full_width - 1);
        y      = clampi(y, im->full_y, im->full_y + im->full_height - 1);
        inside = x >= im->x && x < im->x + im->width && y >= im->y
                 && y < im->y + im->height;
        if (!inside)
            return nullptr;
    }
    return (const float*)((const char*)im->pixels + int64_t(y - im->y) * im->ystride
                          + int64_t(x - im->x) * im->xstride);
}

void
block_bilinear(const orc_image* dst, const orc_image* src)
{
    bool src_is_crop = (src->x > src->full_x || src->y > src->full_y
                        || src->x + src->width < src->full_x + src->full_width
                        || src->y + src->height < src->full_y + src->full_height);
    float xoffset = (float)dst->full_x;
    float yoffset = (float)dst->full_y;
    float xscale  = 1.0f / (float)dst->full_width;
    float yscale  = 1.0f / (float)dst->full_height;
    int nch       = dst->nchannels;
    for (int y = dst->y; y < dst->y + dst->height; ++y) {
        float t = (y + 0.5f) * yscale + yoffset;
        for (int x = dst->x; x < dst->x + dst->width; ++x) {
            float s  = (x + 0.5f) * xscale + xoffset;
            float fx = static_cast<float>(src->full_x)
                       + s * static_cast<float>(src->full_width);
            float fy = static_cast<float>(src->full_y)
                       + t * static_cast<float>(src->full_height);
            fx -= 0.5f;
            fy -= 0.5f;
            int xt, yt;
            float xfrac = floorfrac(fx, &xt);
            float yfrac = floorfrac(fy, &yt);
            const float* p[4];
            for (int k = 0; k < 4; ++k)
                p[k] = src_px(src, xt + (k & 1), yt + (k >> 1), src_is_crop);
            float s1 = 1.0f - xfrac, t1 = 1.0f - yfrac;
            float* d = (float*)((char*)dst->pixels
                                + int64_t(y - dst->y) * dst->ystride
                                + int64_t(x - dst->x) * dst->xstride);
            for (int c = 0; c < nch; ++c) {
                float v0 = p[0] ? p[0][c] : 0.0f, v1 = p[1] ? p[1][c] : 0.0f;
                float v2 = p[2] ? p[2][c] : 0.0f, v3 = p[3] ? p[3][c] : 0.0f;
                d[c] = t1 * (v0 * s1 + v1 * xfrac) + yfrac * (v2 * s1 + v3 * xfrac);
            }
        }
    }
}

void
block_2pass(const orc_image* dst, const orc_image* src)
{
    const int nch  = dst->nchannels;
    const size_t dw = dst->width, dh = dst->height;
    std::vector<float> S0(dw * nch), S1(dw * nch);
    for (size_t y = 0; y < dh; ++y) {
        const float* r0 = (const float*)((const char*)src->pixels
                                         + int64_t(2 * y) * src->ystride);
        const float* r1 = (const float*)((const char*)src->pixels
                                         + int64_t(2 * y + 1) * src->ystride);
        for (size_t x = 0; x < dw; ++x)
            for (int c = 0; c < nch; ++c) {
                S0[x * nch + c] = 0.5f
                                  * (float)(r0[(2 * x) * nch + c]
                                            + r0[(2 * x + 1) * nch + c]);
                S1[x * nch + c] = 0.5f
                                  * (float)(r1[(2 * x) * nch + c]
                                            + r1[(2 * x + 1) * nch + c]);
            }
        float* d = (float*)((char*)dst->pixels + int64_t(y) * dst->ystride);
        for (size_t i = 0; i < dw * nch; ++i)
            d[i] = 0.5f * (S0[i] + S1[i]);
    }
}

}  // namespace

extern "C" int
orc_resize_block(const orc_image* dst, const orc_image* src, int envlatlmode,
                 int allow_shift, int /*nthreads*/)
{
    if (dst->basetype != ORC_FLOAT || src->basetype != ORC_FLOAT
        || dst->nchannels != src->nchannels || envlatlmode)
        return 1;
    bool contiguous = src->xstride == int64_t(4) * src->nchannels
                      && dst->xstride == int64_t(4) * dst->nchannels;
    bool fast = contiguous && (2 * dst->width) == src->width && dst->x == 0
                && dst->y == 0 && src->x == 0 && src->y == 0;
    if (fast && !allow_shift && (src->width % 2 || src->height % 2))
        fast = false;
    if (fast)
        block_2pass(dst, src);
    else
        block_bilinear(dst, src);
    return 0;
}
