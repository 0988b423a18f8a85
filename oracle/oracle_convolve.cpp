// oracle_convolve.cpp -- CPU restatement of ImageBufAlgo::make_kernel,
// convolve and unsharp_mask (src/libOpenImageIO/imagebufalgo.cpp:772-1031),
// the sharpening step maketx wraps around each MIP level's resize
// (maketexture.cpp:909-931) and oiiotool --kernel/--convolve/--blur/--unsharp.
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  Pinned against
// testsuite/oiiotool/ref/{bspline-blur,gauss5x5-blur,unsharp}.tif
// (tests/golden/conv_ref.npz, tests/test_cpu.py).  The "median" kernel of
// unsharp_mask is not restated (not on the path).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <strings.h>
#include <vector>

struct OrcFilter2D {
    int kind;
    float w, h;
};
bool
orc_make_filter2d(const char* name, float w, float h, OrcFilter2D* f);
float
orc_f2d_eval(const OrcFilter2D& f, float x, float y);
void
orc_parallel_bands(orc_roi roi, int nthreads, const std::function<void(orc_roi)>& f);

namespace {

inline int
esize(int bt)
{
    return bt == ORC_UINT8 ? 1 : (bt == ORC_FLOAT ? 4 : 2);
}
inline int
clampi(int v, int lo, int hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}
// imagebufalgo.cpp:855-861
inline float
binomial(int n, int k)
{
    float p = 1;
    for (int i = 1; i <= k; ++i)
        p *= float(n - (k - i)) / i;
    return p;
}

// ConstIterator with WrapClamp: clamp to the display window, black if that
// still misses the data window (imagebuf.cpp:3283-3322)
const unsigned char*
clamped_pixel(const orc_image& im, int x, int y)
{
    bool inside = x >= im.x && x < im.x + im.width && y >= im.y && y < im.y + im.height;
    if (!inside) {
        x = clampi(x, im.full_x, im.full_x + im.full_width - 1);
        y = clampi(y, im.full_y, im.full_y + im.full_height - 1);
        inside = x >= im.x && x < im.x + im.width && y >= im.y && y < im.y + im.height;
        if (!inside)
            return nullptr;
    }
    return (const unsigned char*)im.pixels + int64_t(y - im.y) * im.ystride
           + int64_t(x - im.x) * im.xstride;
}

}  // namespace

// make_kernel (:864-957), depth 1.  Returns 0 on success, 1 for an unknown
// kernel (the reference then hands back a box AND an error).  out may be null
// to query the size.
extern "C" int
orc_make_kernel(const char* name, float width, float height, int normalize, float* out, int* kw,
                int* kh)
{
    OrcFilter2D f;
    const bool have = orc_make_filter2d(name, width, height, &f);
    if (have) {
        if (width == 0.0f)
            width = f.w;
        if (height == 0.0f)
            height = f.h;
    } else {
        if (width == 0.0f)
            width = 4.0f;
        if (height == 0.0f)
            height = 4.0f;
    }
    int w = std::max(1, (int)ceilf(width));
    int h = std::max(1, (int)ceilf(height));
    w |= 1;
    h |= 1;
    *kw = w;
    *kh = h;
    if (!out)
        return 0;
    const int x0 = -w / 2, y0 = -h / 2;
    int rc = 0;
    if (have) {
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i)
                out[j * w + i] = orc_f2d_eval(f, float(x0 + i), float(y0 + j));
    } else if (!strcmp(name, "binomial")) {
        std::vector<float> wf((size_t)std::max(1.0f, width)), hf((size_t)std::max(1.0f, height));
        for (int i = 0; i < width; ++i)
            wf[i] = binomial(width - 1, i);
        for (int i = 0; i < height; ++i)
            hf[i] = binomial(height - 1, i);
        for (int j = 0; j < h; ++j)
            for (int i = 0; i < w; ++i)
                out[j * w + i] = wf[i] * hf[j] * 1.0f;
    } else if (!strcasecmp(name, "laplacian") && w == 3 && h == 3) {
        const float vals[9] = { 0, 1, 0, 1, -4, 1, 0, 1, 0 };
        memcpy(out, vals, sizeof(vals));
        normalize = 0;
    } else {
        float val = normalize ? 1.0f / (w * h * 1) : 1.0f;
        for (int i = 0; i < w * h; ++i)
            out[i] = val;
        return 1;
    }
    if (normalize) {
        float sum = 0;
        for (int i = 0; i < w * h; ++i)
            sum += out[i];
        if (sum != 0.0f)
            for (int i = 0; i < w * h; ++i)
                out[i] = out[i] / sum;
    }
    return rc;
}

// convolve_ (:772-812).  The kernel's window is [-kw/2, -kw/2+kw) x [-kh/2, ...).
extern "C" int
orc_convolve(const orc_image* dst, const orc_image* src, const float* kernel, int kw, int kh,
             int normalize, orc_roi roi, int chbegin, int chend, int nthreads)
{
    if (!dst || !src || !kernel)
        return 1;
    chend = std::min(chend, std::min(dst->nchannels, src->nchannels));
    const int kx0 = -kw / 2, ky0 = -kh / 2;
    const int ses = esize(src->basetype), des = esize(dst->basetype);
    orc_parallel_bands(roi, nthreads, [&](orc_roi r) {
        float scale = 1.0f;
        if (normalize) {
            scale = 0.0f;
            for (int i = 0; i < kw * kh; ++i)
                scale += kernel[i];
            scale = 1.0f / scale;
        }
        std::vector<float> sum((size_t)std::max(chend, 1));
        for (int y = r.ybegin; y < r.yend; ++y)
            for (int x = r.xbegin; x < r.xend; ++x) {
                if (x < dst->x || x >= dst->x + dst->width || y < dst->y
                    || y >= dst->y + dst->height)
                    continue;
                for (int c = chbegin; c < chend; ++c)
                    sum[c] = 0.0f;
                const float* k = kernel;
                for (int j = 0; j < kh; ++j)
                    for (int i = 0; i < kw; ++i, ++k) {
                        const unsigned char* p = clamped_pixel(*src, x + kx0 + i, y + ky0 + j);
                        for (int c = chbegin; c < chend; ++c)
                            sum[c] += k[0]
                                      * (p ? orc_load_as_float(p + c * ses, src->basetype) : 0.0f);
                    }
                unsigned char* dp = (unsigned char*)dst->pixels
                                    + int64_t(y - dst->y) * dst->ystride
                                    + int64_t(x - dst->x) * dst->xstride;
                for (int c = chbegin; c < chend; ++c)
                    orc_store_from_float(dp + c * des, dst->basetype, scale * sum[c]);
            }
    });
    return 0;
}

// unsharp_mask (:975-1029) for every kernel but "median".
extern "C" int
orc_unsharp_mask(const orc_image* dst, const orc_image* src, const char* kernel, float width,
                 float contrast, float threshold, orc_roi roi, int nthreads, char* err,
                 int errlen)
{
    if (!strcmp(kernel, "median")) {
        snprintf(err, errlen, "median not restated");
        return 2;
    }
    const int nc = src->nchannels;
    // Blurry / fst_pass: float images with src's windows, zero filled
    std::vector<float> b1((size_t)src->width * src->height * nc, 0.0f), b2;
    orc_image B = *src;
    B.pixels    = b1.data();
    B.basetype  = ORC_FLOAT;
    B.xstride   = int64_t(nc) * 4;
    B.ystride   = B.xstride * src->width;
    int kw, kh;
    if (width > 3.0) {
        if (orc_make_kernel(kernel, 1, width, 1, nullptr, &kw, &kh) == 0) {
        }
        std::vector<float> K((size_t)kw * kh);
        if (orc_make_kernel(kernel, 1, width, 1, K.data(), &kw, &kh)) {
            snprintf(err, errlen, "Unknown kernel \"%s\" %gx%g", kernel, 1.0, (double)width);
            return 1;
        }
        b2.assign(b1.size(), 0.0f);
        orc_image F = B;
        F.pixels    = b2.data();
        orc_convolve(&F, src, K.data(), kw, kh, 1, roi, 0, nc, nthreads);
        // Kt = transpose(K): kh wide, kw tall, same values
        orc_convolve(&B, &F, K.data(), kh, kw, 1, roi, 0, nc, nthreads);
    } else {
        orc_make_kernel(kernel, width, width, 1, nullptr, &kw, &kh);
        std::vector<float> K((size_t)kw * kh);
        if (orc_make_kernel(kernel, width, width, 1, K.data(), &kw, &kh)) {
            snprintf(err, errlen, "Unknown kernel \"%s\" %gx%g", kernel, (double)width,
                     (double)width);
            return 1;
        }
        orc_convolve(&B, src, K.data(), kw, kh, 1, roi, 0, nc, nthreads);
    }
    // unsharp_impl (:961-985)
    const int ses = esize(src->basetype), des = esize(dst->basetype);
    const int chend = std::min(dst->nchannels, nc);
    for (int y = roi.ybegin; y < roi.yend; ++y)
        for (int x = roi.xbegin; x < roi.xend; ++x) {
            if (x < dst->x || x >= dst->x + dst->width || y < dst->y || y >= dst->y + dst->height)
                continue;
            const bool sin = x >= src->x && x < src->x + src->width && y >= src->y
                             && y < src->y + src->height;
            const unsigned char* sp = sin ? (const unsigned char*)src->pixels
                                                + int64_t(y - src->y) * src->ystride
                                                + int64_t(x - src->x) * src->xstride
                                          : nullptr;
            const float* bp = sin ? b1.data() + ((size_t)(y - src->y) * src->width + (x - src->x)) * nc
                                  : nullptr;
            unsigned char* dp = (unsigned char*)dst->pixels + int64_t(y - dst->y) * dst->ystride
                                + int64_t(x - dst->x) * dst->xstride;
            for (int c = 0; c < chend; ++c) {
                const float s    = sp ? orc_load_as_float(sp + c * ses, src->basetype) : 0.0f;
                const float b    = bp ? bp[c] : 0.0f;
                const float diff = s - b;
                float v          = s;
                if (fabsf(diff) > threshold)
                    v = s + contrast * diff;
                orc_store_from_float(dp + c * des, dst->basetype, v);
            }
        }
    return 0;
}
