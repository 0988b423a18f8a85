// oracle_resize.cpp -- CPU restatement of ImageBufAlgo::resize / resample /
// fit geometry / maketx resize_block.  TEST INFRASTRUCTURE ONLY (oracle.h).
//
// Follows, in /root/reference (OpenImageIO 3.2.0.2-dev):
//   src/libOpenImageIO/imagebufalgo_xform.cpp:595-826   resize_<D,S>
//   src/libOpenImageIO/imagebufalgo_xform.cpp:830-860   get_resize_filter
//   src/libOpenImageIO/imagebufalgo_xform.cpp:962-997   fit geometry
//   src/libOpenImageIO/imagebufalgo_xform.cpp:1107-1186 resample_scalar
//   src/libOpenImageIO/imagebufalgo_xform.cpp:1296-1493 resample fast paths
//   src/libOpenImageIO/imagebuf.cpp:3283-3322,3623-3667 WrapClamp / black
//   src/include/OpenImageIO/fmath.h:584-596 floorfrac, :430-440 bilerp,
//       :753-764 scaled_conversion, :1009-1031 convert_type
//   src/include/OpenImageIO/imagebufalgo_util.h:37-80 + libutil/thread.cpp:
//       837-871  row-band threading of parallel_image
//   src/libOpenImageIO/maketexture.cpp:141-334 resize_block (box fast path)
//
// The loop nest, zero-weight skips, normalisation and accumulation order are
// kept as in the reference (single pass, ytaps x xtaps per output pixel).
// Compile with -ffp-contract=off.
#include "oracle.h"

#include <algorithm>
#include <atomic>
#include <string>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct OrcFilter2D {
    int kind;
    float w, h;
};
bool
orc_make_filter2d(const char* name, float w, float h, OrcFilter2D* f);
bool
orc_f2d_separable(const OrcFilter2D& f);
float
orc_f2d_eval(const OrcFilter2D& f, float x, float y);
float
orc_f2d_xfilt(const OrcFilter2D& f, float x);
float
orc_f2d_yfilt(const OrcFilter2D& f, float y);

namespace {

inline float
floorfrac(float x, int* xint)
{
    float f = std::floor(x);
    *xint   = int(f);
    return x - f;
}
inline int
ifloor(float x)
{
    return (int)floorf(x);
}
inline float
clampf(float a, float low, float high)
{
    float val = a;
    if (!(low <= val))
        val = low;
    if (val > high)
        val = high;
    return val;
}
inline int
clampi(int a, int low, int high)
{
    int val = a;
    if (!(low <= val))
        val = low;
    if (val > high)
        val = high;
    return val;
}

inline int
type_size(int bt)
{
    switch (bt) {
    case ORC_UINT8: return 1;
    case ORC_UINT16: return 2;
    case ORC_HALF: return 2;
    case ORC_FLOAT: return 4;
    }
    return 0;
}

inline float
load_as_float(const void* p, int bt)
{
    switch (bt) {
    case ORC_UINT8: return float(*(const uint8_t*)p) * (1.0f / 255.0f);
    case ORC_UINT16: return float(*(const uint16_t*)p) * (1.0f / 65535.0f);
    case ORC_HALF: return float(*(const _Float16*)p);
    default: return *(const float*)p;
    }
}
inline void
store_from_float(void* p, int bt, float v)
{
    switch (bt) {
    case ORC_UINT8: {
        float s = v * 255.0f;
        s += (s < 0 ? -0.5f : 0.5f);
        *(uint8_t*)p = (uint8_t)clampf(s, 0.0f, 255.0f);
        break;
    }
    case ORC_UINT16: {
        float s = v * 65535.0f;
        s += (s < 0 ? -0.5f : 0.5f);
        *(uint16_t*)p = (uint16_t)clampf(s, 0.0f, 65535.0f);
        break;
    }
    case ORC_HALF: *(_Float16*)p = (_Float16)v; break;
    default: *(float*)p = v; break;
    }
}

// Source fetch with the iterator semantics of ConstIterator(src, WrapClamp):
// inside the data window -> that pixel; otherwise clamp (x,y) to the FULL
// window and, if the clamped position is still outside the data window,
// the black pixel.  wrap_black=true gives WrapBlack (no remap).
struct SrcView {
    const orc_image* im;
    int tsize;
    const char* pixel(int x, int y, bool wrap_black) const
    {
        bool inside = x >= im->x && x < im->x + im->width && y >= im->y
                      && y < im->y + im->height;
        if (!inside) {
            if (wrap_black)
                return nullptr;
            x = clampi(x, im->full_x, im->full_x + im->full_width - 1);
            y = clampi(y, im->full_y, im->full_y + im->full_height - 1);
            inside = x >= im->x && x < im->x + im->width && y >= im->y
                     && y < im->y + im->height;
            if (!inside)
                return nullptr;
        }
        return (const char*)im->pixels + int64_t(y - im->y) * im->ystride
               + int64_t(x - im->x) * im->xstride;
    }
};

inline char*
dst_pixel(const orc_image* im, int x, int y)
{
    return (char*)im->pixels + int64_t(y - im->y) * im->ystride
           + int64_t(x - im->x) * im->xstride;
}

// parallel_image + parallel_for_chunked_2D_id: full-width row bands of
// max(1, H/(2*maxthreads)) rows; maxthreads = min(threads, 1+npixels/1024).
void
parallel_bands(orc_roi roi, int nthreads,
               const std::function<void(orc_roi)>& f)
{
    if (nthreads <= 0)
        nthreads = std::max(1u, std::thread::hardware_concurrency());
    int64_t npix = int64_t(roi.xend - roi.xbegin) * (roi.yend - roi.ybegin);
    int maxthreads = std::min<int64_t>(nthreads, 1 + npix / 1024);
    if (maxthreads <= 1) {
        f(roi);
        return;
    }
    int64_t h      = roi.yend - roi.ybegin;
    int64_t ychunk = std::max<int64_t>(1, h / (2 * maxthreads));
    if (ychunk >= h) {
        f(roi);
        return;
    }
    std::vector<orc_roi> bands;
    for (int64_t y = roi.ybegin; y < roi.yend; y += ychunk) {
        orc_roi b = roi;
        b.ybegin  = int(y);
        b.yend    = int(std::min<int64_t>(roi.yend, y + ychunk));
        bands.push_back(b);
    }
    std::vector<std::thread> pool;
    std::atomic<size_t> next { 0 };
    auto worker = [&]() {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= bands.size())
                break;
            f(bands[i]);
        }
    };
    for (int t = 0; t < maxthreads; ++t)
        pool.emplace_back(worker);
    for (auto& t : pool)
        t.join();
}

template<int BT> struct TypedLoad;
template<> struct TypedLoad<ORC_UINT8> {
    static inline float get(const char* p, int c)
    {
        return float(((const uint8_t*)p)[c]) * (1.0f / 255.0f);
    }
};
template<> struct TypedLoad<ORC_UINT16> {
    static inline float get(const char* p, int c)
    {
        return float(((const uint16_t*)p)[c]) * (1.0f / 65535.0f);
    }
};
template<> struct TypedLoad<ORC_HALF> {
    static inline float get(const char* p, int c)
    {
        return float(((const _Float16*)p)[c]);
    }
};
template<> struct TypedLoad<ORC_FLOAT> {
    static inline float get(const char* p, int c) { return ((const float*)p)[c]; }
};

// The separable tap loop of resize_ (xform.cpp:747-761) for one output pixel.
// Same order of operations as the generic code below; the source type is a
// template parameter and, when the whole footprint lies inside the data
// window, pixels are addressed directly instead of through the per-tap
// WrapClamp test (what the reference's iterator does for in-window pixels).
template<int BT>
static inline void
separable_taps(const SrcView& sv, const orc_image* src, int src_x, int src_y,
               int radi, int radj, int xtaps, const float* xfiltval,
               const float* yfiltval, int nchannels, float* pel)
{
    const bool interior = src_x - radi >= src->x && src_x + radi < src->x + src->width
                          && src_y - radj >= src->y
                          && src_y + radj < src->y + src->height;
    for (int j = -radj; j <= radj; ++j) {
        float wy = yfiltval[j + radj];
        if (wy == 0.0f)
            continue;
        const char* row = nullptr;
        if (interior)
            row = (const char*)src->pixels + int64_t(src_y + j - src->y) * src->ystride
                  + int64_t(src_x - radi - src->x) * src->xstride;
        for (int i = 0; i < xtaps; ++i) {
            float w = wy * xfiltval[i];
            if (w) {
                const char* sp = interior ? row + int64_t(i) * src->xstride
                                          : sv.pixel(src_x - radi + i, src_y + j, false);
                if (sp)
                    for (int c = 0; c < nchannels; ++c)
                        pel[c] += w * TypedLoad<BT>::get(sp, c);
            }
        }
    }
}

void
resize_band(const orc_image* dst, const orc_image* src,
            const OrcFilter2D& filter, orc_roi roi)
{
    const int nchannels = dst->nchannels;
    const int stsize    = type_size(src->basetype);
    const int dtsize    = type_size(dst->basetype);
    SrcView sv { src, stsize };

    float srcfx = src->full_x;
    float srcfy = src->full_y;
    float srcfw = src->full_width;
    float srcfh = src->full_height;

    float xratio = float(dst->full_width) / srcfw;
    float yratio = float(dst->full_height) / srcfh;

    float dstfx          = float(dst->full_x);
    float dstfy          = float(dst->full_y);
    float dstfw          = float(dst->full_width);
    float dstfh          = float(dst->full_height);
    float dstpixelwidth  = 1.0f / dstfw;
    float dstpixelheight = 1.0f / dstfh;
    float filterrad      = filter.w / 2.0f;

    int radi       = (int)ceilf(filterrad / xratio);
    int radj       = (int)ceilf(filterrad / yratio);
    int xtaps      = 2 * radi + 1;
    int ytaps      = 2 * radj + 1;
    bool separable = orc_f2d_separable(filter);
    std::vector<float> yfiltval(ytaps);
    std::vector<float> xfiltval_all;
    const int rw = roi.xend - roi.xbegin;
    if (separable) {
        xfiltval_all.resize(size_t(xtaps) * rw);
        for (int x = roi.xbegin; x < roi.xend; ++x) {
            float* xfiltval = xfiltval_all.data() + size_t(x - roi.xbegin) * xtaps;
            float s         = (x - dstfx + 0.5f) * dstpixelwidth;
            float src_xf    = srcfx + s * srcfw;
            int src_x;
            float src_xf_frac   = floorfrac(src_xf, &src_x);
            float totalweight_x = 0.0f;
            for (int i = 0; i < xtaps; ++i) {
                float w     = orc_f2d_xfilt(filter,
                                            xratio * (i - radi - (src_xf_frac - 0.5f)));
                xfiltval[i] = w;
                totalweight_x += w;
            }
            if (totalweight_x != 0.0f)
                for (int i = 0; i < xtaps; ++i)
                    xfiltval[i] /= totalweight_x;
        }
    }

    std::vector<float> pel(nchannels);

    if (separable) {
        for (int y = roi.ybegin; y < roi.yend; ++y) {
            float t      = (y - dstfy + 0.5f) * dstpixelheight;
            float src_yf = srcfy + t * srcfh;
            int src_y;
            float src_yf_frac   = floorfrac(src_yf, &src_y);
            float totalweight_y = 0.0f;
            for (int j = 0; j < ytaps; ++j) {
                float w     = orc_f2d_yfilt(filter,
                                            yratio * (j - radj - (src_yf_frac - 0.5f)));
                yfiltval[j] = w;
                totalweight_y += w;
            }
            if (totalweight_y != 0.0f)
                for (int i = 0; i < ytaps; ++i)
                    yfiltval[i] /= totalweight_y;

            for (int x = roi.xbegin; x < roi.xend; ++x) {
                float s      = (x - dstfx + 0.5f) * dstpixelwidth;
                float src_xf = srcfx + s * srcfw;
                int src_x    = ifloor(src_xf);
                for (int c = 0; c < nchannels; ++c)
                    pel[c] = 0.0f;
                const float* xfiltval = xfiltval_all.data()
                                        + size_t(x - roi.xbegin) * xtaps;
                float totalweight_x = 0.0f;
                for (int i = 0; i < xtaps; ++i)
                    totalweight_x += xfiltval[i];
                if (totalweight_x != 0.0f) {
                    switch (src->basetype) {
                    case ORC_UINT8:
                        separable_taps<ORC_UINT8>(sv, src, src_x, src_y, radi, radj,
                                                  xtaps, xfiltval, yfiltval.data(),
                                                  nchannels, pel.data());
                        break;
                    case ORC_UINT16:
                        separable_taps<ORC_UINT16>(sv, src, src_x, src_y, radi, radj,
                                                   xtaps, xfiltval, yfiltval.data(),
                                                   nchannels, pel.data());
                        break;
                    case ORC_HALF:
                        separable_taps<ORC_HALF>(sv, src, src_x, src_y, radi, radj,
                                                 xtaps, xfiltval, yfiltval.data(),
                                                 nchannels, pel.data());
                        break;
                    default:
                        separable_taps<ORC_FLOAT>(sv, src, src_x, src_y, radi, radj,
                                                  xtaps, xfiltval, yfiltval.data(),
                                                  nchannels, pel.data());
                        break;
                    }
                }
                char* dp = dst_pixel(dst, x, y);
                if (totalweight_y == 0.0f) {
                    for (int c = 0; c < nchannels; ++c)
                        store_from_float(dp + c * dtsize, dst->basetype, 0.0f);
                } else {
                    for (int c = 0; c < nchannels; ++c)
                        store_from_float(dp + c * dtsize, dst->basetype, pel[c]);
                }
            }
        }
    } else {
        for (int y = roi.ybegin; y < roi.yend; ++y) {
            float t      = (y - dstfy + 0.5f) * dstpixelheight;
            float src_yf = srcfy + t * srcfh;
            int src_y;
            float src_yf_frac = floorfrac(src_yf, &src_y);
            for (int x = roi.xbegin; x < roi.xend; ++x) {
                float s      = (x - dstfx + 0.5f) * dstpixelwidth;
                float src_xf = srcfx + s * srcfw;
                int src_x;
                float src_xf_frac = floorfrac(src_xf, &src_x);
                for (int c = 0; c < nchannels; ++c)
                    pel[c] = 0.0f;
                float totalweight = 0.0f;
                // Reference quirk kept: the iterator's y range is built from
                // radi (xform.cpp:793-794) while the loop runs j over radj.
                // The iterator walks its range row-major, so tap (i,j) reads
                // position n = (j+radj)*xtaps + (i+radi) of a window that is
                // xtaps wide and starts at (src_x-radi, src_y-radi).
                int n = 0;
                for (int j = -radj; j <= radj; ++j) {
                    for (int i = -radi; i <= radi; ++i, ++n) {
                        float w = orc_f2d_eval(filter,
                                               xratio * (i - (src_xf_frac - 0.5f)),
                                               yratio * (j - (src_yf_frac - 0.5f)));
                        if (w) {
                            totalweight += w;
                            int px         = src_x - radi + (n % xtaps);
                            int py         = src_y - radi + (n / xtaps);
                            const char* sp = sv.pixel(px, py, false);
                            if (sp)
                                for (int c = 0; c < nchannels; ++c)
                                    pel[c] += w
                                              * load_as_float(sp + c * stsize,
                                                              src->basetype);
                        }
                    }
                }
                char* dp = dst_pixel(dst, x, y);
                if (totalweight == 0.0f) {
                    for (int c = 0; c < nchannels; ++c)
                        store_from_float(dp + c * dtsize, dst->basetype, 0.0f);
                } else {
                    for (int c = 0; c < nchannels; ++c)
                        store_from_float(dp + c * dtsize, dst->basetype,
                                         pel[c] / totalweight);
                }
            }
        }
    }
}

void
seterr(char* err, int errlen, const char* msg)
{
    if (err && errlen > 0)
        snprintf(err, errlen, "%s", msg);
}

bool
legal_pair(int d, int s)
{
    // OIIO_DISPATCH_COMMON_TYPES2, imagebufalgo_util.h:463-552
    if (d == ORC_FLOAT)
        return s == ORC_FLOAT || s == ORC_UINT8 || s == ORC_HALF || s == ORC_UINT16;
    if (d == ORC_UINT8)
        return s == ORC_FLOAT || s == ORC_UINT8;
    if (d == ORC_HALF)
        return s == ORC_FLOAT || s == ORC_HALF;
    if (d == ORC_UINT16)
        return s == ORC_FLOAT || s == ORC_UINT16;
    return false;
}

}  // namespace

extern "C" {

float
orc_load_as_float(const void* p, int basetype)
{
    return load_as_float(p, basetype);
}
void
orc_store_from_float(void* p, int basetype, float v)
{
    store_from_float(p, basetype, v);
}

int
orc_resize(const orc_image* dst, const orc_image* src, const char* filtername,
           float filterwidth, orc_roi roi, int nthreads, char* err, int errlen)
{
    if (!dst || !src || !src->pixels || !dst->pixels) {
        seterr(err, errlen, "Uninitialized input image");
        return 1;
    }
    if (!legal_pair(dst->basetype, src->basetype)) {
        seterr(err, errlen, "oracle: type pair needs a float temporary");
        return 2;
    }
    // get_resize_filter, xform.cpp:830-860
    float wratio = float(dst->full_width) / float(src->full_width);
    float hratio = float(dst->full_height) / float(src->full_height);
    const char* name = filtername;
    if (!name || !name[0])
        name = (wratio > 1.0f || hratio > 1.0f) ? "blackman-harris" : "lanczos3";
    int idx = orc_filter_lookup(name);
    if (idx < 0) {
        char buf[256];
        snprintf(buf, sizeof(buf), "Filter \"%s\" not recognized", name);
        seterr(err, errlen, buf);
        return 3;
    }
    float fdw = orc_filter_default_width(idx);
    float w   = filterwidth > 0.0f ? filterwidth : fdw * std::max(1.0f, wratio);
    float h   = filterwidth > 0.0f ? filterwidth : fdw * std::max(1.0f, hratio);
    OrcFilter2D filter;
    if (!orc_make_filter2d(name, w, h, &filter)) {
        seterr(err, errlen, "oracle: filter create failed");
        return 3;
    }
    // IBAprep: roi shrink-wrapped to dst's data window
    roi.xbegin = std::max(roi.xbegin, dst->x);
    roi.xend   = std::min(roi.xend, dst->x + dst->width);
    roi.ybegin = std::max(roi.ybegin, dst->y);
    roi.yend   = std::min(roi.yend, dst->y + dst->height);
    if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin)
        return 0;
    parallel_bands(roi, nthreads,
                   [&](orc_roi r) { resize_band(dst, src, filter, r); });
    return 0;
}

int
orc_resample(const orc_image* dst, const orc_image* src, int interpolate,
             orc_roi roi, int nthreads, char* err, int errlen)
{
    if (!dst || !src || !src->pixels || !dst->pixels) {
        seterr(err, errlen, "Uninitialized input image");
        return 1;
    }
    if (!legal_pair(dst->basetype, src->basetype)) {
        seterr(err, errlen, "oracle: type pair needs a float temporary");
        return 2;
    }
    roi.xbegin = std::max(roi.xbegin, dst->x);
    roi.xend   = std::min(roi.xend, dst->x + dst->width);
    roi.ybegin = std::max(roi.ybegin, dst->y);
    roi.yend   = std::min(roi.yend, dst->y + dst->height);
    if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin)
        return 0;
    const int nch    = std::min(src->nchannels, dst->nchannels);
    const int stsize = type_size(src->basetype);
    const int dtsize = type_size(dst->basetype);
    const bool data_is_full = src->x == src->full_x && src->y == src->full_y
                              && src->width == src->full_width
                              && src->height == src->full_height;
    parallel_bands(roi, nthreads, [&](orc_roi r) {
        SrcView sv { src, stsize };
        float srcfx = src->full_x, srcfy = src->full_y;
        float srcfw = src->full_width, srcfh = src->full_height;
        float dstfx = dst->full_x, dstfy = dst->full_y;
        float dstpixelwidth  = 1.0f / float(dst->full_width);
        float dstpixelheight = 1.0f / float(dst->full_height);
        std::vector<float> p(4 * nch), pel(nch);
        for (int y = r.ybegin; y < r.yend; ++y) {
            float t      = (y - dstfy + 0.5f) * dstpixelheight;
            float src_yf = srcfy + t * srcfh;
            int src_y    = ifloor(src_yf);
            for (int x = r.xbegin; x < r.xend; ++x) {
                float s      = (x - dstfx + 0.5f) * dstpixelwidth;
                float src_xf = srcfx + s * srcfw;
                int src_x    = ifloor(src_xf);
                char* dp     = dst_pixel(dst, x, y);
                if (interpolate) {
                    float xx = src_xf - 0.5f, yy = src_yf - 0.5f;
                    int xt, yt;
                    float xfrac = floorfrac(xx, &xt);
                    float yfrac = floorfrac(yy, &yt);
                    for (int k = 0; k < 4; ++k) {
                        int px = xt + (k & 1), py = yt + (k >> 1);
                        const char* sp;
                        if (data_is_full) {
                            // resample_bilinear: taps clamped to the data
                            // window (xform.cpp:1343-1347)
                            px = clampi(px, src->x, src->x + src->width - 1);
                            py = clampi(py, src->y, src->y + src->height - 1);
                            sp = sv.pixel(px, py, true);
                        } else {
                            sp = sv.pixel(px, py, false);  // WrapClamp
                        }
                        for (int c = 0; c < nch; ++c)
                            p[k * nch + c]
                                = sp ? load_as_float(sp + c * stsize, src->basetype)
                                     : 0.0f;
                    }
                    float s1 = 1.0f - xfrac, t1 = 1.0f - yfrac;
                    for (int c = 0; c < nch; ++c)
                        pel[c] = t1 * (p[c] * s1 + p[nch + c] * xfrac)
                                 + yfrac
                                       * (p[2 * nch + c] * s1
                                          + p[3 * nch + c] * xfrac);
                } else {
                    const char* sp = sv.pixel(src_x, src_y, true);  // WrapBlack
                    for (int c = 0; c < nch; ++c)
                        pel[c] = sp ? load_as_float(sp + c * stsize, src->basetype)
                                    : 0.0f;
                }
                for (int c = 0; c < nch; ++c)
                    store_from_float(dp + c * dtsize, dst->basetype, pel[c]);
            }
        }
    });
    return 0;
}

void
orc_fit_geometry(int src_full_w, int src_full_h, int fit_w, int fit_h,
                 const char* fillmode, int* resize_w, int* resize_h,
                 int* xoffset, int* yoffset)
{
    // xform.cpp:962-997
    float oldaspect = float(src_full_w) / src_full_h;
    float newaspect = float(fit_w) / fit_h;
    int rw = fit_w, rh = fit_h, xo = 0, yo = 0;
    std::string mode = fillmode ? fillmode : "letterbox";
    if (mode != "height" && mode != "width")
        mode = "letterbox";
    if (mode == "letterbox")
        mode = (newaspect >= oldaspect) ? "height" : "width";
    if (mode == "height") {
        rw = int(rh * oldaspect + 0.5f);
        xo = (fit_w - rw) / 2;
    } else {
        rh = int(rw / oldaspect + 0.5f);
        yo = (fit_h - rh) / 2;
    }
    *resize_w = rw;
    *resize_h = rh;
    *xoffset  = xo;
    *yoffset  = yo;
}

}  // extern "C"


// row-band threading for the other oracle translation units
void
orc_parallel_bands(orc_roi roi, int nthreads, const std::function<void(orc_roi)>& f)
{
    parallel_bands(roi, nthreads, f);
}
