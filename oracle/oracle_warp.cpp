// oracle_warp.cpp -- CPU restatement of ImageBufAlgo::warp / rotate and the
// exact=1 branch of fit (src/libOpenImageIO/imagebufalgo_xform.cpp:263-418,
// 1727-1760, 1012-1027).
//
// TEST INFRASTRUCTURE ONLY (see oracle.h).  The 3x3 matrix algebra is Imath's
// (Matrix33<float>: translate / scale / rotate / operator*= / inverse,
// Imath 3.1 ImathMatrix.h); Imath is an external dependency that is not
// vendored in the reference tree, so its published algorithm is restated here
// and anchored on the reference's own golden images: testsuite/oiiotool-xform/
// ref/{warped,rotated,rotated-offcenter,resizefrom*}.tif, fit4.exr and the
// twelve fit[wh]-<mode>-<size>.exr (tests/golden/warp_ref.npz, tests/test_cpu.py).
#include "oracle.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
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

struct M33 {
    float x[3][3];
};

// Imath Matrix33<T>::inverse() (non-throwing form)
M33
m33_inverse(const M33& m)
{
    const float(*x)[3] = m.x;
    M33 zero;
    memset(&zero, 0, sizeof(zero));
    zero.x[0][0] = zero.x[1][1] = zero.x[2][2] = 1.0f;  // Matrix33() is identity
    const float tiny = 1.17549435e-38f;                 // numeric_limits<float>::min()
    if (x[0][2] != 0 || x[1][2] != 0 || x[2][2] != 1) {
        M33 s;
        s.x[0][0] = x[1][1] * x[2][2] - x[2][1] * x[1][2];
        s.x[0][1] = x[2][1] * x[0][2] - x[0][1] * x[2][2];
        s.x[0][2] = x[0][1] * x[1][2] - x[1][1] * x[0][2];
        s.x[1][0] = x[2][0] * x[1][2] - x[1][0] * x[2][2];
        s.x[1][1] = x[0][0] * x[2][2] - x[2][0] * x[0][2];
        s.x[1][2] = x[1][0] * x[0][2] - x[0][0] * x[1][2];
        s.x[2][0] = x[1][0] * x[2][1] - x[2][0] * x[1][1];
        s.x[2][1] = x[2][0] * x[0][1] - x[0][0] * x[2][1];
        s.x[2][2] = x[0][0] * x[1][1] - x[1][0] * x[0][1];
        float r   = x[0][0] * s.x[0][0] + x[0][1] * s.x[1][0] + x[0][2] * s.x[2][0];
        if (fabsf(r) >= 1) {
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j)
                    s.x[i][j] /= r;
        } else {
            float mr = fabsf(r) / tiny;
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) {
                    if (mr > fabsf(s.x[i][j]))
                        s.x[i][j] /= r;
                    else
                        return zero;
                }
        }
        return s;
    }
    M33 s;
    s.x[0][0] = x[1][1], s.x[0][1] = -x[0][1], s.x[0][2] = 0;
    s.x[1][0] = -x[1][0], s.x[1][1] = x[0][0], s.x[1][2] = 0;
    s.x[2][0] = 0, s.x[2][1] = 0, s.x[2][2] = 1;
    float r = x[0][0] * x[1][1] - x[1][0] * x[0][1];
    if (fabsf(r) >= 1) {
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                s.x[i][j] /= r;
    } else {
        float mr = fabsf(r) / tiny;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) {
                if (mr > fabsf(s.x[i][j]))
                    s.x[i][j] /= r;
                else
                    return zero;
            }
    }
    s.x[2][0] = -x[2][0] * s.x[0][0] - x[2][1] * s.x[1][0];
    s.x[2][1] = -x[2][0] * s.x[0][1] - x[2][1] * s.x[1][1];
    return s;
}

void
m33_identity(M33& m)
{
    memset(&m, 0, sizeof(m));
    m.x[0][0] = m.x[1][1] = m.x[2][2] = 1.0f;
}
// Matrix33::translate(V2)
void
m33_translate(M33& m, float tx, float ty)
{
    m.x[2][0] += tx * m.x[0][0] + ty * m.x[1][0];
    m.x[2][1] += tx * m.x[0][1] + ty * m.x[1][1];
    m.x[2][2] += tx * m.x[0][2] + ty * m.x[1][2];
}
// Matrix33::scale(V2)
void
m33_scale(M33& m, float sx, float sy)
{
    for (int j = 0; j < 3; ++j) {
        m.x[0][j] *= sx;
        m.x[1][j] *= sy;
    }
}
// Matrix33::operator*=
void
m33_mul(M33& a, const M33& b)
{
    M33 t;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float v = 0.0f;
            for (int k = 0; k < 3; ++k)
                v += a.x[i][k] * b.x[k][j];
            t.x[i][j] = v;
        }
    a = t;
}
// Matrix33::rotate(r): *this *= Matrix33().setRotation(r)
void
m33_rotate(M33& m, float r)
{
    float c = cosf(r), s = sinf(r);
    M33 rot;
    m33_identity(rot);
    rot.x[0][0] = c, rot.x[0][1] = s;
    rot.x[1][0] = -s, rot.x[1][1] = c;
    m33_mul(m, rot);
}

// Dual2 (xform.cpp:152-206)
struct Dual2 {
    float v, dx, dy;
};
inline Dual2
d_mulf(Dual2 a, float b)
{
    return { a.v * b, a.dx * b, a.dy * b };
}
inline Dual2
d_add(Dual2 a, Dual2 b)
{
    return { a.v + b.v, a.dx + b.dx, a.dy + b.dy };
}
inline Dual2
d_addf(Dual2 a, float b)
{
    return { a.v + b, a.dx, a.dy };
}
inline Dual2
d_div(Dual2 a, Dual2 b)
{
    float bvalinv   = 1.0f / b.v;
    float aval_bval = a.v * bvalinv;
    return { aval_bval, bvalinv * (a.dx - aval_bval * b.dx), bvalinv * (a.dy - aval_bval * b.dy) };
}

// robust_multVecMatrix (xform.cpp:210-226)
void
xform_point(const M33& M, Dual2 x, Dual2 y, Dual2* ox, Dual2* oy)
{
    Dual2 a = d_addf(d_add(d_mulf(x, M.x[0][0]), d_mulf(y, M.x[1][0])), M.x[2][0]);
    Dual2 b = d_addf(d_add(d_mulf(x, M.x[0][1]), d_mulf(y, M.x[1][1])), M.x[2][1]);
    Dual2 w = d_addf(d_add(d_mulf(x, M.x[0][2]), d_mulf(y, M.x[1][2])), M.x[2][2]);
    if (w.v != 0.0f) {
        *ox = d_div(a, w);
        *oy = d_div(b, w);
    } else {
        *ox = { 0, 0, 0 };
        *oy = { 0, 0, 0 };
    }
}

inline int
clampi(int v, int lo, int hi)
{
    return v < lo ? lo : (v > hi ? hi : v);
}

// ImageBufImpl::do_wrap (imagebuf.cpp:3283-3322) + wrap_periodic / wrap_mirror
// (imageio.cpp:1398-1434).  Returns the pixel address or nullptr (black).
const unsigned char*
wrapped_pixel(const orc_image& im, int x, int y, int wrap)
{
    bool inside = x >= im.x && x < im.x + im.width && y >= im.y && y < im.y + im.height;
    if (!inside) {
        if (wrap == 2) {  // clamp
            x = clampi(x, im.full_x, im.full_x + im.full_width - 1);
            y = clampi(y, im.full_y, im.full_y + im.full_height - 1);
        } else if (wrap == 3) {  // periodic
            x -= im.full_x, x %= im.full_width;
            if (x < 0)
                x += im.full_width;
            x += im.full_x;
            y -= im.full_y, y %= im.full_height;
            if (y < 0)
                y += im.full_height;
            y += im.full_y;
        } else if (wrap == 4) {  // mirror
            auto mirror = [](int c, int origin, int width) {
                c -= origin;
                if (c < 0)
                    c = -c - 1;
                int iter = c / width;
                c -= iter * width;
                if (iter & 1)
                    c = width - 1 - c;
                return c + origin;
            };
            x = mirror(x, im.full_x, im.full_width);
            y = mirror(y, im.full_y, im.full_height);
        } else {
            return nullptr;  // black (and default)
        }
        inside = x >= im.x && x < im.x + im.width && y >= im.y && y < im.y + im.height;
        if (!inside)
            return nullptr;
    }
    return (const unsigned char*)im.pixels + int64_t(y - im.y) * im.ystride
           + int64_t(x - im.x) * im.xstride;
}

int
esize(int bt)
{
    return bt == ORC_UINT8 ? 1 : (bt == ORC_FLOAT ? 4 : 2);
}

// filtered_sample (xform.cpp:263-325)
void
filtered_sample(const orc_image& src, float s, float t, float dsdx, float dtdx, float dsdy,
                float dtdy, const OrcFilter2D& filter, int wrap, bool edgeclamp, float* result)
{
    const int nc = src.nchannels;
    float ds     = std::max(1.0f, std::max(fabsf(dsdx), fabsf(dsdy)));
    float dt     = std::max(1.0f, std::max(fabsf(dtdx), fabsf(dtdy)));
    float ds_inv = 1.0f / ds;
    float dt_inv = 1.0f / dt;
    float filterrad_s = 0.5f * ds * filter.w;
    float filterrad_t = 0.5f * dt * filter.w;  // width() for both (:275-276)
    int smin = (int)floorf(s - filterrad_s);
    int smax = (int)ceilf(s + filterrad_s);
    int tmin = (int)floorf(t - filterrad_t);
    int tmax = (int)ceilf(t + filterrad_t);
    const int xb = src.x, xe = src.x + src.width, yb = src.y, ye = src.y + src.height;
    if (edgeclamp) {
        smin = clampi(smin, xb, xe);
        smax = clampi(smax, xb, xe);
        tmin = clampi(tmin, yb, ye);
        tmax = clampi(tmax, yb, ye);
        if (s < xb - 1 || s >= xe || t < yb - 1 || t >= ye) {
            for (int c = 0; c < nc; ++c)
                result[c] = 0.0f;
            return;
        }
    }
    float sum[64];
    for (int c = 0; c < nc; ++c)
        sum[c] = 0.0f;
    float total_w = 0.0f;
    const int ses = esize(src.basetype);
    for (int y = tmin; y < tmax; ++y) {
        for (int x = smin; x < smax; ++x) {
            float w = orc_f2d_eval(filter, ds_inv * (x + 0.5f - s), dt_inv * (y + 0.5f - t));
            const unsigned char* p = wrapped_pixel(src, x, y, wrap);
            for (int c = 0; c < nc; ++c)
                sum[c] += w * (p ? orc_load_as_float(p + c * ses, src.basetype) : 0.0f);
            total_w += w;
        }
    }
    if (total_w > 0.0f)
        for (int c = 0; c < nc; ++c)
            result[c] = sum[c] / total_w;
    else
        for (int c = 0; c < nc; ++c)
            result[c] = 0.0f;
}

}  // namespace

extern "C" void
orc_m33_inverse(const float* M, float* out)
{
    M33 m;
    memcpy(m.x, M, sizeof(m.x));
    M33 r = m33_inverse(m);
    memcpy(out, r.x, sizeof(r.x));
}

// ImageBufAlgo::rotate's matrix (xform.cpp:1733-1737)
extern "C" void
orc_rotate_matrix(float angle, float cx, float cy, float* out)
{
    M33 M, T;
    m33_identity(M);
    m33_translate(M, -cx, -cy);
    m33_rotate(M, angle);
    m33_identity(T);
    m33_translate(T, cx, cy);
    m33_mul(M, T);
    memcpy(out, M.x, sizeof(M.x));
}

// oiiotool's --resize:from/to matrix (oiiotool.cpp:4704-4707)
extern "C" void
orc_fromto_matrix(float from_x, float from_y, float from_w, float from_h, float to_x,
                  float to_y, float to_w, float to_h, float* out)
{
    M33 M;
    m33_identity(M);
    m33_translate(M, to_x, to_y);
    m33_scale(M, to_w / from_w, to_h / from_h);
    m33_translate(M, -from_x, -from_y);
    memcpy(out, M.x, sizeof(M.x));
}

// transform(M, roi) (xform.cpp:231-253): bounding box of the transformed ROI
extern "C" void
orc_transform_roi(const float* Mf, orc_roi roi, orc_roi* out)
{
    M33 M;
    memcpy(M.x, Mf, sizeof(M.x));
    float px[4] = { roi.xbegin + 0.5f, roi.xend - 0.5f, roi.xbegin + 0.5f, roi.xend - 0.5f };
    float py[4] = { roi.ybegin + 0.5f, roi.ybegin + 0.5f, roi.yend - 0.5f, roi.yend - 0.5f };
    float minx = 0, miny = 0, maxx = 0, maxy = 0;
    for (int i = 0; i < 4; ++i) {
        // Imath multVecMatrix (V2): a, b, w then divide
        float a = px[i] * M.x[0][0] + py[i] * M.x[1][0] + M.x[2][0];
        float b = px[i] * M.x[0][1] + py[i] * M.x[1][1] + M.x[2][1];
        float w = px[i] * M.x[0][2] + py[i] * M.x[1][2] + M.x[2][2];
        a /= w, b /= w;
        if (i == 0) {
            minx = maxx = a, miny = maxy = b;
        } else {
            minx = std::min(minx, a), maxx = std::max(maxx, a);
            miny = std::min(miny, b), maxy = std::max(maxy, b);
        }
    }
    out->xbegin = int(floorf(minx));
    out->ybegin = int(floorf(miny));
    out->xend   = int(floorf(maxx)) + 1;
    out->yend   = int(floorf(maxy)) + 1;
}

// warp_<D,S> (xform.cpp:353-375) over `roi` of dst.  The filter is given
// resolved (name + width + height): warp() resolves (name, width) with
// get_warp_filter (:329-349), fit(exact) with get_resize_filter.
// wrap: 0 default, 1 black, 2 clamp, 3 periodic, 4 mirror.
extern "C" int
orc_warp(const orc_image* dst, const orc_image* src, const float* Mf, const char* filtername,
         float fw, float fh, int wrap, int edgeclamp, orc_roi roi, int nthreads, char* err,
         int errlen)
{
    OrcFilter2D filter;
    if (!orc_make_filter2d(filtername && filtername[0] ? filtername : "lanczos3", fw, fh,
                           &filter)) {
        snprintf(err, errlen, "Filter \"%s\" not recognized", filtername ? filtername : "");
        return 1;
    }
    M33 M;
    memcpy(M.x, Mf, sizeof(M.x));
    const M33 Minv = m33_inverse(M);
    const int nc   = std::min(dst->nchannels, src->nchannels);
    const int des  = esize(dst->basetype);
    orc_parallel_bands(roi, nthreads, [&](orc_roi r) {
        float pel[64];
        for (int y = r.ybegin; y < r.yend; ++y)
            for (int x = r.xbegin; x < r.xend; ++x) {
                if (x < dst->x || x >= dst->x + dst->width || y < dst->y
                    || y >= dst->y + dst->height)
                    continue;
                Dual2 dx { x + 0.5f, 1.0f, 0.0f }, dy { y + 0.5f, 0.0f, 1.0f }, ox, oy;
                xform_point(Minv, dx, dy, &ox, &oy);
                filtered_sample(*src, ox.v, oy.v, ox.dx, oy.dx, ox.dy, oy.dy, filter, wrap,
                                edgeclamp != 0, pel);
                unsigned char* dp = (unsigned char*)dst->pixels
                                    + int64_t(y - dst->y) * dst->ystride
                                    + int64_t(x - dst->x) * dst->xstride;
                for (int c = 0; c < nc; ++c)
                    orc_store_from_float(dp + c * des, dst->basetype, pel[c]);
            }
    });
    return 0;
}

// The exact=1 branch of fit (xform.cpp:962-1027): matrix and output window.
extern "C" void
orc_fit_exact_matrix(int src_full_w, int src_full_h, int fit_w, int fit_h, const char* fillmode,
                     float* out, int* resize_w, int* resize_h)
{
    float oldaspect = float(src_full_w) / src_full_h;
    float newaspect = float(fit_w) / fit_h;
    int rw = fit_w, rh = fit_h;
    float xoff = 0.0f, yoff = 0.0f, scale = 1.0f;
    std::string mode = fillmode ? fillmode : "letterbox";
    if (mode != "height" && mode != "width")
        mode = "letterbox";
    if (mode == "letterbox")
        mode = (newaspect >= oldaspect) ? "height" : "width";
    if (mode == "height") {
        rw    = int(rh * oldaspect + 0.5f);
        scale = float(fit_h) / float(src_full_h);
        xoff  = float(fit_w - scale * src_full_w) / 2.0f;
    } else {
        rh    = int(rw / oldaspect + 0.5f);
        scale = float(fit_w) / float(src_full_w);
        yoff  = float(fit_h - scale * src_full_h) / 2.0f;
    }
    float M[9] = { scale, 0.0f, 0.0f, 0.0f, scale, 0.0f, xoff, yoff, 1.0f };
    memcpy(out, M, sizeof(M));
    *resize_w = rw;
    *resize_h = rh;
}

// st_warp_<D,S,ST> (xform.cpp:1834-1925).  The filter is given resolved.
// Source pixels outside the data window read as black (the iterator's default
// wrap); their weights still count in the normalisation.
extern "C" int
orc_st_warp(const orc_image* dst, const orc_image* src, const orc_image* st, int chan_s,
            int chan_t, int flip_s, int flip_t, const char* filtername, float fw, float fh,
            orc_roi roi, int chbegin, int chend, int nthreads, char* err, int errlen)
{
    OrcFilter2D filter;
    if (!orc_make_filter2d(filtername && filtername[0] ? filtername : "lanczos3", fw, fh,
                           &filter)) {
        snprintf(err, errlen, "Filter \"%s\" not recognized", filtername ? filtername : "");
        return 1;
    }
    const int src_width = src->full_width, src_height = src->full_height;
    const float xscale = float(dst->full_width) / src_width;
    const float yscale = float(dst->full_height) / src_height;
    const int xbegin = src->x, xend = src->x + src->width;
    const int ybegin = src->y, yend = src->y + src->height;
    const int filterrad_x = (int)ceilf(filter.w / 2.0f / xscale);
    const int filterrad_y = (int)ceilf(filter.h / 2.0f / yscale);
    const int ses = esize(src->basetype), des = esize(dst->basetype), tes = esize(st->basetype);
    // the ROI is intersected with the ST image's window (check_st_warp_args)
    roi.xbegin = std::max(roi.xbegin, st->x), roi.xend = std::min(roi.xend, st->x + st->width);
    roi.ybegin = std::max(roi.ybegin, st->y), roi.yend = std::min(roi.yend, st->y + st->height);
    orc_parallel_bands(roi, nthreads, [&](orc_roi r) {
        float acc[64];
        for (int y = r.ybegin; y < r.yend; ++y)
            for (int x = r.xbegin; x < r.xend; ++x) {
                if (x < dst->x || x >= dst->x + dst->width || y < dst->y
                    || y >= dst->y + dst->height)
                    continue;
                const unsigned char* tp = (const unsigned char*)st->pixels
                                          + int64_t(y - st->y) * st->ystride
                                          + int64_t(x - st->x) * st->xstride;
                float src_s = orc_load_as_float(tp + chan_s * tes, st->basetype);
                float src_t = orc_load_as_float(tp + chan_t * tes, st->basetype);
                if (flip_s)
                    src_s = 1.0f - src_s;
                if (flip_t)
                    src_t = 1.0f - src_t;
                const float src_x = src_s * src_width;
                const float src_y = src_t * src_height;
                const int x_min = clampi((int)floorf(src_x - filterrad_x), xbegin, xend);
                const int x_max = clampi((int)ceilf(src_x + filterrad_x), xbegin, xend);
                const int y_min = clampi((int)floorf(src_y - filterrad_y), ybegin, yend);
                const int y_max = clampi((int)ceilf(src_y + filterrad_y), ybegin, yend);
                for (int c = chbegin; c < chend; ++c)
                    acc[c] = 0.0f;
                float total_weight = 0.0f;
                for (int sy = y_min; sy < y_max + 1; ++sy)
                    for (int sx = x_min; sx < x_max + 1; ++sx) {
                        const float weight = orc_f2d_eval(filter, sx - src_x + 0.5f,
                                                          sy - src_y + 0.5f);
                        total_weight += weight;
                        const unsigned char* p = wrapped_pixel(*src, sx, sy, 1);
                        for (int c = chbegin; c < chend; ++c)
                            acc[c] += (p ? orc_load_as_float(p + c * ses, src->basetype) : 0.0f)
                                      * weight;
                    }
                unsigned char* dp = (unsigned char*)dst->pixels
                                    + int64_t(y - dst->y) * dst->ystride
                                    + int64_t(x - dst->x) * dst->xstride;
                for (int c = chbegin; c < chend; ++c)
                    orc_store_from_float(dp + c * des, dst->basetype,
                                         total_weight > 0.0f ? acc[c] / total_weight : 0.0f);
            }
    });
    return 0;
}
