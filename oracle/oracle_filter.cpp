// oracle_filter.cpp -- CPU restatement of the reference's reconstruction
// filters.  TEST INFRASTRUCTURE ONLY (see oracle.h).
//
// Follows /root/reference/src/libutil/filter.cpp (OpenImageIO 3.2.0.2-dev):
//   box :65-77, triangle :94-132, gaussian :146-150, sharp-gaussian :193-197,
//   catmull-rom :233-262, blackman-harris :311-337, sinc :381-394,
//   lanczos3 :444-474, radial-lanczos3 :516-521, mitchell :558-578,
//   b-spline :625-641, disk :689-694, cubic/keys/simon/rifman :711-726,
//   tables :841-860 / :942-963, factories :897-930 / :1000-1039,
// and fast_exp from src/include/OpenImageIO/fmath.h:1938-1978.
//
// Must be compiled with -ffp-contract=off: the reference build does not fuse
// multiply-adds on baseline x86-64 and the polynomial below is sensitive.
#include "oracle.h"

#include <cmath>
#include <cstring>
#include <limits>

namespace {

enum Kind {
    K_BOX,
    K_TRIANGLE,
    K_GAUSSIAN,
    K_SHARPGAUSS,
    K_CATROM,
    K_BLACKMANHARRIS,
    K_SINC,
    K_LANCZOS3,
    K_RADIAL_LANCZOS3,
    K_MITCHELL,
    K_BSPLINE,
    K_DISK,
    K_CUBIC,
    K_KEYS,
    K_SIMON,
    K_RIFMAN,
    K_NONE
};

struct Desc {
    const char* name;
    float width;
    Kind kind;
};

// filter2d_list order, filter.cpp:942-963 ("nuke-lanczos6" instantiates the
// separable lanczos3 class, :1023-1025).
const Desc k2d[] = {
    { "box", 1, K_BOX },
    { "triangle", 2, K_TRIANGLE },
    { "gaussian", 3, K_GAUSSIAN },
    { "sharp-gaussian", 2, K_SHARPGAUSS },
    { "catmull-rom", 4, K_CATROM },
    { "blackman-harris", 3, K_BLACKMANHARRIS },
    { "sinc", 4, K_SINC },
    { "lanczos3", 6, K_LANCZOS3 },
    { "radial-lanczos3", 6, K_RADIAL_LANCZOS3 },
    { "nuke-lanczos6", 6, K_LANCZOS3 },
    { "mitchell", 4, K_MITCHELL },
    { "bspline", 4, K_BSPLINE },
    { "disk", 1, K_DISK },
    { "cubic", 4, K_CUBIC },
    { "keys", 4, K_KEYS },
    { "simon", 4, K_SIMON },
    { "rifman", 4, K_RIFMAN },
};
const int n2d = int(sizeof(k2d) / sizeof(k2d[0]));

// Names accepted by Filter2D::create (filter.cpp:1000-1039) -- a superset of
// the table (aliases).
Kind
kind_from_create_name(const char* n)
{
    if (!n)
        return K_NONE;
    if (!strcmp(n, "box"))
        return K_BOX;
    if (!strcmp(n, "triangle"))
        return K_TRIANGLE;
    if (!strcmp(n, "gaussian"))
        return K_GAUSSIAN;
    if (!strcmp(n, "sharp-gaussian"))
        return K_SHARPGAUSS;
    if (!strcmp(n, "catmull-rom") || !strcmp(n, "catrom"))
        return K_CATROM;
    if (!strcmp(n, "blackman-harris"))
        return K_BLACKMANHARRIS;
    if (!strcmp(n, "sinc"))
        return K_SINC;
    if (!strcmp(n, "lanczos3") || !strcmp(n, "lanczos")
        || !strcmp(n, "nuke-lanczos6"))
        return K_LANCZOS3;
    if (!strcmp(n, "radial-lanczos3") || !strcmp(n, "radial-lanczos"))
        return K_RADIAL_LANCZOS3;
    if (!strcmp(n, "mitchell"))
        return K_MITCHELL;
    if (!strcmp(n, "b-spline") || !strcmp(n, "bspline"))
        return K_BSPLINE;
    if (!strcmp(n, "disk"))
        return K_DISK;
    if (!strcmp(n, "cubic"))
        return K_CUBIC;
    if (!strcmp(n, "keys"))
        return K_KEYS;
    if (!strcmp(n, "simon"))
        return K_SIMON;
    if (!strcmp(n, "rifman"))
        return K_RIFMAN;
    return K_NONE;
}

float
default_width(Kind k)
{
    switch (k) {
    case K_BOX: return 1.0f;
    case K_TRIANGLE: return 2.0f;
    case K_GAUSSIAN: return 3.0f;
    case K_SHARPGAUSS: return 2.0f;
    case K_BLACKMANHARRIS: return 3.0f;
    case K_LANCZOS3:
    case K_RADIAL_LANCZOS3: return 6.0f;
    case K_DISK: return 1.0f;
    default: return 4.0f;
    }
}

// fmath.h:1938-1971 (scalar, non-CUDA branch) and :1975-1978.
float
orc_fast_exp2(float xval)
{
    float x = xval;
    if (!(-126.0f <= x))
        x = -126.0f;
    if (x > 126.0f)
        x = 126.0f;
    int m = int(x);
    x -= m;
    x       = 1.0f - (1.0f - x);
    float r = 1.33336498402e-3f;
    r       = x * r + 9.810352697968e-3f;
    r       = x * r + 5.551834031939e-2f;
    r       = x * r + 0.2401793301105f;
    r       = x * r + 0.693144857883f;
    r       = x * r + 1.0f;
    unsigned int bits;
    memcpy(&bits, &r, 4);
    bits += (unsigned(m) << 23);
    memcpy(&r, &bits, 4);
    return r;
}

float
orc_fast_exp(float x)
{
    return orc_fast_exp2(x * float(1 / M_LN2));
}

float
tri1d(float x)
{
    x = fabsf(x);
    return (x < 1.0f) ? (1.0f - x) : 0.0f;
}
float
gauss1d(float x)
{
    x = fabsf(x);
    return (x < 1.0f) ? orc_fast_exp(-2.0f * (x * x)) : 0.0f;
}
float
sharpgauss1d(float x)
{
    x = fabsf(x);
    return (x < 1.0f) ? orc_fast_exp(-4.0f * (x * x)) : 0.0f;
}
float
catrom1d(float x)
{
    x        = fabsf(x);
    float x2 = x * x;
    float x3 = x * x2;
    return (x >= 2.0f) ? 0.0f
                       : ((x < 1.0f) ? (3.0f * x3 - 5.0f * x2 + 2.0f)
                                     : (-x3 + 5.0f * x2 - 8.0f * x + 4.0f));
}
float
bh1d(float x)
{
    if (x < -1.0f || x > 1.0f)
        return 0.0f;
    x              = (x + 1.0f) * 0.5f;
    const float A0 = 0.35875f, A1 = -0.48829f, A2 = 0.14128f, A3 = -0.01168f;
    const float pi = float(M_PI);
    float c2       = cosf(2.f * pi * x);
    float c4       = 2.0f * c2 * c2 - 1.0f;
    float c6       = c2 * (2.0f * c4 - 1.0f);
    return A0 + A1 * c2 + A2 * c4 + A3 * c6;
}
float
sinc1d(float x, float rad)
{
    x = fabsf(x);
    if (x > rad)
        return 0.0f;
    const float pi = float(M_PI);
    return (x < 0.0001f) ? 1.0f : sinf(pi * x) / (pi * x);
}
float
lanczos3(float x)
{
    const float a    = 3.0f;
    const float ainv = 1.0f / a;
    const float pi   = float(M_PI);
    x                = fabsf(x);
    if (x > a)
        return 0.0f;
    if (x < 0.0001f)
        return 1.0f;
    float s1 = sinf(x * ainv * pi);
    float s3 = (-4.0f * s1 * s1 + 3.0f) * s1;
    return a / (x * x * (pi * pi)) * s1 * s3;
}
float
mitchell1d(float x)
{
    x = fabsf(x);
    if (x > 1.0f)
        return 0.0f;
    x *= 2.0f;
    float x2          = x * x;
    const float B     = 1.0f / 3.0f;
    const float C     = 1.0f / 3.0f;
    const float SIXTH = 1.0f / 6.0f;
    if (x >= 1.0f)
        return ((-B - 6.0f * C) * x * x2 + (6.0f * B + 30.0f * C) * x2
                + (-12.0f * B - 48.0f * C) * x + (8.0f * B + 24.0f * C))
               * SIXTH;
    else
        return ((12.0f - 9.0f * B - 6.0f * C) * x * x2
                + (-18.0f + 12.0f * B + 6.0f * C) * x2 + (6.0f - 2.0f * B))
               * SIXTH;
}
float
bspline1d(float x)
{
    x = fabsf(x);
    if (x <= 1.0f) {
        float t = 1.0f - x;
        return 0.5f * t * (t * (1.0f - t) + 1.0f) + 1.0f / 6.0f;
    } else if (x < 2.0f) {
        float t = 2.0f - x;
        return t * t * t / 6.0f;
    }
    return 0.0f;
}
float
cubic1d(float x, float a)
{
    x = fabsf(x);
    if (x > 1.0f)
        return 0.0f;
    x *= 2.0f;
    if (x >= 1.0f)
        return a * (x * (x * (x - 5.0f) + 8.0f) - 4.0f);
    else
        return x * x * ((a + 2.0f) * x - (a + 3.0f)) + 1.0f;
}

float
cubic_a(Kind k)
{
    return k == K_KEYS ? -0.5f : k == K_SIMON ? -0.75f : k == K_RIFMAN ? -1.0f : 0.0f;
}

// One axis of a separable 2-D filter of extent `w` along that axis
// (the xfilt/yfilt members of each Filter*2D class).
float
axis_eval(Kind k, float w, float x)
{
    switch (k) {
    case K_BOX: return fabsf(x) <= w * 0.5f ? 1.0f : 0.0f;
    case K_TRIANGLE: return tri1d(x * (2.0f / w));
    case K_GAUSSIAN: return gauss1d(x * (2.0f / w));
    case K_SHARPGAUSS: return sharpgauss1d(x * (2.0f / w));
    case K_CATROM: return catrom1d(x * (4.0f / w));
    case K_BLACKMANHARRIS: return bh1d(x * (2.0f / w));
    case K_SINC: return sinc1d(x, w / 2.0f);
    case K_LANCZOS3:
    case K_RADIAL_LANCZOS3: return lanczos3(x * (6.0f / w));
    case K_MITCHELL: return mitchell1d(x * (2.0f / w));
    case K_BSPLINE: return bspline1d(x * (4.0f / w));
    case K_CUBIC:
    case K_KEYS:
    case K_SIMON:
    case K_RIFMAN: return cubic1d(x * (2.0f / w), cubic_a(k));
    default: return 0.0f;
    }
}

float
disk2d(float w, float h, float x, float y)
{
    x /= (w * 0.5f);
    y /= (h * 0.5f);
    return ((x * x + y * y) < 1.0f) ? 1.0f : 0.0f;
}

}  // namespace

// ---- internal C++ entry used by oracle_resize.cpp ------------------------
struct OrcFilter2D {
    Kind kind;
    float w, h;
};

bool
orc_make_filter2d(const char* name, float w, float h, OrcFilter2D* f)
{
    Kind k = kind_from_create_name(name);
    if (k == K_NONE)
        return false;
    if (h <= 0.0f)  // Filter2D::create, filter.cpp:1003-1004
        h = w;
    float dw = default_width(k);
    f->kind  = k;
    f->w     = w > 0.0f ? w : dw;
    f->h     = h > 0.0f ? h : dw;
    return true;
}
bool
orc_f2d_separable(const OrcFilter2D& f)
{
    return f.kind != K_RADIAL_LANCZOS3 && f.kind != K_DISK;
}
float
orc_f2d_eval(const OrcFilter2D& f, float x, float y)
{
    if (f.kind == K_DISK)
        return disk2d(f.w, f.h, x, y);
    if (f.kind == K_RADIAL_LANCZOS3) {
        x *= 6.0f / f.w;
        y *= 6.0f / f.h;
        return lanczos3(sqrtf(x * x + y * y));
    }
    return axis_eval(f.kind, f.w, x) * axis_eval(f.kind, f.h, y);
}
float
orc_f2d_xfilt(const OrcFilter2D& f, float x)
{
    if (f.kind == K_DISK)  // base-class default: (*this)(x, 0), filter.h:113
        return disk2d(f.w, f.h, x, 0.0f);
    return axis_eval(f.kind, f.w, x);
}
float
orc_f2d_yfilt(const OrcFilter2D& f, float y)
{
    if (f.kind == K_DISK)
        return disk2d(f.w, f.h, 0.0f, y);
    return axis_eval(f.kind, f.h, y);
}

extern "C" {

int
orc_filter_count(void)
{
    return n2d;
}
const char*
orc_filter_name(int i)
{
    return (i >= 0 && i < n2d) ? k2d[i].name : nullptr;
}
float
orc_filter_default_width(int i)
{
    return (i >= 0 && i < n2d) ? k2d[i].width : 0.0f;
}
int
orc_filter_lookup(const char* name)
{
    if (!name)
        return -1;
    for (int i = 0; i < n2d; ++i)
        if (!strcmp(k2d[i].name, name))
            return i;
    return -1;
}
int
orc_filter_separable(const char* name, float w, float h)
{
    OrcFilter2D f;
    if (!orc_make_filter2d(name, w, h, &f))
        return -1;
    return orc_f2d_separable(f) ? 1 : 0;
}
float
orc_filter_xfilt(const char* name, float w, float h, float x)
{
    OrcFilter2D f;
    if (!orc_make_filter2d(name, w, h, &f))
        return std::numeric_limits<float>::quiet_NaN();
    return orc_f2d_xfilt(f, x);
}
float
orc_filter_yfilt(const char* name, float w, float h, float y)
{
    OrcFilter2D f;
    if (!orc_make_filter2d(name, w, h, &f))
        return std::numeric_limits<float>::quiet_NaN();
    return orc_f2d_yfilt(f, y);
}
float
orc_filter_eval2d(const char* name, float w, float h, float x, float y)
{
    OrcFilter2D f;
    if (!orc_make_filter2d(name, w, h, &f))
        return std::numeric_limits<float>::quiet_NaN();
    return orc_f2d_eval(f, x, y);
}
// Filter1D::create, filter.cpp:897-930.  The 1-D catmull-rom keeps a fixed
// reported width of 4 but scales its argument by 4/width (:235-238).
float
orc_filter1d_eval(const char* name, float w, float x)
{
    Kind k = kind_from_create_name(name);
    if (k == K_NONE || k == K_RADIAL_LANCZOS3 || k == K_DISK)
        return std::numeric_limits<float>::quiet_NaN();
    float ww = w > 0.0f ? w : default_width(k);
    return axis_eval(k, ww, x);
}

}  // extern "C"
