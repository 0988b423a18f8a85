// recon_filter.cpp -- table-driven reconstruction filters (host, float32).
//
// Behavioural contract: values, default widths, name/alias rules and the
// FilterDesc tables of OpenImageIO's src/libutil/filter.cpp (profiles :65-726,
// filter1d_list :841-860, filter2d_list :942-963, create() alias rules
// :897-930 and :1000-1039).  fast_exp follows
// src/include/OpenImageIO/fmath.h:1938-1978.  Every profile keeps the
// reference's float expression order; build with -ffp-contract=off.
#include "recon_filter.h"

#include <cmath>
#include <cstdint>
#include <cstring>

namespace b2r {
namespace detail {

enum class Scale {
    BoxEdge,   // compare |x| against extent/2, no scaling
    Two,       // x * (2/extent)
    Four,      // x * (4/extent)
    Six,       // x * (6/extent)
    SincRad,   // profile(x, extent/2)
    Disk,      // 2-D only
    RadialL3   // 2-D only, lanczos3 of the scaled radius
};

struct Family {
    const char* objname;  // what name() reports
    float defwidth;       // used when width <= 0
    Scale scale;
    float (*profile)(float, float);
    float param;
};

static inline float
exp2_poly(float xval)
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
    uint32_t bits;
    std::memcpy(&bits, &r, 4);
    bits += (uint32_t(m) << 23);
    std::memcpy(&r, &bits, 4);
    return r;
}
static inline float
exp_poly(float x)
{
    return exp2_poly(x * float(1 / M_LN2));
}

static float
p_box(float, float)
{
    return 1.0f;  // edge test done by the caller
}
static float
p_tent(float x, float)
{
    x = fabsf(x);
    return (x < 1.0f) ? (1.0f - x) : 0.0f;
}
static float
p_gauss(float x, float k)
{
    x = fabsf(x);
    return (x < 1.0f) ? exp_poly(k * (x * x)) : 0.0f;
}
static float
p_catrom(float x, float)
{
    x        = fabsf(x);
    float x2 = x * x;
    float x3 = x * x2;
    if (x >= 2.0f)
        return 0.0f;
    return (x < 1.0f) ? (3.0f * x3 - 5.0f * x2 + 2.0f)
                      : (-x3 + 5.0f * x2 - 8.0f * x + 4.0f);
}
static float
p_blackman_harris(float x, float)
{
    if (x < -1.0f || x > 1.0f)
        return 0.0f;
    x              = (x + 1.0f) * 0.5f;
    const float pi = float(M_PI);
    float c1       = cosf(2.f * pi * x);
    float c2       = 2.0f * c1 * c1 - 1.0f;
    float c3       = c1 * (2.0f * c2 - 1.0f);
    return 0.35875f + -0.48829f * c1 + 0.14128f * c2 + -0.01168f * c3;
}
static float
p_sinc(float x, float rad)
{
    x = fabsf(x);
    if (x > rad)
        return 0.0f;
    const float pi = float(M_PI);
    return (x < 0.0001f) ? 1.0f : sinf(pi * x) / (pi * x);
}
static float
p_lanczos3(float x, float)
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
static float
p_mitchell(float x, float)
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
    return ((12.0f - 9.0f * B - 6.0f * C) * x * x2
            + (-18.0f + 12.0f * B + 6.0f * C) * x2 + (6.0f - 2.0f * B))
           * SIXTH;
}
static float
p_bspline(float x, float)
{
    x = fabsf(x);
    if (x <= 1.0f) {
        float t = 1.0f - x;
        return 0.5f * t * (t * (1.0f - t) + 1.0f) + 1.0f / 6.0f;
    }
    if (x < 2.0f) {
        float t = 2.0f - x;
        return t * t * t / 6.0f;
    }
    return 0.0f;
}
static float
p_cubic(float x, float a)
{
    x = fabsf(x);
    if (x > 1.0f)
        return 0.0f;
    x *= 2.0f;
    if (x >= 1.0f)
        return a * (x * (x * (x - 5.0f) + 8.0f) - 4.0f);
    return x * x * ((a + 2.0f) * x - (a + 3.0f)) + 1.0f;
}

// One entry per distinct object the factories can build.
static const Family F_BOX      = { "box", 1.0f, Scale::BoxEdge, p_box, 0.0f };
static const Family F_TRIANGLE = { "triangle", 2.0f, Scale::Two, p_tent, 0.0f };
static const Family F_GAUSS    = { "gaussian", 3.0f, Scale::Two, p_gauss, -2.0f };
// the sharp gaussian reports "gaussian" as its name() (filter.cpp:204,232)
static const Family F_SHARPG = { "gaussian", 2.0f, Scale::Two, p_gauss, -4.0f };
static const Family F_CATROM = { "catmull-rom", 4.0f, Scale::Four, p_catrom, 0.0f };
static const Family F_BH = { "blackman-harris", 3.0f, Scale::Two,
                             p_blackman_harris, 0.0f };
static const Family F_SINC     = { "sinc", 4.0f, Scale::SincRad, p_sinc, 0.0f };
static const Family F_LANCZOS3 = { "lanczos3", 6.0f, Scale::Six, p_lanczos3, 0.0f };
static const Family F_RADIAL = { "radial-lanczos3", 6.0f, Scale::RadialL3,
                                 p_lanczos3, 0.0f };
static const Family F_MITCHELL = { "mitchell", 4.0f, Scale::Two, p_mitchell, 0.0f };
static const Family F_BSPLINE  = { "b-spline", 4.0f, Scale::Four, p_bspline, 0.0f };
static const Family F_DISK     = { "disk", 1.0f, Scale::Disk, p_box, 0.0f };
static const Family F_CUBIC    = { "cubic", 4.0f, Scale::Two, p_cubic, 0.0f };
static const Family F_KEYS     = { "keys", 4.0f, Scale::Two, p_cubic, -0.5f };
static const Family F_SIMON    = { "simon", 4.0f, Scale::Two, p_cubic, -0.75f };
static const Family F_RIFMAN   = { "rifman", 4.0f, Scale::Two, p_cubic, -1.0f };

struct Alias {
    const char* key;
    const Family* fam;
    bool in1d;  // accepted by Filter1D::create
};
static const Alias aliases[] = {
    { "box", &F_BOX, true },
    { "triangle", &F_TRIANGLE, true },
    { "gaussian", &F_GAUSS, true },
    { "sharp-gaussian", &F_SHARPG, true },
    { "catmull-rom", &F_CATROM, true },
    { "catrom", &F_CATROM, true },
    { "blackman-harris", &F_BH, true },
    { "sinc", &F_SINC, true },
    { "lanczos3", &F_LANCZOS3, true },
    { "lanczos", &F_LANCZOS3, true },
    { "nuke-lanczos6", &F_LANCZOS3, true },
    { "radial-lanczos3", &F_RADIAL, false },
    { "radial-lanczos", &F_RADIAL, false },
    { "mitchell", &F_MITCHELL, true },
    { "b-spline", &F_BSPLINE, true },
    { "bspline", &F_BSPLINE, true },
    { "disk", &F_DISK, false },
    { "cubic", &F_CUBIC, true },
    { "keys", &F_KEYS, true },
    { "simon", &F_SIMON, true },
    { "rifman", &F_RIFMAN, true },
};

static const Family*
find_family(std::string_view name, bool for1d)
{
    for (const Alias& a : aliases)
        if (name == a.key && (!for1d || a.in1d))
            return a.fam;
    return nullptr;
}

// profile of one axis of extent `ext`
static inline float
axis_value(const Family* f, float ext, float x)
{
    switch (f->scale) {
    case Scale::BoxEdge: return fabsf(x) <= ext * 0.5f ? 1.0f : 0.0f;
    case Scale::Two: return f->profile(x * (2.0f / ext), f->param);
    case Scale::Four: return f->profile(x * (4.0f / ext), f->param);
    case Scale::Six:
    case Scale::RadialL3: return f->profile(x * (6.0f / ext), f->param);
    case Scale::SincRad: return f->profile(x, ext / 2.0f);
    case Scale::Disk: break;
    }
    return 0.0f;
}

static inline float
disk_value(float w, float h, float x, float y)
{
    x /= (w * 0.5f);
    y /= (h * 0.5f);
    return ((x * x + y * y) < 1.0f) ? 1.0f : 0.0f;
}

}  // namespace detail

using detail::Family;
using detail::Scale;

// ---- descriptor tables ---------------------------------------------------
static const FilterDesc descs1d[] = {
    { "box", 1, 1, false, true, true },
    { "triangle", 1, 2, false, true, true },
    { "gaussian", 1, 3, false, true, true },
    { "sharp-gaussian", 1, 2, false, true, true },
    { "catmull-rom", 1, 4, false, true, true },
    { "blackman-harris", 1, 3, false, true, true },
    { "sinc", 1, 4, false, true, true },
    { "lanczos3", 1, 6, false, true, true },
    { "nuke-lanczos6", 1, 6, false, true, true },
    { "mitchell", 1, 4, false, true, true },
    { "bspline", 1, 4, false, true, true },
    { "cubic", 1, 4, false, true, true },
    { "keys", 1, 4, false, true, true },
    { "simon", 1, 4, false, true, true },
    { "rifman", 1, 4, false, true, true },
};
static const FilterDesc descs2d[] = {
    { "box", 2, 1, false, true, true },
    { "triangle", 2, 2, false, true, true },
    { "gaussian", 2, 3, false, true, true },
    { "sharp-gaussian", 2, 2, false, true, true },
    { "catmull-rom", 2, 4, false, true, true },
    { "blackman-harris", 2, 3, false, true, true },
    { "sinc", 2, 4, false, true, true },
    { "lanczos3", 2, 6, false, true, true },
    { "radial-lanczos3", 2, 6, false, true, false },
    { "nuke-lanczos6", 2, 6, false, true, false },
    { "mitchell", 2, 4, false, true, true },
    { "bspline", 2, 4, false, true, true },
    { "disk", 2, 1, false, true, false },
    { "cubic", 2, 4, false, true, true },
    { "keys", 2, 4, false, true, true },
    { "simon", 2, 4, false, true, true },
    { "rifman", 2, 4, false, true, true },
};

// ---- Filter1D --------------------------------------------------------------
float
Filter1D::operator()(float x) const
{
    return detail::axis_value(m_fam, m_ext, x);
}
std::string_view
Filter1D::name() const
{
    return m_fam->objname;
}
Filter1D*
Filter1D::create(std::string_view filtername, float width)
{
    const Family* f = detail::find_family(filtername, true);
    if (!f)
        return nullptr;
    float ext = width > 0.0f ? width : f->defwidth;
    // the 1-D catmull-rom always reports width 4 (filter.cpp:235-238)
    float reported = (f == &detail::F_CATROM) ? 4.0f : ext;
    return new Filter1D(f, reported, ext);
}
Filter1D::ref
Filter1D::create_shared(std::string_view filtername, float width)
{
    return ref(create(filtername, width),
               [](const Filter1D* p) { delete p; });
}
void
Filter1D::destroy(Filter1D* filt)
{
    delete filt;
}
int
Filter1D::num_filters()
{
    return int(sizeof(descs1d) / sizeof(descs1d[0]));
}
const FilterDesc&
Filter1D::get_filterdesc(int i)
{
    return descs1d[i];
}
void
Filter1D::get_filterdesc(int i, FilterDesc* d)
{
    *d = descs1d[i];
}

// ---- Filter2D --------------------------------------------------------------
bool
Filter2D::separable() const
{
    return m_fam->scale != Scale::Disk && m_fam->scale != Scale::RadialL3;
}
int
Filter2D::nonseparable_kind() const
{
    return m_fam->scale == Scale::Disk ? 1
                                       : (m_fam->scale == Scale::RadialL3 ? 2 : 0);
}
Filter2D::DeviceDesc
Filter2D::device_desc() const
{
    DeviceDesc d;
    float (*const profs[])(float, float)
        = { detail::p_box,  detail::p_tent,     detail::p_gauss,    detail::p_catrom,
            detail::p_blackman_harris, detail::p_sinc, detail::p_lanczos3,
            detail::p_mitchell, detail::p_bspline,  detail::p_cubic };
    d.profile = 0;
    for (int i = 0; i < int(sizeof(profs) / sizeof(profs[0])); ++i)
        if (m_fam->profile == profs[i])
            d.profile = i;
    switch (m_fam->scale) {
    case Scale::BoxEdge: d.scale = 0; break;
    case Scale::Two: d.scale = 1; break;
    case Scale::Four: d.scale = 2; break;
    case Scale::Six: d.scale = 3; break;
    case Scale::SincRad: d.scale = 4; break;
    case Scale::Disk: d.scale = 5; break;
    case Scale::RadialL3: d.scale = 6; break;
    }
    d.param = m_fam->param;
    d.w     = m_w;
    d.h     = m_h;
    return d;
}
float
Filter2D::operator()(float x, float y) const
{
    if (m_fam->scale == Scale::Disk)
        return detail::disk_value(m_w, m_h, x, y);
    if (m_fam->scale == Scale::RadialL3) {
        x *= 6.0f / m_w;
        y *= 6.0f / m_h;
        return m_fam->profile(sqrtf(x * x + y * y), 0.0f);
    }
    return detail::axis_value(m_fam, m_w, x) * detail::axis_value(m_fam, m_h, y);
}
float
Filter2D::xfilt(float x) const
{
    if (m_fam->scale == Scale::Disk)
        return detail::disk_value(m_w, m_h, x, 0.0f);
    return detail::axis_value(m_fam, m_w, x);
}
float
Filter2D::yfilt(float y) const
{
    if (m_fam->scale == Scale::Disk)
        return detail::disk_value(m_w, m_h, 0.0f, y);
    return detail::axis_value(m_fam, m_h, y);
}
std::string_view
Filter2D::name() const
{
    return m_fam->objname;
}
Filter2D*
Filter2D::create(std::string_view filtername, float width, float height)
{
    if (height <= 0.0f)
        height = width;
    const Family* f = detail::find_family(filtername, false);
    if (!f)
        return nullptr;
    return new Filter2D(f, width > 0.0f ? width : f->defwidth,
                        height > 0.0f ? height : f->defwidth);
}
Filter2D::ref
Filter2D::create_shared(std::string_view filtername, float width, float height)
{
    return ref(create(filtername, width, height),
               [](const Filter2D* p) { delete p; });
}
void
Filter2D::destroy(Filter2D* filt)
{
    delete filt;
}
int
Filter2D::num_filters()
{
    return int(sizeof(descs2d) / sizeof(descs2d[0]));
}
const FilterDesc&
Filter2D::get_filterdesc(int i)
{
    return descs2d[i];
}
void
Filter2D::get_filterdesc(int i, FilterDesc* d)
{
    *d = descs2d[i];
}

}  // namespace b2r
