// recon_filter.h -- reconstruction filters by name (host side).
//
// Keeps the API surface of OpenImageIO's Filter1D / Filter2D
// (src/include/OpenImageIO/filter.h:19-156) so code written against the
// reference compiles against this: FilterDesc, create / create_shared /
// destroy, num_filters / get_filterdesc, width / height / separable / xfilt /
// yfilt / operator() / name.  The implementation is table driven: one
// profile function per filter family plus an argument scale, instead of a
// class per filter.
//
// All arithmetic is float32 in the reference's expression order
// (src/libutil/filter.cpp) because resize weight tables are built from these
// values on the host and must match the reference to the bit.  Compile with
// -ffp-contract=off.
#pragma once

#include <memory>
#include <string>
#include <string_view>

namespace b2r {

struct FilterDesc {
    const char* name;
    int dim;
    float width;
    bool fixedwidth;
    bool scalable;
    bool separable;
};

namespace detail {
struct Family;  // profile function + argument-scale rule
}

class Filter1D {
public:
    using ref = std::shared_ptr<const Filter1D>;

    float width() const { return m_w; }
    float operator()(float x) const;
    std::string_view name() const;

    static ref create_shared(std::string_view filtername, float width);
    static Filter1D* create(std::string_view filtername, float width);
    static void destroy(Filter1D* filt);
    static int num_filters();
    static void get_filterdesc(int filternum, FilterDesc* filterdesc);
    static const FilterDesc& get_filterdesc(int filternum);

private:
    Filter1D(const detail::Family* fam, float w, float scale_extent)
        : m_fam(fam)
        , m_w(w)
        , m_ext(scale_extent)
    {
    }
    const detail::Family* m_fam;
    float m_w;    // reported width
    float m_ext;  // extent used to scale the argument
};

class Filter2D {
public:
    using ref = std::shared_ptr<const Filter2D>;

    float width() const { return m_w; }
    float height() const { return m_h; }
    bool separable() const;
    float operator()(float x, float y) const;
    float xfilt(float x) const;
    float yfilt(float y) const;
    std::string_view name() const;

    static ref create_shared(std::string_view filtername, float width,
                             float height);
    static Filter2D* create(std::string_view filtername, float width,
                            float height);
    static void destroy(Filter2D* filt);
    static void no_destroy(const Filter2D*) {}
    static int num_filters();
    static void get_filterdesc(int filternum, FilterDesc* filterdesc);
    static const FilterDesc& get_filterdesc(int filternum);

    // Device-side evaluation of the two non-separable filters needs to know
    // which one this is: 0 separable, 1 disk, 2 radial-lanczos3.
    int nonseparable_kind() const;

    // Everything a device-side evaluation of operator() needs (warp kernel):
    // profile 0 box, 1 tent, 2 gauss, 3 catrom, 4 blackman-harris, 5 sinc,
    // 6 lanczos3, 7 mitchell, 8 b-spline, 9 cubic; scale 0 box-edge, 1 x*2/w,
    // 2 x*4/w, 3 x*6/w, 4 sinc radius, 5 disk, 6 radial lanczos3.
    struct DeviceDesc {
        int profile, scale;
        float param, w, h;
    };
    DeviceDesc device_desc() const;

private:
    Filter2D(const detail::Family* fam, float w, float h)
        : m_fam(fam)
        , m_w(w)
        , m_h(h)
    {
    }
    const detail::Family* m_fam;
    float m_w, m_h;
};

}  // namespace b2r
