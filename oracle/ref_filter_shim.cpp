// ref_filter_shim.cpp -- extern "C" access to the REFERENCE's own Filter1D /
// Filter2D object code (src/libutil/filter.cpp compiled unmodified into
// oracle/_ref/libref_filter.so by oracle/Makefile).  TEST INFRASTRUCTURE ONLY:
// used to prove the oracle's and the product's weight functions bit-exact
// against the reference implementation.
#include <OpenImageIO/filter.h>

#include <cmath>
#include <cstring>

using namespace OIIO;

extern "C" {

int
ref_filter2d_count()
{
    return Filter2D::num_filters();
}

const char*
ref_filter2d_name(int i)
{
    return Filter2D::get_filterdesc(i).name;
}

float
ref_filter2d_default_width(int i)
{
    return Filter2D::get_filterdesc(i).width;
}

int
ref_filter1d_count()
{
    return Filter1D::num_filters();
}

const char*
ref_filter1d_name(int i)
{
    return Filter1D::get_filterdesc(i).name;
}

// Evaluate n samples.  mode: 0 = xfilt, 1 = yfilt, 2 = operator()(x,y).
// Returns 0 on success, -1 if the name is not recognised; *separable and the
// constructed width/height are reported too.
int
ref_filter2d_eval(const char* name, float w, float h, int mode, int n,
                  const float* x, const float* y, float* out, int* separable,
                  float* width_out, float* height_out)
{
    auto f = Filter2D::create_shared(name, w, h);
    if (!f)
        return -1;
    if (separable)
        *separable = f->separable() ? 1 : 0;
    if (width_out)
        *width_out = f->width();
    if (height_out)
        *height_out = f->height();
    for (int i = 0; i < n; ++i) {
        if (mode == 0)
            out[i] = f->xfilt(x[i]);
        else if (mode == 1)
            out[i] = f->yfilt(y[i]);
        else
            out[i] = (*f)(x[i], y[i]);
    }
    return 0;
}

int
ref_filter1d_eval(const char* name, float w, int n, const float* x, float* out,
                  float* width_out)
{
    auto f = Filter1D::create_shared(name, w);
    if (!f)
        return -1;
    if (width_out)
        *width_out = f->width();
    for (int i = 0; i < n; ++i)
        out[i] = (*f)(x[i]);
    return 0;
}

}  // extern "C"
