// conv_kernel.cpp -- see conv_kernel.h.  Float arithmetic in the reference's
// order (compile with -ffp-contract=off): the table is what the device
// multiplies the pixels by, so it has to match to the bit.
#include "conv_kernel.h"

#include "recon_filter.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <strings.h>

namespace b2r {

namespace {
// n-choose-k by the running product the reference uses (imagebufalgo.cpp:855-861)
float
binomial_coeff(int n, int k)
{
    float p = 1;
    for (int i = 1; i <= k; ++i)
        p *= float(n - (k - i)) / i;
    return p;
}
int
odd_extent(float v)
{
    return std::max(1, (int)ceilf(v)) | 1;
}
}  // namespace

ConvKernel
make_conv_kernel(const std::string& name, float width, float height, bool normalize)
{
    ConvKernel K;
    Filter2D::ref filter = Filter2D::create_shared(name, width, height);
    if (filter) {
        if (width == 0.0f)
            width = filter->width();
        if (height == 0.0f)
            height = filter->height();
    } else {
        if (width == 0.0f)
            width = 4.0f;
        if (height == 0.0f)
            height = 4.0f;
    }
    K.w = odd_extent(width);
    K.h = odd_extent(height);
    K.v.assign((size_t)K.w * K.h, 0.0f);
    const int x0 = -K.w / 2, y0 = -K.h / 2;
    if (filter) {
        for (int j = 0; j < K.h; ++j)
            for (int i = 0; i < K.w; ++i)
                K.v[(size_t)j * K.w + i] = (*filter)(float(x0 + i), float(y0 + j));
    } else if (name == "binomial") {
        std::vector<float> row((size_t)K.w, 0.0f), col((size_t)K.h, 0.0f);
        for (int i = 0; i < width; ++i)
            row[i] = binomial_coeff(int(width - 1), i);
        for (int i = 0; i < height; ++i)
            col[i] = binomial_coeff(int(height - 1), i);
        for (int j = 0; j < K.h; ++j)
            for (int i = 0; i < K.w; ++i)
                K.v[(size_t)j * K.w + i] = row[i] * col[j] * 1.0f;
    } else if (!strcasecmp(name.c_str(), "laplacian") && K.w == 3 && K.h == 3) {
        const float lap[9] = { 0, 1, 0, 1, -4, 1, 0, 1, 0 };
        K.v.assign(lap, lap + 9);
        normalize = false;  // sums to zero
    } else {
        const float val = normalize ? 1.0f / (K.w * K.h * 1) : 1.0f;
        std::fill(K.v.begin(), K.v.end(), val);
        K.known = false;
        char msg[256];
        snprintf(msg, sizeof(msg), "Unknown kernel \"%s\" %gx%g", name.c_str(), (double)width,
                 (double)height);
        K.error = msg;
        return K;
    }
    if (normalize) {
        float sum = 0;
        for (float f : K.v)
            sum += f;
        if (sum != 0.0f)
            for (float& f : K.v)
                f = f / sum;
    }
    return K;
}

}  // namespace b2r
