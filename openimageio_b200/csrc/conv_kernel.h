// conv_kernel.h -- host side of ImageBufAlgo::make_kernel
// (src/libOpenImageIO/imagebufalgo.cpp:864-957): a small float image of
// filter values that convolve / unsharp_mask slide over the source.
#pragma once

#include <string>
#include <vector>

namespace b2r {

struct ConvKernel {
    int w = 0, h = 0;         // always odd; window [-w/2, -w/2 + w) x [-h/2, ...)
    std::vector<float> v;     // row-major, h rows of w
    bool known = true;        // false: unknown name -> box + error, like the reference
    std::string error;        // `Unknown kernel "name" WxH`
};

// name: any Filter2D name, "binomial", "laplacian" (3x3 only); anything else
// yields a box and sets `known = false`.  width/height 0 = the filter's own.
ConvKernel
make_conv_kernel(const std::string& name, float width, float height, bool normalize = true);

}  // namespace b2r
