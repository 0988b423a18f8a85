// axis_plan.h -- per-axis weight / footprint tables for the resize kernels.
//
// The reference evaluates, for every destination index d of an axis,
//     s      = (d - dst_full_origin + 0.5f) * (1.0f / dst_full_size)
//     src_f  = src_full_origin + s * src_full_size
//     src_i  = floor(src_f),  frac = src_f - src_i
//     w_i    = filt(ratio * (i - rad - (frac - 0.5f))),  i in [0, 2*rad+1)
//     w_i   /= sum(w)   (when the sum is not zero)
// (src/libOpenImageIO/imagebufalgo_xform.cpp:616-661 for x, :714-730 for y).
// AxisTaps reproduces exactly that, in float32, same expression order; it is
// what the exact kernel consumes.  AxisGather is derived from it for the
// fused kernel: WrapClamp addressing (imagebuf.cpp:3283-3322) resolved per
// tap, zero weights and black taps dropped, duplicates produced by clamping
// merged, so every destination index reads one contiguous run of real source
// indices.
#pragma once

#include <cstdint>
#include <vector>

namespace b2r {

class Filter2D;

// Geometry of one axis.
struct AxisGeom {
    int dst_begin, dst_end;        // ROI range on this axis (dst coords)
    int dst_full_origin, dst_full_size;
    int src_full_origin, src_full_size;
    int src_data_begin, src_data_end;  // data window [begin, end)
};

// The reference's "virtual" taps: window [src_i - rad, src_i + rad] per index.
struct AxisTaps {
    int rad   = 0;
    int ntaps = 0;               // 2*rad+1
    float ratio = 1.0f;          // dst_full_size / src_full_size
    std::vector<int32_t> center; // src_i per dst index   [n]
    std::vector<float> frac;     // frac per dst index    [n]  (non-separable)
    std::vector<float> w;        // normalised weights    [n * ntaps]
    std::vector<float> wsum;     // sum of the normalised weights, as the
                                 // reference re-adds them (:740-742)  [n]
    int n() const { return int(center.size()); }
};

// is_y selects yfilt (true) or xfilt (false).  filterrad follows the
// reference: filter->width()/2 for BOTH axes (xform.cpp:626-631).
AxisTaps
build_axis_taps(const AxisGeom& g, const Filter2D& filter, bool is_y);

// Contiguous-run gather form.  Offsets are relative to src_data_begin.
struct AxisGather {
    int n     = 0;
    int maxcount = 0;
    std::vector<int32_t> first;  // first real source index          [n]
    std::vector<int32_t> count;  // run length (0 = output is zero)  [n]
    std::vector<int32_t> woff;   // offset of the run's weights in w [n]
    std::vector<float> w;
    // parallel to w: 1 when the weight is the sum of clamped taps of BOTH signs
    // (an Inf under it is NaN in the reference: w1*Inf + w2*Inf)
    std::vector<uint8_t> mixed;
    bool interior_zero = false;  // some run contains a zero weight
};

// Valid when the source data window lies inside its full window on this
// axis (then WrapClamp addressing is separable); returns false otherwise.
// With overscan on either axis (data window beyond the display window) only
// ranges whose contributing taps ALL lie inside the data window are accepted
// (pass the other axis' axis_overscan() as `other_axis_overscan`).
bool
axis_overscan(const AxisGeom& g);
bool
build_axis_gather(const AxisGeom& g, const AxisTaps& taps, AxisGather* out,
                  bool other_axis_overscan = false);

}  // namespace b2r
