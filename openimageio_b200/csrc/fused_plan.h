// fused_plan.h -- host-side tiling and table layout for resize_fused_kernel.
#pragma once

#include "axis_plan.h"
#include "kernels.h"

#include <vector>

namespace b2r {

struct FusedPlan {
    std::vector<Strip> strips;
    std::vector<Band> bands;
    bool expand = false;  // tables are for resize_expand_kernel
    int w2 = 1;           // 2: strips sized for the stream kernel's 512 V threads
    int ngroups = 0;      // expand: groups of four dst columns in the ROI
    bool vrow_linear = false;  // expand: V-row layout (see expand_kernel.cuh)
    // H pass
    int ht = 0;
    std::vector<int32_t> hfirst;
    std::vector<float> hw;
    std::vector<int32_t> hmix;  // stream kernel: mixed-sign merged taps per pair
    // V pass
    int vk = 0;  // ring size, 0 = gather
    std::vector<float> vring;
    std::vector<float> vstream;  // stream kernel: rotated ring weights + completion counts
    std::vector<int32_t> vlast;
    int vt = 0;
    std::vector<int32_t> vfirst;
    std::vector<float> vw;
    int max_band_rows = 0;
    int max_band_dy   = 0;
    int max_nvec      = 0;
    bool interior_zero = false;
};

// gx / gy: gather tables of the two axes over the dst ROI.  cta_slots: how
// many CTAs the device keeps resident (SMs x CTAs per SM); the band count is
// chosen so the grid fills one wave.  Returns false when the shape is outside
// what the fused kernel instantiates (then the exact kernel runs).
// stream_sms > 0 (persistent stream kernel, one CTA per SM walking a static
// item list): the band count is searched for the shortest schedule of
// nstrips x nbands x nframes items over stream_sms CTAs instead.
bool
plan_fused(const AxisGather& gx, const AxisGather& gy, int nch, int src_w,
           int src_h, int cta_slots, FusedPlan* plan, int e0_align = 4, int w2 = 1,
           int stream_sms = 0, int nframes = 1);

}  // namespace b2r
