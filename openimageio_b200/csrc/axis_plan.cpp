// axis_plan.cpp -- host-side weight tables (float32, reference expression
// order; compile with -ffp-contract=off).  See axis_plan.h.
#include "axis_plan.h"

#include "recon_filter.h"

#include <algorithm>
#include <cmath>

namespace b2r {

AxisTaps
build_axis_taps(const AxisGeom& g, const Filter2D& filter, bool is_y)
{
    AxisTaps t;
    const float srcf0 = float(g.src_full_origin);
    const float srcfs = float(g.src_full_size);
    const float dstf0 = float(g.dst_full_origin);
    const float dstfs = float(g.dst_full_size);
    t.ratio           = dstfs / srcfs;
    const float dstpixel  = 1.0f / dstfs;
    const float filterrad = filter.width() / 2.0f;
    t.rad                 = (int)ceilf(filterrad / t.ratio);
    t.ntaps               = 2 * t.rad + 1;
    const int n           = std::max(0, g.dst_end - g.dst_begin);
    t.center.resize(n);
    t.frac.resize(n);
    t.wsum.resize(n);
    t.w.resize(size_t(n) * t.ntaps);
    for (int k = 0; k < n; ++k) {
        int d        = g.dst_begin + k;
        float s      = (d - dstf0 + 0.5f) * dstpixel;
        float src_f  = srcf0 + s * srcfs;
        float fl     = std::floor(src_f);
        int src_i    = int(fl);
        float fr     = src_f - fl;
        t.center[k]  = src_i;
        t.frac[k]    = fr;
        float* w     = t.w.data() + size_t(k) * t.ntaps;
        float total  = 0.0f;
        for (int i = 0; i < t.ntaps; ++i) {
            float a  = t.ratio * (i - t.rad - (fr - 0.5f));
            float wi = is_y ? filter.yfilt(a) : filter.xfilt(a);
            w[i]     = wi;
            total += wi;
        }
        if (total != 0.0f)
            for (int i = 0; i < t.ntaps; ++i)
                w[i] /= total;
        float resum = 0.0f;
        for (int i = 0; i < t.ntaps; ++i)
            resum += w[i];
        // y: the reference tests the un-normalised total (:728, :765);
        // x: it re-adds the normalised taps (:740-743).  Both are zero or
        // non-zero together except for pathological cancellation; keep each
        // axis' own rule.
        t.wsum[k] = is_y ? total : resum;
    }
    return t;
}

bool
axis_overscan(const AxisGeom& g)
{
    return g.src_data_begin < g.src_full_origin
           || g.src_data_end > g.src_full_origin + g.src_full_size;
}

bool
build_axis_gather(const AxisGeom& g, const AxisTaps& taps, AxisGather* out,
                  bool other_axis_overscan)
{
    const int full_lo = g.src_full_origin;
    const int full_hi = g.src_full_origin + g.src_full_size - 1;
    // Overscan (data window beyond the display window): a tap outside the
    // data window is clamped in BOTH axes, which is not separable.  But where
    // every contributing tap of every output lies inside the data window the
    // reference just reads the pixels themselves: such ranges (the interior of
    // the image) are accepted, anything else is left to the exact kernel.
    // The same holds when only the OTHER axis is overscanned: a tap outside
    // the data window on this axis drags the other coordinate to the display
    // window too.
    const bool overscan = other_axis_overscan || axis_overscan(g);
    const int n = taps.n();
    out->n      = n;
    out->first.assign(n, 0);
    out->count.assign(n, 0);
    out->woff.assign(n, 0);
    out->w.clear();
    out->mixed.clear();
    out->maxcount      = 0;
    out->interior_zero = false;
    std::vector<float> merged;
    std::vector<uint8_t> signs;
    for (int k = 0; k < n; ++k) {
        out->woff[k] = int(out->w.size());
        if (taps.wsum[k] == 0.0f)
            continue;  // whole footprint contributes nothing
        const float* w = taps.w.data() + size_t(k) * taps.ntaps;
        int lo = 1 << 30, hi = -(1 << 30);
        for (int i = 0; i < taps.ntaps; ++i) {
            if (w[i] == 0.0f)
                continue;
            int idx = taps.center[k] - taps.rad + i;
            if (overscan && (idx < g.src_data_begin || idx >= g.src_data_end))
                return false;
            int c = overscan ? idx : std::min(std::max(idx, full_lo), full_hi);
            if (c < g.src_data_begin || c >= g.src_data_end)
                continue;  // black pixel
            lo = std::min(lo, c);
            hi = std::max(hi, c);
        }
        if (hi < lo)
            continue;
        int cnt = hi - lo + 1;
        merged.assign(cnt, 0.0f);
        signs.assign(cnt, 0);
        for (int i = 0; i < taps.ntaps; ++i) {
            if (w[i] == 0.0f)
                continue;
            int idx = taps.center[k] - taps.rad + i;
            int c   = overscan ? idx : std::min(std::max(idx, full_lo), full_hi);
            if (c < g.src_data_begin || c >= g.src_data_end)
                continue;
            merged[c - lo] += w[i];
            signs[c - lo] |= (w[i] > 0.0f) ? 1 : 2;
        }
        for (int i = 0; i < cnt; ++i)
            if (merged[i] == 0.0f)
                out->interior_zero = true;
        out->first[k] = lo - g.src_data_begin;
        out->count[k] = cnt;
        out->w.insert(out->w.end(), merged.begin(), merged.end());
        for (int i = 0; i < cnt; ++i)
            out->mixed.push_back(signs[i] == 3 ? 1 : 0);
        out->maxcount = std::max(out->maxcount, cnt);
    }
    return true;
}

}  // namespace b2r
