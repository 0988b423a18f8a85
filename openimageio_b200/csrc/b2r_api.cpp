// b2r_api.cpp -- the C ABI (include/b2r.h): argument checks in the reference's
// words, host<->device staging, plan cache, kernel launches.
//
// Mirrors, for the resize path, what ImageBufAlgo::resize does after IBAprep:
// get_resize_filter (imagebufalgo_xform.cpp:830-860), the (dst,src) type-pair
// rule of OIIO_DISPATCH_COMMON_TYPES2 (imagebufalgo_util.h:463-552) and the
// call into resize_<D,S> (:595-826) -- here two CUDA kernels instead.
#include "api_internal.h"
#include "host_copy.h"

#include "axis_plan.h"
#include "fused_plan.h"
#include "kernels.h"
#include "recon_filter.h"
#include "conv_kernel.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <memory>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace b2r;
using namespace b2r::api;

namespace {

// ---------------------------------------------------------------------------
// A device blob holding every table of one resize shape.  Cached (LRU) so a
// sequence of frames with the same geometry uploads once
// ("precomputed on the host ... uploaded once per (src,dst,filter) shape").
// ---------------------------------------------------------------------------
struct PlanKey {
    int device;
    int g[16];
    float fw, fh;
    std::string fname;
    int nch, st;
    bool operator==(const PlanKey& o) const
    {
        return device == o.device && !memcmp(g, o.g, sizeof(g)) && fw == o.fw
               && fh == o.fh && fname == o.fname && nch == o.nch && st == o.st;
    }
};

struct DevicePlan {
    PlanKey key;
    int device = 0;
    void* blob = nullptr;
    // exact tables
    int radi = 0, radj = 0, xtaps = 0, ytaps = 0;
    float xratio = 1, yratio = 1;
    int src_row0 = 0, src_row1 = 0;  // source rows the ROI can touch (data-window
                                     // relative, inclusive): band + filter halo
    const int* xcenter = nullptr;
    const float* xw = nullptr;
    const float* xwsum = nullptr;
    const float* xfrac = nullptr;
    const int* ycenter = nullptr;
    const float* yw = nullptr;
    const float* ywsum = nullptr;
    const float* yfrac = nullptr;
    // wide (two plain passes over the gather tables; any window width)
    bool wide = false;
    int wide_x0 = 0, wide_w = 0;  // source columns the H pass reads
    const int *gyfirst = nullptr, *gycount = nullptr, *gywoff = nullptr;
    const float* gyw = nullptr;
    const unsigned char* gymixed = nullptr;
    const int *gxfirst = nullptr, *gxcount = nullptr, *gxwoff = nullptr;
    const float* gxw = nullptr;
    const unsigned char* gxmixed = nullptr;
    // fused
    bool fused = false;
    FusedPlan fp;  // host copies of strips/bands kept for sizes only
    const Strip* strips = nullptr;
    const Band* bands = nullptr;
    const int* hfirst = nullptr;
    const float* hw = nullptr;
    const float* vring = nullptr;
    const float* vstream = nullptr;
    const int* hmix = nullptr;
    const int* vlast = nullptr;
    const int* vfirst = nullptr;
    const float* vw = nullptr;
    ~DevicePlan()
    {
        if (blob) {
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(device);
            cudaFree(blob);
            cudaSetDevice(cur);
        }
    }
};

// frames of the batch being planned (b2r_resize_batch): a small frame alone
// cannot fill the SMs with wide strips, a batch of them can
thread_local int g_batch_frames = 1;

std::mutex g_plan_mutex;
std::list<std::shared_ptr<DevicePlan>> g_plans;  // most recent first
// above the band count of the pipelined host path (<= 64 bands, one plan each)
const size_t kMaxPlans = 192;

struct BlobBuilder {
    std::vector<unsigned char> host;
    size_t add(const void* p, size_t bytes)
    {
        size_t off = (host.size() + 255) & ~size_t(255);
        host.resize(off + bytes);
        if (bytes)
            memcpy(host.data() + off, p, bytes);
        return off;
    }
    template<class T> size_t add(const std::vector<T>& v)
    {
        return add(v.data(), v.size() * sizeof(T));
    }
};

// Source rows (data-window relative, inclusive) a dst ROI can touch: rows
// [min(center)-rad, max(center)+rad], resolved the way the taps are -- clamped
// to the full window (WrapClamp), then cut to the data window.  The
// non-separable branch walks a window that starts radi (not radj) above the
// centre (xform.cpp:793-794), so it takes the larger radius and extent.
void
source_row_span(const b2r_image& src, const Filter2D& filter, const AxisTaps& tx,
                const AxisTaps& ty, int* row0, int* row1)
{
    int lo = INT_MAX, hi = INT_MIN;
    for (int c : ty.center) {
        lo = std::min(lo, c);
        hi = std::max(hi, c);
    }
    if (ty.center.empty())
        lo = hi = src.y;
    int rad = std::max(tx.rad, ty.rad);
    lo -= rad;
    hi += rad + (filter.separable() ? 0 : 2 * rad);
    // a tap inside the data window is used where it is (even outside the
    // display window: overscan); a tap outside it is clamped to the display
    // window first.  Cover both readings of the extreme taps.
    const int flo = src.full_y, fhi = src.full_y + src.full_height - 1;
    int lo_c = std::min(std::max(lo, flo), fhi);
    int hi_c = std::min(std::max(hi, flo), fhi);
    lo = std::max(std::min(lo, lo_c), src.y);
    hi = std::min(std::max(hi, hi_c), src.y + src.height - 1);
    if (hi < lo)
        lo = hi = src.y;  // nothing but black is visible: stage one row
    // The kernels' row windows are rounded up (gather / expand read an even
    // number of rows, shifted back inside the image at its bottom edge) and
    // read those extra rows with zero weights: keep them inside what is staged.
    lo = std::max(lo - 8, src.y);
    hi = std::min(hi + 8, src.y + src.height - 1);
    *row0 = lo - src.y;
    *row1 = hi - src.y;
}

// Is there a kernel for this plan?  Ring-mode plans with a compile-time H
// window belong to the stream kernel (their source rows are TMA-fetchable:
// contiguous pixels, re-pitched if need be), everything else to the expand /
// fused kernels.
static bool
plan_supported(int basetype, int nch, const FusedPlan& fp, bool tma_ok)
{
    if (fp.expand)
        return expand_supported(basetype, nch, fp.vt, fp.ht);
    if (fp.vk > 0 && fp.ht > 0 && fp.ht <= 16)
        return tma_ok && stream_supported(basetype, nch, fp.vk, fp.ht, fp.w2);
    return fused_supported(basetype, nch, fp.vk, fp.ht);
}

int
get_plan(int device, const b2r_image& dst, const b2r_image& src,
         const b2r_roi& roi, const Filter2D& filter, int nch, int cta_slots,
         cudaStream_t stream, std::shared_ptr<DevicePlan>* out, char* err,
         size_t errlen, bool tma_ok = false)
{
    PlanKey key;
    key.device = device;
    int g[16]  = { roi.xbegin, roi.xend, roi.ybegin, roi.yend,
                   dst.full_x, dst.full_y, dst.full_width, dst.full_height,
                   src.full_x, src.full_y, src.full_width, src.full_height,
                   src.x, src.y, src.width, src.height };
    memcpy(key.g, g, sizeof(g));
    key.fw    = filter.width();
    key.fh    = filter.height();
    key.fname = std::string(filter.name());
    key.nch   = nch;
    // source rows TMA can fetch (16-byte aligned base and pitch) get the stream
    // kernel's strip geometry: part of the key
    // (the band count of a batch plan depends on the number of frames)
    key.st    = src.basetype | (tma_ok ? 0x100 : 0) | (std::min(g_batch_frames, 0xffff) << 10);
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        for (auto it = g_plans.begin(); it != g_plans.end(); ++it) {
            if ((*it)->key == key) {
                *out = *it;
                g_plans.splice(g_plans.begin(), g_plans, it);
                return B2R_OK;
            }
        }
    }
    auto plan    = std::make_shared<DevicePlan>();
    plan->key    = key;
    plan->device = device;

    AxisGeom gx { roi.xbegin, roi.xend, dst.full_x, dst.full_width, src.full_x,
                  src.full_width, src.x, src.x + src.width };
    AxisGeom gy { roi.ybegin, roi.yend, dst.full_y, dst.full_height, src.full_y,
                  src.full_height, src.y, src.y + src.height };
    AxisTaps tx = build_axis_taps(gx, filter, false);
    AxisTaps ty = build_axis_taps(gy, filter, true);
    plan->radi  = tx.rad;
    plan->radj  = ty.rad;
    plan->xtaps = tx.ntaps;
    plan->ytaps = ty.ntaps;
    plan->xratio = tx.ratio;
    plan->yratio = ty.ratio;
    source_row_span(src, filter, tx, ty, &plan->src_row0, &plan->src_row1);

    BlobBuilder bb;
    size_t o_xc = bb.add(tx.center), o_xw = bb.add(tx.w), o_xs = bb.add(tx.wsum),
           o_xf = bb.add(tx.frac);
    size_t o_yc = bb.add(ty.center), o_yw = bb.add(ty.w), o_ys = bb.add(ty.wsum),
           o_yf = bb.add(ty.frac);

    size_t o_strips = 0, o_bands = 0, o_hf = 0, o_hw = 0, o_vr = 0, o_vl = 0,
           o_vf = 0, o_vw = 0, o_vs = 0, o_hm = 0;
    size_t o_gy[5] = { 0, 0, 0, 0, 0 }, o_gx[5] = { 0, 0, 0, 0, 0 };
    if (filter.separable()) {
        AxisGather ax, ay;
        const int e0a = 16 / basetype_size(src.basetype);
        const int sms = cta_slots / 3;
        bool ok = build_axis_gather(gx, tx, &ax, axis_overscan(gy))
                  && build_axis_gather(gy, ty, &ay, axis_overscan(gx));
        const bool gather_ok = ok;
        // first choice: the persistent TMA kernel with 512 V threads (wide
        // strips, two work items per SM); it needs a ring-mode shape with a
        // compile-time H window and enough work to fill the SMs that way
        bool stream_plan = false;
        if (ok && tma_ok) {
            stream_plan = plan_fused(ax, ay, nch, src.width, src.height, sms * 2, &plan->fp,
                                     e0a, 2, sms, g_batch_frames)
                          && plan->fp.w2 == 2 && !plan->fp.expand && plan->fp.vk > 0
                          && stream_supported(src.basetype, nch, plan->fp.vk, plan->fp.ht, 2)
                          && (long long)plan->fp.strips.size() * (long long)plan->fp.bands.size()
                                     * g_batch_frames
                                 >= sms;
        }
        ok = ok
             && (stream_plan
                 || (plan_fused(ax, ay, nch, src.width, src.height, cta_slots, &plan->fp, e0a)
                     && plan_supported(src.basetype, nch, plan->fp, tma_ok)));
        if (ok && !stream_plan) {
            // the band count was sized for an assumed number of resident CTAs
            // per SM; ask the runtime what this shape really gets and re-plan
            // when it differs (one wave, no ragged tail)
            FusedParams q;
            memset(&q, 0, sizeof(q));
            q.srctype = src.basetype;
            q.nch     = nch;
            q.vk      = plan->fp.vk;
            q.ht      = plan->fp.ht;
            q.vt      = plan->fp.vt;
            q.expand  = plan->fp.expand ? 1 : 0;
            int per   = fused_ctas_per_sm(q, plan->fp.max_band_rows,
                                          plan->fp.max_band_dy, plan->fp.max_nvec);
            if (per > 0 && per != 3)
                ok = plan_fused(ax, ay, nch, src.width, src.height, sms * per,
                                &plan->fp, e0a)
                     && plan_supported(src.basetype, nch, plan->fp, tma_ok);
        }
        if (ok) {
            plan->fused = true;
            o_strips    = bb.add(plan->fp.strips);
            o_bands     = bb.add(plan->fp.bands);
            o_hf        = bb.add(plan->fp.hfirst);
            o_hw        = bb.add(plan->fp.hw);
            o_vr        = bb.add(plan->fp.vring);
            o_vs        = bb.add(plan->fp.vstream);
            o_hm        = bb.add(plan->fp.hmix);
            o_vl        = bb.add(plan->fp.vlast);
            o_vf        = bb.add(plan->fp.vfirst);
            o_vw        = bb.add(plan->fp.vw);
            // the big host tables are no longer needed
            plan->fp.hw.clear();
            plan->fp.hw.shrink_to_fit();
            plan->fp.vring.clear();
            plan->fp.vring.shrink_to_fit();
            plan->fp.vstream.clear();
            plan->fp.vstream.shrink_to_fit();
        } else if (gather_ok && (ax.maxcount > 16 || ay.maxcount > 16)) {
            // a shape no fused instantiation covers because a window is too wide
            // (thumbnail-scale minification): two plain passes over the gather
            // tables instead of the O(taps^2) exact kernel
            int lo = INT_MAX, hi = 0;
            for (int k = 0; k < ax.n; ++k)
                if (ax.count[k] > 0) {
                    lo = std::min(lo, ax.first[k]);
                    hi = std::max(hi, ax.first[k] + ax.count[k]);
                }
            if (hi > lo) {
                plan->wide    = true;
                plan->wide_x0 = lo;
                plan->wide_w  = hi - lo;
                // a column without taps reads nothing; keep its offset in range
                for (int k = 0; k < ax.n; ++k)
                    if (ax.count[k] <= 0)
                        ax.first[k] = lo;
                o_gy[0] = bb.add(ay.first), o_gy[1] = bb.add(ay.count), o_gy[2] = bb.add(ay.woff);
                o_gy[3] = bb.add(ay.w), o_gy[4] = bb.add(ay.mixed);
                o_gx[0] = bb.add(ax.first), o_gx[1] = bb.add(ax.count), o_gx[2] = bb.add(ax.woff);
                o_gx[3] = bb.add(ax.w), o_gx[4] = bb.add(ax.mixed);
            }
        }
    }
    CUDA_TRY(cudaMalloc(&plan->blob, std::max<size_t>(bb.host.size(), 256)));
    CUDA_TRY(cudaMemcpyAsync(plan->blob, bb.host.data(), bb.host.size(),
                             cudaMemcpyHostToDevice, stream));
    // The plan is about to be published in the cache, where a call on ANOTHER
    // stream can pick it up: the tables must have landed by then (a pageable
    // async copy only guarantees that the data has been staged).  Once per shape.
    CUDA_TRY(cudaStreamSynchronize(stream));
    unsigned char* base = (unsigned char*)plan->blob;
    plan->xcenter = (const int*)(base + o_xc);
    plan->xw      = (const float*)(base + o_xw);
    plan->xwsum   = (const float*)(base + o_xs);
    plan->xfrac   = (const float*)(base + o_xf);
    if (plan->wide) {
        plan->gyfirst = (const int*)(base + o_gy[0]);
        plan->gycount = (const int*)(base + o_gy[1]);
        plan->gywoff  = (const int*)(base + o_gy[2]);
        plan->gyw     = (const float*)(base + o_gy[3]);
        plan->gymixed = base + o_gy[4];
        plan->gxfirst = (const int*)(base + o_gx[0]);
        plan->gxcount = (const int*)(base + o_gx[1]);
        plan->gxwoff  = (const int*)(base + o_gx[2]);
        plan->gxw     = (const float*)(base + o_gx[3]);
        plan->gxmixed = base + o_gx[4];
    }
    plan->ycenter = (const int*)(base + o_yc);
    plan->yw      = (const float*)(base + o_yw);
    plan->ywsum   = (const float*)(base + o_ys);
    plan->yfrac   = (const float*)(base + o_yf);
    if (plan->fused) {
        plan->strips = (const Strip*)(base + o_strips);
        plan->bands  = (const Band*)(base + o_bands);
        plan->hfirst = (const int*)(base + o_hf);
        plan->hw     = (const float*)(base + o_hw);
        plan->vring  = (const float*)(base + o_vr);
        plan->vstream = (const float*)(base + o_vs);
        plan->hmix    = (const int*)(base + o_hm);
        plan->vlast  = (const int*)(base + o_vl);
        plan->vfirst = (const int*)(base + o_vf);
        plan->vw     = (const float*)(base + o_vw);
    }
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        g_plans.push_front(plan);
        while (g_plans.size() > kMaxPlans)
            g_plans.pop_back();
    }
    *out = plan;
    return B2R_OK;
}

// Ring of redo slots per device.  A call gets a slot -- one flag plus a dirty
// map -- and a fresh, non-zero epoch: the fused kernel writes the epoch into
// the flag when it sees a non-finite output and into the map cell (128 columns
// x 2^yshift rows of the ROI) that holds the pixel; the gated exact kernel
// runs only if the flag holds that epoch and re-does only the cells that do.
// Stale values of earlier calls never match, so nothing is cleared and no
// memset node is needed per call.
constexpr int REDO_SLOTS     = 256;
constexpr int REDO_MAP_CELLS = 16384;  // ints per slot (64 KB)
struct RedoSlot {
    int* flag  = nullptr;
    int* dirty = nullptr;
    int epoch  = 0;
};
RedoSlot
next_flag(int dev)
{
    struct Ring {
        int* flags = nullptr;
        int* maps  = nullptr;
        unsigned next = 0;
    };
    static Ring rings[64];
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    Ring& r = rings[dev & 63];
    RedoSlot s;
    if (!r.flags) {
        const size_t map_bytes = size_t(REDO_SLOTS) * REDO_MAP_CELLS * sizeof(int);
        if (cudaMalloc((void**)&r.flags, REDO_SLOTS * sizeof(int)) != cudaSuccess)
            return s;
        if (cudaMalloc((void**)&r.maps, map_bytes) != cudaSuccess) {
            cudaFree(r.flags);
            r.flags = nullptr;
            return s;
        }
        cudaMemset(r.flags, 0, REDO_SLOTS * sizeof(int));
        cudaMemset(r.maps, 0, map_bytes);
    }
    unsigned n = r.next++;
    s.epoch    = int((n & 0x3fffffffu) + 1);
    s.flag     = r.flags + (n % REDO_SLOTS);
    s.dirty    = r.maps + size_t(n % REDO_SLOTS) * REDO_MAP_CELLS;
    return s;
}

// one mutex per device: the three streams are shared by all callers
std::mutex&
pipe_mutex(int dev)
{
    static std::mutex m[64];
    return m[dev & 63];
}

int
resize_host_pipelined(const b2r_image& dst, const b2r_image& src,
                      const Filter2D& filter, const b2r_roi& roi, int device,
                      int path, b2r_report* report, char* err, size_t errlen);

// b2r_resize_batch: the first frame goes through resize_impl with this set; at
// the point where the stream kernel would be launched the parameters are
// handed over instead (the batch entry launches once for all frames).
struct BatchCapture {
    bool want = false, got = false;
    FusedParams fp;
    int w2 = 1, sms = 0;
    std::shared_ptr<DevicePlan> plan;  // keeps the tables alive
};
thread_local BatchCapture* g_capture = nullptr;

// private return code of resize_impl: "highlight compensation was requested
// fused (options.reserved[3]) but this call does not end in the stream kernel
// with float pixels on both sides" -- nothing has been launched
constexpr int HC_NOT_FUSED = 1000;

// options.reserved[] is private to the library:
//   [0] == 1  build + upload the tables only, no launch (MIP chain pass 1)
//   [1] == 1  sub-call of the pipelined host path: never time, never sync
//   [3] == 1  fuse highlight compensation into the stream kernel or return
//             HC_NOT_FUSED before anything is launched
int
resize_impl(const b2r_image* dst_in, const b2r_image* src_in,
            const Filter2D& filter, const b2r_roi* roi_in,
            const b2r_options* opt, b2r_report* report, char* err, size_t errlen)
{
    if (report)
        memset(report, 0, sizeof(*report));
    int rc = check_image(src_in, true, err, errlen);
    if (rc)
        return rc;
    rc = check_image(dst_in, false, err, errlen);
    if (rc)
        return rc;
    const b2r_image& dst = *dst_in;
    const b2r_image& src = *src_in;
    if (src.nchannels < dst.nchannels) {
        set_err(err, errlen,
                "resize: source has %d channels, destination %d",
                src.nchannels, dst.nchannels);
        return B2R_ERR_INVALID;
    }
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 resize path has no CPU "
                "fallback");
        return B2R_ERR_NO_DEVICE;
    }
    // ROI: shrink-wrapped to dst's data window (IBAprep, imagebufalgo.cpp:105-111)
    b2r_roi roi;
    if (roi_in) {
        roi        = *roi_in;
        roi.xbegin = std::max(roi.xbegin, dst.x);
        roi.xend   = std::min(roi.xend, dst.x + dst.width);
        roi.ybegin = std::max(roi.ybegin, dst.y);
        roi.yend   = std::min(roi.yend, dst.y + dst.height);
    } else {
        roi = b2r_roi { dst.x, dst.x + dst.width, dst.y, dst.y + dst.height, 0,
                        dst.nchannels };
    }
    if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin)
        return B2R_OK;
    const int roi_w = roi.xend - roi.xbegin, roi_h = roi.yend - roi.ybegin;
    const int nch = dst.nchannels;  // resize_ writes every dst channel (:603)
    const int ses = basetype_size(src.basetype);
    const int des = basetype_size(dst.basetype);

    int sdev = opt ? opt->device : 0, ddev = sdev;
    const int smem = memspace_of(&src, &sdev);
    const int dmem = memspace_of(&dst, &ddev);
    int device     = opt ? opt->device : 0;
    if (smem == B2R_MEM_DEVICE)
        device = sdev;
    else if (dmem == B2R_MEM_DEVICE)
        device = ddev;
    if (device < 0 || device >= device_count_cached()) {
        set_err(err, errlen, "invalid CUDA device ordinal %d", device);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(device);
    DeviceInfo& di      = device_info(device);
    cudaStream_t stream = opt ? (cudaStream_t)opt->cuda_stream : nullptr;
    const int path      = opt ? opt->path : B2R_PATH_AUTO;

    const bool want_hc = opt && opt->reserved[3] >= 1;  // 2: + maketx's clamp
    if (want_hc
        && !((src.basetype == B2R_FLOAT || src.basetype == B2R_HALF)
             && (dst.basetype == B2R_FLOAT || dst.basetype == B2R_HALF) && smem == B2R_MEM_DEVICE
             && dmem == B2R_MEM_DEVICE && nch <= 4 && src.nchannels == nch && filter.separable()
             && path != B2R_PATH_EXACT))
        return HC_NOT_FUSED;
    // (dst,src) pairs outside OIIO_DISPATCH_COMMON_TYPES2's ten native ones
    // are computed through a float temporary and converted back
    // (imagebufalgo_util.h:463-478, 540-546) -- here on the device: float
    // ROI rectangle, then convert_type<float,D> into dst.
    if (!legal_pair(dst.basetype, src.basetype)) {
        if (dst.xstride != (int64_t)dst.nchannels * des) {
            set_err(err, errlen, "destination pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        void* ftmp = nullptr;
        const size_t frow = size_t(roi_w) * dst.nchannels * 4;
        CUDA_TRY(cudaMallocAsync(&ftmp, frow * roi_h, stream));
        b2r_image F = dst;
        F.pixels    = ftmp;
        F.basetype  = B2R_FLOAT;
        F.x         = roi.xbegin;
        F.y         = roi.ybegin;
        F.width     = roi_w;
        F.height    = roi_h;
        F.xstride   = (int64_t)dst.nchannels * 4;
        F.ystride   = (int64_t)frow;
        F.memspace  = B2R_MEM_DEVICE;
        F.device    = device;
        b2r_options sub;
        memset(&sub, 0, sizeof(sub));
        sub.device      = device;
        sub.path        = path;
        sub.cuda_stream = (void*)stream;
        sub.reserved[1] = 1;
        rc = resize_impl(&F, &src, filter, &roi, &sub, report, err, errlen);
        if (rc == B2R_OK) {
            const size_t drow = size_t(roi_w) * dst.xstride;
            unsigned char* out = (unsigned char*)dst.pixels
                                 + (long long)(roi.ybegin - dst.y) * dst.ystride
                                 + (long long)(roi.xbegin - dst.x) * dst.xstride;
            if (dmem == B2R_MEM_DEVICE) {
                launch_convert_from_float((const float*)ftmp, (long long)frow, out,
                                          dst.ystride, roi_w * dst.nchannels, roi_h,
                                          dst.basetype, stream);
            } else {
                void* dtmp = nullptr;
                CUDA_TRY(cudaMallocAsync(&dtmp, drow * roi_h, stream));
                launch_convert_from_float((const float*)ftmp, (long long)frow, dtmp,
                                          (long long)drow, roi_w * dst.nchannels,
                                          roi_h, dst.basetype, stream);
                CUDA_TRY(copy_d2h_2d(out, dst.ystride, dtmp, drow, drow, roi_h, stream));
                CUDA_TRY(cudaFreeAsync(dtmp, stream));
                if (report)
                    report->d2h_bytes += (int64_t)drow * roi_h;
            }
            if (report)
                report->kernels_launched += 1;
        }
        cudaFreeAsync(ftmp, stream);
        if (rc == B2R_OK && (dmem == B2R_MEM_HOST || smem == B2R_MEM_HOST))
            CUDA_TRY(cudaStreamSynchronize(stream));
        return rc;
    }

    // Overscan (source data window beyond its display window): only outputs
    // whose taps all lie inside the data window are separable (axis_plan.cpp).
    // On the device, cut the ROI into that interior -- which the fused kernel
    // takes -- and the thin frame around it, which stays with the exact kernel.
    if (smem == B2R_MEM_DEVICE && dmem == B2R_MEM_DEVICE && filter.separable()
        && path == B2R_PATH_AUTO && !(opt && (opt->reserved[0] || opt->reserved[2]))
        && !want_hc
        && (src.x < src.full_x || src.y < src.full_y
            || src.x + src.width > src.full_x + src.full_width
            || src.y + src.height > src.full_y + src.full_height)) {
        AxisGeom gx { roi.xbegin, roi.xend, dst.full_x, dst.full_width, src.full_x,
                      src.full_width, src.x, src.x + src.width };
        AxisGeom gy { roi.ybegin, roi.yend, dst.full_y, dst.full_height, src.full_y,
                      src.full_height, src.y, src.y + src.height };
        AxisTaps tx = build_axis_taps(gx, filter, false);
        AxisTaps ty = build_axis_taps(gy, filter, true);
        auto interior = [](const AxisTaps& t, int data_begin, int data_end, int* a, int* b) {
            const int n = t.n();
            int lo = 0, hi = n;
            while (lo < n && t.center[lo] - t.rad < data_begin)
                ++lo;
            while (hi > lo && t.center[hi - 1] + t.rad > data_end - 1)
                --hi;
            for (int k = lo; k < hi; ++k)  // centres are monotone; be safe anyway
                if (t.center[k] - t.rad < data_begin || t.center[k] + t.rad > data_end - 1)
                    hi = k;
            *a = lo, *b = hi;
        };
        int xa, xb, ya, yb;
        interior(tx, src.x, src.x + src.width, &xa, &xb);
        interior(ty, src.y, src.y + src.height, &ya, &yb);
        if (xb - xa >= 64 && yb - ya >= 64) {
            const int X0 = roi.xbegin + xa, X1 = roi.xbegin + xb;
            const int Y0 = roi.ybegin + ya, Y1 = roi.ybegin + yb;
            const b2r_roi parts[5] = {
                { X0, X1, Y0, Y1, roi.chbegin, roi.chend },                  // interior
                { roi.xbegin, roi.xend, roi.ybegin, Y0, roi.chbegin, roi.chend },  // top
                { roi.xbegin, roi.xend, Y1, roi.yend, roi.chbegin, roi.chend },    // bottom
                { roi.xbegin, X0, Y0, Y1, roi.chbegin, roi.chend },          // left
                { X1, roi.xend, Y0, Y1, roi.chbegin, roi.chend },            // right
            };
            b2r_options sub;
            memset(&sub, 0, sizeof(sub));
            if (opt)
                sub = *opt;
            sub.device      = device;
            sub.cuda_stream = (void*)stream;
            sub.reserved[2] = 1;  // private: do not split again
            b2r_report first, part;
            memset(&first, 0, sizeof(first));
            int launched = 0;
            for (int i = 0; i < 5 && rc == B2R_OK; ++i) {
                if (parts[i].xend <= parts[i].xbegin || parts[i].yend <= parts[i].ybegin)
                    continue;
                rc = resize_impl(dst_in, src_in, filter, &parts[i], &sub,
                                 report ? &part : nullptr, err, errlen);
                if (report) {
                    launched += part.kernels_launched;
                    if (i == 0)
                        first = part;
                }
            }
            if (report && rc == B2R_OK) {
                *report                  = first;
                report->kernels_launched = launched;
            }
            return rc;
        }
    }

    // More than four channels (multi-layer EXRs): the fused kernels handle up
    // to four, and the single-pass exact kernel is ~170x slower on a large
    // frame.  Split into groups of <= 4 channels on the device, run each group
    // through the normal path, put the results back.  Separable filters only
    // (the exact kernel keeps the non-separable ones).
    // The same split serves a source with MORE channels than the destination
    // (resize_ reads the first dst.nchannels of them): the group copy drops
    // the extra ones.
    // ... and images whose pixels are not contiguous in x (a channel sub-view
    // of a wider buffer): the group copy packs them, so they no longer fall to
    // the exact kernel (a host DESTINATION with gaps between its pixels is
    // left out: its rows are staged whole, and the gaps must survive).
    const bool src_gaps = src.xstride != (int64_t)src.nchannels * ses;
    const bool dst_gaps = dst.xstride != (int64_t)nch * des;
    if ((nch > 4 || src.nchannels > nch || src_gaps || dst_gaps) && filter.separable()
        && path != B2R_PATH_EXACT && !(opt && opt->reserved[0])
        && !(dst_gaps && dmem == B2R_MEM_HOST) && !(src_gaps && smem == B2R_MEM_HOST)) {
        void *s_all = nullptr, *d_all = nullptr, *ts = nullptr, *td = nullptr;
        auto cleanup = [&]() {
            for (void* p : { s_all, d_all, ts, td })
                if (p)
                    cudaFreeAsync(p, stream);
        };
        // the whole interleaved source on the device
        const unsigned char* sbase = (const unsigned char*)src.pixels;
        long long s_ys             = src.ystride;
        if (smem == B2R_MEM_HOST) {
            const size_t rowb = size_t(src.width) * src.xstride;
            if (cudaMallocAsync(&s_all, rowb * src.height, stream) != cudaSuccess
                || copy_h2d_2d(s_all, rowb, src.pixels, src.ystride, rowb, src.height, stream)
                       != cudaSuccess) {
                cleanup();
                set_err(err, errlen, "CUDA error staging the source");
                return B2R_ERR_CUDA;
            }
            sbase = (const unsigned char*)s_all;
            s_ys  = (long long)rowb;
            if (report)
                report->h2d_bytes += (int64_t)rowb * src.height;
        }
        // the interleaved dst ROI on the device
        unsigned char* dbase = (unsigned char*)dst.pixels
                               + (long long)(roi.ybegin - dst.y) * dst.ystride
                               + (long long)(roi.xbegin - dst.x) * dst.xstride;
        long long d_ys = dst.ystride;
        const size_t drowb = size_t(roi_w) * dst.xstride;
        if (dmem == B2R_MEM_HOST) {
            if (cudaMallocAsync(&d_all, drowb * roi_h, stream) != cudaSuccess) {
                cleanup();
                set_err(err, errlen, "CUDA error allocating the destination");
                return B2R_ERR_CUDA;
            }
            dbase = (unsigned char*)d_all;
            d_ys  = (long long)drowb;
        }
        if (cudaMallocAsync(&ts, size_t(src.width) * src.height * 4 * ses, stream) != cudaSuccess
            || cudaMallocAsync(&td, size_t(roi_w) * roi_h * 4 * des, stream) != cudaSuccess) {
            cleanup();
            set_err(err, errlen, "CUDA error allocating the channel-group images");
            return B2R_ERR_CUDA;
        }
        b2r_options sub;
        memset(&sub, 0, sizeof(sub));
        sub.device      = device;
        sub.path        = path;
        sub.cuda_stream = (void*)stream;
        sub.reserved[1] = 1;
        int launched    = 0;
        b2r_report last;
        memset(&last, 0, sizeof(last));
        for (int c0 = 0; c0 < nch && rc == B2R_OK; c0 += 4) {
            const int ns = std::min(4, nch - c0);
            b2r_image TS = src, TD = dst;
            TS.pixels    = ts;
            TS.nchannels = ns;
            TS.xstride   = (int64_t)ns * ses;
            TS.ystride   = TS.xstride * src.width;
            TS.memspace  = B2R_MEM_DEVICE;
            TS.device    = device;
            TD.pixels    = td;
            TD.nchannels = ns;
            TD.x         = roi.xbegin;
            TD.y         = roi.ybegin;
            TD.width     = roi_w;
            TD.height    = roi_h;
            TD.xstride   = (int64_t)ns * des;
            TD.ystride   = TD.xstride * roi_w;
            TD.memspace  = B2R_MEM_DEVICE;
            TD.device    = device;
            launch_pixel_bytes_copy(ts, TS.xstride, TS.ystride, sbase + (size_t)c0 * ses,
                                    src.xstride, s_ys, src.width, src.height, ns * ses, stream);
            b2r_roi r = roi;
            r.chbegin = 0;
            r.chend   = ns;
            rc = resize_impl(&TD, &TS, filter, &r, &sub, &last, err, errlen);
            launch_pixel_bytes_copy(dbase + (size_t)c0 * des, dst.xstride, d_ys, td, TD.xstride,
                                    TD.ystride, roi_w, roi_h, ns * des, stream);
            launched += 2 + last.kernels_launched;
        }
        if (rc == B2R_OK && dmem == B2R_MEM_HOST) {
            unsigned char* hp = (unsigned char*)dst.pixels
                                + (long long)(roi.ybegin - dst.y) * dst.ystride
                                + (long long)(roi.xbegin - dst.x) * dst.xstride;
            if (copy_d2h_2d(hp, dst.ystride, d_all, drowb, drowb, roi_h, stream) != cudaSuccess) {
                set_err(err, errlen, "CUDA error copying the result back");
                rc = B2R_ERR_CUDA;
            }
            if (report)
                report->d2h_bytes += (int64_t)drowb * roi_h;
        }
        cleanup();
        if (rc == B2R_OK && (smem == B2R_MEM_HOST || dmem == B2R_MEM_HOST))
            CUDA_TRY(cudaStreamSynchronize(stream));
        if (report && rc == B2R_OK) {
            const int64_t h2d = report->h2d_bytes, d2h = report->d2h_bytes;
            *report                  = last;
            report->h2d_bytes        = h2d;
            report->d2h_bytes        = d2h;
            report->kernels_launched = launched;
        }
        return rc;
    }

    // Large host -> host resize: run it as a pipeline of row bands so the
    // upload of later source rows, the kernels and the download of finished
    // rows overlap (PCIe is full duplex; the copies dominate end to end).
    if (smem == B2R_MEM_HOST && dmem == B2R_MEM_HOST && roi_h >= 256
        && src.xstride == (int64_t)src.nchannels * ses
        && dst.xstride == (int64_t)dst.nchannels * des
        && std::max((int64_t)src.width * src.height * src.xstride,
                    (int64_t)roi_w * roi_h * dst.xstride)
               >= (int64_t(32) << 20)
        && !(opt && (opt->reserved[0] || opt->reserved[1])))
        return resize_host_pipelined(dst, src, filter, roi, device, path, report,
                                     err, errlen);

    // TMA needs 16-byte aligned source rows: host sources are staged with such
    // a pitch; a device source that is not aligned (a packed 7679-pixel RGB
    // image, a sub-view) is re-pitched into a temporary first when its shape
    // belongs to the stream kernel (one copy pass at HBM speed instead of the
    // 2-3x slower unaligned paths)
    const bool src_contig  = src.xstride == (int64_t)src.nchannels * ses;
    const bool src_aligned = ((uintptr_t)src.pixels & 15) == 0 && (src.ystride & 15) == 0;
    const bool tma_ok      = src_contig;
    std::shared_ptr<DevicePlan> plan;
    rc = get_plan(device, dst, src, roi, filter, nch, di.sms * 3, stream, &plan,
                  err, errlen, tma_ok);
    if (rc)
        return rc;
    if (opt && opt->reserved[0] == 1)
        return B2R_OK;  // private: build + upload the tables only (MIP chain)
    if (want_hc
        && !(plan->fused && !plan->fp.expand && plan->fp.vk > 0 && plan->fp.ht > 0
             && plan->fp.ht <= 16 && src_contig && dst.xstride == (int64_t)nch * des
             && tma_ok
             && stream_supported(src.basetype, nch, plan->fp.vk, plan->fp.ht, plan->fp.w2, 1)))
        return HC_NOT_FUSED;

    // ---- stage host images ------------------------------------------------
    void* dsrc       = src.pixels;
    void* ddst       = dst.pixels;
    b2r_image srcd   = src;
    b2r_image dstd   = dst;
    size_t src_bytes = 0, dst_bytes = 0;
    bool src_staged = false, dst_staged = false, src_repitched = false;
    void* src_alloc = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // every early return below (CUDA_TRY and the explicit error paths) runs
    // this: staging buffers back to the pool, stream-ordered; events destroyed.
    // The normal exit frees them itself and clears the pointers first.
    struct StagingGuard {
        void*& src_alloc;
        void*& ddst;
        bool& dst_staged;
        cudaEvent_t& ev0;
        cudaEvent_t& ev1;
        cudaStream_t stream;
        ~StagingGuard()
        {
            if (src_alloc)
                cudaFreeAsync(src_alloc, stream);
            if (dst_staged && ddst)
                cudaFreeAsync(ddst, stream);
            if (ev0)
                cudaEventDestroy(ev0);
            if (ev1)
                cudaEventDestroy(ev1);
        }
    } staging_guard { src_alloc, ddst, dst_staged, ev0, ev1, stream };
    if (smem == B2R_MEM_HOST) {
        if (src.xstride != (int64_t)src.nchannels * ses) {
            set_err(err, errlen, "host source pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        // Only the rows this ROI can touch travel: its band plus the filter
        // halo.  A caller that splits dst into row bands (one per GPU) gets
        // band + halo staging for free, straight from the host source.
        const int r0 = plan->src_row0, nrows = plan->src_row1 - plan->src_row0 + 1;
        size_t rowbytes = size_t(src.width) * src.xstride;
        size_t pitch    = (rowbytes + 15) & ~size_t(15);
        src_bytes       = pitch * nrows;
        CUDA_TRY(cudaMallocAsync(&src_alloc, src_bytes, stream));
        src_staged = true;
        CUDA_TRY(copy_h2d_2d(src_alloc, pitch,
                                   (const unsigned char*)src.pixels
                                       + (long long)r0 * src.ystride,
                                   src.ystride, rowbytes, nrows, stream));
        // virtual origin: row r of the data window lives at dsrc + r*pitch;
        // rows outside [r0, r0+nrows) are never addressed
        dsrc         = (unsigned char*)src_alloc - (long long)r0 * (long long)pitch;
        srcd.ystride = (int64_t)pitch;
        if (report)
            report->h2d_bytes += (int64_t)rowbytes * nrows;
    }
    const bool stream_shape = plan->fused && !plan->fp.expand && plan->fp.vk > 0
                              && plan->fp.ht > 0 && plan->fp.ht <= 16;
    if (smem == B2R_MEM_DEVICE && stream_shape && !src_aligned && src_contig
        && path != B2R_PATH_EXACT) {
        const int r0 = plan->src_row0, nrows = plan->src_row1 - plan->src_row0 + 1;
        size_t rowbytes = size_t(src.width) * src.xstride;
        size_t pitch    = (rowbytes + 15) & ~size_t(15);
        CUDA_TRY(cudaMallocAsync(&src_alloc, pitch * nrows, stream));
        src_repitched = true;
        CUDA_TRY(cudaMemcpy2DAsync(src_alloc, pitch,
                                   (const unsigned char*)src.pixels
                                       + (long long)r0 * src.ystride,
                                   src.ystride, rowbytes, nrows, cudaMemcpyDeviceToDevice,
                                   stream));
        dsrc         = (unsigned char*)src_alloc - (long long)r0 * (long long)pitch;
        srcd.ystride = (int64_t)pitch;
    }
    size_t dpitch = 0, drowbytes = 0;
    if (dmem == B2R_MEM_HOST) {
        if (dst.xstride != (int64_t)dst.nchannels * des) {
            set_err(err, errlen, "host destination pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        // stage only the ROI rectangle
        drowbytes = size_t(roi_w) * dst.xstride;
        dpitch    = (drowbytes + 15) & ~size_t(15);
        dst_bytes = dpitch * roi_h;
        CUDA_TRY(cudaMallocAsync(&ddst, dst_bytes, stream));
        dst_staged   = true;
        dstd.x       = roi.xbegin;
        dstd.y       = roi.ybegin;
        dstd.width   = roi_w;
        dstd.height  = roi_h;
        dstd.ystride = (int64_t)dpitch;
    }

    const bool timed = report != nullptr && !(opt && opt->reserved[1]);
    if (timed) {
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventRecord(ev0, stream);
    }

    // ---- launch -----------------------------------------------------------
    DevImage D = to_dev(dstd, ddst);
    DevImage S = to_dev(srcd, dsrc);
    ExactParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.dst = D;
    ep.src = S;
    ep.roi_x0 = roi.xbegin;
    ep.roi_x1 = roi.xend;
    ep.roi_y0 = roi.ybegin;
    ep.roi_y1 = roi.yend;
    ep.nch    = nch;
    ep.radi   = plan->radi;
    ep.radj   = plan->radj;
    ep.xtaps  = plan->xtaps;
    ep.ytaps  = plan->ytaps;
    ep.xratio = plan->xratio;
    ep.yratio = plan->yratio;
    ep.xcenter = plan->xcenter;
    ep.xw      = plan->xw;
    ep.xwsum   = plan->xwsum;
    ep.xfrac   = plan->xfrac;
    ep.ycenter = plan->ycenter;
    ep.yw      = plan->yw;
    ep.ywsum   = plan->ywsum;
    ep.yfrac   = plan->yfrac;
    ep.nonsep_kind = filter.nonseparable_kind();
    ep.filt_w      = filter.width();
    ep.filt_h      = filter.height();
    ep.gate        = nullptr;

    bool fused_ok = plan->fused && src.nchannels == dst.nchannels
                    && srcd.xstride == (int64_t)nch * ses
                    && dstd.xstride == (int64_t)nch * des;
    int launched = 0;
    int used     = B2R_PATH_EXACT;
    int* flag    = nullptr;
    int epoch    = 0;
    if (path == B2R_PATH_FUSED && !fused_ok) {
        set_err(err, errlen, "resize: shape is not eligible for the fused kernel");
        return B2R_ERR_UNSUPPORTED;
    }
    const bool src_is_fp = src.basetype == B2R_FLOAT || src.basetype == B2R_HALF;
    if (fused_ok && path != B2R_PATH_EXACT) {
        FusedParams fp;
        memset(&fp, 0, sizeof(fp));
        fp.src         = (const unsigned char*)dsrc;
        fp.src_ystride = srcd.ystride;
        fp.srctype     = src.basetype;
        fp.src_w       = src.width;
        fp.src_h       = src.height;
        fp.nch         = nch;
        fp.dst         = (unsigned char*)ddst
                 + (long long)(roi.ybegin - dstd.y) * dstd.ystride
                 + (long long)(roi.xbegin - dstd.x) * dstd.xstride;
        fp.dst_ystride = dstd.ystride;
        fp.dsttype     = dst.basetype;
        fp.src_vec_ok  = ((uintptr_t)dsrc % (4 * ses) == 0)
                        && (srcd.ystride % (4 * ses) == 0);
        fp.dst_vec_ok = ((uintptr_t)fp.dst % (4 * des) == 0)
                        && (dstd.ystride % (4 * des) == 0);
        // zero weights are multiplied, not skipped, in this kernel: if the
        // source can hold Inf/NaN the outputs are screened and the exact
        // kernel re-does the image when one shows up
        fp.check_nonfinite = src_is_fp ? 1 : 0;
        fp.nonfinite_flag  = nullptr;
        fp.nstrips         = (int)plan->fp.strips.size();
        fp.nbands          = (int)plan->fp.bands.size();
        fp.strips          = plan->strips;
        fp.bands           = plan->bands;
        fp.hfirst          = plan->hfirst;
        fp.hw              = plan->hw;
        fp.ht              = plan->fp.ht;
        fp.vk              = plan->fp.vk;
        fp.vring           = plan->vring;
        fp.vstream         = plan->vstream;
        fp.hmix            = plan->hmix;
        fp.vlast           = plan->vlast;
        fp.vt              = plan->fp.vt;
        fp.vfirst          = plan->vfirst;
        fp.vw              = plan->vw;
        fp.expand          = plan->fp.expand ? 1 : 0;
        fp.ngroups         = plan->fp.ngroups;
        fp.vrow_linear     = plan->fp.vrow_linear ? 1 : 0;
        {
            // alignment every group of four dst pixels is guaranteed to have
            const uintptr_t a = (uintptr_t)fp.dst | (uintptr_t)dstd.ystride;
            fp.dst_align = (a % 16 == 0 && (4 * nch * des) % 16 == 0) ? 16
                           : (a % 4 == 0 ? 4 : 0);
            fp.dst_pow2  = 16;
            while (fp.dst_pow2 > 1 && (a & uintptr_t(fp.dst_pow2 - 1)))
                fp.dst_pow2 >>= 1;
        }
        if (src_is_fp) {
            // per-call flag from a per-device ring (plans are shared between
            // concurrent calls); the gated exact kernel leaves it cleared, so
            // no memset node is needed per call
            RedoSlot slot = next_flag(device);
            flag          = slot.flag;
            epoch         = slot.epoch;
            if (!flag) {
                set_err(err, errlen, "CUDA error: cannot allocate the redo flags");
                return B2R_ERR_CUDA;
            }
            fp.nonfinite_flag  = flag;
            fp.nonfinite_epoch = epoch;
            // dirty-map geometry: 128-column cells, as few rows per cell as
            // the map has room for (16 at least)
            fp.dirty        = slot.dirty;
            fp.dirty_xcells = (roi_w + 127) / 128;
            fp.dirty_yshift = 4;
            while ((long long)fp.dirty_xcells * (((long long)roi_h >> fp.dirty_yshift) + 1)
                   > REDO_MAP_CELLS)
                ++fp.dirty_yshift;
        }
        // Ring-mode shapes whose source rows are 16-byte aligned take the
        // persistent TMA kernel (stream_kernel.cuh); it resolves zero weights
        // under non-finite pixels itself, so no gated exact kernel follows.
        bool streamed = false;
        const int w2 = plan->fp.w2;
        if (!fp.expand && fp.vk > 0 && tma_ok
            && stream_supported(src.basetype, nch, fp.vk, fp.ht, w2, want_hc ? 1 : 0)) {
            const bool temp = src_staged || src_repitched;
            const int trow0 = temp ? plan->src_row0 : 0;
            const int trows = temp ? plan->src_row1 - plan->src_row0 + 1 : src.height;
            const unsigned char* tbase
                = (const unsigned char*)dsrc + (long long)trow0 * srcd.ystride;
            fp.tma_row0        = trow0;
            fp.check_nonfinite = 0;
            if (want_hc) {
                fp.hicomp  = 1;
                fp.hc_skip = 0;
                if (opt->alpha_channel1 > 0)
                    fp.hc_skip |= 1 << (opt->alpha_channel1 - 1);
                if (opt->z_channel1 > 0)
                    fp.hc_skip |= 1 << (opt->z_channel1 - 1);
                fp.hc_clamp = opt->reserved[3] == 2 ? 1 : 0;
                fp.hc_alpha = opt->alpha_channel1 - 1;
            }
            if (g_capture && g_capture->want && !temp && !dst_staged
                && !(opt && (opt->reserved[1] || opt->reserved[2]))) {
                // batch entry: hand the launch parameters over, launch nothing
                g_capture->fp   = fp;
                g_capture->w2   = w2;
                g_capture->sms  = di.sms;
                g_capture->plan = plan;
                g_capture->got  = true;
                return B2R_OK;
            }
            streamed = launch_resize_stream(fp, w2, tbase, trows, di.sms, stream) == 0;
            if (!streamed)
                fp.check_nonfinite = src_is_fp ? 1 : 0;
        }
        if (!streamed && w2 != 1) {
            set_err(err, errlen, "resize: the stream kernel could not be launched");
            return B2R_ERR_CUDA;
        }
        if (!streamed)
            launch_resize_fused(fp, plan->fp.max_band_rows, plan->fp.max_band_dy,
                                plan->fp.max_nvec, stream);
        ++launched;
        used = B2R_PATH_FUSED;
        if (src_is_fp && streamed) {
            flag  = nullptr;
            epoch = 0;
        }
        if (src_is_fp && !streamed) {
            ep.gate         = flag;  // runs only if a non-finite output was seen
            ep.gate_value   = epoch;
            ep.dirty        = fp.dirty;  // ... and only over the cells that hold one
            ep.dirty_xcells = fp.dirty_xcells;
            ep.dirty_yshift = fp.dirty_yshift;
            launch_resize_exact(ep, stream);
            ++launched;
        }
        if (report) {
            report->fused_vmode = fp.expand ? -fp.vt : fp.vk;
            report->fused_htaps = fp.ht;
        }
    } else if (plan->wide && path == B2R_PATH_AUTO && filter.separable()
               && src.nchannels >= nch && roi_h <= 65535
               && (size_t)roi_h * plan->wide_w * nch * sizeof(float) <= (size_t(8) << 30)) {
        // windows no fused instantiation holds: V pass into a float
        // intermediate, H pass from it (wide_kernel.cu)
        WideParams wp;
        memset(&wp, 0, sizeof(wp));
        wp.src         = (const unsigned char*)dsrc;
        wp.src_ystride = srcd.ystride;
        wp.src_xstride = srcd.xstride;
        wp.srctype     = src.basetype;
        wp.src_es      = ses;
        wp.dst         = (unsigned char*)ddst + (long long)(roi.ybegin - dstd.y) * dstd.ystride
                 + (long long)(roi.xbegin - dstd.x) * dstd.xstride;
        wp.dst_ystride = dstd.ystride;
        wp.dst_xstride = dstd.xstride;
        wp.dsttype     = dst.basetype;
        wp.dst_es      = des;
        wp.nch         = nch;
        wp.roi_w       = roi_w;
        wp.roi_h       = roi_h;
        wp.x0          = plan->wide_x0;
        wp.tmp_w       = plan->wide_w;
        wp.yfirst = plan->gyfirst, wp.ycount = plan->gycount, wp.ywoff = plan->gywoff;
        wp.yw = plan->gyw, wp.ymixed = plan->gymixed;
        wp.xfirst = plan->gxfirst, wp.xcount = plan->gxcount, wp.xwoff = plan->gxwoff;
        wp.xw = plan->gxw, wp.xmixed = plan->gxmixed;
        void* tmp = nullptr;
        CUDA_TRY(cudaMallocAsync(&tmp, (size_t)roi_h * wp.tmp_w * nch * sizeof(float), stream));
        wp.tmp = (float*)tmp;
        launch_resize_wide(wp, stream);
        cudaFreeAsync(tmp, stream);
        launched += 2;
        used = B2R_PATH_FUSED;  // two-pass arithmetic, the fused kernels' tolerance
    } else {
        launch_resize_exact(ep, stream);
        ++launched;
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA launch error: %s", cudaGetErrorString(e));
            return B2R_ERR_CUDA;
        }
    }
    if (timed)
        cudaEventRecord(ev1, stream);

    // ---- results back -----------------------------------------------------
    if (dst_staged) {
        unsigned char* hp = (unsigned char*)dst.pixels
                            + (long long)(roi.ybegin - dst.y) * dst.ystride
                            + (long long)(roi.xbegin - dst.x) * dst.xstride;
        CUDA_TRY(copy_d2h_2d(hp, dst.ystride, ddst, dpitch, drowbytes, roi_h, stream));
        if (report)
            report->d2h_bytes += (int64_t)drowbytes * roi_h;
    }
    int redo = 0;
    if (report && used == B2R_PATH_FUSED && src_is_fp && flag
        && (dst_staged || src_staged))
        CUDA_TRY(cudaMemcpyAsync(&redo, flag, sizeof(int),
                                 cudaMemcpyDeviceToHost, stream));
    if (src_alloc) {
        void* p   = src_alloc;
        src_alloc = nullptr;
        CUDA_TRY(cudaFreeAsync(p, stream));
    }
    if (dst_staged) {
        void* p = ddst;
        ddst    = nullptr;
        CUDA_TRY(cudaFreeAsync(p, stream));
    }
    if (dst_staged || src_staged || timed)
        CUDA_TRY(cudaStreamSynchronize(stream));
    if (report) {
        report->path_used        = used;
        report->kernels_launched = launched;
        report->xtaps            = plan->xtaps;
        report->ytaps            = plan->ytaps;
        report->nonfinite_redo   = (epoch != 0 && redo == epoch) ? 1 : 0;
        if (timed) {
            float ms = 0.0f;
            cudaEventElapsedTime(&ms, ev0, ev1);
            report->kernel_ms = ms;
        }
    }
    return B2R_OK;  // staging_guard destroys the events
}

// ---------------------------------------------------------------------------
// Host -> host resize as a 3-stream pipeline over row bands of the dst ROI:
//   stream in : H2D of the source rows band i still needs (band + halo,
//               each source row travels once)
//   stream k  : the band's kernels (device-resident sub-call)
//   stream out: D2H of the band's finished rows
// ---------------------------------------------------------------------------
struct PipeStreams {
    cudaStream_t in = nullptr, k = nullptr, out = nullptr;
};
PipeStreams&
pipe_streams(int dev)
{
    static PipeStreams s[64];
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    PipeStreams& p = s[dev & 63];
    if (!p.in) {
        cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&p.k, cudaStreamNonBlocking);
        cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking);
    }
    return p;
}

int
resize_host_pipelined(const b2r_image& dst, const b2r_image& src,
                      const Filter2D& filter, const b2r_roi& roi, int device,
                      int path, b2r_report* report, char* err, size_t errlen)
{
    const int roi_w = roi.xend - roi.xbegin, roi_h = roi.yend - roi.ybegin;
    const bool src_pageable = !host_is_pinned(src.pixels);
    const bool dst_pageable = !host_is_pinned(dst.pixels);
    // (measured on C0: 16, 24, 32 and 48 bands all take 10.4-10.5 ms per frame --
    // the exposed head and tail are not what separates the call from the plain
    // copy ceiling, the concurrent download on the same link is)
    int nbands = std::max(2, std::min(16, roi_h / 128));
    if (src_pageable || dst_pageable) {
        // keep a bounce slot around 32 MB (3 pinned slots per direction)
        const long long bytes = (long long)src.width * src.height * src.xstride;
        const int want        = (int)std::min<long long>(64, (bytes + (32 << 20) - 1) / (32 << 20));
        nbands                = std::max(nbands, std::min(want, std::max(2, roi_h / 32)));
    }
    // one mutex per device: the three streams are shared by all callers
    std::lock_guard<std::mutex> lock(pipe_mutex(device));
    PipeStreams& ps = pipe_streams(device);
    Bounce& bounce          = pipeline_bounce(device);

    // source rows each band touches (band + filter halo)
    std::vector<b2r_roi> bands(nbands);
    std::vector<int> r0(nbands), r1(nbands);
    AxisGeom gx { roi.xbegin, roi.xbegin, dst.full_x, dst.full_width, src.full_x,
                  src.full_width, src.x, src.x + src.width };
    AxisTaps tx = build_axis_taps(gx, filter, false);
    for (int i = 0; i < nbands; ++i) {
        b2r_band_split(&roi, nbands, i, &bands[i]);
        AxisGeom gy { bands[i].ybegin, bands[i].yend, dst.full_y, dst.full_height,
                      src.full_y, src.full_height, src.y, src.y + src.height };
        AxisTaps ty = build_axis_taps(gy, filter, true);
        source_row_span(src, filter, tx, ty, &r0[i], &r1[i]);
    }
    int lo = r0[0], hi = r1[0];
    for (int i = 1; i < nbands; ++i) {
        lo = std::min(lo, r0[i]);
        hi = std::max(hi, r1[i]);
    }
    const size_t srow = size_t(src.width) * src.xstride;
    const size_t spitch = (srow + 15) & ~size_t(15);
    const size_t drow = size_t(roi_w) * dst.xstride;
    const size_t dpitch = (drow + 15) & ~size_t(15);
    void *salloc = nullptr, *dalloc = nullptr;
    CUDA_TRY(cudaMallocAsync(&salloc, spitch * size_t(hi - lo + 1), ps.k));
    CUDA_TRY(cudaMallocAsync(&dalloc, dpitch * size_t(roi_h), ps.k));
    cudaEvent_t ready;  // the allocations exist before any copy touches them
    cudaEventCreateWithFlags(&ready, cudaEventDisableTiming);
    cudaEventRecord(ready, ps.k);
    cudaStreamWaitEvent(ps.in, ready, 0);
    cudaStreamWaitEvent(ps.out, ready, 0);

    b2r_image S = src, D = dst;
    S.pixels   = (unsigned char*)salloc - (long long)lo * (long long)spitch;
    S.ystride  = (int64_t)spitch;
    S.memspace = B2R_MEM_DEVICE;
    S.device   = device;
    D.pixels   = dalloc;
    D.x        = roi.xbegin;
    D.y        = roi.ybegin;
    D.width    = roi_w;
    D.height   = roi_h;
    D.ystride  = (int64_t)dpitch;
    D.memspace = B2R_MEM_DEVICE;
    D.device   = device;

    b2r_options sub;
    memset(&sub, 0, sizeof(sub));
    sub.device      = device;
    sub.path        = path;
    sub.cuda_stream = (void*)ps.k;
    sub.reserved[1] = 1;

    std::vector<cudaEvent_t> evs;
    auto new_event = [&]() {
        cudaEvent_t e;
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
        evs.push_back(e);
        return e;
    };
    int rc = B2R_OK;
    int uploaded = lo;  // next source row to upload
    b2r_report last;
    memset(&last, 0, sizeof(last));
    // bounce-buffer bookkeeping (pageable host memory)
    int in_slot = 0, out_slot = 0;
    // the rings are idle here (the previous call drained them): size every slot
    // for the largest band once, so nothing is reallocated while DMAs fly
    size_t max_in = 0, max_out = 0;
    {
        int up = lo;
        for (int i = 0; i < nbands; ++i) {
            if (bands[i].yend <= bands[i].ybegin)
                continue;
            if (r1[i] >= up) {
                max_in = std::max(max_in, srow * size_t(r1[i] - up + 1));
                up     = r1[i] + 1;
            }
            max_out = std::max(max_out, drow * size_t(bands[i].yend - bands[i].ybegin));
        }
    }
    const bool use_in_bounce  = src_pageable && max_in > 0 && bounce.in.reserve(max_in);
    const bool use_out_bounce = dst_pageable && max_out > 0 && bounce.out.reserve(max_out);
    struct PendingOut {
        unsigned char* host;
        int rows;
    };
    PendingOut pending[BOUNCE_SLOTS] = { { nullptr, 0 }, { nullptr, 0 }, { nullptr, 0 } };
    auto drain_out = [&](int slot) {
        if (!pending[slot].host)
            return;
        const unsigned char* pb = (const unsigned char*)bounce.out.acquire(slot);  // waits
        parallel_copy_rows(pending[slot].host, (size_t)dst.ystride, pb, drow, drow,
                           pending[slot].rows);
        pending[slot].host = nullptr;
    };
    for (int i = 0; i < nbands && rc == B2R_OK; ++i) {
        if (bands[i].yend <= bands[i].ybegin)
            continue;
        // bands go top to bottom and their spans are monotone: everything
        // below `uploaded` is already on its way, only [uploaded, r1] is new
        const int need = r1[i];
        if (need >= uploaded) {
            const int nrows = need - uploaded + 1;
            const unsigned char* hsrc = (const unsigned char*)src.pixels
                                        + (long long)uploaded * src.ystride;
            size_t hpitch = (size_t)src.ystride;
            int used_slot = -1;
            if (use_in_bounce) {
                // host threads -> pinned slot, then DMA from the slot
                used_slot = in_slot++ % BOUNCE_SLOTS;
                unsigned char* pb = (unsigned char*)bounce.in.acquire(used_slot);
                parallel_copy_rows(pb, srow, hsrc, (size_t)src.ystride, srow, nrows);
                hsrc   = pb;
                hpitch = srow;
            }
            cudaError_t e = cudaMemcpy2DAsync(
                (unsigned char*)salloc + size_t(uploaded - lo) * spitch, spitch, hsrc, hpitch,
                srow, size_t(nrows), cudaMemcpyHostToDevice, ps.in);
            if (used_slot >= 0)
                bounce.in.release_after(used_slot, ps.in);
            if (e != cudaSuccess) {
                set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
                rc = B2R_ERR_CUDA;
                break;
            }
            if (report)
                report->h2d_bytes += (int64_t)srow * (need - uploaded + 1);
            uploaded = need + 1;
        }
        cudaEvent_t e_in = new_event();
        cudaEventRecord(e_in, ps.in);
        cudaStreamWaitEvent(ps.k, e_in, 0);
        rc = resize_impl(&D, &S, filter, &bands[i], &sub, &last, err, errlen);
        if (rc)
            break;
        cudaEvent_t e_k = new_event();
        cudaEventRecord(e_k, ps.k);
        cudaStreamWaitEvent(ps.out, e_k, 0);
        const int by = bands[i].ybegin - roi.ybegin, bh = bands[i].yend - bands[i].ybegin;
        unsigned char* hp = (unsigned char*)dst.pixels
                            + (long long)(bands[i].ybegin - dst.y) * dst.ystride
                            + (long long)(roi.xbegin - dst.x) * dst.xstride;
        cudaError_t e;
        if (use_out_bounce) {
            // DMA into a pinned slot; host threads move it into the caller's
            // rows once it has landed -- one band later, so the copy-out of
            // band i-1 overlaps the DMA of band i
            const int slot = out_slot++ % BOUNCE_SLOTS;
            drain_out(slot);  // the slot's previous band, if still pending
            unsigned char* pb = (unsigned char*)bounce.out.buf[slot];
            e = cudaMemcpy2DAsync(pb, drow, (unsigned char*)dalloc + size_t(by) * dpitch,
                                  dpitch, drow, size_t(bh), cudaMemcpyDeviceToHost, ps.out);
            bounce.out.release_after(slot, ps.out);
            pending[slot] = PendingOut { hp, bh };
            if (out_slot >= 2)
                drain_out((out_slot - 2) % BOUNCE_SLOTS);
        } else {
            e = cudaMemcpy2DAsync(hp, dst.ystride, (unsigned char*)dalloc + size_t(by) * dpitch,
                                  dpitch, drow, size_t(bh), cudaMemcpyDeviceToHost, ps.out);
        }
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
            rc = B2R_ERR_CUDA;
            break;
        }
        if (report) {
            report->d2h_bytes += (int64_t)drow * bh;
            report->kernels_launched += last.kernels_launched;
        }
    }
    for (int q = 0; q < BOUNCE_SLOTS; ++q)  // oldest first
        drain_out((out_slot + q) % BOUNCE_SLOTS);
    cudaStreamSynchronize(ps.in);
    cudaStreamSynchronize(ps.k);
    cudaError_t es = cudaStreamSynchronize(ps.out);
    bounce.in.idle();  // everything has completed
    bounce.out.idle();
    cudaFreeAsync(salloc, ps.k);
    cudaFreeAsync(dalloc, ps.k);
    cudaEventDestroy(ready);
    for (cudaEvent_t e : evs)
        cudaEventDestroy(e);
    if (rc == B2R_OK && es != cudaSuccess) {
        set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(es));
        rc = B2R_ERR_CUDA;
    }
    if (report && rc == B2R_OK) {
        report->path_used   = last.path_used;
        report->xtaps       = last.xtaps;
        report->ytaps       = last.ytaps;
        report->fused_vmode = last.fused_vmode;
        report->fused_htaps = last.fused_htaps;
    }
    return rc;
}

}  // namespace

// =========================================================================
extern "C" {

const char*
b2r_version(void)
{
    return "b2r 0.1 (sm_100a)";
}

void
b2r_trim_caches(void)
{
    {
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        g_plans.clear();  // ~DevicePlan frees the table blobs
    }
    trim_band_blocks();
    trim_tile_slots();
}

int
b2r_device_count(void)
{
    return device_count_cached();
}

int
b2r_filter_count(int dim)
{
    return dim == 1 ? Filter1D::num_filters()
                    : (dim == 2 ? Filter2D::num_filters() : 0);
}

int
b2r_filter_desc(int dim, int index, const char** name, float* width,
                int* fixedwidth, int* scalable, int* separable)
{
    if ((dim != 1 && dim != 2) || index < 0 || index >= b2r_filter_count(dim))
        return B2R_ERR_INVALID;
    const FilterDesc& d = dim == 1 ? Filter1D::get_filterdesc(index)
                                   : Filter2D::get_filterdesc(index);
    if (name)
        *name = d.name;
    if (width)
        *width = d.width;
    if (fixedwidth)
        *fixedwidth = d.fixedwidth;
    if (scalable)
        *scalable = d.scalable;
    if (separable)
        *separable = d.separable;
    return B2R_OK;
}

b2r_filter2d*
b2r_filter2d_create(const char* name, float width, float height)
{
    Filter2D* f = Filter2D::create(name ? name : "", width, height);
    if (!f)
        return nullptr;
    return new b2r_filter2d { f };
}
void
b2r_filter2d_destroy(b2r_filter2d* f)
{
    if (f) {
        Filter2D::destroy(f->f);
        delete f;
    }
}
float
b2r_filter2d_width(const b2r_filter2d* f)
{
    return f->f->width();
}
float
b2r_filter2d_height(const b2r_filter2d* f)
{
    return f->f->height();
}
int
b2r_filter2d_separable(const b2r_filter2d* f)
{
    return f->f->separable() ? 1 : 0;
}
const char*
b2r_filter2d_name(const b2r_filter2d* f)
{
    return f->f->name().data();
}
float
b2r_filter2d_eval(const b2r_filter2d* f, float x, float y)
{
    return (*f->f)(x, y);
}
float
b2r_filter2d_xfilt(const b2r_filter2d* f, float x)
{
    return f->f->xfilt(x);
}
float
b2r_filter2d_yfilt(const b2r_filter2d* f, float y)
{
    return f->f->yfilt(y);
}
b2r_filter1d*
b2r_filter1d_create(const char* name, float width)
{
    Filter1D* f = Filter1D::create(name ? name : "", width);
    if (!f)
        return nullptr;
    return new b2r_filter1d { f };
}
void
b2r_filter1d_destroy(b2r_filter1d* f)
{
    if (f) {
        Filter1D::destroy(f->f);
        delete f;
    }
}
float
b2r_filter1d_width(const b2r_filter1d* f)
{
    return f->f->width();
}
const char*
b2r_filter1d_name(const b2r_filter1d* f)
{
    return f->f->name().data();
}
float
b2r_filter1d_eval(const b2r_filter1d* f, float x)
{
    return (*f->f)(x);
}

namespace {
int
resize_highlightcomp(const b2r_image* dst, const b2r_image* src, const Filter2D& filter,
                     const b2r_roi* roi, const b2r_options* opt, b2r_report* report,
                     char* err, size_t errlen);

int
resize_entry(const b2r_image* dst, const b2r_image* src, const Filter2D& filter,
             const b2r_roi* roi, const b2r_options* opt, b2r_report* report, char* err,
             size_t errlen)
{
    if (opt && opt->highlightcomp)
        return resize_highlightcomp(dst, src, filter, roi, opt, report, err, errlen);
    return resize_impl(dst, src, filter, roi, opt, report, err, errlen);
}
}  // namespace

int
b2r_resize_filter(const b2r_image* dst, const b2r_image* src,
                  const b2r_filter2d* filter, const b2r_roi* roi,
                  const b2r_options* opt, b2r_report* report, char* err,
                  size_t errlen)
{
    if (!filter || !filter->f) {
        set_err(err, errlen, "resize: null filter");
        return B2R_ERR_INVALID;
    }
    return resize_entry(dst, src, *filter->f, roi, opt, report, err, errlen);
}

int
b2r_resize(const b2r_image* dst, const b2r_image* src, const char* filtername,
           float filterwidth, const b2r_roi* roi, const b2r_options* opt,
           b2r_report* report, char* err, size_t errlen)
{
    int rc = check_image(src, true, err, errlen);
    if (rc)
        return rc;
    rc = check_image(dst, false, err, errlen);
    if (rc)
        return rc;
    // get_resize_filter, imagebufalgo_xform.cpp:830-860
    float wratio = float(dst->full_width) / float(src->full_width);
    float hratio = float(dst->full_height) / float(src->full_height);
    std::string name = filtername ? filtername : "";
    if (name.empty())
        name = (wratio > 1.0f || hratio > 1.0f) ? "blackman-harris" : "lanczos3";
    std::unique_ptr<Filter2D, void (*)(Filter2D*)> filter(nullptr,
                                                          Filter2D::destroy);
    for (int i = 0, e = Filter2D::num_filters(); i < e; ++i) {
        const FilterDesc& fd = Filter2D::get_filterdesc(i);
        if (name == fd.name) {
            float w = filterwidth > 0.0f ? filterwidth
                                         : fd.width * std::max(1.0f, wratio);
            float h = filterwidth > 0.0f ? filterwidth
                                         : fd.width * std::max(1.0f, hratio);
            filter.reset(Filter2D::create(name, w, h));
            break;
        }
    }
    if (!filter) {
        set_err(err, errlen, "Filter \"%s\" not recognized", name.c_str());
        return B2R_ERR_FILTER;
    }
    return resize_entry(dst, src, *filter, roi, opt, report, err, errlen);
}

// A batch of frames of ONE shape: a single persistent launch of the stream
// kernel walks (frame, strip, band) work items -- no launch gaps, no ramp-up
// and tail per frame, which is what bounds small frames (a 4K -> 1080p uint16
// frame is 11 us of HBM time).  Frames that do not qualify (host memory,
// shapes of other kernels, differing geometry) run one by one.
int
b2r_resize_batch(const b2r_image* dst, const b2r_image* src, int nframes,
                 const char* filtername, float filterwidth, const b2r_roi* roi,
                 const b2r_options* opt, b2r_report* report, char* err, size_t errlen)
{
    if (report)
        memset(report, 0, sizeof(*report));
    if (nframes < 0 || (nframes > 0 && (!dst || !src))) {
        set_err(err, errlen, "resize_batch: bad frame arrays");
        return B2R_ERR_INVALID;
    }
    if (nframes == 0)
        return B2R_OK;
    auto same_shape = [](const b2r_image& a, const b2r_image& b) {
        return a.basetype == b.basetype && a.nchannels == b.nchannels && a.x == b.x && a.y == b.y
               && a.width == b.width && a.height == b.height && a.full_x == b.full_x
               && a.full_y == b.full_y && a.full_width == b.full_width
               && a.full_height == b.full_height && a.xstride == b.xstride
               && a.ystride == b.ystride;
    };
    bool one_launch = nframes > 1 && !(opt && opt->highlightcomp)
                      && (!opt || opt->path == B2R_PATH_AUTO || opt->path == B2R_PATH_FUSED);
    int dev0 = opt ? opt->device : 0;
    for (int i = 0; i < nframes && one_launch; ++i) {
        if (!dst[i].pixels || !src[i].pixels || !same_shape(dst[i], dst[0])
            || !same_shape(src[i], src[0]))
            one_launch = false;
        int ds = dev0, dd = dev0;
        if (one_launch
            && (memspace_of(&src[i], &ds) != B2R_MEM_DEVICE
                || memspace_of(&dst[i], &dd) != B2R_MEM_DEVICE || ds != dd
                || (i > 0 && ds != dev0)))
            one_launch = false;
        if (i == 0)
            dev0 = ds;
        // TMA: 16-byte aligned rows (frame 0 could be re-pitched, a batch is not)
        if (one_launch
            && ((((uintptr_t)src[i].pixels) & 15) != 0 || (src[i].ystride & 15) != 0))
            one_launch = false;
    }
    if (one_launch) {
        BatchCapture cap;
        cap.want       = true;
        g_capture      = &cap;
        g_batch_frames = nframes;
        int rc = b2r_resize(&dst[0], &src[0], filtername, filterwidth, roi, opt, nullptr, err,
                            errlen);
        g_capture      = nullptr;
        g_batch_frames = 1;
        if (rc)
            return rc;
        if (cap.got) {
            DeviceGuard guard(dev0);
            cudaStream_t stream = opt ? (cudaStream_t)opt->cuda_stream : nullptr;
            FusedParams fp      = cap.fp;
            const long long doff = (const unsigned char*)fp.dst
                                   - (const unsigned char*)dst[0].pixels;
            // per frame: the tensor map over its source rows, its ROI origin
            const size_t map_bytes = size_t(nframes) * 128;
            std::vector<unsigned char> host(map_bytes + 64 + size_t(nframes) * sizeof(void*));
            unsigned char* maps_h
                = (unsigned char*)(((uintptr_t)host.data() + 63) & ~uintptr_t(63));
            std::vector<unsigned char*> dsts(nframes);
            for (int i = 0; i < nframes; ++i) {
                FusedParams q = fp;
                q.src         = (const unsigned char*)src[i].pixels;
                if (stream_encode_map(maps_h + size_t(i) * 128, q, cap.w2, src[i].pixels,
                                      src[i].height)
                    != 0) {
                    set_err(err, errlen, "resize_batch: cannot encode the tensor map of frame %d",
                            i);
                    return B2R_ERR_CUDA;
                }
                dsts[i] = (unsigned char*)dst[i].pixels + doff;
            }
            void* scratch = nullptr;
            CUDA_TRY(cudaMallocAsync(&scratch, map_bytes + size_t(nframes) * sizeof(void*),
                                     stream));
            cudaEvent_t ev0 = nullptr, ev1 = nullptr;
            CUDA_TRY(cudaMemcpyAsync(scratch, maps_h, map_bytes, cudaMemcpyHostToDevice, stream));
            CUDA_TRY(cudaMemcpyAsync((unsigned char*)scratch + map_bytes, dsts.data(),
                                     size_t(nframes) * sizeof(void*), cudaMemcpyHostToDevice,
                                     stream));
            if (report) {
                cudaEventCreate(&ev0);
                cudaEventCreate(&ev1);
                cudaEventRecord(ev0, stream);
            }
            fp.nframes    = nframes;
            fp.batch_maps = scratch;
            fp.batch_dst  = (unsigned char* const*)((unsigned char*)scratch + map_bytes);
            int lrc = launch_resize_stream_batch(fp, cap.w2, cap.sms, stream);
            cudaError_t e = cudaGetLastError();
            if (report)
                cudaEventRecord(ev1, stream);
            cudaFreeAsync(scratch, stream);
            if (lrc != 0 || e != cudaSuccess) {
                set_err(err, errlen, "resize_batch: launch failed (%s)",
                        e != cudaSuccess ? cudaGetErrorString(e) : "no kernel");
                return B2R_ERR_CUDA;
            }
            if (report) {
                CUDA_TRY(cudaStreamSynchronize(stream));
                float ms = 0.0f;
                cudaEventElapsedTime(&ms, ev0, ev1);
                cudaEventDestroy(ev0);
                cudaEventDestroy(ev1);
                report->path_used        = B2R_PATH_FUSED;
                report->kernels_launched = 1;
                report->kernel_ms        = ms;
                report->fused_vmode      = fp.vk;
                report->fused_htaps      = fp.ht;
            }
            return B2R_OK;
        }
        // frame 0 took another kernel (and is done): the rest one by one
        int launched = 0;
        for (int i = 1; i < nframes; ++i) {
            b2r_report r;
            rc = b2r_resize(&dst[i], &src[i], filtername, filterwidth, roi, opt,
                            report ? &r : nullptr, err, errlen);
            if (rc)
                return rc;
            if (report) {
                launched += r.kernels_launched;
                *report = r;
            }
        }
        if (report)
            report->kernels_launched = launched;
        return B2R_OK;
    }
    int launched = 0;
    for (int i = 0; i < nframes; ++i) {
        b2r_report r;
        int rc = b2r_resize(&dst[i], &src[i], filtername, filterwidth, roi, opt,
                            report ? &r : nullptr, err, errlen);
        if (rc)
            return rc;
        if (report) {
            launched += r.kernels_launched;
            const int64_t h2d = report->h2d_bytes + r.h2d_bytes, d2h = report->d2h_bytes + r.d2h_bytes;
            *report            = r;
            report->h2d_bytes  = h2d;
            report->d2h_bytes  = d2h;
        }
    }
    if (report)
        report->kernels_launched = launched;
    return B2R_OK;
}

int
b2r_band_split(const b2r_roi* roi, int nbands, int band, b2r_roi* out)
{
    if (!roi || !out || nbands <= 0 || band < 0 || band >= nbands)
        return B2R_ERR_INVALID;
    long long h = (long long)roi->yend - roi->ybegin;
    *out        = *roi;
    out->ybegin = roi->ybegin + int(h * band / nbands);
    out->yend   = roi->ybegin + int(h * (band + 1) / nbands);
    return B2R_OK;
}

int
b2r_resize_banded(const b2r_image* dst, const b2r_image* src, const char* filtername,
                  float filterwidth, const b2r_roi* roi_in, const int* devices, int ndevices,
                  const b2r_options* opt, char* err, size_t errlen)
{
    if (!dst || !src || ndevices < 1 || ndevices > 64) {
        set_err(err, errlen, "resize_banded: bad arguments");
        return B2R_ERR_INVALID;
    }
    int d0 = 0, d1 = 0;
    if (memspace_of(dst, &d0) != B2R_MEM_HOST || memspace_of(src, &d1) != B2R_MEM_HOST) {
        set_err(err, errlen, "resize_banded shards host images; device-resident images are "
                             "already on one GPU");
        return B2R_ERR_UNSUPPORTED;
    }
    const int ndev_all = device_count_cached();
    if (ndev_all <= 0) {
        set_err(err, errlen, "no CUDA device available: the B200 resize path has no CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    b2r_roi roi = roi_in ? *roi_in
                         : b2r_roi { dst->x, dst->x + dst->width, dst->y, dst->y + dst->height, 0,
                                     dst->nchannels };
    std::vector<int> rcs(ndevices, B2R_OK);
    std::vector<std::string> errs(ndevices);
    auto work = [&](int g) {
        b2r_roi band;
        if (b2r_band_split(&roi, ndevices, g, &band) || band.yend <= band.ybegin)
            return;
        b2r_options o;
        memset(&o, 0, sizeof(o));
        if (opt)
            o = *opt;
        o.device      = devices ? devices[g] : g;
        o.cuda_stream = nullptr;
        if (o.device < 0 || o.device >= ndev_all) {
            rcs[g]  = B2R_ERR_INVALID;
            errs[g] = "invalid CUDA device ordinal " + std::to_string(o.device);
            return;
        }
        char e[512] = "";
        rcs[g] = b2r_resize(dst, src, filtername, filterwidth, &band, &o, nullptr, e, sizeof(e));
        if (rcs[g])
            errs[g] = e;
    };
    std::vector<std::thread> th;
    int started = 1;
    try {
        for (; started < ndevices; ++started)
            th.emplace_back(work, started);
    } catch (...) {
    }
    work(0);
    for (int g = started; g < ndevices; ++g)
        work(g);  // threads that could not be had
    for (auto& t : th)
        t.join();
    for (int g = 0; g < ndevices; ++g)
        if (rcs[g]) {
            set_err(err, errlen, "%s", errs[g].c_str());
            return rcs[g];
        }
    return B2R_OK;
}

int
b2r_source_rows(const b2r_image* dst, const b2r_image* src, const char* filtername,
                float filterwidth, const b2r_roi* roi_in, int* row_begin,
                int* row_end, char* err, size_t errlen)
{
    if (!dst || !src || dst->full_width <= 0 || dst->full_height <= 0
        || src->full_width <= 0 || src->full_height <= 0 || src->width <= 0
        || src->height <= 0) {
        set_err(err, errlen, "Uninitialized input image");
        return B2R_ERR_INVALID;
    }
    float wratio = float(dst->full_width) / float(src->full_width);
    float hratio = float(dst->full_height) / float(src->full_height);
    std::string name = filtername ? filtername : "";
    if (name.empty())
        name = (wratio > 1.0f || hratio > 1.0f) ? "blackman-harris" : "lanczos3";
    std::unique_ptr<Filter2D, void (*)(Filter2D*)> filter(nullptr,
                                                          Filter2D::destroy);
    for (int i = 0, e = Filter2D::num_filters(); i < e; ++i) {
        const FilterDesc& fd = Filter2D::get_filterdesc(i);
        if (name == fd.name) {
            float w = filterwidth > 0.0f ? filterwidth
                                         : fd.width * std::max(1.0f, wratio);
            float h = filterwidth > 0.0f ? filterwidth
                                         : fd.width * std::max(1.0f, hratio);
            filter.reset(Filter2D::create(name, w, h));
            break;
        }
    }
    if (!filter) {
        set_err(err, errlen, "Filter \"%s\" not recognized", name.c_str());
        return B2R_ERR_FILTER;
    }
    b2r_roi roi = roi_in ? *roi_in
                         : b2r_roi { dst->x, dst->x + dst->width, dst->y,
                                     dst->y + dst->height, 0, dst->nchannels };
    // only the tap geometry is needed (centres and radii), for a 1-wide column
    AxisGeom gx { roi.xbegin, roi.xbegin, dst->full_x, dst->full_width, src->full_x,
                  src->full_width, src->x, src->x + src->width };
    AxisGeom gy { roi.ybegin, roi.yend, dst->full_y, dst->full_height, src->full_y,
                  src->full_height, src->y, src->y + src->height };
    AxisTaps tx = build_axis_taps(gx, *filter, false);
    AxisTaps ty = build_axis_taps(gy, *filter, true);
    int r0 = 0, r1 = 0;
    source_row_span(*src, *filter, tx, ty, &r0, &r1);
    if (row_begin)
        *row_begin = r0;
    if (row_end)
        *row_end = r1 + 1;
    return B2R_OK;
}

void
b2r_fit_geometry(int src_full_width, int src_full_height, int fit_width,
                 int fit_height, const char* fillmode, int* resize_width,
                 int* resize_height, int* xoffset, int* yoffset)
{
    // imagebufalgo_xform.cpp:962-997
    float oldaspect = float(src_full_width) / src_full_height;
    float newaspect = float(fit_width) / fit_height;
    int rw = fit_width, rh = fit_height, xo = 0, yo = 0;
    std::string mode = fillmode ? fillmode : "letterbox";
    if (mode != "height" && mode != "width")
        mode = "letterbox";
    if (mode == "letterbox")
        mode = (newaspect >= oldaspect) ? "height" : "width";
    if (mode == "height") {
        rw = int(rh * oldaspect + 0.5f);
        xo = (fit_width - rw) / 2;
    } else {
        rh = int(rw / oldaspect + 0.5f);
        yo = (fit_height - rh) / 2;
    }
    if (resize_width)
        *resize_width = rw;
    if (resize_height)
        *resize_height = rh;
    if (xoffset)
        *xoffset = xo;
    if (yoffset)
        *yoffset = yo;
}

}  // extern "C"

// =========================================================================
// rangecompress / rangeexpand
// =========================================================================
namespace {

// Device-side worker shared by the two entry points and by the
// highlight-compensated resize: images already resident on `device`.
void
range_on_device(bool expand, const b2r_image& dst, void* dpix, const b2r_image& src,
                const void* spix, const b2r_roi& roi, bool useluma, int alpha_channel,
                int z_channel, cudaStream_t stream)
{
    RangeParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.dst    = to_dev(dst, dpix);
    rp.src    = to_dev(src, const_cast<void*>(spix));
    rp.roi_x0 = roi.xbegin;
    rp.roi_x1 = roi.xend;
    rp.roi_y0 = roi.ybegin;
    rp.roi_y1 = roi.yend;
    rp.chbegin = roi.chbegin;
    rp.chend   = roi.chend;
    rp.expand  = expand ? 1 : 0;
    // "No way to use luma" (pixelmath.cpp:614-621, 682-689)
    const int nc = roi.chend - roi.chbegin;
    if (nc < 3 || (alpha_channel >= roi.chbegin && alpha_channel < roi.chbegin + 3)
        || (z_channel >= roi.chbegin && z_channel < roi.chbegin + 3))
        useluma = false;
    rp.useluma       = useluma ? 1 : 0;
    rp.alpha_channel = alpha_channel;
    rp.z_channel     = z_channel;
    rp.in_place      = (dpix == spix) ? 1 : 0;
    launch_range(rp, stream);
}

int
range_impl(bool expand, const b2r_image* dst_in, const b2r_image* src_in, int useluma,
           int alpha_channel, int z_channel, const b2r_roi* roi_in,
           const b2r_options* opt, b2r_report* report, char* err, size_t errlen)
{
    if (report)
        memset(report, 0, sizeof(*report));
    int rc = check_image(src_in, true, err, errlen);
    if (rc)
        return rc;
    rc = check_image(dst_in, false, err, errlen);
    if (rc)
        return rc;
    const b2r_image& dst = *dst_in;
    const b2r_image& src = *src_in;
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 range compression path has no "
                "CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    b2r_roi roi;
    if (roi_in) {
        roi         = *roi_in;
        roi.xbegin  = std::max(roi.xbegin, dst.x);
        roi.xend    = std::min(roi.xend, dst.x + dst.width);
        roi.ybegin  = std::max(roi.ybegin, dst.y);
        roi.yend    = std::min(roi.yend, dst.y + dst.height);
        roi.chbegin = std::max(roi.chbegin, 0);
        roi.chend   = std::min(roi.chend, dst.nchannels);
    } else {
        roi = b2r_roi { dst.x, dst.x + dst.width, dst.y, dst.y + dst.height, 0,
                        dst.nchannels };
    }
    roi.chend = std::min(roi.chend, src.nchannels);  // IBAprep_CLAMP_MUTUAL_NCHANNELS
    if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin || roi.chend <= roi.chbegin)
        return B2R_OK;
    const int roi_w = roi.xend - roi.xbegin, roi_h = roi.yend - roi.ybegin;
    const int ses = basetype_size(src.basetype);
    const int des = basetype_size(dst.basetype);

    int sdev = opt ? opt->device : 0, ddev = sdev;
    const int smem = memspace_of(&src, &sdev);
    const int dmem = memspace_of(&dst, &ddev);
    int device     = opt ? opt->device : 0;
    if (smem == B2R_MEM_DEVICE)
        device = sdev;
    else if (dmem == B2R_MEM_DEVICE)
        device = ddev;
    if (device < 0 || device >= device_count_cached()) {
        set_err(err, errlen, "invalid CUDA device ordinal %d", device);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(device);
    device_info(device);
    cudaStream_t stream = opt ? (cudaStream_t)opt->cuda_stream : nullptr;

    // host images: stage the part of the ROI that exists on each side
    b2r_image srcd = src, dstd = dst;
    void* dsrc = src.pixels;
    void* ddst = dst.pixels;
    void *src_alloc = nullptr, *dst_alloc = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    ScopeCleanup cleanup { stream, { &src_alloc, &dst_alloc, nullptr }, { &ev0, &ev1 } };
    size_t dpitch = 0, drowbytes = 0;
    if (smem == B2R_MEM_HOST) {
        if (src.xstride != (int64_t)src.nchannels * ses) {
            set_err(err, errlen, "host source pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        const int x0 = std::max(roi.xbegin, src.x), x1 = std::min(roi.xend, src.x + src.width);
        const int y0 = std::max(roi.ybegin, src.y), y1 = std::min(roi.yend, src.y + src.height);
        srcd.x = x0, srcd.y = y0;
        srcd.width = std::max(0, x1 - x0), srcd.height = std::max(0, y1 - y0);
        if (srcd.width > 0 && srcd.height > 0) {
            size_t rowbytes = size_t(srcd.width) * src.xstride;
            size_t pitch    = (rowbytes + 15) & ~size_t(15);
            CUDA_TRY(cudaMallocAsync(&src_alloc, pitch * srcd.height, stream));
            const unsigned char* hp = (const unsigned char*)src.pixels
                                      + (long long)(y0 - src.y) * src.ystride
                                      + (long long)(x0 - src.x) * src.xstride;
            CUDA_TRY(copy_h2d_2d(src_alloc, pitch, hp, src.ystride, rowbytes,
                                       srcd.height, stream));
            srcd.ystride = (int64_t)pitch;
            if (report)
                report->h2d_bytes += (int64_t)rowbytes * srcd.height;
        } else {
            srcd.width = srcd.height = 0;  // every pixel reads as black
            CUDA_TRY(cudaMallocAsync(&src_alloc, 16, stream));
        }
        dsrc = src_alloc;
    }
    unsigned char* host_out = nullptr;
    if (dmem == B2R_MEM_HOST) {
        if (dst.xstride != (int64_t)dst.nchannels * des) {
            set_err(err, errlen, "host destination pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        drowbytes = size_t(roi_w) * dst.xstride;
        dpitch    = (drowbytes + 15) & ~size_t(15);
        CUDA_TRY(cudaMallocAsync(&dst_alloc, dpitch * roi_h, stream));
        host_out = (unsigned char*)dst.pixels + (long long)(roi.ybegin - dst.y) * dst.ystride
                   + (long long)(roi.xbegin - dst.x) * dst.xstride;
        if (roi.chbegin != 0 || roi.chend != dst.nchannels) {
            // channels outside the range keep what dst holds
            CUDA_TRY(copy_h2d_2d(dst_alloc, dpitch, host_out, dst.ystride, drowbytes,
                                       roi_h, stream));
            if (report)
                report->h2d_bytes += (int64_t)drowbytes * roi_h;
        }
        ddst         = dst_alloc;
        dstd.x       = roi.xbegin;
        dstd.y       = roi.ybegin;
        dstd.width   = roi_w;
        dstd.height  = roi_h;
        dstd.ystride = (int64_t)dpitch;
    }
    const bool timed = report != nullptr;
    if (timed) {
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventRecord(ev0, stream);
    }
    range_on_device(expand, dstd, ddst, srcd, dsrc, roi, useluma != 0, alpha_channel,
                    z_channel, stream);
    if (timed)
        cudaEventRecord(ev1, stream);
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA launch error: %s", cudaGetErrorString(e));
            return B2R_ERR_CUDA;
        }
    }
    if (dst_alloc) {
        CUDA_TRY(copy_d2h_2d(host_out, dst.ystride, dst_alloc, dpitch, drowbytes,
                                   roi_h, stream));
        if (report)
            report->d2h_bytes += (int64_t)drowbytes * roi_h;
    }
    const bool staged = dst_alloc || src_alloc;
    for (void** b : { &dst_alloc, &src_alloc })
        if (*b) {
            void* q = *b;
            *b      = nullptr;
            CUDA_TRY(cudaFreeAsync(q, stream));
        }
    if (staged || timed)
        CUDA_TRY(cudaStreamSynchronize(stream));
    if (timed) {
        cudaEventElapsedTime(&report->kernel_ms, ev0, ev1);
        report->kernels_launched = 1;
    }
    return B2R_OK;
}

}  // namespace

namespace {
// oiiotool's ":highlightcomp=1" recipe (oiiotool.cpp:4644-4663) on the device:
// tmp = rangecompress(src) in src's own type, resize(dst, tmp),
// rangeexpand(dst, dst) over the ROI just written.
int
resize_highlightcomp(const b2r_image* dst_in, const b2r_image* src_in, const Filter2D& filter,
                     const b2r_roi* roi_in, const b2r_options* opt, b2r_report* report,
                     char* err, size_t errlen)
{
    if (report)
        memset(report, 0, sizeof(*report));
    int rc = check_image(src_in, true, err, errlen);
    if (rc)
        return rc;
    rc = check_image(dst_in, false, err, errlen);
    if (rc)
        return rc;
    const b2r_image& dst = *dst_in;
    const b2r_image& src = *src_in;
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 resize path has no CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    b2r_roi roi;
    if (roi_in) {
        roi        = *roi_in;
        roi.xbegin = std::max(roi.xbegin, dst.x);
        roi.xend   = std::min(roi.xend, dst.x + dst.width);
        roi.ybegin = std::max(roi.ybegin, dst.y);
        roi.yend   = std::min(roi.yend, dst.y + dst.height);
    } else {
        roi = b2r_roi { dst.x, dst.x + dst.width, dst.y, dst.y + dst.height, 0,
                        dst.nchannels };
    }
    roi.chbegin = 0;
    roi.chend   = dst.nchannels;
    if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin)
        return B2R_OK;
    const int roi_w = roi.xend - roi.xbegin, roi_h = roi.yend - roi.ybegin;
    const int ses = basetype_size(src.basetype), des = basetype_size(dst.basetype);
    int sdev = opt->device, ddev = sdev;
    const int smem = memspace_of(&src, &sdev);
    const int dmem = memspace_of(&dst, &ddev);
    int device     = opt->device;
    if (smem == B2R_MEM_DEVICE)
        device = sdev;
    else if (dmem == B2R_MEM_DEVICE)
        device = ddev;
    if (device < 0 || device >= device_count_cached()) {
        set_err(err, errlen, "invalid CUDA device ordinal %d", device);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(device);
    device_info(device);
    cudaStream_t stream = (cudaStream_t)opt->cuda_stream;
    const int alpha = opt->alpha_channel1 - 1, z = opt->z_channel1 - 1;

    b2r_options sub = *opt;
    sub.highlightcomp = 0;
    sub.device        = device;
    if (opt->reserved[0] == 1)  // tables only (MIP chain pass 1)
        return resize_impl(dst_in, src_in, filter, roi_in, &sub, report, err, errlen);

    // float / half images on the device: compress on load, expand on store,
    // inside the stream kernel (one pass over HBM instead of three)
    {
        b2r_options fused = sub;
        // highlightcomp == 2 (private to the MIP chain): the level is clamped
        // afterwards (maketexture.cpp:936-938) -- in the kernel's store when the
        // call is fused, by clamp_kernel after the three passes otherwise
        fused.reserved[3] = opt->highlightcomp == 2 ? 2 : 1;
        rc = resize_impl(dst_in, src_in, filter, roi_in, &fused, report, err, errlen);
        if (rc != HC_NOT_FUSED)
            return rc;
        rc = B2R_OK;
    }

    // compressed copy of the source, same type and windows, on the device
    b2r_image T = src;
    void* tpix  = nullptr;
    const size_t trow = (size_t(src.width) * src.nchannels * ses + 15) & ~size_t(15);
    CUDA_TRY(cudaMallocAsync(&tpix, trow * src.height, stream));
    T.pixels   = tpix;
    T.xstride  = (int64_t)src.nchannels * ses;
    T.ystride  = (int64_t)trow;
    T.memspace = B2R_MEM_DEVICE;
    T.device   = device;
    b2r_report sub_report;
    rc = b2r_rangecompress(&T, &src, 0, alpha, z, nullptr, &sub, report ? &sub_report : nullptr,
                           err, errlen);
    if (rc == B2R_OK && report)
        report->h2d_bytes += sub_report.h2d_bytes;

    // the destination ROI on the device
    b2r_image D = dst;
    void* dpix  = nullptr;
    if (rc == B2R_OK && dmem == B2R_MEM_HOST) {
        if (dst.xstride != (int64_t)dst.nchannels * des) {
            set_err(err, errlen, "host destination pixels must be contiguous in x");
            rc = B2R_ERR_UNSUPPORTED;
        } else {
            const size_t drow = (size_t(roi_w) * dst.xstride + 15) & ~size_t(15);
            cudaError_t e = cudaMallocAsync(&dpix, drow * roi_h, stream);
            if (e != cudaSuccess) {
                set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
                rc = B2R_ERR_CUDA;
            }
            D.pixels   = dpix;
            D.x        = roi.xbegin;
            D.y        = roi.ybegin;
            D.width    = roi_w;
            D.height   = roi_h;
            D.ystride  = (int64_t)drow;
            D.memspace = B2R_MEM_DEVICE;
            D.device   = device;
        }
    }
    if (rc == B2R_OK) {
        rc = resize_impl(&D, &T, filter, &roi, &sub, report ? &sub_report : nullptr, err,
                         errlen);
        if (rc == B2R_OK && report) {
            int64_t h2d = report->h2d_bytes;
            *report             = sub_report;
            report->h2d_bytes  += h2d;
            report->kernels_launched += 2;
        }
    }
    if (rc == B2R_OK)
        range_on_device(true, D, D.pixels, D, D.pixels, roi, false, alpha, z, stream);
    if (rc == B2R_OK && opt->highlightcomp == 2 && !dpix && D.basetype == B2R_FLOAT
        && D.xstride == (int64_t)D.nchannels * 4
        && D.ystride == (int64_t)D.width * D.nchannels * 4 && roi_w == D.width
        && roi_h == D.height)
        launch_clamp((float*)D.pixels, (long long)D.width * D.height * D.nchannels,
                     D.nchannels, alpha, 0.0f, 3.402823466e+38f, stream);
    if (rc == B2R_OK && dpix) {
        unsigned char* hp = (unsigned char*)dst.pixels
                            + (long long)(roi.ybegin - dst.y) * dst.ystride
                            + (long long)(roi.xbegin - dst.x) * dst.xstride;
        cudaError_t e = copy_d2h_2d(hp, dst.ystride, dpix, D.ystride,
                                          size_t(roi_w) * dst.xstride, roi_h, stream);
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
            rc = B2R_ERR_CUDA;
        } else if (report) {
            report->d2h_bytes += (int64_t)roi_w * dst.xstride * roi_h;
        }
    }
    if (dpix)
        cudaFreeAsync(dpix, stream);
    cudaFreeAsync(tpix, stream);
    if (rc == B2R_OK && (dmem == B2R_MEM_HOST || smem == B2R_MEM_HOST))
        CUDA_TRY(cudaStreamSynchronize(stream));
    return rc;
}
}  // namespace

extern "C" int
b2r_rangecompress(const b2r_image* dst, const b2r_image* src, int useluma,
                  int alpha_channel, int z_channel, const b2r_roi* roi,
                  const b2r_options* opt, b2r_report* report, char* err, size_t errlen)
{
    return range_impl(false, dst, src, useluma, alpha_channel, z_channel, roi, opt, report,
                      err, errlen);
}

extern "C" int
b2r_rangeexpand(const b2r_image* dst, const b2r_image* src, int useluma, int alpha_channel,
                int z_channel, const b2r_roi* roi, const b2r_options* opt,
                b2r_report* report, char* err, size_t errlen)
{
    return range_impl(true, dst, src, useluma, alpha_channel, z_channel, roi, opt, report, err,
                      errlen);
}

// =========================================================================
// resample
// =========================================================================
extern "C" int
b2r_resample(const b2r_image* dst_in, const b2r_image* src_in, int interpolate,
             const b2r_roi* roi_in, const b2r_options* opt, b2r_report* report,
             char* err, size_t errlen)
{
    if (report)
        memset(report, 0, sizeof(*report));
    int rc = check_image(src_in, true, err, errlen);
    if (rc)
        return rc;
    rc = check_image(dst_in, false, err, errlen);
    if (rc)
        return rc;
    const b2r_image& dst = *dst_in;
    const b2r_image& src = *src_in;
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 resample path has no CPU "
                "fallback");
        return B2R_ERR_NO_DEVICE;
    }
    b2r_roi roi;
    if (roi_in) {
        roi        = *roi_in;
        roi.xbegin = std::max(roi.xbegin, dst.x);
        roi.xend   = std::min(roi.xend, dst.x + dst.width);
        roi.ybegin = std::max(roi.ybegin, dst.y);
        roi.yend   = std::min(roi.yend, dst.y + dst.height);
        roi.chbegin = std::max(roi.chbegin, 0);
        roi.chend   = std::min(roi.chend, dst.nchannels);
    } else {
        roi = b2r_roi { dst.x, dst.x + dst.width, dst.y, dst.y + dst.height, 0,
                        dst.nchannels };
    }
    roi.chend = std::min(roi.chend, src.nchannels);
    if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin || roi.chend <= roi.chbegin)
        return B2R_OK;
    const int roi_w = roi.xend - roi.xbegin, roi_h = roi.yend - roi.ybegin;
    const int ses = basetype_size(src.basetype);
    const int des = basetype_size(dst.basetype);

    int sdev = opt ? opt->device : 0, ddev = sdev;
    const int smem = memspace_of(&src, &sdev);
    const int dmem = memspace_of(&dst, &ddev);
    int device     = opt ? opt->device : 0;
    if (smem == B2R_MEM_DEVICE)
        device = sdev;
    else if (dmem == B2R_MEM_DEVICE)
        device = ddev;
    if (device < 0 || device >= device_count_cached()) {
        set_err(err, errlen, "invalid CUDA device ordinal %d", device);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(device);
    device_info(device);
    cudaStream_t stream = opt ? (cudaStream_t)opt->cuda_stream : nullptr;

    // non-native (dst,src) pair: float temporary + convert back, as above
    if (!legal_pair(dst.basetype, src.basetype)) {
        if (dst.xstride != (int64_t)dst.nchannels * des) {
            set_err(err, errlen, "destination pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        if (roi.chbegin != 0 || roi.chend < dst.nchannels) {
            set_err(err, errlen,
                    "resample: a channel sub-range needs a native (dst,src) pair");
            return B2R_ERR_TYPES;
        }
        void* ftmp = nullptr;
        const size_t frow = size_t(roi_w) * dst.nchannels * 4;
        CUDA_TRY(cudaMallocAsync(&ftmp, frow * roi_h, stream));
        b2r_image F = dst;
        F.pixels    = ftmp;
        F.basetype  = B2R_FLOAT;
        F.x         = roi.xbegin;
        F.y         = roi.ybegin;
        F.width     = roi_w;
        F.height    = roi_h;
        F.xstride   = (int64_t)dst.nchannels * 4;
        F.ystride   = (int64_t)frow;
        F.memspace  = B2R_MEM_DEVICE;
        F.device    = device;
        b2r_options sub;
        memset(&sub, 0, sizeof(sub));
        sub.device      = device;
        sub.cuda_stream = (void*)stream;
        rc = b2r_resample(&F, &src, interpolate, &roi, &sub, nullptr, err, errlen);
        if (rc == B2R_OK) {
            const size_t drow = size_t(roi_w) * dst.xstride;
            unsigned char* out = (unsigned char*)dst.pixels
                                 + (long long)(roi.ybegin - dst.y) * dst.ystride
                                 + (long long)(roi.xbegin - dst.x) * dst.xstride;
            if (dmem == B2R_MEM_DEVICE) {
                launch_convert_from_float((const float*)ftmp, (long long)frow, out,
                                          dst.ystride, roi_w * dst.nchannels, roi_h,
                                          dst.basetype, stream);
            } else {
                void* dtmp = nullptr;
                CUDA_TRY(cudaMallocAsync(&dtmp, drow * roi_h, stream));
                launch_convert_from_float((const float*)ftmp, (long long)frow, dtmp,
                                          (long long)drow, roi_w * dst.nchannels,
                                          roi_h, dst.basetype, stream);
                CUDA_TRY(copy_d2h_2d(out, dst.ystride, dtmp, drow, drow, roi_h, stream));
                CUDA_TRY(cudaFreeAsync(dtmp, stream));
            }
        }
        cudaFreeAsync(ftmp, stream);
        if (rc == B2R_OK && (dmem == B2R_MEM_HOST || smem == B2R_MEM_HOST))
            CUDA_TRY(cudaStreamSynchronize(stream));
        return rc;
    }

    void* dsrc     = src.pixels;
    void* ddst     = dst.pixels;
    b2r_image srcd = src, dstd = dst;
    bool src_staged = false, dst_staged = false;
    size_t dpitch = 0, drowbytes = 0;
    if (smem == B2R_MEM_HOST) {
        if (src.xstride != (int64_t)src.nchannels * ses) {
            set_err(err, errlen, "host source pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        size_t rowbytes = size_t(src.width) * src.xstride;
        size_t pitch    = (rowbytes + 15) & ~size_t(15);
        CUDA_TRY(cudaMallocAsync(&dsrc, pitch * src.height, stream));
        src_staged = true;
        CUDA_TRY(copy_h2d_2d(dsrc, pitch, src.pixels, src.ystride, rowbytes,
                                   src.height, stream));
        srcd.ystride = (int64_t)pitch;
        if (report)
            report->h2d_bytes += (int64_t)rowbytes * src.height;
    }
    if (dmem == B2R_MEM_HOST) {
        if (dst.xstride != (int64_t)dst.nchannels * des) {
            set_err(err, errlen, "host destination pixels must be contiguous in x");
            if (src_staged)
                cudaFreeAsync(dsrc, stream);
            return B2R_ERR_UNSUPPORTED;
        }
        drowbytes = size_t(roi_w) * dst.xstride;
        dpitch    = (drowbytes + 15) & ~size_t(15);
        CUDA_TRY(cudaMallocAsync(&ddst, dpitch * roi_h, stream));
        dst_staged = true;
        // channels outside [chbegin,chend) must survive: bring the ROI over
        if (roi.chbegin > 0 || roi.chend < dst.nchannels) {
            unsigned char* hp = (unsigned char*)dst.pixels
                                + (long long)(roi.ybegin - dst.y) * dst.ystride
                                + (long long)(roi.xbegin - dst.x) * dst.xstride;
            CUDA_TRY(copy_h2d_2d(ddst, dpitch, hp, dst.ystride, drowbytes,
                                       roi_h, stream));
        }
        dstd.x       = roi.xbegin;
        dstd.y       = roi.ybegin;
        dstd.width   = roi_w;
        dstd.height  = roi_h;
        dstd.ystride = (int64_t)dpitch;
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (report) {
        cudaEventCreate(&ev0);
        cudaEventCreate(&ev1);
        cudaEventRecord(ev0, stream);
    }
    ResampleParams rp;
    memset(&rp, 0, sizeof(rp));
    rp.dst          = to_dev(dstd, ddst);
    rp.src          = to_dev(srcd, dsrc);
    rp.roi_x0       = roi.xbegin;
    rp.roi_x1       = roi.xend;
    rp.roi_y0       = roi.ybegin;
    rp.roi_y1       = roi.yend;
    rp.nch          = roi.chend - roi.chbegin;
    rp.interpolate  = interpolate ? 1 : 0;
    rp.data_is_full = (src.x == src.full_x && src.y == src.full_y
                       && src.width == src.full_width
                       && src.height == src.full_height)
                          ? 1
                          : 0;
    launch_resample(rp, roi.chbegin, roi.chend, stream);
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA launch error: %s", cudaGetErrorString(e));
            return B2R_ERR_CUDA;
        }
    }
    if (report)
        cudaEventRecord(ev1, stream);
    if (dst_staged) {
        unsigned char* hp = (unsigned char*)dst.pixels
                            + (long long)(roi.ybegin - dst.y) * dst.ystride
                            + (long long)(roi.xbegin - dst.x) * dst.xstride;
        CUDA_TRY(copy_d2h_2d(hp, dst.ystride, ddst, dpitch, drowbytes, roi_h, stream));
        if (report)
            report->d2h_bytes += (int64_t)drowbytes * roi_h;
    }
    if (src_staged)
        CUDA_TRY(cudaFreeAsync(dsrc, stream));
    if (dst_staged)
        CUDA_TRY(cudaFreeAsync(ddst, stream));
    if (dst_staged || src_staged || report)
        CUDA_TRY(cudaStreamSynchronize(stream));
    if (report) {
        report->path_used        = B2R_PATH_EXACT;
        report->kernels_launched = 1;
        float ms                 = 0.0f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        report->kernel_ms = ms;
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
    }
    return B2R_OK;
}

// =========================================================================
// device-resident MIP chain (write_mipmap level loop, maketexture.cpp:823-986)
// =========================================================================

extern "C" int
b2r_mipchain_create(b2r_mipchain** out, const b2r_image* level0,
                    const char* filtername, int clamp_half, int alpha_channel,
                    const b2r_options* opt, char* err, size_t errlen)
{
    b2r_mip_options mo;
    memset(&mo, 0, sizeof(mo));
    mo.filtername    = filtername;
    mo.clamp_half    = clamp_half;
    mo.alpha_channel = alpha_channel;
    return b2r_mipchain_create_ex(out, level0, &mo, opt, err, errlen);
}

extern "C" int
b2r_mipchain_create_ex(b2r_mipchain** out, const b2r_image* level0,
                       const b2r_mip_options* mo, const b2r_options* opt, char* err,
                       size_t errlen)
{
    if (!mo) {
        set_err(err, errlen, "mipchain: null options");
        return B2R_ERR_INVALID;
    }
    const char* filtername  = mo->filtername;
    const int clamp_half    = mo->clamp_half;
    const int alpha_channel = mo->alpha_channel;
    const float sharpen     = mo->sharpen;
    const bool sharpen_first = mo->sharpen_first != 0;
    const std::string sharpenfilt = (mo->sharpenfilt && mo->sharpenfilt[0]) ? mo->sharpenfilt
                                                                             : "gaussian";
    if (sharpen > 0.0f) {
        if (sharpenfilt == "median") {
            set_err(err, errlen, "mipchain: the median sharpening kernel is not supported");
            return B2R_ERR_UNSUPPORTED;
        }
        ConvKernel K = make_conv_kernel(sharpenfilt, 3.0f, 3.0f);
        if (!K.known) {
            set_err(err, errlen, "%s", K.error.c_str());
            return B2R_ERR_FILTER;
        }
    }
    if (!out) {
        set_err(err, errlen, "mipchain: null output handle");
        return B2R_ERR_INVALID;
    }
    *out   = nullptr;
    int rc = check_image(level0, true, err, errlen);
    if (rc)
        return rc;
    if (level0->basetype != B2R_FLOAT) {
        set_err(err, errlen,
                "mipchain: levels are float32 (maketx:forcefloat); convert first");
        return B2R_ERR_TYPES;
    }
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 MIP path has no CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    std::string fname = (filtername && filtername[0]) ? filtername : "box";
    if (fname != "box") {
        bool known = false;
        for (int i = 0, e = Filter2D::num_filters(); i < e; ++i)
            known |= (fname == Filter2D::get_filterdesc(i).name);
        if (!known) {
            set_err(err, errlen, "Could not make filter \"%s\"", fname.c_str());
            return B2R_ERR_FILTER;
        }
    }
    int dev0      = opt ? opt->device : 0;
    const int mem = memspace_of(level0, &dev0);
    int device    = (mem == B2R_MEM_DEVICE) ? dev0 : (opt ? opt->device : 0);
    if (device < 0 || device >= device_count_cached()) {
        set_err(err, errlen, "invalid CUDA device ordinal %d", device);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(device);
    device_info(device);
    cudaStream_t stream = opt ? (cudaStream_t)opt->cuda_stream : nullptr;

    std::unique_ptr<b2r_mipchain> chain(new b2r_mipchain);
    chain->device = device;
    chain->nch    = level0->nchannels;
    const int nch = level0->nchannels;
    auto fail = [&](int code) {
        for (auto& l : chain->levels)
            if (l.owned && l.px)
                cudaFree(l.px);
        return code;
    };

    // level 0 on the device, contiguous
    b2r_mipchain::Level l0 { level0->width, level0->height, nullptr, false };
    const size_t row0 = size_t(level0->width) * nch * 4;
    if (mem == B2R_MEM_DEVICE && level0->xstride == (int64_t)nch * 4
        && level0->ystride == (int64_t)row0) {
        l0.px = level0->pixels;  // adopted, not owned
    } else {
        if (level0->xstride != (int64_t)nch * 4) {
            set_err(err, errlen, "mipchain: level 0 pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        if (cudaMalloc(&l0.px, row0 * level0->height) != cudaSuccess) {
            set_err(err, errlen, "mipchain: out of device memory for level 0");
            return B2R_ERR_CUDA;
        }
        l0.owned = true;
        chain->levels.push_back(l0);
        // host level 0 (pageable or pinned) or a strided device image
        cudaError_t e = (mem == B2R_MEM_HOST)
                            ? copy_h2d_2d(l0.px, row0, level0->pixels, level0->ystride, row0,
                                          level0->height, stream)
                            : cudaMemcpy2DAsync(l0.px, row0, level0->pixels, level0->ystride,
                                                row0, level0->height, cudaMemcpyDefault,
                                                stream);
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
            return fail(B2R_ERR_CUDA);
        }
        chain->levels.pop_back();
    }
    chain->levels.push_back(l0);

    // Pass 1: allocate every level and build + upload every weight table
    // (host work), so that pass 2 -- what kernel_ms times -- is back-to-back
    // launches with no host gaps and no allocation on the stream.
    {
        int w = l0.w, h = l0.h;
        while (w > 1 || h > 1) {
            const int sw = w > 1 ? w / 2 : w;
            const int sh = h > 1 ? h / 2 : h;
            b2r_mipchain::Level lv { sw, sh, nullptr, true };
            if (cudaMalloc(&lv.px, size_t(sw) * sh * nch * 4) != cudaSuccess) {
                set_err(err, errlen, "mipchain: out of device memory");
                return fail(B2R_ERR_CUDA);
            }
            chain->levels.push_back(lv);
            w = sw;
            h = sh;
        }
    }
    b2r_options sub;
    memset(&sub, 0, sizeof(sub));
    sub.device      = device;
    sub.path        = opt ? opt->path : B2R_PATH_AUTO;
    sub.cuda_stream = (void*)stream;
    // maketx --hicomp: only the filtered branch is wrapped (:889-935)
    // sharpening takes the filtered branch even for "box" (:880-881)
    // ... and so does an image that came in offset, cropped or overscanned
    // (orig_was_overscan, :700-703, :881): its levels go through resize() with
    // the "box" Filter2D, level 0 keeping its origin with display == data
    // (img->set_full, :876-877)
    const bool orig_was_overscan
        = level0->x || level0->y || level0->full_x || level0->full_y
          || level0->full_width != level0->width || level0->full_height != level0->height;
    const bool box_path = fname == "box" && !(sharpen > 0.0f) && !orig_was_overscan;
    const bool hicomp   = opt && opt->highlightcomp && !box_path
                        && chain->levels.size() > 1;
    const bool scratch_needed = (hicomp && sharpen > 0.0f)
                                || (sharpen > 0.0f && sharpen_first && !box_path
                                    && chain->levels.size() > 1);
    const int hc_z    = opt ? opt->z_channel1 - 1 : -1;
    void* hc_tmp      = nullptr;
    if (scratch_needed
        && cudaMalloc(&hc_tmp, size_t(l0.w) * l0.h * nch * 4) != cudaSuccess) {
        set_err(err, errlen, "mipchain: out of device memory for the highlight "
                             "compensation scratch image");
        return fail(B2R_ERR_CUDA);
    }
    struct ScratchFree {
        void* p;
        ~ScratchFree()
        {
            if (p)
                cudaFree(p);
        }
    } hc_free { hc_tmp };
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) {
            cudaEventCreate(&ev0);
            cudaEventCreate(&ev1);
            cudaEventRecord(ev0, stream);
        }
        sub.reserved[0] = (pass == 0) ? 1 : 0;  // private: plan only, no launch
        for (size_t li = 1; li < chain->levels.size(); ++li) {
            const b2r_mipchain::Level& big = chain->levels[li - 1];
            const b2r_mipchain::Level& lv  = chain->levels[li];
            const int w = big.w, h = big.h, sw = lv.w, sh = lv.h;
            // "Trick" of write_mipmap (:866-879): both images get zero-origin
            // windows with display == data
            b2r_image S;
            memset(&S, 0, sizeof(S));
            S.pixels    = big.px;
            S.basetype  = B2R_FLOAT;
            S.nchannels = nch;
            S.width = S.full_width = w;
            S.height = S.full_height = h;
            S.xstride  = (int64_t)nch * 4;
            S.ystride  = (int64_t)w * nch * 4;
            S.memspace = B2R_MEM_DEVICE;
            S.device   = device;
            b2r_image D = S;
            if (li == 1) {
                // level 0 keeps its pixel origin; its display window is set
                // to its data window (:876-877).  Levels >= 1 sit at the origin.
                S.x = S.full_x = level0->x;
                S.y = S.full_y = level0->y;
            }
            D.pixels    = lv.px;
            D.width = D.full_width = sw;
            D.height = D.full_height = sh;
            D.ystride = (int64_t)sw * nch * 4;
            if (box_path) {
                if (pass == 0)
                    continue;
                BoxParams bp;
                bp.dst = to_dev(D, D.pixels);
                bp.src = to_dev(S, S.pixels);
                // resize_block (:304-334): 2-pass needs 2*dw == sw and, without
                // allow_pixel_shift, even source sizes (:265-266)
                bp.two_pass = (2 * sw == w) && !(w % 2 || h % 2);
            bp.force_clamp = 0;
                launch_box_downsample(bp, stream);
                chain->launches += 1;
            } else {
                b2r_image R = S;  // what the resize reads
                if (hicomp && !(sharpen > 0.0f)) {
                    // rangecompress -> resize -> rangeexpand in ONE kernel when
                    // the level's shape belongs to the stream kernel (else the
                    // three passes, inside b2r_resize), then the clamp (:936-938)
                    b2r_options hc = sub;
                    hc.highlightcomp  = 2;  // + the clamp, wherever the call ends up
                    hc.alpha_channel1 = alpha_channel + 1;
                    hc.z_channel1     = hc_z + 1;
                    rc = b2r_resize(&D, &S, fname.c_str(), 0.0f, nullptr, &hc, nullptr, err,
                                    errlen);
                    if (rc) {
                        if (ev0)
                            cudaEventDestroy(ev0);
                        if (ev1)
                            cudaEventDestroy(ev1);
                        return fail(rc);
                    }
                    if (pass == 1)
                        chain->launches += 1;
                } else {
                if (hicomp) {
                    // rangecompress(*img) (:907-909); the level itself stays
                    // intact for the caller, the compressed copy is scratch
                    R.pixels = hc_tmp;
                    if (pass == 1) {
                        b2r_roi all { 0, w, 0, h, 0, nch };
                        range_on_device(false, R, R.pixels, S, S.pixels, all, false,
                                        alpha_channel, hc_z, stream);
                        chain->launches += 1;
                    }
                }
                if (sharpen > 0.0f && sharpen_first && pass == 1) {
                    // unsharp_mask(*sharp, *img, sharpenfilt, 3, sharpen, 0) and
                    // swap (:910-918): the sharpened copy feeds the resize, the
                    // level itself stays as it was written
                    b2r_image T = S;
                    T.pixels    = hc_tmp;
                    rc = b2r_unsharp_mask(&T, &R, sharpenfilt.c_str(), 3.0f, sharpen, 0.0f,
                                          nullptr, &sub, nullptr, err, errlen);
                    R = T;
                    chain->launches += 2;
                } else if (sharpen > 0.0f && sharpen_first) {
                    R.pixels = hc_tmp;
                }
                if (!rc)
                    rc = b2r_resize(&D, &R, fname.c_str(), 0.0f, nullptr, &sub, nullptr,
                                    err, errlen);
                if (!rc && sharpen > 0.0f && !sharpen_first && pass == 1) {
                    // sharpen the new level in place of itself (:923-931)
                    rc = b2r_unsharp_mask(&D, &D, sharpenfilt.c_str(), 3.0f, sharpen, 0.0f,
                                          nullptr, &sub, nullptr, err, errlen);
                    chain->launches += 2;
                }
                if (rc) {
                    if (ev0)
                        cudaEventDestroy(ev0);
                    if (ev1)
                        cudaEventDestroy(ev1);
                    return fail(rc);
                }
                if (pass == 1)
                    chain->launches += 2;
                if (hicomp && pass == 1) {
                    // rangeexpand(*small) in place (:933-935)
                    b2r_roi all { 0, sw, 0, sh, 0, nch };
                    range_on_device(true, D, D.pixels, D, D.pixels, all, false,
                                    alpha_channel, hc_z, stream);
                    // ... and clamp(*small, 0, FLT_MAX, clampalpha01) (:936-938)
                    launch_clamp((float*)lv.px, (long long)sw * sh * nch, nch,
                                 alpha_channel, 0.0f, 3.402823466e+38f, stream);
                    chain->launches += 2;
                }
                }
            }
            if (pass == 1 && clamp_half) {
                launch_clamp((float*)lv.px, (long long)sw * sh * nch, nch,
                             alpha_channel, -65504.0f, 65504.0f, stream);
                chain->launches += 1;
            }
        }
    }
    cudaEventRecord(ev1, stream);
    cudaError_t e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess)
        e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
        cudaEventDestroy(ev0);
        cudaEventDestroy(ev1);
        return fail(B2R_ERR_CUDA);
    }
    cudaEventElapsedTime(&chain->kernel_ms, ev0, ev1);
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    *out = chain.release();
    return B2R_OK;
}

// =========================================================================
// The chain for a HOST level 0 with its levels delivered to HOST memory, as one
// pipeline (write_mipmap end to end, maketexture.cpp:823-986, without the
// file): level 0 goes up in row bands on one stream; as soon as the rows a band
// of level 1 reads have landed, that band is computed on a second stream and
// comes down on a third while the next rows are still going up (PCIe is full
// duplex, and level 1 is three quarters of everything that comes back);
// levels >= 2 follow from the device-resident level 1.  A host thread of its
// own issues the downloads, so a pageable destination (whose copies block
// while they drain the pinned bounce slots) never holds up the uploads.
// =========================================================================
extern "C" int
b2r_mipchain_create_streamed(b2r_mipchain** out, const b2r_image* level0,
                             const b2r_mip_options* mo, const b2r_options* opt,
                             const b2r_image* out_levels, int nout, char* err, size_t errlen)
{
    if (!out || !mo || !level0) {
        set_err(err, errlen, "mipchain: null argument");
        return B2R_ERR_INVALID;
    }
    *out = nullptr;
    int rc = check_image(level0, true, err, errlen);
    if (rc)
        return rc;
    int dev0 = opt ? opt->device : 0;
    const bool plain = level0->basetype == B2R_FLOAT && memspace_of(level0, &dev0) == B2R_MEM_HOST
                       && !(mo->sharpen > 0.0f) && !(opt && opt->highlightcomp) && !level0->x
                       && !level0->y && !level0->full_x && !level0->full_y
                       && level0->full_width == level0->width
                       && level0->full_height == level0->height
                       && level0->xstride == (int64_t)level0->nchannels * 4
                       && level0->width >= 2 && level0->height >= 256;
    auto download_all = [&](b2r_mipchain* c) -> int {
        for (int k = 1; k < (int)c->levels.size() && k - 1 < nout; ++k) {
            if (!out_levels || !out_levels[k - 1].pixels)
                continue;
            int r = b2r_mipchain_download(c, k, &out_levels[k - 1], err, errlen);
            if (r)
                return r;
        }
        return B2R_OK;
    };
    if (!plain) {
        // everything the pipeline does not cover: the chain, then the downloads
        rc = b2r_mipchain_create_ex(out, level0, mo, opt, err, errlen);
        if (rc)
            return rc;
        rc = download_all(*out);
        if (rc) {
            b2r_mipchain_destroy(*out);
            *out = nullptr;
        }
        return rc;
    }
    std::string fname = (mo->filtername && mo->filtername[0]) ? mo->filtername : "box";
    if (fname != "box") {
        bool known = false;
        for (int i = 0, e = Filter2D::num_filters(); i < e; ++i)
            known |= (fname == Filter2D::get_filterdesc(i).name);
        if (!known) {
            set_err(err, errlen, "Could not make filter \"%s\"", fname.c_str());
            return B2R_ERR_FILTER;
        }
    }
    if (device_count_cached() <= 0) {
        set_err(err, errlen, "no CUDA device available: the B200 MIP path has no CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    const int device = opt ? opt->device : 0;
    if (device < 0 || device >= device_count_cached()) {
        set_err(err, errlen, "invalid CUDA device ordinal %d", device);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(device);
    device_info(device);
    const int nch = level0->nchannels;
    std::unique_ptr<b2r_mipchain> chain(new b2r_mipchain);
    chain->device = device;
    chain->nch    = nch;
    auto fail = [&](int code) {
        for (auto& l : chain->levels)
            if (l.owned && l.px)
                cudaFree(l.px);
        chain->levels.clear();
        return code;
    };
    for (int w = level0->width, h = level0->height;;) {
        b2r_mipchain::Level lv { w, h, nullptr, true };
        if (cudaMalloc(&lv.px, size_t(w) * h * nch * 4) != cudaSuccess) {
            cudaGetLastError();
            set_err(err, errlen, "mipchain: out of device memory");
            return fail(B2R_ERR_CUDA);
        }
        chain->levels.push_back(lv);
        if (w <= 1 && h <= 1)
            break;
        w = w > 1 ? w / 2 : w;
        h = h > 1 ? h / 2 : h;
    }
    const int nlev = (int)chain->levels.size();
    for (int k = 1; k < nlev && k - 1 < nout; ++k) {
        const b2r_image* o = out_levels ? &out_levels[k - 1] : nullptr;
        if (!o || !o->pixels)
            continue;
        const auto& l = chain->levels[k];
        int odev = 0;
        if (o->basetype != B2R_FLOAT || o->nchannels != nch || o->width != l.w || o->height != l.h
            || o->xstride != (int64_t)nch * 4 || memspace_of(o, &odev) != B2R_MEM_HOST) {
            set_err(err, errlen, "mipchain: output image does not match level %d", k);
            return fail(B2R_ERR_INVALID);
        }
    }
    auto image_of = [&](int k) {
        const auto& l = chain->levels[k];
        b2r_image im;
        memset(&im, 0, sizeof(im));
        im.pixels    = l.px;
        im.basetype  = B2R_FLOAT;
        im.nchannels = nch;
        im.width = im.full_width = l.w;
        im.height = im.full_height = l.h;
        im.xstride  = (int64_t)nch * 4;
        im.ystride  = (int64_t)l.w * nch * 4;
        im.memspace = B2R_MEM_DEVICE;
        im.device   = device;
        return im;
    };

    std::lock_guard<std::mutex> lock(pipe_mutex(device));
    PipeStreams& ps = pipe_streams(device);
    const auto& L0  = chain->levels[0];
    const auto& L1  = chain->levels[1];
    const int nb    = std::max(2, std::min(16, L1.h / 128));
    std::vector<cudaEvent_t> ev_up(nb, nullptr), ev_k(nlev + nb, nullptr);
    auto free_events = [&]() {
        for (auto e : ev_up)
            if (e)
                cudaEventDestroy(e);
        for (auto e : ev_k)
            if (e)
                cudaEventDestroy(e);
    };
    for (auto& e : ev_up)
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto& e : ev_k)
        cudaEventCreateWithFlags(&e, cudaEventDisableTiming);

    // the download thread: task i = rows [y0, y1) of level k, ready after ev
    struct Task {
        int level, y0, y1;
        cudaEvent_t ready;
    };
    std::vector<Task> tasks;
    std::mutex tm;
    std::condition_variable tcv;
    bool closed    = false;
    int dl_rc      = B2R_OK;
    std::string dl_err;
    std::thread downloader([&]() {
        cudaSetDevice(device);
        size_t next = 0;
        for (;;) {
            Task t;
            {
                std::unique_lock<std::mutex> lk(tm);
                tcv.wait(lk, [&] { return next < tasks.size() || closed; });
                if (next >= tasks.size())
                    return;
                t = tasks[next++];
            }
            const b2r_image* o = (out_levels && t.level - 1 < nout) ? &out_levels[t.level - 1]
                                                                    : nullptr;
            if (!o || !o->pixels || dl_rc)
                continue;
            const auto& l   = chain->levels[t.level];
            const size_t rb = size_t(l.w) * nch * 4;
            cudaStreamWaitEvent(ps.out, t.ready, 0);
            cudaError_t e = copy_d2h_2d((unsigned char*)o->pixels + (long long)t.y0 * o->ystride,
                                        o->ystride, (const unsigned char*)l.px + size_t(t.y0) * rb,
                                        rb, rb, size_t(t.y1 - t.y0), ps.out);
            if (e != cudaSuccess) {
                dl_rc  = B2R_ERR_CUDA;
                dl_err = std::string("CUDA error: ") + cudaGetErrorString(e);
            }
        }
    });
    auto post = [&](int level, int y0, int y1, cudaEvent_t ready) {
        {
            std::lock_guard<std::mutex> lk(tm);
            tasks.push_back({ level, y0, y1, ready });
        }
        tcv.notify_one();
    };
    auto finish = [&](int code) {
        {
            std::lock_guard<std::mutex> lk(tm);
            closed = true;
        }
        tcv.notify_one();
        downloader.join();
        cudaStreamSynchronize(ps.in);
        cudaStreamSynchronize(ps.k);
        cudaStreamSynchronize(ps.out);
        free_events();
        return code;
    };

    b2r_options sub;
    memset(&sub, 0, sizeof(sub));
    sub.device      = device;
    sub.path        = opt ? opt->path : B2R_PATH_AUTO;
    sub.cuda_stream = (void*)ps.k;
    const size_t row0 = size_t(L0.w) * nch * 4;
    b2r_image S0 = image_of(0), D1 = image_of(1);
    int uploaded = 0;  // rows of level 0 on their way so far
    for (int b = 0; b < nb; ++b) {
        const int y0 = int((long long)L1.h * b / nb), y1 = int((long long)L1.h * (b + 1) / nb);
        int r0 = 0, r1 = 0;
        if (fname == "box") {
            const bool two_pass = (2 * L1.w == L0.w) && !(L0.w % 2 || L0.h % 2);
            r1 = two_pass ? 2 * y1
                          : std::min(L0.h, int(((long long)y1 * L0.h + L1.h - 1) / L1.h) + 2);
        } else {
            b2r_roi rr { 0, L1.w, y0, y1, 0, nch };
            rc = b2r_source_rows(&D1, &S0, fname.c_str(), 0.0f, &rr, &r0, &r1, err, errlen);
            if (rc)
                return fail(finish(rc));
        }
        if (b == nb - 1)
            r1 = L0.h;
        r1 = std::min(L0.h, std::max(r1, uploaded));
        if (r1 > uploaded) {
            cudaError_t e = copy_h2d_2d(
                (unsigned char*)L0.px + size_t(uploaded) * row0, row0,
                (const unsigned char*)level0->pixels + (long long)uploaded * level0->ystride,
                level0->ystride, row0, size_t(r1 - uploaded), ps.in);
            if (e != cudaSuccess) {
                set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
                return fail(finish(B2R_ERR_CUDA));
            }
            uploaded = r1;
        }
        cudaEventRecord(ev_up[b], ps.in);
        cudaStreamWaitEvent(ps.k, ev_up[b], 0);
        b2r_roi roi { 0, L1.w, y0, y1, 0, nch };
        rc = b2r_mip_level(&D1, &S0, fname.c_str(), &roi, &sub, err, errlen);
        if (rc)
            return fail(finish(rc));
        if (mo->clamp_half)
            launch_clamp((float*)L1.px + size_t(y0) * L1.w * nch, (long long)(y1 - y0) * L1.w * nch,
                         nch, mo->alpha_channel, -65504.0f, 65504.0f, ps.k);
        chain->launches += 1;
        cudaEventRecord(ev_k[b], ps.k);
        post(1, y0, y1, ev_k[b]);
    }
    for (int k = 2; k < nlev; ++k) {
        b2r_image S = image_of(k - 1), D = image_of(k);
        rc = b2r_mip_level(&D, &S, fname.c_str(), nullptr, &sub, err, errlen);
        if (rc)
            return fail(finish(rc));
        if (mo->clamp_half)
            launch_clamp((float*)D.pixels, (long long)D.width * D.height * nch, nch,
                         mo->alpha_channel, -65504.0f, 65504.0f, ps.k);
        chain->launches += 1;
        cudaEventRecord(ev_k[nb + k], ps.k);
        post(k, 0, D.height, ev_k[nb + k]);
    }
    rc = finish(B2R_OK);
    cudaError_t ce = cudaGetLastError();
    if (dl_rc) {
        set_err(err, errlen, "%s", dl_err.c_str());
        return fail(dl_rc);
    }
    if (ce != cudaSuccess) {
        set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(ce));
        return fail(B2R_ERR_CUDA);
    }
    *out = chain.release();
    return rc;
}

extern "C" int
b2r_mipchain_nlevels(const b2r_mipchain* chain)
{
    return chain ? (int)chain->levels.size() : 0;
}

extern "C" int
b2r_mipchain_level(const b2r_mipchain* chain, int level, int* width, int* height,
                   void** device_pixels)
{
    if (!chain || level < 0 || level >= (int)chain->levels.size())
        return B2R_ERR_INVALID;
    if (width)
        *width = chain->levels[level].w;
    if (height)
        *height = chain->levels[level].h;
    if (device_pixels)
        *device_pixels = chain->levels[level].px;
    return B2R_OK;
}

extern "C" int
b2r_mipchain_download(const b2r_mipchain* chain, int level, const b2r_image* out,
                      char* err, size_t errlen)
{
    if (!chain || level < 0 || level >= (int)chain->levels.size() || !out
        || !out->pixels) {
        set_err(err, errlen, "mipchain: bad level or output image");
        return B2R_ERR_INVALID;
    }
    const auto& l = chain->levels[level];
    if (out->basetype != B2R_FLOAT || out->nchannels != chain->nch
        || out->width != l.w || out->height != l.h
        || out->xstride != (int64_t)chain->nch * 4) {
        set_err(err, errlen, "mipchain: output image does not match level %d",
                level);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(chain->device);
    size_t row = size_t(l.w) * chain->nch * 4;
    int odev = 0;
    if (memspace_of(out, &odev) == B2R_MEM_HOST) {
        CUDA_TRY(copy_d2h_2d(out->pixels, out->ystride, l.px, row, row, l.h, nullptr));
        CUDA_TRY(cudaStreamSynchronize(nullptr));
    } else {
        CUDA_TRY(cudaMemcpy2D(out->pixels, out->ystride, l.px, row, row, l.h,
                              cudaMemcpyDefault));
    }
    return B2R_OK;
}

extern "C" float
b2r_mipchain_kernel_ms(const b2r_mipchain* chain)
{
    return chain ? chain->kernel_ms : 0.0f;
}

extern "C" void
b2r_mipchain_destroy(b2r_mipchain* chain)
{
    if (!chain)
        return;
    DeviceGuard guard(chain->device);
    for (auto& l : chain->levels)
        if (l.owned && l.px)
            cudaFree(l.px);
    delete chain;
}

// =========================================================================
// one MIP level step on arbitrary windows (row-band sharding of the chain)
// =========================================================================
extern "C" int
b2r_mip_level(const b2r_image* dst, const b2r_image* src, const char* filtername,
              const b2r_roi* roi_in, const b2r_options* opt, char* err, size_t errlen)
{
    int rc = check_image(src, true, err, errlen);
    if (rc)
        return rc;
    rc = check_image(dst, false, err, errlen);
    if (rc)
        return rc;
    std::string fname = (filtername && filtername[0]) ? filtername : "box";
    if (fname != "box")
        return b2r_resize(dst, src, fname.c_str(), 0.0f, roi_in, opt, nullptr, err,
                          errlen);
    if (dst->basetype != B2R_FLOAT || src->basetype != B2R_FLOAT
        || dst->nchannels != src->nchannels) {
        set_err(err, errlen, "mip_level: box levels are float32 with equal channels");
        return B2R_ERR_TYPES;
    }
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 MIP path has no CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    int sdev = 0, ddev = 0;
    if (memspace_of(src, &sdev) != B2R_MEM_DEVICE
        || memspace_of(dst, &ddev) != B2R_MEM_DEVICE || sdev != ddev) {
        set_err(err, errlen, "mip_level: box levels must be device resident");
        return B2R_ERR_UNSUPPORTED;
    }
    b2r_roi roi = roi_in ? *roi_in
                         : b2r_roi { dst->x, dst->x + dst->width, dst->y,
                                     dst->y + dst->height, 0, dst->nchannels };
    roi.xbegin = std::max(roi.xbegin, dst->x);
    roi.xend   = std::min(roi.xend, dst->x + dst->width);
    roi.ybegin = std::max(roi.ybegin, dst->y);
    roi.yend   = std::min(roi.yend, dst->y + dst->height);
    if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin)
        return B2R_OK;
    DeviceGuard guard(sdev);
    cudaStream_t stream = opt ? (cudaStream_t)opt->cuda_stream : nullptr;
    const int nch = dst->nchannels;
    // resize_block (maketexture.cpp:304-334): the 2x2 average applies when the
    // WHOLE level halves an even-sized, zero-origin image; a band of it is the
    // same arithmetic on rows 2y, 2y+1
    const bool two_pass
        = 2 * dst->full_width == src->full_width && !(src->full_width % 2)
          && !(src->full_height % 2) && dst->full_x == 0 && dst->full_y == 0
          && src->full_x == 0 && src->full_y == 0 && roi.xbegin == 0
          && roi.xend == dst->full_width && dst->x == 0 && src->x == 0
          && src->width == src->full_width && dst->width == dst->full_width;
    BoxParams bp;
    if (two_pass) {
        if (2 * roi.ybegin < src->y || 2 * roi.yend > src->y + src->height) {
            set_err(err, errlen, "mip_level: source rows [%d,%d) not present",
                    2 * roi.ybegin, 2 * roi.yend);
            return B2R_ERR_INVALID;
        }
        b2r_image D = *dst, S = *src;
        D.pixels = (unsigned char*)dst->pixels
                   + (long long)(roi.ybegin - dst->y) * dst->ystride;
        D.y      = 0;
        D.height = roi.yend - roi.ybegin;
        S.pixels = (unsigned char*)src->pixels
                   + (long long)(2 * roi.ybegin - src->y) * src->ystride;
        S.y      = 0;
        S.height = 2 * D.height;
        bp.dst      = to_dev(D, D.pixels);
        bp.src      = to_dev(S, S.pixels);
        bp.two_pass = 1;
        bp.force_clamp = 0;
    } else {
        // bilinear at NDC centres over the ROI; the chain's images are never
        // crops (windows are zeroed per level), so WrapClamp even though a
        // band's data window is smaller than the level
        b2r_image D = *dst;
        D.pixels = (unsigned char*)dst->pixels
                   + (long long)(roi.ybegin - dst->y) * dst->ystride
                   + (long long)(roi.xbegin - dst->x) * dst->xstride;
        D.x      = roi.xbegin;
        D.y      = roi.ybegin;
        D.width  = roi.xend - roi.xbegin;
        D.height = roi.yend - roi.ybegin;
        bp.dst      = to_dev(D, D.pixels);
        bp.src      = to_dev(*src, src->pixels);
        bp.two_pass = 0;
        bp.force_clamp = 1;
    }
    (void)nch;
    launch_box_downsample(bp, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_err(err, errlen, "CUDA launch error: %s", cudaGetErrorString(e));
        return B2R_ERR_CUDA;
    }
    return B2R_OK;
}
