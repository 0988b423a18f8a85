// b2r_api_warp.cpp -- part of the C ABI (include/b2r.h): warp / rotate / fit(exact) / --resize:from/to and st_warp.
#include "api_internal.h"
#include "host_copy.h"

#include "conv_kernel.h"
#include "xform_matrix.h"

#include <cmath>
#include <string>
#include <vector>

using namespace b2r;
using namespace b2r::api;

// =========================================================================
// warp / rotate / fit(exact) / --resize:from/to
// (imagebufalgo_xform.cpp:263-591, 1012-1027, 1727-1760; oiiotool.cpp:4684-4720)
// =========================================================================

namespace {

int
warp_impl(const b2r_image* dst_in, const b2r_image* src_in, const float* M9,
          const Filter2D& filter, int wrap, int edgeclamp, const b2r_roi* roi_in,
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
    if (!M9) {
        set_err(err, errlen, "warp: no matrix");
        return B2R_ERR_INVALID;
    }
    if (wrap < 0 || wrap > 4) {
        set_err(err, errlen, "warp: unknown wrap mode %d", wrap);
        return B2R_ERR_INVALID;
    }
    const b2r_image& dst = *dst_in;
    const b2r_image& src = *src_in;
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 warp path has no CPU fallback");
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
    roi.chend = std::min(roi.chend, src.nchannels);  // warp_impl, xform.cpp:395
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

    // Host images: the whole source data window goes up (a warp may read any
    // of it), the dst ROI comes back.
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
        const size_t rowbytes = size_t(src.width) * src.xstride;
        const size_t pitch    = (rowbytes + 15) & ~size_t(15);
        CUDA_TRY(cudaMallocAsync(&src_alloc, pitch * src.height, stream));
        CUDA_TRY(copy_h2d_2d(src_alloc, pitch, src.pixels, src.ystride, rowbytes,
                                   src.height, stream));
        srcd.ystride = (int64_t)pitch;
        dsrc         = src_alloc;
        if (report)
            report->h2d_bytes += (int64_t)rowbytes * src.height;
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
    WarpParams wp;
    memset(&wp, 0, sizeof(wp));
    wp.dst     = to_dev(dstd, ddst);
    wp.src     = to_dev(srcd, dsrc);
    wp.roi_x0  = roi.xbegin;
    wp.roi_x1  = roi.xend;
    wp.roi_y0  = roi.ybegin;
    wp.roi_y1  = roi.yend;
    wp.chbegin = roi.chbegin;
    wp.chend   = roi.chend;
    Mat33(M9).inverse().store(wp.minv);  // warp_: Minv = M.inverse()
    const Filter2D::DeviceDesc fd = filter.device_desc();
    wp.filt_profile = fd.profile;
    wp.filt_scale   = fd.scale;
    wp.filt_param   = fd.param;
    wp.filt_w       = fd.w;
    wp.filt_h       = fd.h;
    wp.wrap         = wrap;
    wp.edgeclamp    = edgeclamp ? 1 : 0;
    wp.contract_fma = (opt && opt->contract_fma) ? 1 : 0;
    launch_warp(wp, stream);
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
        report->path_used        = B2R_PATH_EXACT;
    }
    return B2R_OK;
}

}  // namespace

extern "C" int
b2r_warp_filter(const b2r_image* dst, const b2r_image* src, const float* M,
                const b2r_filter2d* filter, int wrap, int edgeclamp, const b2r_roi* roi,
                const b2r_options* opt, b2r_report* report, char* err, size_t errlen)
{
    if (!filter || !filter->f) {
        // warp_impl: "If no filter was provided, punt and use lanczos3" (:403-408)
        std::unique_ptr<Filter2D, void (*)(Filter2D*)> f(Filter2D::create("lanczos3", 6.0f,
                                                                          6.0f),
                                                         Filter2D::destroy);
        return warp_impl(dst, src, M, *f, wrap, edgeclamp, roi, opt, report, err, errlen);
    }
    return warp_impl(dst, src, M, *filter->f, wrap, edgeclamp, roi, opt, report, err, errlen);
}

extern "C" int
b2r_warp(const b2r_image* dst, const b2r_image* src, const float* M, const char* filtername,
         float filterwidth, int wrap, int edgeclamp, const b2r_roi* roi,
         const b2r_options* opt, b2r_report* report, char* err, size_t errlen)
{
    // get_warp_filter, imagebufalgo_xform.cpp:329-349
    std::string name = (filtername && filtername[0]) ? filtername : "lanczos3";
    std::unique_ptr<Filter2D, void (*)(Filter2D*)> filter(nullptr, Filter2D::destroy);
    for (int i = 0, e = Filter2D::num_filters(); i < e; ++i) {
        const FilterDesc& fd = Filter2D::get_filterdesc(i);
        if (name == fd.name) {
            const float w = filterwidth > 0.0f ? filterwidth : fd.width;
            filter.reset(Filter2D::create(name, w, w));
            break;
        }
    }
    if (!filter) {
        set_err(err, errlen, "Filter \"%s\" not recognized", name.c_str());
        return B2R_ERR_FILTER;
    }
    return warp_impl(dst, src, M, *filter, wrap, edgeclamp, roi, opt, report, err, errlen);
}

extern "C" void
b2r_m33_inverse(const float* M, float* out)
{
    Mat33(M).inverse().store(out);
}

extern "C" void
b2r_rotate_matrix(float angle, float center_x, float center_y, float* out)
{
    rotation_about(angle, center_x, center_y).store(out);
}

extern "C" void
b2r_fromto_matrix(float from_x, float from_y, float from_w, float from_h, float to_x,
                  float to_y, float to_w, float to_h, float* out)
{
    fromto_matrix(from_x, from_y, from_w, from_h, to_x, to_y, to_w, to_h).store(out);
}

extern "C" void
b2r_transform_roi(const float* M, const b2r_roi* roi, b2r_roi* out)
{
    int b[4];
    transform_bounds(Mat33(M), roi->xbegin, roi->xend, roi->ybegin, roi->yend, b);
    *out        = *roi;
    out->xbegin = b[0], out->xend = b[1], out->ybegin = b[2], out->yend = b[3];
}

extern "C" void
b2r_fit_exact(int src_full_width, int src_full_height, int fit_width, int fit_height,
              const char* fillmode, float* M, int* resize_width, int* resize_height)
{
    // imagebufalgo_xform.cpp:962-997 (scale / offsets) and :1023
    const float oldaspect = float(src_full_width) / src_full_height;
    const float newaspect = float(fit_width) / fit_height;
    int rw = fit_width, rh = fit_height;
    float xoff = 0.0f, yoff = 0.0f, scale = 1.0f;
    std::string mode = fillmode ? fillmode : "letterbox";
    if (mode != "height" && mode != "width")
        mode = "letterbox";
    if (mode == "letterbox")
        mode = (newaspect >= oldaspect) ? "height" : "width";
    if (mode == "height") {
        rw    = int(rh * oldaspect + 0.5f);
        scale = float(fit_height) / float(src_full_height);
        xoff  = float(fit_width - scale * src_full_width) / 2.0f;
    } else {
        rh    = int(rw / oldaspect + 0.5f);
        scale = float(fit_width) / float(src_full_width);
        yoff  = float(fit_height - scale * src_full_height) / 2.0f;
    }
    if (M) {
        const float m[9] = { scale, 0.0f, 0.0f, 0.0f, scale, 0.0f, xoff, yoff, 1.0f };
        memcpy(M, m, sizeof(m));
    }
    if (resize_width)
        *resize_width = rw;
    if (resize_height)
        *resize_height = rh;
}


// =========================================================================
// st_warp (imagebufalgo_xform.cpp:1834-2058)
// =========================================================================
namespace {
int
st_warp_impl(const b2r_image* dst, const b2r_image* src, const b2r_image* st,
             const Filter2D& filter, int chan_s, int chan_t, int flip_s, int flip_t,
             const b2r_roi* roi_in, const b2r_options* opt, b2r_report* report, char* err,
             size_t errlen)
{
    // check_st_warp_args (:1929-1972)
    if (!st || !st->pixels || st->width <= 0 || st->height <= 0 || st->nchannels <= 0) {
        set_err(err, errlen, "ImageBufAlgo::st_warp : Uninitialized ST buffer");
        return B2R_ERR_INVALID;
    }
    if (!valid_basetype(st->basetype)) {
        set_err(err, errlen, "unsupported pixel data type %d", st->basetype);
        return B2R_ERR_TYPES;
    }
    if (chan_s >= st->nchannels || chan_s < 0) {
        set_err(err, errlen, "ImageBufAlgo::st_warp : Out-of-range S channel index: %d",
                chan_s);
        return B2R_ERR_INVALID;
    }
    if (chan_t >= st->nchannels || chan_t < 0) {
        set_err(err, errlen, "ImageBufAlgo::st_warp : Out-of-range T channel index: %d",
                chan_t);
        return B2R_ERR_INVALID;
    }
    StencilIO io;
    int rc = io.begin(dst, src, roi_in, opt, report, "st_warp", err, errlen, false);
    if (rc || io.empty) {
        io.release();
        return rc;
    }
    io.roi.chend = std::min(io.roi.chend, src->nchannels);
    // the warp is only defined where the ST image has pixels
    b2r_roi r = io.roi;
    r.xbegin  = std::max(r.xbegin, st->x), r.xend = std::min(r.xend, st->x + st->width);
    r.ybegin  = std::max(r.ybegin, st->y), r.yend = std::min(r.yend, st->y + st->height);
    if (r.xend <= r.xbegin || r.yend <= r.ybegin) {
        io.release();
        set_err(err, errlen,
                "ImageBufAlgo::st_warp : Output ROI does not intersect ST buffer.");
        return B2R_ERR_INVALID;
    }
    if (io.dst_alloc && (r.xbegin != io.roi.xbegin || r.xend != io.roi.xend
                         || r.ybegin != io.roi.ybegin || r.yend != io.roi.yend
                         || io.roi.chbegin != 0 || io.roi.chend != dst->nchannels)) {
        // part of the staged dst ROI will not be written: bring its pixels over
        copy_h2d_2d(io.dst_alloc, io.dpitch, io.host_out, dst->ystride, io.drowbytes,
                          io.roi_h, io.stream);
    }
    b2r_image std_;
    void* dst_pix = nullptr;
    rc = io.stage_extra(*st, &std_, &dst_pix, report, err, errlen);
    if (rc) {
        io.release();
        return rc;
    }
    StWarpParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.dst     = to_dev(io.dstd, io.ddst);
    sp.src     = to_dev(io.srcd, io.dsrc);
    sp.st      = to_dev(std_, dst_pix);
    sp.roi_x0  = r.xbegin;
    sp.roi_x1  = r.xend;
    sp.roi_y0  = r.ybegin;
    sp.roi_y1  = r.yend;
    sp.chbegin = io.roi.chbegin;
    sp.chend   = io.roi.chend;
    sp.chan_s  = chan_s;
    sp.chan_t  = chan_t;
    sp.flip_s  = flip_s ? 1 : 0;
    sp.flip_t  = flip_t ? 1 : 0;
    const float xscale = float(dst->full_width) / src->full_width;
    const float yscale = float(dst->full_height) / src->full_height;
    sp.filterrad_x     = (int)ceilf(filter.width() / 2.0f / xscale);
    sp.filterrad_y     = (int)ceilf(filter.height() / 2.0f / yscale);
    const Filter2D::DeviceDesc fd = filter.device_desc();
    sp.filt_profile = fd.profile;
    sp.filt_scale   = fd.scale;
    sp.filt_param   = fd.param;
    sp.filt_w       = fd.w;
    sp.filt_h       = fd.h;
    launch_st_warp(sp, io.stream);
    return io.end(*dst, report, 1, err, errlen);
}
}  // namespace

extern "C" int
b2r_st_warp_filter(const b2r_image* dst, const b2r_image* src, const b2r_image* st,
                   const b2r_filter2d* filter, int chan_s, int chan_t, int flip_s,
                   int flip_t, const b2r_roi* roi, const b2r_options* opt,
                   b2r_report* report, char* err, size_t errlen)
{
    if (!filter || !filter->f) {  // :1994-1998
        std::unique_ptr<Filter2D, void (*)(Filter2D*)> f(Filter2D::create("lanczos3", 6.0f,
                                                                          6.0f),
                                                         Filter2D::destroy);
        return st_warp_impl(dst, src, st, *f, chan_s, chan_t, flip_s, flip_t, roi, opt, report,
                            err, errlen);
    }
    return st_warp_impl(dst, src, st, *filter->f, chan_s, chan_t, flip_s, flip_t, roi, opt,
                        report, err, errlen);
}

extern "C" int
b2r_st_warp(const b2r_image* dst, const b2r_image* src, const b2r_image* st,
            const char* filtername, float filterwidth, int chan_s, int chan_t, int flip_s,
            int flip_t, const b2r_roi* roi, const b2r_options* opt, b2r_report* report,
            char* err, size_t errlen)
{
    std::string name = (filtername && filtername[0]) ? filtername : "lanczos3";
    std::unique_ptr<Filter2D, void (*)(Filter2D*)> filter(nullptr, Filter2D::destroy);
    for (int i = 0, e = Filter2D::num_filters(); i < e; ++i) {
        const FilterDesc& fd = Filter2D::get_filterdesc(i);
        if (name == fd.name) {
            const float w = filterwidth > 0.0f ? filterwidth : fd.width;
            filter.reset(Filter2D::create(name, w, w));
            break;
        }
    }
    if (!filter) {
        set_err(err, errlen, "Filter \"%s\" not recognized", name.c_str());
        return B2R_ERR_FILTER;
    }
    return st_warp_impl(dst, src, st, *filter, chan_s, chan_t, flip_s, flip_t, roi, opt,
                        report, err, errlen);
}
