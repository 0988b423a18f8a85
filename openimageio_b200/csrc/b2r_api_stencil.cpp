// b2r_api_stencil.cpp -- part of the C ABI (include/b2r.h): make_kernel / convolve / unsharp_mask, fixNonFinite, fillholes_pushpull,
// computePixelStats.
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
// make_kernel / convolve / unsharp_mask  (imagebufalgo.cpp:772-1031)
// =========================================================================

namespace {

// upload a host float table; freed stream-ordered by the caller
int
upload_floats(const float* host, size_t n, cudaStream_t stream, float** out, char* err,
              size_t errlen)
{
    CUDA_TRY(cudaMallocAsync((void**)out, n * sizeof(float), stream));
    CUDA_TRY(cudaMemcpyAsync(*out, host, n * sizeof(float), cudaMemcpyHostToDevice, stream));
    return B2R_OK;
}

float
kernel_scale(const float* k, int n, bool normalize)
{
    if (!normalize)
        return 1.0f;
    float s = 0.0f;
    for (int i = 0; i < n; ++i)
        s += k[i];
    return 1.0f / s;
}

void
run_convolve(const b2r_image& dstd, void* ddst, const b2r_image& srcd, void* dsrc,
             const b2r_roi& roi, const float* dkernel, int kw, int kh, float scale,
             cudaStream_t stream)
{
    ConvParams cp;
    memset(&cp, 0, sizeof(cp));
    cp.dst     = to_dev(dstd, ddst);
    cp.src     = to_dev(srcd, dsrc);
    cp.roi_x0  = std::max(roi.xbegin, dstd.x);
    cp.roi_x1  = std::min(roi.xend, dstd.x + dstd.width);
    cp.roi_y0  = std::max(roi.ybegin, dstd.y);
    cp.roi_y1  = std::min(roi.yend, dstd.y + dstd.height);
    cp.chbegin = roi.chbegin;
    cp.chend   = roi.chend;
    cp.kernel  = dkernel;
    cp.kw      = kw;
    cp.kh      = kh;
    cp.scale   = scale;
    launch_convolve(cp, stream);
}

}  // namespace

/* ImageBufAlgo::make_kernel (imagebufalgo.cpp:864-957), depth 1. */
extern "C" int
b2r_make_kernel(const char* name, float width, float height, int normalize, float* out,
                int* kw, int* kh, char* err, size_t errlen)
{
    ConvKernel K = make_conv_kernel(name ? name : "", width, height, normalize != 0);
    if (kw)
        *kw = K.w;
    if (kh)
        *kh = K.h;
    if (out)
        memcpy(out, K.v.data(), K.v.size() * sizeof(float));
    if (!K.known) {
        set_err(err, errlen, "%s", K.error.c_str());
        return B2R_ERR_FILTER;
    }
    return B2R_OK;
}

extern "C" int
b2r_convolve(const b2r_image* dst, const b2r_image* src, const float* kernel, int kw, int kh,
             int normalize, const b2r_roi* roi, const b2r_options* opt, b2r_report* report,
             char* err, size_t errlen)
{
    if (!kernel || kw <= 0 || kh <= 0) {
        set_err(err, errlen, "convolve: no kernel");
        return B2R_ERR_INVALID;
    }
    StencilIO io;
    int rc = io.begin(dst, src, roi, opt, report, "convolve", err, errlen);
    if (rc || io.empty) {
        io.release();
        return rc;
    }
    if (io.same_pixels) {
        io.release();
        set_err(err, errlen, "convolve: dst and src must be different images");
        return B2R_ERR_INVALID;
    }
    float* dk = nullptr;
    rc        = upload_floats(kernel, (size_t)kw * kh, io.stream, &dk, err, errlen);
    if (rc) {
        io.release();
        return rc;
    }
    run_convolve(io.dstd, io.ddst, io.srcd, io.dsrc, io.roi, dk, kw, kh,
                 kernel_scale(kernel, kw * kh, normalize != 0), io.stream);
    cudaFreeAsync(dk, io.stream);
    return io.end(*dst, report, 1, err, errlen);
}

extern "C" int
b2r_unsharp_mask(const b2r_image* dst, const b2r_image* src, const char* kernel, float width,
                 float contrast, float threshold, const b2r_roi* roi, const b2r_options* opt,
                 b2r_report* report, char* err, size_t errlen)
{
    const std::string kname = (kernel && kernel[0]) ? kernel : "gaussian";
    if (kname == "median") {
        set_err(err, errlen,
                "unsharp_mask: the median kernel is not part of the B200 resampling path");
        return B2R_ERR_UNSUPPORTED;
    }
    // the kernels first: an unknown name is an error before anything is staged
    const bool two_pass = width > 3.0;
    ConvKernel K = two_pass ? make_conv_kernel(kname, 1, width) : make_conv_kernel(kname, width, width);
    if (!K.known) {
        set_err(err, errlen, "%s", K.error.c_str());
        return B2R_ERR_FILTER;
    }
    StencilIO io;
    int rc = io.begin(dst, src, roi, opt, report, "unsharp_mask", err, errlen);
    if (rc || io.empty) {
        io.release();
        return rc;
    }
    // Blurry (and fst_pass): float images with src's windows, zero filled (:988-992)
    const b2r_image& S = io.srcd;
    const size_t brow  = size_t(S.width) * S.nchannels * sizeof(float);
    const size_t bsize = brow * S.height;
    void *blur = nullptr, *first = nullptr;
    float* dk  = nullptr;
    auto fail  = [&](int code) {
        if (blur)
            cudaFreeAsync(blur, io.stream);
        if (first)
            cudaFreeAsync(first, io.stream);
        if (dk)
            cudaFreeAsync(dk, io.stream);
        io.release();
        return code;
    };
    rc = upload_floats(K.v.data(), K.v.size(), io.stream, &dk, err, errlen);
    if (rc)
        return fail(rc);
    b2r_roi croi  = io.roi;  // convolve runs over every channel (:1004-1019 pass roi as is)
    const float s = kernel_scale(K.v.data(), K.w * K.h, true);
    int launched  = 0;
    if (!two_pass && !io.same_pixels && !io.dst_alloc_is_src()) {
        // one 2-D pass and distinct images: blur and combine in ONE kernel,
        // the blur image never exists (same arithmetic, same result)
        ConvParams cp;
        memset(&cp, 0, sizeof(cp));
        cp.dst       = to_dev(io.dstd, io.ddst);
        cp.src       = to_dev(S, io.dsrc);
        cp.roi_x0    = io.roi.xbegin, cp.roi_x1 = io.roi.xend;
        cp.roi_y0    = io.roi.ybegin, cp.roi_y1 = io.roi.yend;
        cp.chbegin   = io.roi.chbegin;
        cp.chend     = io.roi.chend;
        cp.kernel    = dk;
        cp.kw        = K.w;
        cp.kh        = K.h;
        cp.scale     = s;
        cp.unsharp   = 1;
        cp.contrast  = contrast;
        cp.threshold = threshold;
        launch_convolve(cp, io.stream);
        cudaFreeAsync(dk, io.stream);
        return io.end(*dst, report, 1, err, errlen);
    }
    if (cudaMallocAsync(&blur, bsize, io.stream) != cudaSuccess
        || cudaMemsetAsync(blur, 0, bsize, io.stream) != cudaSuccess) {
        set_err(err, errlen, "CUDA error allocating the blur image");
        return fail(B2R_ERR_CUDA);
    }
    b2r_image B = S;
    B.pixels    = blur;
    B.basetype  = B2R_FLOAT;
    B.xstride   = (int64_t)S.nchannels * 4;
    B.ystride   = (int64_t)brow;
    B.memspace  = B2R_MEM_DEVICE;
    if (two_pass) {
        if (cudaMallocAsync(&first, bsize, io.stream) != cudaSuccess
            || cudaMemsetAsync(first, 0, bsize, io.stream) != cudaSuccess) {
            set_err(err, errlen, "CUDA error allocating the first-pass image");
            return fail(B2R_ERR_CUDA);
        }
        b2r_image F = B;
        F.pixels    = first;
        run_convolve(F, first, S, io.dsrc, croi, dk, K.w, K.h, s, io.stream);
        // Kt = transpose(K): same values, K.h wide and K.w tall
        run_convolve(B, blur, F, first, croi, dk, K.h, K.w, s, io.stream);
        launched += 2;
    } else {
        run_convolve(B, blur, S, io.dsrc, croi, dk, K.w, K.h, s, io.stream);
        launched += 1;
    }
    UnsharpParams up;
    memset(&up, 0, sizeof(up));
    up.dst       = to_dev(io.dstd, io.ddst);
    up.src       = to_dev(S, io.dsrc);
    up.blur      = to_dev(B, blur);
    up.roi_x0    = io.roi.xbegin;
    up.roi_x1    = io.roi.xend;
    up.roi_y0    = io.roi.ybegin;
    up.roi_y1    = io.roi.yend;
    up.chbegin   = io.roi.chbegin;
    up.chend     = io.roi.chend;
    up.contrast  = contrast;
    up.threshold = threshold;
    launch_unsharp_combine(up, io.stream);
    ++launched;
    cudaFreeAsync(blur, io.stream);
    if (first)
        cudaFreeAsync(first, io.stream);
    cudaFreeAsync(dk, io.stream);
    return io.end(*dst, report, launched, err, errlen);
}



// =========================================================================
// fixNonFinite / check_nan  (imagebufalgo_pixelmath.cpp:1424-1610,
// maketexture.cpp:340-360, 1666-1705)
// =========================================================================
extern "C" int
b2r_fix_nonfinite(const b2r_image* dst, const b2r_image* src, int mode, int* pixels_fixed,
                  const b2r_roi* roi, const b2r_options* opt, b2r_report* report, char* err,
                  size_t errlen)
{
    if (pixels_fixed)
        *pixels_fixed = 0;
    if (mode != 0 && mode != 1 && mode != 2 && mode != 100) {
        set_err(err, errlen, "fixNonFinite: unknown repair mode");
        return B2R_ERR_INVALID;
    }
    StencilIO io;
    int rc = io.begin(dst, src, roi, opt, report, "fixNonFinite", err, errlen);
    if (rc || io.empty) {
        io.release();
        return rc;
    }
    if (dst->basetype != src->basetype) {
        io.release();
        set_err(err, errlen, "fixNonFinite: dst and src must have the same pixel type");
        return B2R_ERR_TYPES;
    }
    const bool fp    = src->basetype == B2R_FLOAT || src->basetype == B2R_HALF;
    const int es     = basetype_size(src->basetype);
    const bool count_only = (mode == 0 || mode == 100);
    int launched     = 0;
    int* counter     = nullptr;
    void* inplace_tmp = nullptr;
    b2r_image S      = io.srcd;
    void* spix       = io.dsrc;
    // Integer types cannot hold NaN/Inf: fixNonFinite is just the copy
    // (:1586-1587); the counting modes copy too.  (The repair kernel writes
    // every channel of every ROI pixel itself.)
    if ((!fp || count_only) && (io.dst_alloc || !io.same_pixels)) {
        if (S.xstride != (int64_t)S.nchannels * es
            || io.dstd.xstride != (int64_t)S.nchannels * es) {
            io.release();
            set_err(err, errlen, "fixNonFinite: pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        if (io.dst_alloc)
            cudaMemset2DAsync(io.dst_alloc, io.dpitch, 0, io.drowbytes, io.roi_h, io.stream);
        const int x0 = std::max(io.roi.xbegin, S.x), x1 = std::min(io.roi.xend, S.x + S.width);
        const int y0 = std::max(io.roi.ybegin, S.y), y1 = std::min(io.roi.yend, S.y + S.height);
        if (x1 > x0 && y1 > y0) {
            unsigned char* dp = (unsigned char*)io.ddst
                                + (long long)(y0 - io.dstd.y) * io.dstd.ystride
                                + (long long)(x0 - io.dstd.x) * io.dstd.xstride;
            const unsigned char* sp = (const unsigned char*)spix
                                      + (long long)(y0 - S.y) * S.ystride
                                      + (long long)(x0 - S.x) * S.xstride;
            cudaMemcpy2DAsync(dp, io.dstd.ystride, sp, S.ystride, size_t(x1 - x0) * S.xstride,
                              y1 - y0, cudaMemcpyDeviceToDevice, io.stream);
        }
    }
    if (fp) {
        if (cudaMallocAsync((void**)&counter, sizeof(int), io.stream) != cudaSuccess
            || cudaMemsetAsync(counter, 0, sizeof(int), io.stream) != cudaSuccess) {
            io.release();
            set_err(err, errlen, "CUDA error allocating the counter");
            return B2R_ERR_CUDA;
        }
        if (io.same_pixels && mode == 2 && !io.src_alloc) {
            // box3 in place on the device: neighbours must be read unrepaired
            const size_t rowb = size_t(S.width) * S.nchannels * es;
            if (cudaMallocAsync(&inplace_tmp, rowb * S.height, io.stream) != cudaSuccess) {
                cudaFreeAsync(counter, io.stream);
                io.release();
                set_err(err, errlen, "CUDA error allocating the in-place copy");
                return B2R_ERR_CUDA;
            }
            cudaMemcpy2DAsync(inplace_tmp, rowb, spix, S.ystride, rowb, S.height,
                              cudaMemcpyDeviceToDevice, io.stream);
            S.ystride = (int64_t)rowb;
            S.xstride = (int64_t)S.nchannels * es;
            spix      = inplace_tmp;
        }
        NonFiniteParams np;
        memset(&np, 0, sizeof(np));
        np.dst     = to_dev(io.dstd, io.ddst);
        np.src     = to_dev(S, spix);
        np.roi_x0  = io.roi.xbegin;
        np.roi_x1  = io.roi.xend;
        np.roi_y0  = io.roi.ybegin;
        np.roi_y1  = io.roi.yend;
        np.chbegin = io.roi.chbegin;
        np.chend   = io.roi.chend;
        np.mode    = mode;
        np.write   = count_only ? 0 : 1;
        np.counter = counter;
        launch_fix_nonfinite(np, io.stream);
        launched = 1;
    }
    int host_count = 0;
    if (counter) {
        cudaMemcpyAsync(&host_count, counter, sizeof(int), cudaMemcpyDeviceToHost, io.stream);
        cudaFreeAsync(counter, io.stream);
    }
    if (inplace_tmp)
        cudaFreeAsync(inplace_tmp, io.stream);
    rc = io.end(*dst, report, launched, err, errlen);
    if (rc)
        return rc;
    if (counter) {
        cudaError_t e = cudaStreamSynchronize(io.stream);  // the count is a return value
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
            return B2R_ERR_CUDA;
        }
    }
    if (pixels_fixed)
        *pixels_fixed = host_count;
    if (mode == 100 && host_count) {
        set_err(err, errlen, "Nonfinite pixel values found");
        return B2R_ERR_INVALID;
    }
    return B2R_OK;
}


// =========================================================================
// fillholes_pushpull (imagebufalgo.cpp:1600-1670): a float pyramid made with
// the triangle-filter resize, pushed down (divide by alpha) and pulled back up
// (over), entirely on the device.
// =========================================================================
extern "C" int
b2r_fillholes_pushpull(const b2r_image* dst, const b2r_image* src, int alpha_channel,
                       const b2r_options* opt, b2r_report* report, char* err, size_t errlen)
{
    int rc = check_image(src, true, err, errlen);
    if (rc)
        return rc;
    if (alpha_channel < 0 || alpha_channel >= src->nchannels) {
        set_err(err, errlen, "images must have alpha channels");
        return B2R_ERR_INVALID;
    }
    // the result is pasted back at the source's origin (:1659): the region
    // written is src's data window (clipped to dst's)
    b2r_roi roi { src->x, src->x + src->width, src->y, src->y + src->height, 0,
                  src->nchannels };
    StencilIO io;
    rc = io.begin(dst, src, &roi, opt, report, "fillholes_pushpull", err, errlen);
    if (rc || io.empty) {
        io.release();
        return rc;
    }
    const int nch = src->nchannels;
    const int W = src->width, H = src->height;
    struct Lv {
        int w, h;
        float* px;
    };
    std::vector<Lv> pyr;
    float* blow = nullptr;
    auto fail = [&](int code) {
        for (auto& l : pyr)
            if (l.px)
                cudaFreeAsync(l.px, io.stream);
        if (blow)
            cudaFreeAsync(blow, io.stream);
        io.release();
        return code;
    };
    for (int w = W, h = H;;) {
        Lv l { w, h, nullptr };
        if (cudaMallocAsync((void**)&l.px, size_t(w) * h * nch * 4, io.stream) != cudaSuccess) {
            set_err(err, errlen, "fillholes_pushpull: out of device memory");
            return fail(B2R_ERR_CUDA);
        }
        pyr.push_back(l);
        if (w <= 1 && h <= 1)
            break;
        w = std::max(1, w / 2);
        h = std::max(1, h / 2);
    }
    auto describe = [&](const Lv& l) {
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
        im.device   = io.device;
        return im;
    };
    b2r_options sub;
    memset(&sub, 0, sizeof(sub));
    sub.device      = io.device;
    sub.cuda_stream = (void*)io.stream;
    int launched    = 0;
    {   // top = float copy of src shifted to the origin (paste, :1620-1634):
        // a nearest resample between equal windows is the identity + conversion
        b2r_image T = describe(pyr[0]);
        b2r_image S = io.srcd;
        S.pixels    = io.dsrc;
        S.x = S.y = S.full_x = S.full_y = 0;
        S.full_width  = S.width;
        S.full_height = S.height;
        S.memspace    = B2R_MEM_DEVICE;
        S.device      = io.device;
        rc = b2r_resample(&T, &S, 0, nullptr, &sub, nullptr, err, errlen);
        if (rc)
            return fail(rc);
        ++launched;
    }
    for (size_t i = 1; i < pyr.size(); ++i) {
        b2r_image D = describe(pyr[i]), S = describe(pyr[i - 1]);
        rc = b2r_resize(&D, &S, "triangle", 0.0f, nullptr, &sub, nullptr, err, errlen);
        if (rc)
            return fail(rc);
        launch_divide_by_alpha(pyr[i].px, (long long)pyr[i].w * pyr[i].h, nch, alpha_channel,
                               io.stream);
        launched += 3;
    }
    if (pyr.size() > 1) {
        if (cudaMallocAsync((void**)&blow, size_t(W) * H * nch * 4, io.stream) != cudaSuccess) {
            set_err(err, errlen, "fillholes_pushpull: out of device memory");
            return fail(B2R_ERR_CUDA);
        }
        for (int i = (int)pyr.size() - 2; i >= 0; --i) {
            Lv b { pyr[i].w, pyr[i].h, blow };
            b2r_image D = describe(b), S = describe(pyr[i + 1]);
            rc = b2r_resize(&D, &S, "triangle", 0.0f, nullptr, &sub, nullptr, err, errlen);
            if (rc)
                return fail(rc);
            launch_over_inplace(pyr[i].px, blow, (long long)pyr[i].w * pyr[i].h, nch,
                                alpha_channel, io.stream);
            launched += 3;
        }
    }
    {   // paste the finished top level back at src's origin, in dst's type
        const b2r_image& D = io.dstd;
        const int x0 = io.roi.xbegin, y0 = io.roi.ybegin;  // inside src's window
        if (D.xstride != (int64_t)nch * basetype_size(D.basetype)) {
            set_err(err, errlen, "destination pixels must be contiguous in x");
            return fail(B2R_ERR_UNSUPPORTED);
        }
        unsigned char* dp = (unsigned char*)io.ddst + (long long)(y0 - D.y) * D.ystride
                            + (long long)(x0 - D.x) * D.xstride;
        const float* sp = pyr[0].px + ((size_t)(y0 - src->y) * W + (x0 - src->x)) * nch;
        launch_convert_from_float(sp, (long long)W * nch * 4, dp, D.ystride, io.roi_w * nch,
                                  io.roi_h, D.basetype, io.stream);
        ++launched;
    }
    for (auto& l : pyr)
        cudaFreeAsync(l.px, io.stream);
    pyr.clear();
    if (blow)
        cudaFreeAsync(blow, io.stream);
    blow = nullptr;
    return io.end(*dst, report, launched, err, errlen);
}


// =========================================================================
// computePixelStats / isMonochrome (imagebufalgo_compare.cpp:81-215, 539-600)
// =========================================================================
extern "C" int
b2r_pixel_stats(const b2r_image* src, const b2r_roi* roi_in, b2r_pixel_stats_t* stats,
                int nstats, int* is_monochrome, float mono_threshold, const b2r_options* opt,
                char* err, size_t errlen)
{
    int rc = check_image(src, true, err, errlen);
    if (rc)
        return rc;
    if (!stats || nstats < src->nchannels) {
        set_err(err, errlen, "pixel_stats: need one stats record per channel (%d)",
                src->nchannels);
        return B2R_ERR_INVALID;
    }
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 pixel statistics path has no CPU "
                "fallback");
        return B2R_ERR_NO_DEVICE;
    }
    const int nch = src->nchannels;
    b2r_roi roi;
    if (roi_in) {
        roi        = *roi_in;
        roi.chbegin = std::max(roi.chbegin, 0);
        roi.chend  = std::min(roi.chend, nch);
    } else {
        roi = b2r_roi { src->x, src->x + src->width, src->y, src->y + src->height, 0, nch };
    }
    // PixelStats::reset
    for (int c = 0; c < nch; ++c) {
        memset(&stats[c], 0, sizeof(stats[c]));
    }
    if (is_monochrome)
        *is_monochrome = 1;
    // pixels of the ROI outside the data window read as black through the
    // iterator; restrict to the window and account for them as zeros below
    const int x0 = std::max(roi.xbegin, src->x), x1 = std::min(roi.xend, src->x + src->width);
    const int y0 = std::max(roi.ybegin, src->y), y1 = std::min(roi.yend, src->y + src->height);
    const long long roi_px = (long long)std::max(0, roi.xend - roi.xbegin)
                             * std::max(0, roi.yend - roi.ybegin);
    const long long in_px = (long long)std::max(0, x1 - x0) * std::max(0, y1 - y0);
    std::vector<StatsPartial> total((size_t)nch);
    for (auto& t : total) {
        t.sum = t.sum2 = 0.0;
        t.mn  = INFINITY;
        t.mx  = -INFINITY;
        t.finite = t.nan = t.inf = 0;
    }
    bool mono = true;
    if (in_px > 0 && roi.chend > roi.chbegin) {
        int dev = opt ? opt->device : 0;
        const int mem = memspace_of(src, &dev);
        if (dev < 0 || dev >= device_count_cached()) {
            set_err(err, errlen, "invalid CUDA device ordinal %d", dev);
            return B2R_ERR_INVALID;
        }
        DeviceGuard guard(dev);
        device_info(dev);
        cudaStream_t stream = opt ? (cudaStream_t)opt->cuda_stream : nullptr;
        b2r_image S = *src;
        void* spix  = src->pixels;
        void* stage = nullptr;
        const int es = basetype_size(src->basetype);
        if (mem == B2R_MEM_HOST) {
            if (src->xstride != (int64_t)nch * es) {
                set_err(err, errlen, "host source pixels must be contiguous in x");
                return B2R_ERR_UNSUPPORTED;
            }
            const size_t rowbytes = size_t(x1 - x0) * src->xstride;
            const size_t pitch    = (rowbytes + 15) & ~size_t(15);
            CUDA_TRY(cudaMallocAsync(&stage, pitch * (y1 - y0), stream));
            const unsigned char* hp = (const unsigned char*)src->pixels
                                      + (long long)(y0 - src->y) * src->ystride
                                      + (long long)(x0 - src->x) * src->xstride;
            CUDA_TRY(copy_h2d_2d(stage, pitch, hp, src->ystride, rowbytes, y1 - y0, stream));
            S.x = x0, S.y = y0, S.width = x1 - x0, S.height = y1 - y0;
            S.ystride = (int64_t)pitch;
            spix      = stage;
        }
        const int nblocks = stats_grid_blocks(y1 - y0);
        StatsPartial* dpart = nullptr;
        int* dmono          = nullptr;
        CUDA_TRY(cudaMallocAsync((void**)&dpart, sizeof(StatsPartial) * nblocks * nch, stream));
        CUDA_TRY(cudaMallocAsync((void**)&dmono, sizeof(int) * nblocks, stream));
        StatsParams sp;
        memset(&sp, 0, sizeof(sp));
        sp.src       = to_dev(S, spix);
        sp.roi_x0    = x0, sp.roi_x1 = x1, sp.roi_y0 = y0, sp.roi_y1 = y1;
        sp.chbegin   = roi.chbegin;
        sp.chend     = roi.chend;
        sp.nchannels = nch;
        sp.partials  = dpart;
        sp.want_mono = is_monochrome ? 1 : 0;
        sp.mono_threshold = mono_threshold;
        sp.mono_flags     = dmono;
        launch_pixel_stats(sp, nblocks, stream);
        std::vector<StatsPartial> hpart((size_t)nblocks * nch);
        std::vector<int> hmono((size_t)nblocks, 1);
        CUDA_TRY(cudaMemcpyAsync(hpart.data(), dpart, sizeof(StatsPartial) * hpart.size(),
                                 cudaMemcpyDeviceToHost, stream));
        if (is_monochrome)
            CUDA_TRY(cudaMemcpyAsync(hmono.data(), dmono, sizeof(int) * nblocks,
                                     cudaMemcpyDeviceToHost, stream));
        cudaFreeAsync(dpart, stream);
        cudaFreeAsync(dmono, stream);
        if (stage)
            cudaFreeAsync(stage, stream);
        CUDA_TRY(cudaStreamSynchronize(stream));
        {
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) {
                set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(e));
                return B2R_ERR_CUDA;
            }
        }
        for (int b = 0; b < nblocks; ++b) {  // PixelStats::merge, in block order
            for (int c = roi.chbegin; c < roi.chend; ++c) {
                const StatsPartial& q = hpart[(size_t)b * nch + c];
                StatsPartial& t       = total[c];
                t.mn = std::min(t.mn, q.mn);
                t.mx = std::max(t.mx, q.mx);
                t.nan += q.nan, t.inf += q.inf, t.finite += q.finite;
                t.sum += q.sum, t.sum2 += q.sum2;
            }
            mono = mono && hmono[b] != 0;
        }
    }
    // black pixels of the ROI beyond the data window: value 0 in every channel
    if (roi_px > in_px)
        for (int c = roi.chbegin; c < roi.chend; ++c) {
            total[c].finite += (unsigned long long)(roi_px - in_px);
            total[c].mn = std::min(total[c].mn, 0.0f);
            total[c].mx = std::max(total[c].mx, 0.0f);
        }
    for (int c = 0; c < nch; ++c) {  // finalize (:101-118)
        const StatsPartial& t = total[c];
        b2r_pixel_stats_t& o  = stats[c];
        o.nancount    = (int64_t)t.nan;
        o.infcount    = (int64_t)t.inf;
        o.finitecount = (int64_t)t.finite;
        o.sum         = t.sum;
        o.sum2        = t.sum2;
        if (t.finite == 0) {
            o.min = o.max = o.avg = o.stddev = 0.0f;
        } else {
            const double count = double(t.finite);
            const double davg  = t.sum / count;
            const double var   = t.sum2 / count - davg * davg;
            o.min    = t.mn;
            o.max    = t.mx;
            o.avg    = float(davg);
            o.stddev = float(var > 0.0 ? sqrt(var) : 0.0);
        }
    }
    if (is_monochrome)
        *is_monochrome = (nch < 2 || mono) ? 1 : 0;
    return B2R_OK;
}
