// api_internal.h -- helpers shared by the translation units that implement
// the C ABI (include/b2r.h): error text, image checks, device bookkeeping and
// the host<->device staging used by the stencil-style entry points.
// Internal to libb2r.so; nothing here is exported.
#pragma once

#include "b2r.h"

#include "host_copy.h"
#include "kernels.h"
#include "recon_filter.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <vector>

struct b2r_filter2d {
    b2r::Filter2D* f;
};
struct b2r_filter1d {
    b2r::Filter1D* f;
};


// device-resident MIP chain (b2r_mipchain_create*, b2r_api.cpp)
struct b2r_mipchain {
    int device = 0;
    int nch    = 0;
    struct Level {
        int w, h;
        void* px;
        bool owned;
    };
    std::vector<Level> levels;
    float kernel_ms = 0.0f;
    int launches    = 0;
};

namespace b2r {
namespace api {


inline void
set_err(char* err, size_t errlen, const char* fmt, ...)
{
    if (!err || !errlen)
        return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

#define CUDA_TRY(call)                                                        \
    do {                                                                      \
        cudaError_t e_ = (call);                                              \
        if (e_ != cudaSuccess) {                                              \
            set_err(err, errlen, "CUDA error: %s (%s)", cudaGetErrorString(e_), \
                    #call);                                                   \
            return B2R_ERR_CUDA;                                              \
        }                                                                     \
    } while (0)

inline bool
valid_basetype(int bt)
{
    return bt == B2R_UINT8 || bt == B2R_UINT16 || bt == B2R_HALF
           || bt == B2R_FLOAT;
}

// OIIO_DISPATCH_COMMON_TYPES2: the ten native (dst,src) pairs.
inline bool
legal_pair(int d, int s)
{
    if (d == B2R_FLOAT)
        return valid_basetype(s);
    return s == B2R_FLOAT || s == d;
}

inline int
device_count_cached()
{
    static int n = -1;
    static std::once_flag once;
    std::call_once(once, [] {
        int c = 0;
        if (cudaGetDeviceCount(&c) != cudaSuccess)
            c = 0;
        cudaGetLastError();
        n = c;
    });
    return n;
}

struct DeviceInfo {
    int sms = 0;
    bool pool_ready = false;
};
inline DeviceInfo&
device_info(int dev)
{
    static DeviceInfo info[64];
    static std::mutex m;
    std::lock_guard<std::mutex> lock(m);
    DeviceInfo& d = info[dev & 63];
    if (!d.sms) {
        cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;  // keep freed blocks cached
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            d.pool_ready = true;
        }
    }
    return d;
}

inline int
memspace_of(const b2r_image* im, int* device)
{
    if (im->memspace == B2R_MEM_HOST)
        return B2R_MEM_HOST;
    if (im->memspace == B2R_MEM_DEVICE) {
        *device = im->device;
        return B2R_MEM_DEVICE;
    }
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, im->pixels) == cudaSuccess
        && (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged)) {
        *device = a.device;
        return B2R_MEM_DEVICE;
    }
    cudaGetLastError();
    return B2R_MEM_HOST;
}

inline DevImage
to_dev(const b2r_image& im, void* pixels)
{
    DevImage d;
    d.pixels      = pixels;
    d.basetype    = im.basetype;
    d.nchannels   = im.nchannels;
    d.x           = im.x;
    d.y           = im.y;
    d.width       = im.width;
    d.height      = im.height;
    d.full_x      = im.full_x;
    d.full_y      = im.full_y;
    d.full_width  = im.full_width;
    d.full_height = im.full_height;
    d.xstride     = im.xstride;
    d.ystride     = im.ystride;
    return d;
}

inline int
check_image(const b2r_image* im, bool is_src, char* err, size_t errlen)
{
    if (!im || !im->pixels || im->width <= 0 || im->height <= 0
        || im->nchannels <= 0) {
        set_err(err, errlen,
                is_src ? "Uninitialized input image"
                       : "Uninitialized destination image");
        return B2R_ERR_INVALID;
    }
    if (!valid_basetype(im->basetype)) {
        set_err(err, errlen, "unsupported pixel data type %d", im->basetype);
        return B2R_ERR_TYPES;
    }
    if (im->full_width <= 0 || im->full_height <= 0) {
        set_err(err, errlen, "image has an empty display window");
        return B2R_ERR_INVALID;
    }
    return B2R_OK;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev)
            cudaSetDevice(dev);
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
    }
};


// caches behind b2r_trim_caches (b2r_api_banded.cpp, b2r_api_tiles.cpp)
void
trim_band_blocks();
void
trim_tile_slots();

// Runs at every exit of an entry point: staging blocks that are still set go
// back to the pool (stream-ordered), events that are still set are destroyed.
// The normal path frees / destroys them itself and clears the variables.
struct ScopeCleanup {
    cudaStream_t stream;
    void** blocks[3];
    cudaEvent_t* events[2];
    ~ScopeCleanup()
    {
        for (void** b : blocks)
            if (b && *b) {
                cudaFreeAsync(*b, stream);
                *b = nullptr;
            }
        for (cudaEvent_t* e : events)
            if (e && *e) {
                cudaEventDestroy(*e);
                *e = nullptr;
            }
    }
};

// Host<->device staging shared by the stencil entry points: the whole source
// data window goes up (a stencil may read any of it), the dst ROI comes back.
struct StencilIO {
    b2r_roi roi;
    b2r_image srcd, dstd;
    void *dsrc = nullptr, *ddst = nullptr;
    void *src_alloc = nullptr, *dst_alloc = nullptr;
    unsigned char* host_out = nullptr;
    size_t dpitch = 0, drowbytes = 0;
    int roi_w = 0, roi_h = 0, device = 0;
    bool empty       = false;
    bool same_pixels = false;  // dst and src are one image (in place)
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    std::unique_ptr<DeviceGuard> guard;

    int begin(const b2r_image* dst_in, const b2r_image* src_in, const b2r_roi* roi_in,
              const b2r_options* opt, b2r_report* report, const char* what, char* err,
              size_t errlen, bool same_nchannels = true)
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
        if (same_nchannels && dst.nchannels != src.nchannels) {
            set_err(err, errlen, "images must have the same number of channels");
            return B2R_ERR_INVALID;
        }
        if (device_count_cached() <= 0) {
            set_err(err, errlen,
                    "no CUDA device available: the B200 %s path has no CPU fallback", what);
            return B2R_ERR_NO_DEVICE;
        }
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
        if (roi.xend <= roi.xbegin || roi.yend <= roi.ybegin || roi.chend <= roi.chbegin) {
            empty = true;
            return B2R_OK;
        }
        roi_w = roi.xend - roi.xbegin, roi_h = roi.yend - roi.ybegin;
        const int ses = basetype_size(src.basetype);
        const int des = basetype_size(dst.basetype);
        int sdev = opt ? opt->device : 0, ddev = sdev;
        const int smem = memspace_of(&src, &sdev);
        const int dmem = memspace_of(&dst, &ddev);
        device         = opt ? opt->device : 0;
        if (smem == B2R_MEM_DEVICE)
            device = sdev;
        else if (dmem == B2R_MEM_DEVICE)
            device = ddev;
        if (device < 0 || device >= device_count_cached()) {
            set_err(err, errlen, "invalid CUDA device ordinal %d", device);
            return B2R_ERR_INVALID;
        }
        guard.reset(new DeviceGuard(device));
        device_info(device);
        stream      = opt ? (cudaStream_t)opt->cuda_stream : nullptr;
        same_pixels = dst.pixels == src.pixels;
        srcd = src, dstd = dst;
        dsrc = src.pixels, ddst = dst.pixels;
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
        if (dmem == B2R_MEM_HOST) {
            if (dst.xstride != (int64_t)dst.nchannels * des) {
                set_err(err, errlen, "host destination pixels must be contiguous in x");
                return B2R_ERR_UNSUPPORTED;
            }
            drowbytes = size_t(roi_w) * dst.xstride;
            dpitch    = (drowbytes + 15) & ~size_t(15);
            CUDA_TRY(cudaMallocAsync(&dst_alloc, dpitch * roi_h, stream));
            host_out = (unsigned char*)dst.pixels
                       + (long long)(roi.ybegin - dst.y) * dst.ystride
                       + (long long)(roi.xbegin - dst.x) * dst.xstride;
            if (roi.chbegin != 0 || roi.chend != dst.nchannels) {
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
        if (report) {
            cudaEventCreate(&ev0);
            cudaEventCreate(&ev1);
            cudaEventRecord(ev0, stream);
        }
        return B2R_OK;
    }

    int end(const b2r_image& dst, b2r_report* report, int launched, char* err, size_t errlen)
    {
        if (report)
            cudaEventRecord(ev1, stream);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            set_err(err, errlen, "CUDA launch error: %s", cudaGetErrorString(e));
            return B2R_ERR_CUDA;
        }
        if (dst_alloc) {
            CUDA_TRY(copy_d2h_2d(host_out, dst.ystride, dst_alloc, dpitch, drowbytes,
                                       roi_h, stream));
            if (report)
                report->d2h_bytes += (int64_t)drowbytes * roi_h;
        }
        const bool staged = dst_alloc || src_alloc || extra_alloc;
        release();
        if (staged || report)
            CUDA_TRY(cudaStreamSynchronize(stream));
        if (report) {
            cudaEventElapsedTime(&report->kernel_ms, ev0, ev1);
            cudaEventDestroy(ev0);
            cudaEventDestroy(ev1);
            ev0 = ev1                = nullptr;
            report->kernels_launched = launched;
            report->path_used        = B2R_PATH_EXACT;
        }
        return B2R_OK;
    }

    // true when dst and src share device pixels even though the caller's host
    // pointers differ (never: each host image gets its own staging buffer)
    bool dst_alloc_is_src() const { return ddst == dsrc; }

    // a further input image (st_warp's ST buffer): whole data window up
    void* extra_alloc = nullptr;
    int stage_extra(const b2r_image& im, b2r_image* imd, void** dpix, b2r_report* report,
                    char* err, size_t errlen)
    {
        *imd  = im;
        *dpix = im.pixels;
        int dev = device;
        if (memspace_of(&im, &dev) != B2R_MEM_HOST)
            return B2R_OK;
        const int es = basetype_size(im.basetype);
        if (im.xstride != (int64_t)im.nchannels * es) {
            set_err(err, errlen, "host pixels must be contiguous in x");
            return B2R_ERR_UNSUPPORTED;
        }
        const size_t rowbytes = size_t(im.width) * im.xstride;
        const size_t pitch    = (rowbytes + 15) & ~size_t(15);
        CUDA_TRY(cudaMallocAsync(&extra_alloc, pitch * im.height, stream));
        CUDA_TRY(copy_h2d_2d(extra_alloc, pitch, im.pixels, im.ystride, rowbytes,
                                   im.height, stream));
        imd->ystride = (int64_t)pitch;
        *dpix        = extra_alloc;
        if (report)
            report->h2d_bytes += (int64_t)rowbytes * im.height;
        return B2R_OK;
    }

    void release()
    {
        if (dst_alloc)
            cudaFreeAsync(dst_alloc, stream);
        if (src_alloc)
            cudaFreeAsync(src_alloc, stream);
        if (extra_alloc)
            cudaFreeAsync(extra_alloc, stream);
        dst_alloc = src_alloc = extra_alloc = nullptr;
    }
    // every early return (CUDA_TRY, argument errors after begin) ends here: the
    // staging buffers go back to the pool stream-ordered, the events are destroyed
    ~StencilIO()
    {
        release();
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
    }
};


}  // namespace api
}  // namespace b2r
