// kernels_aux.cu -- the small bandwidth-bound kernels around the resize:
// resample (nearest / bilinear), maketx's box downsample, clamp.
//
// All index and blend arithmetic uses explicit round-to-nearest mul/add
// intrinsics: the reference computes them unfused, and a fused multiply-add
// moves results across integer / rounding boundaries
// (imagebufalgo_xform.cpp:1337-1341 warns about exactly that).
#include "dev_convert.cuh"
#include "kernels.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

namespace b2r {

namespace {

__device__ __forceinline__ float
ld_elem(const unsigned char* p, int bt)
{
    switch (bt) {
    case BT_UINT8: return __fmul_rn(float(*p), 1.0f / 255.0f);
    case BT_UINT16:
        return __fmul_rn(float(*(const unsigned short*)p), 1.0f / 65535.0f);
    case BT_HALF: return __half2float(*(const __half*)p);
    default: return *(const float*)p;
    }
}

__device__ __forceinline__ unsigned int
q_int(float v, float mx)
{
    float s = __fmul_rn(v, mx);
    s       = __fadd_rn(s, (s < 0.0f) ? -0.5f : 0.5f);
    if (!(0.0f <= s))
        s = 0.0f;
    if (s > mx)
        s = mx;
    return (unsigned int)s;
}

__device__ __forceinline__ void
st_elem(unsigned char* p, int bt, float v)
{
    switch (bt) {
    case BT_UINT8: *p = (unsigned char)q_int(v, 255.0f); break;
    case BT_UINT16: *(unsigned short*)p = (unsigned short)q_int(v, 65535.0f); break;
    case BT_HALF: *(__half*)p = __float2half_rn(v); break;
    default: *(float*)p = v; break;
    }
}

__device__ __forceinline__ int
esize(int bt)
{
    return bt == BT_UINT8 ? 1 : (bt == BT_FLOAT ? 4 : 2);
}

__device__ __forceinline__ bool
in_data(const DevImage& im, int x, int y)
{
    return x >= im.x && x < im.x + im.width && y >= im.y && y < im.y + im.height;
}

// WrapBlack (wrap_clamp=false) or WrapClamp addressing; nullptr = black.
__device__ __forceinline__ const unsigned char*
src_px(const DevImage& im, int x, int y, bool wrap_clamp)
{
    if (!in_data(im, x, y)) {
        if (!wrap_clamp)
            return nullptr;
        x = min(max(x, im.full_x), im.full_x + im.full_width - 1);
        y = min(max(y, im.full_y), im.full_y + im.full_height - 1);
        if (!in_data(im, x, y))
            return nullptr;
    }
    return (const unsigned char*)im.pixels + (long long)(y - im.y) * im.ystride
           + (long long)(x - im.x) * im.xstride;
}

// bilerp, fmath.h:430-440, unfused
__device__ __forceinline__ float
bilerp1(float v0, float v1, float v2, float v3, float s, float t)
{
    float s1 = __fsub_rn(1.0f, s);
    float t1 = __fsub_rn(1.0f, t);
    float top = __fadd_rn(__fmul_rn(v0, s1), __fmul_rn(v1, s));
    float bot = __fadd_rn(__fmul_rn(v2, s1), __fmul_rn(v3, s));
    return __fadd_rn(__fmul_rn(t1, top), __fmul_rn(t, bot));
}

// ---------------------------------------------------------------------------
// resample: imagebufalgo_xform.cpp:1107-1186 (scalar), :1379-1493 (tables)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
resample_kernel(ResampleParams p, int chbegin, int chend, bool vec4)
{
    pdl_prologue();
    // 2-D grid (rows in blockIdx.y, grid-strided): no 64-bit div/mod
    const int kx = blockIdx.x * blockDim.x + threadIdx.x;
    if (kx >= p.roi_x1 - p.roi_x0)
        return;
    const int x = p.roi_x0 + kx;
    for (int y = p.roi_y0 + blockIdx.y; y < p.roi_y1; y += gridDim.y) {

    const float srcfx = float(p.src.full_x), srcfy = float(p.src.full_y);
    const float srcfw = float(p.src.full_width), srcfh = float(p.src.full_height);
    const float dstfx = float(p.dst.full_x), dstfy = float(p.dst.full_y);
    const float dpw   = __fdiv_rn(1.0f, float(p.dst.full_width));
    const float dph   = __fdiv_rn(1.0f, float(p.dst.full_height));

    float t      = __fmul_rn(__fadd_rn(__fsub_rn(float(y), dstfy), 0.5f), dph);
    float src_yf = __fadd_rn(srcfy, __fmul_rn(t, srcfh));
    float s      = __fmul_rn(__fadd_rn(__fsub_rn(float(x), dstfx), 0.5f), dpw);
    float src_xf = __fadd_rn(srcfx, __fmul_rn(s, srcfw));

    const int ses = esize(p.src.basetype), des = esize(p.dst.basetype);
    unsigned char* dp = (unsigned char*)p.dst.pixels
                        + (long long)(y - p.dst.y) * p.dst.ystride
                        + (long long)(x - p.dst.x) * p.dst.xstride;
    if (vec4) {
        // RGBA float -> RGBA float, 16-byte aligned: whole-pixel loads/stores,
        // same arithmetic per channel
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        if (!p.interpolate) {
            int sx = (int)floorf(src_xf), sy = (int)floorf(src_yf);
            const float4* sp = (const float4*)src_px(p.src, sx, sy, false);
            *(float4*)dp     = sp ? __ldg(sp) : zero;
            continue;
        }
        float xx = __fsub_rn(src_xf, 0.5f), yy = __fsub_rn(src_yf, 0.5f);
        float fxf = floorf(xx), fyf = floorf(yy);
        int xt = (int)fxf, yt = (int)fyf;
        float xfrac = __fsub_rn(xx, fxf), yfrac = __fsub_rn(yy, fyf);
        float4 v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int px = xt + (k & 1), py = yt + (k >> 1);
            const float4* q;
            if (p.data_is_full) {
                px = min(max(px, p.src.x), p.src.x + p.src.width - 1);
                py = min(max(py, p.src.y), p.src.y + p.src.height - 1);
                q  = (const float4*)src_px(p.src, px, py, false);
            } else {
                q = (const float4*)src_px(p.src, px, py, true);
            }
            v[k] = q ? __ldg(q) : zero;
        }
        float4 o;
        o.x = bilerp1(v[0].x, v[1].x, v[2].x, v[3].x, xfrac, yfrac);
        o.y = bilerp1(v[0].y, v[1].y, v[2].y, v[3].y, xfrac, yfrac);
        o.z = bilerp1(v[0].z, v[1].z, v[2].z, v[3].z, xfrac, yfrac);
        o.w = bilerp1(v[0].w, v[1].w, v[2].w, v[3].w, xfrac, yfrac);
        *(float4*)dp = o;
        continue;
    }
    if (!p.interpolate) {
        int sx = (int)floorf(src_xf), sy = (int)floorf(src_yf);
        const unsigned char* sp = src_px(p.src, sx, sy, false);
        for (int c = chbegin; c < chend; ++c)
            st_elem(dp + c * des, p.dst.basetype,
                    sp ? ld_elem(sp + c * ses, p.src.basetype) : 0.0f);
        continue;
    }
    float xx = __fsub_rn(src_xf, 0.5f), yy = __fsub_rn(src_yf, 0.5f);
    float fxf = floorf(xx), fyf = floorf(yy);
    int xt = (int)fxf, yt = (int)fyf;
    float xfrac = __fsub_rn(xx, fxf), yfrac = __fsub_rn(yy, fyf);
    const unsigned char* q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int px = xt + (k & 1), py = yt + (k >> 1);
        if (p.data_is_full) {
            px   = min(max(px, p.src.x), p.src.x + p.src.width - 1);
            py   = min(max(py, p.src.y), p.src.y + p.src.height - 1);
            q[k] = src_px(p.src, px, py, false);
        } else {
            q[k] = src_px(p.src, px, py, true);
        }
    }
    for (int c = chbegin; c < chend; ++c) {
        float v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            v[k] = q[k] ? ld_elem(q[k] + c * ses, p.src.basetype) : 0.0f;
        st_elem(dp + c * des, p.dst.basetype,
                bilerp1(v[0], v[1], v[2], v[3], xfrac, yfrac));
    }
    }  // row loop
}

// ---------------------------------------------------------------------------
// maketx box downsample (float32 only), maketexture.cpp:141-334
// ---------------------------------------------------------------------------
// resize_block_2pass (:260-300): 0.5f*(0.5f*(a+b) + 0.5f*(c+d)).  One thread
// per 4 consecutive dst elements when the row is 16-byte friendly.
__global__ void __launch_bounds__(256)
box2_kernel(const float* __restrict__ src, long long src_ystride_f,
            float* __restrict__ dst, long long dst_ystride_f, int dw, int dh,
            int nch)
{
    pdl_prologue();
    const long long row_elems = (long long)dw * nch;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= row_elems * dh)
        return;
    const int y       = int(idx / row_elems);
    const int e       = int(idx % row_elems);
    const int x       = e / nch;
    const int c       = e - x * nch;
    const float* r0   = src + (long long)(2 * y) * src_ystride_f;
    const float* r1   = r0 + src_ystride_f;
    const int a       = (2 * x) * nch + c;
    float h0 = __fmul_rn(0.5f, __fadd_rn(__ldg(r0 + a), __ldg(r0 + a + nch)));
    float h1 = __fmul_rn(0.5f, __fadd_rn(__ldg(r1 + a), __ldg(r1 + a + nch)));
    dst[(long long)y * dst_ystride_f + e] = __fmul_rn(0.5f, __fadd_rn(h0, h1));
}

// RGBA fast path of the same arithmetic: one thread per dst pixel, two
// 32-byte source reads, one 16-byte store.
__global__ void __launch_bounds__(256)
box2_rgba_kernel(const float4* __restrict__ src, long long src_ystride_4,
                 float4* __restrict__ dst, long long dst_ystride_4, int dw, int dh)
{
    pdl_prologue();
    // 2-D grid: no 64-bit div/mod per thread (they dominated a 1-D version)
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= dw)
        return;
    for (int y = blockIdx.y; y < dh; y += gridDim.y) {
    const float4* r0 = src + (long long)(2 * y) * src_ystride_4 + 2 * x;
    const float4* r1 = r0 + src_ystride_4;
    float4 a = __ldg(r0), b = __ldg(r0 + 1), c = __ldg(r1), d = __ldg(r1 + 1);
    float4 o;
    o.x = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, __fadd_rn(a.x, b.x)),
                                    __fmul_rn(0.5f, __fadd_rn(c.x, d.x))));
    o.y = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, __fadd_rn(a.y, b.y)),
                                    __fmul_rn(0.5f, __fadd_rn(c.y, d.y))));
    o.z = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, __fadd_rn(a.z, b.z)),
                                    __fmul_rn(0.5f, __fadd_rn(c.z, d.z))));
    o.w = __fmul_rn(0.5f, __fadd_rn(__fmul_rn(0.5f, __fadd_rn(a.w, b.w)),
                                    __fmul_rn(0.5f, __fadd_rn(c.w, d.w))));
    dst[(long long)y * dst_ystride_4 + x] = o;
    }
}

// resize_block_ (:205-241) + interppixel_NDC (:141-198, non-envlatl)
__global__ void __launch_bounds__(256)
box_bilinear_kernel(BoxParams p)
{
    pdl_prologue();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)p.dst.width * p.dst.height)
        return;
    const int x = p.dst.x + int(idx % p.dst.width);
    const int y = p.dst.y + int(idx / p.dst.width);
    const DevImage& S = p.src;
    const bool src_is_crop
        = (S.x > S.full_x || S.y > S.full_y
           || S.x + S.width < S.full_x + S.full_width
           || S.y + S.height < S.full_y + S.full_height);
    const float xoffset = float(p.dst.full_x), yoffset = float(p.dst.full_y);
    const float xscale  = __fdiv_rn(1.0f, float(p.dst.full_width));
    const float yscale  = __fdiv_rn(1.0f, float(p.dst.full_height));
    float t  = __fadd_rn(__fmul_rn(__fadd_rn(float(y), 0.5f), yscale), yoffset);
    float s  = __fadd_rn(__fmul_rn(__fadd_rn(float(x), 0.5f), xscale), xoffset);
    float fx = __fadd_rn(float(S.full_x), __fmul_rn(s, float(S.full_width)));
    float fy = __fadd_rn(float(S.full_y), __fmul_rn(t, float(S.full_height)));
    fx       = __fsub_rn(fx, 0.5f);
    fy       = __fsub_rn(fy, 0.5f);
    float flx = floorf(fx), fly = floorf(fy);
    int xt = (int)flx, yt = (int)fly;
    float xfrac = __fsub_rn(fx, flx), yfrac = __fsub_rn(fy, fly);
    const float* q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
        q[k] = (const float*)src_px(S, xt + (k & 1), yt + (k >> 1),
                                    p.force_clamp || !src_is_crop);
    float* d = (float*)((unsigned char*)p.dst.pixels
                        + (long long)(y - p.dst.y) * p.dst.ystride
                        + (long long)(x - p.dst.x) * p.dst.xstride);
    for (int c = 0; c < p.dst.nchannels; ++c) {
        float v0 = q[0] ? q[0][c] : 0.0f, v1 = q[1] ? q[1][c] : 0.0f;
        float v2 = q[2] ? q[2][c] : 0.0f, v3 = q[3] ? q[3][c] : 0.0f;
        d[c]     = bilerp1(v0, v1, v2, v3, xfrac, yfrac);
    }
}

// ImageBufAlgo::clamp(img, img, lo, hi, clampalpha01=true)
// (imagebufalgo_pixelmath.cpp:222-240) on contiguous float pixels.
__global__ void __launch_bounds__(256)
clamp_kernel(float* px, long long n, int nch, int alpha, float lo, float hi)
{
    pdl_prologue();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    float v = px[i];
    float r = v;
    if (!(lo <= r))
        r = lo;
    if (r > hi)
        r = hi;
    if (alpha >= 0 && int(i % nch) == alpha) {
        if (!(0.0f <= r))
            r = 0.0f;
        if (r > 1.0f)
            r = 1.0f;
    }
    px[i] = r;
}

// convert_type<float, D> over a rectangle (fmath.h:753-764, 1009-1031): the
// `R.copy(Rtmp, R.pixeltype())` step of OIIO_DISPATCH_COMMON_TYPES2 for
// (dst,src) pairs that are computed through a float temporary.
__global__ void __launch_bounds__(256)
convert_from_float_kernel(const float* __restrict__ src, long long src_ystride_f,
                          unsigned char* __restrict__ dst, long long dst_ystride,
                          int row_elems, int h, int bt)
{
    pdl_prologue();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= row_elems)
        return;
    const int es = esize(bt);
    for (int y = blockIdx.y; y < h; y += gridDim.y)
        st_elem(dst + (long long)y * dst_ystride + (long long)e * es, bt,
                src[(long long)y * src_ystride_f + e]);
}

}  // namespace

void
launch_convert_from_float(const float* src, long long src_ystride_bytes, void* dst,
                          long long dst_ystride_bytes, int row_elems, int height,
                          int dst_basetype, void* stream)
{
    if (row_elems <= 0 || height <= 0)
        return;
    dim3 grid((row_elems + 255) / 256, std::min(height, 32768), 1);
    launch_pdl(convert_from_float_kernel, grid, dim3(256), 0, (cudaStream_t)stream,
        src, src_ystride_bytes / 4, (unsigned char*)dst, dst_ystride_bytes, row_elems,
        height, dst_basetype);
}

void
launch_resample(const ResampleParams& p, int chbegin, int chend, void* stream)
{
    const int w = p.roi_x1 - p.roi_x0, h = p.roi_y1 - p.roi_y0;
    if (w <= 0 || h <= 0 || chend <= chbegin)
        return;
    dim3 grid((w + 255) / 256, std::min(h, 32768), 1);
    const bool vec4 = p.src.basetype == BT_FLOAT && p.dst.basetype == BT_FLOAT
                      && p.src.nchannels == 4 && p.dst.nchannels == 4 && chbegin == 0
                      && chend == 4 && p.src.xstride == 16 && p.dst.xstride == 16
                      && (uintptr_t)p.src.pixels % 16 == 0
                      && (uintptr_t)p.dst.pixels % 16 == 0 && p.src.ystride % 16 == 0
                      && p.dst.ystride % 16 == 0;
    launch_pdl(resample_kernel, grid, dim3(256), 0, (cudaStream_t)stream, p, chbegin, chend, vec4);
}

void
launch_box_downsample(const BoxParams& p, void* stream)
{
    const DevImage& D = p.dst;
    const DevImage& S = p.src;
    cudaStream_t st   = (cudaStream_t)stream;
    if (p.two_pass) {
        const bool rgba = D.nchannels == 4 && ((uintptr_t)S.pixels % 16 == 0)
                          && ((uintptr_t)D.pixels % 16 == 0) && S.ystride % 16 == 0
                          && D.ystride % 16 == 0;
        if (rgba) {
            dim3 grid((D.width + 255) / 256, std::min(D.height, 32768), 1);
            launch_pdl(box2_rgba_kernel, grid, dim3(256), 0, st,
                (const float4*)S.pixels, S.ystride / 16, (float4*)D.pixels,
                D.ystride / 16, D.width, D.height);
        } else {
            long long n = (long long)D.width * D.nchannels * D.height;
            launch_pdl(box2_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st,
                (const float*)S.pixels, S.ystride / 4, (float*)D.pixels,
                D.ystride / 4, D.width, D.height, D.nchannels);
        }
    } else {
        long long n = (long long)D.width * D.height;
        launch_pdl(box_bilinear_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, p);
    }
}

// ------------------------------------------------------------------------
// rangecompress / rangeexpand (imagebufalgo_pixelmath.cpp:560-735): the
// highlight-compensation point ops oiiotool and maketx wrap around a resize.
// ------------------------------------------------------------------------
namespace {

template<bool EXPAND>
__device__ __forceinline__ float
dev_range(float v)
{
    return EXPAND ? dev_rangeexpand(v) : dev_rangecompress(v);
}

// one thread per pixel of the ROI, any of the four types on either side
template<bool EXPAND>
__global__ void __launch_bounds__(256)
range_kernel(const RangeParams p)
{
    pdl_prologue();
    const int w = p.roi_x1 - p.roi_x0;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= w)
        return;
    const int px = p.roi_x0 + x, py = p.roi_y0 + y;
    unsigned char* dp = (unsigned char*)p.dst.pixels + (long long)(py - p.dst.y) * p.dst.ystride
                        + (long long)(px - p.dst.x) * p.dst.xstride;
    const bool inside = px >= p.src.x && px < p.src.x + p.src.width && py >= p.src.y
                        && py < p.src.y + p.src.height;
    const unsigned char* sp = (const unsigned char*)p.src.pixels
                              + (long long)(py - p.src.y) * p.src.ystride
                              + (long long)(px - p.src.x) * p.src.xstride;
    const int ses = p.src.basetype == BT_UINT8 ? 1 : (p.src.basetype == BT_FLOAT ? 4 : 2);
    const int des = p.dst.basetype == BT_UINT8 ? 1 : (p.dst.basetype == BT_FLOAT ? 4 : 2);
    float scale = 0.0f;
    if (p.useluma) {
        float r = inside ? ld_elem(sp + (p.chbegin) * ses, p.src.basetype) : 0.0f;
        float g = inside ? ld_elem(sp + (p.chbegin + 1) * ses, p.src.basetype) : 0.0f;
        float b = inside ? ld_elem(sp + (p.chbegin + 2) * ses, p.src.basetype) : 0.0f;
        float luma = __fadd_rn(__fadd_rn(__fmul_rn(0.21264f, r), __fmul_rn(0.71517f, g)),
                               __fmul_rn(0.07219f, b));
        scale = luma > 0.0f ? __fdiv_rn(dev_range<EXPAND>(luma), luma) : 0.0f;
    }
    for (int c = p.chbegin; c < p.chend; ++c) {
        float v = inside ? ld_elem(sp + c * ses, p.src.basetype) : 0.0f;
        if (c == p.alpha_channel || c == p.z_channel) {
            if (p.in_place)
                continue;
        } else if (p.useluma) {
            v = __fmul_rn(v, scale);
        } else {
            v = dev_range<EXPAND>(v);
        }
        st_elem(dp + c * des, p.dst.basetype, v);
    }
}

// contiguous float rows, no luma: one thread per 4 elements
template<bool EXPAND>
__global__ void __launch_bounds__(256)
range_f32_vec4_kernel(float* dst, long long dst_ystride, const float* src,
                      long long src_ystride, int row_vec4, int nch, int alpha, int z)
{
    pdl_prologue();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= row_vec4)
        return;
    const float4 v = reinterpret_cast<const float4*>(
        (const unsigned char*)src + blockIdx.y * src_ystride)[i];
    float o[4] = { v.x, v.y, v.z, v.w };
    int c = (4 * i) % nch;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (c != alpha && c != z)
            o[k] = dev_range<EXPAND>(o[k]);
        c = (c + 1 == nch) ? 0 : c + 1;
    }
    reinterpret_cast<float4*>((unsigned char*)dst + blockIdx.y * dst_ystride)[i]
        = make_float4(o[0], o[1], o[2], o[3]);
}

}  // namespace

void
launch_range(const RangeParams& p, void* stream)
{
    const int w = p.roi_x1 - p.roi_x0, h = p.roi_y1 - p.roi_y0;
    if (w <= 0 || h <= 0 || p.chend <= p.chbegin)
        return;
    cudaStream_t st = (cudaStream_t)stream;
    // fast path: whole-pixel float rows, every channel, both sides 16-byte aligned
    const int nch = p.dst.nchannels;
    const bool vec = !p.useluma && p.dst.basetype == BT_FLOAT && p.src.basetype == BT_FLOAT
                     && p.src.nchannels == nch && p.chbegin == 0 && p.chend == nch
                     && p.dst.xstride == 4LL * nch && p.src.xstride == 4LL * nch
                     && (w * nch) % 4 == 0 && p.roi_x0 >= p.src.x
                     && p.roi_x1 <= p.src.x + p.src.width && p.roi_y0 >= p.src.y
                     && p.roi_y1 <= p.src.y + p.src.height && h <= 65535;
    if (vec) {
        float* d = (float*)((unsigned char*)p.dst.pixels
                            + (long long)(p.roi_y0 - p.dst.y) * p.dst.ystride
                            + (long long)(p.roi_x0 - p.dst.x) * p.dst.xstride);
        const float* s = (const float*)((const unsigned char*)p.src.pixels
                                        + (long long)(p.roi_y0 - p.src.y) * p.src.ystride
                                        + (long long)(p.roi_x0 - p.src.x) * p.src.xstride);
        if (((uintptr_t)d | (uintptr_t)s | (uintptr_t)p.dst.ystride | (uintptr_t)p.src.ystride)
                % 16
            == 0) {
            const int n4 = w * nch / 4;
            dim3 grid((n4 + 255) / 256, h);
            if (p.expand)
                launch_pdl(range_f32_vec4_kernel<true>, grid, dim3(256), 0, st,
                    d, p.dst.ystride, s, p.src.ystride, n4, nch, p.alpha_channel, p.z_channel);
            else
                launch_pdl(range_f32_vec4_kernel<false>, grid, dim3(256), 0, st,
                    d, p.dst.ystride, s, p.src.ystride, n4, nch, p.alpha_channel, p.z_channel);
            return;
        }
    }
    for (int y0 = 0; y0 < h; y0 += 65535) {
        RangeParams q = p;
        q.roi_y0      = p.roi_y0 + y0;
        q.roi_y1      = std::min(p.roi_y1, q.roi_y0 + 65535);
        dim3 grid((w + 255) / 256, q.roi_y1 - q.roi_y0);
        if (p.expand)
            launch_pdl(range_kernel<true>, grid, dim3(256), 0, st, q);
        else
            launch_pdl(range_kernel<false>, grid, dim3(256), 0, st, q);
    }
}

void
launch_clamp(float* pixels, long long n, int nch, int alpha, float lo, float hi,
             void* stream)
{
    if (n <= 0)
        return;
    launch_pdl(clamp_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream,
        pixels, n, nch, alpha, lo, hi);
}

// ------------------------------------------------------------------------
// Channel-group copy: `nbytes` bytes of every pixel from src to dst (each
// with its own pixel / row strides).  Splits an N-channel image into <= 4
// channel groups for the fused kernel and puts the results back.
// ------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256)
pixel_bytes_copy_kernel(unsigned char* dst, long long dxs, long long dys,
                        const unsigned char* src, long long sxs, long long sys, int w, int h,
                        int nbytes)
{
    pdl_prologue();
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= w)
        return;
    for (int y = blockIdx.y; y < h; y += gridDim.y) {
        unsigned char* d       = dst + y * dys + x * dxs;
        const unsigned char* s = src + y * sys + x * sxs;
        if (((nbytes | (uintptr_t)d | (uintptr_t)s) & 3) == 0) {
            for (int i = 0; i < nbytes; i += 4)
                *(unsigned*)(d + i) = __ldg((const unsigned*)(s + i));
        } else if (((nbytes | (uintptr_t)d | (uintptr_t)s) & 1) == 0) {
            for (int i = 0; i < nbytes; i += 2)
                *(unsigned short*)(d + i) = __ldg((const unsigned short*)(s + i));
        } else {
            for (int i = 0; i < nbytes; ++i)
                d[i] = __ldg(s + i);
        }
    }
}
}  // namespace

void
launch_pixel_bytes_copy(void* dst, long long dxs, long long dys, const void* src, long long sxs,
                        long long sys, int w, int h, int nbytes, void* stream)
{
    if (w <= 0 || h <= 0 || nbytes <= 0)
        return;
    dim3 grid((w + 255) / 256, std::min(h, 32768), 1);
    launch_pdl(pixel_bytes_copy_kernel, grid, dim3(256), 0, (cudaStream_t)stream,
               (unsigned char*)dst, dxs, dys, (const unsigned char*)src, sxs, sys, w, h, nbytes);
}

}  // namespace b2r
