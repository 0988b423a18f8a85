// kernels.cu -- sm_100a kernels of the filtered resampler.
//
//  resize_exact_kernel   single pass, the reference's own tap order and
//                        zero-weight skips (imagebufalgo_xform.cpp:732-823);
//                        bit-exact against the CPU algorithm.  Used for
//                        non-separable filters, overscanned sources, channel
//                        mismatches, very wide footprints, and as the re-do
//                        path when the fused kernel meets non-finite pixels.
//  resize_fused_kernel   the hot path: vertical pass straight from HBM into
//                        registers (each source row of a strip is loaded
//                        once, 128-bit, coalesced), V-filtered rows staged in
//                        shared memory, horizontal pass from shared memory
//                        with per-thread register weights, quantise + store.
//  resample_kernel       nearest / bilinear (xform.cpp:1107-1186,1379-1493).
//  box_* kernels         maketx resize_block (maketexture.cpp:141-334).
//
// No tensor cores: this is a banded gather-FMA bounded by HBM bandwidth.
#include "kernels.h"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>

namespace b2r {

// ------------------------------------------------------------------------
// type conversion: convert_type<S,float> / convert_type<float,D>
// (src/include/OpenImageIO/fmath.h:753-764, 1009-1031).  Explicit _rn
// intrinsics keep nvcc from contracting v*max+0.5 into one FMA, which would
// round differently from the reference's multiply-then-add.
// ------------------------------------------------------------------------
__device__ __forceinline__ float
load_elem(const unsigned char* p, int bt)
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
quant_int(float v, float mx)
{
    float s = __fmul_rn(v, mx);
    s       = __fadd_rn(s, (s < 0.0f) ? -0.5f : 0.5f);
    if (!(0.0f <= s))  // also catches NaN -> 0
        s = 0.0f;
    if (s > mx)
        s = mx;
    return (unsigned int)s;  // truncation
}

__device__ __forceinline__ void
store_elem(unsigned char* p, int bt, float v)
{
    switch (bt) {
    case BT_UINT8: *p = (unsigned char)quant_int(v, 255.0f); break;
    case BT_UINT16: *(unsigned short*)p = (unsigned short)quant_int(v, 65535.0f); break;
    case BT_HALF: *(__half*)p = __float2half_rn(v); break;
    default: *(float*)p = v; break;
    }
}

__device__ __forceinline__ bool
nonfinite(float v)
{
    return (__float_as_uint(v) & 0x7f800000u) == 0x7f800000u;
}

// ------------------------------------------------------------------------
// exact kernel
// ------------------------------------------------------------------------
__device__ __forceinline__ const unsigned char*
src_pixel_wrapclamp(const DevImage& im, int x, int y)
{
    // ConstIterator(src, WrapClamp): imagebuf.cpp:3623-3667 + do_wrap :3283-3322
    bool inside = x >= im.x && x < im.x + im.width && y >= im.y
                  && y < im.y + im.height;
    if (!inside) {
        x      = min(max(x, im.full_x), im.full_x + im.full_width - 1);
        y      = min(max(y, im.full_y), im.full_y + im.full_height - 1);
        inside = x >= im.x && x < im.x + im.width && y >= im.y
                 && y < im.y + im.height;
        if (!inside)
            return nullptr;
    }
    return (const unsigned char*)im.pixels + (long long)(y - im.y) * im.ystride
           + (long long)(x - im.x) * im.xstride;
}

__device__ __forceinline__ float
dev_lanczos3(float x)
{
    const float a    = 3.0f;
    const float ainv = 1.0f / a;
    const float pi   = 3.14159265358979323846f;
    x                = fabsf(x);
    if (x > a)
        return 0.0f;
    if (x < 0.0001f)
        return 1.0f;
    float s1 = sinf(__fmul_rn(__fmul_rn(x, ainv), pi));
    float s3 = __fmul_rn(__fadd_rn(__fmul_rn(__fmul_rn(-4.0f, s1), s1), 3.0f), s1);
    float den = __fmul_rn(__fmul_rn(x, x), __fmul_rn(pi, pi));
    return __fmul_rn(__fmul_rn(__fdiv_rn(a, den), s1), s3);
}

__device__ __forceinline__ float
dev_nonsep_filter(int kind, float fw, float fh, float x, float y)
{
    if (kind == 1) {  // disk, filter.cpp:689-694
        x = __fdiv_rn(x, __fmul_rn(fw, 0.5f));
        y = __fdiv_rn(y, __fmul_rn(fh, 0.5f));
        return (__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)) < 1.0f) ? 1.0f : 0.0f;
    }
    // radial-lanczos3, filter.cpp:516-521
    x = __fmul_rn(x, __fdiv_rn(6.0f, fw));
    y = __fmul_rn(y, __fdiv_rn(6.0f, fh));
    return dev_lanczos3(__fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))));
}

constexpr int EXACT_CH = 4;  // channels per thread

__global__ void __launch_bounds__(128)
resize_exact_kernel(ExactParams p)
{
    if (p.gate && *p.gate == 0)
        return;
    const int roi_w = p.roi_x1 - p.roi_x0;
    const int xblocks = (roi_w + blockDim.x - 1) / blockDim.x;
    const int kx      = (blockIdx.x % xblocks) * blockDim.x + threadIdx.x;
    const int ky      = blockIdx.x / xblocks;
    const int c0    = blockIdx.z * EXACT_CH;
    if (kx >= roi_w)
        return;
    const int nc  = min(EXACT_CH, p.nch - c0);
    const int ses = (p.src.basetype == BT_UINT8) ? 1 : (p.src.basetype == BT_FLOAT ? 4 : 2);
    const int des = (p.dst.basetype == BT_UINT8) ? 1 : (p.dst.basetype == BT_FLOAT ? 4 : 2);

    float pel[EXACT_CH] = { 0.0f, 0.0f, 0.0f, 0.0f };
    bool zero_out       = false;

    if (p.nonsep_kind == 0) {
        const int src_x   = p.xcenter[kx];
        const int src_y   = p.ycenter[ky];
        const float* xw   = p.xw + (long long)kx * p.xtaps;
        const float* yw   = p.yw + (long long)ky * p.ytaps;
        if (p.xwsum[kx] != 0.0f) {
            for (int j = 0; j < p.ytaps; ++j) {
                float wy = yw[j];
                if (wy == 0.0f)
                    continue;
                for (int i = 0; i < p.xtaps; ++i) {
                    float w = __fmul_rn(wy, xw[i]);
                    if (w != 0.0f) {
                        const unsigned char* sp
                            = src_pixel_wrapclamp(p.src, src_x - p.radi + i,
                                                  src_y - p.radj + j);
                        if (sp) {
                            sp += c0 * ses;
#pragma unroll
                            for (int c = 0; c < EXACT_CH; ++c)
                                if (c < nc)
                                    pel[c] = __fadd_rn(
                                        pel[c],
                                        __fmul_rn(w, load_elem(sp + c * ses,
                                                               p.src.basetype)));
                        }
                    }
                }
            }
        }
        zero_out = (p.ywsum[ky] == 0.0f);
    } else {
        const int src_x     = p.xcenter[kx];
        const int src_y     = p.ycenter[ky];
        const float fx      = p.xfrac[kx];
        const float fy      = p.yfrac[ky];
        float totalweight   = 0.0f;
        int n               = 0;
        for (int j = -p.radj; j <= p.radj; ++j) {
            for (int i = -p.radi; i <= p.radi; ++i, ++n) {
                float ax = __fmul_rn(p.xratio, __fsub_rn(float(i), __fsub_rn(fx, 0.5f)));
                float ay = __fmul_rn(p.yratio, __fsub_rn(float(j), __fsub_rn(fy, 0.5f)));
                float w  = dev_nonsep_filter(p.nonsep_kind, p.filt_w, p.filt_h, ax, ay);
                if (w != 0.0f) {
                    totalweight = __fadd_rn(totalweight, w);
                    // the reference iterator's window is xtaps wide and
                    // starts at (src_x-radi, src_y-radi) (xform.cpp:793-794)
                    int px = src_x - p.radi + (n % p.xtaps);
                    int py = src_y - p.radi + (n / p.xtaps);
                    const unsigned char* sp = src_pixel_wrapclamp(p.src, px, py);
                    if (sp) {
                        sp += c0 * ses;
#pragma unroll
                        for (int c = 0; c < EXACT_CH; ++c)
                            if (c < nc)
                                pel[c] = __fadd_rn(
                                    pel[c],
                                    __fmul_rn(w, load_elem(sp + c * ses,
                                                           p.src.basetype)));
                    }
                }
            }
        }
        if (totalweight == 0.0f) {
            zero_out = true;
        } else {
#pragma unroll
            for (int c = 0; c < EXACT_CH; ++c)
                pel[c] = __fdiv_rn(pel[c], totalweight);
        }
    }

    unsigned char* dp = (unsigned char*)p.dst.pixels
                        + (long long)(p.roi_y0 + ky - p.dst.y) * p.dst.ystride
                        + (long long)(p.roi_x0 + kx - p.dst.x) * p.dst.xstride
                        + c0 * des;
#pragma unroll
    for (int c = 0; c < EXACT_CH; ++c)
        if (c < nc)
            store_elem(dp + c * des, p.dst.basetype, zero_out ? 0.0f : pel[c]);
}

void
launch_resize_exact(const ExactParams& p, void* stream)
{
    int roi_w = p.roi_x1 - p.roi_x0, roi_h = p.roi_y1 - p.roi_y0;
    if (roi_w <= 0 || roi_h <= 0)
        return;
    dim3 block(128, 1, 1);
    dim3 grid(((roi_w + 127) / 128) * roi_h, 1, (p.nch + EXACT_CH - 1) / EXACT_CH);
    resize_exact_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(p);
}


// ------------------------------------------------------------------------
// fused two-pass kernel
// ------------------------------------------------------------------------
// Raw 4-element group as loaded from global memory (not yet converted):
// float -> 4 words, half/uint16 -> 2 words, uint8 -> 1 word.
__device__ __forceinline__ uint4
ldg_stream_v4(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2
ldg_stream_v2(const void* p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];"
                 : "=r"(r.x), "=r"(r.y)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned int
ldg_stream_u32(const void* p)
{
    unsigned int r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

// Load elements [e, e+4) of a source row.  STREAM: bypass L1 (each row is
// read once per strip); otherwise allocate in L1 (gather mode re-reads rows).
template<bool STREAM>
__device__ __forceinline__ uint4
load_raw4(const unsigned char* row, int e, int row_elems, int bt, bool vec_ok)
{
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (vec_ok && e + 4 <= row_elems) {
        if (bt == BT_FLOAT) {
            const void* q = row + (size_t)e * 4;
            r = STREAM ? ldg_stream_v4(q) : __ldg((const uint4*)q);
        } else if (bt == BT_UINT8) {
            const void* q = row + e;
            r.x = STREAM ? ldg_stream_u32(q) : __ldg((const unsigned int*)q);
        } else {
            const void* q = row + (size_t)e * 2;
            uint2 t = STREAM ? ldg_stream_v2(q) : __ldg((const uint2*)q);
            r.x     = t.x;
            r.y     = t.y;
        }
    } else {
        unsigned int w[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (e + k < row_elems) {
                if (bt == BT_FLOAT)
                    w[k] = __ldg((const unsigned int*)(row + (size_t)(e + k) * 4));
                else if (bt == BT_UINT8)
                    w[k] = __ldg(row + e + k);
                else
                    w[k] = __ldg((const unsigned short*)(row + (size_t)(e + k) * 2));
            }
        }
        if (bt == BT_FLOAT)
            r = make_uint4(w[0], w[1], w[2], w[3]);
        else if (bt == BT_UINT8)
            r.x = w[0] | (w[1] << 8) | (w[2] << 16) | (w[3] << 24);
        else {
            r.x = w[0] | (w[1] << 16);
            r.y = w[2] | (w[3] << 16);
        }
    }
    return r;
}

__device__ __forceinline__ float4
convert_raw4(uint4 r, int bt)
{
    float4 v;
    if (bt == BT_FLOAT) {
        v = make_float4(__uint_as_float(r.x), __uint_as_float(r.y),
                        __uint_as_float(r.z), __uint_as_float(r.w));
    } else if (bt == BT_HALF) {
        __half2 a = *reinterpret_cast<__half2*>(&r.x);
        __half2 b = *reinterpret_cast<__half2*>(&r.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        v = make_float4(fa.x, fa.y, fb.x, fb.y);
    } else if (bt == BT_UINT16) {
        const float s = 1.0f / 65535.0f;
        v = make_float4(__fmul_rn(float(r.x & 0xffffu), s),
                        __fmul_rn(float(r.x >> 16), s),
                        __fmul_rn(float(r.y & 0xffffu), s),
                        __fmul_rn(float(r.y >> 16), s));
    } else {
        const float s = 1.0f / 255.0f;
        v = make_float4(__fmul_rn(float(r.x & 0xffu), s),
                        __fmul_rn(float((r.x >> 8) & 0xffu), s),
                        __fmul_rn(float((r.x >> 16) & 0xffu), s),
                        __fmul_rn(float(r.x >> 24), s));
    }
    return v;
}

// Rotate each block of 8 float4 slots by its block index so that the
// stride-2 (2:1 downsample) and stride-1 access patterns of the H pass are
// both free of bank conflicts.
__device__ __forceinline__ int
swz(int q)
{
    return (q & ~7) | ((q + (q >> 3)) & 7);
}

__device__ __forceinline__ void
fma4(float4& a, float w, const float4& v)
{
    a.x = fmaf(w, v.x, a.x);
    a.y = fmaf(w, v.y, a.y);
    a.z = fmaf(w, v.z, a.z);
    a.w = fmaf(w, v.w, a.w);
}

constexpr int FUSED_PREFETCH = 4;  // source rows in flight per thread

// H pass over `nb` buffered rows; dst rows dyb .. dyb+nb-1 of the ROI.
template<int NCH, int HT>
__device__ __forceinline__ void
fused_hpass(const FusedParams& p, const float* vrows, int pitch_f, int nb,
            int dyb, int dxg, bool col_active, const int (&hoff)[HT],
            const float (&hwt)[HT], int hr)
{
    if (!col_active)
        return;
    const int des = (p.dsttype == BT_UINT8) ? 1 : (p.dsttype == BT_FLOAT ? 4 : 2);
    for (int b = hr; b < nb; b += FUSED_THREADS / FUSED_HCOLS) {
        const float* row = vrows + (size_t)b * pitch_f;
        unsigned char* dp = p.dst + (long long)(dyb + b) * p.dst_ystride
                            + (long long)dxg * (NCH * des);
        if (NCH == 4) {
            const float4* row4 = reinterpret_cast<const float4*>(row);
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int i = 0; i < HT; ++i)
                fma4(a, hwt[i], row4[hoff[i]]);
            if (p.check_nonfinite
                && (nonfinite(a.x) || nonfinite(a.y) || nonfinite(a.z)
                    || nonfinite(a.w)))
                *p.nonfinite_flag = 1;
            if (p.dst_vec_ok) {
                if (p.dsttype == BT_FLOAT) {
                    *reinterpret_cast<float4*>(dp) = a;
                } else if (p.dsttype == BT_HALF) {
                    __half2 h0 = __floats2half2_rn(a.x, a.y);
                    __half2 h1 = __floats2half2_rn(a.z, a.w);
                    uint2 u;
                    u.x = *reinterpret_cast<unsigned int*>(&h0);
                    u.y = *reinterpret_cast<unsigned int*>(&h1);
                    *reinterpret_cast<uint2*>(dp) = u;
                } else if (p.dsttype == BT_UINT16) {
                    uint2 u;
                    u.x = quant_int(a.x, 65535.0f) | (quant_int(a.y, 65535.0f) << 16);
                    u.y = quant_int(a.z, 65535.0f) | (quant_int(a.w, 65535.0f) << 16);
                    *reinterpret_cast<uint2*>(dp) = u;
                } else {
                    unsigned int u = quant_int(a.x, 255.0f)
                                     | (quant_int(a.y, 255.0f) << 8)
                                     | (quant_int(a.z, 255.0f) << 16)
                                     | (quant_int(a.w, 255.0f) << 24);
                    *reinterpret_cast<unsigned int*>(dp) = u;
                }
            } else {
                store_elem(dp, p.dsttype, a.x);
                store_elem(dp + des, p.dsttype, a.y);
                store_elem(dp + 2 * des, p.dsttype, a.z);
                store_elem(dp + 3 * des, p.dsttype, a.w);
            }
        } else {
            float a[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                a[c] = 0.0f;
#pragma unroll
            for (int i = 0; i < HT; ++i) {
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                    a[c] = fmaf(hwt[i], row[hoff[i] + c], a[c]);
            }
            bool bad = false;
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                bad |= nonfinite(a[c]);
                store_elem(dp + c * des, p.dsttype, a[c]);
            }
            if (p.check_nonfinite && bad)
                *p.nonfinite_flag = 1;
        }
    }
}

template<int NCH, int VK, int HT>
__global__ void __launch_bounds__(FUSED_THREADS, 2)
resize_fused_kernel(const FusedParams p, int pitch_f)
{
    extern __shared__ __align__(16) float smem[];
    float* vrows  = smem;                                   // FUSED_BATCH * pitch_f
    float* ringw  = smem + (size_t)FUSED_BATCH * pitch_f;   // band rows * VK

    const int tid     = threadIdx.x;
    const Strip strip = p.strips[blockIdx.x];
    const Band band   = p.bands[blockIdx.y];

    // ---- H-pass per-thread constants: this thread's dst column ----------
    const int hc          = tid % FUSED_HCOLS;
    const int hr          = tid / FUSED_HCOLS;
    const bool col_active = hc < strip.ndx;
    const int dxg         = strip.dx0 + hc;
    int hoff[HT];
    float hwt[HT];
    {
        int f = col_active ? p.hfirst[dxg] : 0;
#pragma unroll
        for (int i = 0; i < HT; ++i) {
            hwt[i] = col_active ? p.hw[(size_t)dxg * HT + i] : 0.0f;
            if (NCH == 4) {
                int q   = f + i - strip.e0 / 4;
                hoff[i] = col_active ? swz(q) : 0;
            } else {
                hoff[i] = col_active ? (f + i) * NCH - strip.e0 : 0;
            }
        }
    }

    // ---- V-pass per-thread constants: this thread's 4 source elements ---
    const bool vactive   = tid < strip.nvec;
    const int e          = strip.e0 + 4 * tid;
    const int row_elems  = p.src_w * p.nch;
    const bool vec_ok    = p.src_vec_ok != 0;
    const int vslot      = (NCH == 4) ? swz(tid) * 4 : tid * 4;  // float index
    const int dy_end     = band.dy0 + band.ndy;
    int dy               = band.dy0;
    int nb               = 0;

    if (VK > 0) {
        // stage this band's ring weights
        const float* gw = p.vring + (size_t)band.woff * VK;
        for (int i = tid; i < band.nrows * VK; i += FUSED_THREADS)
            ringw[i] = gw[i];
        __syncthreads();

        float4 acc[VK > 0 ? VK : 1];
#pragma unroll
        for (int k = 0; k < VK; ++k)
            acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

        const int rend = band.r0 + band.nrows;
        int slot       = 0;
        int nextlast   = p.vlast[dy];
        const unsigned char* rowp = p.src + (long long)band.r0 * p.src_ystride;

        uint4 cur[FUSED_PREFETCH];
#pragma unroll
        for (int q = 0; q < FUSED_PREFETCH; ++q) {
            cur[q] = make_uint4(0u, 0u, 0u, 0u);
            if (vactive && band.r0 + q < rend)
                cur[q] = load_raw4<true>(rowp + (long long)q * p.src_ystride, e,
                                         row_elems, p.srctype, vec_ok);
        }
        for (int r = band.r0; r < rend; r += FUSED_PREFETCH) {
            uint4 nxt[FUSED_PREFETCH];
#pragma unroll
            for (int q = 0; q < FUSED_PREFETCH; ++q) {
                nxt[q] = make_uint4(0u, 0u, 0u, 0u);
                if (vactive && r + FUSED_PREFETCH + q < rend)
                    nxt[q] = load_raw4<true>(
                        rowp + (long long)(FUSED_PREFETCH + q) * p.src_ystride, e,
                        row_elems, p.srctype, vec_ok);
            }
            rowp += (long long)FUSED_PREFETCH * p.src_ystride;
#pragma unroll
            for (int q = 0; q < FUSED_PREFETCH; ++q) {
                const int rr = r + q;
                if (rr < rend) {
                    const float4 v = convert_raw4(cur[q], p.srctype);
                    const float* w = ringw + (size_t)(rr - band.r0) * VK;
#pragma unroll
                    for (int k = 0; k < VK; ++k)
                        fma4(acc[k], w[k], v);
                    while (dy < dy_end && nextlast == rr) {
                        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
                        switch (slot) {
#define B2R_EMIT_CASE(K_)                                 \
    case K_:                                              \
        if (K_ < VK) {                                    \
            o       = acc[K_ < VK ? K_ : 0];              \
            acc[K_ < VK ? K_ : 0] = make_float4(0.f, 0.f, 0.f, 0.f); \
        }                                                 \
        break;
                            B2R_EMIT_CASE(0)
                            B2R_EMIT_CASE(1)
                            B2R_EMIT_CASE(2)
                            B2R_EMIT_CASE(3)
                            B2R_EMIT_CASE(4)
                            B2R_EMIT_CASE(5)
                            B2R_EMIT_CASE(6)
                            B2R_EMIT_CASE(7)
#undef B2R_EMIT_CASE
                        default: break;
                        }
                        slot = (slot + 1 == VK) ? 0 : slot + 1;
                        if (vactive)
                            *reinterpret_cast<float4*>(vrows + (size_t)nb * pitch_f
                                                       + vslot)
                                = o;
                        ++nb;
                        ++dy;
                        nextlast = (dy < dy_end) ? p.vlast[dy] : -1;
                        if (nb == FUSED_BATCH || dy == dy_end) {
                            __syncthreads();
                            fused_hpass<NCH, HT>(p, vrows, pitch_f, nb, dy - nb, dxg,
                                                 col_active, hoff, hwt, hr);
                            __syncthreads();
                            nb = 0;
                        }
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < FUSED_PREFETCH; ++q)
                cur[q] = nxt[q];
        }
    } else {
        // gather mode: each dst row reads its own vt source rows (upsampling:
        // consecutive dst rows share rows, which stay in L1/L2)
        for (; dy < dy_end; ++dy) {
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (vactive) {
                const int first = p.vfirst[dy];
                const float* w  = p.vw + (size_t)dy * p.vt;
                const unsigned char* rowp
                    = p.src + (long long)first * p.src_ystride;
                for (int j = 0; j < p.vt; ++j) {
                    uint4 raw = load_raw4<false>(rowp, e, row_elems, p.srctype,
                                                 vec_ok);
                    fma4(a, __ldg(w + j), convert_raw4(raw, p.srctype));
                    rowp += p.src_ystride;
                }
                *reinterpret_cast<float4*>(vrows + (size_t)nb * pitch_f + vslot) = a;
            }
            ++nb;
            if (nb == FUSED_BATCH || dy + 1 == dy_end) {
                __syncthreads();
                fused_hpass<NCH, HT>(p, vrows, pitch_f, nb, dy + 1 - nb, dxg,
                                     col_active, hoff, hwt, hr);
                __syncthreads();
                nb = 0;
            }
        }
    }
}

typedef void (*FusedKernel)(const FusedParams, int);

template<int NCH, int VK>
static FusedKernel
pick_ht(int ht)
{
    switch (ht) {
    case 4: return resize_fused_kernel<NCH, VK, 4>;
    case 8: return resize_fused_kernel<NCH, VK, 8>;
    case 12: return resize_fused_kernel<NCH, VK, 12>;
    case 16: return resize_fused_kernel<NCH, VK, 16>;
    }
    return nullptr;
}
template<int NCH>
static FusedKernel
pick_vk(int vk, int ht)
{
    switch (vk) {
    case 0: return pick_ht<NCH, 0>(ht);
    case 2: return pick_ht<NCH, 2>(ht);
    case 4: return pick_ht<NCH, 4>(ht);
    case 6: return pick_ht<NCH, 6>(ht);
    case 8: return pick_ht<NCH, 8>(ht);
    }
    return nullptr;
}
static FusedKernel
pick_fused(int nch, int vk, int ht)
{
    switch (nch) {
    case 1: return pick_vk<1>(vk, ht);
    case 2: return pick_vk<2>(vk, ht);
    case 3: return pick_vk<3>(vk, ht);
    case 4: return pick_vk<4>(vk, ht);
    }
    return nullptr;
}

bool
fused_supported(int nch, int vk, int ht)
{
    return pick_fused(nch, vk, ht) != nullptr;
}

static int
fused_pitch(int max_nvec)
{
    return ((max_nvec + 7) / 8) * 8 * 4;  // floats; multiple of 8 float4 slots
}

size_t
fused_smem_bytes(const FusedParams& p, int max_band_rows, int max_nvec)
{
    size_t f = (size_t)FUSED_BATCH * fused_pitch(max_nvec);
    if (p.vk > 0)
        f += (size_t)max_band_rows * p.vk;
    return f * sizeof(float);
}

void
launch_resize_fused(const FusedParams& p, int max_band_rows, int max_nvec,
                    void* stream)
{
    FusedKernel k = pick_fused(p.nch, p.vk, p.ht);
    if (!k || p.nstrips <= 0 || p.nbands <= 0)
        return;
    size_t smem = fused_smem_bytes(p, max_band_rows, max_nvec);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute((const void*)k,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    dim3 grid(p.nstrips, p.nbands, 1);
    k<<<grid, FUSED_THREADS, smem, (cudaStream_t)stream>>>(p, fused_pitch(max_nvec));
}

}  // namespace b2r
