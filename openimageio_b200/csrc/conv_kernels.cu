// conv_kernels.cu -- ImageBufAlgo::convolve and the combine step of
// unsharp_mask on the device (src/libOpenImageIO/imagebufalgo.cpp:772-812,
// 961-985): what maketx's --sharpen wraps around every MIP level's resize
// (maketexture.cpp:909-931) and oiiotool --convolve/--blur/--unsharp run.
//
// convolve: one thread per destination pixel, channels in passes of four,
// taps in the reference's order (kernel rows outer, columns inner) with
// multiply-then-add -- no contraction -- so results equal the CPU's to the bit
// (the kernel table is built on the host).  Source pixels are wrapped like a
// WrapClamp iterator: clamp to the display window, black if that still misses
// the data window.  Neighbouring threads share almost all of their taps, so
// the source goes through L1 (ld.global.nc); a 3x3 or 5x5 blur is HBM-bound
// (one read + one write per pixel), 15x15 is FP32-bound.
#include "dev_convert.cuh"

#include <algorithm>

namespace b2r {
namespace {

constexpr int CONV_BX = 32, CONV_BY = 8;

// One source pixel's (up to) four channels as float.
template<bool VEC_F32>
__device__ __forceinline__ void
conv_load(const unsigned char* sp, const DevImage& S, int c0, int ses, float v[4])
{
    if (VEC_F32) {
        const float4 q = __ldg((const float4*)sp);
        v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            v[c] = (c0 + c < S.nchannels) ? load_elem(sp + c * ses, S.basetype) : 0.0f;
    }
}

// Store one output pixel's channel pass: the blur itself, or -- unsharp -- the
// combine step fused behind it (unsharp_impl, :961-985: the blurred value is
// what convolve would have stored in the float `Blurry` image, so the result
// is the same to the bit).
__device__ __forceinline__ void
conv_store(const ConvParams& p, int x, int y, int c0, int ses, int des, const float (&sum)[4])
{
    unsigned char* dp = (unsigned char*)p.dst.pixels + (long long)(y - p.dst.y) * p.dst.ystride
                        + (long long)(x - p.dst.x) * p.dst.xstride;
    const DevImage& S = p.src;
    if (p.unsharp) {
        const bool in = x >= S.x && x < S.x + S.width && y >= S.y && y < S.y + S.height;
        const unsigned char* cp = (const unsigned char*)S.pixels
                                  + (long long)(y - S.y) * S.ystride
                                  + (long long)(x - S.x) * S.xstride;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int ch = c0 + c;
            if (ch >= p.chbegin && ch < p.chend) {
                const float sv   = in ? load_elem(cp + ch * ses, S.basetype) : 0.0f;
                const float b    = in ? __fmul_rn(p.scale, sum[c]) : 0.0f;
                const float diff = __fsub_rn(sv, b);
                float v          = sv;
                if (fabsf(diff) > p.threshold)
                    v = __fadd_rn(sv, __fmul_rn(p.contrast, diff));
                store_elem(dp + ch * des, p.dst.basetype, v);
            }
        }
        return;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const int ch = c0 + c;
        if (ch >= p.chbegin && ch < p.chend)
            store_elem(dp + ch * des, p.dst.basetype, __fmul_rn(p.scale, sum[c]));
    }
}

// One output pixel, any position: taps in the reference's order, WrapClamp.
template<bool VEC_F32>
__device__ __forceinline__ void
conv_pixel_generic(const ConvParams& p, int x, int y, int c0, int kw, int kh, int ses,
                   float (&sum)[4])
{
    const DevImage& S = p.src;
    const int kx0 = -(kw / 2), ky0 = -(kh / 2);
    const int fx0 = S.full_x, fx1 = S.full_x + S.full_width - 1;
    const int fy0 = S.full_y, fy1 = S.full_y + S.full_height - 1;
    const bool interior = x + kx0 >= S.x && x + kx0 + kw <= S.x + S.width && y + ky0 >= S.y
                          && y + ky0 + kh <= S.y + S.height;
    const float* k = p.kernel;
    if (interior) {
        const unsigned char* rowp = (const unsigned char*)S.pixels
                                    + (long long)(y + ky0 - S.y) * S.ystride
                                    + (long long)(x + kx0 - S.x) * S.xstride + c0 * ses;
#pragma unroll 1
        for (int j = 0; j < kh; ++j, rowp += S.ystride) {
            const unsigned char* sp = rowp;
#pragma unroll 1
            for (int i = 0; i < kw; ++i, sp += S.xstride) {
                const float w = __ldg(k++);
                float v[4];
                conv_load<VEC_F32>(sp, S, c0, ses, v);
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    sum[c] = __fadd_rn(sum[c], __fmul_rn(w, v[c]));
            }
        }
        return;
    }
#pragma unroll 1
    for (int j = 0; j < kh; ++j) {
        const int ty = y + ky0 + j;
#pragma unroll 1
        for (int i = 0; i < kw; ++i, ++k) {
            const float w = __ldg(k);
            int px = x + kx0 + i, py = ty;
            bool ok = px >= S.x && px < S.x + S.width && py >= S.y && py < S.y + S.height;
            if (!ok) {
                px = min(max(px, fx0), fx1);
                py = min(max(py, fy0), fy1);
                ok = px >= S.x && px < S.x + S.width && py >= S.y && py < S.y + S.height;
            }
            float v[4] = { 0.f, 0.f, 0.f, 0.f };
            if (ok)
                conv_load<VEC_F32>((const unsigned char*)S.pixels
                                       + (long long)(py - S.y) * S.ystride
                                       + (long long)(px - S.x) * S.xstride + c0 * ses,
                                   S, c0, ses, v);
            // a black tap still adds k * 0 (sign of zero aside: no effect)
#pragma unroll
            for (int c = 0; c < 4; ++c)
                sum[c] = __fadd_rn(sum[c], __fmul_rn(w, v[c]));
        }
    }
}

// KW, KH > 0: compile-time kernel extent -- taps fully unrolled, weights in
// registers, and CONV_RPT vertically adjacent output pixels per thread so that
// each source pixel of the (KH + RPT - 1) x KW window is loaded once and feeds
// every output it belongs to (each output still receives its taps in the
// reference's order: rows ascending, columns ascending).  0: run-time extent,
// one output per thread.
constexpr int CONV_RPT = 4;

template<bool VEC_F32, int KW, int KH>
__global__ void __launch_bounds__(CONV_BX* CONV_BY)
convolve_kernel(const ConvParams p)
{
    pdl_prologue();
    const int x = p.roi_x0 + blockIdx.x * CONV_BX + threadIdx.x;
    if (x >= p.roi_x1)
        return;
    const DevImage& S = p.src;
    const int ses = S.basetype == BT_UINT8 ? 1 : (S.basetype == BT_FLOAT ? 4 : 2);
    const int des = p.dst.basetype == BT_UINT8 ? 1 : (p.dst.basetype == BT_FLOAT ? 4 : 2);
    constexpr int RPT = KW > 0 ? CONV_RPT : 1;
    if constexpr (KW > 0) {
        float wreg[KW > 0 ? KW * KH : 1];
#pragma unroll
        for (int i = 0; i < (KW > 0 ? KW * KH : 1); ++i)
            wreg[i] = __ldg(p.kernel + i);
        const int kx0 = -(KW / 2), ky0 = -(KH / 2);
        const bool x_inside = x + kx0 >= S.x && x + kx0 + KW <= S.x + S.width;
        for (int y0 = p.roi_y0 + (blockIdx.y * CONV_BY + threadIdx.y) * RPT; y0 < p.roi_y1;
             y0 += gridDim.y * CONV_BY * RPT) {
            const int nr = min(RPT, p.roi_y1 - y0);
            const bool interior = x_inside && nr == RPT && y0 + ky0 >= S.y
                                  && y0 + RPT - 1 + ky0 + KH <= S.y + S.height;
            for (int c0 = p.chbegin & ~3; c0 < p.chend; c0 += 4) {
                if (interior) {
                    float sum[RPT][4];
#pragma unroll
                    for (int r = 0; r < RPT; ++r)
                        sum[r][0] = sum[r][1] = sum[r][2] = sum[r][3] = 0.0f;
                    const unsigned char* rowp = (const unsigned char*)S.pixels
                                                + (long long)(y0 + ky0 - S.y) * S.ystride
                                                + (long long)(x + kx0 - S.x) * S.xstride
                                                + c0 * ses;
#pragma unroll
                    for (int j = 0; j < (KW > 0 ? KH + RPT - 1 : 1); ++j, rowp += S.ystride) {
                        const unsigned char* sp = rowp;
#pragma unroll
                        for (int i = 0; i < (KW > 0 ? KW : 1); ++i, sp += S.xstride) {
                            float v[4];
                            conv_load<VEC_F32>(sp, S, c0, ses, v);
#pragma unroll
                            for (int r = 0; r < RPT; ++r) {
                                const int jj = j - r;  // this row's tap row for output r
                                if (jj >= 0 && jj < KH) {
                                    const float w = wreg[KW > 0 ? jj * KW + i : 0];
#pragma unroll
                                    for (int c = 0; c < 4; ++c)
                                        sum[r][c] = __fadd_rn(sum[r][c], __fmul_rn(w, v[c]));
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < RPT; ++r)
                        conv_store(p, x, y0 + r, c0, ses, des, sum[r]);
                } else {
                    for (int r = 0; r < nr; ++r) {
                        float sum[4] = { 0.f, 0.f, 0.f, 0.f };
                        conv_pixel_generic<VEC_F32>(p, x, y0 + r, c0, KW, KH, ses, sum);
                        conv_store(p, x, y0 + r, c0, ses, des, sum);
                    }
                }
            }
        }
    } else {
        for (int y = p.roi_y0 + blockIdx.y * CONV_BY + threadIdx.y; y < p.roi_y1;
             y += gridDim.y * CONV_BY) {
            for (int c0 = p.chbegin & ~3; c0 < p.chend; c0 += 4) {
                float sum[4] = { 0.f, 0.f, 0.f, 0.f };
                conv_pixel_generic<VEC_F32>(p, x, y, c0, p.kw, p.kh, ses, sum);
                conv_store(p, x, y, c0, ses, des, sum);
            }
        }
    }
}

// unsharp_impl: d = s + contrast * (s - b) where |s - b| > threshold, else s
__global__ void __launch_bounds__(256)
unsharp_combine_kernel(const UnsharpParams p)
{
    pdl_prologue();
    const int x = p.roi_x0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= p.roi_x1)
        return;
    const DevImage &S = p.src, &B = p.blur;
    const int ses = S.basetype == BT_UINT8 ? 1 : (S.basetype == BT_FLOAT ? 4 : 2);
    const int des = p.dst.basetype == BT_UINT8 ? 1 : (p.dst.basetype == BT_FLOAT ? 4 : 2);
    for (int y = p.roi_y0 + blockIdx.y; y < p.roi_y1; y += gridDim.y) {
        const bool in = x >= S.x && x < S.x + S.width && y >= S.y && y < S.y + S.height;
        const unsigned char* sp = (const unsigned char*)S.pixels
                                  + (long long)(y - S.y) * S.ystride
                                  + (long long)(x - S.x) * S.xstride;
        const float* bp = (const float*)((const unsigned char*)B.pixels
                                         + (long long)(y - B.y) * B.ystride
                                         + (long long)(x - B.x) * B.xstride);
        unsigned char* dp = (unsigned char*)p.dst.pixels
                            + (long long)(y - p.dst.y) * p.dst.ystride
                            + (long long)(x - p.dst.x) * p.dst.xstride;
        for (int c = p.chbegin; c < p.chend; ++c) {
            const float s    = in ? load_elem(sp + c * ses, S.basetype) : 0.0f;
            const float b    = in ? __ldg(bp + c) : 0.0f;
            const float diff = __fsub_rn(s, b);
            float v          = s;
            if (fabsf(diff) > p.threshold)
                v = __fadd_rn(s, __fmul_rn(p.contrast, diff));
            store_elem(dp + c * des, p.dst.basetype, v);
        }
    }
}

}  // namespace

void
launch_convolve(const ConvParams& p, void* stream)
{
    const int w = p.roi_x1 - p.roi_x0, h = p.roi_y1 - p.roi_y0;
    if (w <= 0 || h <= 0)
        return;
    const bool vec = p.src.basetype == BT_FLOAT && p.src.nchannels == 4 && p.src.xstride == 16
                     && (uintptr_t)p.src.pixels % 16 == 0 && p.src.ystride % 16 == 0;
    const bool k3 = p.kw == 3 && p.kh == 3, k5 = p.kw == 5 && p.kh == 5;
    const bool unrolled = k3 || (vec && k5);
    const int rows_per_cta = CONV_BY * (unrolled ? CONV_RPT : 1);
    dim3 grid((w + CONV_BX - 1) / CONV_BX,
              std::min((h + rows_per_cta - 1) / rows_per_cta, 32768), 1);
    cudaStream_t st = (cudaStream_t)stream;
    const dim3 block(CONV_BX, CONV_BY);
    if (vec && k3)  // unsharp_mask's default, maketx --sharpen
        launch_pdl(convolve_kernel<true, 3, 3>, grid, block, 0, st, p);
    else if (vec && k5)
        launch_pdl(convolve_kernel<true, 5, 5>, grid, block, 0, st, p);
    else if (vec)
        launch_pdl(convolve_kernel<true, 0, 0>, grid, block, 0, st, p);
    else if (k3)
        launch_pdl(convolve_kernel<false, 3, 3>, grid, block, 0, st, p);
    else
        launch_pdl(convolve_kernel<false, 0, 0>, grid, block, 0, st, p);
}

void
launch_unsharp_combine(const UnsharpParams& p, void* stream)
{
    const int w = p.roi_x1 - p.roi_x0, h = p.roi_y1 - p.roi_y0;
    if (w <= 0 || h <= 0)
        return;
    dim3 grid((w + 255) / 256, std::min(h, 32768), 1);
    launch_pdl(unsharp_combine_kernel, grid, dim3(256), 0, (cudaStream_t)stream, p);
}

}  // namespace b2r

// ===========================================================================
// fixNonFinite (imagebufalgo_pixelmath.cpp:1424-1500) and maketx's
// check_nan_block (maketexture.cpp:340-360): count the pixels that hold a
// NaN/Inf and, per mode, overwrite those values with 0 or with the average of
// the finite values in the 3x3 neighbourhood (clipped to the image).  The
// neighbourhood is read from the SOURCE image, so the result does not depend
// on any scan order (the reference repairs in place, band by band; the two
// agree wherever non-finite values are not adjacent).  Half images sum in
// half precision like the reference's `T sum`.
// ===========================================================================
namespace b2r {
namespace {

template<bool HALF>
__device__ __forceinline__ float
nf_load(const unsigned char* p)
{
    return HALF ? __half2float(*(const __half*)p) : *(const float*)p;
}

template<bool HALF>
__global__ void __launch_bounds__(256)
fix_nonfinite_kernel(const NonFiniteParams p)
{
    pdl_prologue();
    const int x = p.roi_x0 + blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = x < p.roi_x1;
    const int es    = HALF ? 2 : 4;
    const DevImage& S = p.src;
    int fixed_here    = 0;
    for (int y = p.roi_y0 + blockIdx.y; y < p.roi_y1; y += gridDim.y) {
        bool fixed = false;
        if (live) {
            const bool in = x >= S.x && x < S.x + S.width && y >= S.y && y < S.y + S.height;
            const unsigned char* sp = (const unsigned char*)S.pixels
                                      + (long long)(y - S.y) * S.ystride
                                      + (long long)(x - S.x) * S.xstride;
            unsigned char* dp = (unsigned char*)p.dst.pixels
                                + (long long)(y - p.dst.y) * p.dst.ystride
                                + (long long)(x - p.dst.x) * p.dst.xstride;
            for (int c = 0; c < p.dst.nchannels; ++c) {
                float v = in ? nf_load<HALF>(sp + c * es) : 0.0f;
                if (c >= p.chbegin && c < p.chend && nonfinite(v)) {
                    fixed = true;
                    if (p.mode == 1) {
                        v = 0.0f;
                    } else if (p.mode == 2) {
                        int numvals = 0;
                        float sum   = 0.0f;
                        // roi2 = 3x3 around the pixel, clipped to dst's window
                        const int j0 = max(y - 1, p.dst.y), j1 = min(y + 2, p.dst.y + p.dst.height);
                        const int i0 = max(x - 1, p.dst.x), i1 = min(x + 2, p.dst.x + p.dst.width);
                        for (int j = j0; j < j1; ++j)
                            for (int i = i0; i < i1; ++i) {
                                const bool nin = i >= S.x && i < S.x + S.width && j >= S.y
                                                 && j < S.y + S.height;
                                const float nv
                                    = nin ? nf_load<HALF>((const unsigned char*)S.pixels
                                                          + (long long)(j - S.y) * S.ystride
                                                          + (long long)(i - S.x) * S.xstride
                                                          + c * es)
                                          : 0.0f;
                                if (!nonfinite(nv)) {
                                    sum = __fadd_rn(sum, nv);
                                    if (HALF)
                                        sum = __half2float(__float2half_rn(sum));
                                    ++numvals;
                                }
                            }
                        v = numvals ? __fdiv_rn(sum, float(numvals)) : 0.0f;
                    }
                }
                if (p.write) {
                    if (HALF)
                        *(__half*)(dp + c * es) = __float2half_rn(v);
                    else
                        *(float*)(dp + c * es) = v;
                }
            }
        }
        fixed_here += fixed ? 1 : 0;
    }
    // one atomic per warp
    const unsigned m = __ballot_sync(0xffffffffu, fixed_here != 0);
    int total = fixed_here;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        total += __shfl_xor_sync(0xffffffffu, total, o);
    if (m && (threadIdx.x & 31) == 0)
        atomicAdd(p.counter, total);
}

}  // namespace

void
launch_fix_nonfinite(const NonFiniteParams& p, void* stream)
{
    const int w = p.roi_x1 - p.roi_x0, h = p.roi_y1 - p.roi_y0;
    if (w <= 0 || h <= 0)
        return;
    dim3 grid((w + 255) / 256, std::min(h, 4096), 1);
    if (p.src.basetype == BT_HALF)
        launch_pdl(fix_nonfinite_kernel<true>, grid, dim3(256), 0, (cudaStream_t)stream, p);
    else
        launch_pdl(fix_nonfinite_kernel<false>, grid, dim3(256), 0, (cudaStream_t)stream, p);
}

}  // namespace b2r

// ===========================================================================
// The two point ops of fillholes_pushpull (imagebufalgo.cpp:1578-1596 and
// imagebufalgo_pixelmath.cpp:1687-1692, 1733-1739) on contiguous float
// images: divide the pixels with nonzero alpha by their alpha, and
// big = big OVER blowup (big + (1 - clamp(alpha_big, 0, 1)) * blowup).
// ===========================================================================
namespace b2r {
namespace {

__global__ void __launch_bounds__(256)
divide_by_alpha_kernel(float* px, long long npixels, int nch, int alpha_channel)
{
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npixels;
         i += (long long)gridDim.x * blockDim.x) {
        float* p      = px + i * nch;
        const float a = p[alpha_channel];
        if (a != 0.0f)
            for (int c = 0; c < nch; ++c)
                p[c] = __fdiv_rn(p[c], a);
    }
}

__global__ void __launch_bounds__(256)
over_inplace_kernel(float* big, const float* blowup, long long npixels, int nch,
                    int alpha_channel)
{
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < npixels;
         i += (long long)gridDim.x * blockDim.x) {
        float* a       = big + i * nch;
        const float* b = blowup + i * nch;
        const float alpha = fminf(fmaxf(a[alpha_channel], 0.0f), 1.0f);
        const float oma   = __fsub_rn(1.0f, alpha);
        for (int c = 0; c < nch; ++c)
            a[c] = __fadd_rn(a[c], __fmul_rn(oma, b[c]));
    }
}

unsigned
point_grid(long long n)
{
    return (unsigned)std::min<long long>((n + 255) / 256, 148 * 16);
}

}  // namespace

void
launch_divide_by_alpha(float* px, long long npixels, int nch, int alpha_channel, void* stream)
{
    if (npixels > 0)
        launch_pdl(divide_by_alpha_kernel, dim3(point_grid(npixels)), dim3(256), 0,
                   (cudaStream_t)stream, px, npixels, nch, alpha_channel);
}

void
launch_over_inplace(float* big, const float* blowup, long long npixels, int nch,
                    int alpha_channel, void* stream)
{
    if (npixels > 0)
        launch_pdl(over_inplace_kernel, dim3(point_grid(npixels)), dim3(256), 0,
                   (cudaStream_t)stream, big, blowup, npixels, nch, alpha_channel);
}

}  // namespace b2r
