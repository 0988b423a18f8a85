// warp_kernel.cu -- ImageBufAlgo::warp on the device: one thread per
// destination pixel, the reference's own footprint and tap order.
//
// Follows warp_<D,S> + filtered_sample<S> + robust_multVecMatrix
// (src/libOpenImageIO/imagebufalgo_xform.cpp:353-375, 263-325, 210-226) and the
// wrap rules of ImageBufImpl::do_wrap (imagebuf.cpp:3283-3322, imageio.cpp:
// 1398-1434).  Every float operation that decides WHICH source pixels a
// destination pixel reads (the projective transform with its differentials,
// the footprint bounds) is an explicit round-to-nearest multiply / add /
// divide in the reference's expression order, so the footprint is the
// reference's to the pixel; the taps are accumulated in the reference's order
// (rows outer, columns inner, multiply then add).  The only arithmetic that
// can differ from the CPU is sinf / cosf inside the lanczos3 / blackman-harris
// / sinc profiles (device libm, <= 2 ulp).
//
// A separable Filter2D is xfilt(x) * yfilt(y): the row factor is evaluated
// once per footprint row and the column factors once per pixel (kept in a
// small per-thread array) instead of once per tap -- same values, same
// product, about footprint-width times fewer profile evaluations.
//
// Memory: neighbouring threads' footprints overlap almost entirely, so the
// source goes through the read-only L1 path (ld.global.nc) from 32x8-pixel
// thread blocks; whole pixels are fetched with one vector load when the
// layout allows.  The kernel is FP32/L1 bound, not HBM bound (49 taps per
// pixel at the default lanczos3).  Pixels whose footprint lies inside the data
// window (almost all of them) take a fast path without the per-tap wrap tests.
#include "dev_convert.cuh"
#include "dev_filters.cuh"

#include <algorithm>

namespace b2r {
namespace {

constexpr int WARP_BX  = 32;
constexpr int WARP_BY  = 8;
constexpr int WARP_WXC = 24;  // column factors cached per thread

__device__ __forceinline__ int
wrap_mirror(int c, int origin, int width)
{
    c -= origin;
    if (c < 0)
        c = -c - 1;
    int iter = c / width;
    c -= iter * width;
    if (iter & 1)
        c = width - 1 - c;
    return c + origin;
}

__device__ __forceinline__ int
wrap_periodic(int c, int origin, int width)
{
    c -= origin;
    c %= width;
    if (c < 0)
        c += width;
    return c + origin;
}

// wrap: 0 default, 1 black, 2 clamp, 3 periodic, 4 mirror.
__device__ __forceinline__ bool
inside(int v, int b, int n)
{
    return v >= b && v < b + n;
}

template<int BT> struct SrcElem;
template<> struct SrcElem<BT_UINT8> {
    typedef unsigned char T;
    typedef uchar4 V4;
};
template<> struct SrcElem<BT_UINT16> {
    typedef unsigned short T;
    typedef ushort4 V4;
};
template<> struct SrcElem<BT_HALF> {
    typedef __half T;
    typedef uint2 V4;
};
template<> struct SrcElem<BT_FLOAT> {
    typedef float T;
    typedef float4 V4;
};

template<int BT>
__device__ __forceinline__ float
to_float(typename SrcElem<BT>::T v);
template<>
__device__ __forceinline__ float
to_float<BT_UINT8>(unsigned char v)
{
    return __fmul_rn(float(v), 1.0f / 255.0f);
}
template<>
__device__ __forceinline__ float
to_float<BT_UINT16>(unsigned short v)
{
    return __fmul_rn(float(v), 1.0f / 65535.0f);
}
template<>
__device__ __forceinline__ float
to_float<BT_HALF>(__half v)
{
    return __half2float(v);
}
template<>
__device__ __forceinline__ float
to_float<BT_FLOAT>(float v)
{
    return v;
}

// up to four channels of one source pixel, as float
template<int BT, bool VEC>
__device__ __forceinline__ void
load_px(const unsigned char* p, int n, float v[4])
{
    typedef typename SrcElem<BT>::T T;
    if (VEC) {
        if (BT == BT_FLOAT) {
            float4 q = __ldg((const float4*)p);
            v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
        } else if (BT == BT_HALF) {
            uint2 q    = __ldg((const uint2*)p);
            __half2 lo = *(__half2*)&q.x, hi = *(__half2*)&q.y;
            v[0] = __low2float(lo), v[1] = __high2float(lo);
            v[2] = __low2float(hi), v[3] = __high2float(hi);
        } else if (BT == BT_UINT16) {
            ushort4 q = __ldg((const ushort4*)p);
            v[0] = to_float<BT_UINT16>(q.x), v[1] = to_float<BT_UINT16>(q.y);
            v[2] = to_float<BT_UINT16>(q.z), v[3] = to_float<BT_UINT16>(q.w);
        } else {
            uchar4 q = __ldg((const uchar4*)p);
            v[0] = to_float<BT_UINT8>(q.x), v[1] = to_float<BT_UINT8>(q.y);
            v[2] = to_float<BT_UINT8>(q.z), v[3] = to_float<BT_UINT8>(q.w);
        }
        return;
    }
    const T* q = (const T*)p;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        v[c] = (c < n) ? to_float<BT>(__ldg(q + c)) : 0.0f;
}

// sum += w * v over the four channels.  FMA = false: multiply, then add, each
// rounded -- the reference's sequence (warp_ accumulates through filtered_sample's
// `sum += w * pixel` compiled without contraction).  FMA = true (b2r_options::
// contract_fma, "tolerance mode"): two packed fused multiply-adds -- a quarter
// of the floating-point instructions of a kernel that is issue-bound on them;
// results move by an ulp of the accumulator, inside north_star's 1e-5 / 1 LSB.
template<bool FMA>
__device__ __forceinline__ void
accum4(float (&sum)[4], float w, const float (&v)[4])
{
    if (FMA) {
        const float2 ww = make_float2(w, w);
        const float2 lo = __ffma2_rn(ww, make_float2(v[0], v[1]), make_float2(sum[0], sum[1]));
        const float2 hi = __ffma2_rn(ww, make_float2(v[2], v[3]), make_float2(sum[2], sum[3]));
        sum[0] = lo.x, sum[1] = lo.y, sum[2] = hi.x, sum[3] = hi.y;
    } else {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            sum[c] = __fadd_rn(sum[c], __fmul_rn(w, v[c]));
    }
}

template<int BT, bool VEC, bool FMA>
__global__ void __launch_bounds__(WARP_BX* WARP_BY)
warp_kernel(const WarpParams p)
{
    pdl_prologue();
    const int x = p.roi_x0 + blockIdx.x * WARP_BX + threadIdx.x;
    if (x >= p.roi_x1)
        return;
    const DevImage& S = p.src;
    const int ses     = (BT == BT_UINT8) ? 1 : (BT == BT_FLOAT ? 4 : 2);
    const int des     = p.dst.basetype == BT_UINT8 ? 1 : (p.dst.basetype == BT_FLOAT ? 4 : 2);
    DevFilter filt;
    filt.profile = p.filt_profile, filt.scale = p.filt_scale;
    filt.param = p.filt_param, filt.w = p.filt_w, filt.h = p.filt_h;
    const bool separable = p.filt_scale < 5;
    const DevAxis ax = dev_make_axis(filt, filt.w), ay = dev_make_axis(filt, filt.h);
    const int sxb = S.x, sxe = S.x + S.width, syb = S.y, sye = S.y + S.height;

    for (int y = p.roi_y0 + blockIdx.y * WARP_BY + threadIdx.y; y < p.roi_y1;
         y += gridDim.y * WARP_BY) {
        // ---- robust_multVecMatrix on Dual2 (x + .5, 1, 0), (y + .5, 0, 1)
        const float xv = __fadd_rn(float(x), 0.5f), yv = __fadd_rn(float(y), 0.5f);
        const float* M = p.minv;  // M[i*3+j] = Imath x[i][j]
        float av  = __fadd_rn(__fadd_rn(__fmul_rn(xv, M[0]), __fmul_rn(yv, M[3])), M[6]);
        float adx = __fadd_rn(__fmul_rn(1.0f, M[0]), __fmul_rn(0.0f, M[3]));
        float ady = __fadd_rn(__fmul_rn(0.0f, M[0]), __fmul_rn(1.0f, M[3]));
        float bv  = __fadd_rn(__fadd_rn(__fmul_rn(xv, M[1]), __fmul_rn(yv, M[4])), M[7]);
        float bdx = __fadd_rn(__fmul_rn(1.0f, M[1]), __fmul_rn(0.0f, M[4]));
        float bdy = __fadd_rn(__fmul_rn(0.0f, M[1]), __fmul_rn(1.0f, M[4]));
        float wv  = __fadd_rn(__fadd_rn(__fmul_rn(xv, M[2]), __fmul_rn(yv, M[5])), M[8]);
        float wdx = __fadd_rn(__fmul_rn(1.0f, M[2]), __fmul_rn(0.0f, M[5]));
        float wdy = __fadd_rn(__fmul_rn(0.0f, M[2]), __fmul_rn(1.0f, M[5]));
        float s = 0.0f, t = 0.0f, dsdx = 0.0f, dsdy = 0.0f, dtdx = 0.0f, dtdy = 0.0f;
        if (wv != 0.0f) {
            const float winv = __fdiv_rn(1.0f, wv);
            s                = __fmul_rn(av, winv);
            dsdx             = __fmul_rn(winv, __fsub_rn(adx, __fmul_rn(s, wdx)));
            dsdy             = __fmul_rn(winv, __fsub_rn(ady, __fmul_rn(s, wdy)));
            t                = __fmul_rn(bv, winv);
            dtdx             = __fmul_rn(winv, __fsub_rn(bdx, __fmul_rn(t, wdx)));
            dtdy             = __fmul_rn(winv, __fsub_rn(bdy, __fmul_rn(t, wdy)));
        }
        // ---- filtered_sample: footprint
        const float ds     = fmaxf(1.0f, fmaxf(fabsf(dsdx), fabsf(dsdy)));
        const float dt     = fmaxf(1.0f, fmaxf(fabsf(dtdx), fabsf(dtdy)));
        const float ds_inv = __fdiv_rn(1.0f, ds), dt_inv = __fdiv_rn(1.0f, dt);
        const float rad_s  = __fmul_rn(__fmul_rn(0.5f, ds), p.filt_w);
        const float rad_t  = __fmul_rn(__fmul_rn(0.5f, dt), p.filt_w);  // width() for both
        int smin = (int)floorf(__fsub_rn(s, rad_s)), smax = (int)ceilf(__fadd_rn(s, rad_s));
        int tmin = (int)floorf(__fsub_rn(t, rad_t)), tmax = (int)ceilf(__fadd_rn(t, rad_t));
        bool black = false;
        if (p.edgeclamp) {
            smin = min(max(smin, sxb), sxe), smax = min(max(smax, sxb), sxe);
            tmin = min(max(tmin, syb), sye), tmax = min(max(tmax, syb), sye);
            black = (s < float(sxb - 1) || s >= float(sxe) || t < float(syb - 1)
                     || t >= float(sye));
        }
        unsigned char* dp = (unsigned char*)p.dst.pixels
                            + (long long)(y - p.dst.y) * p.dst.ystride
                            + (long long)(x - p.dst.x) * p.dst.xstride;
        const int ntx    = smax - smin;
        const bool cache = separable && ntx <= WARP_WXC;
        float wxc[WARP_WXC];
        if (cache && !black) {
#pragma unroll 1
            for (int i = 0; i < ntx; ++i)
                wxc[i] = dev_axis_eval(
                    ax,
                    __fmul_rn(ds_inv, __fsub_rn(__fadd_rn(float(smin + i), 0.5f), s)));
        }
        // channels in passes of four (one pass for everything up to RGBA)
        for (int c0 = p.chbegin & ~3; c0 < p.chend; c0 += 4) {
            const int ncs = min(4, S.nchannels - c0);  // source channels in this pass
            float sum[4]  = { 0.0f, 0.0f, 0.0f, 0.0f };
            float total_w = 0.0f;
            // Interior fast path: the whole footprint lies inside the data
            // window and the column factors are cached -- no wrap tests, one
            // row pointer per footprint row, the same taps in the same order.
            const bool interior = cache && !black && smin >= sxb && smax <= sxe
                                  && tmin >= syb && tmax <= sye;
            if (interior) {
                const unsigned char* rowp = (const unsigned char*)S.pixels
                                            + (long long)(tmin - syb) * S.ystride
                                            + (long long)(smin - sxb) * S.xstride + c0 * ses;
#pragma unroll 1
                for (int ty = tmin; ty < tmax; ++ty, rowp += S.ystride) {
                    const float yarg
                        = __fmul_rn(dt_inv, __fsub_rn(__fadd_rn(float(ty), 0.5f), t));
                    const float wy          = dev_axis_eval(ay, yarg);
                    const unsigned char* sp = rowp;
#pragma unroll 1
                    for (int i = 0; i < ntx; ++i, sp += S.xstride) {
                        const float w = __fmul_rn(wxc[i], wy);
                        float v[4];
                        load_px<BT, VEC>(sp, ncs, v);
                        accum4<FMA>(sum, w, v);
                        total_w = __fadd_rn(total_w, w);
                    }
                }
            } else if (!black) {
#pragma unroll 1
                for (int ty = tmin; ty < tmax; ++ty) {
                    const float yarg
                        = __fmul_rn(dt_inv, __fsub_rn(__fadd_rn(float(ty), 0.5f), t));
                    const float wy = separable ? dev_axis_eval(ay, yarg) : 0.0f;
                    // do_wrap: a pixel inside the data window is itself; one
                    // outside is wrapped in BOTH axes against the full window
                    // and is black if it still misses the data window
                    const bool row_in = inside(ty, syb, S.height);
                    int wyc           = ty;
                    if (p.wrap == 2)
                        wyc = min(max(ty, S.full_y), S.full_y + S.full_height - 1);
                    else if (p.wrap == 3)
                        wyc = wrap_periodic(ty, S.full_y, S.full_height);
                    else if (p.wrap == 4)
                        wyc = wrap_mirror(ty, S.full_y, S.full_height);
#pragma unroll 1
                    for (int tx = smin; tx < smax; ++tx) {
                        float w;
                        if (cache)
                            w = __fmul_rn(wxc[tx - smin], wy);
                        else {
                            const float xarg = __fmul_rn(
                                ds_inv, __fsub_rn(__fadd_rn(float(tx), 0.5f), s));
                            w = separable
                                    ? __fmul_rn(dev_axis_eval(ax, xarg), wy)
                                    : dev_filter2d(filt, xarg, yarg);
                        }
                        int px = tx, py = ty;
                        bool ok = row_in && inside(tx, sxb, S.width);
                        if (!ok && p.wrap >= 2) {
                            py = wyc;
                            if (p.wrap == 2)
                                px = min(max(tx, S.full_x), S.full_x + S.full_width - 1);
                            else if (p.wrap == 3)
                                px = wrap_periodic(tx, S.full_x, S.full_width);
                            else
                                px = wrap_mirror(tx, S.full_x, S.full_width);
                            ok = inside(px, sxb, S.width) && inside(py, syb, S.height);
                        }
                        if (ok) {
                            const unsigned char* sp = (const unsigned char*)S.pixels
                                                      + (long long)(py - syb) * S.ystride
                                                      + (long long)(px - sxb) * S.xstride
                                                      + c0 * ses;
                            float v[4];
                            load_px<BT, VEC>(sp, ncs, v);
                            accum4<FMA>(sum, w, v);
                        }
                        total_w = __fadd_rn(total_w, w);
                    }
                }
            }
            const bool norm = total_w > 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ch = c0 + c;
                if (ch >= p.chbegin && ch < p.chend) {
                    // channels the source does not have stay at warp_'s memset 0
                    const float r = (norm && c < ncs) ? __fdiv_rn(sum[c], total_w) : 0.0f;
                    store_elem(dp + ch * des, p.dst.basetype, r);
                }
            }
        }
    }
}

// st_warp_<D,S,ST> (imagebufalgo_xform.cpp:1834-1925): the source position of
// every destination pixel comes from two channels of an ST image; the
// footprint is filter radius / scale source pixels around it, clipped to the
// data window, inclusive of its upper bound (rerange(x_min, x_max + 1, ...)),
// and a tap that falls outside the data window reads black.
template<int BT, bool VEC>
__global__ void __launch_bounds__(WARP_BX* WARP_BY)
st_warp_kernel(const StWarpParams p)
{
    pdl_prologue();
    const int x = p.roi_x0 + blockIdx.x * WARP_BX + threadIdx.x;
    if (x >= p.roi_x1)
        return;
    const DevImage& S = p.src;
    const int ses     = (BT == BT_UINT8) ? 1 : (BT == BT_FLOAT ? 4 : 2);
    const int des     = p.dst.basetype == BT_UINT8 ? 1 : (p.dst.basetype == BT_FLOAT ? 4 : 2);
    const int tes     = p.st.basetype == BT_UINT8 ? 1 : (p.st.basetype == BT_FLOAT ? 4 : 2);
    DevFilter filt;
    filt.profile = p.filt_profile, filt.scale = p.filt_scale;
    filt.param = p.filt_param, filt.w = p.filt_w, filt.h = p.filt_h;
    const bool separable = p.filt_scale < 5;
    const DevAxis ax = dev_make_axis(filt, filt.w), ay = dev_make_axis(filt, filt.h);
    const int sxb = S.x, sxe = S.x + S.width, syb = S.y, sye = S.y + S.height;
    const float fsw = float(S.full_width), fsh = float(S.full_height);
    const float radx = float(p.filterrad_x), rady = float(p.filterrad_y);

    for (int y = p.roi_y0 + blockIdx.y * WARP_BY + threadIdx.y; y < p.roi_y1;
         y += gridDim.y * WARP_BY) {
        const unsigned char* tp = (const unsigned char*)p.st.pixels
                                  + (long long)(y - p.st.y) * p.st.ystride
                                  + (long long)(x - p.st.x) * p.st.xstride;
        float src_s = load_elem(tp + p.chan_s * tes, p.st.basetype);
        float src_t = load_elem(tp + p.chan_t * tes, p.st.basetype);
        if (p.flip_s)
            src_s = __fsub_rn(1.0f, src_s);
        if (p.flip_t)
            src_t = __fsub_rn(1.0f, src_t);
        const float src_x = __fmul_rn(src_s, fsw), src_y = __fmul_rn(src_t, fsh);
        const int x_min   = min(max((int)floorf(__fsub_rn(src_x, radx)), sxb), sxe);
        const int x_max   = min(max((int)ceilf(__fadd_rn(src_x, radx)), sxb), sxe);
        const int y_min   = min(max((int)floorf(__fsub_rn(src_y, rady)), syb), sye);
        const int y_max   = min(max((int)ceilf(__fadd_rn(src_y, rady)), syb), sye);
        unsigned char* dp = (unsigned char*)p.dst.pixels
                            + (long long)(y - p.dst.y) * p.dst.ystride
                            + (long long)(x - p.dst.x) * p.dst.xstride;
        const int ntx    = x_max + 1 - x_min;
        const bool cache = separable && ntx <= WARP_WXC;
        float wxc[WARP_WXC];
        if (cache) {
#pragma unroll 1
            for (int i = 0; i < ntx; ++i)
                wxc[i] = dev_axis_eval(ax,
                                        __fadd_rn(__fsub_rn(float(x_min + i), src_x), 0.5f));
        }
        for (int c0 = p.chbegin & ~3; c0 < p.chend; c0 += 4) {
            const int ncs = min(4, S.nchannels - c0);
            float sum[4]  = { 0.0f, 0.0f, 0.0f, 0.0f };
            float total_w = 0.0f;
#pragma unroll 1
            for (int ty = y_min; ty <= y_max; ++ty) {
                const float yarg = __fadd_rn(__fsub_rn(float(ty), src_y), 0.5f);
                const float wy   = separable ? dev_axis_eval(ay, yarg) : 0.0f;
                const bool row_in = ty < sye;  // ty >= syb by construction
                const unsigned char* sp = (const unsigned char*)S.pixels
                                          + (long long)(ty - syb) * S.ystride
                                          + (long long)(x_min - sxb) * S.xstride + c0 * ses;
#pragma unroll 1
                for (int tx = x_min; tx <= x_max; ++tx, sp += S.xstride) {
                    float w;
                    if (cache)
                        w = __fmul_rn(wxc[tx - x_min], wy);
                    else {
                        const float xarg = __fadd_rn(__fsub_rn(float(tx), src_x), 0.5f);
                        w = separable ? __fmul_rn(dev_axis_eval(ax, xarg), wy)
                                      : dev_filter2d(filt, xarg, yarg);
                    }
                    total_w = __fadd_rn(total_w, w);
                    if (row_in && tx < sxe) {
                        float v[4];
                        load_px<BT, VEC>(sp, ncs, v);
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            sum[c] = __fadd_rn(sum[c], __fmul_rn(v[c], w));
                    }
                }
            }
            const bool norm = total_w > 0.0f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int ch = c0 + c;
                if (ch >= p.chbegin && ch < p.chend)
                    store_elem(dp + ch * des, p.dst.basetype,
                               (norm && c < ncs) ? __fdiv_rn(sum[c], total_w) : 0.0f);
            }
        }
    }
}

template<int BT>
void
launch_st_bt(const StWarpParams& p, bool vec, dim3 grid, cudaStream_t st)
{
    if (vec)
        launch_pdl(st_warp_kernel<BT, true>, grid, dim3(WARP_BX, WARP_BY), 0, st, p);
    else
        launch_pdl(st_warp_kernel<BT, false>, grid, dim3(WARP_BX, WARP_BY), 0, st, p);
}

template<int BT>
void
launch_bt(const WarpParams& p, bool vec, dim3 grid, cudaStream_t st)
{
    if (p.contract_fma) {
        if (vec)
            launch_pdl(warp_kernel<BT, true, true>, grid, dim3(WARP_BX, WARP_BY), 0, st, p);
        else
            launch_pdl(warp_kernel<BT, false, true>, grid, dim3(WARP_BX, WARP_BY), 0, st, p);
    } else if (vec)
        launch_pdl(warp_kernel<BT, true, false>, grid, dim3(WARP_BX, WARP_BY), 0, st, p);
    else
        launch_pdl(warp_kernel<BT, false, false>, grid, dim3(WARP_BX, WARP_BY), 0, st, p);
}

}  // namespace

void
launch_warp(const WarpParams& p, void* stream)
{
    const int w = p.roi_x1 - p.roi_x0, h = p.roi_y1 - p.roi_y0;
    if (w <= 0 || h <= 0)
        return;
    dim3 grid((w + WARP_BX - 1) / WARP_BX, std::min((h + WARP_BY - 1) / WARP_BY, 32768), 1);
    const int es   = basetype_size(p.src.basetype);
    const bool vec = p.src.nchannels == 4 && p.src.xstride == 4 * es
                     && (uintptr_t)p.src.pixels % (4 * es) == 0 && p.src.ystride % (4 * es) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (p.src.basetype) {
    case BT_UINT8: launch_bt<BT_UINT8>(p, vec, grid, st); break;
    case BT_UINT16: launch_bt<BT_UINT16>(p, vec, grid, st); break;
    case BT_HALF: launch_bt<BT_HALF>(p, vec, grid, st); break;
    default: launch_bt<BT_FLOAT>(p, vec, grid, st); break;
    }
}

void
launch_st_warp(const StWarpParams& p, void* stream)
{
    const int w = p.roi_x1 - p.roi_x0, h = p.roi_y1 - p.roi_y0;
    if (w <= 0 || h <= 0)
        return;
    dim3 grid((w + WARP_BX - 1) / WARP_BX, std::min((h + WARP_BY - 1) / WARP_BY, 32768), 1);
    const int es   = basetype_size(p.src.basetype);
    const bool vec = p.src.nchannels == 4 && p.src.xstride == 4 * es
                     && (uintptr_t)p.src.pixels % (4 * es) == 0 && p.src.ystride % (4 * es) == 0;
    cudaStream_t st = (cudaStream_t)stream;
    switch (p.src.basetype) {
    case BT_UINT8: launch_st_bt<BT_UINT8>(p, vec, grid, st); break;
    case BT_UINT16: launch_st_bt<BT_UINT16>(p, vec, grid, st); break;
    case BT_HALF: launch_st_bt<BT_HALF>(p, vec, grid, st); break;
    default: launch_st_bt<BT_FLOAT>(p, vec, grid, st); break;
    }
}

}  // namespace b2r
