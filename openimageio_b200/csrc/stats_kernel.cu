// stats_kernel.cu -- ImageBufAlgo::computePixelStats + isMonochrome in ONE pass
// over the image (src/libOpenImageIO/imagebufalgo_compare.cpp:81-215, 539-600):
// what make_texture_impl runs over the whole source before anything else
// (maketexture.cpp:1396-1409, 1469-1476; compute_average is on by default).
//
// Per channel: min / max (float), sum and sum of squares (double, like
// PixelStats::sum / sum2), counts of finite / NaN / Inf values; plus one flag,
// "every pixel has all its channels within `threshold` of its first".
// A thread walks whole pixels with register accumulators (channels in passes
// of four), a CTA reduces through shuffles and shared memory and writes one
// partial per (block, channel); the host merges the partials in block order
// (PixelStats::merge), so the result is deterministic.  HBM-bound: one read of
// the image, nothing written but a few kilobytes.
#include "dev_convert.cuh"

#include <algorithm>

namespace b2r {
namespace {

constexpr int STATS_THREADS = 256;

struct Acc {
    double sum, sum2;
    float mn, mx;
    unsigned long long finite, nan, inf;
};

__device__ __forceinline__ void
acc_init(Acc& a)
{
    a.sum = a.sum2 = 0.0;
    a.mn           = __int_as_float(0x7f800000);
    a.mx           = __int_as_float(0xff800000);
    a.finite = a.nan = a.inf = 0ull;
}

__device__ __forceinline__ void
acc_val(Acc& a, float v)
{
    if (v != v) {
        ++a.nan;
    } else if (fabsf(v) == __int_as_float(0x7f800000)) {
        ++a.inf;
    } else {
        ++a.finite;
        a.sum += double(v);
        a.sum2 += double(v) * double(v);
        a.mn = fminf(v, a.mn);
        a.mx = fmaxf(v, a.mx);
    }
}

__device__ __forceinline__ void
acc_merge(Acc& a, const Acc& b)
{
    a.sum += b.sum;
    a.sum2 += b.sum2;
    a.mn = fminf(a.mn, b.mn);
    a.mx = fmaxf(a.mx, b.mx);
    a.finite += b.finite;
    a.nan += b.nan;
    a.inf += b.inf;
}

__device__ __forceinline__ Acc
acc_shfl_xor(const Acc& a, int o)
{
    Acc b;
    b.sum    = __shfl_xor_sync(0xffffffffu, a.sum, o);
    b.sum2   = __shfl_xor_sync(0xffffffffu, a.sum2, o);
    b.mn     = __shfl_xor_sync(0xffffffffu, a.mn, o);
    b.mx     = __shfl_xor_sync(0xffffffffu, a.mx, o);
    b.finite = __shfl_xor_sync(0xffffffffu, a.finite, o);
    b.nan    = __shfl_xor_sync(0xffffffffu, a.nan, o);
    b.inf    = __shfl_xor_sync(0xffffffffu, a.inf, o);
    return b;
}

template<bool VEC_F32>
__global__ void __launch_bounds__(STATS_THREADS)
pixel_stats_kernel(const StatsParams p)
{
    pdl_prologue();
    __shared__ Acc sh[STATS_THREADS / 32];
    __shared__ int sh_mono;
    const DevImage& S = p.src;
    const int es  = S.basetype == BT_UINT8 ? 1 : (S.basetype == BT_FLOAT ? 4 : 2);
    int mono = 1;
    if (threadIdx.x == 0)
        sh_mono = 1;
    for (int c0 = p.chbegin; c0 < p.chend; c0 += 4) {
        Acc a[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            acc_init(a[k]);
        const int nc = min(4, p.chend - c0);
        // a CTA takes whole rows (no div/mod per pixel), threads stride along x
        for (int y = p.roi_y0 + blockIdx.x; y < p.roi_y1; y += gridDim.x)
        for (int x = p.roi_x0 + threadIdx.x; x < p.roi_x1; x += STATS_THREADS) {
            const unsigned char* sp = (const unsigned char*)S.pixels
                                      + (long long)(y - S.y) * S.ystride
                                      + (long long)(x - S.x) * S.xstride;
            float v[4] = { 0.f, 0.f, 0.f, 0.f };
            if (VEC_F32) {
                const float4 q = __ldg((const float4*)sp);
                v[0] = q.x, v[1] = q.y, v[2] = q.z, v[3] = q.w;
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nc)
                        v[k] = load_elem(sp + (c0 + k) * es, S.basetype);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (k < nc)
                    acc_val(a[k], v[k]);
            if (p.want_mono) {
                // every channel of the ROI against the first one (:554-571)
                const float first = (c0 == p.chbegin) ? v[0]
                                                      : load_elem(sp + p.chbegin * es, S.basetype);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < nc && c0 + k > p.chbegin) {
                        const bool same = p.mono_threshold == 0.0f
                                              ? (v[k] == first)
                                              : !(fabsf(__fsub_rn(v[k], first)) > p.mono_threshold);
                        mono &= same ? 1 : 0;
                    }
            }
        }
        // CTA reduction, one channel at a time
        for (int k = 0; k < nc; ++k) {
            Acc r = a[k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc_merge(r, acc_shfl_xor(r, o));
            __syncthreads();
            if ((threadIdx.x & 31) == 0)
                sh[threadIdx.x >> 5] = r;
            __syncthreads();
            if (threadIdx.x == 0) {
                Acc t = sh[0];
                for (int q = 1; q < STATS_THREADS / 32; ++q)
                    acc_merge(t, sh[q]);
                StatsPartial& out = p.partials[(size_t)blockIdx.x * p.nchannels + c0 + k];
                out.sum = t.sum, out.sum2 = t.sum2;
                out.mn = t.mn, out.mx = t.mx;
                out.finite = t.finite, out.nan = t.nan, out.inf = t.inf;
            }
        }
    }
    if (p.want_mono) {
        if (!mono)
            sh_mono = 0;  // benign race: everyone writes 0
        __syncthreads();
        if (threadIdx.x == 0)
            p.mono_flags[blockIdx.x] = sh_mono;
    }
}

}  // namespace

int
stats_grid_blocks(int rows)
{
    return std::max(1, std::min(rows, 148 * 8));
}

void
launch_pixel_stats(const StatsParams& p, int nblocks, void* stream)
{
    const bool vec = p.src.basetype == BT_FLOAT && p.src.nchannels == 4 && p.chbegin == 0
                     && p.chend == 4 && p.src.xstride == 16
                     && (uintptr_t)p.src.pixels % 16 == 0 && p.src.ystride % 16 == 0;
    if (vec)
        launch_pdl(pixel_stats_kernel<true>, dim3(nblocks), dim3(STATS_THREADS), 0,
                   (cudaStream_t)stream, p);
    else
        launch_pdl(pixel_stats_kernel<false>, dim3(nblocks), dim3(STATS_THREADS), 0,
                   (cudaStream_t)stream, p);
}

}  // namespace b2r
