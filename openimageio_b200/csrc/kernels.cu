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
#include "dev_convert.cuh"

#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace b2r {

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
    // double-precision sine rounded to float: the correctly rounded value, which
    // is what glibc's sinf returns in all but rare cases.  radial-lanczos3
    // weights are normalised by their sum, which nearly cancels for narrow fixed
    // widths, so a last-bit difference in a weight can show up amplified.
    float s1 = (float)sin((double)__fmul_rn(__fmul_rn(x, ainv), pi));
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

// One destination pixel (kx, ky: ROI-relative), channels [c0, c0 + EXACT_CH).
__device__ __forceinline__ void
exact_pixel(const ExactParams& p, int kx, int ky, int c0)
{
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

__global__ void __launch_bounds__(128)
resize_exact_kernel(ExactParams p)
{
    pdl_prologue();
    if (p.gate && *p.gate != p.gate_value)
        return;
    const int roi_w = p.roi_x1 - p.roi_x0;
    const int roi_h = p.roi_y1 - p.roi_y0;
    const int c0    = blockIdx.z * EXACT_CH;
    if (p.gate && p.dirty) {
        // redo pass: only the cells the fused kernel marked (a cell is 128
        // columns x 2^yshift rows; blockDim.x == 128)
        const int cell_h = 1 << p.dirty_yshift;
        const int ycells = (roi_h + cell_h - 1) >> p.dirty_yshift;
        for (int cell = blockIdx.x; cell < p.dirty_xcells * ycells; cell += gridDim.x) {
            if (p.dirty[cell] != p.gate_value)
                continue;
            const int cy = cell / p.dirty_xcells, cx = cell - cy * p.dirty_xcells;
            const int kx = (cx << DIRTY_CELL_SHIFT_X) + threadIdx.x;
            if (kx >= roi_w)
                continue;
            // the rows of a cell are spread over gridDim.y CTAs
            const int y1 = min(roi_h, (cy + 1) << p.dirty_yshift);
            for (int ky = (cy << p.dirty_yshift) + blockIdx.y; ky < y1; ky += gridDim.y)
                exact_pixel(p, kx, ky, c0);
        }
        return;
    }
    const int xblocks = (roi_w + blockDim.x - 1) / blockDim.x;
    for (long long tile = blockIdx.x; tile < (long long)xblocks * roi_h;
         tile += gridDim.x) {
        const int kx = int(tile % xblocks) * blockDim.x + threadIdx.x;
        const int ky = int(tile / xblocks);
        if (kx < roi_w)
            exact_pixel(p, kx, ky, c0);
    }
}

void
launch_resize_exact(const ExactParams& p, void* stream)
{
    int roi_w = p.roi_x1 - p.roi_x0, roi_h = p.roi_y1 - p.roi_y0;
    if (roi_w <= 0 || roi_h <= 0)
        return;
    dim3 block(128, 1, 1);
    long long tiles = (long long)((roi_w + 127) / 128) * roi_h;
    // grid-stride: enough CTAs to fill the machine, few enough that a launch
    // gated off by *p.gate == 0 retires in a few microseconds
    // a gated launch (redo path) uses a small grid so that the usual case --
    // flag clear, every CTA exits at once -- costs as little as possible
    // (the redo pass walks dirty-map cells: a few per CTA at most)
    const long long cap = p.gate ? 148 * 4 : 148 * 16;
    dim3 grid((unsigned)std::min<long long>(tiles, cap), 1,
              (p.nch + EXACT_CH - 1) / EXACT_CH);
    if (p.gate && p.dirty) {
        // redo pass: x walks the dirty-map cells, y splits a cell's rows
        const int cell_h = 1 << p.dirty_yshift;
        const long long ncells
            = (long long)p.dirty_xcells * ((roi_h + cell_h - 1) >> p.dirty_yshift);
        grid.y = 16;
        grid.x = (unsigned)std::max<long long>(1, std::min<long long>(ncells, cap / 16));
    }
    launch_pdl(resize_exact_kernel, grid, block, 0, (cudaStream_t)stream, p);
}



// ------------------------------------------------------------------------
// fused kernel dispatch: the instantiations live in fused_inst.cu, compiled
// once per (source type, channel count) so the build parallelises.
// ------------------------------------------------------------------------
#define B2R_DECL_LOOKUP(ST, NCH) \
    FusedKernel fused_lookup_##ST##_##NCH(int vk, int ht);
#define B2R_DECL_ST(ST) \
    B2R_DECL_LOOKUP(ST, 1) B2R_DECL_LOOKUP(ST, 2) B2R_DECL_LOOKUP(ST, 3) B2R_DECL_LOOKUP(ST, 4)
B2R_DECL_ST(0) B2R_DECL_ST(1) B2R_DECL_ST(2) B2R_DECL_ST(3)

static int
st_index(int basetype)
{
    switch (basetype) {
    case BT_UINT8: return 0;
    case BT_UINT16: return 1;
    case BT_HALF: return 2;
    case BT_FLOAT: return 3;
    }
    return -1;
}

static FusedKernel
pick_fused(int srctype, int nch, int vk, int ht)
{
    typedef FusedKernel (*Lookup)(int, int);
#define B2R_ROW(ST) \
    { fused_lookup_##ST##_1, fused_lookup_##ST##_2, fused_lookup_##ST##_3, fused_lookup_##ST##_4 }
    static const Lookup table[4][4] = { B2R_ROW(0), B2R_ROW(1), B2R_ROW(2), B2R_ROW(3) };
    int st = st_index(srctype);
    if (st < 0 || nch < 1 || nch > 4)
        return nullptr;
    return table[st][nch - 1](vk, ht);
}

#define B2R_XDECL(n) \
    ExpandKernel expand_lookup_##n##_2(int), expand_lookup_##n##_4(int), \
        expand_lookup_##n##_6(int), expand_lookup_##n##_8(int);
B2R_XDECL(1) B2R_XDECL(2) B2R_XDECL(3) B2R_XDECL(4)
static ExpandKernel
pick_expand(int srctype, int nch, int vt, int ht)
{
    typedef ExpandKernel (*Lookup)(int);
#define B2R_XROW(n) \
    { expand_lookup_##n##_2, expand_lookup_##n##_4, expand_lookup_##n##_6, expand_lookup_##n##_8 }
    static const Lookup table[4][4] = { B2R_XROW(1), B2R_XROW(2), B2R_XROW(3), B2R_XROW(4) };
    if (st_index(srctype) < 0 || nch < 1 || nch > 4)
        return nullptr;
    if (vt != 2 && vt != 4 && vt != 6 && vt != 8)
        return nullptr;
    return table[nch - 1][vt / 2 - 1](ht);
}

bool
fused_supported(int srctype, int nch, int vk, int ht)
{
    return pick_fused(srctype, nch, vk, ht) != nullptr;
}

bool
expand_supported(int srctype, int nch, int vt, int ht)
{
    return pick_expand(srctype, nch, vt, ht) != nullptr;
}

int
fused_ctas_per_sm(const FusedParams& p, int max_band_rows, int max_band_dy,
                  int max_nvec)
{
    const void* k = p.expand ? (const void*)pick_expand(p.srctype, p.nch, p.vt, p.ht)
                             : (const void*)pick_fused(p.srctype, p.nch, p.vk, p.ht);
    if (!k)
        return 0;
    size_t smem = fused_smem_bytes(p, max_band_rows, max_band_dy, max_nvec);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute((const void*)k,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, (const void*)k,
                                                      FUSED_THREADS, smem)
        != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

size_t
fused_smem_bytes(const FusedParams& p, int max_band_rows, int max_band_dy,
                 int max_nvec)
{
    (void)max_nvec;
    if (p.expand) {
        size_t b = (size_t)EXPAND_BATCH * (FUSED_THREADS * 4 + 32) * 4;  // vrows
        b += (size_t)p.ht * EXPAND_GROUPS * 16;                          // hws
        b += (size_t)EXPAND_FIFO * FUSED_THREADS * 4 * basetype_size(p.srctype);
        b += (((size_t)(max_band_dy + 1) * p.vt + 3) & ~size_t(3)) * 4
             + (size_t)((max_band_dy + 1 + 3) & ~3) * 4;  // vws, vfs (+1 row read ahead)
        return (b + 15) & ~size_t(15);
    }
    size_t f = (size_t)fused_batch(p.vk) * (FUSED_THREADS * 4 + 32);  // FUSED_PITCH_B (4224 B)
    f += (size_t)p.ht * FUSED_HCOLS;  // H weights
    if (p.vk > 0) {
        f += (size_t)max_band_rows * ((p.vk + 3) & ~3);
        f += (size_t)((max_band_dy + 3) & ~3);
        f += (size_t)fused_fifo_rows(basetype_size(p.srctype), p.vk) * FUSED_THREADS
             * basetype_size(p.srctype);  // 4*es bytes per thread
    }
    return f * sizeof(float);
}

void
launch_resize_fused(const FusedParams& p, int max_band_rows, int max_band_dy,
                    int max_nvec, void* stream)
{
    if (p.expand) {
        ExpandKernel k = pick_expand(p.srctype, p.nch, p.vt, p.ht);
        if (!k || p.nstrips <= 0 || p.nbands <= 0)
            return;
        size_t smem = fused_smem_bytes(p, max_band_rows, max_band_dy, max_nvec);
        if (smem > 48 * 1024)
            cudaFuncSetAttribute((const void*)k,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        dim3 grid(p.nstrips, p.nbands, 1);
        launch_pdl(k, grid, dim3(FUSED_THREADS), smem, (cudaStream_t)stream, p, max_band_dy);
        return;
    }
    FusedKernel k = pick_fused(p.srctype, p.nch, p.vk, p.ht);
    if (!k || p.nstrips <= 0 || p.nbands <= 0)
        return;
    size_t smem = fused_smem_bytes(p, max_band_rows, max_band_dy, max_nvec);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute((const void*)k,
                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem);
    dim3 grid(p.nstrips, p.nbands, 1);
    launch_pdl(k, grid, dim3(FUSED_THREADS), smem, (cudaStream_t)stream, p, max_band_rows,
               max_band_dy);
}

}  // namespace b2r
