// tile_kernel.cu -- re-lays rows of a finished MIP level into the tiles a
// tiled texture file stores (maketx: 64 x 64, maketexture.cpp:1100-1106,
// 1599-1602), converting float -> half on the way when the file's data type is
// half (maketx -d half).  Output: tile rows [ty0, ty1) of the level, tiles in
// raster order, each tile tile_h x tile_w x nch elements, contiguous; pixels of
// edge tiles that fall outside the level are zero (what ImageOutput::write_tile
// pads with).  HBM-bound streaming: algorithmic bytes = level rows read once +
// tiles written once.
#include "dev_convert.cuh"
#include "kernels.h"

#include <algorithm>

namespace b2r {

namespace {

// One thread = four consecutive samples of a tile row (tile_w * nch is a
// multiple of 4 in the vector kernels; other tile shapes take the scalar one).
struct TileQuad {
    const float* row;  // first sample of source row y (valid when inside)
    bool inside;       // y < h
    int x0;            // first source column of the tile
    int count;         // first of the four samples within the tile row
    long long out_row; // index of the tile row's first sample in the tiled output
};

__device__ __forceinline__ TileQuad
tile_quad(long long i, const float* __restrict__ src, int w, int h, int nch, int tile_w,
          int tile_h, int ntx, int ty0)
{
    const int row_quads   = tile_w * nch / 4;
    const int tile_quads  = row_quads * tile_h;
    // 64-bit divisions are emulated (~100 instructions each, most of this
    // kernel before): every index of a launch below 2^32 quads -- a 64 MB chunk
    // of the encode pipeline, or a whole 32K^2 RGBA level -- divides in 32 bits
    long long t;
    int e, trow, tcol;
    if ((unsigned long long)i >> 32) {
        t    = i / tile_quads;
        e    = int(i - t * tile_quads);
        trow = int(t / ntx);
        tcol = int(t - (long long)trow * ntx);
    } else {
        const unsigned iu = (unsigned)i;
        const unsigned tu = iu / (unsigned)tile_quads;
        e                 = int(iu - tu * (unsigned)tile_quads);
        trow              = int(tu / (unsigned)ntx);
        tcol              = int(tu - (unsigned)trow * (unsigned)ntx);
        t                 = tu;
    }
    const int yy = e / row_quads;
    const int y  = (ty0 + trow) * tile_h + yy;
    TileQuad q;
    q.x0      = tcol * tile_w;
    q.count   = (e - yy * row_quads) * 4;
    q.inside  = y < h;
    q.row     = src + (long long)y * w * nch;
    q.out_row = (t * tile_h + yy) * (long long)(tile_w * nch);
    return q;
}

// samples count .. count+3 of the tile row (zero outside the level)
__device__ __forceinline__ float4
tile_load4(const TileQuad& q, int w, int nch, int count)
{
    float4 v = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    if (!q.inside)
        return v;
    const long long first = (long long)q.x0 * nch + count;  // sample index within the source row
    const long long valid = (long long)w * nch;
    const float* p        = q.row + first;
    if (first + 3 < valid && (reinterpret_cast<unsigned long long>(p) & 15ull) == 0)
        return __ldg(reinterpret_cast<const float4*>(p));
    if (first < valid)
        v.x = __ldg(p);
    if (first + 1 < valid)
        v.y = __ldg(p + 1);
    if (first + 2 < valid)
        v.z = __ldg(p + 2);
    if (first + 3 < valid)
        v.w = __ldg(p + 3);
    return v;
}

template<bool HALF>
__device__ __forceinline__ unsigned
tile_bits(float v)
{
    return HALF ? (unsigned)__half_as_ushort(__float2half_rn(v)) : __float_as_uint(v);
}

template<bool HALF>
__global__ void __launch_bounds__(256)
tile_relayout_vec_kernel(const float* __restrict__ src, int w, int h, int nch, int tile_w,
                         int tile_h, int ntx, int ty0, long long total_quads,
                         void* __restrict__ out)
{
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads;
         i += (long long)gridDim.x * blockDim.x) {
        const TileQuad q = tile_quad(i, src, w, h, nch, tile_w, tile_h, ntx, ty0);
        const float4 v   = tile_load4(q, w, nch, q.count);
        if (HALF) {
            uint2 o;
            o.x = tile_bits<true>(v.x) | (tile_bits<true>(v.y) << 16);
            o.y = tile_bits<true>(v.z) | (tile_bits<true>(v.w) << 16);
            reinterpret_cast<uint2*>(out)[i] = o;
        } else {
            reinterpret_cast<float4*>(out)[i] = v;
        }
    }
}

// scalar form for tile rows that are not a multiple of four samples:
// one thread = one element of the tiled output
template<bool HALF>
__global__ void __launch_bounds__(256)
tile_relayout_kernel(const float* __restrict__ src, int w, int h, int nch, int tile_w, int tile_h,
                     int ntx, int ty0, long long total, void* __restrict__ out)
{
    pdl_prologue();
    const long long tile_elems = (long long)tile_w * tile_h * nch;
    const int row_elems        = tile_w * nch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / tile_elems;
        const int e       = int(i - t * tile_elems);
        const int ty      = ty0 + int(t / ntx);
        const int tx      = int(t % ntx);
        const int yy      = e / row_elems;
        const int xe      = e - yy * row_elems;
        const int y       = ty * tile_h + yy;
        const int x       = tx * tile_w + xe / nch;
        float v           = 0.0f;
        if (y < h && x < w)
            v = __ldg(src + ((long long)y * w + x) * nch + (xe % nch));
        if (HALF)
            reinterpret_cast<__half*>(out)[i] = __float2half_rn(v);
        else
            reinterpret_cast<float*>(out)[i] = v;
    }
}

// The same tiles with TIFF's floating-point predictor applied to every tile row
// (PREDICTOR_FLOATINGPOINT, what the reference's TIFF writer sets for float /
// half data under zip: tiffoutput.cpp:683-690; libtiff's fpDiff): the row's
// samples are split into byte planes, most significant byte first, and every
// byte of the re-ordered row has the byte `nch` positions before it subtracted
// -- the same byte of the same channel of the previous pixel, or, for the
// first pixel of the row, the next more significant byte of the row's LAST
// pixel (the end of the previous plane).  A thread's four samples give four
// consecutive bytes in each plane: one 32-bit store per plane, coalesced
// across the warp; loads are the four samples (128-bit when aligned) and their
// four predecessors.
template<bool HALF>
__global__ void __launch_bounds__(256)
tile_relayout_fp_predictor_kernel(const float* __restrict__ src, int w, int h, int nch,
                                  int tile_w, int tile_h, int ntx, int ty0, long long total_quads,
                                  unsigned* __restrict__ out)
{
    pdl_prologue();
    constexpr int BPS   = HALF ? 2 : 4;
    const int wc        = tile_w * nch;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads;
         i += (long long)gridDim.x * blockDim.x) {
        const TileQuad q = tile_quad(i, src, w, h, nch, tile_w, tile_h, ntx, ty0);
        const float4 v   = tile_load4(q, w, nch, q.count);
        const unsigned cur[4] = { tile_bits<HALF>(v.x), tile_bits<HALF>(v.y), tile_bits<HALF>(v.z),
                                  tile_bits<HALF>(v.w) };
        unsigned prev[4];      // the sample `nch` positions earlier ...
        bool wrapped[4];       // ... or the row's last pixel when there is none
        if (q.count >= nch) {
            const float4 u = tile_load4(q, w, nch, q.count - nch);
            prev[0] = tile_bits<HALF>(u.x), prev[1] = tile_bits<HALF>(u.y);
            prev[2] = tile_bits<HALF>(u.z), prev[3] = tile_bits<HALF>(u.w);
            wrapped[0] = wrapped[1] = wrapped[2] = wrapped[3] = false;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int a = q.count + k - nch;
                wrapped[k]  = a < 0;
                const int s = a < 0 ? a + wc : a;  // sample index within the tile row
                const long long first = (long long)q.x0 * nch + s;
                float f = 0.0f;
                if (q.inside && first < (long long)w * nch)
                    f = __ldg(q.row + first);
                prev[k] = tile_bits<HALF>(f);
            }
        }
        // byte planes of this row in the output, in 32-bit words
        unsigned* orow = out + q.out_row * BPS / 4 + q.count / 4;
#pragma unroll
        for (int plane = 0; plane < BPS; ++plane) {
            const int shift = 8 * (BPS - 1 - plane);
            unsigned word   = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                unsigned b = cur[k] >> shift;
                if (!wrapped[k])
                    b -= prev[k] >> shift;
                else if (plane > 0)
                    b -= prev[k] >> (shift + 8);
                word |= (b & 0xffu) << (8 * k);
            }
            orow[plane * (wc / 4)] = word;
        }
    }
}

// Integer tiles (maketx of a uint8 / uint16 texture writes its float levels
// through convert_type<float,D>, fmath.h:753-764 -- quant_u8 / quant_u16 here,
// bit for bit), optionally with TIFF's horizontal predictor (PREDICTOR_HORIZONTAL,
// what the reference's TIFF writer sets for 8- and 16-bit data under zip:
// tiffoutput.cpp:691-694; libtiff horDiff8 / horDiff16): every sample of a tile
// row minus the sample `nch` positions before it, modulo the sample width.
template<bool U16, bool HDIFF>
__global__ void __launch_bounds__(256)
tile_relayout_int_kernel(const float* __restrict__ src, int w, int h, int nch, int tile_w,
                         int tile_h, int ntx, int ty0, long long total_quads,
                         void* __restrict__ out)
{
    pdl_prologue();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_quads;
         i += (long long)gridDim.x * blockDim.x) {
        const TileQuad q = tile_quad(i, src, w, h, nch, tile_w, tile_h, ntx, ty0);
        const float4 v   = tile_load4(q, w, nch, q.count);
        unsigned s[4];
        s[0] = U16 ? quant_u16(v.x) : quant_u8(v.x);
        s[1] = U16 ? quant_u16(v.y) : quant_u8(v.y);
        s[2] = U16 ? quant_u16(v.z) : quant_u8(v.z);
        s[3] = U16 ? quant_u16(v.w) : quant_u8(v.w);
        if (HDIFF) {
            unsigned p[4] = { 0u, 0u, 0u, 0u };
            if (q.count >= nch) {
                const float4 u = tile_load4(q, w, nch, q.count - nch);
                p[0] = U16 ? quant_u16(u.x) : quant_u8(u.x);
                p[1] = U16 ? quant_u16(u.y) : quant_u8(u.y);
                p[2] = U16 ? quant_u16(u.z) : quant_u8(u.z);
                p[3] = U16 ? quant_u16(u.w) : quant_u8(u.w);
            } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int a = q.count + k - nch;  // < 0: first pixel of the row, kept as is
                    if (a >= 0) {
                        const long long first = (long long)q.x0 * nch + a;
                        float f = 0.0f;
                        if (q.inside && first < (long long)w * nch)
                            f = __ldg(q.row + first);
                        p[k] = U16 ? quant_u16(f) : quant_u8(f);
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < 4; ++k)
                s[k] = (s[k] - p[k]) & (U16 ? 0xffffu : 0xffu);
        }
        if (U16) {
            uint2 o;
            o.x = s[0] | (s[1] << 16);
            o.y = s[2] | (s[3] << 16);
            reinterpret_cast<uint2*>(out)[i] = o;
        } else {
            reinterpret_cast<unsigned*>(out)[i] = s[0] | (s[1] << 8) | (s[2] << 16) | (s[3] << 24);
        }
    }
}

}  // namespace

void
launch_tile_relayout_int(const float* level, int w, int h, int nch, int tile_w, int tile_h,
                         int ty0, int ty1, int out_u16, int hdiff, void* out, void* stream)
{
    const int ntx         = (w + tile_w - 1) / tile_w;
    const long long total = (long long)(ty1 - ty0) * ntx * tile_w * tile_h * nch / 4;
    if (total <= 0)
        return;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    void (*k)(const float*, int, int, int, int, int, int, int, long long, void*)
        = out_u16 ? (hdiff ? tile_relayout_int_kernel<true, true>
                           : tile_relayout_int_kernel<true, false>)
                  : (hdiff ? tile_relayout_int_kernel<false, true>
                           : tile_relayout_int_kernel<false, false>);
    launch_pdl(k, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, level, w, h, nch, tile_w,
               tile_h, ntx, ty0, total, out);
}

void
launch_tile_relayout_fp_predictor(const float* level, int w, int h, int nch, int tile_w,
                                  int tile_h, int ty0, int ty1, int out_half, void* out,
                                  void* stream)
{
    const int ntx         = (w + tile_w - 1) / tile_w;
    const long long total = (long long)(ty1 - ty0) * ntx * tile_w * tile_h * nch / 4;
    if (total <= 0)
        return;
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    if (out_half)
        launch_pdl(tile_relayout_fp_predictor_kernel<true>, dim3(blocks), dim3(256), 0,
                   (cudaStream_t)stream, level, w, h, nch, tile_w, tile_h, ntx, ty0, total,
                   (unsigned*)out);
    else
        launch_pdl(tile_relayout_fp_predictor_kernel<false>, dim3(blocks), dim3(256), 0,
                   (cudaStream_t)stream, level, w, h, nch, tile_w, tile_h, ntx, ty0, total,
                   (unsigned*)out);
}

void
launch_tile_relayout(const float* level, int w, int h, int nch, int tile_w, int tile_h,
                     int ty0, int ty1, int out_half, void* out, void* stream)
{
    const int ntx         = (w + tile_w - 1) / tile_w;
    const long long total = (long long)(ty1 - ty0) * ntx * tile_w * tile_h * nch;
    if (total <= 0)
        return;
    if ((tile_w * nch) % 4 == 0 && (reinterpret_cast<unsigned long long>(out) & 15ull) == 0) {
        const long long quads = total / 4;
        const int blocks = (int)std::min<long long>((quads + 255) / 256, 148LL * 32);
        if (out_half)
            launch_pdl(tile_relayout_vec_kernel<true>, dim3(blocks), dim3(256), 0,
                       (cudaStream_t)stream, level, w, h, nch, tile_w, tile_h, ntx, ty0, quads, out);
        else
            launch_pdl(tile_relayout_vec_kernel<false>, dim3(blocks), dim3(256), 0,
                       (cudaStream_t)stream, level, w, h, nch, tile_w, tile_h, ntx, ty0, quads, out);
        return;
    }
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    if (out_half)
        launch_pdl(tile_relayout_kernel<true>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                   level, w, h, nch, tile_w, tile_h, ntx, ty0, total, out);
    else
        launch_pdl(tile_relayout_kernel<false>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                   level, w, h, nch, tile_w, tile_h, ntx, ty0, total, out);
}

}  // namespace b2r
