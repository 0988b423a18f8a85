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

// one thread = one element of the tiled output (coalesced stores; the loads
// are tile_w * nch contiguous floats per tile row)
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
// byte of the re-ordered row has the byte `nch` positions before it subtracted.
// One thread = four consecutive bytes of a predicted tile row (one 32-bit store);
// a byte's two samples come from the level through L1.
template<bool HALF>
__device__ __forceinline__ unsigned
tile_sample_byte(const float* __restrict__ src, int w, int h, int nch, int y, int x0,
                 int count, int plane)
{
    const int x = x0 + count / nch;
    float v     = 0.0f;
    if (y < h && x < w)
        v = __ldg(src + ((long long)y * w + x) * nch + (count % nch));
    if (HALF)
        return ((unsigned)__half_as_ushort(__float2half_rn(v)) >> (8 * (1 - plane))) & 0xffu;
    return (__float_as_uint(v) >> (8 * (3 - plane))) & 0xffu;
}

template<bool HALF>
__global__ void __launch_bounds__(256)
tile_relayout_fp_predictor_kernel(const float* __restrict__ src, int w, int h, int nch,
                                  int tile_w, int tile_h, int ntx, int ty0, long long total_words,
                                  unsigned* __restrict__ out)
{
    pdl_prologue();
    const int bps       = HALF ? 2 : 4;
    const int wc        = tile_w * nch;  // samples of a tile row
    const int row_words = wc * bps / 4;  // tile_w is a multiple of 16
    const long long tile_words = (long long)row_words * tile_h;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_words;
         i += (long long)gridDim.x * blockDim.x) {
        const long long t = i / tile_words;
        const int e       = int(i - t * tile_words);
        const int yy      = e / row_words;
        const int b0      = (e - yy * row_words) * 4;  // first byte of the predicted row
        const int y       = (ty0 + int(t / ntx)) * tile_h + yy;
        const int x0      = int(t % ntx) * tile_w;
        unsigned word     = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int b = b0 + k;
            unsigned v  = tile_sample_byte<HALF>(src, w, h, nch, y, x0, b % wc, b / wc);
            if (b >= nch) {
                const int a = b - nch;
                v -= tile_sample_byte<HALF>(src, w, h, nch, y, x0, a % wc, a / wc);
            }
            word |= (v & 0xffu) << (8 * k);
        }
        out[i] = word;
    }
}

}  // namespace

void
launch_tile_relayout_fp_predictor(const float* level, int w, int h, int nch, int tile_w,
                                  int tile_h, int ty0, int ty1, int out_half, void* out,
                                  void* stream)
{
    const int ntx         = (w + tile_w - 1) / tile_w;
    const long long total = (long long)(ty1 - ty0) * ntx * tile_w * tile_h * nch
                            * (out_half ? 2 : 4) / 4;
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
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148LL * 32);
    if (out_half)
        launch_pdl(tile_relayout_kernel<true>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                   level, w, h, nch, tile_w, tile_h, ntx, ty0, total, out);
    else
        launch_pdl(tile_relayout_kernel<false>, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
                   level, w, h, nch, tile_w, tile_h, ntx, ty0, total, out);
}

}  // namespace b2r
