// wide_kernel.cu -- separable resize with windows of ANY width (thumbnail-scale
// minification: 8K -> 128 x 72 with lanczos3 is 361 taps per axis).
//
// resize_ (imagebufalgo_xform.cpp:595-826) is O(xtaps * ytaps) per pixel; the
// exact kernel follows it and takes 20-80 ms for such a thumbnail.  The fused
// kernels are two-pass but keep their H window in shared memory / registers
// (<= 128 taps).  Here the two passes are two plain kernels over the
// contiguous-run gather tables of axis_plan.h (WrapClamp addressing resolved,
// zero weights and black taps dropped, clamped duplicates merged):
//
//   V: tmp[y][x][c]  = sum_j wy[y][j] * src[first_y[y] + j][x][c]     (float)
//   H: dst[y][x][c]  = quantise( sum_i wx[x][i] * tmp[y][first_x[x] + i][c] )
//
// The source is read ~ytaps / yratio^-1 times through L2 (each source row
// contributes to a handful of destination rows), the intermediate is
// dst_h x src_w floats.  Same accumulation model as the fused kernels
// (w_y and w_x applied in two steps instead of as one product): results within
// 1e-5 / 1 LSB of the reference, not bit-exact.
#include "dev_convert.cuh"
#include "kernels.h"

#include <cuda_fp16.h>

namespace b2r {

namespace {

__device__ __forceinline__ float
wide_load(const unsigned char* p, int bt)
{
    switch (bt) {
    case BT_UINT8: return __fmul_rn(float(*p), 1.0f / 255.0f);
    case BT_UINT16: return __fmul_rn(float(*(const unsigned short*)p), 1.0f / 65535.0f);
    case BT_HALF: return __half2float(*(const __half*)p);
    default: return *(const float*)p;
    }
}

// one thread = one (dst row, source column, group of <= 4 channels)
__global__ void __launch_bounds__(256)
resize_wide_v_kernel(const WideParams p)
{
    pdl_prologue();
    const int x  = blockIdx.x * blockDim.x + threadIdx.x;  // column of the needed range
    const int y  = blockIdx.y;                             // dst row of the ROI
    const int c0 = blockIdx.z * 4;
    if (x >= p.tmp_w)
        return;
    const int nc    = min(4, p.nch - c0);
    const int first = p.yfirst[y], count = p.ycount[y];
    const float* w  = p.yw + p.ywoff[y];
    const unsigned char* mixed = p.ymixed ? p.ymixed + p.ywoff[y] : nullptr;
    const int ses   = p.src_es;
    const unsigned char* sp = p.src + (long long)first * p.src_ystride
                              + (long long)(p.x0 + x) * p.src_xstride + (long long)c0 * ses;
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
    const bool fp = p.srctype == BT_FLOAT || p.srctype == BT_HALF;
    for (int j = 0; j < count; ++j, sp += p.src_ystride) {
        const float wj = __ldg(w + j);
        float v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c)
            v[c] = c < nc ? wide_load(sp + c * ses, p.srctype) : 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            acc[c] = fmaf(wj, v[c], acc[c]);
        // a weight that merges clamped taps of both signs: w1*Inf + w2*Inf is NaN
        if (fp && mixed && mixed[j]) {
#pragma unroll
            for (int c = 0; c < 4; ++c)
                acc[c] = fmaf(0.0f, v[c], acc[c]);
        }
    }
    float* t = p.tmp + ((long long)y * p.tmp_w + x) * p.nch + c0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < nc)
            t[c] = acc[c];
}

// one thread = one (dst pixel, group of <= 4 channels)
__global__ void __launch_bounds__(256)
resize_wide_h_kernel(const WideParams p)
{
    pdl_prologue();
    const int x  = blockIdx.x * blockDim.x + threadIdx.x;  // dst column of the ROI
    const int y  = blockIdx.y;
    const int c0 = blockIdx.z * 4;
    if (x >= p.roi_w)
        return;
    const int nc    = min(4, p.nch - c0);
    const int count = p.xcount[x];
    const float* w  = p.xw + p.xwoff[x];
    const unsigned char* mixed = p.xmixed ? p.xmixed + p.xwoff[x] : nullptr;
    const float* t  = p.tmp + ((long long)y * p.tmp_w + (p.xfirst[x] - p.x0)) * p.nch + c0;
    const bool fp   = p.srctype == BT_FLOAT || p.srctype == BT_HALF;
    float acc[4] = { 0.f, 0.f, 0.f, 0.f };
    // a destination row with no contributing source row is written as zeros
    // (xform.cpp:765-772); so is a column with none
    const bool row_empty = p.ycount[y] == 0;
    if (!row_empty) {
        for (int i = 0; i < count; ++i, t += p.nch) {
            const float wi = __ldg(w + i);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nc)
                    acc[c] = fmaf(wi, t[c], acc[c]);
            if (fp && mixed && mixed[i]) {
#pragma unroll
                for (int c = 0; c < 4; ++c)
                    if (c < nc)
                        acc[c] = fmaf(0.0f, t[c], acc[c]);
            }
        }
    }
    unsigned char* dp = p.dst + (long long)y * p.dst_ystride + (long long)x * p.dst_xstride
                        + (long long)c0 * p.dst_es;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        if (c < nc)
            store_elem(dp + c * p.dst_es, p.dsttype, acc[c]);
}

}  // namespace

void
launch_resize_wide(const WideParams& p, void* stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    const int groups = (p.nch + 3) / 4;
    dim3 gv((p.tmp_w + 255) / 256, p.roi_h, groups);
    launch_pdl(resize_wide_v_kernel, gv, dim3(256), 0, st, p);
    dim3 gh((p.roi_w + 255) / 256, p.roi_h, groups);
    launch_pdl(resize_wide_h_kernel, gh, dim3(256), 0, st, p);
}

}  // namespace b2r
