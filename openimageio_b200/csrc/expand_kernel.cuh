// expand_kernel.cuh -- the fused V->H resampler for UPSIZING in y.
//
// When the destination has more rows than the source, few source rows feed
// many destination rows and the output side dominates: C3 (1920x1080 RGB
// uint8 -> 7680x4320) reads 6 MB and writes 100 MB.  The strip/ring layout of
// resize_fused_kernel leaves most of a CTA idle in the V pass there (a 256
// column strip spans ~64 source pixels) and spends its instructions on
// per-tap type conversion and per-element stores.  This variant keeps the
// same CTA = column strip x row band decomposition and the same planar V-row
// layout, with
//
//  * strips up to 1024 dst columns wide, so the source span of a strip fills
//    the CTA (thread t owns 4 consecutive source elements, as in the ring
//    kernel);
//  * source rows streamed through a per-thread cp.async FIFO and converted to
//    float ONCE into a per-thread ring of EXPAND_RING rows in shared memory;
//    every destination row gathers its vt taps from that ring (private
//    LDS.128 + 2 FFMA2 per tap, no barrier);
//  * an H pass in which a thread computes a GROUP of four adjacent dst
//    columns from one shared window of HT source pixels (the four footprints
//    nearly coincide when upsizing), all channels packed into FFMA2 pairs;
//  * quantisation without the F2I-side clamps, and the group's 4*NCH
//    elements packed into 32- or 128-bit stores.
//
// Arithmetic per output is the same as in resize_fused_kernel (V then H, fp32
// FMA in tap order, zero weights multiplied and screened).
#pragma once

#include "fused_kernel.cuh"

namespace b2r {

// convert_type<float,uint>(v) (fmath.h:753-764, 1009-1031) for v*mx computed
// with a separate multiply: s = v*mx; s += 0.5 (s >= 0) ; clamp ; truncate.
// For s < 0 the reference adds -0.5 and clamps to 0: s + 0.5 is below 0.5
// there and truncates (or saturates) to 0 as well, so one FADD serves both
// signs; the float->u32 conversion saturates negatives and NaN to 0 and the
// upper clamp is an integer min.
__device__ __forceinline__ unsigned
quant_fast(float v, float mx, unsigned imx)
{
    float s = __fadd_rn(__fmul_rn(v, mx), 0.5f);
    unsigned u;
    asm("cvt.rzi.u32.f32 %0, %1;" : "=r"(u) : "f"(s));
    return min(u, imx);
}

__device__ __forceinline__ float2
ffma2(float2 a, float2 b, float2 c)
{
    return __ffma2_rn(a, b, c);
}

// Accumulators of one row of a group (4 columns x NCH channels) as FFMA2
// pairs, and the order they come out in.
template<int NCH> struct GroupAcc;
template<> struct GroupAcc<4> {
    float2 a[8];  // a[2k] = (x,y) of column k, a[2k+1] = (z,w)
    __device__ __forceinline__ void tap(const float2 (&wd)[4], float4, const float4& v)
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[2 * k]     = ffma2(wd[k], make_float2(v.x, v.y), a[2 * k]);
            a[2 * k + 1] = ffma2(wd[k], make_float2(v.z, v.w), a[2 * k + 1]);
        }
    }
    __device__ __forceinline__ void get(float (&o)[16]) const
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[4 * k]     = a[2 * k].x;
            o[4 * k + 1] = a[2 * k].y;
            o[4 * k + 2] = a[2 * k + 1].x;
            o[4 * k + 3] = a[2 * k + 1].y;
        }
    }
};
template<> struct GroupAcc<3> {
    float2 a[6];  // a[k] = (x,y) of column k; a[4] = (z0,z1), a[5] = (z2,z3)
    __device__ __forceinline__ void tap(const float2 (&wd)[4], float4 w, const float4& v)
    {
        const float2 xy = make_float2(v.x, v.y), zz = make_float2(v.z, v.z);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            a[k] = ffma2(wd[k], xy, a[k]);
        a[4] = ffma2(make_float2(w.x, w.y), zz, a[4]);
        a[5] = ffma2(make_float2(w.z, w.w), zz, a[5]);
    }
    __device__ __forceinline__ void get(float (&o)[12]) const
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[3 * k]     = a[k].x;
            o[3 * k + 1] = a[k].y;
        }
        o[2]  = a[4].x;
        o[5]  = a[4].y;
        o[8]  = a[5].x;
        o[11] = a[5].y;
    }
};
template<> struct GroupAcc<2> {
    float2 a[4];  // a[k] = (x,y) of column k
    __device__ __forceinline__ void tap(const float2 (&wd)[4], float4, const float4& v)
    {
        const float2 xy = make_float2(v.x, v.y);
#pragma unroll
        for (int k = 0; k < 4; ++k)
            a[k] = ffma2(wd[k], xy, a[k]);
    }
    __device__ __forceinline__ void get(float (&o)[8]) const
    {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            o[2 * k]     = a[k].x;
            o[2 * k + 1] = a[k].y;
        }
    }
};
template<> struct GroupAcc<1> {
    float2 a[2];  // (x0,x1), (x2,x3)
    __device__ __forceinline__ void tap(const float2 (&)[4], float4 w, const float4& v)
    {
        const float2 xx = make_float2(v.x, v.x);
        a[0] = ffma2(make_float2(w.x, w.y), xx, a[0]);
        a[1] = ffma2(make_float2(w.z, w.w), xx, a[1]);
    }
    __device__ __forceinline__ void get(float (&o)[4]) const
    {
        o[0] = a[0].x;
        o[1] = a[0].y;
        o[2] = a[1].x;
        o[3] = a[1].y;
    }
};

// Quantise the N = 4*NCH elements of a group to the dst type and store them.
// dp: address of the group's first element.  `align`: 16 / 4 / 0 as
// guaranteed by the host for every group address (FusedParams::dst_align).
template<int N>
__device__ __forceinline__ void
store_group(unsigned char* dp, int dsttype, int align, int nvalid, const float (&o)[N])
{
    if (nvalid == N && align >= 4) {
        if (dsttype == BT_FLOAT) {
            if (align == 16) {
#pragma unroll
                for (int i = 0; i < N / 4; ++i)
                    reinterpret_cast<float4*>(dp)[i]
                        = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < N; ++i)
                    reinterpret_cast<float*>(dp)[i] = o[i];
            }
        } else if (dsttype == BT_UINT8) {
            unsigned w[N / 4];
#pragma unroll
            for (int i = 0; i < N / 4; ++i) {
                unsigned lo = quant_fast(o[4 * i], 255.0f, 255u)
                              | (quant_fast(o[4 * i + 1], 255.0f, 255u) << 8);
                unsigned hi = quant_fast(o[4 * i + 2], 255.0f, 255u)
                              | (quant_fast(o[4 * i + 3], 255.0f, 255u) << 8);
                w[i] = lo | (hi << 16);
            }
            if constexpr (N == 16) {
                if (align == 16) {
                    *reinterpret_cast<uint4*>(dp) = make_uint4(w[0], w[1], w[2], w[3]);
                    return;
                }
            }
            {
#pragma unroll
                for (int i = 0; i < N / 4; ++i)
                    reinterpret_cast<unsigned*>(dp)[i] = w[i];
            }
        } else {
            unsigned w[N / 2];
#pragma unroll
            for (int i = 0; i < N / 2; ++i) {
                if (dsttype == BT_HALF) {
                    __half2 h = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
                    w[i]      = *reinterpret_cast<unsigned*>(&h);
                } else {
                    w[i] = quant_fast(o[2 * i], 65535.0f, 65535u)
                           | (quant_fast(o[2 * i + 1], 65535.0f, 65535u) << 16);
                }
            }
            if (N % 8 == 0 && align == 16) {
#pragma unroll
                for (int i = 0; i < N / 8; ++i)
                    reinterpret_cast<uint4*>(dp)[i]
                        = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < N / 2; ++i)
                    reinterpret_cast<unsigned*>(dp)[i] = w[i];
            }
        }
    } else {
        const int des = (dsttype == BT_UINT8) ? 1 : (dsttype == BT_FLOAT ? 4 : 2);
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < nvalid)
                store_elem(dp + i * des, dsttype, o[i]);
    }
}

// H pass over the nb buffered rows (dst rows dyb .. dyb+nb-1 of the ROI).
// hws: this group's weights, hws[j * EXPAND_GROUPS] = (w of columns 0..3 for
// window tap j).  hb[]: byte offsets of window taps 0..3 in V row 0.
template<int NCH, int HT>
__device__ __forceinline__ void
expand_hpass(const FusedParams& p, const unsigned char* vrows, int nb, int dyb, int dx0,
             int ncols, const int (&hb)[4], const float4* hws)
{
    constexpr int NR = (NCH == 4) ? 2 : 4;  // rows filtered together
    constexpr int N  = 4 * NCH;
    const int des    = (p.dsttype == BT_UINT8) ? 1 : (p.dsttype == BT_FLOAT ? 4 : 2);
    unsigned char* dp = p.dst + (long long)dyb * p.dst_ystride + (long long)dx0 * (NCH * des);
    bool bad = false;
#pragma unroll 1
    for (int rb = 0; rb < nb; rb += NR) {
        GroupAcc<NCH> acc[NR];
#pragma unroll
        for (int k = 0; k < NR; ++k)
#pragma unroll
            for (int i = 0; i < int(sizeof(acc[k].a) / sizeof(float2)); ++i)
                acc[k].a[i] = make_float2(0.f, 0.f);
        const unsigned char* vr = vrows + rb * FUSED_PITCH_B;
#pragma unroll
        for (int j = 0; j < HT; ++j) {
            const float4 w      = hws[j * EXPAND_GROUPS];
            const float2 wd[4] = { make_float2(w.x, w.x), make_float2(w.y, w.y),
                                   make_float2(w.z, w.z), make_float2(w.w, w.w) };
#pragma unroll
            for (int k = 0; k < NR; ++k) {
                // rows past nb hold stale data: computed and discarded
                const float4 v = *reinterpret_cast<const float4*>(
                    vr + k * FUSED_PITCH_B + hb[j & 3] + (j >> 2) * 16);
                acc[k].tap(wd, w, v);
            }
        }
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            if (rb + k < nb) {
                float o[N];
                acc[k].get(o);
                if (p.check_nonfinite) {
                    float t = 0.0f;
#pragma unroll
                    for (int i = 0; i < N; ++i)
                        t += o[i];  // Inf/NaN in any element survives the sum
                    bad |= !(fabsf(t) <= 3.402823466e+38f);
                }
                store_group<N>(dp + (long long)(rb + k) * p.dst_ystride, p.dsttype,
                               p.dst_align, ncols * NCH, o);
            }
        }
    }
    if (bad)
        *p.nonfinite_flag = p.nonfinite_epoch;
}

// Shared memory carve-up (bytes):
//   vrows  EXPAND_BATCH x FUSED_PITCH_B             V-filtered rows
//   hws    HT x EXPAND_GROUPS x 16                  H weights, [tap][group]
//   ring   EXPAND_RING x FUSED_THREADS x 16         converted source rows
//   fifo   EXPAND_FIFO x FUSED_THREADS x 4*es       raw source rows in flight
//   vws    max_band_dy x vt x 4                     V weights of the band
//   vfs    max_band_dy x 4                          first source row per dst row
template<int NCH, int HT>
__global__ void __launch_bounds__(FUSED_THREADS, 2)
resize_expand_kernel(const FusedParams p, int max_band_dy)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int BATCH = EXPAND_BATCH;
    const int es        = (p.srctype == BT_UINT8) ? 1 : (p.srctype == BT_FLOAT ? 4 : 2);
    const int TB        = 4 * es;              // raw bytes per thread per row
    const int SLOT      = FUSED_THREADS * TB;  // bytes between FIFO rows
    unsigned char* vrows = smem_raw;
    float4* hws_all      = reinterpret_cast<float4*>(smem_raw + BATCH * FUSED_PITCH_B);
    float4* ring         = hws_all + HT * EXPAND_GROUPS;
    unsigned char* fifo  = reinterpret_cast<unsigned char*>(ring + EXPAND_RING * FUSED_THREADS);
    float* vws           = reinterpret_cast<float*>(fifo + EXPAND_FIFO * SLOT);
    int* vfs             = reinterpret_cast<int*>(vws + (size_t)max_band_dy * p.vt);

    const int tid     = threadIdx.x;
    const Strip strip = p.strips[blockIdx.x];
    const Band band   = p.bands[blockIdx.y];
    const int vt      = p.vt;

    // ---- H-pass constants: this thread's group of four dst columns -------
    const int ncols = min(4, strip.ndx - 4 * tid);  // <= 0: no group
    const int gg    = (strip.dx0 >> 2) + tid;       // group index in the ROI
    int hb[4]       = { 0, 0, 0, 0 };
    {
        const float4* gw = reinterpret_cast<const float4*>(p.hw);
#pragma unroll
        for (int j = 0; j < HT; ++j)
            hws_all[j * EXPAND_GROUPS + tid]
                = (ncols > 0) ? gw[(size_t)j * p.ngroups + gg] : make_float4(0.f, 0.f, 0.f, 0.f);
        if (ncols > 0) {
            const int q = p.hfirst[gg] - strip.e0 / NCH;  // window start in the strip
#pragma unroll
            for (int k = 0; k < 4; ++k)
                hb[k] = planar_byte(q + k);
        }
    }
    const float4* hws = hws_all + tid;

    // ---- V-pass constants: this thread's 4 source elements ---------------
    const bool vactive  = tid < strip.nvec;
    const int e         = strip.e0 + 4 * tid;
    const int row_elems = p.src_w * NCH;
    VPark<NCH> park;
    park.init(strip.e0, tid);
    if (NCH < 4) {
        // padded pixels: the unused lanes are never written, they must not
        // hold Inf/NaN bit patterns
        for (int i = tid; i < BATCH * (FUSED_PITCH_B / 16); i += FUSED_THREADS)
            reinterpret_cast<uint4*>(vrows)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // stage the band's V weights and (band-relative) first rows
    {
        const float* gw = p.vw + (size_t)band.dy0 * vt;
        for (int i = tid; i < band.ndy * vt; i += FUSED_THREADS)
            vws[i] = gw[i];
        for (int i = tid; i < band.ndy; i += FUSED_THREADS)
            vfs[i] = p.vfirst[band.dy0 + i] - band.r0;
    }

    // cp.async needs aligned global addresses and groups inside the row (see
    // resize_fused_kernel); otherwise rows are loaded synchronously
    const bool async_ok = p.src_vec_ok != 0
                          && ((row_elems & 3) == 0 || strip.e0 + 4 * strip.nvec <= row_elems);
    const bool vfetch = vactive && e + 4 <= row_elems;
    const unsigned fifo_s = smem_u32(fifo) + tid * TB;
    if (async_ok && !vfetch) {
        for (int q = 0; q < EXPAND_FIFO; ++q)
            for (int b = 0; b < TB; b += 4)
                *reinterpret_cast<unsigned*>(fifo + q * SLOT + tid * TB + b) = 0u;
    }
    const unsigned char* gnext = p.src + (size_t)e * es + (long long)band.r0 * p.src_ystride;
    int rnext = 0;  // next row to fetch, band-relative
    int slot_next = 0, slot_cur = 0;
    auto issue_next = [&]() {
        if (async_ok) {
            if (vfetch && rnext < band.nrows) {
                if (es == 4)
                    cp_async_copy<16>(fifo_s + slot_next, gnext);
                else if (es == 2)
                    cp_async_copy<8>(fifo_s + slot_next, gnext);
                else
                    cp_async_copy<4>(fifo_s + slot_next, gnext);
            }
            cp_async_commit();
            gnext += p.src_ystride;
            slot_next = (slot_next + SLOT == EXPAND_FIFO * SLOT) ? 0 : slot_next + SLOT;
        }
        ++rnext;
    };
    // next source row, converted to float
    int loaded = 0;  // rows converted so far (band-relative index of the next)
    auto pop_row = [&]() -> float4 {
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);
        if (async_ok) {
            cp_async_wait<EXPAND_FIFO - 1>();
            if (es == 4)
                raw = lds_v4(fifo_s + slot_cur);
            else if (es == 2) {
                uint2 t = lds_v2(fifo_s + slot_cur);
                raw     = make_uint4(t.x, t.y, 0u, 0u);
            } else
                raw.x = lds_u32(fifo_s + slot_cur);
            slot_cur = (slot_cur + SLOT == EXPAND_FIFO * SLOT) ? 0 : slot_cur + SLOT;
        } else if (vactive) {
            const unsigned char* rowp = p.src + (long long)(band.r0 + loaded) * p.src_ystride;
            if (p.srctype == BT_FLOAT)
                raw = load_raw4<ST_FLOAT>(rowp, e, row_elems, false);
            else if (p.srctype == BT_HALF)
                raw = load_raw4<ST_HALF>(rowp, e, row_elems, false);
            else if (p.srctype == BT_UINT16)
                raw = load_raw4<ST_U16>(rowp, e, row_elems, false);
            else
                raw = load_raw4<ST_U8>(rowp, e, row_elems, false);
        }
        if (p.srctype == BT_FLOAT)
            return convert_raw4<ST_FLOAT>(raw);
        if (p.srctype == BT_HALF)
            return convert_raw4<ST_HALF>(raw);
        if (p.srctype == BT_UINT16)
            return convert_raw4<ST_U16>(raw);
        return convert_raw4<ST_U8>(raw);
    };

    if (async_ok) {
#pragma unroll
        for (int q = 0; q < EXPAND_FIFO; ++q)
            issue_next();
    }
    __syncthreads();  // vws / vfs / hws / cleared vrows visible

    float4* myring = ring + tid;
    const int dy_end = band.dy0 + band.ndy;
    for (int dyb = band.dy0; dyb < dy_end; dyb += BATCH) {
        const int nb = min(BATCH, dy_end - dyb);
        for (int row = 0; row < nb; ++row) {
            const int li = dyb - band.dy0 + row;
            const int f  = vfs[li];
            while (loaded < f + vt) {
                myring[(loaded & (EXPAND_RING - 1)) * FUSED_THREADS] = pop_row();
                issue_next();  // refills the FIFO slot just consumed
                ++loaded;
            }
            const float* w = vws + li * vt;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int j = 0; j < vt; ++j)
                fma4(a, w[j], myring[((f + j) & (EXPAND_RING - 1)) * FUSED_THREADS]);
            if (vactive)
                park.store(vrows + (size_t)row * FUSED_PITCH_B, a);
        }
        __syncthreads();
        if (ncols > 0)
            expand_hpass<NCH, HT>(p, vrows, nb, dyb, strip.dx0 + 4 * tid, ncols, hb, hws);
        __syncthreads();
    }
    if (async_ok)
        cp_async_wait<0>();
}

}  // namespace b2r
