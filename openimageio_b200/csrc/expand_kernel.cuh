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
//    float ONCE into a sliding window of VT rows held in registers; every
//    destination row is 2*VT FFMA2 on that window (consecutive dst rows share
//    it: the window slides once per ~4 dst rows at 4x);
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

// Quantise the N = 4*NCH elements of a group to the dst type DT and store
// them.  dp: address of the group's first element.  `align`: 16 / 4 / 0 as
// guaranteed by the host for every group address (FusedParams::dst_align).
template<int N, int DT>
__device__ __forceinline__ void
store_group(unsigned char* dp, int align, int nvalid, const float (&o)[N])
{
    if (nvalid == N && align >= 4) {
        if (DT == BT_FLOAT) {
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
        } else if (DT == BT_UINT8) {
            // clamp + one packed FMA per pair, no cvt (dev_convert.cuh: quant2_magic)
            unsigned w[N / 4];
#pragma unroll
            for (int i = 0; i < N / 4; ++i) {
                const float2 a = quant2_magic_u8(o[4 * i], o[4 * i + 1]);
                const float2 b = quant2_magic_u8(o[4 * i + 2], o[4 * i + 3]);
                unsigned lo = __byte_perm(__float_as_uint(a.x), __float_as_uint(a.y), 0x0040);
                unsigned hi = __byte_perm(__float_as_uint(b.x), __float_as_uint(b.y), 0x0040);
                w[i]        = __byte_perm(lo, hi, 0x5410);
            }
            if constexpr (N == 16) {
                if (align == 16) {
                    *reinterpret_cast<uint4*>(dp) = make_uint4(w[0], w[1], w[2], w[3]);
                    return;
                }
            }
#pragma unroll
            for (int i = 0; i < N / 4; ++i)
                reinterpret_cast<unsigned*>(dp)[i] = w[i];
        } else {
            unsigned w[N / 2];
#pragma unroll
            for (int i = 0; i < N / 2; ++i) {
                if (DT == BT_HALF) {
                    __half2 h = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
                    w[i]      = *reinterpret_cast<unsigned*>(&h);
                } else {
                    const float2 a = quant2_magic(o[2 * i], o[2 * i + 1], 65535.0f);
                    w[i] = __byte_perm(__float_as_uint(a.x), __float_as_uint(a.y), 0x5410);
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
        constexpr int des = (DT == BT_UINT8) ? 1 : (DT == BT_FLOAT ? 4 : 2);
#pragma unroll
        for (int i = 0; i < N; ++i)
            if (i < nvalid)
                store_elem(dp + i * des, DT, o[i]);
    }
}

// The NR rows a thread has just filtered: screen, quantise, store.
template<int NCH, int NR, int DT>
__device__ __forceinline__ bool
store_rows(const FusedParams& p, const GroupAcc<NCH> (&acc)[NR], unsigned char* dp,
           int nrows, int ncols, int x0, int y0)
{
    constexpr int N = 4 * NCH;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        if (k < nrows) {
            float o[N];
            acc[k].get(o);
            if (p.check_nonfinite) {
                float t = 0.0f;
#pragma unroll
                for (int i = 0; i < N; ++i)
                    t += o[i];  // Inf/NaN in any element survives the sum
                if (!(fabsf(t) <= 3.402823466e+38f)) {
                    bad = true;
                    // mark the cell(s) of the dirty map this group of four
                    // columns touches: the gated exact kernel re-does only those
                    int* row = p.dirty + ((y0 + k) >> p.dirty_yshift) * p.dirty_xcells;
                    row[x0 >> DIRTY_CELL_SHIFT_X]                   = p.nonfinite_epoch;
                    row[(x0 + ncols - 1) >> DIRTY_CELL_SHIFT_X] = p.nonfinite_epoch;
                }
            }
            store_group<N, DT>(dp, p.dst_align, ncols * NCH, o);
        }
        dp += p.dst_ystride;
    }
    return bad;
}

// H pass over the nb buffered rows (dst rows dyb .. dyb+nb-1 of the ROI).
// hws: this group's weights, hws[j * EXPAND_GROUPS] = (w of columns 0..3 for
// window tap j).  hb[j]: byte offset of window tap j in V row 0.
template<int NCH, int HT>
__device__ __forceinline__ void
expand_hpass(const FusedParams& p, const unsigned char* vrows, int nb, int dyb, int dx0,
             int ncols, const int (&hb)[HT], const float4* hws)
{
    constexpr int NR = (NCH >= 3) ? 2 : 4;  // rows filtered together
    const int dt     = p.dsttype;
    const int des    = (dt == BT_UINT8) ? 1 : (dt == BT_FLOAT ? 4 : 2);
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
                const float4 v
                    = *reinterpret_cast<const float4*>(vr + k * FUSED_PITCH_B + hb[j]);
                acc[k].tap(wd, w, v);
            }
        }
        // one type dispatch per NR rows
        const int nrows = nb - rb;
        if (dt == BT_UINT8)
            bad |= store_rows<NCH, NR, BT_UINT8>(p, acc, dp, nrows, ncols, dx0, dyb + rb);
        else if (dt == BT_FLOAT)
            bad |= store_rows<NCH, NR, BT_FLOAT>(p, acc, dp, nrows, ncols, dx0, dyb + rb);
        else if (dt == BT_HALF)
            bad |= store_rows<NCH, NR, BT_HALF>(p, acc, dp, nrows, ncols, dx0, dyb + rb);
        else
            bad |= store_rows<NCH, NR, BT_UINT16>(p, acc, dp, nrows, ncols, dx0, dyb + rb);
        dp += (long long)NR * p.dst_ystride;
    }
    if (bad)
        *p.nonfinite_flag = p.nonfinite_epoch;
}

// Shared memory carve-up (bytes):
//   vrows  EXPAND_BATCH x FUSED_PITCH_B             V-filtered rows
//   hws    HT x EXPAND_GROUPS x 16                  H weights, [tap][group]
//   vws    max_band_dy x VT x 4                     V weights of the band
//   vfs    max_band_dy x 4                          first source row per dst row
//   fifo   EXPAND_FIFO x FUSED_THREADS x 4*es       raw source rows in flight
//
// V rows are laid out linearly (pixel q at q*16) when the windows of adjacent
// groups start an odd number of pixels apart (4x upsizing: 1 pixel -- a
// quarter-warp reads 8 consecutive pixels), else in the four planes of
// resize_fused_kernel (even strides); the host picks (FusedParams::vrow_linear).
template<int NCH, int VT, int HT>
__global__ void __launch_bounds__(FUSED_THREADS, 2)
resize_expand_kernel(const FusedParams p, int max_band_dy)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_prologue();
    constexpr int BATCH = EXPAND_BATCH;
    const int es        = (p.srctype == BT_UINT8) ? 1 : (p.srctype == BT_FLOAT ? 4 : 2);
    const int TB        = 4 * es;              // raw bytes per thread per row
    const int SLOT      = FUSED_THREADS * TB;  // bytes between FIFO rows
    unsigned char* vrows = smem_raw;
    float4* hws_all      = reinterpret_cast<float4*>(smem_raw + BATCH * FUSED_PITCH_B);
    // per-row tables sit before the FIFO so their offsets do not depend on es
    float* vws           = reinterpret_cast<float*>(hws_all + HT * EXPAND_GROUPS);
    int* vfs             = reinterpret_cast<int*>(vws + (((size_t)(max_band_dy + 1) * VT + 3) & ~size_t(3)));
    unsigned char* fifo  = reinterpret_cast<unsigned char*>(vfs + ((max_band_dy + 1 + 3) & ~3));

    const int tid     = threadIdx.x;
    const Strip strip = p.strips[blockIdx.x];
    const Band band   = p.bands[blockIdx.y];
    const bool linear = p.vrow_linear != 0;

    // ---- H-pass constants: this thread's group of four dst columns -------
    const int ncols = min(4, strip.ndx - 4 * tid);  // <= 0: no group
    const int gg    = (strip.dx0 >> 2) + tid;       // group index in the ROI
    int hb[HT];
    {
        // tables are staged with cp.async (one group, committed before the
        // first source rows are requested) so their latency overlaps
        const float4* gw = reinterpret_cast<const float4*>(p.hw);
#pragma unroll
        for (int j = 0; j < HT; ++j) {
            if (ncols > 0)
                cp_async_copy<16>(smem_u32(hws_all + j * EXPAND_GROUPS + tid),
                                  gw + (size_t)j * p.ngroups + gg);
            else
                hws_all[j * EXPAND_GROUPS + tid] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    int hfirst_v = 0;
    if (ncols > 0)
        asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(hfirst_v) : "l"(p.hfirst + gg));
    const float4* hws = hws_all + tid;

    // ---- V-pass constants: this thread's 4 source elements ---------------
    const bool vactive  = tid < strip.nvec;
    const int e         = strip.e0 + 4 * tid;
    const int row_elems = p.src_w * NCH;
    VPark<NCH> park;
    park.init(strip.e0, tid, linear);
    // 1-3 channels: the four scalar park stores of a warp pile up on a few
    // banks (6-way for RGB, linear rows) when every thread stores element k in
    // instruction k.  Rotating the thread's elements by (tid/4)%4 -- once per
    // source row, at conversion -- brings that down to the 2-way floor.
    const int rot = (NCH < 4 && (NCH == 2 || linear)) ? (tid >> 2) & 3 : 0;
    if (NCH < 4 && rot) {
        int o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            o[k] = park.off[NCH == 4 ? 0 : k];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int sel = (k + rot) & 3;
            park.off[NCH == 4 ? 0 : k] = sel == 0 ? o[0] : sel == 1 ? o[1] : sel == 2 ? o[2] : o[3];
        }
    }
    if (NCH < 4) {
        // padded pixels: the unused lanes are never written, they must not
        // hold Inf/NaN bit patterns
        for (int i = tid; i < BATCH * (FUSED_PITCH_B / 16); i += FUSED_THREADS)
            reinterpret_cast<uint4*>(vrows)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // stage the band's V weights and first rows (VT is even: 8-byte copies)
    {
        const float* gw = p.vw + (size_t)band.dy0 * VT;
        for (int i = tid; i < band.ndy * (VT / 2); i += FUSED_THREADS)
            cp_async_copy<8>(smem_u32(vws + 2 * i), gw + 2 * i);
        for (int i = tid; i < band.ndy; i += FUSED_THREADS)
            cp_async_copy<4>(smem_u32(vfs + i), p.vfirst + band.dy0 + i);
        cp_async_commit();
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
    int loaded = band.r0;  // next source row to consume
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
            const unsigned char* rowp = p.src + (long long)loaded * p.src_ystride;
            if (p.srctype == BT_FLOAT)
                raw = load_raw4<ST_FLOAT>(rowp, e, row_elems, false);
            else if (p.srctype == BT_HALF)
                raw = load_raw4<ST_HALF>(rowp, e, row_elems, false);
            else if (p.srctype == BT_UINT16)
                raw = load_raw4<ST_U16>(rowp, e, row_elems, false);
            else
                raw = load_raw4<ST_U8>(rowp, e, row_elems, false);
        }
        float4 v;
        if (p.srctype == BT_FLOAT)
            v = convert_raw4<ST_FLOAT>(raw);
        else if (p.srctype == BT_HALF)
            v = convert_raw4<ST_HALF>(raw);
        else if (p.srctype == BT_UINT16)
            v = convert_raw4<ST_U16>(raw);
        else
            v = convert_raw4<ST_U8>(raw);
        if (NCH < 4) {
            if (rot & 1)
                v = make_float4(v.y, v.z, v.w, v.x);
            if (rot & 2)
                v = make_float4(v.z, v.w, v.x, v.y);
        }
        return v;
    };

    if (async_ok) {
#pragma unroll
        for (int q = 0; q < EXPAND_FIFO; ++q)
            issue_next();
        cp_async_wait<EXPAND_FIFO>();  // the table group; source rows stay in flight
    } else {
        cp_async_wait<0>();
    }
    {
        const int q = hfirst_v - strip.e0 / NCH;  // window start in the strip
#pragma unroll
        for (int j = 0; j < HT; ++j)
            hb[j] = linear ? (q + j) * 16 : planar_byte(q + j);
    }
    __syncthreads();  // vws / vfs / hws / cleared vrows visible

    // the VT source rows under the current dst row, converted, in registers:
    // win[j] = row (loaded - VT + j) of the band
    float4 win[VT];
#pragma unroll
    for (int j = 0; j < VT; ++j)
        win[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int dy_end = band.dy0 + band.ndy;
    // walking shared-window addresses of the current row's weights / first row
    // (kept opaque so they stay in registers instead of being recomputed)
    unsigned vw_s, vf_s;
    asm volatile("mov.u32 %0, %1;" : "=r"(vw_s) : "r"(smem_u32(vws)));
    asm volatile("mov.u32 %0, %1;" : "=r"(vf_s) : "r"(smem_u32(vfs)));
    // a row's first-source-row entry is fetched one row ahead (the loop
    // branches on it; one entry past the band is read and unused)
    auto load_w = [&](float (&w)[VT]) {
        if (VT % 4 == 0) {
#pragma unroll
            for (int j = 0; j < VT / 4; ++j) {
                uint4 t = lds_v4(vw_s + 16 * j);
                w[4 * j]     = __uint_as_float(t.x);
                w[4 * j + 1] = __uint_as_float(t.y);
                w[4 * j + 2] = __uint_as_float(t.z);
                w[4 * j + 3] = __uint_as_float(t.w);
            }
        } else {
#pragma unroll
            for (int j = 0; j < VT / 2; ++j) {
                uint2 t = lds_v2(vw_s + 8 * j);
                w[2 * j]     = __uint_as_float(t.x);
                w[2 * j + 1] = __uint_as_float(t.y);
            }
        }
        vw_s += 4 * VT;
    };
    int fn = (int)lds_u32(vf_s);
    vf_s += 4;
    const unsigned vrows_s = smem_u32(vrows);
    for (int dyb = band.dy0; dyb < dy_end; dyb += BATCH) {
        const int nb = min(BATCH, dy_end - dyb);
        for (int row = 0; row < nb; ++row) {
            const int f = fn;
            float w[VT];
            load_w(w);
            fn = (int)lds_u32(vf_s);
            vf_s += 4;
            while (loaded < f + VT) {  // uniform: slide the window down one row
#pragma unroll
                for (int j = 0; j + 1 < VT; ++j)
                    win[j] = win[j + 1];
                win[VT - 1] = pop_row();
                issue_next();  // refills the FIFO slot just consumed
                ++loaded;
            }
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < VT; ++j)
                fma4(a, w[j], win[j]);
            if (vactive)
                park.store_s(vrows_s + row * FUSED_PITCH_B, a);
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
