// fused_kernel.cuh -- the hot path: one kernel, vertical pass then horizontal
// pass, the intermediate never leaves the SM.
//
// One CTA = one column strip x one row band of the destination ROI.
//
//  V pass  Thread t owns 4 consecutive source elements (one RGBA float pixel,
//          or 4 flat elements of an N-channel row) of the strip.  Source rows
//          stream HBM -> shared memory through a per-thread private cp.async
//          FIFO (LDGSTS, 16 B per thread per row, FUSED_FIFO rows in flight,
//          no CTA barrier because a thread only reads back what it fetched).
//          "Ring" mode (downsizing): every source row is read ONCE and
//          scattered into the VK destination rows that are open at that time
//          (VK register accumulators, slot = dst row mod VK, weights per
//          source row staged in shared memory).  "Gather" mode (upsizing):
//          each destination row gathers its few source rows through L1.
//  H pass  FUSED_BATCH V-filtered rows sit in shared memory; thread (column,
//          row group) gathers HT taps with weights it keeps in registers for
//          the whole band, quantises to the destination type and stores.
//
// Zero weights are multiplied rather than skipped (the reference skips them,
// imagebufalgo_xform.cpp:749-759); for float/half sources the outputs are
// screened for Inf/NaN and the exact kernel re-does the image if one appears.
#pragma once

#include "dev_convert.cuh"
#include "kernels.h"

namespace b2r {

// source type index used as a template parameter
enum : int { ST_U8 = 0, ST_U16 = 1, ST_HALF = 2, ST_FLOAT = 3 };

template<int ST> struct SrcTraits;
template<> struct SrcTraits<ST_U8> { static constexpr int es = 1, bt = BT_UINT8; };
template<> struct SrcTraits<ST_U16> { static constexpr int es = 2, bt = BT_UINT16; };
template<> struct SrcTraits<ST_HALF> { static constexpr int es = 2, bt = BT_HALF; };
template<> struct SrcTraits<ST_FLOAT> { static constexpr int es = 4, bt = BT_FLOAT; };

// Exact uint -> float for values below 2^23 without the (slow) I2F unit:
// 0x4B000000 | v is the float 8388608 + v; subtracting 8388608 is exact.
__device__ __forceinline__ float
small_uint_to_float(unsigned v)
{
    return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f);
}

// ---- raw 4-element groups ------------------------------------------------
// float -> 4 words, half/uint16 -> 2 words, uint8 -> 1 word.
template<int ST>
__device__ __forceinline__ float4
convert_raw4(uint4 r)
{
    if (ST == ST_FLOAT) {
        return make_float4(__uint_as_float(r.x), __uint_as_float(r.y),
                           __uint_as_float(r.z), __uint_as_float(r.w));
    } else if (ST == ST_HALF) {
        __half2 a = *reinterpret_cast<__half2*>(&r.x);
        __half2 b = *reinterpret_cast<__half2*>(&r.y);
        float2 fa = __half22float2(a), fb = __half22float2(b);
        return make_float4(fa.x, fa.y, fb.x, fb.y);
    } else if (ST == ST_U16) {
        // convert_type<uint16,float> is float(u) * (1/65535) (fmath.h:1009-1031).
        // PRMT builds the float 2^23 + u (0x4B00'0000 | u), then ONE packed FMA
        // per element pair computes (2^23 + u) * s - 2^23 * s: the product
        // 2^23 * s is exact, so the FMA rounds exactly u * s -- the same bits as
        // the multiply, without the subtraction and the shifts / masks.
        const float s    = 1.0f / 65535.0f;
        const float2 ss  = make_float2(s, s);
        const float2 off = make_float2(-8388608.0f * s, -8388608.0f * s);
        const float2 lo  = __ffma2_rn(
            make_float2(__uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7410)),
                        __uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7432))),
            ss, off);
        const float2 hi = __ffma2_rn(
            make_float2(__uint_as_float(__byte_perm(r.y, 0x4B000000u, 0x7410)),
                        __uint_as_float(__byte_perm(r.y, 0x4B000000u, 0x7432))),
            ss, off);
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    } else {
        const float s    = 1.0f / 255.0f;
        const float2 ss  = make_float2(s, s);
        const float2 off = make_float2(-8388608.0f * s, -8388608.0f * s);
        const float2 lo  = __ffma2_rn(
            make_float2(__uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7440)),
                        __uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7441))),
            ss, off);
        const float2 hi = __ffma2_rn(
            make_float2(__uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7442)),
                        __uint_as_float(__byte_perm(r.x, 0x4B000000u, 0x7443))),
            ss, off);
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
}

// Synchronous load of elements [e, e+4) of a row (gather mode and the
// unaligned fallback): vector load when allowed, guarded scalars otherwise.
template<int ST>
__device__ __forceinline__ uint4
load_raw4(const unsigned char* row, int e, int row_elems, bool vec_ok)
{
    constexpr int es = SrcTraits<ST>::es;
    uint4 r = make_uint4(0u, 0u, 0u, 0u);
    if (vec_ok && e + 4 <= row_elems) {
        const void* q = row + (size_t)e * es;
        if (es == 4) {
            r = __ldg((const uint4*)q);
        } else if (es == 2) {
            uint2 t = __ldg((const uint2*)q);
            r.x = t.x;
            r.y = t.y;
        } else {
            r.x = __ldg((const unsigned int*)q);
        }
    } else {
        unsigned int w[4] = { 0u, 0u, 0u, 0u };
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (e + k < row_elems) {
                if (es == 4)
                    w[k] = __ldg((const unsigned int*)(row + (size_t)(e + k) * 4));
                else if (es == 2)
                    w[k] = __ldg((const unsigned short*)(row + (size_t)(e + k) * 2));
                else
                    w[k] = __ldg(row + e + k);
            }
        }
        if (es == 4)
            r = make_uint4(w[0], w[1], w[2], w[3]);
        else if (es == 2) {
            r.x = w[0] | (w[1] << 16);
            r.y = w[2] | (w[3] << 16);
        } else {
            r.x = w[0] | (w[1] << 8) | (w[2] << 16) | (w[3] << 24);
        }
    }
    return r;
}

// ---- cp.async (LDGSTS) ---------------------------------------------------
template<int BYTES>
__device__ __forceinline__ void
cp_async_copy(unsigned smem_addr, const void* gsrc)
{
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr),
                     "l"(gsrc)
                     : "memory");
    else if (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr),
                     "l"(gsrc)
                     : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_addr),
                     "l"(gsrc)
                     : "memory");
}
__device__ __forceinline__ unsigned
smem_u32(const void* p)
{
    return (unsigned)__cvta_generic_to_shared(p);
}
// volatile: the FIFO is written behind the compiler's back by cp.async
__device__ __forceinline__ uint4
lds_v4(unsigned addr)
{
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"(addr));
    return v;
}
__device__ __forceinline__ uint2
lds_v2(unsigned addr)
{
    uint2 v;
    asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned
lds_u32(unsigned addr)
{
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void
cp_async_commit()
{
    asm volatile("cp.async.commit_group;" ::: "memory");
}
template<int N>
__device__ __forceinline__ void
cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// V rows are parked as float4 pixels in four planes by (pixel index mod 4);
// plane p starts 66 slots (2 mod 8) after plane p-1, so planes sit in
// different 16-byte bank groups.  The H pass handles two adjacent dst columns
// per thread, i.e. a quarter-warp reads pixels c, c+4, c+8, ... (2:1
// downsample: one plane, consecutive slots) or c, c+2, ... / c, c+1, ...
// (milder ratios: the planes interleave bank groups) -- conflict free -- and a
// thread's taps are base[j & 3] + immediate.
constexpr int FUSED_PLANE_SLOTS = FUSED_THREADS / 4 + 2;
__device__ __forceinline__ int
planar_byte(int q)
{
    return ((q >> 2) + (q & 3) * FUSED_PLANE_SLOTS) * 16;
}

// Blackwell packed FP32: one FFMA2 instruction does two fused multiply-adds
// on 64-bit register pairs -- half the issue slots of scalar FFMA for the
// same arithmetic (each lane result is the same correctly rounded fma).
__device__ __forceinline__ void
fma4(float4& a, float2 ww, const float4& v)
{
    float2 lo = __ffma2_rn(ww, make_float2(v.x, v.y), make_float2(a.x, a.y));
    float2 hi = __ffma2_rn(ww, make_float2(v.z, v.w), make_float2(a.z, a.w));
    a         = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ void
fma4(float4& a, float w, const float4& v)
{
    fma4(a, make_float2(w, w), v);
}

// Where a V thread parks its 4 filtered elements (group v of the strip) in a
// V row: RGBA -> one planar float4; RGB -> scattered into RGBX planar pixels
// (elements straddle pixels); 1-2 channels -> flat floats.
template<int NCH> struct VPark {
    int off[NCH == 4 ? 1 : 4];
    // linear: pixels in order (16 bytes each) instead of the four planes
    __device__ __forceinline__ void init(int e0, int v, bool linear = false)
    {
        if (NCH == 4) {
            off[0] = linear ? v * 16 : planar_byte(v);
        } else {
            const int px0 = e0 / NCH;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                int E  = e0 + 4 * v + k;
                int px = E / NCH;
                off[NCH == 4 ? 0 : k] = (linear ? (px - px0) * 16 : planar_byte(px - px0))
                                        + (E - NCH * px) * 4;
            }
        }
    }
    // same, `row` given as a 32-bit shared-window address
    __device__ __forceinline__ void store_s(unsigned row, const float4& a) const
    {
        if (NCH == 4) {
            asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + off[0]), "f"(a.x),
                         "f"(a.y), "f"(a.z), "f"(a.w)
                         : "memory");
        } else {
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[0]), "f"(a.x) : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[NCH == 4 ? 0 : 1]), "f"(a.y)
                         : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[NCH == 4 ? 0 : 2]), "f"(a.z)
                         : "memory");
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[NCH == 4 ? 0 : 3]), "f"(a.w)
                         : "memory");
        }
    }
    __device__ __forceinline__ void store(unsigned char* row, const float4& a) const
    {
        if (NCH == 4) {
            *reinterpret_cast<float4*>(row + off[0]) = a;
        } else {
            *reinterpret_cast<float*>(row + off[0])                  = a.x;
            *reinterpret_cast<float*>(row + off[NCH == 4 ? 0 : 1]) = a.y;
            *reinterpret_cast<float*>(row + off[NCH == 4 ? 0 : 2]) = a.z;
            *reinterpret_cast<float*>(row + off[NCH == 4 ? 0 : 3]) = a.w;
        }
    }
};

// ---- H pass over `nb` buffered rows; dst rows dyb .. dyb+nb-1 of the ROI ---
// The row pitch is a compile-time constant so every LDS of the H pass is base
// register + immediate.
constexpr int FUSED_PITCH_B = 4 * FUSED_PLANE_SLOTS * 16;  // bytes per V row
constexpr int FUSED_HPAIRS  = FUSED_HCOLS / 2;             // column pairs per strip
constexpr int FUSED_HROWG_MAX = FUSED_THREADS / FUSED_HPAIRS;  // row groups (4)
// row groups used for a batch of BATCH rows: a divisor of BATCH so no thread
// filters rows that do not exist (6 rows -> 3 groups x 2 rows, a quarter of the
// threads sit the H pass out; 8 rows -> 4 groups x 2 rows)
__host__ __device__ constexpr int
fused_hrowg(int batch)
{
    return batch % FUSED_HROWG_MAX == 0 ? FUSED_HROWG_MAX : 3;
}

__device__ __forceinline__ bool
any_nonfinite(float a, float b, float c, float d)
{
    float t = (a + b) + (c + d);  // Inf/NaN in any term survives the sum
    return !(fabsf(t) <= 3.402823466e+38f);
}

template<int NCH>
__device__ __forceinline__ void
store_pixel(const FusedParams& p, unsigned char* dp, int des, const float (&a)[NCH])
{
    if (NCH == 4 && p.dst_vec_ok) {
        if (p.dsttype == BT_FLOAT) {
            *reinterpret_cast<float4*>(dp) = make_float4(a[0], a[1], a[2], a[NCH - 1]);
        } else if (p.dsttype == BT_HALF) {
            __half2 h0 = __floats2half2_rn(a[0], a[1]);
            __half2 h1 = __floats2half2_rn(a[2], a[NCH - 1]);
            uint2 u;
            u.x = *reinterpret_cast<unsigned int*>(&h0);
            u.y = *reinterpret_cast<unsigned int*>(&h1);
            *reinterpret_cast<uint2*>(dp) = u;
        } else if (p.dsttype == BT_UINT16) {
            uint2 u;
            u.x = quant_u16(a[0]) | (quant_u16(a[1]) << 16);
            u.y = quant_u16(a[2]) | (quant_u16(a[NCH - 1]) << 16);
            *reinterpret_cast<uint2*>(dp) = u;
        } else {
            unsigned int u = quant_u8(a[0]) | (quant_u8(a[1]) << 8) | (quant_u8(a[2]) << 16)
                             | (quant_u8(a[NCH - 1]) << 24);
            *reinterpret_cast<unsigned int*>(dp) = u;
        }
    } else {
#pragma unroll
        for (int c = 0; c < NCH; ++c)
            store_elem(dp + c * des, p.dsttype, a[c]);
    }
}

// Two adjacent dst columns (A, B) per thread from one shared window of HT
// source pixels; the thread's rows of the batch (hr, hr+4, ...) are filtered
// interleaved, tap by tap, so independent LDS -> FFMA2 chains hide the
// shared-memory latency.  hb[] are BYTE offsets of window taps 0..3 in this
// thread's first V row (row hr); tap j is hb[j & 3] + (j >> 2) * 16, i.e. base
// register + immediate.  hws points at this pair's (wA, wB) weights:
// hws[j * FUSED_HPAIRS] for tap j (staged once per CTA, lane == pair).
template<int NCH, int HT, int BATCH>
__device__ __forceinline__ void
fused_hpass(const FusedParams& p, const unsigned char* vrows, int nb, int dyb,
            int dxA, bool pair_active, bool has_b, const int (&hb)[4],
            const float2* hws, int hr)
{
    if (!pair_active)
        return;
    const int des = (p.dsttype == BT_UINT8) ? 1 : (p.dsttype == BT_FLOAT ? 4 : 2);
    constexpr int ROWG = fused_hrowg(BATCH);
    constexpr int NR   = (BATCH + ROWG - 1) / ROWG;  // rows per thread
    float4 aA[NR], aB[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k)
        aA[k] = aB[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto tap = [&](int j) {
        const float2 w   = hws[j * FUSED_HPAIRS];
        const float2 wwA = make_float2(w.x, w.x), wwB = make_float2(w.y, w.y);
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            // rows past nb hold stale data from an earlier batch: computed
            // and discarded
            const float4 v = *reinterpret_cast<const float4*>(
                vrows + k * ROWG * FUSED_PITCH_B + hb[j & 3] + (j >> 2) * 16);
            fma4(aA[k], wwA, v);
            fma4(aB[k], wwB, v);
        }
    };
    if constexpr (HT > 0) {
#pragma unroll
        for (int j = 0; j < HT; ++j)
            tap(j);
    } else {
        // HT == 0: wide windows (strong minification), tap count at run time
        // (a multiple of 4), four taps per trip
#pragma unroll 1
        for (int j = 0; j < p.ht; j += 4) {
            tap(j);
            tap(j + 1);
            tap(j + 2);
            tap(j + 3);
        }
    }
    unsigned char* dp = p.dst + (long long)(dyb + hr) * p.dst_ystride
                        + (long long)dxA * (NCH * des);
    bool bad = false;
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        if (hr + k * ROWG < nb) {
            const float ca[4] = { aA[k].x, aA[k].y, aA[k].z, aA[k].w };
            const float cb[4] = { aB[k].x, aB[k].y, aB[k].z, aB[k].w };
            float t = 0.0f, oa[NCH], ob[NCH];
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                oa[c] = ca[c];
                ob[c] = cb[c];
                t += ca[c] + cb[c];  // Inf/NaN in any channel survives the sum
            }
            if (p.check_nonfinite && !(fabsf(t) <= 3.402823466e+38f)) {
                bad = true;
                // the gated exact kernel re-does only the marked cells
                p.dirty[((dyb + hr + k * ROWG) >> p.dirty_yshift) * p.dirty_xcells
                        + (dxA >> DIRTY_CELL_SHIFT_X)]
                    = p.nonfinite_epoch;
            }
            store_pixel<NCH>(p, dp, des, oa);
            if (has_b)
                store_pixel<NCH>(p, dp + NCH * des, des, ob);
        }
        dp += (long long)ROWG * p.dst_ystride;
    }
    if (p.check_nonfinite && bad)
        *p.nonfinite_flag = p.nonfinite_epoch;
}

// Shared memory carve-up (bytes):
//   vrows  fused_batch(VK) x FUSED_PITCH_B       V-filtered rows for the H pass
//   hws    HT x FUSED_HPAIRS x 8                 H weights (wA,wB), [tap][pair]
//   fifo   FUSED_FIFO x FUSED_THREADS x 4*es     per-thread prefetch   (ring)
//   ringw  max_band_rows x VKP x 4               ring weights            "
//   lastr  max_band_dy x 4                       completion rows         "
template<int ST, int NCH, int VK, int HT>
__global__ void __launch_bounds__(FUSED_THREADS, fused_min_ctas(SrcTraits<ST>::es, VK))
resize_fused_kernel(const FusedParams p, int max_band_rows, int max_band_dy)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    pdl_prologue();
    unsigned char* vrows  = smem_raw;
    constexpr int BATCH   = (VK == 6) ? 12 : 8;  // == fused_batch(VK), kernels.h
    constexpr int pitch_b = FUSED_PITCH_B;
    constexpr int es      = SrcTraits<ST>::es;

    const int tid     = threadIdx.x;
    const Strip strip = p.strips[blockIdx.x];
    const Band band   = p.bands[blockIdx.y];

    // ---- H-pass per-thread constants: this thread's pair of dst columns --
    const int hp           = tid % FUSED_HPAIRS;
    const int hr           = tid / FUSED_HPAIRS;
    const bool pair_active = 2 * hp < strip.ndx && hr < fused_hrowg(BATCH);
    const bool has_b       = 2 * hp + 1 < strip.ndx;
    const int dxA          = strip.dx0 + 2 * hp;
    int hb[4] = { 0, 0, 0, 0 };  // byte offsets of window taps 0..3 (row hr)
    float2* hws_all = reinterpret_cast<float2*>(smem_raw + BATCH * FUSED_PITCH_B);
    const float2* hws = hws_all + hp;
    {
        const int gp = (strip.dx0 >> 1) + hp;  // pair index in the ROI
        // ring mode: staged with cp.async (joins the group of the ring
        // weights below, completes with the first FIFO row) so the table
        // loads overlap the first source rows' latency
        const int ht = HT > 0 ? HT : p.ht;  // HT == 0: run-time window (<= 64 taps)
        for (int j = hr; j < ht; j += FUSED_HROWG_MAX) {
            const float2* g = reinterpret_cast<const float2*>(p.hw) + (size_t)gp * ht + j;
            if (2 * hp >= strip.ndx)
                hws_all[j * FUSED_HPAIRS + hp] = make_float2(0.f, 0.f);
            else if (VK > 0)
                cp_async_copy<8>(smem_u32(hws_all + j * FUSED_HPAIRS + hp), g);
            else
                hws_all[j * FUSED_HPAIRS + hp] = *g;
        }
    }
    // the pair's window start: loaded here, used after the first source rows
    // have been requested (see hb[] below)
    int hfirst_v = 0;
    if (pair_active)
        asm volatile("ld.global.nc.s32 %0, [%1];"
                     : "=r"(hfirst_v)
                     : "l"(p.hfirst + (strip.dx0 >> 1) + hp));
    auto finish_hb = [&]() {
        if (pair_active) {
            const int q = hfirst_v - strip.e0 / NCH;  // window start in the strip
#pragma unroll
            for (int k = 0; k < 4; ++k)
                hb[k] = planar_byte(q + k) + hr * pitch_b;
        }
    };

    // ---- V-pass per-thread constants: this thread's 4 source elements ---
    const bool vactive  = tid < strip.nvec;
    const int e         = strip.e0 + 4 * tid;
    const int row_elems = p.src_w * NCH;
    const bool vec_ok   = p.src_vec_ok != 0;
    // where this thread parks its 4 V-filtered elements in a V row
    VPark<NCH> park;
    park.init(strip.e0, tid);
    if (NCH < 4) {
        // padded rows: the unused lanes are never written, they must read as zero
        for (int i = tid; i < BATCH * (FUSED_PITCH_B / 16); i += FUSED_THREADS)
            reinterpret_cast<uint4*>(vrows)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
    }
    const int dy_end    = band.dy0 + band.ndy;
    int nb              = 0;

    if constexpr (VK > 0) {
        // ring row: VK weights padded to float4 loads (duplicated into FFMA2
        // operand pairs in registers: a MOV is cheaper than a shared-memory
        // wavefront here)
        constexpr int VKP  = (VK + 3) & ~3;
        // FIFO geometry: thread t owns bytes [t*4*es, (t+1)*4*es) of every row
        // slot (dense, so 8- and 4-byte accesses stay bank-conflict free)
        constexpr int NFIFO = fused_fifo_rows(es, VK);  // rows in flight per thread
        constexpr int TB   = 4 * es;             // bytes per thread per row
        constexpr int SLOT = FUSED_THREADS * TB;  // bytes between FIFO rows
        // fixed-offset regions first, so their addresses are compile-time
        // offsets from the dynamic shared memory base
        unsigned char* fifo = smem_raw + BATCH * FUSED_PITCH_B
                              + (HT > 0 ? HT : p.ht) * FUSED_HPAIRS * 8;
        float* ringw = reinterpret_cast<float*>(fifo + NFIFO * SLOT);
        int* lastr   = reinterpret_cast<int*>(ringw + (size_t)max_band_rows * VKP);

        // stage this band's ring weights and completion rows
        const float* gw = p.vring + (size_t)band.woff * VKP;
        for (int i = tid; i < band.nrows * (VKP / 4); i += FUSED_THREADS)
            cp_async_copy<16>(smem_u32(ringw + 4 * i), gw + 4 * i);
        for (int i = tid; i < band.ndy; i += FUSED_THREADS)
            cp_async_copy<4>(smem_u32(lastr + i), p.vlast + band.dy0 + i);
        cp_async_commit();  // hws + ringw + lastr: one group, older than every FIFO row

        float4 acc[VK];
#pragma unroll
        for (int k = 0; k < VK; ++k)
            acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);

        const int rend = band.r0 + band.nrows;
        // cp.async needs the global address aligned to the copy size (the host
        // sets src_vec_ok only when rows and strips are 4-element aligned) and
        // every staged group inside the row; a ragged last strip takes the
        // synchronous guarded loads instead (uniform per CTA)
        const bool async_ok
            = vec_ok
              && ((row_elems & 3) == 0 || strip.e0 + 4 * strip.nvec <= row_elems);
        // groups that lie entirely past the end of the row (tap padding of the
        // last strip) are never fetched: their FIFO slots stay zero
        const bool vfetch = vactive && e + 4 <= row_elems;
        // this thread's FIFO column, as a 32-bit shared-window address
        unsigned fifo_s;
        {
            // opaque move: keeps the address in a register (under the register
            // cap the compiler otherwise recomputes it every row)
            unsigned t = smem_u32(fifo) + tid * TB;
            asm volatile("mov.u32 %0, %1;" : "=r"(fifo_s) : "r"(t));
        }

        // running state of the prefetcher: next row to fetch, its address,
        // its FIFO slot
        const unsigned char* gnext = p.src + (size_t)e * es
                                     + (long long)band.r0 * p.src_ystride;
        int rnext     = band.r0;
        // In steady state the FIFO is full: a fetch goes into the slot the
        // previous pop freed, so one walking index serves both ends.
        int slot_free = 0;
        auto issue_next = [&]() {
            if (vfetch && rnext < rend)
                cp_async_copy<4 * es>(fifo_s + slot_free, gnext);
            cp_async_commit();
            gnext += p.src_ystride;
            ++rnext;
        };
        int slot_cur = 0;
        auto fifo_pop = [&]() -> uint4 {
            uint4 raw;
            if (es == 4)
                raw = lds_v4(fifo_s + slot_cur);
            else if (es == 2) {
                uint2 t = lds_v2(fifo_s + slot_cur);
                raw     = make_uint4(t.x, t.y, 0u, 0u);
            } else {
                raw = make_uint4(lds_u32(fifo_s + slot_cur), 0u, 0u, 0u);
            }
            slot_free = slot_cur;
            slot_cur  = (slot_cur + SLOT) & (NFIFO * SLOT - 1);
            return raw;
        };
        // one source row: scatter it into the VK open destination rows
        auto accumulate = [&](const uint4& raw, const float4* wr) {
            const float4 v = convert_raw4<ST>(raw);
            float w[VKP];
#pragma unroll
            for (int k = 0; k < VKP / 4; ++k) {
                float4 t     = wr[k];
                w[4 * k]     = t.x;
                w[4 * k + 1] = t.y;
                w[4 * k + 2] = t.z;
                w[4 * k + 3] = t.w;
            }
#pragma unroll
            for (int k = 0; k < VK; ++k)
                fma4(acc[k], w[k], v);
        };
        // a destination row is complete: park it for the H pass
        auto emit = [&](int s) {
            if (vactive)
                park.store(vrows + (size_t)nb * pitch_b, acc[s]);
            acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
            ++nb;
        };
        // BATCH is a multiple of VK, so the buffer can only fill up at the end
        // of a slot-unrolled group: the H pass is called from ONE place per
        // loop instead of VK (its code is large; keep the kernel in I-cache)
        auto flush_if_full = [&](int dy_done) {
            if (nb + VK > BATCH || dy_done >= dy_end) {
                __syncthreads();
                fused_hpass<NCH, HT, BATCH>(p, vrows, nb, dy_done - nb, dxA, pair_active,
                                            has_b, hb, hws, hr);
                __syncthreads();
                nb = 0;
            }
        };

        int r              = band.r0;
        const float4* wrow = reinterpret_cast<const float4*>(ringw);
        const int* lastp   = lastr;
        if (async_ok) {
            if (!vfetch) {
#pragma unroll
                for (int q = 0; q < NFIFO; ++q)
                    if (es == 4)
                        *reinterpret_cast<uint4*>(fifo + tid * TB + q * SLOT)
                            = make_uint4(0u, 0u, 0u, 0u);
                    else if (es == 2)
                        *reinterpret_cast<uint2*>(fifo + tid * TB + q * SLOT)
                            = make_uint2(0u, 0u);
                    else
                        *reinterpret_cast<unsigned*>(fifo + tid * TB + q * SLOT) = 0u;
            }
#pragma unroll
            for (int q = 0; q < NFIFO; ++q) {
                slot_free = q * SLOT;
                issue_next();
            }
            finish_hb();
            cp_async_wait<NFIFO - 1>();  // tables and first row landed
            uint4 rawA = fifo_pop(), rawB = make_uint4(0u, 0u, 0u, 0u);
            __syncthreads();  // ringw / lastr visible
            // Software pipeline: the row being accumulated was popped one
            // step earlier; its FIFO slot is refilled and the NEXT row popped
            // before the FMAs so the LDS latency hides behind them.  Two
            // steps per trip so the two register sets swap roles for free.
            auto step = [&](const uint4& cur, uint4& nxt) {
                issue_next();
                cp_async_wait<NFIFO - 1>();
                nxt = fifo_pop();
                accumulate(cur, wrow);
                wrow += VKP / 4;
            };
            for (int dy = band.dy0; dy < dy_end; dy += VK) {
#pragma unroll
                for (int s = 0; s < VK; ++s) {
                    if (dy + s < dy_end) {
                        const int last = *lastp++;
                        for (; r < last; r += 2) {
                            step(rawA, rawB);
                            step(rawB, rawA);
                        }
                        if (r == last) {
                            step(rawA, rawB);
                            rawA = rawB;
                            ++r;
                        }
                        emit(s);
                    }
                }
                flush_if_full(min(dy + VK, dy_end));
            }
        } else {
            finish_hb();
            cp_async_wait<0>();
            __syncthreads();  // ringw / lastr visible
            for (int dy = band.dy0; dy < dy_end; dy += VK) {
#pragma unroll
                for (int s = 0; s < VK; ++s) {
                    if (dy + s < dy_end) {
                        const int last = *lastp++;
                        for (; r <= last; ++r) {
                            uint4 raw = make_uint4(0u, 0u, 0u, 0u);
                            if (vactive)
                                raw = load_raw4<ST>(p.src + (long long)r * p.src_ystride,
                                                    e, row_elems, false);
                            accumulate(raw, wrow);
                            wrow += VKP / 4;
                        }
                        emit(s);
                    }
                }
                flush_if_full(min(dy + VK, dy_end));
            }
        }
        cp_async_wait<0>();
    } else {
        finish_hb();
        // gather mode (upsizing): each dst row gathers its own vt source rows
        // through L1 (consecutive dst rows share rows).  A strip holds few
        // source pixels here, so the (row, group) items of a whole batch are
        // spread over all threads instead of leaving most of the CTA idle.
        const int nvec = strip.nvec;
        // the thread's first item of every batch is the same (row, group):
        // its parking offsets are computed once
        const int row0 = tid / nvec, v0 = tid - row0 * nvec;
        VPark<NCH> park0;
        park0.init(strip.e0, v0);
        for (int dyb = band.dy0; dyb < dy_end; dyb += BATCH) {
            const int nrows = min(BATCH, dy_end - dyb);
            for (int it = tid; it < nrows * nvec; it += FUSED_THREADS) {
                int row = row0, v = v0;
                if (it != tid) {
                    row = it / nvec;
                    v   = it - row * nvec;
                }
                const int dy    = dyb + row;
                const int ev    = strip.e0 + 4 * v;
                const int first = p.vfirst[dy];
                const float* w  = p.vw + (size_t)dy * p.vt;
                const unsigned char* rowp = p.src + (long long)first * p.src_ystride;
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int j = 0; j < p.vt; ++j) {
                    uint4 raw = load_raw4<ST>(rowp, ev, row_elems, vec_ok);
                    fma4(a, __ldg(w + j), convert_raw4<ST>(raw));
                    rowp += p.src_ystride;
                }
                if (it == tid) {
                    park0.store(vrows + (size_t)row * pitch_b, a);
                } else {
                    VPark<NCH> pk;
                    pk.init(strip.e0, v);
                    pk.store(vrows + (size_t)row * pitch_b, a);
                }
            }
            __syncthreads();
            fused_hpass<NCH, HT, BATCH>(p, vrows, nrows, dyb, dxA, pair_active, has_b,
                                        hb, hws, hr);
            __syncthreads();
        }
    }
}

}  // namespace b2r
