// stream_kernel.cuh -- the hot path for downsizing (ring-mode) shapes:
// a persistent, warp-specialised version of the fused two-pass resize.
//
// One CTA per SM walks a static list of work items (column strip x row band
// of the dst ROI, the same decomposition as resize_fused_kernel).  Three
// roles run concurrently inside the CTA:
//
//   producer  warp 0, one elected lane (its warpgroup gives its registers
//             to the H warps with setmaxnreg).  Streams the band's source rows
//             HBM -> shared memory with TMA: per stage STREAM_ROWS rows of the
//             strip as 2-D tensor-map boxes (cp.async.bulk.tensor.2d, 256
//             elements x STREAM_ROWS rows each) plus the ring weights of those
//             rows (cp.async.bulk), all completing on the stage's `full`
//             mbarrier.  A stage is re-armed when the V warps have arrived on
//             its `empty` mbarrier.  The producer runs ahead across item
//             boundaries, so the pipe never drains between items.
//   V warps   8 (W2 = 1) or 16 (W2 = 2) warps; thread t owns 4 consecutive
//             source elements of the strip.  Every source row is read from shared memory ONCE and
//             scattered into the VK destination rows open at that moment
//             (register accumulators, ring weights from the stage).  A
//             completed destination row is parked in one of two V-row
//             buffers for the H warps.
//   H warps   8 warps; thread = two adjacent dst columns x NR rows of the
//             batch, H weights held in registers for the whole item; the
//             quantise + store epilogue of resize_fused_kernel.
//
// V and H overlap: while the H warps filter batch n the V warps fill batch
// n+1 (named barriers 1..4 hand the two buffers back and forth).
//
// Zero weights / non-finite pixels.  The reference never multiplies a zero
// weight (imagebufalgo_xform.cpp:749-759).  Here a V thread tests every
// float / half source group it loads; a group with an Inf/NaN takes a
// predicated accumulate that skips zero ring weights (warp-divergent, rare).
// The H pass screens its outputs; a thread with a non-finite one re-filters
// its rows from the parked V rows with the zero taps skipped.  So the result
// has the reference's finite / Inf / NaN pattern without a second kernel.
#pragma once

#include "fused_kernel.cuh"

#include <cuda.h>  // CUtensorMap
#include <cuda_fp16.h>

#include <type_traits>

namespace b2r {

// ---- mbarrier / TMA wrappers (PTX) ---------------------------------------
__device__ __forceinline__ void
mbar_init(unsigned bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void
mbar_expect_tx(unsigned bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void
mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void
mbar_wait(unsigned bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred p;\n"
                 "B2R_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@p bra B2R_DONE;\n"
                 "bra B2R_WAIT;\n"
                 "B2R_DONE:\n"
                 "}" ::"r"(bar),
                 "r"(parity)
                 : "memory");
}
__device__ __forceinline__ void
tma_load_2d(unsigned smem_dst, const CUtensorMap* map, int c0, int c1, unsigned bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
                 "l"(map), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void
bulk_load_1d(unsigned smem_dst, const void* gsrc, unsigned bytes, unsigned bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes"
                 " [%0], [%1], %2, [%3];" ::"r"(smem_dst),
                 "l"(gsrc), "r"(bytes), "r"(bar)
                 : "memory");
}
// one lane of a converged warp (the TMA issue slot)
__device__ __forceinline__ bool
elect_one()
{
    unsigned pred;
    asm volatile("{\n"
                 ".reg .pred P;\n"
                 "elect.sync _|P, 0xffffffff;\n"
                 "selp.u32 %0, 1, 0, P;\n"
                 "}"
                 : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void
named_bar_sync(int id, int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void
named_bar_arrive(int id, int count)
{
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- geometry ---------------------------------------------------------------
// W2 = 1: 256 V threads, the strip geometry of the fused kernel (<= 128 dst
// columns, <= 256 source pixels); W2 = 2: 512 V threads, strips twice as wide,
// half as many rows per hand-over batch.  A V thread always owns 4 elements.
// E2: four-element groups per V thread -- 2 for the 1- and 2-byte source types
// of the wide geometry (16 / 8 bytes per thread per row: half the per-element
// control and load overhead), 1 otherwise.
// NCH < 4: a V row holds 256 * W2 pixels, i.e. fewer elements than 256 * W2 V
// threads would cover (RGB, wide: 1536 of 2048) -- the CTA runs only the V
// warpgroups that have elements, so no warp spends its issue slots on padding
// (stream_vthreads() in kernels.h, shared with the planner and the launcher).
template<int W2, int E2 = 1, int NCH = 4> struct StreamGeom {
    static constexpr int NVT     = stream_vthreads(NCH, W2) / E2;  // V threads
    static constexpr int VPIX    = STREAM_VTHREADS * W2;      // pixels a V row holds
    static constexpr int HCOLS   = FUSED_HCOLS * W2;          // dst columns per strip
    static constexpr int HPAIRS  = HCOLS / 2;
    static constexpr int ROWG    = STREAM_HTHREADS / HPAIRS;  // H row groups (4 / 2)
    static constexpr int PLANE   = VPIX / 4 + 2;              // slots per plane
    static constexpr int PITCH_B = 4 * PLANE * 16;            // bytes per V row
    static constexpr int THREADS = STREAM_PTHREADS + NVT + STREAM_HTHREADS;
    static constexpr int ROWS    = stream_rows(W2);           // source rows per stage
    // registers after setmaxnreg (the kernel is compiled for 65536 / THREADS)
    // LREGS: what a thread holds at launch (__launch_bounds__(THREADS, 1)).
    //   NVT 512: 72 at launch -> P 24, V 64, H 112   (128 x 24 + 512 x 64 + 256 x 112 = 64512)
    //   NVT 384: 80           -> P 24, V 72, H 120   (61440 of 61440; at 112 the
    //                                                 16-tap H windows spill)
    //   NVT 256: 96           -> P 24, V 96, H 120   (58368 of 61440)
    //   NVT 128: 128          -> P 24, V 96, H stays at 128
    static constexpr int LREGS   = (65536 / THREADS) / 8 * 8;
    static constexpr int VREGS   = NVT >= 512 ? 64 : (NVT == 384 ? 72 : 96);
    static constexpr int HREGS   = NVT >= 512 ? 112 : 120;
};
template<int W2>
__device__ __forceinline__ int
stream_planar_byte(int q)
{
    return ((q >> 2) + (q & 3) * StreamGeom<W2>::PLANE) * 16;
}

// Where a V thread parks its 4 * E2 filtered elements.  NCH == 4: E2 planar
// float4 pixels; otherwise one scalar offset per element (padded pixels).
template<int NCH, int W2, int E2> struct StreamPark {
    static constexpr int NOFF = NCH == 4 ? E2 : 4 * E2;
    int off[NOFF];
    __device__ __forceinline__ void init(int e0, int v)
    {
        if (NCH == 4) {
#pragma unroll
            for (int k = 0; k < E2; ++k)
                off[k] = stream_planar_byte<W2>(E2 * v + k);
        } else {
            const int px0 = e0 / NCH;
#pragma unroll
            for (int k = 0; k < NOFF; ++k) {
                int E  = e0 + 4 * E2 * v + k;
                int px = E / NCH;
                off[k] = stream_planar_byte<W2>(px - px0) + (E - NCH * px) * 4;
            }
        }
    }
    __device__ __forceinline__ void store(unsigned row, const float4 (&a)[E2]) const
    {
#pragma unroll
        for (int h = 0; h < E2; ++h) {
            if (NCH == 4) {
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(row + off[h]),
                             "f"(a[h].x), "f"(a[h].y), "f"(a[h].z), "f"(a[h].w)
                             : "memory");
            } else {
                constexpr int S = NCH == 4 ? 0 : 1;  // keeps the indices in range for NCH == 4
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[S * (4 * h)]), "f"(a[h].x)
                             : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[S * (4 * h + 1)]),
                             "f"(a[h].y)
                             : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[S * (4 * h + 2)]),
                             "f"(a[h].z)
                             : "memory");
                asm volatile("st.shared.f32 [%0], %1;" ::"r"(row + off[S * (4 * h + 3)]),
                             "f"(a[h].w)
                             : "memory");
            }
        }
    }
};

// Raw group of 4 * E2 elements -> E2 float4 values.
template<int ST, int E2>
__device__ __forceinline__ void
stream_convert(const uint4& r, float4 (&v)[E2])
{
    if (E2 == 1) {
        v[0] = convert_raw4<ST>(r);
    } else if (SrcTraits<ST>::es == 2) {
        v[0]      = convert_raw4<ST>(make_uint4(r.x, r.y, 0u, 0u));
        v[E2 - 1] = convert_raw4<ST>(make_uint4(r.z, r.w, 0u, 0u));
    } else {
        v[0]      = convert_raw4<ST>(make_uint4(r.x, 0u, 0u, 0u));
        v[E2 - 1] = convert_raw4<ST>(make_uint4(r.y, 0u, 0u, 0u));
    }
}

// ---- H-pass epilogue: quantise and store two adjacent pixels ---------------
// convert_type<float,D> (fmath.h:753-764, 1009-1031) per channel, then the
// pair's 2 * NCH elements leave as the widest words their alignment allows
// (RGB: three words of two elements; `palign` = power of two dividing the ROI
// origin and the row pitch).
__device__ __forceinline__ unsigned
pack_half2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<unsigned*>(&h);
}
template<int NCH>
__device__ __forceinline__ void
stream_store_pair(const FusedParams& p, unsigned char* dp, int des, const float4& A,
                  const float4& B, bool has_b)
{
    const float a[4] = { A.x, A.y, A.z, A.w };
    const float b[4] = { B.x, B.y, B.z, B.w };
    if (NCH == 4 || !has_b || p.dst_pow2 < 2 * des * (NCH == 3 ? 1 : NCH)) {
        float oa[NCH], ob[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            oa[c] = a[c];
            ob[c] = b[c];
        }
        store_pixel<NCH>(p, dp, des, oa);
        if (has_b)
            store_pixel<NCH>(p, dp + NCH * des, des, ob);
        return;
    }
    // the pair's elements in memory order
    float e[2 * NCH];
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        e[c]       = a[c];
        e[NCH + c] = b[c];
    }
    if (p.dsttype == BT_FLOAT) {
        if (NCH == 2)
            *reinterpret_cast<float4*>(dp) = make_float4(e[0], e[1], e[2 % (2 * NCH)],
                                                         e[3 % (2 * NCH)]);
        else
#pragma unroll
            for (int k = 0; k < NCH; ++k)
                reinterpret_cast<float2*>(dp)[k] = make_float2(e[2 * k], e[2 * k + 1]);
    } else if (p.dsttype == BT_HALF) {
        unsigned wds[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
            wds[k] = pack_half2(e[2 * k], e[2 * k + 1]);
        if (NCH == 2)
            *reinterpret_cast<uint2*>(dp) = make_uint2(wds[0], wds[1 % NCH]);
        else
#pragma unroll
            for (int k = 0; k < NCH; ++k)
                reinterpret_cast<unsigned*>(dp)[k] = wds[k];
    } else if (p.dsttype == BT_UINT16) {
        unsigned wds[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
            wds[k] = quant_u16(e[2 * k]) | (quant_u16(e[2 * k + 1]) << 16);
        if (NCH == 2)
            *reinterpret_cast<uint2*>(dp) = make_uint2(wds[0], wds[1 % NCH]);
        else
#pragma unroll
            for (int k = 0; k < NCH; ++k)
                reinterpret_cast<unsigned*>(dp)[k] = wds[k];
    } else {
        unsigned short wds[NCH];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
            wds[k] = (unsigned short)(quant_u8(e[2 * k]) | (quant_u8(e[2 * k + 1]) << 8));
        if (NCH == 2)
            *reinterpret_cast<unsigned*>(dp) = unsigned(wds[0]) | (unsigned(wds[1 % NCH]) << 16);
        else
#pragma unroll
            for (int k = 0; k < NCH; ++k)
                reinterpret_cast<unsigned short*>(dp)[k] = wds[k];
    }
}

// ---- H pass over one batch ---------------------------------------------------
// Same arithmetic as fused_hpass (two adjacent dst columns per thread from one
// window of HT parked pixels); the weights stay in registers for the whole
// item and the thread's NR rows are filtered one after the other (HT
// independent loads and four FFMA2 chains per row are enough to keep the
// pipes fed; interleaving rows would not fit the register budget).
// FP: screen the outputs and re-filter with the zero taps skipped.
template<int NCH, int HT, int W2, int NR, bool FP, bool HC>
__device__ __forceinline__ void
stream_hpass(const FusedParams& p, unsigned char* dbase, unsigned vrows, int nb, int dyb,
             int dxA, bool has_b, const int (&hb)[4], const float2 (&hw)[HT], int hmix, int hr)
{
    using G = StreamGeom<W2>;
    const int des = (p.dsttype == BT_UINT8) ? 1 : (p.dsttype == BT_FLOAT ? 4 : 2);
    unsigned char* dp = dbase + (long long)(dyb + hr) * p.dst_ystride
                        + (long long)dxA * (NCH * des);
#pragma unroll 1
    for (int k = 0; k < NR; ++k) {
        if (hr + k * G::ROWG >= nb)
            break;
        auto ldv = [&](int j) -> float4 {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                         : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                         : "r"(vrows + hb[j & 3] + (j >> 2) * 16));
            return v;
        };
        float4 aA = make_float4(0.f, 0.f, 0.f, 0.f), aB = aA;
#pragma unroll
        for (int j = 0; j < HT; ++j) {
            const float4 v = ldv(j);
            fma4(aA, hw[j].x, v);
            fma4(aB, hw[j].y, v);
        }
        if (FP) {
            const float t = ((aA.x + aA.y) + (aA.z + aA.w)) + ((aB.x + aB.y) + (aB.z + aB.w));
            if (!(fabsf(t) <= 3.402823466e+38f)) {
                // an Inf/NaN reached this window: filter again the way the
                // reference does, never touching a zero weight
                aA = aB = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int j = 0; j < HT; ++j) {
                    const float4 v = ldv(j);
                    if (hw[j].x != 0.0f)
                        fma4(aA, hw[j].x, v);
                    if (hw[j].y != 0.0f)
                        fma4(aB, hw[j].y, v);
                    // a weight that stands for clamped taps of both signs:
                    // w1*Inf + w2*Inf is NaN in the reference
                    if ((hmix >> j) & 1)
                        fma4(aA, 0.0f, v);
                    if ((hmix >> (16 + j)) & 1)
                        fma4(aB, 0.0f, v);
                }
            }
        }
        if (HC) {
            // rangeexpand the results (not alpha / z)
            float* fa[4] = { &aA.x, &aA.y, &aA.z, &aA.w };
            float* fb[4] = { &aB.x, &aB.y, &aB.z, &aB.w };
            // (a half destination holds the resize's result as a half before
            // rangeexpand reads it back in place, oiiotool.cpp:4661-4662)
            const bool hdst = p.dsttype == BT_HALF;
            auto exp1 = [hdst](float y) -> float {
                if (hdst)
                    y = __half2float(__float2half_rn(y));
                return dev_rangeexpand_fast(y);
            };
#pragma unroll
            for (int c = 0; c < NCH; ++c)
                if (!((p.hc_skip >> c) & 1)) {
                    *fa[c] = exp1(*fa[c]);
                    *fb[c] = exp1(*fb[c]);
                }
            if (p.hc_clamp) {
                // ImageBufAlgo::clamp(small, small, 0, FLT_MAX, clampalpha01 = true)
                // as clamp_kernel does it (NaN -> 0)
                auto cl = [](float r, bool alpha) -> float {
                    if (!(0.0f <= r))
                        r = 0.0f;
                    if (r > 3.402823466e+38f)
                        r = 3.402823466e+38f;
                    if (alpha && r > 1.0f)
                        r = 1.0f;
                    return r;
                };
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    *fa[c] = cl(*fa[c], c == p.hc_alpha);
                    *fb[c] = cl(*fb[c], c == p.hc_alpha);
                }
            }
        }
        stream_store_pair<NCH>(p, dp, des, aA, aB, has_b);
        dp += (long long)G::ROWG * p.dst_ystride;
        vrows += G::ROWG * G::PITCH_B;
    }
}

// Shared memory carve-up (bytes), all regions 128-byte aligned:
//   stages  NST x STAGE_B   STAGE_B = NBOX boxes of (ROWS x 256 elements) + ROWS ring-table rows
//   vbuf    2 x BATCH x PITCH_B
//   bars    2 x NST mbarriers
template<int ST> struct StreamE2 {
    // Measured (C2 half RGBA, C5 uint16 RGB): two groups per thread with half
    // as many V warps is 5-6 % SLOWER than one group per thread -- the V pass is
    // bound by latency per row, not by its instruction count -- so every type
    // runs one group per thread; the E2 = 2 code stays for the record.
    template<int W2> static constexpr int of() { return 1; }
};
template<int ST, int VK, int W2, int NCH = 4> struct StreamSmem {
    static constexpr int E2       = StreamE2<ST>::template of<W2>();
    using G = StreamGeom<W2, E2, NCH>;
    static constexpr int es       = SrcTraits<ST>::es;
    static constexpr int ROWS     = G::ROWS;
    static constexpr int BOXW     = 256;  // elements per box row (the TMA limit)
    static constexpr int NBOX     = G::NVT * 4 * E2 / BOXW;
    static constexpr int BOX_ROWB = BOXW * es;
    static constexpr int BOX_B    = ROWS * BOX_ROWB;
    static constexpr int WROW     = stream_wrow(VK);  // floats per ring-table row
    static constexpr int WROWB    = WROW * 4;
    static constexpr int DATA_B   = NBOX * BOX_B;
    static constexpr int STAGE_B  = stream_stage_bytes(es, VK, W2);
    static constexpr int BATCH    = stream_batch(VK, W2);
    static constexpr int NR       = BATCH / G::ROWG;
    static constexpr int VBUF_B   = BATCH * G::PITCH_B;
    static constexpr int NST      = stream_stages(es, VK, W2);
    static constexpr int BARS_OFF = NST * STAGE_B + 2 * VBUF_B;
    static constexpr int TOTAL    = BARS_OFF + 2 * NST * 8 + 64;
    static_assert(DATA_B + ROWS * WROWB <= STAGE_B, "stage layout");
    static_assert(NST >= 2, "at least two stages");
};

// HC: highlight compensation fused in (float / half sources) -- a template
// parameter because the V loop has no register to spare for a run-time flag
// (the run-time version cost the plain C0 frame 115 -> 136 us).
template<int ST, int NCH, int VK, int HT, int W2, bool HC = false>
__global__ void __launch_bounds__(StreamSmem<ST, VK, W2, NCH>::G::THREADS, 1)
resize_stream_kernel(const __grid_constant__ CUtensorMap tmap, const FusedParams p)
{
    using S = StreamSmem<ST, VK, W2, NCH>;
    using G = typename S::G;
    constexpr int E2    = S::E2;
    constexpr int es    = S::es;
    constexpr int ROWS  = S::ROWS;
    constexpr int NST   = S::NST;
    constexpr int WROW  = S::WROW;
    constexpr int BATCH = S::BATCH;
    constexpr int NR    = S::NR;
    constexpr int NSYNC = G::NVT + STREAM_HTHREADS;  // threads on the V <-> H barriers
    constexpr bool FP   = (ST == ST_FLOAT || ST == ST_HALF);
    // stage-specialised V pass (see the V warps below)
    constexpr bool JIT  = W2 == 2 && FP && VK == 6;
    constexpr bool SPEC = W2 == 2 && (!FP || VK <= 6);
    static_assert(!HC || FP, "highlight compensation is fused for float / half sources");
    static_assert(BATCH % G::ROWG == 0, "batch rows must split evenly over the H row groups");

    extern __shared__ __align__(1024) unsigned char stream_smem[];
    pdl_prologue();
    const unsigned smem0  = smem_u32(stream_smem);
    const unsigned vbuf0  = smem0 + NST * S::STAGE_B;
    const unsigned bars   = smem0 + S::BARS_OFF;  // full[NST], empty[NST]
    const int warp        = threadIdx.x >> 5;
    const int lane        = threadIdx.x & 31;
    // work items: (frame, strip, band), frame-major -- a batch of frames of one
    // shape runs as ONE persistent launch (p.nframes > 1)
    const int per_frame   = p.nstrips * p.nbands;
    const int nitems      = per_frame * (p.nframes > 1 ? p.nframes : 1);

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) {
            mbar_init(bars + 8 * s, 1);
            mbar_init(bars + 8 * (NST + s), G::NVT / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (NCH < 4) {
        // padded V rows: the unused lanes are never written, they must read as zero
        for (int i = threadIdx.x; i < 2 * S::VBUF_B / 16; i += G::THREADS)
            reinterpret_cast<uint4*>(stream_smem + NST * S::STAGE_B)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    __syncthreads();

    if (warp >= (G::NVT + STREAM_HTHREADS) / 32) {
        // ================= producer ========================================
        // The LAST warpgroup: the issue arbiter favours high warp ids, and the
        // ring must never wait for its refill requests.  One warp runs the
        // loop (converged; an elected lane issues the copies); the warpgroup
        // hands its registers to the H warps.
        asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
        if (warp == (G::NVT + STREAM_HTHREADS) / 32) {
            int stage = 0;
            unsigned phase = 0;
            for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
                const int frame   = item / per_frame;
                const int fitem   = item - frame * per_frame;
                const Strip strip = p.strips[fitem % p.nstrips];
                const Band band   = p.bands[fitem / p.nstrips];
                const CUtensorMap* map
                    = p.batch_maps ? reinterpret_cast<const CUtensorMap*>(p.batch_maps) + frame
                                   : &tmap;
                const int nst     = (band.nrows + ROWS - 1) / ROWS;
                const int nbox    = min(S::NBOX, (strip.nvec * 4 + S::BOXW - 1) / S::BOXW);
                const unsigned tx = nbox * S::BOX_B + ROWS * S::WROWB;
                const float* gw   = p.vstream + (size_t)band.soff * WROW;
                int row           = band.r0 - p.tma_row0;
                for (int i = 0; i < nst; ++i) {
                    const unsigned full  = bars + 8 * stage;
                    const unsigned empty = bars + 8 * (NST + stage);
                    mbar_wait(empty, phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(full, tx);
                        const unsigned sb = smem0 + stage * S::STAGE_B;
#pragma unroll
                        for (int b = 0; b < S::NBOX; ++b)
                            if (b < nbox)
                                tma_load_2d(sb + b * S::BOX_B, map, strip.e0 + b * S::BOXW, row,
                                            full);
                        bulk_load_1d(sb + S::DATA_B, gw, ROWS * S::WROWB, full);
                    }
                    __syncwarp();
                    gw += ROWS * WROW;
                    row += ROWS;
                    if (++stage == NST) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp < G::NVT / 32) {
        // ================= V warps =========================================
        if (G::VREGS < G::LREGS)
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(G::VREGS));
        const int tid = threadIdx.x;
        int stage = 0;
        unsigned phase = 0;
        int nbatch = 0;  // batches handed over so far (whole kernel)
        // this thread's 4*E2*es bytes inside a stage
        constexpr int EPT = 4 * E2;
        const int toff = ((tid * EPT) / S::BOXW) * S::BOX_B + ((tid * EPT) % S::BOXW) * es;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int fitem   = item % per_frame;
            const Strip strip = p.strips[fitem % p.nstrips];
            const Band band   = p.bands[fitem / p.nstrips];
            const bool vactive = tid * E2 < strip.nvec;
            // (Measured: letting V warps that have no elements in the strip --
            // RGB strips fill 3/4 of the V threads -- skip the arithmetic and only
            // follow the barrier protocol is SLOWER, in line or out of line: the
            // second path costs the hot loop registers (16 bytes of spills per
            // stage); C0 111 -> 145 us, C5 26.9 -> 31.0 us.)
            StreamPark<NCH, W2, E2> park;
            park.init(strip.e0, tid);
            // fused highlight compensation (float sources): which of this
            // thread's elements are range-compressed on load
            unsigned hcmask = 0;
            if (HC) {
#pragma unroll
                for (int k = 0; k < 4 * E2; ++k) {
                    const int ch = (strip.e0 + 4 * E2 * tid + k) % NCH;
                    if (!((p.hc_skip >> ch) & 1))
                        hcmask |= 1u << k;
                }
            }
            // raw group -> float values (+ rangecompress)
            auto load_values = [&](const uint4& r, float4 (&v)[E2]) {
                stream_convert<ST, E2>(r, v);
                if (HC && hcmask) {
                    // a half source is compressed into a half temporary by the
                    // three passes (tmpimg has the source's type,
                    // oiiotool.cpp:4646-4649): round the way that store does
                    auto cmp = [](float x) -> float {
                        const float y = dev_rangecompress_fast(x);
                        return ST == ST_HALF ? __half2float(__float2half_rn(y)) : y;
                    };
#pragma unroll
                    for (int h = 0; h < E2; ++h) {
                        if (hcmask & (1u << (4 * h)))
                            v[h].x = cmp(v[h].x);
                        if (hcmask & (2u << (4 * h)))
                            v[h].y = cmp(v[h].y);
                        if (hcmask & (4u << (4 * h)))
                            v[h].z = cmp(v[h].z);
                        if (hcmask & (8u << (4 * h)))
                            v[h].w = cmp(v[h].w);
                    }
                }
            };
            // acc[0] is the oldest open destination row (the ring table is
            // laid out the same way).  When a row completes it is parked and
            // the others move up one place -- not with register moves: the
            // NEXT source row's FFMA2s read acc[k+1] and write acc[k]
            // (`shift`), so the rotation costs nothing.
            float4 acc[VK][E2];
#pragma unroll
            for (int k = 0; k < VK; ++k)
#pragma unroll
                for (int h = 0; h < E2; ++h)
                    acc[k][h] = make_float4(0.f, 0.f, 0.f, 0.f);
            bool shift = false;  // acc[0] was parked after the previous source row
            int ndone  = 0;      // destination rows of the band completed so far
            unsigned prow = vbuf0 + (nbatch & 1) * S::VBUF_B;  // where the next row is parked
            // park acc[0] as the next destination row; `first` / `last`: the
            // row opens / closes a hand-over batch
            auto park_oldest = [&](bool first, bool last) {
                if (first && nbatch >= 2)  // the H warps are done with this buffer
                    named_bar_sync(3 + (nbatch & 1), NSYNC);
                if (vactive)
                    park.store(prow, acc[0]);
                prow += G::PITCH_B;
                ++ndone;
                if (last) {
                    named_bar_arrive(1 + (nbatch & 1), NSYNC);
                    ++nbatch;
                    prow = vbuf0 + (nbatch & 1) * S::VBUF_B;
                }
            };
            // explicit rotation (only when a second row completes with the
            // same source row: edges, ratios near 1)
            auto rotate = [&]() {
#pragma unroll
                for (int k = 0; k + 1 < VK; ++k)
#pragma unroll
                    for (int h = 0; h < E2; ++h)
                        acc[k][h] = acc[k + 1][h];
#pragma unroll
                for (int h = 0; h < E2; ++h)
                    acc[VK - 1][h] = make_float4(0.f, 0.f, 0.f, 0.f);
            };
            // The band's rows in the ring table are padded to whole stages with
            // all-zero rows, so every stage is processed in full: a padded row
            // multiplies whatever the TMA fetched by zero (and a non-finite
            // value takes the predicated path, which skips zero weights).
            for (int rbase = 0; rbase < band.nrows; rbase += ROWS) {
                const unsigned char* sb = stream_smem + stage * S::STAGE_B;
                mbar_wait(bars + 8 * stage, phase);
                // The stage's rows of this thread.  JIT (float / half rings of 6-8
                // rows, where the accumulators leave no room): a row is loaded
                // where it is used -- the specialised stage code is straight-line,
                // so ptxas hoists the loads as far as the registers allow; the
                // other instantiations fetch the whole stage up front.
                auto load_raw_at = [&](int j) -> uint4 {
                    const unsigned char* a = sb + toff + j * S::BOX_ROWB;
                    if (EPT * es == 16)
                        return *reinterpret_cast<const uint4*>(a);
                    else if (EPT * es == 8) {
                        uint2 u = *reinterpret_cast<const uint2*>(a);
                        return make_uint4(u.x, u.y, 0u, 0u);
                    } else
                        return make_uint4(*reinterpret_cast<const unsigned*>(a), 0u, 0u, 0u);
                };
                uint4 raw[JIT ? 1 : ROWS];
                if (!JIT) {
#pragma unroll
                    for (int j = 0; j < ROWS; ++j)
                        raw[JIT ? 0 : j] = load_raw_at(j);
                }
                auto load_raw = [&](int j) -> uint4 { return JIT ? load_raw_at(j) : raw[JIT ? 0 : j]; };
                // ring-table row j of the stage (broadcast loads)
                auto load_w = [&](int j, float (&w)[WROW]) {
#pragma unroll
                    for (int k = 0; k < WROW / 4; ++k) {
                        const float4 t = *reinterpret_cast<const float4*>(
                            sb + S::DATA_B + j * S::WROWB + k * 16);
                        w[4 * k]     = t.x;
                        w[4 * k + 1] = t.y;
                        w[4 * k + 2] = t.z;
                        w[4 * k + 3] = t.w;
                    }
                };
                // acc[k] (+)= w[k] * v, reading acc[k + 1] when the oldest row
                // was parked after the previous source row
                auto fma_row = [&](const float (&w)[WROW], const float4 (&v)[E2], bool shifted) {
                    if (shifted) {
#pragma unroll
                        for (int k = 0; k < VK; ++k) {
#pragma unroll
                            for (int h = 0; h < E2; ++h) {
                                const float4 c = (k + 1 < VK) ? acc[k + 1][h]
                                                              : make_float4(0.f, 0.f, 0.f, 0.f);
                                const float2 ww = make_float2(w[k], w[k]);
                                float2 lo = __ffma2_rn(ww, make_float2(v[h].x, v[h].y),
                                                       make_float2(c.x, c.y));
                                float2 hi = __ffma2_rn(ww, make_float2(v[h].z, v[h].w),
                                                       make_float2(c.z, c.w));
                                acc[k][h] = make_float4(lo.x, lo.y, hi.x, hi.y);
                            }
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < VK; ++k)
#pragma unroll
                            for (int h = 0; h < E2; ++h)
                                fma4(acc[k][h], w[k], v[h]);
                    }
                };
                // Inf/NaN in this thread's group: never let it meet a zero
                // weight; a weight that stands for clamped taps of both signs
                // turns it into NaN (w1*Inf + w2*Inf).  Rare, divergent.
                auto slow_row = [&](const float (&w)[WROW], const float4 (&v)[E2], int info,
                                    bool shifted) {
                    if (shifted)
                        rotate();
#pragma unroll
                    for (int k = 0; k < VK; ++k) {
#pragma unroll
                        for (int h = 0; h < E2; ++h) {
                            if (w[k] != 0.0f)
                                fma4(acc[k][h], w[k], v[h]);
                            if ((info >> (8 + k)) & 1)
                                fma4(acc[k][h], 0.0f, v[h]);
                        }
                    }
                };
                // Stage descriptor (slot VK of the stage's first table row): bits
                // 0..ROWS-1 = exactly one destination row completes with row j;
                // 0x100 = irregular stage (several rows complete at once).
                // Regular stages of the wide geometry run straight-line code
                // specialised on (shift at entry, completion mask): no per-row
                // flag decoding, no shift / completion branches; float / half
                // stages are screened for Inf/NaN as a whole first.  Measured:
                // C2 96.7 -> 85.8 us, C5 36.0 -> 32.2 us.  Float / half rings of
                // 6 rows only fit the specialised code in 64 registers when the
                // stage's rows are loaded where they are used (JIT) -- with the
                // rows held across the stage ptxas spills and C0 loses 10 us;
                // with JIT it gains 0.6 %.  Rings of 8 rows spill either way and
                // keep the row-by-row path.
                int sdesc = 0x100;
                if (SPEC)
                    sdesc = __float_as_int(
                        *reinterpret_cast<const float*>(sb + S::DATA_B + VK * 4));
                bool slow = (sdesc & 0x100) != 0;
                if (FP && SPEC) {
                    // any Inf/NaN in this warp's rows of the stage -> the per-row
                    // path with its zero-weight handling; otherwise the stage is
                    // straight-line code with no per-row branches at all
                    if (ST == ST_HALF && E2 == 1) {
                        // on the raw words: x * 0 is NaN exactly for Inf / NaN, so
                        // one packed half FMA per word leaves q at +0 unless the
                        // stage holds a non-finite value
                        __half2 q = __floats2half2_rn(0.f, 0.f);
                        const __half2 z = q;
#pragma unroll
                        for (int j = 0; j < ROWS; ++j) {
                            const uint4 rr = load_raw(j);
                            q = __hfma2(*reinterpret_cast<const __half2*>(&rr.x), z, q);
                            q = __hfma2(*reinterpret_cast<const __half2*>(&rr.y), z, q);
                        }
                        const unsigned qb = *reinterpret_cast<const unsigned*>(&q);
                        slow = slow || __any_sync(0xffffffffu, (qb & 0x7fff7fffu) != 0u);
                    } else {
                        float q = 0.0f;
#pragma unroll
                        for (int j = 0; j < ROWS; ++j) {
                            float4 v[E2];
                            stream_convert<ST, E2>(load_raw(j), v);
#pragma unroll
                            for (int h = 0; h < E2; ++h)
                                q += (v[h].x + v[h].y) + (v[h].z + v[h].w);
                        }
                        slow = slow
                               || __any_sync(0xffffffffu, !(fabsf(q) <= 3.402823466e+38f));
                    }
                }
                if (SPEC && !slow) {
                    auto fast = [&](auto code) {
                        constexpr int C = decltype(code)::value;
#pragma unroll
                        for (int j = 0; j < ROWS; ++j) {
                            // shift state is static: at entry from C, then from the mask
                            constexpr int dummy = 0;
                            (void)dummy;
                            const bool shifted = j == 0 ? ((C >> ROWS) & 1) != 0
                                                        : ((C >> (j - 1)) & 1) != 0;
                            float w[WROW];
                            load_w(j, w);
                            float4 v[E2];
                            load_values(load_raw(j), v);
                            fma_row(w, v, shifted);
                            if ((C >> j) & 1) {
                                const int info = __float_as_int(w[WROW - 1]);
                                park_oldest((info >> 16) & 1, (info >> 17) & 1);
                            }
                        }
                        shift = ((C >> (ROWS - 1)) & 1) != 0;
                    };
                    const int code = (sdesc & ((1 << ROWS) - 1)) | (shift ? (1 << ROWS) : 0);
                    static_assert(ROWS <= 3 || W2 == 1, "16 specialisations cover three rows");
                    switch (code) {
                    case 0: fast(std::integral_constant<int, 0>{}); break;
                    case 1: fast(std::integral_constant<int, 1>{}); break;
                    case 2: fast(std::integral_constant<int, 2>{}); break;
                    case 3: fast(std::integral_constant<int, 3>{}); break;
                    case 4: fast(std::integral_constant<int, 4>{}); break;
                    case 5: fast(std::integral_constant<int, 5>{}); break;
                    case 6: fast(std::integral_constant<int, 6>{}); break;
                    case 7: fast(std::integral_constant<int, 7>{}); break;
                    case 8: fast(std::integral_constant<int, 8>{}); break;
                    case 9: fast(std::integral_constant<int, 9>{}); break;
                    case 10: fast(std::integral_constant<int, 10>{}); break;
                    case 11: fast(std::integral_constant<int, 11>{}); break;
                    case 12: fast(std::integral_constant<int, 12>{}); break;
                    case 13: fast(std::integral_constant<int, 13>{}); break;
                    case 14: fast(std::integral_constant<int, 14>{}); break;
                    default: fast(std::integral_constant<int, 15>{}); break;
                    }
                } else {
#pragma unroll
                for (int j = 0; j < ROWS; ++j) {
                    float w[WROW];
                    load_w(j, w);
                    // bits 0-7: rows completing with this source row; 8-15:
                    // slots whose weight merges clamped taps of both signs;
                    // 16 / 17: the (single) completing row opens / closes a batch
                    const int info = __float_as_int(w[WROW - 1]);
                    float4 v[E2];
                    load_values(load_raw(j), v);
                    bool bad = false;
                    if (FP) {
                        float q = 0.0f;
#pragma unroll
                        for (int h = 0; h < E2; ++h)
                            q += (v[h].x + v[h].y) + (v[h].z + v[h].w);
                        bad = !(fabsf(q) <= 3.402823466e+38f);
                    }
                    if (FP && bad)
                        slow_row(w, v, info, shift);
                    else
                        fma_row(w, v, shift);
                    if (FP)
                        __syncwarp();
                    // destination rows that complete with this source row
                    int ne = info & 0xff;
                    shift  = ne > 0;
                    if (ne == 1) {
                        park_oldest((info >> 16) & 1, (info >> 17) & 1);
                    } else if (ne > 1) {
#pragma unroll 1
                        for (;;) {
                            park_oldest(ndone % BATCH == 0,
                                        ndone % BATCH == BATCH - 1 || ndone == band.ndy - 1);
                            if (--ne == 0)
                                break;
                            rotate();
                        }
                    }
                }
                }
                // stage consumed: hand it back to the producer
                __syncwarp();
                if (lane == 0)
                    mbar_arrive(bars + 8 * (NST + stage));
                if (++stage == NST) {
                    stage = 0;
                    phase ^= 1;
                }
            }
        }
    } else {
        // ================= H warps =========================================
        if (G::HREGS > G::LREGS)
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(G::HREGS));
        const int tid = threadIdx.x - G::NVT;
        const int hp  = tid % G::HPAIRS;
        const int hr  = tid / G::HPAIRS;
        int nbatch    = 0;
        // total batches of this CTA: the last two hand-backs have no taker
        int total = 0;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x)
            total += (p.bands[(item % per_frame) / p.nstrips].ndy + BATCH - 1) / BATCH;
        for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
            const int frame   = item / per_frame;
            const int fitem   = item - frame * per_frame;
            const Strip strip = p.strips[fitem % p.nstrips];
            const Band band   = p.bands[fitem / p.nstrips];
            unsigned char* dbase = p.batch_dst ? p.batch_dst[frame] : p.dst;
            const bool pair_active = 2 * hp < strip.ndx;
            const bool has_b       = 2 * hp + 1 < strip.ndx;
            const int dxA          = strip.dx0 + 2 * hp;
            const int gp           = (strip.dx0 >> 1) + hp;  // pair index in the ROI
            float2 hw[HT];
            int hb[4] = { 0, 0, 0, 0 };
            int hmix  = 0;
            if (pair_active) {
                if (FP)
                    hmix = __ldg(p.hmix + gp);
                const float2* g = reinterpret_cast<const float2*>(p.hw) + (size_t)gp * HT;
#pragma unroll
                for (int j = 0; j < HT; ++j)
                    hw[j] = __ldg(g + j);
                const int q = __ldg(p.hfirst + gp) - strip.e0 / NCH;  // window start in the strip
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    hb[k] = stream_planar_byte<W2>(q + k) + hr * G::PITCH_B;
            } else {
#pragma unroll
                for (int j = 0; j < HT; ++j)
                    hw[j] = make_float2(0.f, 0.f);
            }
            const int dy_end = band.dy0 + band.ndy;
            for (int dyb = band.dy0; dyb < dy_end; dyb += BATCH) {
                const int nb = min(BATCH, dy_end - dyb);
                const int b  = nbatch & 1;
                named_bar_sync(1 + b, NSYNC);
                if (pair_active)
                    stream_hpass<NCH, HT, W2, NR, FP, HC>(p, dbase, vbuf0 + b * S::VBUF_B, nb, dyb,
                                                      dxA, has_b, hb, hw, hmix, hr);
                ++nbatch;
                if (nbatch + 2 <= total)
                    named_bar_arrive(3 + b, NSYNC);
            }
        }
    }
}

}  // namespace b2r
