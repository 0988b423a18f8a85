// kernels.h -- launch interface between the host planner (b2r_api.cpp) and the
// sm_100a kernels (kernels.cu).  Plain structs, device pointers.
#pragma once

#include <cstdint>
#if defined(__CUDACC__)
#    include <cuda_runtime.h>
#endif

#if defined(__CUDACC__)
struct CUtensorMap_st;  // <cuda.h>
#endif

namespace b2r {

// Programmatic dependent launch (PDL): every resize kernel is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization and starts with
//     griddepcontrol.launch_dependents   -- the next kernel may be scheduled now
//     griddepcontrol.wait                -- predecessor grids complete + flushed
// so the launch latency and CTA ramp-up of kernel n+1 overlap the tail of
// kernel n (the fused kernel and the gated exact kernel behind it, the levels
// of a MIP chain, back-to-back frames).  No global memory is touched before
// the wait.  Both instructions are no-ops in a kernel launched without the
// attribute.
#if defined(__CUDACC__)
__device__ __forceinline__ void
pdl_prologue()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// Launch with programmatic stream serialization allowed.
template<class... KArgs, class... Args>
static inline void
launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
           cudaStream_t stream, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim            = grid;
    cfg.blockDim           = block;
    cfg.dynamicSmemBytes   = smem;
    cfg.stream             = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs    = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

enum : int { BT_UINT8 = 2, BT_UINT16 = 4, BT_HALF = 10, BT_FLOAT = 11 };

inline int
basetype_size(int bt)
{
    return bt == BT_UINT8 ? 1 : (bt == BT_FLOAT ? 4 : 2);
}

// Image as the kernels see it: `pixels` is the data-window origin.
struct DevImage {
    void* pixels;
    int basetype;
    int nchannels;
    int x, y, width, height;
    int full_x, full_y, full_width, full_height;
    long long xstride, ystride;  // bytes
};

// ----------------------------------------------------------------- wide ----
// Two plain passes over the gather tables for windows of any width
// (wide_kernel.cu).  Table indices are relative to the ROI's first row /
// column; source indices to the data window.
struct WideParams {
    const unsigned char* src;  // data-window row 0, column 0 (may be virtual for staged rows)
    long long src_ystride, src_xstride;
    int srctype, src_es;
    unsigned char* dst;        // first pixel of the ROI
    long long dst_ystride, dst_xstride;
    int dsttype, dst_es;
    int nch, roi_w, roi_h;
    float* tmp;                // [roi_h][tmp_w][nch]
    int x0, tmp_w;             // source columns [x0, x0 + tmp_w) the H pass reads
    const int *yfirst, *ycount, *ywoff;
    const float* yw;
    const unsigned char* ymixed;
    const int *xfirst, *xcount, *xwoff;
    const float* xw;
    const unsigned char* xmixed;
};
void
launch_resize_wide(const WideParams& p, void* stream);

// ---------------------------------------------------------------- exact ----
// Single-pass kernel in the reference's own tap order (bit-exact path).
struct ExactParams {
    DevImage dst, src;
    int roi_x0, roi_x1, roi_y0, roi_y1;  // dst coords
    int nch;                             // channels to compute (= dst.nchannels)
    int radi, radj, xtaps, ytaps;
    float xratio, yratio;
    // separable: virtual taps
    const int* xcenter;    // [roi_w]
    const float* xw;       // [roi_w * xtaps] normalised
    const float* xwsum;    // [roi_w]
    const int* ycenter;    // [roi_h]
    const float* yw;       // [roi_h * ytaps]
    const float* ywsum;    // [roi_h]
    // non-separable: per-index fractions + filter description
    const float* xfrac;    // [roi_w]
    const float* yfrac;    // [roi_h]
    int nonsep_kind;       // 0 separable, 1 disk, 2 radial-lanczos3
    float filt_w, filt_h;
    // when not null: run only if *gate == gate_value (non-finite redo)
    const int* gate;
    int gate_value;
    // gated runs re-do only the cells of the ROI the fused kernel marked:
    // cell (cx, cy) = 128 columns x (1 << dirty_yshift) rows, marked when
    // dirty[cy * dirty_xcells + cx] == gate_value
    const int* dirty;
    int dirty_xcells, dirty_yshift;
};
void
launch_resize_exact(const ExactParams& p, void* stream);

// ---------------------------------------------------------------- fused ----
// Two-pass (vertical then horizontal) kernel; the intermediate stays in
// shared memory.  One CTA = one column strip x one row band of the dst ROI.
struct Strip {
    int dx0;    // first dst column (index into the ROI, 0-based)
    int ndx;    // columns in this strip (<= FUSED_HCOLS; even except the last)
    int e0;     // first flat source element (px*nch + c) staged, multiple of 4
    int nvec;   // number of 4-element groups staged (<= FUSED_THREADS)
};
struct Band {
    int dy0;    // first dst row (index into the ROI)
    int ndy;    // rows
    int r0;     // first source row consumed (relative to src data origin)
    int nrows;  // source rows consumed (ring mode)
    int woff;   // ring mode: row offset of this band's weights in vring
    int soff;   // ring mode: row offset of this band in vstream (bands padded to
                // whole stages there)
};

constexpr int FUSED_THREADS = 256;
constexpr int FUSED_HCOLS   = 128;  // dst columns per strip (H-pass threads per row)
// V-filtered rows buffered between the passes: a multiple of the ring size so
// the H pass sits outside the slot-unrolled row loop (one copy in the code)
constexpr int
fused_batch(int vk)
{
    return vk == 6 ? 12 : 8;
}
constexpr int FUSED_FIFO    = 8;    // source rows in flight per thread (power of 2)
constexpr int FUSED_MAX_HT  = 128;  // widest H window (taps) of the run-time-HT variant
#if defined(__CUDACC__)
#    define B2R_HD __host__ __device__
#else
#    define B2R_HD
#endif
// Per (element size, ring size): rows in flight per thread and resident CTAs
// the kernel is compiled for.  Measured on C2 (half, VK = 4, 100.0 us): 4 rows
// with 4 CTAs/SM 104.0 us, 16 rows 105.1 us -- 8 rows / 3 CTAs stays.
B2R_HD constexpr int
fused_fifo_rows(int es, int vk)
{
    return (void)es, (void)vk, FUSED_FIFO;
}
B2R_HD constexpr int
fused_min_ctas(int es, int vk)
{
    return (void)es, (void)vk, 3;
}

// Stream variant (stream_kernel.cuh): persistent, warp-specialised -- one TMA
// producer warp, w2 * STREAM_VTHREADS V-pass threads, STREAM_HTHREADS H-pass
// threads per CTA, one CTA per SM.
constexpr int STREAM_VTHREADS = 256;  // V threads per unit of strip width (w2)
constexpr int STREAM_HTHREADS = 256;
// V threads of a stream-kernel CTA.  A V row holds 256 * w2 pixels, i.e.
// 64 * w2 * nch four-element groups: with fewer than four channels a CTA of
// 256 * w2 V threads would run the rest of its warps on padding.  Whole
// warpgroups only (setmaxnreg is executed per warpgroup, and the H warps that
// follow must start on a warpgroup boundary).
B2R_HD constexpr int
stream_vthreads(int nch, int w2)
{
    return (64 * w2 * (nch < 1 ? 1 : (nch > 4 ? 4 : nch)) + 127) / 128 * 128;
}
// warpgroup 0 holds the producer warp (its other three warps give their
// registers away with setmaxnreg and exit), then the V warps, then the H warps
constexpr int STREAM_PTHREADS = 128;
constexpr int STREAM_SMEM_MAX = 227 * 1024;
// a band's rows in the stream kernel's ring table are padded to a multiple of
// this (whole stages for either geometry)
constexpr int STREAM_ROWPAD = 12;
// source rows per TMA stage (w2 = 1: 4 rows of 256 chunks; w2 = 2: 3 rows of 512)
B2R_HD constexpr int
stream_rows(int w2)
{
    return w2 == 1 ? 4 : 3;
}
// V-filtered rows per hand-over batch
B2R_HD constexpr int
stream_batch(int vk, int w2)
{
    return (void)vk, w2 == 1 ? 12 : 6;
}
B2R_HD constexpr int
stream_pitch_bytes(int w2)
{
    return 4 * (STREAM_VTHREADS * w2 / 4 + 2) * 16;
}
// floats per source row of the stream kernel's ring table: vk weights (oldest
// open destination row first) + the completion count, padded to float4 loads
B2R_HD constexpr int
stream_wrow(int vk)
{
    return (vk + 1 + 3) & ~3;
}
B2R_HD constexpr int
stream_stage_bytes(int es, int vk, int w2)
{
    return (STREAM_VTHREADS * w2 * 4 * es * stream_rows(w2)
            + stream_rows(w2) * stream_wrow(vk) * 4 + 127)
           & ~127;
}
// stages of the TMA ring: what fits beside the two V-row buffers, at most 8
B2R_HD constexpr int
stream_stages(int es, int vk, int w2)
{
    return (STREAM_SMEM_MAX - 2 * stream_batch(vk, w2) * stream_pitch_bytes(w2) - 1024)
                       / stream_stage_bytes(es, vk, w2)
                   >= 8
               ? 8
               : (STREAM_SMEM_MAX - 2 * stream_batch(vk, w2) * stream_pitch_bytes(w2) - 1024)
                     / stream_stage_bytes(es, vk, w2);
}

// Expand variant (upsizing in y: the V pass gathers): strips are up to
// EXPAND_GROUPS groups of four dst columns wide, a group per H-pass thread.
constexpr int EXPAND_GROUPS = 256;
constexpr int EXPAND_BATCH  = 8;    // V-filtered rows between the passes
constexpr int EXPAND_MAXVT  = 8;    // widest V window (rows held in registers)
constexpr int EXPAND_FIFO   = 4;    // raw source rows in flight per thread

constexpr int DIRTY_CELL_SHIFT_X = 7;  // 128 dst columns per dirty-map cell

struct FusedParams {
    const unsigned char* src;  // data-window origin, pixels contiguous in x
    long long src_ystride;
    int srctype, src_w, src_h, nch;
    unsigned char* dst;  // address of the ROI's first pixel
    long long dst_ystride;
    int dsttype;
    int src_vec_ok, dst_vec_ok;  // 4-element vector loads / stores allowed
    int expand;                  // 1: resize_expand_kernel tables (groups of 4 columns)
    int dst_align;               // expand: 16 / 4 / 0 = alignment dst rows guarantee
    int dst_pow2;                // stream: largest power of two (<= 16) dividing the ROI
                                 // origin's address and the dst row pitch
    int vrow_linear;             // expand: V rows laid out linearly, not in planes
    int check_nonfinite;         // float/half sources: flag non-finite outputs
    int* nonfinite_flag;
    int nonfinite_epoch;         // value written to the flag (unique per call)
    // ... and to the cell of the dirty map that holds the offending pixel
    // (128 columns x (1 << dirty_yshift) rows of the ROI per cell)
    int* dirty;
    int dirty_xcells, dirty_yshift;

    int nstrips, nbands;
    const Strip* strips;
    const Band* bands;

    // H pass: two adjacent dst columns (A, B) share a window of HT source
    // pixels; per pair the window start and per tap the weights (wA, wB)
    const int* hfirst;  // [ceil(roi_w/2)] first source pixel of the window
    const float* hw;    // [ceil(roi_w/2) * HT * 2]
    // expand: hfirst[ceil(roi_w/4)], hw[(tap * ngroups + group) * 4 + column]
    int ngroups;
    int ht;             // window taps (== template HT)

    // V pass, ring mode (vk > 0): per band, per source row, K slot weights
    int vk;
    const float* vring;  // [(woff + r - r0) * vk + slot]
    const int* vlast;    // [roi_h] source row after which dst row completes
    const float* vstream;  // stream kernel: [(woff + r - r0) * stream_wrow(vk) + i]
    const int* hmix;       // stream kernel: [ceil(roi_w/2)] mixed-sign merged H taps
    // V pass, gather mode (vk == 0)
    int vt;
    const int* vfirst;   // [roi_h]
    const float* vw;     // [roi_h * vt]
    // stream kernel: data-window row that is row 0 of the TMA tensor map
    int tma_row0;
    // stream kernel, float -> float: highlight compensation fused in --
    // rangecompress on the source values as the V pass loads them, rangeexpand
    // on the results as the H pass stores them (oiiotool.cpp:4644-4663,
    // maketexture.cpp:907-935); bit c of hc_skip = channel c is left alone
    // (alpha, z)
    int hicomp;
    int hc_skip;
    // ... followed by maketx's clamp (maketexture.cpp:936-938): every channel
    // to [0, FLT_MAX] (NaN -> 0), channel hc_alpha (>= 0) to [0, 1]
    int hc_clamp;
    int hc_alpha;
    // stream kernel, a batch of frames of one shape in one launch (nframes > 1):
    // per frame a tensor map (device array of CUtensorMap, 64-byte aligned) and
    // the address of the ROI's first pixel
    int nframes;
    const void* batch_maps;
    unsigned char* const* batch_dst;
};
// Returns false when no instantiation matches (srctype, nch, vk, ht).
bool
fused_supported(int srctype, int nch, int vk, int ht);
bool
expand_supported(int srctype, int nch, int vt, int ht);
typedef void (*FusedKernel)(const FusedParams, int, int);
// stream kernel (TMA): ring-mode shapes with a compile-time H window whose
// source rows are 16-byte aligned.  `w2`: 1 or 2 (4 or 8 elements per V thread).
#if defined(__CUDACC__)
typedef void (*StreamKernel)(const ::CUtensorMap_st, const FusedParams);
#endif
bool
stream_supported(int srctype, int nch, int vk, int ht, int w2, int hicomp = 0);
// tma_base: address of element 0 of data-window row p.tma_row0; tma_rows: rows
// that exist from there on.  Returns 0, or a negative value when the tensor
// map could not be made (the caller falls back to the fused kernel).
int
launch_resize_stream(const FusedParams& p, int w2, const void* tma_base, int tma_rows,
                     int sms, void* stream);
// Fill `map` (128 bytes, 64-byte aligned) with the tensor map the stream kernel
// reads source rows through.  0 on success.
int
stream_encode_map(void* map, const FusedParams& p, int w2, const void* tma_base,
                  int tma_rows);
// nframes > 1: p.batch_maps / p.batch_dst are device arrays (one entry per frame)
int
launch_resize_stream_batch(const FusedParams& p, int w2, int sms, void* stream);
typedef void (*ExpandKernel)(const FusedParams, int);
// resident CTAs per SM for this shape (occupancy query), 0 on error
int
fused_ctas_per_sm(const FusedParams& p, int max_band_rows, int max_band_dy,
                  int max_nvec);
size_t
fused_smem_bytes(const FusedParams& p, int max_band_rows, int max_band_dy,
                 int max_nvec);
void
launch_resize_fused(const FusedParams& p, int max_band_rows, int max_band_dy,
                    int max_nvec, void* stream);

// ------------------------------------------------------------- resample ----
struct ResampleParams {
    DevImage dst, src;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int nch;
    int interpolate;
    int data_is_full;
};
void
launch_resample(const ResampleParams& p, int chbegin, int chend, void* stream);

// ------------------------------------------------------------- box mip ----
// maketx resize_block (maketexture.cpp:141-334) on float32 images.
struct BoxParams {
    DevImage dst, src;
    int two_pass;  // 1: exact 2x2 average path (:260-300); 0: bilinear NDC
    int force_clamp;  // bilinear: WrapClamp even if src's data window is a crop
                      // (a row band of an uncropped level)
};
void
launch_box_downsample(const BoxParams& p, void* stream);

// clamp a float image to [lo, hi] in place (ImageBufAlgo::clamp with
// clampalpha01=true as maketx calls it, maketexture.cpp:943-945).
void
launch_clamp(float* pixels, long long n, int nch, int alpha_channel, float lo,
             float hi, void* stream);

// rangecompress / rangeexpand (imagebufalgo_pixelmath.cpp:560-735) over a ROI.
struct RangeParams {
    DevImage dst, src;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int chbegin, chend;
    int expand;    // 0: rangecompress, 1: rangeexpand
    int useluma;   // caller has already applied the reference's "no way to use luma" rule
    int alpha_channel, z_channel;  // copied through untouched (-1: none)
    int in_place;  // dst and src are the same pixels: skipped channels are not rewritten
};
void
launch_range(const RangeParams& p, void* stream);

// ----------------------------------------------------------------- warp ----
// warp_<D,S> (imagebufalgo_xform.cpp:353-375): dst pixel -> Minv -> filtered
// sample of src with a per-pixel footprint.
struct WarpParams {
    DevImage dst, src;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int chbegin, chend;
    float minv[9];  // inverse matrix, [i*3+j] = Imath x[i][j]
    int filt_profile, filt_scale;  // Filter2D::DeviceDesc
    float filt_param, filt_w, filt_h;
    int wrap;       // 0 default, 1 black, 2 clamp, 3 periodic, 4 mirror
    int edgeclamp;
    int contract_fma;  // accumulate with packed FMAs (tolerance mode)
};
void
launch_warp(const WarpParams& p, void* stream);

// st_warp_<D,S,ST> (imagebufalgo_xform.cpp:1834-1925)
struct StWarpParams {
    DevImage dst, src, st;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int chbegin, chend;
    int chan_s, chan_t, flip_s, flip_t;
    int filterrad_x, filterrad_y;  // ceil(filter extent / 2 / scale), source pixels
    int filt_profile, filt_scale;
    float filt_param, filt_w, filt_h;
};
void
launch_st_warp(const StWarpParams& p, void* stream);

// ------------------------------------------------- convolve / unsharp ----
// convolve_ (imagebufalgo.cpp:772-812): dst = scale * sum_k kernel[k] * src
struct ConvParams {
    DevImage dst, src;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int chbegin, chend;
    const float* kernel;  // device, kh rows of kw; window origin (-kw/2, -kh/2)
    int kw, kh;
    float scale;          // 1 / sum(kernel) when normalising, else 1
    // unsharp != 0: store src + contrast * (src - blur) where |src - blur| >
    // threshold instead of the blur (unsharp_mask in one pass, no blur image)
    int unsharp;
    float contrast, threshold;
};
void
launch_convolve(const ConvParams& p, void* stream);
// unsharp_impl (imagebufalgo.cpp:961-985); blur is float32 with src's windows
struct UnsharpParams {
    DevImage dst, src, blur;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int chbegin, chend;
    float contrast, threshold;
};
void
launch_unsharp_combine(const UnsharpParams& p, void* stream);

// fixNonFinite_ (imagebufalgo_pixelmath.cpp:1424-1500), float / half images of
// one type.  mode 0 / 100: count only, 1: black, 2: 3x3 average of the finite
// neighbours.  write == 0: nothing is stored (pure count).
struct NonFiniteParams {
    DevImage dst, src;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int chbegin, chend;
    int mode, write;
    int* counter;  // device; += pixels that held a non-finite value
};
void
launch_fix_nonfinite(const NonFiniteParams& p, void* stream);

// fillholes_pushpull's point ops on contiguous float images
// (imagebufalgo.cpp:1578-1596; imagebufalgo_pixelmath.cpp:1687-1692)
void
launch_divide_by_alpha(float* px, long long npixels, int nch, int alpha_channel, void* stream);
void
launch_over_inplace(float* big, const float* blowup, long long npixels, int nch,
                    int alpha_channel, void* stream);

// computePixelStats + isMonochrome in one pass (imagebufalgo_compare.cpp:81-215,
// 539-600).  One partial per (block, channel), merged on the host.
struct StatsPartial {
    double sum, sum2;
    float mn, mx;
    unsigned long long finite, nan, inf;
};
struct StatsParams {
    DevImage src;
    int roi_x0, roi_x1, roi_y0, roi_y1;
    int chbegin, chend, nchannels;
    StatsPartial* partials;  // [nblocks * nchannels]
    int want_mono;
    float mono_threshold;
    int* mono_flags;         // [nblocks]
};
int
stats_grid_blocks(int rows);
void
launch_pixel_stats(const StatsParams& p, int nblocks, void* stream);

// copy `nbytes` bytes of every pixel (a channel group) between two images with
// independent pixel / row strides
void
launch_pixel_bytes_copy(void* dst, long long dst_xstride, long long dst_ystride,
                        const void* src, long long src_xstride, long long src_ystride,
                        int width, int height, int nbytes, void* stream);

// tile rows [ty0, ty1) of a contiguous float level -> tiles (tile_kernel.cu)
void
launch_tile_relayout(const float* level, int w, int h, int nch, int tile_w, int tile_h,
                     int ty0, int ty1, int out_half, void* out, void* stream);
// uint8 / uint16 tiles (convert_type<float,D> of the level), optionally with TIFF's
// horizontal predictor per tile row; tile_w * nch must be a multiple of 4
void
launch_tile_relayout_int(const float* level, int w, int h, int nch, int tile_w, int tile_h,
                         int ty0, int ty1, int out_u16, int hdiff, void* out, void* stream);
// the same with TIFF's floating-point predictor applied per tile row
// (tile_w * nch must be a multiple of 4)
void
launch_tile_relayout_fp_predictor(const float* level, int w, int h, int nch, int tile_w,
                                  int tile_h, int ty0, int ty1, int out_half, void* out,
                                  void* stream);

// float -> dst type over a rectangle of `row_elems` x `height` elements
// (the copy-back step for (dst,src) pairs computed through a float temporary)
void
launch_convert_from_float(const float* src, long long src_ystride_bytes, void* dst,
                          long long dst_ystride_bytes, int row_elems, int height,
                          int dst_basetype, void* stream);

}  // namespace b2r
