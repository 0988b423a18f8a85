/* b2r.h -- C ABI of the B200-native filtered resampler (libb2r.so).
 *
 * This is the drop-in boundary for ONE hot path of OpenImageIO 3.2: the
 * compute kernel behind ImageBufAlgo::resize / fit / resample and the
 * per-level MIP downsample of make_texture.  Every entry point names the
 * reference interface it replaces (paths relative to the OpenImageIO tree).
 * Plain C: pointers, sizes, PODs; no C++ or torch types cross it.
 *
 * There is no CPU fallback: every compute entry point fails with
 * B2R_ERR_NO_DEVICE when no CUDA device is usable.
 *
 * Conventions kept from the reference:
 *  - the caller owns all pixel memory; dst is already allocated (IBAprep,
 *    src/libOpenImageIO/imagebufalgo_xform.cpp:882-884);
 *  - functions return 0 on success, else a B2R_ERR_* code and a
 *    human-readable message in `err` (the text ImageBuf::errorfmt would have
 *    received, e.g. `Filter "foo" not recognized`, xform.cpp:857); they never
 *    throw and never abort;
 *  - calls are synchronous for host pixel memory (dst is complete on return)
 *    and stream-ordered for device pixel memory;
 *  - re-entrant: concurrent calls on distinct dst images are allowed.
 */
#ifndef B2R_H
#define B2R_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#    define B2R_API __declspec(dllexport)
#else
#    define B2R_API __attribute__((visibility("default")))
#endif

/* TypeDesc::BASETYPE values, src/include/OpenImageIO/typedesc.h:57-84 */
enum b2r_basetype {
    B2R_UINT8  = 2,
    B2R_UINT16 = 4,
    B2R_HALF   = 10,
    B2R_FLOAT  = 11
};

enum b2r_memspace {
    B2R_MEM_AUTO   = 0, /* ask the CUDA runtime (cudaPointerGetAttributes) */
    B2R_MEM_HOST   = 1, /* pageable or pinned host memory */
    B2R_MEM_DEVICE = 2  /* device memory on `device` */
};

enum b2r_error {
    B2R_OK                = 0,
    B2R_ERR_INVALID       = 1, /* bad argument / uninitialized image */
    B2R_ERR_FILTER        = 2, /* Filter "..." not recognized */
    B2R_ERR_TYPES         = 3, /* pixel type outside uint8/uint16/half/float */
    B2R_ERR_NO_DEVICE     = 4, /* no usable CUDA device: there is no fallback */
    B2R_ERR_CUDA          = 5, /* CUDA runtime error (message has details) */
    B2R_ERR_UNSUPPORTED   = 6  /* volumes / deep images */
};

/* What resize_<D,S> reads from an ImageBuf + ImageSpec
 * (src/libOpenImageIO/imagebufalgo_xform.cpp:601-625, imagebuf.h:1383-1385). */
typedef struct b2r_image {
    void* pixels;      /* address of the data-window origin pixel (x,y) */
    int32_t basetype;  /* enum b2r_basetype */
    int32_t nchannels;
    int32_t x, y, width, height;                      /* data window */
    int32_t full_x, full_y, full_width, full_height;  /* display window */
    int64_t xstride, ystride; /* bytes: pixel_stride(), scanline_stride() */
    int32_t memspace;         /* enum b2r_memspace */
    int32_t device;           /* CUDA ordinal when memspace == DEVICE */
} b2r_image;

/* ROI, src/include/OpenImageIO/imageio.h:93-108 (z dropped: no volumes). */
typedef struct b2r_roi {
    int32_t xbegin, xend, ybegin, yend, chbegin, chend;
} b2r_roi;

enum b2r_path {
    B2R_PATH_AUTO  = 0, /* fused two-pass kernel when eligible, else exact */
    B2R_PATH_EXACT = 1, /* single-pass kernel in the reference's tap order */
    B2R_PATH_FUSED = 2  /* fail (B2R_ERR_UNSUPPORTED) if not eligible */
};

typedef struct b2r_options {
    int32_t device;    /* CUDA ordinal for host-memory calls (default 0) */
    int32_t path;      /* enum b2r_path */
    void* cuda_stream; /* cudaStream_t; NULL = default stream */
    int32_t reserved[4];
    /* Highlight compensation, as oiiotool --resize/--fit ":highlightcomp=1"
     * (src/oiiotool/oiiotool.cpp:4636-4665) and maketx --hicomp
     * (maketexture.cpp:907-935) wrap it around the resize: rangecompress the
     * source into a temporary of its own type, resize, rangeexpand the
     * result in place -- all on the device.  alpha_channel1 / z_channel1 are
     * 1 + the index of the channel rangecompress must leave alone
     * (ImageSpec::alpha_channel / z_channel), 0 = none. */
    int32_t highlightcomp;
    int32_t alpha_channel1, z_channel1;
    /* Tolerance mode for the warp family (b2r_warp*, b2r_fit_exact's warp):
     * 0 (default) accumulates `sum += w * pixel` as the reference's compiled
     * code does, multiply and add rounded separately, so the footprint AND the
     * sums follow its rounding; 1 lets the kernel use packed fused
     * multiply-adds (a quarter of the floating-point instructions of an
     * issue-bound kernel): results stay within 1e-5 relative / 1 LSB of the
     * reference, the footprint arithmetic is unchanged. */
    int32_t contract_fma;
} b2r_options;

/* Filled on request: what ran, for tests and benchmarks. */
typedef struct b2r_report {
    int32_t path_used;       /* B2R_PATH_EXACT or B2R_PATH_FUSED */
    int32_t kernels_launched;
    int32_t xtaps, ytaps;    /* reference tap counts 2*radi+1, 2*radj+1 */
    int32_t fused_vmode;     /* ring size K (>0), 0 = gather, -VT = expand kernel
                                (upsizing in y, VT-row register window) */
    int32_t fused_htaps;
    int64_t h2d_bytes, d2h_bytes;
    float kernel_ms;         /* CUDA-event time of the kernels, host calls only */
    int32_t nonfinite_redo;  /* 1 if the exact kernel re-did the image */
} b2r_report;

/* ---- library ---------------------------------------------------------- */
B2R_API const char* b2r_version(void);
/* The library keeps device / pinned memory between calls so that a run of
 * same-sized jobs allocates once: the band buffers of destroyed banded chains
 * (up to 12 GB per device), the tile encoder's staging slots, the weight-table
 * plans.  This call gives all of it back to the driver (no call may be in
 * flight).  Host weight tables are rebuilt on the next call of each shape. */
B2R_API void b2r_trim_caches(void);
/* Number of usable CUDA devices (0 = none; compute calls will fail). */
B2R_API int b2r_device_count(void);

/* ---- Filter1D / Filter2D by name --------------------------------------
 * Replaces the statics of src/include/OpenImageIO/filter.h:59-78,130-152 as
 * implemented in src/libutil/filter.cpp:841-1047.  Host arithmetic, float32,
 * identical expression order to the reference. */
B2R_API int b2r_filter_count(int dim /*1 or 2*/);
/* FilterDesc (filter.h:19-27).  Returns 0, or B2R_ERR_INVALID. */
B2R_API int b2r_filter_desc(int dim, int index, const char** name,
                            float* width, int* fixedwidth, int* scalable,
                            int* separable);
/* Opaque filter handles: Filter2D::create / destroy (filter.cpp:1000-1047). */
typedef struct b2r_filter2d b2r_filter2d;
typedef struct b2r_filter1d b2r_filter1d;
B2R_API b2r_filter2d* b2r_filter2d_create(const char* name, float width,
                                          float height);
B2R_API void b2r_filter2d_destroy(b2r_filter2d* f);
B2R_API float b2r_filter2d_width(const b2r_filter2d* f);
B2R_API float b2r_filter2d_height(const b2r_filter2d* f);
B2R_API int b2r_filter2d_separable(const b2r_filter2d* f);
B2R_API const char* b2r_filter2d_name(const b2r_filter2d* f);
B2R_API float b2r_filter2d_eval(const b2r_filter2d* f, float x, float y);
B2R_API float b2r_filter2d_xfilt(const b2r_filter2d* f, float x);
B2R_API float b2r_filter2d_yfilt(const b2r_filter2d* f, float y);
B2R_API b2r_filter1d* b2r_filter1d_create(const char* name, float width);
B2R_API void b2r_filter1d_destroy(b2r_filter1d* f);
B2R_API float b2r_filter1d_width(const b2r_filter1d* f);
B2R_API const char* b2r_filter1d_name(const b2r_filter1d* f);
B2R_API float b2r_filter1d_eval(const b2r_filter1d* f, float x);

/* ---- resize ----------------------------------------------------------- */
/* Replaces ImageBufAlgo::resize(dst, src, KWArgs{filtername, filterwidth},
 * roi, nthreads) from the point after IBAprep:
 * src/libOpenImageIO/imagebufalgo_xform.cpp:885-917 (get_resize_filter
 * :830-860 + the OIIO_DISPATCH_COMMON_TYPES2 call into resize_<D,S> :595-826).
 * filtername NULL or "" -> lanczos3 when both ratios <= 1, else
 * blackman-harris; filterwidth <= 0 -> FilterDesc.width * max(1, ratio).
 * roi NULL -> dst's data window; otherwise intersected with it. */
B2R_API int b2r_resize(const b2r_image* dst, const b2r_image* src,
                       const char* filtername, float filterwidth,
                       const b2r_roi* roi, const b2r_options* opt,
                       b2r_report* report, char* err, size_t errlen);

/* Same, with a caller-supplied filter: the KWArgs "filterptr" form
 * (xform.cpp:888, get_filterptr_option :456-465).  The filter is borrowed. */
B2R_API int b2r_resize_filter(const b2r_image* dst, const b2r_image* src,
                              const b2r_filter2d* filter, const b2r_roi* roi,
                              const b2r_options* opt, b2r_report* report,
                              char* err, size_t errlen);

/* A batch of `nframes` independent frames (dst[i] <- src[i]) with one filter.
 * The reference resizes a frame sequence with one ImageBufAlgo::resize call per
 * frame (oiiotool --frames, imagebufalgo_xform.cpp:864-918); here frames of ONE
 * shape (equal windows, types, strides) that live on one device run as a
 * single persistent launch over (frame, strip, band) work items, which removes
 * the per-frame launch gap, ramp-up and tail -- what bounds small frames.
 * Frames that do not qualify (host memory, differing geometry, shapes of the
 * other kernels) are resized one after the other, with the same results.
 * report (optional) describes the batch launch or the last frame. */
B2R_API int b2r_resize_batch(const b2r_image* dst, const b2r_image* src, int nframes,
                             const char* filtername, float filterwidth,
                             const b2r_roi* roi, const b2r_options* opt,
                             b2r_report* report, char* err, size_t errlen);

/* ---- rangecompress / rangeexpand ---------------------------------------
 * Replaces rangecompress_<R,A> / rangeexpand_<R,A>
 * (src/libOpenImageIO/imagebufalgo_pixelmath.cpp:560-735): the log-curve
 * point ops used for highlight compensation around a resize.  Any of the
 * four pixel types on either side; dst may be src (in place).  useluma,
 * alpha_channel and z_channel (-1 = none) as in the reference; roi NULL =
 * dst data window, channel range clamped to both images
 * (IBAprep_CLAMP_MUTUAL_NCHANNELS, :763).  logf/expf are the device's
 * (<= 2 ulp from glibc's), so float results agree to ~1e-6 relative, not to
 * the bit. */
B2R_API int b2r_rangecompress(const b2r_image* dst, const b2r_image* src,
                              int useluma, int alpha_channel, int z_channel,
                              const b2r_roi* roi, const b2r_options* opt,
                              b2r_report* report, char* err, size_t errlen);
B2R_API int b2r_rangeexpand(const b2r_image* dst, const b2r_image* src,
                            int useluma, int alpha_channel, int z_channel,
                            const b2r_roi* roi, const b2r_options* opt,
                            b2r_report* report, char* err, size_t errlen);

/* ---- resample --------------------------------------------------------- */
/* Replaces resample_<D,S> (xform.cpp:1663-1687 and the kernels it picks,
 * :1107-1186, :1379-1493): interpolate=0 nearest, else bilinear. */
B2R_API int b2r_resample(const b2r_image* dst, const b2r_image* src,
                         int interpolate, const b2r_roi* roi,
                         const b2r_options* opt, b2r_report* report, char* err,
                         size_t errlen);

/* ---- fit -------------------------------------------------------------- */
/* The integer geometry of ImageBufAlgo::fit (xform.cpp:962-997): for a
 * source display window sw x sh fitted into fw x fh with fillmode
 * "letterbox" | "width" | "height", the size to resize to and the offset of
 * the result's data window.  The resize itself is b2r_resize. */
B2R_API void b2r_fit_geometry(int src_full_width, int src_full_height,
                              int fit_width, int fit_height,
                              const char* fillmode, int* resize_width,
                              int* resize_height, int* xoffset, int* yoffset);

/* ---- warp / rotate / fit(exact) / --resize:from/to -----------------------
 * Replaces warp_<D,S> + filtered_sample<S> as dispatched by warp_impl
 * (src/libOpenImageIO/imagebufalgo_xform.cpp:263-325, 353-418) -- the path
 * ImageBufAlgo::warp, ImageBufAlgo::rotate (:1727-1760), fit(exact=1)
 * (:1017-1027) and oiiotool --resize:from=/to= (src/oiiotool/oiiotool.cpp:
 * 4652-4656, 4684-4720) all end in.  M: 9 floats, row-major Imath::M33f
 * (p' = p * M, translation in M[6], M[7]).  dst must be allocated (warp_impl
 * does that through IBAprep, :397); roi NULL = dst's data window.
 * wrap: ImageBuf::WrapMode (imagebuf.h:  0 default, 1 black, 2 clamp,
 * 3 periodic, 4 mirror).  Any of the four pixel types on either side.
 * filtername NULL or "" -> lanczos3; filterwidth <= 0 -> the filter's default
 * width (get_warp_filter, :329-349). */
B2R_API int b2r_warp(const b2r_image* dst, const b2r_image* src, const float* M,
                     const char* filtername, float filterwidth, int wrap,
                     int edgeclamp, const b2r_roi* roi, const b2r_options* opt,
                     b2r_report* report, char* err, size_t errlen);
/* The KWArgs "filterptr" form (:478-480; also how fit(exact=1) passes the
 * filter it resolved with get_resize_filter).  filter NULL -> lanczos3, 6x6
 * (:403-408).  The filter is borrowed. */
B2R_API int b2r_warp_filter(const b2r_image* dst, const b2r_image* src,
                            const float* M, const b2r_filter2d* filter, int wrap,
                            int edgeclamp, const b2r_roi* roi,
                            const b2r_options* opt, b2r_report* report,
                            char* err, size_t errlen);
/* ---- st_warp -----------------------------------------------------------
 * Replaces st_warp_<D,S,ST> (imagebufalgo_xform.cpp:1834-1925; argument
 * checks :1929-1972): dst(x,y) = filtered sample of src around
 * (S * src.full_width, T * src.full_height), S/T read from channels chan_s /
 * chan_t of `st` at (x,y).  The ROI is intersected with st's data window.
 * Filter as for b2r_warp. */
B2R_API int b2r_st_warp(const b2r_image* dst, const b2r_image* src,
                        const b2r_image* st, const char* filtername,
                        float filterwidth, int chan_s, int chan_t, int flip_s,
                        int flip_t, const b2r_roi* roi, const b2r_options* opt,
                        b2r_report* report, char* err, size_t errlen);
B2R_API int b2r_st_warp_filter(const b2r_image* dst, const b2r_image* src,
                               const b2r_image* st, const b2r_filter2d* filter,
                               int chan_s, int chan_t, int flip_s, int flip_t,
                               const b2r_roi* roi, const b2r_options* opt,
                               b2r_report* report, char* err, size_t errlen);

/* Host matrix helpers, float32 in Imath's operation order (Imath 3.1
 * Matrix33<float>; Imath is an external dependency of the reference). */
B2R_API void b2r_m33_inverse(const float* M, float* out);
/* ImageBufAlgo::rotate: translate(-c) * rotate(angle) * translate(c), xform.cpp:1733-1737 */
B2R_API void b2r_rotate_matrix(float angle_radians, float center_x,
                               float center_y, float* out);
/* oiiotool --resize:from=:to= : translate(to) scale(to/from) translate(-from),
 * oiiotool.cpp:4704-4707 */
B2R_API void b2r_fromto_matrix(float from_x, float from_y, float from_w,
                               float from_h, float to_x, float to_y, float to_w,
                               float to_h, float* out);
/* transform(M, roi), xform.cpp:231-253 (warp's recompute_roi) */
B2R_API void b2r_transform_roi(const float* M, const b2r_roi* roi, b2r_roi* out);
/* fit(exact=1): the warp matrix (scale, centring offsets) and the size the
 * filter is chosen for, xform.cpp:962-997, 1005-1023 */
B2R_API void b2r_fit_exact(int src_full_width, int src_full_height,
                           int fit_width, int fit_height, const char* fillmode,
                           float* M, int* resize_width, int* resize_height);

/* ---- make_kernel / convolve / unsharp_mask ------------------------------
 * Replaces ImageBufAlgo::make_kernel, convolve_<D,S> and unsharp_mask
 * (src/libOpenImageIO/imagebufalgo.cpp:864-957, 772-838, 961-1031): the
 * sharpening maketx wraps around each MIP level's resize (--sharpen,
 * maketexture.cpp:909-931) and oiiotool --kernel / --convolve / --blur /
 * --unsharp.  Kernel tables are built on the host with the same Filter2D code
 * the resize tables use; the device multiplies and adds in the reference's tap
 * order without contraction, so results match the CPU to the bit. */
/* name: any Filter2D name, "binomial", "laplacian" (3x3).  The table is kh
 * rows of kw floats (both odd), window origin (-kw/2, -kh/2).  out NULL =
 * size query.  An unknown name fills a box AND returns B2R_ERR_FILTER with
 * `Unknown kernel "name" WxH`, as the reference returns a box plus an error. */
B2R_API int b2r_make_kernel(const char* name, float width, float height,
                            int normalize, float* out, int* kw, int* kh,
                            char* err, size_t errlen);
/* dst = (normalize ? 1/sum(kernel) : 1) * sum_k kernel[k] * src, WrapClamp at
 * the edges (:790-807).  kernel: host memory.  dst and src need the same
 * channel count (IBAprep_REQUIRE_SAME_NCHANNELS) and must not alias. */
B2R_API int b2r_convolve(const b2r_image* dst, const b2r_image* src,
                         const float* kernel, int kw, int kh, int normalize,
                         const b2r_roi* roi, const b2r_options* opt,
                         b2r_report* report, char* err, size_t errlen);
/* dst = src + contrast * (src - blur(src)) where |src - blur| > threshold.
 * kernel NULL or "" -> "gaussian"; width <= 3: one 2-D pass with a
 * width x width kernel, else two 1-D passes (:997-1019).  "median" is refused
 * (B2R_ERR_UNSUPPORTED).  dst may be src (in place). */
B2R_API int b2r_unsharp_mask(const b2r_image* dst, const b2r_image* src,
                             const char* kernel, float width, float contrast,
                             float threshold, const b2r_roi* roi,
                             const b2r_options* opt, b2r_report* report,
                             char* err, size_t errlen);

/* ---- fixNonFinite / check_nan ---------------------------------------------
 * Replaces ImageBufAlgo::fixNonFinite (imagebufalgo_pixelmath.cpp:1424-1610)
 * as maketx uses it (--fixnan, maketexture.cpp:1666-1686) and maketx's
 * check_nan_block (--checknan, :340-360, 1690-1705 = mode 0's count).
 * mode: ImageBufAlgo::NonFiniteFixMode -- 0 NONE (count only), 1 BLACK,
 * 2 BOX3 (average of the finite values in the 3x3 neighbourhood), 100 ERROR
 * (count; fails with "Nonfinite pixel values found" if any).  dst gets a copy
 * of src (dst may be src); float and half images are repaired, the other
 * types are only copied.  *pixels_fixed = pixels that held a NaN/Inf.
 * BOX3 reads the neighbourhood from the UNREPAIRED source, so the result does
 * not depend on a scan order; it equals the reference wherever two non-finite
 * values are not neighbours (the reference repairs in place per thread band). */
B2R_API int b2r_fix_nonfinite(const b2r_image* dst, const b2r_image* src,
                              int mode, int* pixels_fixed, const b2r_roi* roi,
                              const b2r_options* opt, b2r_report* report,
                              char* err, size_t errlen);

/* ---- fillholes_pushpull ---------------------------------------------------
 * Replaces ImageBufAlgo::fillholes_pushpull (src/libOpenImageIO/
 * imagebufalgo.cpp:1600-1670), an in-library user of the same resize: a float
 * pyramid built with the triangle-filter resize (each level divided by its
 * alpha where that is nonzero), then pulled back up with level OVER blown-up
 * next level; the finished top level is pasted back at src's origin in dst's
 * type.  Everything stays on the device between the 2*levels resizes.
 * alpha_channel: ImageSpec::alpha_channel ("images must have alpha channels"
 * if < 0).  dst and src need the same channel count; dst may be src. */
B2R_API int b2r_fillholes_pushpull(const b2r_image* dst, const b2r_image* src,
                                   int alpha_channel, const b2r_options* opt,
                                   b2r_report* report, char* err, size_t errlen);

/* ---- computePixelStats / isMonochrome -------------------------------------
 * Replaces ImageBufAlgo::computePixelStats (src/libOpenImageIO/
 * imagebufalgo_compare.cpp:81-215) and, in the same pass over the pixels,
 * ImageBufAlgo::isMonochrome (:539-600): the whole-image reductions
 * make_texture_impl starts with (maketexture.cpp:1396-1409 -- average colour,
 * constant-colour detection as min == max -- and :1469-1476).  One
 * b2r_pixel_stats_t per channel of src (ImageBufAlgo::PixelStats: sum and sum2
 * in double, min/max/avg/stddev in float, counts).  roi NULL = the data
 * window; channels outside [chbegin,chend) keep PixelStats::reset()'s zeros...
 * (min/max are finalised to 0 for channels with no finite value, :103-108).
 * is_monochrome may be NULL; threshold 0 compares exactly (:551-561).
 * Sums are accumulated per thread and merged in a fixed order: deterministic,
 * and equal to the reference's to double rounding (its own merge order depends
 * on thread timing). */
typedef struct b2r_pixel_stats_t {
    float min, max, avg, stddev;
    int64_t nancount, infcount, finitecount;
    double sum, sum2;
} b2r_pixel_stats_t;
B2R_API int b2r_pixel_stats(const b2r_image* src, const b2r_roi* roi,
                            b2r_pixel_stats_t* stats, int nstats,
                            int* is_monochrome, float mono_threshold,
                            const b2r_options* opt, char* err, size_t errlen);

/* ---- row-band sharding (multi-GPU / multi-rank) --------------------------
 * The reference parallelises resize by full-width row bands
 * (parallel_image, src/include/OpenImageIO/imagebufalgo_util.h:37-80).  The
 * same split is used across GPUs: each device runs b2r_resize on its band of
 * the dst ROI; only the source rows that band can touch (band + filter-radius
 * halo, xform.cpp:626-633) are staged on that device, straight from the host
 * source -- no collective, outputs are disjoint.  Host arithmetic only. */
/* Band `band` of `nbands`: rows [ybegin + H*band/nbands, ybegin + H*(band+1)/nbands). */
B2R_API int b2r_band_split(const b2r_roi* roi, int nbands, int band, b2r_roi* out);
/* One large HOST -> HOST resize over the `ndevices` GPUs of a box from this one
 * call (the driver of the split above, for hosts that are not Python): band g
 * of the dst ROI runs as b2r_resize on devices[g] (NULL: 0 .. ndevices-1),
 * issued from one host thread per device; every band stages its own source
 * rows (+ halo) and brings its own rows back over its own PCIe link.  Returns
 * the first band's error, if any.  roi == NULL: dst's data window. */
B2R_API int b2r_resize_banded(const b2r_image* dst, const b2r_image* src,
                              const char* filtername, float filterwidth,
                              const b2r_roi* roi, const int* devices, int ndevices,
                              const b2r_options* opt, char* err, size_t errlen);
/* Source rows [row_begin, row_end), relative to src's data-window origin,
 * that resizing into `roi` of dst reads (what b2r_resize stages for a host
 * source).  Only the images' windows are read, not their pixels. */
B2R_API int b2r_source_rows(const b2r_image* dst, const b2r_image* src,
                            const char* filtername, float filterwidth,
                            const b2r_roi* roi, int* row_begin, int* row_end,
                            char* err, size_t errlen);

/* ---- device-resident MIP chain ---------------------------------------- */
/* Replaces the level loop of write_mipmap,
 * src/libOpenImageIO/maketexture.cpp:823-986: each level is computed on the
 * device from the previous one (no host round trip), float32 pixels
 * (maketx:forcefloat), windows zeroed per level (:848-879).
 * filtername "box" takes the resize_block path (:881-887 -> :141-334), any
 * other name the resize path with setup_filter's width rule (:38-66).
 * Pass level 0 with its REAL windows: like write_mipmap the chain sets its
 * display window to its data window (:876-877) and keeps its origin for the
 * first resize; a level 0 that is offset, cropped or overscanned
 * (orig_was_overscan, :700-703) takes the resize path even for "box" (:881).
 * clamp_half != 0 clamps each level to +-HALF_MAX and, like
 * ImageBufAlgo::clamp(..., clampalpha01=true), the alpha channel
 * (alpha_channel >= 0) to [0,1] (:943-945). */
typedef struct b2r_mipchain b2r_mipchain;
/* level0: float32 image (host or device).  The chain copies/adopts it on
 * `device` and computes every level down to 1x1. */
B2R_API int b2r_mipchain_create(b2r_mipchain** chain, const b2r_image* level0,
                                const char* filtername, int clamp_half,
                                int alpha_channel, const b2r_options* opt,
                                char* err, size_t errlen);
/* The same with maketx's remaining per-level options: --sharpen <contrast>
 * (unsharp_mask with `sharpenfilt`, width 3, threshold 0, before or after
 * each level's resize: maketexture.cpp:909-931; it also turns the "box" fast
 * path off, :880-881).  With opt->highlightcomp the level is additionally
 * clamped to [0, FLT_MAX] (alpha to [0,1]) after rangeexpand (:936-938). */
typedef struct b2r_mip_options {
    const char* filtername;  /* NULL / "" -> "box" */
    int32_t clamp_half;
    int32_t alpha_channel;   /* -1: none */
    float sharpen;           /* <= 0: off */
    int32_t sharpen_first;   /* maketx:sharpenfirst */
    const char* sharpenfilt; /* NULL / "" -> "gaussian" */
    int32_t reserved[4];
} b2r_mip_options;
B2R_API int b2r_mipchain_create_ex(b2r_mipchain** chain, const b2r_image* level0,
                                   const b2r_mip_options* mip,
                                   const b2r_options* opt, char* err,
                                   size_t errlen);
/* The same chain for a HOST level 0 whose levels are wanted in HOST memory, as
 * ONE pipeline -- write_mipmap end to end (maketexture.cpp:823-986) without the
 * file: level 0 goes up in row bands; as soon as the rows a band of level 1
 * reads have landed, the band is computed and comes down on another stream
 * while the next rows are still going up (PCIe is full duplex; level 1 is
 * three quarters of what comes back); levels >= 2 follow from the device.
 * out_levels[k-1] receives level k (float32, the level's size, contiguous in
 * x; entries with NULL pixels are skipped), nout entries.  The chain stays
 * resident (b2r_mipchain_level / _encode_tiles / _destroy as usual).  A level 0
 * the pipeline does not cover (device resident, offset windows, sharpening,
 * highlight compensation, fewer than 256 rows) takes b2r_mipchain_create_ex
 * followed by the downloads: same results. */
B2R_API int b2r_mipchain_create_streamed(b2r_mipchain** chain, const b2r_image* level0,
                                         const b2r_mip_options* mip, const b2r_options* opt,
                                         const b2r_image* out_levels, int nout, char* err,
                                         size_t errlen);
B2R_API int b2r_mipchain_nlevels(const b2r_mipchain* chain);
/* Dimensions and device pointer of a level (level 0 = the input). */
B2R_API int b2r_mipchain_level(const b2r_mipchain* chain, int level,
                               int* width, int* height, void** device_pixels);
/* Copy a finished level into host (or device) memory described by `out`
 * (float32, same nchannels).  Synchronous. */
B2R_API int b2r_mipchain_download(const b2r_mipchain* chain, int level,
                                  const b2r_image* out, char* err,
                                  size_t errlen);
/* Device time of the level loop (CUDA events), ms. */
B2R_API float b2r_mipchain_kernel_ms(const b2r_mipchain* chain);
B2R_API void b2r_mipchain_destroy(b2r_mipchain* chain);

/* ---- pixel hash (maketx oiio:SHA-1) --------------------------------------- */
/* ImageBufAlgo::computePixelHashSHA1 (imagebufalgo_compare.cpp:856-892): SHA-1
 * over the ROI's scanlines in the image's own pixel format, in `blocksize`-row
 * blocks (0 or >= height: one stream) whose hex digests are hashed together,
 * with `extrainfo` appended last.  maketx calls it on the top level with
 * blocksize 256 and "<filter> [sharpen_A=..] [highlightcomp=1 ] .." as extra
 * info (maketexture.cpp:1909-1934).  Runs on `nthreads` HOST threads (0 = all),
 * one block per task; device images are downloaded first.  digest41 receives
 * 40 upper-case hex digits and a terminating NUL. */
B2R_API int b2r_pixel_hash_sha1(const b2r_image* src, const char* extrainfo,
                                const b2r_roi* roi, int blocksize, int nthreads,
                                char* digest41, char* err, size_t errlen);

/* ---- tiled, compressed level output ------------------------ */
/* What write_mipmap does with every finished level -- out->open(AppendMIPLevel),
 * small->write(out), maketexture.cpp:957-974 -- up to the compressed tiles the
 * file plugin stores: each level of `chain` (from first_level on) is re-laid
 * into tile_width x tile_height tiles on the device (float32, or converted to
 * half: maketx -d half), downloaded through pinned staging slots on a copy
 * stream, and deflated tile by tile (zlib; maketx's default "zip" compression,
 * maketexture.cpp:1605) by `nthreads` host threads while the next chunk is in
 * flight.  Edge tiles are zero-padded to full size.  b2r_tileset_write_tiff
 * (below) writes the TIFF container around the tile set; OpenEXR headers and
 * offset tables stay with the reference's plugin. */
typedef struct b2r_tile_options {
    int32_t tile_width, tile_height; /* 0 -> 64 (maketexture.cpp:1100-1106) */
    int32_t out_basetype;            /* B2R_FLOAT (0 -> float), B2R_HALF, or B2R_UINT16 /
                                        B2R_UINT8: the level through convert_type<float,D>
                                        (fmath.h:753-764), what maketx writes for an
                                        integer texture */
    int32_t compression;             /* 0 store, 1 zlib deflate */
    int32_t zlevel;                  /* 1..9, 0 -> 1 */
    int32_t nthreads;                /* 0 -> hardware concurrency */
    int32_t first_level;             /* skip levels below this one */
    int32_t predictor;               /* 0 / 1 none; 3 = TIFF's floating-point predictor
                                        (float / half tiles), 2 = the horizontal one
                                        (uint8 / uint16 tiles) -- what the reference's
                                        TIFF writer sets under zip,
                                        src/tiff.imageio/tiffoutput.cpp:683-698 --
                                        applied per tile row by the re-layout kernel */
    int32_t reserved[2];
} b2r_tile_options;
typedef struct b2r_tileset b2r_tileset;
B2R_API int b2r_mipchain_encode_tiles(const b2r_mipchain* chain,
                                      const b2r_tile_options* opt, b2r_tileset** out,
                                      char* err, size_t errlen);
B2R_API int b2r_tileset_nlevels(const b2r_tileset* ts);
B2R_API int b2r_tileset_level(const b2r_tileset* ts, int level, int* width, int* height,
                              int* ntiles_x, int* ntiles_y);
/* One stored tile (compressed bytes, or raw when compression == 0). */
B2R_API int b2r_tileset_tile(const b2r_tileset* ts, int level, int tx, int ty,
                             const void** data, size_t* bytes);
/* what: 0 wall ms of the encode call, 1 device ms (relayout + downloads),
 * 2 host ms inside the deflate phases, 3 raw bytes, 4 stored bytes,
 * 5 ms spent setting up the staging slots, 6 host ms blocked on downloads. */
B2R_API double b2r_tileset_stat(const b2r_tileset* ts, int what);
B2R_API void b2r_tileset_destroy(b2r_tileset* ts);

/* The device stage of b2r_mipchain_encode_tiles on its own, for hosts that
 * keep the tiles on the device (their own compressor, their own download): one
 * device-resident, contiguous float32 image re-laid into tiles -- raster order,
 * tile_height x tile_width x nchannels each, float32 or half (opt->out_basetype),
 * edge tiles zero padded, opt->predictor as above -- in the caller's device
 * buffer of at least ntiles_x * ntiles_y * tile bytes.  Asynchronous on
 * `stream` (a cudaStream_t, NULL = the default stream). */
B2R_API int b2r_tile_relayout(const b2r_image* src, void* dst_device, size_t dst_bytes,
                              const b2r_tile_options* opt, void* stream, char* err,
                              size_t errlen);

/* The container step the reference leaves to its TIFF plugin
 * (out->open(..., AppendMIPLevel) + small->write(out), maketexture.cpp:957-974;
 * src/tiff.imageio/tiffoutput.cpp): the tiles of `ts` written as a tiled TIFF
 * texture, one directory per MIP level in level order (the layout OpenImageIO's
 * TIFF reader presents as MIP levels when PIXAR_TEXTUREFORMAT is set,
 * tiffinput.cpp:88,1435), contiguous planar config, IEEE float samples (32 or
 * 16 bits) or unsigned integers (16 or 8 bits), Compression = 8 (Adobe deflate: each stored tile is already the
 * zlib stream TIFF wants) or 1, Predictor = 3 / 2 when the tiles were encoded
 * with it, Software / ImageDescription / PIXAR_TEXTUREFORMAT ("Plain Texture") /
 * PIXAR_WRAPMODES.  Classic TIFF below 4 GB, BigTIFF above (or when
 * `bigtiff` != 0).  Tile sizes must be multiples of 16 and every level of the
 * chain must have been encoded (first_level == 0).  Host code only: no
 * device work, no copy of the tile bytes other than the file write. */
typedef struct b2r_tiff_options {
    const char* image_description; /* NULL: none (maketx stores the SHA-1 and stats here) */
    const char* software;          /* NULL: "openimageio_b200" */
    const char* wrapmodes;         /* NULL: "black,black" (maketx's default wrap) */
    int32_t bigtiff;               /* 0 auto, 1 force BigTIFF */
    int32_t reserved[3];
} b2r_tiff_options;
B2R_API int b2r_tileset_write_tiff(const b2r_tileset* ts, const char* path,
                                   const b2r_tiff_options* opt, char* err, size_t errlen);

/* The same for OpenEXR (src/openexr.imageio; maketx's usual container for float
 * / half textures): a single-part tiled file, level mode MIPMAP_LEVELS /
 * ROUND_DOWN (the sizes write_mipmap produces, maketexture.cpp:823, 851-857), ZIP
 * compression, channels Y / Y,A / R,G,B / R,G,B,A of type HALF or FLOAT.  The
 * tile set must hold STORED tiles (compression = 0, predictor = 0): OpenEXR's zip
 * codec wants every tile cropped to the level, planar by channel within a
 * scanline, its bytes interleaved and delta-coded before deflate -- done here on
 * `nthreads` host threads.  attr_names / attr_values: string attributes to add
 * (maketx: "textureformat", "wrapmodes", "oiio:SHA-1", "oiio:AverageColor" ...). */
typedef struct b2r_exr_options {
    const char* const* attr_names;
    const char* const* attr_values;
    int32_t nattrs;
    int32_t zlevel;   /* 1..9, 0 -> 1 */
    int32_t nthreads; /* 0 -> hardware concurrency */
    int32_t reserved[3];
} b2r_exr_options;
B2R_API int b2r_tileset_write_exr(const b2r_tileset* ts, const char* path,
                                  const b2r_exr_options* opt, char* err, size_t errlen);

/* ---- the same chain cut into row bands over several GPUs ---------------- */
/* One process, `ndevices` GPUs of one box (replaces the level loop
 * of write_mipmap, maketexture.cpp:823-986, for textures whose level 0 is too
 * large or too slow to push through one PCIe link).  Level k is split into
 * ndevices row bands; band g lives on devices[g] together with the halo rows
 * level k+1's band g reads.  Level 0 is uploaded band by band straight from
 * the HOST image `level0` (every device pulls its rows over its own PCIe
 * link).  After each level a device pushes the rows its neighbours need into
 * their memory with peer copies over NVLink and the neighbour's next level
 * waits for them on the device (no NCCL, no host synchronisation per level).
 * Levels whose bands would be shorter than 64 rows are gathered on devices[0]
 * and finished there.  Every level equals the single-device chain bit for bit.
 * mip->sharpen, highlight compensation and a level 0 with a non-zero origin
 * or display window != data window are not supported here.
 * devices == NULL: devices 0 .. ndevices-1. */
typedef struct b2r_mipchain_banded b2r_mipchain_banded;
B2R_API int b2r_mipchain_create_banded(b2r_mipchain_banded** chain,
                                       const b2r_image* level0,
                                       const b2r_mip_options* mip,
                                       const int* devices, int ndevices,
                                       char* err, size_t errlen);
B2R_API int b2r_mipchain_banded_nlevels(const b2r_mipchain_banded* chain);
B2R_API int b2r_mipchain_banded_level(const b2r_mipchain_banded* chain, int level,
                                      int* width, int* height);
/* Band `band` of a level: the device that owns it, its rows [row_begin,
 * row_end) and the device address of row_begin (NULL when the device sits
 * the level out). */
B2R_API int b2r_mipchain_banded_band(const b2r_mipchain_banded* chain, int level,
                                     int band, int* device, int* row_begin,
                                     int* row_end, void** device_pixels);
/* Copy a finished level into the host image `out` (float32, same channels);
 * every device sends its own rows.  Synchronous. */
B2R_API int b2r_mipchain_banded_download(const b2r_mipchain_banded* chain, int level,
                                         const b2r_image* out, char* err,
                                         size_t errlen);
/* what: 0 host wall clock of the create call (upload + tables + level loop),
 * 1 level-0 upload and 2 level loop incl. the pushes (device events, max over
 * the devices), 3 bytes pushed between devices; host wall clock of the call's
 * parts: 4 setup (peer access, band buffers), 5 upload threads + tables,
 * 6 level loop issue + final synchronise.  ms (bytes for 3). */
B2R_API float b2r_mipchain_banded_ms(const b2r_mipchain_banded* chain, int what);
B2R_API void b2r_mipchain_banded_destroy(b2r_mipchain_banded* chain);

/* One level step of the chain on arbitrary windows: rows `roi` of level k+1
 * (dst: data window = the rows held, display window = the whole level) from
 * level k (src: data window = the rows available, display window = the whole
 * level).  This is the loop body of write_mipmap (maketexture.cpp:881-922)
 * for a row band; it is what a multi-GPU chain calls per band after the
 * neighbour halo rows have arrived.  "box" levels must be device resident. */
B2R_API int b2r_mip_level(const b2r_image* dst, const b2r_image* src,
                          const char* filtername, const b2r_roi* roi,
                          const b2r_options* opt, char* err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif /* B2R_H */
