/* oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the reference
 * (OpenImageIO 3.2.0.2-dev) filtered-resampling path.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load it.  The product (openimageio_b200/) never links or calls it.
 *
 * Parity status: PINNED.  The restatement is checked against
 *   - the 17 known-answer images testsuite/filters/ref/<filter>.exr,
 *   - testsuite/oiiotool-xform/ref/{resize,resize2,resize64,resize512}.tif,
 *     resized-offset.exr, fit*.tif, resample*.tif,
 *   - testsuite/oiiotool/ref/range{compress,expand}{,-luma}.tif,
 *   - the reference's own src/libutil/filter.cpp object code compiled
 *     unmodified into oracle/_ref/ (bit-exact weight comparison),
 * see tests/test_oracle_golden.py and tests/golden/make_golden.py.
 * MIP levels >= 1 are not pinned by any reference test.
 */
#ifndef OIIO_ORACLE_H
#define OIIO_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* TypeDesc::BASETYPE values (src/include/OpenImageIO/typedesc.h:57-84) */
enum { ORC_UINT8 = 2, ORC_UINT16 = 4, ORC_HALF = 10, ORC_FLOAT = 11 };

typedef struct {
    void* pixels;     /* address of the data-window origin pixel (x,y) */
    int32_t basetype; /* ORC_* */
    int32_t nchannels;
    int32_t x, y, width, height;                       /* data window  */
    int32_t full_x, full_y, full_width, full_height;   /* display window */
    int64_t xstride, ystride;                          /* bytes */
} orc_image;

typedef struct {
    int32_t xbegin, xend, ybegin, yend;
} orc_roi;

/* Filter evaluation (restates src/libutil/filter.cpp). Returns NaN-free float;
 * unknown name -> returns -1 from orc_filter_lookup. */
int orc_filter_count(void);
const char* orc_filter_name(int index);
float orc_filter_default_width(int index);
int orc_filter_lookup(const char* name); /* exact FilterDesc name, else -1 */
int orc_filter_separable(const char* name, float w, float h);
float orc_filter_xfilt(const char* name, float w, float h, float x);
float orc_filter_yfilt(const char* name, float w, float h, float y);
float orc_filter_eval2d(const char* name, float w, float h, float x, float y);
float orc_filter1d_eval(const char* name, float w, float x);

/* ImageBufAlgo::resize restated (imagebufalgo_xform.cpp:595-918).
 * filtername NULL/"" -> default rule; filterwidth 0 -> default rule.
 * nthreads <= 0 -> hardware concurrency.  Returns 0 on success, else nonzero
 * and a message in err. */
int orc_resize(const orc_image* dst, const orc_image* src,
               const char* filtername, float filterwidth, orc_roi roi,
               int nthreads, char* err, int errlen);

/* ImageBufAlgo::resample restated (imagebufalgo_xform.cpp:1107-1186).  */
int orc_resample(const orc_image* dst, const orc_image* src, int interpolate,
                 orc_roi roi, int nthreads, char* err, int errlen);

/* ImageBufAlgo::fit geometry (imagebufalgo_xform.cpp:962-997).  Writes the
 * resize target size and the data-window offset of the result. */
void orc_fit_geometry(int src_full_w, int src_full_h, int fit_w, int fit_h,
                      const char* fillmode, int* resize_w, int* resize_h,
                      int* xoffset, int* yoffset);

/* maketexture.cpp resize_block (box fast path, :141-334): float images only.
 * dst/src are float, same nchannels.  */
int orc_resize_block(const orc_image* dst, const orc_image* src,
                     int envlatlmode, int allow_shift, int nthreads);

/* ImageBufAlgo::rangecompress / rangeexpand restated
 * (imagebufalgo_pixelmath.cpp:560-735).  alpha_channel / z_channel: -1 = none.
 * Pinned against testsuite/oiiotool/ref/range{compress,expand}{,-luma}.tif. */
int orc_rangecompress(const orc_image* dst, const orc_image* src, int useluma,
                      int alpha_channel, int z_channel, orc_roi roi, int chbegin,
                      int chend);
int orc_rangeexpand(const orc_image* dst, const orc_image* src, int useluma,
                    int alpha_channel, int z_channel, orc_roi roi, int chbegin,
                    int chend);

/* ImageBufAlgo::fixNonFinite (imagebufalgo_pixelmath.cpp:1424-1610) and
 * maketx's check_nan_block (maketexture.cpp:340-360).  mode: 0 none (count),
 * 1 black, 2 box3, 100 error (count).  Raster order, single band. */
int orc_fix_nonfinite(const orc_image* dst, const orc_image* src, int mode,
                      orc_roi roi, int chbegin, int chend, int* pixels_fixed);

/* ImageBufAlgo::warp / rotate and fit(exact=1) restated
 * (imagebufalgo_xform.cpp:263-418, 1012-1027, 1727-1760) with Imath's
 * Matrix33<float> algebra.  M: row-major x[i][j] as Imath stores it.
 * wrap: 0 default(black), 1 black, 2 clamp, 3 periodic, 4 mirror.  The filter
 * is passed resolved (name, width, height).  Pinned against
 * testsuite/oiiotool-xform/ref/{warped,rotated*,resizefrom*}.tif, fit4.exr and
 * fit[wh]-<mode>-<size>.exr. */
int orc_warp(const orc_image* dst, const orc_image* src, const float* M,
             const char* filtername, float fw, float fh, int wrap, int edgeclamp,
             orc_roi roi, int nthreads, char* err, int errlen);
/* ImageBufAlgo::st_warp restated (imagebufalgo_xform.cpp:1834-1925): every dst
 * pixel looks its source position up in the ST image.  Pinned against
 * testsuite/oiiotool-xform/ref/st_warped.tif. */
int orc_st_warp(const orc_image* dst, const orc_image* src, const orc_image* st,
                int chan_s, int chan_t, int flip_s, int flip_t,
                const char* filtername, float fw, float fh, orc_roi roi,
                int chbegin, int chend, int nthreads, char* err, int errlen);
void orc_m33_inverse(const float* M, float* out);
void orc_rotate_matrix(float angle, float cx, float cy, float* out);
void orc_fromto_matrix(float from_x, float from_y, float from_w, float from_h,
                       float to_x, float to_y, float to_w, float to_h, float* out);
void orc_transform_roi(const float* M, orc_roi roi, orc_roi* out);
void orc_fit_exact_matrix(int src_full_w, int src_full_h, int fit_w, int fit_h,
                          const char* fillmode, float* M, int* resize_w,
                          int* resize_h);

/* ImageBufAlgo::make_kernel / convolve / unsharp_mask restated
 * (imagebufalgo.cpp:772-1031; "median" excluded).  Pinned against
 * testsuite/oiiotool/ref/{bspline-blur,gauss5x5-blur,unsharp}.tif.
 * orc_make_kernel: out == NULL queries the size; returns 1 for an unknown
 * kernel name (the reference returns a box plus an error). */
int orc_make_kernel(const char* name, float width, float height, int normalize,
                    float* out, int* kw, int* kh);
int orc_convolve(const orc_image* dst, const orc_image* src, const float* kernel,
                 int kw, int kh, int normalize, orc_roi roi, int chbegin, int chend,
                 int nthreads);
int orc_unsharp_mask(const orc_image* dst, const orc_image* src, const char* kernel,
                     float width, float contrast, float threshold, orc_roi roi,
                     int nthreads, char* err, int errlen);

/* float -> dst-type store and src-type -> float load used by the above
 * (fmath.h:753-764, 1009-1031); exposed for unit tests. */
float orc_load_as_float(const void* p, int basetype);
void orc_store_from_float(void* p, int basetype, float v);

#ifdef __cplusplus
}
#endif
#endif
