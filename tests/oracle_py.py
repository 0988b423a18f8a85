"""ctypes access to the CPU oracle (oracle/liboiio_oracle.so) and, when built,
to the reference's own filter object code (oracle/_ref/libref_filter.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg; never by openimageio_b200/.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

UINT8, UINT16, HALF, FLOAT = 2, 4, 10, 11
_NP2BT = {np.dtype(np.uint8): UINT8, np.dtype(np.uint16): UINT16,
          np.dtype(np.float16): HALF, np.dtype(np.float32): FLOAT}


class OrcImage(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("basetype", C.c_int32),
                ("nchannels", C.c_int32),
                ("x", C.c_int32), ("y", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32),
                ("full_x", C.c_int32), ("full_y", C.c_int32),
                ("full_width", C.c_int32), ("full_height", C.c_int32),
                ("xstride", C.c_int64), ("ystride", C.c_int64)]


class OrcRoi(C.Structure):
    _fields_ = [("xbegin", C.c_int32), ("xend", C.c_int32),
                ("ybegin", C.c_int32), ("yend", C.c_int32)]


_lib = None
_ref = None


def build_oracle():
    subprocess.check_call(["make", "-s", "-C", ORACLE_DIR])


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(ORACLE_DIR, "liboiio_oracle.so")
        if not os.path.exists(path):
            build_oracle()
        L = C.CDLL(path)
        L.orc_filter_name.restype = C.c_char_p
        L.orc_filter_default_width.restype = C.c_float
        for fn in (L.orc_filter_xfilt, L.orc_filter_yfilt):
            fn.restype = C.c_float
            fn.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_float]
        L.orc_filter_eval2d.restype = C.c_float
        L.orc_filter_eval2d.argtypes = [C.c_char_p, C.c_float, C.c_float,
                                        C.c_float, C.c_float]
        L.orc_filter1d_eval.restype = C.c_float
        L.orc_filter1d_eval.argtypes = [C.c_char_p, C.c_float, C.c_float]
        L.orc_filter_separable.argtypes = [C.c_char_p, C.c_float, C.c_float]
        L.orc_filter_lookup.argtypes = [C.c_char_p]
        L.orc_resize.argtypes = [C.POINTER(OrcImage), C.POINTER(OrcImage),
                                 C.c_char_p, C.c_float, OrcRoi, C.c_int,
                                 C.c_char_p, C.c_int]
        L.orc_resample.argtypes = [C.POINTER(OrcImage), C.POINTER(OrcImage),
                                   C.c_int, OrcRoi, C.c_int, C.c_char_p,
                                   C.c_int]
        L.orc_resize_block.argtypes = [C.POINTER(OrcImage),
                                       C.POINTER(OrcImage), C.c_int, C.c_int,
                                       C.c_int]
        L.orc_fit_geometry.argtypes = [C.c_int] * 4 + [C.c_char_p] + \
            [C.POINTER(C.c_int)] * 4
        L.orc_load_as_float.restype = C.c_float
        for fn in (L.orc_rangecompress, L.orc_rangeexpand):
            fn.argtypes = [C.POINTER(OrcImage), C.POINTER(OrcImage), C.c_int,
                           C.c_int, C.c_int, OrcRoi, C.c_int, C.c_int]
            fn.restype = C.c_int
        F9 = C.POINTER(C.c_float)
        L.orc_warp.argtypes = [C.POINTER(OrcImage), C.POINTER(OrcImage), F9,
                               C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_int,
                               OrcRoi, C.c_int, C.c_char_p, C.c_int]
        L.orc_st_warp.argtypes = [C.POINTER(OrcImage)] * 3 + [C.c_int] * 4 + \
            [C.c_char_p, C.c_float, C.c_float, OrcRoi, C.c_int, C.c_int, C.c_int,
             C.c_char_p, C.c_int]
        L.orc_m33_inverse.argtypes = [F9, F9]
        L.orc_rotate_matrix.argtypes = [C.c_float] * 3 + [F9]
        L.orc_fromto_matrix.argtypes = [C.c_float] * 8 + [F9]
        L.orc_transform_roi.argtypes = [F9, OrcRoi, C.POINTER(OrcRoi)]
        L.orc_fit_exact_matrix.argtypes = [C.c_int] * 4 + [C.c_char_p, F9,
                                                          C.POINTER(C.c_int),
                                                          C.POINTER(C.c_int)]
        for fn in (L.orc_m33_inverse, L.orc_rotate_matrix, L.orc_fromto_matrix,
                   L.orc_transform_roi, L.orc_fit_exact_matrix):
            fn.restype = None
        L.orc_make_kernel.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_int,
                                      C.POINTER(C.c_float), C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]
        L.orc_convolve.argtypes = [C.POINTER(OrcImage), C.POINTER(OrcImage),
                                   C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                                   OrcRoi, C.c_int, C.c_int, C.c_int]
        L.orc_unsharp_mask.argtypes = [C.POINTER(OrcImage), C.POINTER(OrcImage),
                                       C.c_char_p, C.c_float, C.c_float, C.c_float,
                                       OrcRoi, C.c_int, C.c_char_p, C.c_int]
        L.orc_fix_nonfinite.argtypes = [C.POINTER(OrcImage), C.POINTER(OrcImage),
                                        C.c_int, OrcRoi, C.c_int, C.c_int,
                                        C.POINTER(C.c_int)]
        L.orc_load_as_float.argtypes = [C.c_void_p, C.c_int]
        L.orc_store_from_float.argtypes = [C.c_void_p, C.c_int, C.c_float]
        _lib = L
    return _lib


def ref_filter_lib():
    """The reference's own filter.cpp object code, or None if not built."""
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libref_filter.so")
        if not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.ref_filter2d_name.restype = C.c_char_p
        R.ref_filter1d_name.restype = C.c_char_p
        R.ref_filter2d_default_width.restype = C.c_float
        R.ref_filter2d_eval.argtypes = [
            C.c_char_p, C.c_float, C.c_float, C.c_int, C.c_int,
            C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int),
            C.POINTER(C.c_float), C.POINTER(C.c_float)]
        R.ref_filter1d_eval.argtypes = [C.c_char_p, C.c_float, C.c_int,
                                        C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_float)]
        _ref = R
    return _ref


class Img:
    """numpy pixels (H, W, C) + OpenImageIO-style data / full windows."""

    def __init__(self, arr, x=0, y=0, full=None):
        if arr.ndim == 2:
            arr = arr[:, :, None]
        assert arr.ndim == 3
        self.arr = arr
        self.x, self.y = x, y
        h, w, _ = arr.shape
        self.full = full if full is not None else (x, y, w, h)

    @property
    def nchannels(self):
        return self.arr.shape[2]

    def c(self):
        a = self.arr
        im = OrcImage()
        im.pixels = a.ctypes.data
        im.basetype = _NP2BT[a.dtype]
        im.nchannels = a.shape[2]
        im.x, im.y = self.x, self.y
        im.width, im.height = a.shape[1], a.shape[0]
        im.full_x, im.full_y, im.full_width, im.full_height = self.full
        im.xstride = a.strides[1]
        im.ystride = a.strides[0]
        return im


def resize(dst, src, filtername="", filterwidth=0.0, roi=None, nthreads=0):
    """dst, src: Img.  Writes into dst.arr.  Raises on error."""
    L = lib()
    d, s = dst.c(), src.c()
    if roi is None:
        roi = (dst.x, dst.x + dst.arr.shape[1], dst.y, dst.y + dst.arr.shape[0])
    err = C.create_string_buffer(512)
    rc = L.orc_resize(C.byref(d), C.byref(s),
                      (filtername or "").encode(), float(filterwidth),
                      OrcRoi(*roi), int(nthreads), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return dst


def resample(dst, src, interpolate=True, roi=None, nthreads=0):
    L = lib()
    d, s = dst.c(), src.c()
    if roi is None:
        roi = (dst.x, dst.x + dst.arr.shape[1], dst.y, dst.y + dst.arr.shape[0])
    err = C.create_string_buffer(512)
    rc = L.orc_resample(C.byref(d), C.byref(s), 1 if interpolate else 0,
                        OrcRoi(*roi), int(nthreads), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return dst


def _range(fn, dst, src, useluma, alpha_channel, z_channel, roi, chrange):
    d, s = dst.c(), src.c()
    if roi is None:
        roi = (dst.x, dst.x + dst.arr.shape[1], dst.y, dst.y + dst.arr.shape[0])
    ch0, ch1 = chrange if chrange else (0, dst.arr.shape[2])
    rc = fn(C.byref(d), C.byref(s), int(bool(useluma)), int(alpha_channel),
            int(z_channel), OrcRoi(*roi), int(ch0), int(ch1))
    if rc:
        raise RuntimeError("oracle range op failed")
    return dst


def rangecompress(dst, src, useluma=False, alpha_channel=-1, z_channel=-1,
                  roi=None, chrange=None):
    """dst, src: Img (may wrap the same array: in place)."""
    return _range(lib().orc_rangecompress, dst, src, useluma, alpha_channel,
                  z_channel, roi, chrange)


def rangeexpand(dst, src, useluma=False, alpha_channel=-1, z_channel=-1,
                roi=None, chrange=None):
    return _range(lib().orc_rangeexpand, dst, src, useluma, alpha_channel,
                  z_channel, roi, chrange)


WRAP = {"default": 0, "black": 1, "clamp": 2, "periodic": 3, "mirror": 4}


def _m9(M):
    a = (C.c_float * 9)(*[float(v) for v in np.asarray(M, np.float32).reshape(9)])
    return a


def warp(dst, src, M, filtername="lanczos3", fw=0.0, fh=0.0, wrap="default",
         edgeclamp=False, roi=None, nthreads=0):
    """warp_<D,S> over roi of dst.  The filter is given resolved: width 0 =
    the filter's default width."""
    d, s = dst.c(), src.c()
    if roi is None:
        roi = (dst.x, dst.x + dst.arr.shape[1], dst.y, dst.y + dst.arr.shape[0])
    err = C.create_string_buffer(512)
    rc = lib().orc_warp(C.byref(d), C.byref(s), _m9(M), (filtername or "").encode(),
                        float(fw), float(fh), WRAP[wrap], int(bool(edgeclamp)),
                        OrcRoi(*roi), int(nthreads), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return dst


def st_warp(dst, src, st, filtername="lanczos3", fw=0.0, fh=0.0, chan_s=0,
            chan_t=1, flip_s=False, flip_t=False, roi=None, chrange=None,
            nthreads=0):
    d, s, t = dst.c(), src.c(), st.c()
    cb, ce = chrange if chrange else (0, min(dst.nchannels, src.nchannels))
    err = C.create_string_buffer(512)
    rc = lib().orc_st_warp(C.byref(d), C.byref(s), C.byref(t), chan_s, chan_t,
                           int(flip_s), int(flip_t), (filtername or "").encode(),
                           float(fw), float(fh), _full_roi(dst, roi), cb, ce,
                           int(nthreads), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return dst


def m33_inverse(M):
    out = (C.c_float * 9)()
    lib().orc_m33_inverse(_m9(M), out)
    return np.array(list(out), np.float32).reshape(3, 3)


def rotate_matrix(angle, cx, cy):
    out = (C.c_float * 9)()
    lib().orc_rotate_matrix(float(angle), float(cx), float(cy), out)
    return np.array(list(out), np.float32).reshape(3, 3)


def fromto_matrix(frm, to):
    """frm, to: (x, y, w, h) as oiiotool --resize:from=:to= (oiiotool.cpp:4704-4707)."""
    out = (C.c_float * 9)()
    lib().orc_fromto_matrix(float(frm[0]), float(frm[1]), float(frm[2]), float(frm[3]),
                            float(to[0]), float(to[1]), float(to[2]), float(to[3]), out)
    return np.array(list(out), np.float32).reshape(3, 3)


def transform_roi(M, roi):
    out = OrcRoi()
    lib().orc_transform_roi(_m9(M), OrcRoi(*roi), C.byref(out))
    return (out.xbegin, out.xend, out.ybegin, out.yend)


def fit_exact_matrix(src_full_w, src_full_h, fit_w, fit_h, fillmode="letterbox"):
    out = (C.c_float * 9)()
    rw, rh = C.c_int(), C.c_int()
    lib().orc_fit_exact_matrix(src_full_w, src_full_h, fit_w, fit_h, fillmode.encode(),
                               out, C.byref(rw), C.byref(rh))
    return np.array(list(out), np.float32).reshape(3, 3), rw.value, rh.value


def make_kernel(name, width, height, normalize=True):
    """(kernel array (kh, kw) float32, known) -- ImageBufAlgo::make_kernel."""
    kw, kh = C.c_int(), C.c_int()
    lib().orc_make_kernel(name.encode(), float(width), float(height),
                          int(normalize), None, C.byref(kw), C.byref(kh))
    out = np.zeros((kh.value, kw.value), np.float32)
    rc = lib().orc_make_kernel(name.encode(), float(width), float(height),
                               int(normalize),
                               out.ctypes.data_as(C.POINTER(C.c_float)),
                               C.byref(kw), C.byref(kh))
    return out, rc == 0


def _full_roi(dst, roi):
    if roi is None:
        roi = (dst.x, dst.x + dst.arr.shape[1], dst.y, dst.y + dst.arr.shape[0])
    return OrcRoi(*roi)


def convolve(dst, src, kernel, normalize=True, roi=None, chrange=None, nthreads=0):
    k = np.ascontiguousarray(kernel, np.float32)
    d, s = dst.c(), src.c()
    cb, ce = chrange if chrange else (0, dst.nchannels)
    rc = lib().orc_convolve(C.byref(d), C.byref(s),
                            k.ctypes.data_as(C.POINTER(C.c_float)), k.shape[1],
                            k.shape[0], int(normalize), _full_roi(dst, roi), cb, ce,
                            int(nthreads))
    if rc:
        raise RuntimeError("orc_convolve failed")
    return dst


def unsharp_mask(dst, src, kernel="gaussian", width=3.0, contrast=1.0,
                 threshold=0.0, roi=None, nthreads=0):
    d, s = dst.c(), src.c()
    err = C.create_string_buffer(512)
    rc = lib().orc_unsharp_mask(C.byref(d), C.byref(s), kernel.encode(), float(width),
                                float(contrast), float(threshold),
                                _full_roi(dst, roi), int(nthreads), err, 512)
    if rc:
        raise RuntimeError(err.value.decode())
    return dst


NONFINITE = {"none": 0, "black": 1, "box3": 2, "error": 100}


def fix_nonfinite(dst, src, mode="box3", roi=None, chrange=None):
    """Returns the number of pixels with a non-finite value (= repaired)."""
    d, s = dst.c(), src.c()
    cb, ce = chrange if chrange else (0, dst.nchannels)
    n = C.c_int(0)
    rc = lib().orc_fix_nonfinite(C.byref(d), C.byref(s), NONFINITE[mode],
                                 _full_roi(dst, roi), cb, ce, C.byref(n))
    if rc:
        raise RuntimeError("orc_fix_nonfinite: bad arguments")
    return n.value


def resize_block(dst, src, allow_shift=False):
    L = lib()
    d, s = dst.c(), src.c()
    rc = L.orc_resize_block(C.byref(d), C.byref(s), 0, int(allow_shift), 0)
    if rc:
        raise RuntimeError("orc_resize_block failed")
    return dst


def fit_geometry(sw, sh, fw, fh, fillmode="letterbox"):
    L = lib()
    v = [C.c_int() for _ in range(4)]
    L.orc_fit_geometry(sw, sh, fw, fh, fillmode.encode(), *[C.byref(i) for i in v])
    return tuple(i.value for i in v)


def mip_chain(level0, filtername="box", highlightcomp=False, alpha_channel=-1,
              sharpen=0.0, sharpen_first=False, sharpenfilt="gaussian",
              clamp_half=False):
    """write_mipmap level loop (maketexture.cpp:823-986) on a float32 HxWxC
    array, no file output.  Returns [level1, level2, ...].  highlightcomp:
    maketx --hicomp, rangecompress(img) / rangeexpand(small) + clamp around
    every filtered level (:907-938); sharpen: maketx --sharpen (:909-931)."""
    assert level0.dtype == np.float32
    out = []
    fmax = np.float32(3.402823466e+38)
    img = level0.copy() if highlightcomp else level0
    while img.shape[1] > 1 or img.shape[0] > 1:
        h, w, c = img.shape
        sw = w // 2 if w > 1 else w
        sh = h // 2 if h > 1 else h
        small = np.empty((sh, sw, c), np.float32)
        if filtername == "box" and not sharpen > 0:
            resize_block(Img(small), Img(img))
        else:
            if highlightcomp:
                rangecompress(Img(img), Img(img), alpha_channel=alpha_channel)
            if sharpen > 0 and sharpen_first:
                sharp = np.empty_like(img)
                unsharp_mask(Img(sharp), Img(img), sharpenfilt, 3.0, sharpen, 0.0)
                img = sharp
            resize(Img(small), Img(img), filtername)
            if sharpen > 0 and not sharpen_first:
                sharp = np.empty_like(small)
                unsharp_mask(Img(sharp), Img(small), sharpenfilt, 3.0, sharpen, 0.0)
                small = sharp
            if highlightcomp:
                rangeexpand(Img(small), Img(small), alpha_channel=alpha_channel)
                # ImageBufAlgo::clamp(small, small, 0, FLT_MAX, clampalpha01=true)
                np.clip(small, np.float32(0), fmax, out=small)
                if alpha_channel >= 0:
                    np.clip(small[..., alpha_channel], 0, 1,
                            out=small[..., alpha_channel])
        if clamp_half:
            # clamp(small, small, -HALF_MAX, HALF_MAX, clampalpha01) (:943-945):
            # in place, so the next level is made from the clamped one
            np.clip(small, np.float32(-65504.0), np.float32(65504.0), out=small)
            if alpha_channel >= 0:
                np.clip(small[..., alpha_channel], 0, 1, out=small[..., alpha_channel])
        out.append(small)
        # the reference compresses img in place AFTER the level was written
        # to the file: keep the returned level intact
        img = small.copy() if highlightcomp else small
    return out


def fillholes_pushpull(src_arr, alpha_channel):
    """ImageBufAlgo::fillholes_pushpull (imagebufalgo.cpp:1600-1670) on an
    HxWxC array of any of the four types: float pyramid by triangle resize +
    divide_by_alpha (:1578-1596), pulled back up with `over`
    (imagebufalgo_pixelmath.cpp:1687-1692), converted back to the input type.
    Composition of the oracle's resize with float32 numpy point ops."""
    f32 = np.float32
    h, w, c = src_arr.shape
    top = np.zeros((h, w, c), f32)
    resample(Img(top), Img(src_arr), False)          # paste: convert to float
    pyramid = [top]
    while w > 1 or h > 1:
        w, h = max(1, w // 2), max(1, h // 2)
        small = np.zeros((h, w, c), f32)
        resize(Img(small), Img(pyramid[-1]), "triangle")
        a = small[..., alpha_channel]
        nz = a != 0
        small[nz] = small[nz] / a[nz][:, None]
        pyramid.append(small)
    for i in range(len(pyramid) - 2, -1, -1):
        big, small = pyramid[i], pyramid[i + 1]
        blowup = np.zeros_like(big)
        resize(Img(blowup), Img(small), "triangle")
        alpha = np.clip(big[..., alpha_channel], f32(0), f32(1))
        one_minus = (f32(1) - alpha)[..., None]
        big[...] = big + one_minus * blowup
    return quantize(pyramid[0], src_arr.dtype) if src_arr.dtype != np.float32 \
        else pyramid[0]


def pixel_stats(arr):
    """ImageBufAlgo::computePixelStats (imagebufalgo_compare.cpp:81-215) on an
    HxWxC array: per channel dict of min, max, avg, stddev, nancount,
    infcount, finitecount, sum, sum2.  Values are converted to float the way
    the iterator does; sums run in double (numpy pairwise summation -- the
    reference's own order depends on its thread bands)."""
    h, w, c = arr.shape
    f = np.zeros((h, w, c), np.float32)
    resample(Img(f), Img(arr), False)
    out = []
    for ch in range(c):
        v = f[..., ch].ravel()
        nan = np.isnan(v)
        inf = np.isinf(v)
        fin = v[~(nan | inf)].astype(np.float64)
        d = dict(nancount=int(nan.sum()), infcount=int(inf.sum()),
                 finitecount=int(fin.size), sum=float(fin.sum()),
                 sum2=float((fin * fin).sum()))
        if fin.size == 0:
            d.update(min=0.0, max=0.0, avg=0.0, stddev=0.0)
        else:
            davg = d["sum"] / fin.size
            var = d["sum2"] / fin.size - davg * davg
            d.update(min=float(np.float32(fin.min())), max=float(np.float32(fin.max())),
                     avg=float(np.float32(davg)),
                     stddev=float(np.float32(np.sqrt(var) if var > 0 else 0.0)))
        out.append(d)
    return out


def is_monochrome(arr, threshold=0.0):
    """ImageBufAlgo::isMonochrome (imagebufalgo_compare.cpp:539-600)."""
    if arr.shape[2] < 2:
        return True
    if threshold == 0.0:
        return bool((arr[..., 1:] == arr[..., :1]).all())
    f = np.zeros(arr.shape, np.float32)
    resample(Img(f), Img(arr), False)
    return bool(not (np.abs(f[..., 1:] - f[..., :1]) > np.float32(threshold)).any())


def quantize(arr_f32, dtype):
    """convert_type<float, D> on a whole array through the oracle's store."""
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return arr_f32.copy()
    if dtype == np.float16:
        return arr_f32.astype(np.float16)
    mx = 255.0 if dtype == np.uint8 else 65535.0
    s = arr_f32.astype(np.float32) * np.float32(mx)
    s = s + np.where(s < 0, np.float32(-0.5), np.float32(0.5)).astype(np.float32)
    s = np.where(np.isnan(s), np.float32(0), s)
    s = np.clip(s, np.float32(0), np.float32(mx))
    return s.astype(dtype)
