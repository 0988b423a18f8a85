"""openimageio_b200 -- B200-native filtered resampling behind OpenImageIO's API.

Host-side mirror of the reference's Python surface for ONE path
(`OpenImageIO.ImageBufAlgo.resize / fit / resample`, `Filter2D` by name, the
maketx MIP loop), forwarding to the C ABI in `libb2r.so` (include/b2r.h), which
launches hand-written sm_100a CUDA kernels.  There is no CPU fallback: without
the built extension or without a GPU every compute call raises / returns an
error.

Reference interfaces mirrored (paths in the OpenImageIO tree):
  src/python/py_imagebufalgo.cpp:1470-1560   IBA_resize / IBA_fit / IBA_resample
  src/include/OpenImageIO/imagebufalgo.h:667-762
  src/libOpenImageIO/imagebufalgo.cpp:64-281  IBAprep (ROI / dst allocation)
  src/include/OpenImageIO/imageio.h:93-230    ROI
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

__all__ = ["ROI", "ImageSpec", "ImageBuf", "ImageBufAlgo", "Filter2D",
           "MipChain", "lib", "get_string_attribute", "B2RError", "band_split",
           "source_rows", "resize_banded", "shard_frames"]

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2R_LIBPATH: load another build of the same library (A/B kernel experiments)
_LIBPATH = os.environ.get("B2R_LIBPATH") or os.path.join(_HERE, "libb2r.so")

UINT8, UINT16, HALF, FLOAT = 2, 4, 10, 11
_NAME2BT = {"uint8": UINT8, "uint16": UINT16, "half": HALF, "float": FLOAT}
_BT2NP = {UINT8: np.uint8, UINT16: np.uint16, HALF: np.float16, FLOAT: np.float32}
_NP2BT = {np.dtype(v): k for k, v in _BT2NP.items()}

PATH_AUTO, PATH_EXACT, PATH_FUSED = 0, 1, 2
MEM_AUTO, MEM_HOST, MEM_DEVICE = 0, 1, 2


class B2RError(RuntimeError):
    pass


class _Image(C.Structure):
    _fields_ = [("pixels", C.c_void_p), ("basetype", C.c_int32),
                ("nchannels", C.c_int32),
                ("x", C.c_int32), ("y", C.c_int32),
                ("width", C.c_int32), ("height", C.c_int32),
                ("full_x", C.c_int32), ("full_y", C.c_int32),
                ("full_width", C.c_int32), ("full_height", C.c_int32),
                ("xstride", C.c_int64), ("ystride", C.c_int64),
                ("memspace", C.c_int32), ("device", C.c_int32)]


class _Roi(C.Structure):
    _fields_ = [("xbegin", C.c_int32), ("xend", C.c_int32),
                ("ybegin", C.c_int32), ("yend", C.c_int32),
                ("chbegin", C.c_int32), ("chend", C.c_int32)]


class _Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("path", C.c_int32),
                ("cuda_stream", C.c_void_p), ("reserved", C.c_int32 * 4),
                ("highlightcomp", C.c_int32), ("alpha_channel1", C.c_int32),
                ("z_channel1", C.c_int32), ("contract_fma", C.c_int32)]


class _MipOptions(C.Structure):
    _fields_ = [("filtername", C.c_char_p), ("clamp_half", C.c_int32),
                ("alpha_channel", C.c_int32), ("sharpen", C.c_float),
                ("sharpen_first", C.c_int32), ("sharpenfilt", C.c_char_p),
                ("reserved", C.c_int32 * 4)]


class _TileOptions(C.Structure):
    _fields_ = [("tile_width", C.c_int32), ("tile_height", C.c_int32),
                ("out_basetype", C.c_int32), ("compression", C.c_int32),
                ("zlevel", C.c_int32), ("nthreads", C.c_int32),
                ("first_level", C.c_int32), ("predictor", C.c_int32),
                ("reserved", C.c_int32 * 2)]


_TILE_TYPES = {"float": (11, np.float32), "half": (10, np.float16),
               "uint16": (4, np.uint16), "uint8": (2, np.uint8)}


class _ExrOptions(C.Structure):
    _fields_ = [("attr_names", C.POINTER(C.c_char_p)), ("attr_values", C.POINTER(C.c_char_p)),
                ("nattrs", C.c_int32), ("zlevel", C.c_int32), ("nthreads", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class _TiffOptions(C.Structure):
    _fields_ = [("image_description", C.c_char_p), ("software", C.c_char_p),
                ("wrapmodes", C.c_char_p), ("bigtiff", C.c_int32),
                ("reserved", C.c_int32 * 3)]


class _PixelStatsC(C.Structure):
    _fields_ = [("min", C.c_float), ("max", C.c_float), ("avg", C.c_float),
                ("stddev", C.c_float), ("nancount", C.c_int64),
                ("infcount", C.c_int64), ("finitecount", C.c_int64),
                ("sum", C.c_double), ("sum2", C.c_double)]


class PixelStats:
    """ImageBufAlgo::PixelStats (imagebufalgo.h): per-channel lists."""

    def __init__(self, recs=(), monochrome=None):
        for f, _ in _PixelStatsC._fields_:
            setattr(self, f, [getattr(r, f) for r in recs])
        self.monochrome = monochrome


class Report(C.Structure):
    _fields_ = [("path_used", C.c_int32), ("kernels_launched", C.c_int32),
                ("xtaps", C.c_int32), ("ytaps", C.c_int32),
                ("fused_vmode", C.c_int32), ("fused_htaps", C.c_int32),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("kernel_ms", C.c_float), ("nonfinite_redo", C.c_int32)]


_lib = None


def lib():
    """Load libb2r.so (built in-tree by __graft_entry__.build / csrc/Makefile).
    Fails loudly when the extension is missing: there is no other path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIBPATH):
        raise B2RError(
            "openimageio_b200: %s not built (run `python -c 'import "
            "__graft_entry__ as g; g.build()'` or `make -C "
            "openimageio_b200/csrc`); there is no CPU fallback" % _LIBPATH)
    L = C.CDLL(_LIBPATH)
    L.b2r_version.restype = C.c_char_p
    L.b2r_filter_desc.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_char_p),
                                  C.POINTER(C.c_float), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int)]
    L.b2r_filter2d_create.restype = C.c_void_p
    L.b2r_filter2d_create.argtypes = [C.c_char_p, C.c_float, C.c_float]
    L.b2r_filter2d_destroy.argtypes = [C.c_void_p]
    for n in ("width", "height"):
        f = getattr(L, "b2r_filter2d_" + n)
        f.restype = C.c_float
        f.argtypes = [C.c_void_p]
    L.b2r_filter2d_separable.argtypes = [C.c_void_p]
    L.b2r_filter2d_name.restype = C.c_char_p
    L.b2r_filter2d_name.argtypes = [C.c_void_p]
    L.b2r_filter2d_eval.restype = C.c_float
    L.b2r_filter2d_eval.argtypes = [C.c_void_p, C.c_float, C.c_float]
    for n in ("xfilt", "yfilt"):
        f = getattr(L, "b2r_filter2d_" + n)
        f.restype = C.c_float
        f.argtypes = [C.c_void_p, C.c_float]
    L.b2r_filter1d_create.restype = C.c_void_p
    L.b2r_filter1d_create.argtypes = [C.c_char_p, C.c_float]
    L.b2r_filter1d_destroy.argtypes = [C.c_void_p]
    L.b2r_filter1d_width.restype = C.c_float
    L.b2r_filter1d_width.argtypes = [C.c_void_p]
    L.b2r_filter1d_name.restype = C.c_char_p
    L.b2r_filter1d_name.argtypes = [C.c_void_p]
    L.b2r_filter1d_eval.restype = C.c_float
    L.b2r_filter1d_eval.argtypes = [C.c_void_p, C.c_float]
    L.b2r_resize.argtypes = [C.POINTER(_Image), C.POINTER(_Image), C.c_char_p,
                             C.c_float, C.POINTER(_Roi), C.POINTER(_Options),
                             C.POINTER(Report), C.c_char_p, C.c_size_t]
    L.b2r_resize_batch.argtypes = [C.POINTER(_Image), C.POINTER(_Image), C.c_int,
                                   C.c_char_p, C.c_float, C.POINTER(_Roi),
                                   C.POINTER(_Options), C.POINTER(Report), C.c_char_p,
                                   C.c_size_t]
    if hasattr(L, "b2r_resize_banded"):  # (absent from libraries older than round 2)
        L.b2r_resize_banded.argtypes = [C.POINTER(_Image), C.POINTER(_Image), C.c_char_p,
                                        C.c_float, C.POINTER(_Roi), C.POINTER(C.c_int), C.c_int,
                                        C.POINTER(_Options), C.c_char_p, C.c_size_t]
        L.b2r_mipchain_create_streamed.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Image),
                                                   C.POINTER(_MipOptions), C.POINTER(_Options),
                                                   C.POINTER(_Image), C.c_int, C.c_char_p,
                                                   C.c_size_t]
    L.b2r_resize_filter.argtypes = [C.POINTER(_Image), C.POINTER(_Image),
                                    C.c_void_p, C.POINTER(_Roi),
                                    C.POINTER(_Options), C.POINTER(Report),
                                    C.c_char_p, C.c_size_t]
    L.b2r_resample.argtypes = [C.POINTER(_Image), C.POINTER(_Image), C.c_int,
                               C.POINTER(_Roi), C.POINTER(_Options),
                               C.POINTER(Report), C.c_char_p, C.c_size_t]
    for f in (L.b2r_rangecompress, L.b2r_rangeexpand):
        f.argtypes = [C.POINTER(_Image), C.POINTER(_Image), C.c_int, C.c_int,
                      C.c_int, C.POINTER(_Roi), C.POINTER(_Options),
                      C.POINTER(Report), C.c_char_p, C.c_size_t]
    F9 = C.POINTER(C.c_float)
    L.b2r_warp.argtypes = [C.POINTER(_Image), C.POINTER(_Image), F9, C.c_char_p,
                           C.c_float, C.c_int, C.c_int, C.POINTER(_Roi),
                           C.POINTER(_Options), C.POINTER(Report), C.c_char_p,
                           C.c_size_t]
    L.b2r_warp_filter.argtypes = [C.POINTER(_Image), C.POINTER(_Image), F9,
                                  C.c_void_p, C.c_int, C.c_int, C.POINTER(_Roi),
                                  C.POINTER(_Options), C.POINTER(Report),
                                  C.c_char_p, C.c_size_t]
    L.b2r_st_warp.argtypes = [C.POINTER(_Image)] * 3 + [C.c_char_p, C.c_float] + \
        [C.c_int] * 4 + [C.POINTER(_Roi), C.POINTER(_Options), C.POINTER(Report),
                         C.c_char_p, C.c_size_t]
    L.b2r_st_warp_filter.argtypes = [C.POINTER(_Image)] * 3 + [C.c_void_p] + \
        [C.c_int] * 4 + [C.POINTER(_Roi), C.POINTER(_Options), C.POINTER(Report),
                         C.c_char_p, C.c_size_t]
    L.b2r_m33_inverse.argtypes = [F9, F9]
    L.b2r_rotate_matrix.argtypes = [C.c_float] * 3 + [F9]
    L.b2r_fromto_matrix.argtypes = [C.c_float] * 8 + [F9]
    L.b2r_transform_roi.argtypes = [F9, C.POINTER(_Roi), C.POINTER(_Roi)]
    L.b2r_fit_exact.argtypes = [C.c_int] * 4 + [C.c_char_p, F9,
                                                C.POINTER(C.c_int),
                                                C.POINTER(C.c_int)]
    for f in (L.b2r_m33_inverse, L.b2r_rotate_matrix, L.b2r_fromto_matrix,
              L.b2r_transform_roi, L.b2r_fit_exact):
        f.restype = None
    L.b2r_fit_geometry.argtypes = [C.c_int] * 4 + [C.c_char_p] + \
        [C.POINTER(C.c_int)] * 4
    L.b2r_band_split.argtypes = [C.POINTER(_Roi), C.c_int, C.c_int,
                                 C.POINTER(_Roi)]
    L.b2r_source_rows.argtypes = [C.POINTER(_Image), C.POINTER(_Image),
                                  C.c_char_p, C.c_float, C.POINTER(_Roi),
                                  C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.c_char_p, C.c_size_t]
    L.b2r_mipchain_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Image),
                                      C.c_char_p, C.c_int, C.c_int,
                                      C.POINTER(_Options), C.c_char_p,
                                      C.c_size_t]
    L.b2r_mipchain_create_ex.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Image),
                                         C.POINTER(_MipOptions),
                                         C.POINTER(_Options), C.c_char_p,
                                         C.c_size_t]
    L.b2r_make_kernel.argtypes = [C.c_char_p, C.c_float, C.c_float, C.c_int,
                                  C.POINTER(C.c_float), C.POINTER(C.c_int),
                                  C.POINTER(C.c_int), C.c_char_p, C.c_size_t]
    L.b2r_convolve.argtypes = [C.POINTER(_Image), C.POINTER(_Image),
                               C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int,
                               C.POINTER(_Roi), C.POINTER(_Options),
                               C.POINTER(Report), C.c_char_p, C.c_size_t]
    L.b2r_unsharp_mask.argtypes = [C.POINTER(_Image), C.POINTER(_Image),
                                   C.c_char_p, C.c_float, C.c_float, C.c_float,
                                   C.POINTER(_Roi), C.POINTER(_Options),
                                   C.POINTER(Report), C.c_char_p, C.c_size_t]
    L.b2r_fix_nonfinite.argtypes = [C.POINTER(_Image), C.POINTER(_Image), C.c_int,
                                    C.POINTER(C.c_int), C.POINTER(_Roi),
                                    C.POINTER(_Options), C.POINTER(Report),
                                    C.c_char_p, C.c_size_t]
    L.b2r_fillholes_pushpull.argtypes = [C.POINTER(_Image), C.POINTER(_Image),
                                         C.c_int, C.POINTER(_Options),
                                         C.POINTER(Report), C.c_char_p,
                                         C.c_size_t]
    L.b2r_pixel_stats.argtypes = [C.POINTER(_Image), C.POINTER(_Roi),
                                  C.POINTER(_PixelStatsC), C.c_int,
                                  C.POINTER(C.c_int), C.c_float,
                                  C.POINTER(_Options), C.c_char_p, C.c_size_t]
    L.b2r_mipchain_nlevels.argtypes = [C.c_void_p]
    L.b2r_mipchain_level.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                     C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    L.b2r_mipchain_download.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Image),
                                        C.c_char_p, C.c_size_t]
    L.b2r_mipchain_kernel_ms.restype = C.c_float
    L.b2r_mipchain_kernel_ms.argtypes = [C.c_void_p]
    L.b2r_mipchain_destroy.argtypes = [C.c_void_p]
    L.b2r_pixel_hash_sha1.argtypes = [C.POINTER(_Image), C.c_char_p, C.POINTER(_Roi),
                                      C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
    L.b2r_mipchain_encode_tiles.argtypes = [C.c_void_p, C.POINTER(_TileOptions),
                                            C.POINTER(C.c_void_p), C.c_char_p, C.c_size_t]
    L.b2r_tileset_nlevels.argtypes = [C.c_void_p]
    L.b2r_tileset_level.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 4
    L.b2r_tileset_tile.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]
    L.b2r_tileset_stat.restype = C.c_double
    L.b2r_tileset_stat.argtypes = [C.c_void_p, C.c_int]
    L.b2r_tileset_destroy.argtypes = [C.c_void_p]
    L.b2r_tile_relayout.argtypes = [C.POINTER(_Image), C.c_void_p, C.c_size_t,
                                    C.POINTER(_TileOptions), C.c_void_p, C.c_char_p,
                                    C.c_size_t]
    L.b2r_tileset_write_exr.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(_ExrOptions),
                                        C.c_char_p, C.c_size_t]
    L.b2r_tileset_write_tiff.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(_TiffOptions),
                                         C.c_char_p, C.c_size_t]
    L.b2r_mipchain_create_banded.argtypes = [C.POINTER(C.c_void_p), C.POINTER(_Image),
                                             C.POINTER(_MipOptions), C.POINTER(C.c_int),
                                             C.c_int, C.c_char_p, C.c_size_t]
    L.b2r_mipchain_banded_nlevels.argtypes = [C.c_void_p]
    L.b2r_mipchain_banded_level.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int),
                                            C.POINTER(C.c_int)]
    L.b2r_mipchain_banded_band.argtypes = [C.c_void_p, C.c_int, C.c_int,
                                           C.POINTER(C.c_int), C.POINTER(C.c_int),
                                           C.POINTER(C.c_int), C.POINTER(C.c_void_p)]
    L.b2r_mipchain_banded_download.argtypes = [C.c_void_p, C.c_int, C.POINTER(_Image),
                                               C.c_char_p, C.c_size_t]
    L.b2r_mipchain_banded_ms.restype = C.c_float
    L.b2r_mipchain_banded_ms.argtypes = [C.c_void_p, C.c_int]
    L.b2r_mipchain_banded_destroy.argtypes = [C.c_void_p]
    _lib = L
    return L


def _filter_descs():
    """(name, default width) of every 2-D filter (Filter2D::get_filterdesc)."""
    L = lib()
    out = []
    for i in range(L.b2r_filter_count(2)):
        nm, w = C.c_char_p(), C.c_float()
        L.b2r_filter_desc(2, i, C.byref(nm), C.byref(w), None, None, None)
        out.append((nm.value.decode(), w.value))
    return out


def get_string_attribute(name):
    """OIIO.get_string_attribute: only "filter_list" is on this path
    (imageio.cpp builds it from Filter2D::get_filterdesc)."""
    if name == "filter_list":
        L = lib()
        names = []
        for i in range(L.b2r_filter_count(2)):
            nm = C.c_char_p()
            L.b2r_filter_desc(2, i, C.byref(nm), None, None, None, None)
            names.append(nm.value.decode())
        return ";".join(names)
    if name == "library_list":
        return "b2r:" + lib().b2r_version().decode()
    return ""


# --------------------------------------------------------------------------
class ROI:
    """imageio.h:93-230 (z kept for signature compatibility; depth 1 only)."""
    __slots__ = ("xbegin", "xend", "ybegin", "yend", "zbegin", "zend",
                 "chbegin", "chend")
    All = None  # set below

    def __init__(self, xbegin=-(2**31), xend=0, ybegin=0, yend=0, zbegin=0,
                 zend=1, chbegin=0, chend=10000):
        self.xbegin, self.xend = xbegin, xend
        self.ybegin, self.yend = ybegin, yend
        self.zbegin, self.zend = zbegin, zend
        self.chbegin, self.chend = chbegin, chend

    @property
    def defined(self):
        return self.xbegin != -(2**31)

    @property
    def width(self):
        return self.xend - self.xbegin

    @property
    def height(self):
        return self.yend - self.ybegin

    @property
    def nchannels(self):
        return self.chend - self.chbegin

    def copy(self):
        return ROI(self.xbegin, self.xend, self.ybegin, self.yend, self.zbegin,
                   self.zend, self.chbegin, self.chend)

    def __repr__(self):
        if not self.defined:
            return "ROI.All"
        return "ROI(%d,%d,%d,%d,ch %d-%d)" % (self.xbegin, self.xend,
                                               self.ybegin, self.yend,
                                               self.chbegin, self.chend)


ROI.All = ROI()


class ImageSpec:
    """The subset of ImageSpec the resize path reads (imageio.h: x, y, width,
    height, full_*, nchannels, format, alpha_channel)."""

    def __init__(self, width=0, height=0, nchannels=0, format="float"):
        self.x = self.y = 0
        self.width, self.height = width, height
        self.full_x = self.full_y = 0
        self.full_width, self.full_height = width, height
        self.nchannels = nchannels
        self.depth = 1
        self.deep = False
        self.alpha_channel = 3 if nchannels == 4 else -1
        self.z_channel = -1
        self.set_format(format)

    def set_format(self, format):
        if isinstance(format, str):
            if format not in _NAME2BT:
                raise ValueError("unsupported format " + format)
            self.basetype = _NAME2BT[format]
        elif isinstance(format, int):
            self.basetype = format
        else:
            self.basetype = _NP2BT[np.dtype(format)]

    @property
    def format(self):
        return {v: k for k, v in _NAME2BT.items()}[self.basetype]

    def copy(self):
        s = ImageSpec(self.width, self.height, self.nchannels, self.basetype)
        for k in ("x", "y", "full_x", "full_y", "full_width", "full_height",
                  "depth", "deep", "alpha_channel", "z_channel"):
            setattr(s, k, getattr(self, k))
        return s

    @property
    def roi(self):
        return ROI(self.x, self.x + self.width, self.y, self.y + self.height,
                   0, 1, 0, self.nchannels)

    @property
    def roi_full(self):
        return ROI(self.full_x, self.full_x + self.full_width, self.full_y,
                   self.full_y + self.full_height, 0, 1, 0, self.nchannels)

    def set_roi(self, r):
        self.x, self.y = r.xbegin, r.ybegin
        self.width, self.height = r.width, r.height

    def set_roi_full(self, r):
        self.full_x, self.full_y = r.xbegin, r.ybegin
        self.full_width, self.full_height = r.width, r.height


def _is_torch(t):
    return type(t).__module__.startswith("torch")


class ImageBuf:
    """Pixel container: a spec plus an (H, W, C) array that is either a numpy
    array (host) or a torch CUDA tensor (device-resident, HBM).  Only what the
    resize path needs from src/include/OpenImageIO/imagebuf.h."""

    def __init__(self, spec_or_pixels=None, spec=None):
        self._err = ""
        self.pixels = None
        self.spec_ = None
        if isinstance(spec_or_pixels, ImageSpec):
            self.reset(spec_or_pixels)
        elif spec_or_pixels is not None:
            px = spec_or_pixels
            if px.ndim == 2:
                px = px[:, :, None]
            h, w, c = px.shape
            if spec is None:
                dt = (px.dtype if not _is_torch(px)
                      else np.dtype(str(px.dtype).replace("torch.", "")))
                spec = ImageSpec(w, h, c, dt)
            assert (spec.width, spec.height, spec.nchannels) == (w, h, c)
            self.spec_ = spec
            self.pixels = px

    # -- reference-style accessors ----------------------------------------
    @property
    def initialized(self):
        return self.pixels is not None

    def spec(self):
        return self.spec_

    def specmod(self):
        return self.spec_

    @property
    def nchannels(self):
        return self.spec_.nchannels

    @property
    def roi(self):
        return self.spec_.roi

    @property
    def roi_full(self):
        return self.spec_.roi_full

    @property
    def has_error(self):
        return bool(self._err)

    def geterror(self, clear=True):
        e = self._err
        if clear:
            self._err = ""
        return e

    def errorfmt(self, msg):
        self._err = msg

    def reset(self, spec, device=None):
        self.spec_ = spec.copy()
        shape = (spec.height, spec.width, spec.nchannels)
        if device is None:
            self.pixels = np.zeros(shape, _BT2NP[spec.basetype])
        else:
            import torch
            tdt = {UINT8: torch.uint8, UINT16: torch.uint16,
                   HALF: torch.float16, FLOAT: torch.float32}[spec.basetype]
            self.pixels = torch.zeros(shape, dtype=tdt, device=device)

    @property
    def on_device(self):
        return self.pixels is not None and _is_torch(self.pixels) \
            and self.pixels.is_cuda

    def get_pixels(self):
        """Host numpy copy of the pixels."""
        if _is_torch(self.pixels):
            t = self.pixels
            if str(t.dtype) == "torch.uint16":
                return t.view(__import__("torch").int16).cpu().numpy().view(np.uint16)
            return t.cpu().numpy()
        return self.pixels

    def copy(self, src):
        self.spec_ = src.spec_.copy()
        self.pixels = (src.pixels.clone() if _is_torch(src.pixels)
                       else src.pixels.copy())
        return True

    # -- C view --------------------------------------------------------------
    def _c(self):
        s = self.spec_
        im = _Image()
        px = self.pixels
        if _is_torch(px):
            im.pixels = px.data_ptr()
            es = px.element_size()
            st = px.stride()
            im.ystride, im.xstride = st[0] * es, st[1] * es
            if px.is_cuda:
                im.memspace, im.device = MEM_DEVICE, px.device.index or 0
            else:
                im.memspace = MEM_HOST
        else:
            im.pixels = px.ctypes.data
            im.ystride, im.xstride = px.strides[0], px.strides[1]
            im.memspace = MEM_HOST
        im.basetype, im.nchannels = s.basetype, s.nchannels
        im.x, im.y, im.width, im.height = s.x, s.y, s.width, s.height
        im.full_x, im.full_y = s.full_x, s.full_y
        im.full_width, im.full_height = s.full_width, s.full_height
        return im


class Filter2D:
    """Filter2D by name (src/include/OpenImageIO/filter.h:88-156)."""

    def __init__(self, handle):
        self._h = handle

    @staticmethod
    def create(name, width, height=0.0):
        h = lib().b2r_filter2d_create(name.encode(), float(width), float(height))
        return Filter2D(h) if h else None

    def __del__(self):
        try:
            if self._h:
                lib().b2r_filter2d_destroy(self._h)
        except Exception:
            pass

    def width(self):
        return lib().b2r_filter2d_width(self._h)

    def height(self):
        return lib().b2r_filter2d_height(self._h)

    def separable(self):
        return bool(lib().b2r_filter2d_separable(self._h))

    def name(self):
        return lib().b2r_filter2d_name(self._h).decode()

    def __call__(self, x, y):
        return lib().b2r_filter2d_eval(self._h, float(x), float(y))

    def xfilt(self, x):
        return lib().b2r_filter2d_xfilt(self._h, float(x))

    def yfilt(self, y):
        return lib().b2r_filter2d_yfilt(self._h, float(y))


# --------------------------------------------------------------------------
def _ibaprep(roi, dst, src):
    """IBAprep with IBAprep_NO_SUPPORT_VOLUME | IBAprep_NO_COPY_ROI_FULL
    (imagebufalgo.cpp:64-348), restricted to one input image."""
    if src is None or not src.initialized:
        dst.errorfmt("Uninitialized input image")
        return None
    roi = roi.copy() if roi is not None and roi.defined else None
    if dst.initialized:
        d = dst.spec_.roi
        if roi is not None:
            roi = ROI(max(roi.xbegin, d.xbegin), min(roi.xend, d.xend),
                      max(roi.ybegin, d.ybegin), min(roi.yend, d.yend), 0, 1,
                      max(roi.chbegin, 0), min(roi.chend, d.chend))
        else:
            roi = d
    else:
        full = None
        if roi is None:
            roi, full = src.roi, src.roi_full
        else:
            roi.chend = min(roi.chend, src.nchannels)
        spec = src.spec_.copy()
        spec.set_roi(roi)
        spec.set_roi_full(full if full is not None else roi)
        dev = src.pixels.device if src.on_device else None
        dst.reset(spec, device=dev)
    if dst.spec_.depth > 1 or src.spec_.depth > 1:
        dst.errorfmt("volumes not supported")
        return None
    if dst.spec_.deep or src.spec_.deep:
        dst.errorfmt("deep images not supported")
        return None
    roi.chend = min(roi.chend, max(dst.nchannels, src.nchannels))
    return roi


def _options(device=0, path=PATH_AUTO, stream=None, highlightcomp=False,
             alpha_channel=-1, z_channel=-1, contract_fma=False):
    o = _Options()
    o.contract_fma = 1 if contract_fma else 0
    o.device = device
    o.path = path
    o.cuda_stream = stream
    o.highlightcomp = 1 if highlightcomp else 0
    o.alpha_channel1 = alpha_channel + 1 if alpha_channel >= 0 else 0
    o.z_channel1 = z_channel + 1 if z_channel >= 0 else 0
    return o


def _current_stream(buf):
    if buf.on_device:
        import torch
        return torch.cuda.current_stream(buf.pixels.device).cuda_stream
    return None


class _IBA:
    """ImageBufAlgo functions on this path.  Two call shapes, like the
    reference's Python binding: f(dst, src, ...) -> bool, f(src, ...) ->
    ImageBuf (py_imagebufalgo.cpp:1470-1560)."""

    last_report = None
    force_report = False
    # b2r_options::contract_fma for the warp family (warp / rotate / fit exact /
    # resize from-to): packed FMAs in the accumulation, results within 1e-5 /
    # 1 LSB of the reference instead of its exact rounding sequence
    contract_fma = False

    # ---- resize -----------------------------------------------------------
    def resize(self, *args, filtername="", filterwidth=0.0, roi=None,
               nthreads=0, filterptr=None, path=PATH_AUTO, device=0,
               highlightcomp=False):
        """highlightcomp: oiiotool's `--resize:highlightcomp=1` recipe
        (oiiotool.cpp:4644-4663) run on the device around the resize."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        if rest:  # positional filtername / filterwidth
            filtername = rest[0]
            if len(rest) > 1:
                filterwidth = rest[1]
        if len(bufs) == 2:
            return self._resize(bufs[0], bufs[1], filtername, filterwidth, roi,
                                filterptr, path, device, highlightcomp)
        result = ImageBuf()
        ok = self._resize(result, bufs[0], filtername, filterwidth, roi,
                          filterptr, path, device, highlightcomp)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::resize() error")
        return result

    def resize_batch(self, dsts, srcs, filtername="", filterwidth=0.0, roi=None,
                     device=0, report=False):
        """Resize srcs[i] into dsts[i] (lists of ImageBuf of ONE shape).  Frames
        that live on one device run as a single persistent launch
        (b2r_resize_batch); anything else falls back to frame-by-frame."""
        if len(dsts) != len(srcs):
            raise ValueError("resize_batch: %d destinations for %d sources"
                             % (len(dsts), len(srcs)))
        if not dsts:
            return True
        n = len(dsts)
        r0 = _ibaprep(roi, dsts[0], srcs[0])
        if r0 is None:
            return False
        for d, s_ in zip(dsts[1:], srcs[1:]):
            if _ibaprep(roi, d, s_) is None:
                return False
        L = lib()
        D = (_Image * n)(*[d._c() for d in dsts])
        S = (_Image * n)(*[s_._c() for s_ in srcs])
        r = _Roi(r0.xbegin, r0.xend, r0.ybegin, r0.yend, r0.chbegin, r0.chend)
        stream = _current_stream(dsts[0]) or _current_stream(srcs[0])
        o = _options(device, PATH_AUTO, stream, False, srcs[0].spec_.alpha_channel,
                     srcs[0].spec_.z_channel)
        rep = Report()
        err = C.create_string_buffer(512)
        rc = L.b2r_resize_batch(D, S, n, (filtername or "").encode(), float(filterwidth),
                                C.byref(r), C.byref(o), C.byref(rep) if report else None,
                                err, 512)
        self.last_report = rep if report else None
        if rc:
            dsts[0].errorfmt(err.value.decode())
            return False
        return True

    def _resize(self, dst, src, filtername, filterwidth, roi, filterptr, path,
                device, highlightcomp=False):
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return False
        L = lib()
        d, s = dst._c(), src._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        stream = _current_stream(dst) or _current_stream(src)
        o = _options(device, path, stream, highlightcomp,
                     src.spec_.alpha_channel, src.spec_.z_channel)
        rep = Report()
        err = C.create_string_buffer(512)
        # device-resident calls skip the report (it costs two events and a
        # synchronise) unless force_report is set (tests, probes)
        want_report = self.force_report or not (dst.on_device and src.on_device)
        if filterptr is not None:
            rc = L.b2r_resize_filter(C.byref(d), C.byref(s), filterptr._h,
                                     C.byref(r), C.byref(o),
                                     C.byref(rep) if want_report else None,
                                     err, 512)
        else:
            rc = L.b2r_resize(C.byref(d), C.byref(s),
                              (filtername or "").encode(), float(filterwidth),
                              C.byref(r), C.byref(o),
                              C.byref(rep) if want_report else None, err, 512)
        self.last_report = rep if want_report else None
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    # ---- rangecompress / rangeexpand -------------------------------------
    def rangecompress(self, *args, useluma=False, roi=None, nthreads=0,
                      device=0):
        """ImageBufAlgo::rangecompress (imagebufalgo_pixelmath.cpp:748-759)."""
        return self._range_call(False, args, useluma, roi, device)

    def rangeexpand(self, *args, useluma=False, roi=None, nthreads=0, device=0):
        """ImageBufAlgo::rangeexpand (imagebufalgo_pixelmath.cpp:763-774)."""
        return self._range_call(True, args, useluma, roi, device)

    def _range_call(self, expand, args, useluma, roi, device):
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        if rest:
            useluma = bool(rest[0])
        if len(bufs) == 2:
            return self._range(expand, bufs[0], bufs[1], useluma, roi, device)
        result = ImageBuf()
        ok = self._range(expand, result, bufs[0], useluma, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::%s() error"
                            % ("rangeexpand" if expand else "rangecompress"))
        return result

    def _range(self, expand, dst, src, useluma, roi, device):
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return False
        roi.chend = min(roi.chend, dst.nchannels, src.nchannels)
        L = lib()
        d, s = dst._c(), src._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        stream = _current_stream(dst) or _current_stream(src)
        o = _options(device, PATH_AUTO, stream)
        err = C.create_string_buffer(512)
        f = L.b2r_rangeexpand if expand else L.b2r_rangecompress
        rc = f(C.byref(d), C.byref(s), 1 if useluma else 0,
               src.spec_.alpha_channel, src.spec_.z_channel, C.byref(r),
               C.byref(o), None, err, 512)
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    # ---- resample ---------------------------------------------------------
    def resample(self, *args, interpolate=True, roi=None, nthreads=0, device=0):
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        if rest:
            interpolate = bool(rest[0])
        if len(bufs) == 2:
            return self._resample(bufs[0], bufs[1], interpolate, roi, device)
        result = ImageBuf()
        ok = self._resample(result, bufs[0], interpolate, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::resample() error")
        return result

    def _resample(self, dst, src, interpolate, roi, device):
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return False
        L = lib()
        d, s = dst._c(), src._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        stream = _current_stream(dst) or _current_stream(src)
        o = _options(device, PATH_AUTO, stream)
        err = C.create_string_buffer(512)
        rc = L.b2r_resample(C.byref(d), C.byref(s), 1 if interpolate else 0,
                            C.byref(r), C.byref(o), None, err, 512)
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    # ---- fit --------------------------------------------------------------
    def fit(self, *args, filtername="", filterwidth=0.0, fillmode="letterbox",
            exact=False, roi=None, nthreads=0, device=0):
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        if rest:
            filtername = rest[0]
            if len(rest) > 1:
                filterwidth = rest[1]
        if len(bufs) == 2:
            return self._fit(bufs[0], bufs[1], filtername, filterwidth, fillmode,
                             exact, roi, device)
        result = ImageBuf()
        ok = self._fit(result, bufs[0], filtername, filterwidth, fillmode, exact,
                       roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::fit() error")
        return result

    def _fit(self, dst, src, filtername, filterwidth, fillmode, exact, roi,
             device):
        # imagebufalgo_xform.cpp:933-1062
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return False
        sspec = src.spec_
        L = lib()
        if exact:
            # :1017-1027: partial-pixel fit = a warp by (scale, centring
            # offset) with black wrap and edgeclamp, through the filter
            # get_resize_filter picks for the whole-pixel resize ratios
            M = (C.c_float * 9)()
            rw, rh = C.c_int(), C.c_int()
            L.b2r_fit_exact(sspec.full_width, sspec.full_height, roi.width,
                            roi.height, fillmode.encode(), M, C.byref(rw),
                            C.byref(rh))
            wratio = np.float32(rw.value) / np.float32(sspec.full_width)
            hratio = np.float32(rh.value) / np.float32(sspec.full_height)
            name = filtername or ("blackman-harris"
                                  if (wratio > 1.0 or hratio > 1.0) else "lanczos3")
            filt = None
            for fd in _filter_descs():
                if fd[0] == name:
                    w = filterwidth if filterwidth > 0 else \
                        np.float32(fd[1]) * max(np.float32(1), wratio)
                    h = filterwidth if filterwidth > 0 else \
                        np.float32(fd[1]) * max(np.float32(1), hratio)
                    filt = Filter2D.create(name, float(w), float(h))
                    break
            if filt is None:
                dst.errorfmt('Filter "%s" not recognized' % name)
                return False
            newroi = ROI(roi.xbegin, roi.xend, roi.ybegin, roi.yend, 0, 1, 0,
                         sspec.nchannels)
            newspec = sspec.copy()
            newspec.set_roi(newroi)
            newspec.set_roi_full(newroi)
            dev = src.pixels.device if src.on_device else None
            dst.reset(newspec, device=dev)
            return self._warp(dst, src, list(M), "", 0.0, filt, False, "black",
                              True, None, device)
        rw, rh, xo, yo = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.b2r_fit_geometry(sspec.full_width, sspec.full_height, roi.width,
                           roi.height, fillmode.encode(), C.byref(rw),
                           C.byref(rh), C.byref(xo), C.byref(yo))
        fit_x, fit_y = roi.xbegin, roi.ybegin
        ok = True
        # the filter is chosen from the RESIZE ratios (:1005-1015)
        if (rw.value != sspec.full_width or rh.value != sspec.full_height
                or fit_x != sspec.full_x or fit_y != sspec.full_y):
            resizeroi = ROI(fit_x, fit_x + rw.value, fit_y, fit_y + rh.value,
                            0, 1, 0, sspec.nchannels)
            newspec = sspec.copy()
            newspec.set_roi(resizeroi)
            newspec.set_roi_full(resizeroi)
            dev = src.pixels.device if src.on_device else None
            dst.reset(newspec, device=dev)
            ok = self._resize(dst, src, filtername, filterwidth, resizeroi,
                              None, PATH_AUTO, device)
        else:
            ok = dst.copy(src)
        sp = dst.spec_
        sp.full_width, sp.full_height = roi.width, roi.height
        sp.full_x, sp.full_y = fit_x, fit_y
        sp.x, sp.y = xo.value, yo.value
        return ok


    # ---- make_kernel / convolve / unsharp_mask ---------------------------
    def make_kernel(self, name, width, height=None, depth=1.0, normalize=True):
        """ImageBufAlgo::make_kernel (imagebufalgo.cpp:864-957): a 1-channel
        float ImageBuf whose data window is centred on the origin.  An
        unknown name gives a box with the error set, like the reference."""
        if height is None:
            height = width
        L = lib()
        kw, kh = C.c_int(), C.c_int()
        err = C.create_string_buffer(512)
        L.b2r_make_kernel(name.encode(), float(width), float(height),
                          1 if normalize else 0, None, C.byref(kw), C.byref(kh),
                          err, 512)
        arr = np.zeros((kh.value, kw.value, 1), np.float32)
        rc = L.b2r_make_kernel(name.encode(), float(width), float(height),
                               1 if normalize else 0,
                               arr.ctypes.data_as(C.POINTER(C.c_float)),
                               C.byref(kw), C.byref(kh), err, 512)
        K = ImageBuf(arr)
        sp = K.spec_
        sp.x = sp.full_x = -(kw.value // 2)
        sp.y = sp.full_y = -(kh.value // 2)
        if rc:
            K.errorfmt(err.value.decode())
        return K

    def convolve(self, *args, normalize=True, roi=None, nthreads=0, device=0):
        """ImageBufAlgo::convolve(dst, src, kernel, normalize, roi, nthreads)
        (imagebufalgo.cpp:817-850)."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        if rest:
            normalize = bool(rest[0])
        if len(bufs) == 3:
            return self._convolve(bufs[0], bufs[1], bufs[2], normalize, roi, device)
        result = ImageBuf()
        ok = self._convolve(result, bufs[0], bufs[1], normalize, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::convolve() error")
        return result

    def _stencil_prep(self, roi, dst, src):
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return None
        if dst.nchannels != src.nchannels:
            dst.errorfmt("images must have the same number of channels")
            return None
        roi.chend = min(roi.chend, dst.nchannels)
        return roi

    def _convolve(self, dst, src, kernel, normalize, roi, device):
        roi = self._stencil_prep(roi, dst, src)
        if roi is None:
            return False
        k = np.ascontiguousarray(kernel.get_pixels()[..., 0], np.float32)
        d, s = dst._c(), src._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        o = _options(device, PATH_AUTO,
                     _current_stream(dst) or _current_stream(src))
        err = C.create_string_buffer(512)
        rc = lib().b2r_convolve(C.byref(d), C.byref(s),
                                k.ctypes.data_as(C.POINTER(C.c_float)),
                                k.shape[1], k.shape[0], 1 if normalize else 0,
                                C.byref(r), C.byref(o), None, err, 512)
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    def unsharp_mask(self, *args, kernel="gaussian", width=3.0, contrast=1.0,
                     threshold=0.0, roi=None, nthreads=0, device=0):
        """ImageBufAlgo::unsharp_mask(dst, src, kernel, width, contrast,
        threshold, roi, nthreads) (imagebufalgo.cpp:975-1045)."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        names = ["kernel", "width", "contrast", "threshold"]
        vals = dict(kernel=kernel, width=width, contrast=contrast,
                    threshold=threshold)
        for n, v in zip(names, rest):
            vals[n] = v
        if len(bufs) == 2:
            return self._unsharp(bufs[0], bufs[1], vals, roi, device)
        result = ImageBuf()
        ok = self._unsharp(result, bufs[0], vals, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::unsharp_mask() error")
        return result

    def _unsharp(self, dst, src, v, roi, device):
        roi = self._stencil_prep(roi, dst, src)
        if roi is None:
            return False
        d, s = dst._c(), src._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        o = _options(device, PATH_AUTO,
                     _current_stream(dst) or _current_stream(src))
        err = C.create_string_buffer(512)
        rc = lib().b2r_unsharp_mask(C.byref(d), C.byref(s), v["kernel"].encode(),
                                    float(v["width"]), float(v["contrast"]),
                                    float(v["threshold"]), C.byref(r), C.byref(o),
                                    None, err, 512)
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    # ---- computePixelStats / isMonochrome ---------------------------------
    def computePixelStats(self, src, roi=None, nthreads=0, device=0,
                          mono_threshold=0.0):
        """ImageBufAlgo::computePixelStats(src, roi, nthreads)
        (imagebufalgo_compare.cpp:187-208); the same pass also answers
        isMonochrome (`.monochrome`).  An empty PixelStats on error."""
        if src is None or not src.initialized:
            return PixelStats()
        n = src.nchannels
        recs = (_PixelStatsC * n)()
        mono = C.c_int(1)
        s = src._c()
        o = _options(device, PATH_AUTO, _current_stream(src))
        err = C.create_string_buffer(512)
        r = None
        if roi is not None and roi.defined:
            r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                     min(roi.chend, n))
        rc = lib().b2r_pixel_stats(C.byref(s), C.byref(r) if r else None, recs, n,
                                   C.byref(mono), float(mono_threshold),
                                   C.byref(o), err, 512)
        if rc:
            src.errorfmt(err.value.decode())
            return PixelStats()
        return PixelStats(list(recs), bool(mono.value))

    def isMonochrome(self, src, threshold=0.0, roi=None, nthreads=0, device=0):
        """ImageBufAlgo::isMonochrome (imagebufalgo_compare.cpp:579-600)."""
        st = self.computePixelStats(src, roi, device=device,
                                    mono_threshold=threshold)
        return bool(st.monochrome)

    def isConstantColor(self, src, threshold=0.0, roi=None, nthreads=0, device=0):
        """The exact (threshold 0) form maketx uses: constant iff min == max
        in every channel (maketexture.cpp:1425).  Returns the colour or None."""
        if threshold != 0.0:
            raise B2RError("isConstantColor: only threshold 0 is on this path")
        st = self.computePixelStats(src, roi, device=device)
        if st.min and st.min == st.max and not any(st.nancount) \
                and not any(st.infcount):
            return list(st.min)
        return None

    # ---- fillholes_pushpull ----------------------------------------------
    def computePixelHashSHA1(self, src, extrainfo="", roi=None, blocksize=0, nthreads=0):
        """ImageBufAlgo::computePixelHashSHA1 (imagebufalgo_compare.cpp:856-892):
        upper-case hex SHA-1 of the pixels (+ extrainfo), in `blocksize`-row
        blocks when blocksize > 0.  Host threads; needs no GPU for host images."""
        if not isinstance(src, ImageBuf):
            src = ImageBuf(src)
        im = src._c()
        r = None
        if roi is not None and roi.defined:
            r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin, roi.chend)
        out = C.create_string_buffer(41)
        err = C.create_string_buffer(512)
        rc = lib().b2r_pixel_hash_sha1(C.byref(im), (extrainfo or "").encode(),
                                       C.byref(r) if r is not None else None,
                                       int(blocksize), int(nthreads), out, err, 512)
        if rc:
            raise B2RError(err.value.decode())
        return out.value.decode()

    def fillholes_pushpull(self, *args, roi=None, nthreads=0, device=0):
        """ImageBufAlgo::fillholes_pushpull(dst, src, roi, nthreads)
        (imagebufalgo.cpp:1600-1682)."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        if len(bufs) == 2:
            return self._fillholes(bufs[0], bufs[1], roi, device)
        result = ImageBuf()
        ok = self._fillholes(result, bufs[0], roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::fillholes_pushpull() error")
        return result

    def _fillholes(self, dst, src, roi, device):
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return False
        if dst.nchannels != src.nchannels:
            dst.errorfmt("images must have the same number of channels")
            return False
        if src.spec_.alpha_channel < 0 or dst.spec_.alpha_channel < 0:
            dst.errorfmt("images must have alpha channels")
            return False
        d, s = dst._c(), src._c()
        o = _options(device, PATH_AUTO,
                     _current_stream(dst) or _current_stream(src))
        err = C.create_string_buffer(512)
        rc = lib().b2r_fillholes_pushpull(C.byref(d), C.byref(s),
                                          src.spec_.alpha_channel, C.byref(o),
                                          None, err, 512)
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    # ---- fixNonFinite -----------------------------------------------------
    last_pixels_fixed = 0

    def fixNonFinite(self, *args, mode=None, roi=None, nthreads=0, device=0):
        """ImageBufAlgo::fixNonFinite(dst, src, mode=NONFINITE_BOX3, roi,
        nthreads) (imagebufalgo_pixelmath.cpp:1562-1625).  The number of
        repaired pixels (the C++ `pixelsFixed` out-parameter) is left in
        `ImageBufAlgo.last_pixels_fixed`."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        if mode is None:
            mode = rest[0] if rest else NONFINITE_BOX3
        if len(bufs) == 2:
            return self._fixnf(bufs[0], bufs[1], mode, roi, device)
        result = ImageBuf()
        ok = self._fixnf(result, bufs[0], mode, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("fixNonFinite error")
        return result

    def _fixnf(self, dst, src, mode, roi, device):
        if mode not in (NONFINITE_NONE, NONFINITE_BLACK, NONFINITE_BOX3,
                        NONFINITE_ERROR):
            dst.errorfmt("fixNonFinite: unknown repair mode")
            return False
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return False
        roi.chend = min(roi.chend, dst.nchannels, src.nchannels)
        d, s = dst._c(), src._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        o = _options(device, PATH_AUTO,
                     _current_stream(dst) or _current_stream(src))
        err = C.create_string_buffer(512)
        n = C.c_int(0)
        rc = lib().b2r_fix_nonfinite(C.byref(d), C.byref(s), int(mode),
                                     C.byref(n), C.byref(r), C.byref(o), None,
                                     err, 512)
        self.last_pixels_fixed = n.value
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    # ---- warp / rotate ----------------------------------------------------
    def warp(self, *args, filtername="", filterwidth=0.0, recompute_roi=False,
             wrap="default", edgeclamp=False, filterptr=None, roi=None,
             nthreads=0, device=0):
        """ImageBufAlgo::warp(dst, src, M, KWArgs{filtername, filterwidth,
        wrap, edgeclamp, recompute_roi, filterptr}, roi, nthreads)
        (imagebufalgo_xform.cpp:468-503).  M: 9 floats, Imath::M33f order."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        M = rest[0]
        if len(rest) > 1:
            filtername = rest[1]
        if len(rest) > 2:
            filterwidth = rest[2]
        if len(bufs) == 2:
            return self._warp(bufs[0], bufs[1], M, filtername, filterwidth,
                              filterptr, recompute_roi, wrap, edgeclamp, roi,
                              device)
        result = ImageBuf()
        ok = self._warp(result, bufs[0], M, filtername, filterwidth, filterptr,
                        recompute_roi, wrap, edgeclamp, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::warp() error")
        return result

    def _warp(self, dst, src, M, filtername, filterwidth, filterptr,
              recompute_roi, wrap, edgeclamp, roi, device):
        # warp_impl (imagebufalgo_xform.cpp:379-418)
        if src is None or not src.initialized:
            dst.errorfmt("Uninitialized input image")
            return False
        L = lib()
        m9 = (C.c_float * 9)(*[float(v) for v in np.asarray(M, np.float32).reshape(9)])
        if filterptr is None:
            name = filtername or "lanczos3"
            if name not in [fd[0] for fd in _filter_descs()]:
                dst.errorfmt('Filter "%s" not recognized' % name)
                return False
        if isinstance(wrap, str):
            if wrap not in WRAP_MODES:
                wrap = "default"  # WrapMode_from_string
            wrap = WRAP_MODES[wrap]
        have_roi = roi is not None and roi.defined
        if dst.initialized:
            dst_roi = roi.copy() if have_roi else dst.roi
        elif have_roi:
            dst_roi = roi.copy()
        elif recompute_roi:
            sr = src.roi
            r = _Roi(sr.xbegin, sr.xend, sr.ybegin, sr.yend, sr.chbegin, sr.chend)
            o = _Roi()
            L.b2r_transform_roi(m9, C.byref(r), C.byref(o))
            dst_roi = ROI(o.xbegin, o.xend, o.ybegin, o.yend, 0, 1, sr.chbegin,
                          sr.chend)
        else:
            dst_roi = src.roi
        dst_roi.chend = min(dst_roi.chend, src.nchannels)
        if not dst.initialized:
            # IBAprep sets full = src's full window for an uninitialized dst
            spec = src.spec_.copy()
            spec.set_roi(dst_roi)
            spec.set_roi_full(src.roi_full)
            spec.nchannels = dst_roi.chend
            dev = src.pixels.device if src.on_device else None
            dst.reset(spec, device=dev)
        roi = _ibaprep(dst_roi, dst, src)
        if roi is None:
            return False
        roi.chend = min(roi.chend, src.nchannels)
        d, s = dst._c(), src._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        stream = _current_stream(dst) or _current_stream(src)
        o = _options(device, PATH_AUTO, stream, contract_fma=self.contract_fma)
        err = C.create_string_buffer(512)
        rep = Report()
        want_report = self.force_report or not (dst.on_device and src.on_device)
        if filterptr is not None:
            rc = L.b2r_warp_filter(C.byref(d), C.byref(s), m9, filterptr._h,
                                   int(wrap), 1 if edgeclamp else 0, C.byref(r),
                                   C.byref(o),
                                   C.byref(rep) if want_report else None, err, 512)
        else:
            rc = L.b2r_warp(C.byref(d), C.byref(s), m9,
                            (filtername or "").encode(), float(filterwidth),
                            int(wrap), 1 if edgeclamp else 0, C.byref(r),
                            C.byref(o), C.byref(rep) if want_report else None,
                            err, 512)
        self.last_report = rep if want_report else None
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    def st_warp(self, *args, filtername="", filterwidth=0.0, chan_s=0, chan_t=1,
                flip_s=False, flip_t=False, filterptr=None, roi=None, nthreads=0,
                device=0):
        """ImageBufAlgo::st_warp(dst, src, stbuf, filtername, filterwidth,
        chan_s, chan_t, flip_s, flip_t, roi, nthreads)
        (imagebufalgo_xform.cpp:1976-2058)."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        if rest:
            filtername = rest[0]
        if len(rest) > 1:
            filterwidth = rest[1]
        if len(bufs) == 3:
            return self._st_warp(bufs[0], bufs[1], bufs[2], filtername, filterwidth,
                                 filterptr, chan_s, chan_t, flip_s, flip_t, roi,
                                 device)
        result = ImageBuf()
        ok = self._st_warp(result, bufs[0], bufs[1], filtername, filterwidth,
                           filterptr, chan_s, chan_t, flip_s, flip_t, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::st_warp : Unknown error")
        return result

    def _st_warp(self, dst, src, st, filtername, filterwidth, filterptr, chan_s,
                 chan_t, flip_s, flip_t, roi, device):
        if filterptr is None:
            name = filtername or "lanczos3"
            if name not in [fd[0] for fd in _filter_descs()]:
                dst.errorfmt('Filter "%s" not recognized' % name)
                return False
        # check_st_warp_args (:1929-1972)
        if st is None or not st.initialized:
            dst.errorfmt("ImageBufAlgo::st_warp : Uninitialized ST buffer")
            return False
        if chan_s >= st.nchannels:
            dst.errorfmt("ImageBufAlgo::st_warp : Out-of-range S channel index: %d"
                         % chan_s)
            return False
        if chan_t >= st.nchannels:
            dst.errorfmt("ImageBufAlgo::st_warp : Out-of-range T channel index: %d"
                         % chan_t)
            return False
        roi = _ibaprep(roi, dst, src)
        if roi is None:
            return False
        roi.chend = min(roi.chend, dst.nchannels, src.nchannels)
        d, s, t = dst._c(), src._c(), st._c()
        r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin,
                 roi.chend)
        stream = _current_stream(dst) or _current_stream(src)
        o = _options(device, PATH_AUTO, stream)
        err = C.create_string_buffer(512)
        if filterptr is not None:
            rc = lib().b2r_st_warp_filter(C.byref(d), C.byref(s), C.byref(t),
                                          filterptr._h, chan_s, chan_t,
                                          1 if flip_s else 0, 1 if flip_t else 0,
                                          C.byref(r), C.byref(o), None, err, 512)
        else:
            rc = lib().b2r_st_warp(C.byref(d), C.byref(s), C.byref(t),
                                   (filtername or "").encode(), float(filterwidth),
                                   chan_s, chan_t, 1 if flip_s else 0,
                                   1 if flip_t else 0, C.byref(r), C.byref(o), None,
                                   err, 512)
        if rc:
            dst.errorfmt(err.value.decode())
            return False
        return True

    def rotate(self, *args, filtername="", filterwidth=0.0, recompute_roi=False,
               center=None, filterptr=None, roi=None, nthreads=0, device=0):
        """ImageBufAlgo::rotate(dst, src, angle[, center_x, center_y], ...)
        (imagebufalgo_xform.cpp:1727-1830): angle in radians, default centre =
        centre of src's display window, wrap black."""
        bufs = [a for a in args if isinstance(a, ImageBuf)]
        rest = [a for a in args if not isinstance(a, ImageBuf)]
        angle = float(rest[0])
        if len(rest) >= 3 and not isinstance(rest[1], str):
            center = (float(rest[1]), float(rest[2]))
            rest = rest[:1] + rest[3:]
        if len(rest) > 1:
            filtername = rest[1]
        if len(rest) > 2:
            filterwidth = rest[2]
        src = bufs[-1]
        if src is None or not src.initialized:
            dst = bufs[0] if len(bufs) == 2 else ImageBuf()
            dst.errorfmt("Uninitialized input image")
            return False if len(bufs) == 2 else dst
        if center is None:
            f = src.roi_full
            center = (np.float32(0.5) * np.float32(f.xbegin + f.xend),
                      np.float32(0.5) * np.float32(f.ybegin + f.yend))
        M = rotate_matrix(angle, center[0], center[1])
        if len(bufs) == 2:
            return self._warp(bufs[0], src, M, filtername, filterwidth,
                              filterptr, recompute_roi, "black", False, roi,
                              device)
        result = ImageBuf()
        ok = self._warp(result, src, M, filtername, filterwidth, filterptr,
                        recompute_roi, "black", False, roi, device)
        if not ok and not result.has_error:
            result.errorfmt("ImageBufAlgo::rotate() error")
        return result


ImageBufAlgo = _IBA()

# ImageBufAlgo::NonFiniteFixMode (imagebufalgo.h)
NONFINITE_NONE, NONFINITE_BLACK, NONFINITE_BOX3, NONFINITE_ERROR = 0, 1, 2, 100
# ImageBufAlgo::MakeTextureMode (imagebufalgo.h)
MakeTxTexture, MakeTxShadow, MakeTxEnvLatl = 0, 1, 2

# ImageBuf::WrapMode (src/include/OpenImageIO/imagebuf.h)
WRAP_MODES = {"default": 0, "black": 1, "clamp": 2, "periodic": 3, "mirror": 4}


def _m9out():
    return (C.c_float * 9)()


def m33_inverse(M):
    """Imath::M33f::inverse() as warp_ uses it."""
    out = _m9out()
    m = (C.c_float * 9)(*[float(v) for v in np.asarray(M, np.float32).reshape(9)])
    lib().b2r_m33_inverse(m, out)
    return np.array(list(out), np.float32).reshape(3, 3)


def rotate_matrix(angle, center_x, center_y):
    """The matrix ImageBufAlgo::rotate builds (imagebufalgo_xform.cpp:1733-1737)."""
    out = _m9out()
    lib().b2r_rotate_matrix(float(angle), float(center_x), float(center_y), out)
    return np.array(list(out), np.float32).reshape(3, 3)


def fromto_matrix(frm, to):
    """oiiotool --resize:from=:to= matrix (oiiotool.cpp:4704-4707);
    frm, to = (x, y, w, h)."""
    out = _m9out()
    lib().b2r_fromto_matrix(*[float(v) for v in frm], *[float(v) for v in to], out)
    return np.array(list(out), np.float32).reshape(3, 3)


def oiiotool_resize(src, width, height, filtername="", from_geom=None,
                    to_geom=None, highlightcomp=False, edgeclamp=False,
                    device=0):
    """The compute step of `oiiotool --resize[:from=..][:to=..][:filter=..]
    [:highlightcomp=1][:edgeclamp=1] WxH` (OpResize, src/oiiotool/oiiotool.cpp:
    4581-4735) for an image whose display window starts at the origin:
    a separable resize, or -- when from/to differ from the full windows -- a
    warp by the from->to matrix.  from_geom / to_geom: (x, y, w, h)."""
    A = src.spec_
    newspec = A.copy()
    newspec.full_width, newspec.full_height = width, height
    newspec.x, newspec.y = newspec.full_x, newspec.full_y
    newspec.width, newspec.height = width, height
    frm = (A.full_x, A.full_y, A.full_width, A.full_height)
    to = (newspec.full_x, newspec.full_y, width, height)
    f = tuple(float(v) for v in (from_geom or frm))
    t = tuple(float(v) for v in (to_geom or to))
    do_warp = f != tuple(map(float, frm)) or t != tuple(map(float, to))
    IBA = ImageBufAlgo
    if not do_warp:
        wr = np.float32(width) / np.float32(A.full_width)
        hr = np.float32(height) / np.float32(A.full_height)
        newspec.x = newspec.full_x + int(np.floor(np.float32(A.x - A.full_x) * wr))
        newspec.y = newspec.full_y + int(np.floor(np.float32(A.y - A.full_y) * hr))
        newspec.width = int(np.ceil(np.float32(A.width) * wr))
        newspec.height = int(np.ceil(np.float32(A.height) * hr))
    dst = ImageBuf()
    dst.reset(newspec, device=src.pixels.device if src.on_device else None)
    if not do_warp:
        ok = IBA.resize(dst, src, filtername=filtername, roi=dst.roi,
                        device=device, highlightcomp=highlightcomp)
        return dst if ok else dst
    s = src
    if highlightcomp:
        s = IBA.rangecompress(src, device=device)
    ok = IBA.warp(dst, s, fromto_matrix(f, t), filtername=filtername,
                  recompute_roi=False, edgeclamp=edgeclamp, device=device)
    if highlightcomp and ok:
        IBA.rangeexpand(dst, dst, device=device)
    return dst


# --------------------------------------------------------------------------
# Row-band sharding across GPUs / ranks.  The reference's only parallel
# strategy is full-width row bands (parallel_image,
# imagebufalgo_util.h:37-80); the same split is used across devices.  Bands
# are independent: each device stages its band of source rows plus the filter
# halo straight from the host image, so there is no collective in the data
# path.
def band_split(roi, nbands, band):
    """Band `band` of `nbands` of a dst ROI (rows only)."""
    r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin, roi.chend)
    o = _Roi()
    if lib().b2r_band_split(C.byref(r), nbands, band, C.byref(o)):
        raise B2RError("band_split: bad arguments")
    return ROI(o.xbegin, o.xend, o.ybegin, o.yend, 0, 1, o.chbegin, o.chend)


def source_rows(dst, src, filtername="", filterwidth=0.0, roi=None):
    """Source rows [begin, end) (relative to src's data window) that resizing
    into `roi` of dst reads: the band plus its filter-radius halo.  Host
    arithmetic only; works without a GPU."""
    d, s = dst._c(), src._c()
    if roi is None:
        roi = dst.roi
    r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin, roi.chend)
    b, e = C.c_int(), C.c_int()
    err = C.create_string_buffer(512)
    rc = lib().b2r_source_rows(C.byref(d), C.byref(s),
                               (filtername or "").encode(), float(filterwidth),
                               C.byref(r), C.byref(b), C.byref(e), err, 512)
    if rc:
        raise B2RError(err.value.decode())
    return b.value, e.value


def resize_banded(dst, src, devices, filtername="", filterwidth=0.0, roi=None):
    """ImageBufAlgo.resize of one large HOST image over several GPUs of a
    box (b2r_resize_banded): dst rows are split into len(devices) bands, one
    b2r_resize per band on its own device, issued from one host thread each
    inside the library.  Returns True / False like resize()."""
    roi = _ibaprep(roi, dst, src)
    if roi is None:
        return False
    if dst.on_device or src.on_device:
        dst.errorfmt("resize_banded shards host images; device-resident "
                     "images are already on one GPU")
        return False
    d, s = dst._c(), src._c()
    r = _Roi(roi.xbegin, roi.xend, roi.ybegin, roi.yend, roi.chbegin, roi.chend)
    devs = (C.c_int * len(devices))(*[int(x) for x in devices])
    o = _options(0, PATH_AUTO, None)
    err = C.create_string_buffer(512)
    # the C ABI drives the bands: one host thread per device inside b2r_resize_banded
    rc = lib().b2r_resize_banded(C.byref(d), C.byref(s), (filtername or "").encode(),
                                 float(filterwidth), C.byref(r), devs, len(devices),
                                 C.byref(o), err, 512)
    if rc:
        dst.errorfmt(err.value.decode())
        return False
    return True


def trim_caches():
    """b2r_trim_caches: give the library's cached device / pinned memory (band
    buffers of destroyed banded chains, tile staging slots, weight-table plans)
    back to the driver."""
    L = lib()
    L.b2r_trim_caches.restype = None
    L.b2r_trim_caches()


def shard_frames(nframes, world_size, rank):
    """Frames of a batch are independent: round-robin over ranks (one
    process per GPU), no exchange."""
    return list(range(rank, nframes, world_size))


class MipChain:
    """Device-resident MIP pyramid: the level loop of write_mipmap
    (src/libOpenImageIO/maketexture.cpp:823-986) without host round trips."""

    def __init__(self, level0, filtername="box", clamp_half=False, device=0,
                 highlightcomp=False, sharpen=0.0, sharpen_first=False,
                 sharpenfilt="gaussian"):
        """highlightcomp: maketx --hicomp (rangecompress / rangeexpand + clamp
        around every filtered level, maketexture.cpp:907-938).  sharpen:
        maketx --sharpen <contrast> (unsharp mask before / after each level's
        resize, :909-931)."""
        if not isinstance(level0, ImageBuf):
            level0 = ImageBuf(level0)
        L = lib()
        h = C.c_void_p()
        im = level0._c()
        o = _options(device, PATH_AUTO, _current_stream(level0), highlightcomp,
                     level0.spec_.alpha_channel,
                     getattr(level0.spec_, "z_channel", -1))
        err = C.create_string_buffer(512)
        mo = _MipOptions()
        mo.filtername = filtername.encode()
        mo.clamp_half = 1 if clamp_half else 0
        mo.alpha_channel = level0.spec_.alpha_channel
        mo.sharpen = float(sharpen)
        mo.sharpen_first = 1 if sharpen_first else 0
        mo.sharpenfilt = (sharpenfilt or "gaussian").encode()
        rc = L.b2r_mipchain_create_ex(C.byref(h), C.byref(im), C.byref(mo),
                                      C.byref(o), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        self._h = h
        self._keep = level0  # level 0 may be adopted, keep it alive
        self.nchannels = level0.nchannels

    def __del__(self):
        try:
            if self._h:
                lib().b2r_mipchain_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @classmethod
    def streamed(cls, level0, filtername="box", clamp_half=False, device=0, outputs=None):
        """b2r_mipchain_create_streamed: a HOST level 0 up in row bands, level 1
        computed and downloaded band by band while the upload is still running,
        levels >= 2 from the device.  Returns (chain, [level 1, level 2, ...]) with
        the levels as float32 numpy arrays (`outputs` supplies them; else they
        are allocated here)."""
        if not isinstance(level0, ImageBuf):
            level0 = ImageBuf(level0)
        L = lib()
        sizes = []
        w, h = level0.spec_.width, level0.spec_.height
        while w > 1 or h > 1:
            w, h = (w // 2 if w > 1 else w), (h // 2 if h > 1 else h)
            sizes.append((h, w))
        c = level0.nchannels
        if outputs is None:
            outputs = [np.empty((hh, ww, c), np.float32) for hh, ww in sizes]
        bufs = [ImageBuf(a) for a in outputs]
        arr = (type(level0._c()) * len(bufs))(*[b._c() for b in bufs])
        self = cls.__new__(cls)
        hnd = C.c_void_p()
        im = level0._c()
        o = _options(device, PATH_AUTO, None)
        mo = _MipOptions()
        mo.filtername = filtername.encode()
        mo.clamp_half = 1 if clamp_half else 0
        mo.alpha_channel = level0.spec_.alpha_channel
        mo.sharpenfilt = b"gaussian"
        err = C.create_string_buffer(512)
        rc = L.b2r_mipchain_create_streamed(C.byref(hnd), C.byref(im), C.byref(mo), C.byref(o),
                                            arr, len(bufs), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        self._h = hnd
        self._keep = level0
        self.nchannels = c
        return self, outputs

    @property
    def nlevels(self):
        return lib().b2r_mipchain_nlevels(self._h)

    @property
    def kernel_ms(self):
        return lib().b2r_mipchain_kernel_ms(self._h)

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        lib().b2r_mipchain_level(self._h, level, C.byref(w), C.byref(h), None)
        return w.value, h.value

    def level_ptr(self, level):
        p = C.c_void_p()
        lib().b2r_mipchain_level(self._h, level, None, None, C.byref(p))
        return p.value

    def download(self, level):
        w, h = self.level_size(level)
        out = np.empty((h, w, self.nchannels), np.float32)
        ib = ImageBuf(out)
        im = ib._c()
        err = C.create_string_buffer(512)
        rc = lib().b2r_mipchain_download(self._h, level, C.byref(im), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        return out


class TileSet:
    """Compressed tiles of every level of a MipChain (b2r_mipchain_encode_tiles:
    the level output stage of write_mipmap, maketexture.cpp:957-974, up to the
    compressed tiles a file plugin stores)."""

    def __init__(self, chain, tile=64, half=False, compression=True, zlevel=1,
                 nthreads=0, first_level=0, predictor=0, out_type=None):
        """out_type: "float" (default), "half" (same as half=True), "uint16" or
        "uint8" (the levels through convert_type<float,D>).  predictor: 0, 3
        (TIFF floating point; float / half) or 2 (TIFF horizontal; integers)."""
        out_type = out_type or ("half" if half else "float")
        half = out_type == "half"
        o = _TileOptions()
        o.predictor = int(predictor)
        o.tile_width = o.tile_height = int(tile)
        o.out_basetype = _TILE_TYPES[out_type][0]
        o.compression = 1 if compression else 0
        o.zlevel = int(zlevel)
        o.nthreads = int(nthreads)
        o.first_level = int(first_level)
        h = C.c_void_p()
        err = C.create_string_buffer(512)
        rc = lib().b2r_mipchain_encode_tiles(chain._h, C.byref(o), C.byref(h), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        self._h = h
        self.tile = int(tile)
        self.half = bool(half)
        self.compressed = bool(compression)
        self.predictor = int(predictor) if int(predictor) in (2, 3) else 0
        self.out_type = out_type
        self.nchannels = chain.nchannels

    def __del__(self):
        try:
            if self._h:
                lib().b2r_tileset_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def nlevels(self):
        return lib().b2r_tileset_nlevels(self._h)

    def level_info(self, level):
        v = [C.c_int() for _ in range(4)]
        lib().b2r_tileset_level(self._h, level, *[C.byref(x) for x in v])
        return tuple(x.value for x in v)  # width, height, ntiles_x, ntiles_y

    def tile_bytes(self, level, tx, ty):
        p, n = C.c_void_p(), C.c_size_t()
        if lib().b2r_tileset_tile(self._h, level, tx, ty, C.byref(p), C.byref(n)):
            raise B2RError("no such tile")
        return C.string_at(p.value, n.value)

    def tile_pixels(self, level, tx, ty):
        """The tile decoded back into a tile x tile x nchannels array."""
        import zlib
        b = self.tile_bytes(level, tx, ty)
        if self.compressed:
            b = zlib.decompress(b)
        dt = _TILE_TYPES[self.out_type][1]
        if self.predictor == 2:
            # undo TIFF's horizontal predictor (libtiff horAcc8 / horAcc16)
            px = np.frombuffer(b, dt).reshape(self.tile, self.tile, self.nchannels)
            return np.cumsum(px, axis=1, dtype=dt)
        if self.predictor == 3:
            # undo TIFF's floating-point predictor row by row (libtiff fpAcc)
            bps = 2 if self.half else 4
            wc = self.tile * self.nchannels
            rows = np.frombuffer(b, np.uint8).reshape(self.tile, wc * bps).copy()
            for k in range(self.nchannels, wc * bps):
                rows[:, k] += rows[:, k - self.nchannels]
            planes = rows.reshape(self.tile, bps, wc)[:, ::-1, :]
            b = np.ascontiguousarray(planes.transpose(0, 2, 1)).tobytes()
        return np.frombuffer(b, dt).reshape(self.tile, self.tile, self.nchannels)

    def write_tiff(self, path, image_description=None, software=None, wrapmodes=None,
                   bigtiff=False):
        """The tile set as a tiled TIFF texture, one directory per MIP level
        (b2r_tileset_write_tiff; what `out->open(AppendMIPLevel)` +
        `small->write(out)` leave to the TIFF plugin, maketexture.cpp:957-974)."""
        o = _TiffOptions()
        o.image_description = image_description.encode() if image_description else None
        o.software = software.encode() if software else None
        o.wrapmodes = wrapmodes.encode() if wrapmodes else None
        o.bigtiff = 1 if bigtiff else 0
        err = C.create_string_buffer(512)
        rc = lib().b2r_tileset_write_tiff(self._h, str(path).encode(), C.byref(o), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        return True

    def write_exr(self, path, attributes=None, zlevel=1, nthreads=0):
        """The tile set (stored tiles: compression=False, float or half) as a
        tiled, MIP-mapped, zip-compressed OpenEXR file (b2r_tileset_write_exr).
        attributes: {name: string} extra header attributes."""
        items = list((attributes or {}).items())
        o = _ExrOptions()
        names = (C.c_char_p * max(1, len(items)))(*[str(k).encode() for k, _ in items])
        vals = (C.c_char_p * max(1, len(items)))(*[str(v).encode() for _, v in items])
        o.attr_names = C.cast(names, C.POINTER(C.c_char_p))
        o.attr_values = C.cast(vals, C.POINTER(C.c_char_p))
        o.nattrs = len(items)
        o.zlevel, o.nthreads = int(zlevel), int(nthreads)
        err = C.create_string_buffer(512)
        rc = lib().b2r_tileset_write_exr(self._h, str(path).encode(), C.byref(o), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        return True

    def stat(self, what):
        k = {"total_ms": 0, "device_ms": 1, "deflate_ms": 2, "raw_bytes": 3,
             "stored_bytes": 4, "setup_ms": 5, "wait_ms": 6}[what]
        return lib().b2r_tileset_stat(self._h, k)


def tile_relayout(src, tile=64, half=False, predictor=0, out_type=None):
    """b2r_tile_relayout: a contiguous float32 CUDA tensor (h, w, c) re-laid into
    `tile` x `tile` tiles on the device -> uint8 CUDA tensor of shape
    (ntiles_y, ntiles_x, tile bytes); asynchronous on the current torch stream."""
    import torch
    img = src if isinstance(src, ImageBuf) else ImageBuf(src)
    if not img.on_device:
        raise B2RError("tile_relayout: the source is a contiguous float32 device image")
    h, w, c = img.spec_.height, img.spec_.width, img.spec_.nchannels
    ntx, nty = -(-w // tile), -(-h // tile)
    out_type = out_type or ("half" if half else "float")
    tb = tile * tile * c * np.dtype(_TILE_TYPES[out_type][1]).itemsize
    out = torch.empty((nty, ntx, tb), dtype=torch.uint8, device=img.pixels.device)
    o = _TileOptions()
    o.tile_width = o.tile_height = int(tile)
    o.out_basetype = _TILE_TYPES[out_type][0]
    o.predictor = int(predictor)
    im = img._c()
    err = C.create_string_buffer(512)
    rc = lib().b2r_tile_relayout(C.byref(im), C.c_void_p(out.data_ptr()),
                                 C.c_size_t(out.numel()), C.byref(o),
                                 _current_stream(img), err, 512)
    if rc:
        raise B2RError(err.value.decode())
    return out


class MipChainBanded:
    """The MIP chain cut into row bands over several GPUs of one box, driven
    from this process (b2r_mipchain_create_banded: per-band upload of level 0
    from the host image, halo rows pushed between devices over NVLink, tail of
    the chain gathered on the first device).  Levels equal MipChain's bit for
    bit.  `devices`: CUDA ordinals, one band per entry."""

    def __init__(self, level0, devices, filtername="box", clamp_half=False):
        if not isinstance(level0, ImageBuf):
            level0 = ImageBuf(level0)
        L = lib()
        h = C.c_void_p()
        im = level0._c()
        mo = _MipOptions()
        mo.filtername = filtername.encode()
        mo.clamp_half = 1 if clamp_half else 0
        mo.alpha_channel = level0.spec_.alpha_channel
        devs = (C.c_int * len(devices))(*devices)
        err = C.create_string_buffer(512)
        rc = L.b2r_mipchain_create_banded(C.byref(h), C.byref(im), C.byref(mo),
                                          devs, len(devices), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        self._h = h
        self.nchannels = level0.nchannels
        self.devices = list(devices)

    def __del__(self):
        try:
            if self._h:
                lib().b2r_mipchain_banded_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def nlevels(self):
        return lib().b2r_mipchain_banded_nlevels(self._h)

    def ms(self, what="total"):
        """'total': host wall clock of the construction; 'upload' / 'chain':
        device events, max over the devices; 'halo_bytes'."""
        k = {"total": 0, "upload": 1, "chain": 2, "halo_bytes": 3, "setup": 4,
             "phase_a": 5, "phase_b": 6}[what]
        return lib().b2r_mipchain_banded_ms(self._h, k)

    def level_size(self, level):
        w, h = C.c_int(), C.c_int()
        lib().b2r_mipchain_banded_level(self._h, level, C.byref(w), C.byref(h))
        return w.value, h.value

    def band(self, level, band):
        d, y0, y1, p = C.c_int(), C.c_int(), C.c_int(), C.c_void_p()
        lib().b2r_mipchain_banded_band(self._h, level, band, C.byref(d), C.byref(y0),
                                       C.byref(y1), C.byref(p))
        return d.value, y0.value, y1.value, p.value

    def download(self, level, out=None):
        w, h = self.level_size(level)
        if out is None:
            out = np.empty((h, w, self.nchannels), np.float32)
        ib = ImageBuf(out)
        im = ib._c()
        err = C.create_string_buffer(512)
        rc = lib().b2r_mipchain_banded_download(self._h, level, C.byref(im), err, 512)
        if rc:
            raise B2RError(err.value.decode())
        return out


# --------------------------------------------------------------------------
# make_texture: the pixel pipeline of make_texture_impl + write_mipmap
# (src/libOpenImageIO/maketexture.cpp:1396-2040, 811-987) for plain textures,
# without the file: every stage that touches pixels runs on the device and the
# finished levels are handed back instead of being written.
class Texture:
    """What make_texture computed: `levels[0]` is the top level (after
    --fixnan / --resize), `levels[1:]` the MIP chain; `format` the pixel type
    the reference would write them as."""

    def __init__(self):
        self.levels = []
        self.format = "float"
        self.pixels_fixed = 0
        self.stats = None           # PixelStats of the (repaired) source
        self.average_color = None   # "oiio:AverageColor"
        self.constant_color = None  # "oiio:ConstantColor" when min == max
        self.monochrome = None
        self.error = ""

    @property
    def ok(self):
        return not self.error


def _ceil2(x):
    p = 1
    while p < x:
        p <<= 1
    return p


def _write_texture_file(T, outputfilename, cfg):
    """The tail of make_texture_impl (maketexture.cpp:1884-1990 metadata,
    :957-974 level output): every level as 64x64 (config "tile_width") tiles,
    deflated, floating-point predictor, into a tiled TIFF whose
    ImageDescription carries oiio:SHA-1 / ConstantColor / AverageColor the way
    the reference crams them in for formats without arbitrary metadata."""
    if T.format not in _TILE_TYPES:
        T.error = ("make_texture: \"%s\" output is not on the B200 path (tiles are "
                   "float, half, uint16 or uint8)" % T.format)
        return T
    desc = []
    top = T.levels[0]
    if int(cfg.get("maketx:hash", 1)) and not top.on_device:
        # addlHashData (:1909-1926): the options that change the MIP levels
        extra = (cfg.get("maketx:filtername", "box") or "box") + " "
        if float(cfg.get("maketx:sharpen", 0.0) or 0.0):
            extra += "sharpen_A=%g " % float(cfg["maketx:sharpen"])
        if cfg.get("maketx:highlightcomp"):
            extra += "highlightcomp=1 "
        if cfg.get("maketx:keepaspect"):
            extra += "keepaspect=1 "
        T.sha1 = ImageBufAlgo.computePixelHashSHA1(top, extra, blocksize=256)
        desc.append("oiio:SHA-1=" + T.sha1)
    join = lambda v: ",".join(str(np.float32(x)) for x in v)
    if T.constant_color is not None:
        desc.append("oiio:ConstantColor=" + join(T.constant_color))
    if T.average_color is not None and int(cfg.get("maketx:compute_average", 1)):
        desc.append("oiio:AverageColor=" + join(T.average_color))
    try:
        if str(outputfilename).lower().endswith(".exr"):
            if T.format not in ("float", "half"):
                raise B2RError("make_texture: OpenEXR textures are float or half, not \"%s\""
                               % T.format)
            # OpenEXR holds arbitrary metadata: one attribute each (:1936-1975)
            attrs = {"textureformat": "Plain Texture",
                     "wrapmodes": cfg.get("wrapmodes", None) or "black,black"}
            for d in desc:
                k, v = d.split("=", 1)
                attrs[k] = v
            ts = TileSet(T.chain, tile=int(cfg.get("tile_width", 64) or 64),
                         out_type=T.format, compression=False)
            ts.write_exr(outputfilename, attributes=attrs)
        else:
            ts = TileSet(T.chain, tile=int(cfg.get("tile_width", 64) or 64),
                         out_type=T.format, zlevel=int(cfg.get("tiff:zipquality", 1) or 1),
                         predictor=3 if T.format in ("float", "half") else 2)
            ts.write_tiff(outputfilename, image_description=" ".join(desc) or None,
                          wrapmodes=cfg.get("wrapmodes", None))
    except B2RError as e:
        T.error = str(e)
        return T
    T.tiles = ts
    T.filename = str(outputfilename)
    return T


def make_texture(mode, src, config=None, device=0, outputfilename=None):
    """ImageBufAlgo::make_texture(mode, input, outputfilename, configspec) for
    mode MakeTxTexture, returning a `Texture`; with `outputfilename` the
    levels are also written as a tiled TIFF texture (_write_texture_file).
    config: dict of the maketx attributes that select pixel work --
    "maketx:filtername" ("box"; "post-" / "unsharp-<f>" prefixes as in
    maketexture.cpp:759-768), "maketx:highlightcomp", "maketx:sharpen",
    "maketx:fixnan" ("none" | "black" | "box3"), "maketx:checknan",
    "maketx:resize" (round up to a power of two), "maketx:nomipmap",
    "maketx:allow_pixel_shift", "format" (output type: levels written as half
    are clamped to +-HALF_MAX, :709-714, 943-945)."""
    cfg = dict(config or {})
    T = Texture()
    IBA = ImageBufAlgo
    if mode != MakeTxTexture:
        T.error = "make_texture: only MakeTxTexture is on the B200 resampling path"
        return T
    if src is None or not src.initialized:
        T.error = "make_texture: uninitialized input image"
        return T
    sspec = src.spec_
    # --fixnan / --checknan (:1666-1705)
    fixnan = cfg.get("maketx:fixnan", "") or "none"
    if fixnan not in ("none", "black", "box3"):
        T.error = 'Unknown fixnan mode "%s"' % fixnan
        return T
    isfp = sspec.basetype in (HALF, FLOAT)
    if fixnan != "none" and isfp:
        fixed = IBA.fixNonFinite(src, {"black": NONFINITE_BLACK,
                                       "box3": NONFINITE_BOX3}[fixnan],
                                 device=device)
        if fixed.has_error:
            T.error = "Error fixing nans/infs."
            return T
        T.pixels_fixed = IBA.last_pixels_fixed
        src = fixed
    if cfg.get("maketx:checknan") and isfp:
        probe = ImageBuf()
        IBA.fixNonFinite(probe, src, NONFINITE_NONE, device=device)
        if IBA.last_pixels_fixed:
            T.error = "maketx ERROR: Nan/Inf at %d pixels" % IBA.last_pixels_fixed
            return T
    # pixel statistics (:1396-1409): average colour, constant-colour and
    # monochrome detection, one pass on the device
    if any(cfg.get(k, d) for k, d in (("maketx:compute_average", 1),
                                      ("maketx:constant_color_detect", 0),
                                      ("maketx:opaque_detect", 0),
                                      ("maketx:monochrome_detect", 0))):
        T.stats = IBA.computePixelStats(src, device=device)
        if not T.stats.min:
            T.error = src.geterror() or "computePixelStats failed"
            return T
        T.average_color = list(T.stats.avg)
        T.constant_color = (list(T.stats.min) if T.stats.min == T.stats.max
                            else None)
        T.monochrome = T.stats.monochrome
    filtername = cfg.get("maketx:filtername", "box") or "box"
    # top level: resize to a power of two and/or convert to float (:1791-1877)
    dw, dh = sspec.width, sspec.height
    if cfg.get("maketx:resize"):
        dw, dh = _ceil2(dw), _ceil2(dh)
    do_resize = (dw, dh) != (sspec.width, sspec.height)
    allow_shift = bool(cfg.get("maketx:allow_pixel_shift"))
    dev = src.pixels.device if src.on_device else None
    top_spec = sspec.copy()
    top_spec.width = top_spec.full_width = dw
    top_spec.height = top_spec.full_height = dh
    top_spec.x = top_spec.y = top_spec.full_x = top_spec.full_y = 0
    top_spec.set_format("float")          # maketx:forcefloat (default 1)
    s0 = src.spec_.copy()
    s0.x = s0.y = s0.full_x = s0.full_y = 0
    s0.full_width, s0.full_height = s0.width, s0.height
    src0 = ImageBuf(src.pixels, s0)
    # float copy of the source (maketx:forcefloat; resize_block needs float):
    # copy_pixels == a nearest resample at 1:1
    if s0.basetype != FLOAT:
        fspec = s0.copy()
        fspec.set_format("float")
        srcf = ImageBuf()
        srcf.reset(fspec, device=dev)
        if not IBA.resample(srcf, src0, interpolate=False, device=device):
            T.error = srcf.geterror()
            return T
    else:
        srcf = src0
    if not do_resize:
        top = srcf
    else:
        top = ImageBuf()
        rf = "lanczos3" if filtername.lower().startswith("unsharp-") else filtername
        if rf in ("box", "triangle"):
            # resize_block (:141-334) works on device-resident float levels
            import torch
            if srcf.on_device:
                sdev = srcf
            else:
                sdev = ImageBuf(torch.from_numpy(srcf.pixels).to("cuda:%d" % device),
                                srcf.spec_.copy())
            top.reset(top_spec, device=sdev.pixels.device)
            if not _mip_level(top, sdev, "box", device):
                T.error = top.geterror()
                return T
            if not srcf.on_device:
                top = ImageBuf(top.pixels.cpu().numpy(), top_spec)
        else:
            top.reset(top_spec, device=dev)
            if not IBA.resize(top, srcf, filtername=rf, device=device):
                T.error = top.geterror()
                return T
    out_format = cfg.get("format", None) or {UINT8: "uint8", UINT16: "uint16",
                                             HALF: "half", FLOAT: "float"}[sspec.basetype]
    T.format = out_format
    clamp_half = out_format == "half"
    T.levels = [top]
    if cfg.get("maketx:nomipmap"):
        if outputfilename:
            T.error = ("make_texture: file output of a texture without MIP levels "
                       "(maketx:nomipmap) is not on the B200 path")
        return T
    sharpen = float(cfg.get("maketx:sharpen", 0.0) or 0.0)
    sharpen_first, sharpenfilt = True, "gaussian"
    fn = filtername
    if fn.lower().startswith("post-"):
        sharpen_first, fn = False, fn[5:]
    if fn.lower().startswith("unsharp-"):
        sharpenfilt, fn = fn[8:], "lanczos3"
    try:
        chain = MipChain(top, fn, clamp_half=clamp_half, device=device,
                         highlightcomp=bool(cfg.get("maketx:highlightcomp")),
                         sharpen=sharpen, sharpen_first=sharpen_first,
                         sharpenfilt=sharpenfilt)
    except B2RError as e:
        T.error = str(e)
        return T
    for lvl in range(1, chain.nlevels):
        T.levels.append(ImageBuf(chain.download(lvl)))
    T.chain = chain
    if outputfilename:
        return _write_texture_file(T, outputfilename, cfg)
    return T


def _mip_level(dst, src, filtername, device):
    """One write_mipmap level step on whole images (b2r_mip_level)."""
    d, s = dst._c(), src._c()
    o = _options(device, PATH_AUTO, _current_stream(dst) or _current_stream(src))
    err = C.create_string_buffer(512)
    rc = lib().b2r_mip_level(C.byref(d), C.byref(s), filtername.encode(), None,
                             C.byref(o), err, 512)
    if rc:
        dst.errorfmt(err.value.decode())
        return False
    return True
