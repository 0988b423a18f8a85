#!/usr/bin/env python
"""Convert the reference's own golden images for the resize/resample/fit path
into compact .npz fixtures (the GPU box has no /root/reference).

Run in the build container:   python tests/golden/make_golden.py

Sources (all under /root/reference/testsuite, OpenImageIO 3.2.0.2-dev):
  filters/ref/<filter>.exr            17 known-answer images written by
                                      filters/run.py (delta 16x16 -> 80x80, half)
  common/grid.tif                     1000x1000 RGBA uint8 input
  common/tahoe-tiny.tif               input of fit3
  oiiotool-xform/ref/resize.tif       --resize 256x256
  oiiotool-xform/ref/resize2.tif      --resize 25%
  oiiotool-xform/ref/resize64.tif     --resize 64x64
  oiiotool-xform/ref/resize512.tif    resize64 -> --resize 512x512
  oiiotool-xform/ref/resized-offset.exr  nonzero origin, half
  oiiotool-xform/ref/resample.tif, resample-nearest.tif   --resample 128x128
  oiiotool-xform/ref/fit.tif, fit2.tif, fit3.tif           --fit
  common/tahoe-small.tif + oiiotool/ref/range{compress,expand}{,-luma}.tif
  oiiotool-xform/ref/{warped,rotated,rotated-offcenter,resizefrom*}.tif,
  oiiotool-xform/src/target1.exr, ref/fit4.exr, two of ref/fit[wh]-*.exr
  oiiotool/ref/bspline-blur.tif, gauss5x5-blur.tif, unsharp.tif
  oiiotool/src/hole.tif, oiiotool/ref/tahoe-filled.tif, growholes.tif
Only pixel arrays are stored (no reference source code is copied).
"""
import os
import sys

os.environ["OPENCV_IO_ENABLE_OPENEXR"] = "1"
import cv2  # noqa: E402
import numpy as np  # noqa: E402

REF = "/root/reference/testsuite"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(path):
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    if img is None:
        raise RuntimeError("cannot read " + path)
    if img.ndim == 3:
        if img.shape[2] == 3:
            img = img[:, :, ::-1]
        elif img.shape[2] == 4:
            img = img[:, :, [2, 1, 0, 3]]
    return np.ascontiguousarray(img)


def main():
    filt = {}
    fdir = os.path.join(REF, "filters/ref")
    for f in sorted(os.listdir(fdir)):
        if f.endswith(".exr"):
            a = read(os.path.join(fdir, f))
            if a.ndim == 3:  # single-channel EXR may be expanded by cv2
                assert np.array_equal(a[..., 0], a[..., 1])
                a = a[..., 0]
            # stored as half on disk; cv2 returns float32 -- exact
            h = a.astype(np.float16)
            assert np.array_equal(h.astype(np.float32), a)
            filt[f[:-4]] = h
    np.savez_compressed(os.path.join(HERE, "filters_ref.npz"), **filt)
    print("filters:", sorted(filt))

    x = {}
    x["grid"] = read(os.path.join(REF, "common/grid.tif"))
    x["tahoe_tiny"] = read(os.path.join(REF, "common/tahoe-tiny.tif"))
    xdir = os.path.join(REF, "oiiotool-xform/ref")
    for name in ["resize", "resize2", "resize64", "resize512", "resample",
                 "resample-nearest", "fit", "fit2", "fit3"]:
        x[name.replace("-", "_")] = read(os.path.join(xdir, name + ".tif"))
    ro = read(os.path.join(xdir, "resized-offset.exr"))
    x["resized_offset"] = ro.astype(np.float16)
    assert np.array_equal(x["resized_offset"].astype(np.float32), ro)
    for k, v in x.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(HERE, "xform_ref.npz"), **x)

    # rangecompress / rangeexpand known answers (testsuite/oiiotool/run.py:28-32)
    pm = {"tahoe_small": read(os.path.join(REF, "common/tahoe-small.tif"))}
    for name in ["rangecompress", "rangeexpand", "rangecompress-luma",
                 "rangeexpand-luma"]:
        pm[name.replace("-", "_")] = read(os.path.join(REF, "oiiotool/ref",
                                                       name + ".tif"))
    for k, v in pm.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(HERE, "pixelmath_ref.npz"), **pm)

    # warp / rotate / resize:from-to / fit:exact known answers
    # (testsuite/oiiotool-xform/run.py); inputs resize.tif and grid.tif are in
    # xform_ref.npz already
    wp = {}
    for name in ["warped", "rotated", "rotated-offcenter", "st_warped", "resizefrom",
                 "resizefromto", "resizefromtooffset"]:
        wp[name.replace("-", "_")] = read(os.path.join(xdir, name + ".tif"))
    for name, path in [("target1", "oiiotool-xform/src/target1.exr"),
                       ("fit4", "oiiotool-xform/ref/fit4.exr"),
                       ("fitw_letterbox_200", "oiiotool-xform/ref/fitw-letterbox-200x200.exr"),
                       ("fith_width_300", "oiiotool-xform/ref/fith-width-300x300.exr")]:
        a = read(os.path.join(REF, path))
        h = a.astype(np.float16)
        assert np.array_equal(h.astype(np.float32), a), name
        wp[name] = h
    for k, v in wp.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(HERE, "warp_ref.npz"), **wp)

    # make_kernel / convolve / blur / unsharp known answers
    # (testsuite/oiiotool/run.py:159-182); the input tahoe-small.tif is in
    # pixelmath_ref.npz
    cv = {}
    # (bsplinekernel.exr has a negative data-window origin that cv2's EXR
    # reader refuses; bspline-blur.tif pins the same kernel through convolve)
    for name in ["bspline-blur", "gauss5x5-blur", "unsharp"]:
        cv[name.replace("-", "_")] = read(os.path.join(REF, "oiiotool/ref", name + ".tif"))
    for k, v in cv.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(HERE, "conv_ref.npz"), **cv)

    # fillholes_pushpull known answers (testsuite/oiiotool/run.py:124-126)
    fh = {"hole": read(os.path.join(REF, "oiiotool/src/hole.tif")),
          "tahoe_filled": read(os.path.join(REF, "oiiotool/ref/tahoe-filled.tif")),
          "growholes": read(os.path.join(REF, "oiiotool/ref/growholes.tif"))}
    for k, v in fh.items():
        print(k, v.shape, v.dtype)
    np.savez_compressed(os.path.join(HERE, "fillholes_ref.npz"), **fh)


if __name__ == "__main__":
    sys.exit(main())
