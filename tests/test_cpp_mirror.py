"""The C++ host mirror (include/b2r_imagebufalgo.h): compiles against the C
ABI, reproduces the reference's error behaviour, and (GPU) matches the oracle."""
import os
import subprocess

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def iba_cli(oiio, tmp_path_factory):
    out = str(tmp_path_factory.mktemp("cpp") / "iba_cli")
    pkg = os.path.join(ROOT, "openimageio_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I",
                           os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "iba_cli.cpp"),
                           "-o", out, "-L", pkg, "-lb2r", "-Wl,-rpath," + pkg])
    return out


def test_cpp_error_behaviour(iba_cli):
    out = subprocess.check_output([iba_cli, "errors"], text=True).splitlines()
    assert out[0] == '1:Filter "catrom" not recognized'
    assert out[1] == "2:Uninitialized input image"
    assert out[2] == "3:volumes not supported"
    assert out[3] == '4:Filter "lanczos" not recognized'
    assert out[4] == '5:Filter "nope" not recognized'


@pytest.mark.gpu
def test_cpp_resize_matches_oracle(iba_cli, oracle, tmp_path):
    src = synth.image(180, 120, 4, np.float32, seed=91)
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    src.tofile(fin)
    subprocess.check_call([iba_cli, "resize", fin, "180", "120", "4", "11",
                           "90", "60", "lanczos3", fout])
    got = np.fromfile(fout, np.float32).reshape(60, 90, 4)
    ref = np.zeros_like(got)
    oracle.resize(oracle.Img(ref), oracle.Img(src), "lanczos3")
    assert np.abs(got - ref).max() <= 1e-5


@pytest.mark.gpu
def test_cpp_highlightcomp_three_calls_equal_the_fused_option(iba_cli, oracle, oiio,
                                                               tmp_path):
    """rangecompress -> resize -> rangeexpand written as three ImageBufAlgo
    calls (what oiiotool does, oiiotool.cpp:4644-4663) against the oracle, and
    against the one-call `highlightcomp` option of the ABI."""
    src = (synth.image(180, 120, 4, np.float32, seed=92) ** 4 * np.float32(25)
           ).astype(np.float32)
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    src.tofile(fin)
    subprocess.check_call([iba_cli, "hicomp", fin, "180", "120", "4", "11",
                           "90", "60", "lanczos3", fout])
    got = np.fromfile(fout, np.float32).reshape(60, 90, 4)
    tmp = src.copy()
    oracle.rangecompress(oracle.Img(tmp), oracle.Img(src), alpha_channel=3)
    ref = np.zeros_like(got)
    oracle.resize(oracle.Img(ref), oracle.Img(tmp), "lanczos3")
    oracle.rangeexpand(oracle.Img(ref), oracle.Img(ref), alpha_channel=3)
    scale = max(1.0, float(np.abs(ref).max()))
    assert np.abs(got - ref).max() <= 5e-5 * scale
    one = np.zeros_like(got)
    assert oiio.ImageBufAlgo.resize(oiio.ImageBuf(one), oiio.ImageBuf(src),
                                    filtername="lanczos3", highlightcomp=True)
    assert np.array_equal(one, got)


@pytest.mark.gpu
def test_cpp_rotate_and_fit_exact_match_oracle(iba_cli, oracle, tmp_path):
    import math
    src = synth.image(120, 90, 3, np.float32, seed=93)
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    src.tofile(fin)
    subprocess.check_call([iba_cli, "rotate", fin, "120", "90", "3", "11", "120", "90",
                           "lanczos3", fout])
    got = np.fromfile(fout, np.float32).reshape(90, 120, 3)
    ref = np.zeros_like(got)
    M = oracle.rotate_matrix(np.float32(30.0) * np.float32(math.pi / 180.0), 60.0, 45.0)
    oracle.warp(oracle.Img(ref), oracle.Img(src), M, "lanczos3", wrap="black")
    assert np.abs(got - ref).max() <= 1e-5
    subprocess.check_call([iba_cli, "fitexact", fin, "120", "90", "3", "11", "100", "50",
                           "", fout])
    got = np.fromfile(fout, np.float32).reshape(50, 100, 3)
    Mx, rw, rh = oracle.fit_exact_matrix(120, 90, 100, 50)
    ref = np.zeros_like(got)
    oracle.warp(oracle.Img(ref), oracle.Img(src), Mx, "lanczos3", 6.0, 6.0, wrap="black",
                edgeclamp=True)
    assert np.abs(got - ref).max() <= 1e-5


@pytest.mark.gpu
def test_cpp_widened_surface(iba_cli, oracle, tmp_path):
    """fixNonFinite -> unsharp_mask -> convolve -> computePixelStats written
    against the C++ mirror, checked bit for bit against the oracle chain."""
    src = synth.image(90, 70, 3, np.float32, seed=94)
    src[10, 20, 1] = np.nan
    src[40, 50, 2] = np.inf
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    src.tofile(fin)
    p = subprocess.run([iba_cli, "sharpen", fin, "90", "70", "3", "11", "90", "70", "box",
                        fout], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    got = np.fromfile(fout, np.float32).reshape(70, 90, 3)
    a = np.zeros_like(src)
    assert oracle.fix_nonfinite(oracle.Img(a), oracle.Img(src), "box3") == 2
    b = np.zeros_like(src)
    oracle.unsharp_mask(oracle.Img(b), oracle.Img(a), "gaussian", 3.0, 0.7, 0.0)
    K, _ = oracle.make_kernel("box", 3, 3)
    ref = np.zeros_like(src)
    oracle.convolve(oracle.Img(ref), oracle.Img(b), K)
    assert np.array_equal(got, ref)
    st = oracle.pixel_stats(ref)
    assert "fixed=2" in p.stderr and "finite0=%d" % st[0]["finitecount"] in p.stderr
    assert "mono=0" in p.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("filt,hicomp", [("box", 0), ("lanczos3", 0), ("lanczos3", 1)])
def test_cpp_make_texture(iba_cli, oracle, tmp_path, filt, hicomp):
    """ImageBufAlgo::make_texture of the C++ mirror: the levels against the
    oracle's write_mipmap loop, oiio:SHA-1 against a hashlib restatement of
    computePixelHashSHA1 (256-row blocks + "<filter> [highlightcomp=1 ]")."""
    import hashlib
    src = synth.image(300, 520, 4, np.float32, seed=93)
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    src.tofile(fin)
    out = subprocess.check_output([iba_cli, "maketx", fin, "300", "520", "4", "11", filt,
                                   str(hicomp), fout], text=True)
    kv = dict(t.split("=") for t in out.split())
    levels = [src] + oracle.mip_chain(src, filt, highlightcomp=bool(hicomp), alpha_channel=3)
    assert int(kv["levels"]) == len(levels)
    got = np.fromfile(fout, np.float32)
    off = 0
    for lv, ref in enumerate(levels):
        n = ref.size
        g = got[off:off + n].reshape(ref.shape)
        off += n
        if filt == "box" or lv == 0:
            assert np.array_equal(g, ref), lv
        else:
            assert np.abs(g - ref).max() <= (5e-5 if hicomp else 1e-5) * max(1, lv), lv
    s = hashlib.sha1()
    for y in range(0, 520, 256):
        s.update(hashlib.sha1(src[y:y + 256].tobytes()).hexdigest().upper().encode())
    s.update((filt + " " + ("highlightcomp=1 " if hicomp else "")).encode())
    assert kv["sha1"] == s.hexdigest().upper()
    assert abs(float(kv["avg0"]) - float(src[..., 0].astype(np.float64).mean())) < 1e-6


@pytest.mark.gpu
def test_cpp_make_texture_writes_exr(iba_cli, tmp_path):
    """make_texture(..., "out.exr") of the C++ mirror: every level of the OpenEXR
    file equals the float levels it returned; SHA-1 as a string attribute."""
    import exr_mini
    src = synth.image(200, 330, 4, np.float32, seed=97)
    fin, fout, fexr = (str(tmp_path / n) for n in ("in.raw", "out.raw", "out.exr"))
    src.tofile(fin)
    out = subprocess.check_output([iba_cli, "maketx", fin, "200", "330", "4", "11", "lanczos3",
                                   "0", fout, fexr], text=True)
    kv = dict(t.split("=") for t in out.split())
    levels, attrs = exr_mini.read(fexr)
    assert len(levels) == int(kv["levels"])
    got = np.fromfile(fout, np.float32)
    off = 0
    for lv in levels:
        assert lv.dtype == np.float32 and np.array_equal(lv.ravel(), got[off:off + lv.size])
        off += lv.size
    assert off == got.size
    assert attrs["oiio:SHA-1"][1].decode() == kv["sha1"]
    assert attrs["textureformat"] == ("string", b"Plain Texture")
    l0 = exr_mini.read_openexr_level0(fexr, 4)
    if l0 is not None:
        assert np.array_equal(l0, src)


@pytest.mark.gpu
def test_cpp_make_texture_writes_uint8_tiff(iba_cli, tmp_path):
    """The same for a uint8 picture: the file holds convert_type<float,uint8>
    of the float levels (fmath.h:753-764), Predictor 2, SampleFormat uint."""
    import tiff_mini
    src = synth.image(200, 330, 3, np.uint8, seed=96)
    fin, fout, ftif = (str(tmp_path / n) for n in ("in.raw", "out.raw", "out.tif"))
    src.tofile(fin)
    out = subprocess.check_output([iba_cli, "maketx", fin, "200", "330", "3", "2", "lanczos3",
                                   "0", fout, ftif], text=True)
    kv = dict(t.split("=") for t in out.split())
    dirs = tiff_mini.read(ftif)
    assert len(dirs) == int(kv["levels"])
    got = np.fromfile(fout, np.float32)
    off = 0
    for lv, d in enumerate(dirs):
        n = d["pixels"].size
        s = got[off:off + n] * np.float32(255) + np.float32(0.5)
        want = np.clip(s, np.float32(0), np.float32(255)).astype(np.uint8)
        assert d["pixels"].dtype == np.uint8 and np.array_equal(d["pixels"].ravel(), want), lv
        off += n
        assert d["tags"][317] == [2] and d["tags"][339] == [1, 1, 1] and d["tags"][258] == [8] * 3
    assert np.array_equal(dirs[0]["pixels"], src)
    theirs = tiff_mini.read_libtiff(ftif, 3)
    if theirs is not None:
        assert all(np.array_equal(a, d["pixels"]) for a, d in zip(theirs, dirs))


@pytest.mark.gpu
def test_cpp_make_texture_writes_tiff(iba_cli, tmp_path):
    """make_texture(..., outputfilename) of the C++ mirror: the file's
    directories are the levels it returned (float, Predictor 3, 64x64 deflate
    tiles), the ImageDescription carries the SHA-1 it reports."""
    import tiff_mini
    src = synth.image(200, 330, 3, np.float32, seed=95)
    fin, fout, ftif = (str(tmp_path / n) for n in ("in.raw", "out.raw", "out.tif"))
    src.tofile(fin)
    out = subprocess.check_output([iba_cli, "maketx", fin, "200", "330", "3", "11", "lanczos3",
                                   "0", fout, ftif], text=True)
    kv = dict(t.split("=") for t in out.split())
    dirs = tiff_mini.read(ftif)
    assert len(dirs) == int(kv["levels"])
    got = np.fromfile(fout, np.float32)
    off = 0
    for d in dirs:
        n = d["pixels"].size
        assert np.array_equal(d["pixels"].ravel(), got[off:off + n])
        off += n
        assert d["tags"][317] == [3] and d["tags"][322] == [64] and d["tags"][259] == [8]
    assert off == got.size
    assert ("oiio:SHA-1=" + kv["sha1"]) in dirs[0]["tags"][270]
    assert "oiio:AverageColor=" in dirs[0]["tags"][270]
    theirs = tiff_mini.read_libtiff(ftif, 3)
    if theirs is not None:
        assert all(np.array_equal(a, d["pixels"]) for a, d in zip(theirs, dirs))


@pytest.mark.gpu
def test_cpp_oiiotool_resize(iba_cli, oracle, tmp_path):
    src = synth.image(180, 120, 3, np.float32, seed=94)
    fin, fout = str(tmp_path / "in.raw"), str(tmp_path / "out.raw")
    src.tofile(fin)
    subprocess.check_call([iba_cli, "otresize", fin, "180", "120", "3", "11", "90", "60",
                           "lanczos3", "0", fout])
    got = np.fromfile(fout, np.float32).reshape(60, 90, 3)
    ref = np.zeros_like(got)
    oracle.resize(oracle.Img(ref), oracle.Img(src), "lanczos3")
    assert np.abs(got - ref).max() <= 1e-5
