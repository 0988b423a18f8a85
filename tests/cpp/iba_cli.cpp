// iba_cli -- tiny driver over include/b2r_imagebufalgo.h, written the way a
// program against OpenImageIO's C++ API would be:
//   ImageBuf src(spec, pixels); ImageBuf dst(dstspec);
//   ImageBufAlgo::resize(dst, src, {{"filtername", f}}, roi);
// usage: iba_cli errors
//        iba_cli resize in.raw w h nch basetype dw dh filter out.raw
//        iba_cli rotate | fitexact | sharpen  (same arguments)
//        iba_cli hicomp in.raw w h nch basetype dw dh filter out.raw
//            (oiiotool --resize:highlightcomp=1 spelled with three IBA calls)
#include "b2r_imagebufalgo.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

using namespace b2r;

static int
run_errors()
{
    ImageBuf src(ImageSpec(16, 16, 3, TypeDesc::FLOAT));
    ImageBuf bad = ImageBufAlgo::resize(src, { { "filtername", "catrom" } },
                                        ROI(0, 8, 0, 8, 0, 1, 0, 3));
    printf("1:%s\n", bad.geterror().c_str());
    ImageBuf empty;
    ImageBuf r2 = ImageBufAlgo::resize(empty, { { "filtername", "box" } });
    printf("2:%s\n", r2.geterror().c_str());
    ImageSpec vol(8, 8, 1, TypeDesc::FLOAT);
    vol.depth = 2;
    ImageBuf v(vol);
    ImageBuf r3 = ImageBufAlgo::resize(v, {}, ROI(0, 4, 0, 4, 0, 1, 0, 1));
    printf("3:%s\n", r3.geterror().c_str());
    const float ident[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
    ImageBuf r4 = ImageBufAlgo::warp(src, ident, { { "filtername", "lanczos" } });
    printf("4:%s\n", r4.geterror().c_str());
    // unrecognised options are ignored (xform.cpp:880)
    ImageBuf r5 = ImageBufAlgo::resize(src, { { "bogus", 1 }, { "filtername", "nope" } },
                                       ROI(0, 8, 0, 8, 0, 1, 0, 3));
    printf("5:%s\n", r5.geterror().c_str());
    printf("devices:%d\n", b2r_device_count());
    return 0;
}

// iba_cli maketx in.raw w h nch basetype filter hicomp(0|1) out.raw [out.tif]
//   ImageBufAlgo::make_texture: every level appended to out.raw (float32),
//   "levels=N sha1=... avg0=..." on stdout
static int
run_maketx(char** argv)
{
    int w = atoi(argv[3]), h = atoi(argv[4]), c = atoi(argv[5]), bt = atoi(argv[6]);
    ImageSpec sspec(w, h, c, TypeDesc(bt));
    std::vector<unsigned char> px(size_t(w) * h * c * sspec.format.size());
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(px.data(), 1, px.size(), f) != px.size())
        return 3;
    fclose(f);
    ImageBuf src(sspec, px.data(), B2R_MEM_HOST);
    Texture T = ImageBufAlgo::make_texture(ImageBufAlgo::MakeTxTexture, src,
                                           { { "maketx:filtername", argv[7] },
                                             { "maketx:highlightcomp", atoi(argv[8]) } },
                                           argv[10] ? argv[10] : "");
    if (!T.ok()) {
        fprintf(stderr, "%s\n", T.error.c_str());
        return 4;
    }
    f = fopen(argv[9], "wb");
    for (auto& L : T.levels)
        fwrite(L.localpixels(), 4, size_t(L.spec().width) * L.spec().height * c, f);
    fclose(f);
    printf("levels=%d sha1=%s avg0=%.9g\n", int(T.levels.size()), T.sha1.c_str(),
           T.average_color[0]);
    return 0;
}

// iba_cli otresize in.raw w h nch basetype dw dh filter hicomp out.raw
//   oiiotool --resize[:highlightcomp=1] WxH
static int
run_otresize(char** argv)
{
    int w = atoi(argv[3]), h = atoi(argv[4]), c = atoi(argv[5]), bt = atoi(argv[6]);
    int dw = atoi(argv[7]), dh = atoi(argv[8]);
    ImageSpec sspec(w, h, c, TypeDesc(bt));
    std::vector<unsigned char> px(size_t(w) * h * c * sspec.format.size());
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(px.data(), 1, px.size(), f) != px.size())
        return 3;
    fclose(f);
    ImageBuf src(sspec, px.data(), B2R_MEM_HOST);
    ImageBuf dst = oiiotool_resize(src, dw, dh, argv[9], nullptr, nullptr, atoi(argv[10]) != 0);
    if (dst.has_error()) {
        fprintf(stderr, "%s\n", dst.geterror().c_str());
        return 4;
    }
    f = fopen(argv[11], "wb");
    fwrite(dst.localpixels(), sspec.format.size(), size_t(dw) * dh * c, f);
    fclose(f);
    return 0;
}

int
main(int argc, char** argv)
{
    if (argc >= 2 && std::string(argv[1]) == "errors")
        return run_errors();
    if ((argc == 10 || argc == 11) && std::string(argv[1]) == "maketx")
        return run_maketx(argv);
    if (argc == 12 && std::string(argv[1]) == "otresize")
        return run_otresize(argv);
    const bool hicomp = argc >= 2 && std::string(argv[1]) == "hicomp";
    const bool rotate = argc >= 2 && std::string(argv[1]) == "rotate";
    const bool fitx   = argc >= 2 && std::string(argv[1]) == "fitexact";
    const bool sharp  = argc >= 2 && std::string(argv[1]) == "sharpen";
    if (argc != 11
        || (std::string(argv[1]) != "resize" && !hicomp && !rotate && !fitx && !sharp)) {
        fprintf(stderr, "usage: see source\n");
        return 2;
    }
    int w = atoi(argv[3]), h = atoi(argv[4]), c = atoi(argv[5]), bt = atoi(argv[6]);
    int dw = atoi(argv[7]), dh = atoi(argv[8]);
    ImageSpec sspec(w, h, c, TypeDesc(bt));
    std::vector<unsigned char> px(size_t(w) * h * c * sspec.format.size());
    FILE* f = fopen(argv[2], "rb");
    if (!f || fread(px.data(), 1, px.size(), f) != px.size())
        return 3;
    fclose(f);
    ImageBuf src(sspec, px.data(), B2R_MEM_HOST);
    ImageBuf tmpimg;
    const ImageBuf* in = &src;
    if (hicomp) {  // oiiotool.cpp:4644-4650
        if (!ImageBufAlgo::rangecompress(tmpimg, src))
            return 5;
        in = &tmpimg;
    }
    ImageBuf dst;
    if (rotate) {  // 30 degrees about the display-window centre, dst = src's window
        dst = ImageBufAlgo::rotate(src, 30.0f * float(M_PI / 180.0), argv[9]);
        dw = w, dh = h;
    } else if (sharp) {
        // fixNonFinite -> unsharp_mask -> convolve with a 3x3 box, plus the
        // statistics of the result on stderr: the widened C++ surface
        ImageBuf fixed, sharpened;
        int nfixed = 0;
        if (!ImageBufAlgo::fixNonFinite(fixed, src, ImageBufAlgo::NONFINITE_BOX3, &nfixed)
            || !ImageBufAlgo::unsharp_mask(sharpened, fixed, "gaussian", 3.0f, 0.7f, 0.0f)) {
            fprintf(stderr, "%s%s\n", fixed.geterror().c_str(), sharpened.geterror().c_str());
            return 7;
        }
        ImageBuf K = ImageBufAlgo::make_kernel(argv[9], 3.0f, 3.0f);
        if (!ImageBufAlgo::convolve(dst, sharpened, K))
            return 8;
        bool mono = true;
        ImageBufAlgo::PixelStats st = ImageBufAlgo::computePixelStats(dst, {}, 0, &mono);
        fprintf(stderr, "fixed=%d avg0=%.9g finite0=%lld mono=%d\n", nfixed, st.avg[0],
                st.finitecount[0], int(mono));
        dw = w, dh = h;
    } else if (fitx) {
        dst = ImageBufAlgo::fit(src, { { "exact", 1 }, { "filtername", argv[9] } },
                                ROI(0, dw, 0, dh, 0, 1, 0, c));
    } else {
        dst = ImageBufAlgo::resize(*in, { { "filtername", argv[9] } },
                                   ROI(0, dw, 0, dh, 0, 1, 0, c));
    }
    if (hicomp && !dst.has_error() && !ImageBufAlgo::rangeexpand(dst, dst))
        return 6;
    if (dst.has_error()) {
        fprintf(stderr, "%s\n", dst.geterror().c_str());
        return 4;
    }
    f = fopen(argv[10], "wb");
    fwrite(dst.localpixels(), 1, size_t(dw) * dh * c * sspec.format.size(), f);
    fclose(f);
    return 0;
}
