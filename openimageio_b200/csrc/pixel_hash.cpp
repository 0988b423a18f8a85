// pixel_hash.cpp -- ImageBufAlgo::computePixelHashSHA1
// (src/libOpenImageIO/imagebufalgo_compare.cpp:817-892), the hash maketx stores
// as oiio:SHA-1 (maketexture.cpp:1909-1934: top level, 256-row blocks, the
// filter name etc. appended as `extrainfo`).
//
// SHA-1 is one serial stream per block: 64 blocks for a 16K^2 texture, far too
// few for a GPU.  The top level maketx hashes is the image the HOST hands to
// the chain, so the hash runs on host threads (one block per task, as the
// reference's parallel_for_chunked does) while the device uploads the image
// and builds the pyramid -- it costs no wall time.  A device image is brought
// down block by block.
#include "api_internal.h"

#include <atomic>
#include <string>
#include <thread>
#include <vector>

using namespace b2r;
using namespace b2r::api;

namespace {

// FIPS 180-1.
struct Sha1 {
    uint32_t h[5] = { 0x67452301u, 0xEFCDAB89u, 0x98BADCFEu, 0x10325476u, 0xC3D2E1F0u };
    uint64_t nbytes = 0;
    unsigned char buf[64];
    size_t fill = 0;

    static uint32_t rol(uint32_t v, int s) { return (v << s) | (v >> (32 - s)); }
    void block(const unsigned char* p)
    {
        uint32_t w[80];
        for (int i = 0; i < 16; ++i)
            w[i] = (uint32_t(p[4 * i]) << 24) | (uint32_t(p[4 * i + 1]) << 16)
                   | (uint32_t(p[4 * i + 2]) << 8) | uint32_t(p[4 * i + 3]);
        for (int i = 16; i < 80; ++i)
            w[i] = rol(w[i - 3] ^ w[i - 8] ^ w[i - 14] ^ w[i - 16], 1);
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4];
        for (int i = 0; i < 80; ++i) {
            uint32_t f, k;
            if (i < 20) {
                f = (b & c) | (~b & d);
                k = 0x5A827999u;
            } else if (i < 40) {
                f = b ^ c ^ d;
                k = 0x6ED9EBA1u;
            } else if (i < 60) {
                f = (b & c) | (b & d) | (c & d);
                k = 0x8F1BBCDCu;
            } else {
                f = b ^ c ^ d;
                k = 0xCA62C1D6u;
            }
            uint32_t t = rol(a, 5) + f + e + k + w[i];
            e = d;
            d = c;
            c = rol(b, 30);
            b = a;
            a = t;
        }
        h[0] += a;
        h[1] += b;
        h[2] += c;
        h[3] += d;
        h[4] += e;
    }
    void append(const void* data, size_t n)
    {
        const unsigned char* p = (const unsigned char*)data;
        nbytes += n;
        if (fill) {
            size_t take = std::min(n, 64 - fill);
            memcpy(buf + fill, p, take);
            fill += take;
            p += take;
            n -= take;
            if (fill == 64) {
                block(buf);
                fill = 0;
            }
        }
        for (; n >= 64; p += 64, n -= 64)
            block(p);
        if (n) {
            memcpy(buf, p, n);
            fill = n;
        }
    }
    // upper-case hex, no separators: CSHA1::REPORT_HEX_SHORT (SHA1.cpp:264-279)
    std::string digest()
    {
        const uint64_t bits = nbytes * 8;
        unsigned char pad[72] = { 0x80 };
        size_t padlen = (fill < 56) ? 56 - fill : 120 - fill;
        unsigned char len[8];
        for (int i = 0; i < 8; ++i)
            len[i] = (unsigned char)(bits >> (56 - 8 * i));
        append(pad, padlen);
        append(len, 8);
        static const char* hex = "0123456789ABCDEF";
        std::string s(40, '0');
        for (int i = 0; i < 20; ++i) {
            unsigned char b = (unsigned char)(h[i / 4] >> (24 - 8 * (i % 4)));
            s[2 * i]     = hex[b >> 4];
            s[2 * i + 1] = hex[b & 15];
        }
        return s;
    }
};

// simplePixelHashSHA1 (:817-853) over rows [y0, y1) of the ROI columns; rows
// are hashed in the image's own pixel format, scanline by scanline
std::string
hash_rows(const unsigned char* base, long long ystride, size_t row_bytes, int y0, int y1,
          const char* extra, size_t extralen)
{
    Sha1 s;
    if ((long long)row_bytes == ystride)
        s.append(base + (long long)y0 * ystride, row_bytes * size_t(y1 - y0));
    else
        for (int y = y0; y < y1; ++y)
            s.append(base + (long long)y * ystride, row_bytes);
    if (extralen)
        s.append(extra, extralen);
    return s.digest();
}

}  // namespace

extern "C" int
b2r_pixel_hash_sha1(const b2r_image* src, const char* extrainfo, const b2r_roi* roi_in,
                    int blocksize, int nthreads, char* digest41, char* err, size_t errlen)
{
    int rc = check_image(src, true, err, errlen);
    if (rc)
        return rc;
    if (!digest41) {
        set_err(err, errlen, "pixel_hash: null digest buffer");
        return B2R_ERR_INVALID;
    }
    const int es = basetype_size(src->basetype);
    if (src->xstride != (int64_t)src->nchannels * es) {
        set_err(err, errlen, "pixel_hash: pixels must be contiguous in x");
        return B2R_ERR_UNSUPPORTED;
    }
    b2r_roi roi = roi_in ? *roi_in
                         : b2r_roi { src->x, src->x + src->width, src->y, src->y + src->height,
                                     0, src->nchannels };
    roi.xbegin = std::max(roi.xbegin, src->x);
    roi.xend   = std::min(roi.xend, src->x + src->width);
    roi.ybegin = std::max(roi.ybegin, src->y);
    roi.yend   = std::min(roi.yend, src->y + src->height);
    const int rw = roi.xend - roi.xbegin, rh = roi.yend - roi.ybegin;
    if (rw <= 0 || rh <= 0) {
        set_err(err, errlen, "pixel_hash: empty region");
        return B2R_ERR_INVALID;
    }
    const std::string extra = extrainfo ? extrainfo : "";
    const size_t row_bytes  = size_t(rw) * src->xstride;

    // a device image comes down first (block hashes need host bytes)
    std::vector<unsigned char> staged;
    const unsigned char* base = (const unsigned char*)src->pixels
                                + (long long)(roi.ybegin - src->y) * src->ystride
                                + (long long)(roi.xbegin - src->x) * src->xstride;
    long long ystride = src->ystride;
    int dev           = 0;
    if (memspace_of(src, &dev) == B2R_MEM_DEVICE) {
        DeviceGuard guard(dev);
        staged.resize(row_bytes * size_t(rh));
        CUDA_TRY(copy_d2h_2d(staged.data(), row_bytes, base, (size_t)ystride, row_bytes, rh,
                             nullptr));
        CUDA_TRY(cudaStreamSynchronize(nullptr));
        base    = staged.data();
        ystride = (long long)row_bytes;
    }
    std::string out;
    if (blocksize <= 0 || blocksize >= rh) {
        out = hash_rows(base, ystride, row_bytes, 0, rh, extra.data(), extra.size());
    } else {
        const int nblocks = (rh + blocksize - 1) / blocksize;
        std::vector<std::string> results(nblocks);
        int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
        nt     = std::max(1, std::min(nt, nblocks));
        std::atomic<int> next(0);
        auto work = [&]() {
            for (;;) {
                const int b = next.fetch_add(1);
                if (b >= nblocks)
                    break;
                results[b] = hash_rows(base, ystride, row_bytes, b * blocksize,
                                       std::min(rh, (b + 1) * blocksize), nullptr, 0);
            }
        };
        std::vector<std::thread> th;
        try {
            for (int k = 1; k < nt; ++k)
                th.emplace_back(work);
        } catch (...) {
        }
        work();
        for (auto& x : th)
            x.join();
        // the block digests (as text), then the extra info (:881-884)
        Sha1 s;
        for (const auto& r : results)
            s.append(r.data(), r.size());
        s.append(extra.data(), extra.size());
        out = s.digest();
    }
    memcpy(digest41, out.c_str(), 41);
    return B2R_OK;
}
