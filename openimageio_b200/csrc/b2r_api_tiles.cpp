// b2r_api_tiles.cpp -- the output stage behind the MIP chain:
// what write_mipmap does per level with out->open(AppendMIPLevel) and
// small->write(out) (src/libOpenImageIO/maketexture.cpp:957-974) up to the
// point where compressed tiles are handed to the file: every level is re-laid
// into tiles on the device (optionally float -> half, maketx -d half), comes
// down over a copy stream into pinned staging slots, and a pool of host
// threads deflates the tiles (zlib, the "zip" compression maketx sets by
// default, maketexture.cpp:1605) while the next chunk is still on its way:
//
//   device : relayout(chunk c+2) | D2H(chunk c+1)      (stream-ordered, 3 slots)
//   host   :                                deflate(chunk c)   (nthreads workers)
//
// The result is an in-memory tile set (per level: a grid of compressed tiles).
// b2r_tileset_write_tiff puts the TIFF container around it (one directory per
// MIP level, the layout of a maketx .tx file: src/tiff.imageio/tiffoutput.cpp);
// with predictor = 3 the re-layout kernel has already applied the floating-point
// predictor the reference's TIFF writer uses (tiffoutput.cpp:683-690).  OpenEXR
// containers stay with the reference's plugin.
#include "api_internal.h"

#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include <atomic>
#include <mutex>
#include <chrono>
#include <thread>

using namespace b2r;
using namespace b2r::api;

struct b2r_tileset {
    struct Level {
        int w = 0, h = 0, ntx = 0, nty = 0;
        std::vector<std::vector<unsigned char>> tiles;  // nty * ntx, raster order
    };
    int nch = 0, tile_w = 0, tile_h = 0, out_basetype = 0, compression = 0, predictor = 0;
    int first_level = 0;
    std::vector<Level> levels;
    long long raw_bytes = 0, stored_bytes = 0;
    float total_ms = 0, device_ms = 0, deflate_busy_ms = 0, setup_ms = 0, wait_ms = 0;
};

namespace {

int
tile_elem_size(int obt)
{
    return obt == B2R_UINT8 ? 1 : (obt == B2R_FLOAT ? 4 : 2);
}

// out_basetype / predictor / tile-row rules shared by encode_tiles and tile_relayout:
// predictor 3 (floating point) goes with float / half tiles, 2 (horizontal) with
// uint8 / uint16 ones -- the pairing the reference's TIFF writer uses
// (tiffoutput.cpp:683-698)
int
check_tile_format(const char* who, int obt, int predictor, int tile_w, int nch, int* pred,
                  char* err, size_t errlen)
{
    const bool isint = obt == B2R_UINT8 || obt == B2R_UINT16;
    if (!isint && obt != B2R_FLOAT && obt != B2R_HALF) {
        set_err(err, errlen, "%s: tiles are float32, half, uint16 or uint8", who);
        return B2R_ERR_TYPES;
    }
    if (predictor != 0 && predictor != 1 && predictor != 2 && predictor != 3) {
        set_err(err, errlen,
                "%s: predictor must be 0, 1 (none), 2 (horizontal) or 3 (floating point)", who);
        return B2R_ERR_INVALID;
    }
    if ((predictor == 2 && !isint) || (predictor == 3 && isint)) {
        set_err(err, errlen,
                "%s: predictor 2 goes with uint8 / uint16 tiles, predictor 3 with float / half",
                who);
        return B2R_ERR_INVALID;
    }
    *pred = predictor >= 2 ? predictor : 0;
    if ((*pred || isint) && (size_t(tile_w) * nch) % 4 != 0) {
        set_err(err, errlen,
                "%s: predictor and integer tiles need tile rows of a multiple of 4 samples", who);
        return B2R_ERR_INVALID;
    }
    return B2R_OK;
}

void
launch_tiles(const float* px, int w, int h, int nch, int tile_w, int tile_h, int ty0, int ty1,
             int obt, int pred, void* out, void* stream)
{
    if (obt == B2R_UINT8 || obt == B2R_UINT16)
        launch_tile_relayout_int(px, w, h, nch, tile_w, tile_h, ty0, ty1, obt == B2R_UINT16,
                                 pred == 2, out, stream);
    else if (pred == 3)
        launch_tile_relayout_fp_predictor(px, w, h, nch, tile_w, tile_h, ty0, ty1,
                                          obt == B2R_HALF, out, stream);
    else
        launch_tile_relayout(px, w, h, nch, tile_w, tile_h, ty0, ty1, obt == B2R_HALF, out,
                             stream);
}

// The three device + three pinned host staging slots, kept per device between
// calls (grow-only): pinning and unpinning 3 x 64 MB cost 90-300 ms per call,
// a third of an 8K^2 chain's whole output stage.  One encode at a time per
// device holds them.
constexpr int NSLOT = 3;
struct SlotCache {
    std::mutex mu;
    void* d[NSLOT] = { nullptr, nullptr, nullptr };
    void* h[NSLOT] = { nullptr, nullptr, nullptr };
    size_t bytes   = 0;
    bool reserve(size_t need)
    {
        if (need <= bytes)
            return true;
        for (int i = 0; i < NSLOT; ++i) {
            if (d[i])
                cudaFree(d[i]);
            if (h[i])
                cudaFreeHost(h[i]);
            d[i] = h[i] = nullptr;
        }
        bytes = 0;
        for (int i = 0; i < NSLOT; ++i) {
            if (cudaMalloc(&d[i], need) != cudaSuccess
                || cudaHostAlloc(&h[i], need, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return false;
            }
        }
        bytes = need;
        return true;
    }
};
SlotCache&
slot_cache(int device)
{
    static SlotCache caches[64];
    return caches[device & 63];
}

}  // namespace

namespace b2r {
namespace api {
// b2r_trim_caches: the tile encoder's staging slots (device + pinned host)
void
trim_tile_slots()
{
    int prev = 0;
    cudaGetDevice(&prev);
    for (int d = 0; d < 64; ++d) {
        SlotCache& c = slot_cache(d);
        std::lock_guard<std::mutex> lock(c.mu);
        if (!c.bytes)
            continue;
        cudaSetDevice(d);
        for (int i = 0; i < NSLOT; ++i) {
            if (c.d[i])
                cudaFree(c.d[i]);
            if (c.h[i])
                cudaFreeHost(c.h[i]);
            c.d[i] = c.h[i] = nullptr;
        }
        c.bytes = 0;
    }
    cudaSetDevice(prev);
}
}  // namespace api
}  // namespace b2r

extern "C" int
b2r_mipchain_encode_tiles(const b2r_mipchain* chain, const b2r_tile_options* to,
                          b2r_tileset** out, char* err, size_t errlen)
{
    if (!chain || !out) {
        set_err(err, errlen, "encode_tiles: null argument");
        return B2R_ERR_INVALID;
    }
    *out = nullptr;
    const int tile_w = (to && to->tile_width > 0) ? to->tile_width : 64;
    const int tile_h = (to && to->tile_height > 0) ? to->tile_height : 64;
    const int obt    = (to && to->out_basetype) ? to->out_basetype : B2R_FLOAT;
    const int comp   = to ? to->compression : 1;
    const int zlevel = (to && to->zlevel > 0) ? std::min(9, (int)to->zlevel) : 1;
    const int first  = (to && to->first_level > 0) ? to->first_level : 0;
    int pred = 0;
    {
        const int rc = check_tile_format("encode_tiles", obt, to ? to->predictor : 0, tile_w,
                                         chain->nch, &pred, err, errlen);
        if (rc)
            return rc;
    }
    int nthreads = (to && to->nthreads > 0) ? to->nthreads
                                            : (int)std::thread::hardware_concurrency();
    nthreads     = std::max(1, std::min(nthreads, 256));
    const int nch = chain->nch;
    const int oes = tile_elem_size(obt);
    const size_t tile_bytes = size_t(tile_w) * tile_h * nch * oes;

    auto t_begin = std::chrono::steady_clock::now();
    DeviceGuard guard(chain->device);
    std::unique_ptr<b2r_tileset> ts(new b2r_tileset);
    ts->nch          = nch;
    ts->tile_w       = tile_w;
    ts->tile_h       = tile_h;
    ts->out_basetype = obt;
    ts->compression  = comp;
    ts->predictor    = pred;
    ts->first_level  = first;

    // chunks of whole tile rows, at most ~64 MB of tiles each
    struct Chunk {
        int level, ty0, ty1;
    };
    std::vector<Chunk> chunks;
    for (int lv = 0; lv < (int)chain->levels.size(); ++lv) {
        b2r_tileset::Level L;
        L.w   = chain->levels[lv].w;
        L.h   = chain->levels[lv].h;
        L.ntx = (L.w + tile_w - 1) / tile_w;
        L.nty = (L.h + tile_h - 1) / tile_h;
        if (lv >= first)
            L.tiles.resize(size_t(L.ntx) * L.nty);
        ts->levels.push_back(std::move(L));
        if (lv < first)
            continue;
        const auto& T   = ts->levels.back();
        const size_t rowb = size_t(T.ntx) * tile_bytes;
        const int per     = (int)std::max<size_t>(1, (size_t(64) << 20) / rowb);
        for (int ty = 0; ty < T.nty; ty += per)
            chunks.push_back({ lv, ty, std::min(T.nty, ty + per) });
    }
    size_t slot_bytes = 0;
    for (const Chunk& c : chunks)
        slot_bytes = std::max(slot_bytes,
                              size_t(c.ty1 - c.ty0) * ts->levels[c.level].ntx * tile_bytes);
    SlotCache& cache = slot_cache(chain->device);
    std::lock_guard<std::mutex> cache_lock(cache.mu);
    void** dslot = cache.d;
    void** hslot = cache.h;
    cudaEvent_t landed[NSLOT] = { nullptr, nullptr, nullptr };
    cudaStream_t sk = nullptr, sc = nullptr;
    cudaEvent_t relaid = nullptr, ev0 = nullptr, ev1 = nullptr;
    auto cleanup = [&]() {
        for (int i = 0; i < NSLOT; ++i)
            if (landed[i])
                cudaEventDestroy(landed[i]);
        if (relaid)
            cudaEventDestroy(relaid);
        if (ev0)
            cudaEventDestroy(ev0);
        if (ev1)
            cudaEventDestroy(ev1);
        if (sk)
            cudaStreamDestroy(sk);
        if (sc)
            cudaStreamDestroy(sc);
    };
    if (slot_bytes && !cache.reserve(slot_bytes)) {
        set_err(err, errlen, "encode_tiles: out of memory for the staging slots");
        return B2R_ERR_CUDA;
    }
    for (int i = 0; i < NSLOT; ++i)
        cudaEventCreateWithFlags(&landed[i], cudaEventDisableTiming);
    cudaStreamCreateWithFlags(&sk, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&sc, cudaStreamNonBlocking);
    cudaEventCreateWithFlags(&relaid, cudaEventDisableTiming);
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    cudaEventRecord(ev0, sk);
    ts->setup_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now()
                                                            - t_begin)
                       .count();

    // device side of chunk i: relayout on `sk`, download on `sc`
    auto issue = [&](int i) {
        const Chunk& c = chunks[i];
        const auto& CL = chain->levels[c.level];
        const auto& TL = ts->levels[c.level];
        const int s    = i % NSLOT;
        launch_tiles((const float*)CL.px, CL.w, CL.h, nch, tile_w, tile_h, c.ty0, c.ty1, obt, pred,
                     dslot[s], sk);
        cudaEventRecord(relaid, sk);
        cudaStreamWaitEvent(sc, relaid, 0);
        cudaMemcpyAsync(hslot[s], dslot[s], size_t(c.ty1 - c.ty0) * TL.ntx * tile_bytes,
                        cudaMemcpyDeviceToHost, sc);
        cudaEventRecord(landed[s], sc);
        // the next relayout into this device slot must not overtake the copy
        cudaStreamWaitEvent(sk, landed[s], 0);
    };
    const int nchunks = (int)chunks.size();
    for (int i = 0; i < std::min(nchunks, NSLOT - 1); ++i)
        issue(i);
    std::atomic<int> failed(0);
    double busy_ms = 0, wait_ms = 0;
    for (int i = 0; i < nchunks; ++i) {
        const Chunk& c = chunks[i];
        auto& TL       = ts->levels[c.level];
        const int s    = i % NSLOT;
        if (i + NSLOT - 1 < nchunks)
            issue(i + NSLOT - 1);  // its slot was deflated in the previous iteration
        auto tw = std::chrono::steady_clock::now();
        const cudaError_t landed_rc = cudaEventSynchronize(landed[s]);
        wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw)
                       .count();
        if (landed_rc != cudaSuccess) {
            cleanup();
            set_err(err, errlen, "encode_tiles: CUDA error: %s",
                    cudaGetErrorString(cudaGetLastError()));
            return B2R_ERR_CUDA;
        }
        if (i == nchunks - 1)
            cudaEventRecord(ev1, sc);
        auto t0 = std::chrono::steady_clock::now();
        const int ntiles = (c.ty1 - c.ty0) * TL.ntx;
        const unsigned char* base = (const unsigned char*)hslot[s];
        std::atomic<int> next(0);
        auto work = [&]() {
            std::vector<unsigned char> tmp(comp ? compressBound((uLong)tile_bytes) : 0);
            for (;;) {
                const int t = next.fetch_add(1);
                if (t >= ntiles)
                    break;
                auto& dst = TL.tiles[size_t(c.ty0) * TL.ntx + t];
                const unsigned char* src = base + size_t(t) * tile_bytes;
                if (comp) {
                    uLongf n = (uLongf)tmp.size();
                    if (compress2(tmp.data(), &n, src, (uLong)tile_bytes, zlevel) != Z_OK) {
                        failed = 1;
                        break;
                    }
                    dst.assign(tmp.data(), tmp.data() + n);
                } else {
                    dst.assign(src, src + tile_bytes);
                }
            }
        };
        const int nt = std::min(nthreads, ntiles);
        std::vector<std::thread> th;
        try {
            for (int k = 1; k < nt; ++k)
                th.emplace_back(work);
        } catch (...) {
        }
        work();
        for (auto& x : th)
            x.join();
        busy_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0)
                       .count();
    }
    if (nchunks == 0)
        cudaEventRecord(ev1, sc);
    cudaStreamSynchronize(sk);
    cudaStreamSynchronize(sc);
    cudaEventElapsedTime(&ts->device_ms, ev0, ev1);
    cleanup();
    if (failed) {
        set_err(err, errlen, "encode_tiles: zlib failed");
        return B2R_ERR_INVALID;
    }
    for (const auto& L : ts->levels)
        for (const auto& t : L.tiles) {
            ts->raw_bytes += (long long)tile_bytes;
            ts->stored_bytes += (long long)t.size();
        }
    ts->deflate_busy_ms = (float)busy_ms;
    ts->wait_ms         = (float)wait_ms;
    ts->total_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now()
                                                            - t_begin)
                       .count();
    *out = ts.release();
    return B2R_OK;
}

extern "C" int
b2r_tileset_nlevels(const b2r_tileset* ts)
{
    return ts ? (int)ts->levels.size() : 0;
}

extern "C" int
b2r_tileset_level(const b2r_tileset* ts, int level, int* width, int* height, int* ntiles_x,
                  int* ntiles_y)
{
    if (!ts || level < 0 || level >= (int)ts->levels.size())
        return B2R_ERR_INVALID;
    const auto& L = ts->levels[level];
    if (width)
        *width = L.w;
    if (height)
        *height = L.h;
    if (ntiles_x)
        *ntiles_x = L.ntx;
    if (ntiles_y)
        *ntiles_y = L.nty;
    return B2R_OK;
}

extern "C" int
b2r_tileset_tile(const b2r_tileset* ts, int level, int tx, int ty, const void** data,
                 size_t* bytes)
{
    if (!ts || level < 0 || level >= (int)ts->levels.size())
        return B2R_ERR_INVALID;
    const auto& L = ts->levels[level];
    if (tx < 0 || ty < 0 || tx >= L.ntx || ty >= L.nty || L.tiles.empty())
        return B2R_ERR_INVALID;
    const auto& t = L.tiles[size_t(ty) * L.ntx + tx];
    if (data)
        *data = t.data();
    if (bytes)
        *bytes = t.size();
    return B2R_OK;
}

extern "C" double
b2r_tileset_stat(const b2r_tileset* ts, int what)
{
    if (!ts)
        return 0.0;
    switch (what) {
    case 0: return ts->total_ms;         // host wall clock of the encode call
    case 1: return ts->device_ms;        // relayout + downloads, device events
    case 2: return ts->deflate_busy_ms;  // host time spent in the deflate phases
    case 3: return (double)ts->raw_bytes;
    case 4: return (double)ts->stored_bytes;
    case 5: return ts->setup_ms;  // staging slots, streams
    case 6: return ts->wait_ms;   // host time blocked on the downloads
    }
    return 0.0;
}

extern "C" void
b2r_tileset_destroy(b2r_tileset* ts)
{
    delete ts;
}


extern "C" int
b2r_tile_relayout(const b2r_image* src, void* dst, size_t dst_bytes, const b2r_tile_options* to,
                  void* stream, char* err, size_t errlen)
{
    if (!src || !src->pixels || !dst) {
        set_err(err, errlen, "tile_relayout: null argument");
        return B2R_ERR_INVALID;
    }
    if (device_count_cached() <= 0) {
        set_err(err, errlen,
                "no CUDA device available: the B200 tile output path has no CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    const int nch    = src->nchannels;
    const int tile_w = (to && to->tile_width > 0) ? to->tile_width : 64;
    const int tile_h = (to && to->tile_height > 0) ? to->tile_height : 64;
    const int obt    = (to && to->out_basetype) ? to->out_basetype : B2R_FLOAT;
    int pred         = 0;
    if (src->memspace != B2R_MEM_DEVICE || src->basetype != B2R_FLOAT || nch < 1
        || src->xstride != (int64_t)nch * 4 || src->ystride != (int64_t)src->width * nch * 4) {
        set_err(err, errlen, "tile_relayout: the source is a contiguous float32 device image");
        return B2R_ERR_INVALID;
    }
    {
        const int rc = check_tile_format("tile_relayout", obt, to ? to->predictor : 0, tile_w, nch,
                                         &pred, err, errlen);
        if (rc)
            return rc;
    }
    const int ntx = (src->width + tile_w - 1) / tile_w, nty = (src->height + tile_h - 1) / tile_h;
    const size_t need = size_t(ntx) * nty * tile_w * tile_h * nch * tile_elem_size(obt);
    if (dst_bytes < need) {
        set_err(err, errlen, "tile_relayout: destination holds %zu bytes, the tiles need %zu",
                dst_bytes, need);
        return B2R_ERR_INVALID;
    }
    DeviceGuard guard(src->device);
    launch_tiles((const float*)src->pixels, src->width, src->height, nch, tile_w, tile_h, 0, nty,
                 obt, pred, dst, stream);
    const cudaError_t rc = cudaGetLastError();
    if (rc != cudaSuccess) {
        set_err(err, errlen, "tile_relayout: CUDA error: %s", cudaGetErrorString(rc));
        return B2R_ERR_CUDA;
    }
    return B2R_OK;
}

// ---------------------------------------------------------------------------
// TIFF container around the stored tiles
// ---------------------------------------------------------------------------
namespace {

struct TiffTag {
    uint16_t tag, type;  // type: 2 ASCII, 3 SHORT, 4 LONG, 16 LONG8
    uint64_t count;
    std::vector<unsigned char> data;
};

template<typename T>
void
put(std::vector<unsigned char>& v, T x)
{
    // the file is little-endian ("II"), like every host this library runs on
    const unsigned char* p = reinterpret_cast<const unsigned char*>(&x);
    v.insert(v.end(), p, p + sizeof(T));
}

TiffTag
tag_shorts(uint16_t tag, const std::vector<uint16_t>& vals)
{
    TiffTag t { tag, 3, vals.size(), {} };
    for (uint16_t x : vals)
        put(t.data, x);
    return t;
}

TiffTag
tag_long(uint16_t tag, uint32_t v)
{
    TiffTag t { tag, 4, 1, {} };
    put(t.data, v);
    return t;
}

TiffTag
tag_ascii(uint16_t tag, const std::string& s)
{
    TiffTag t { tag, 2, s.size() + 1, {} };
    t.data.assign(s.begin(), s.end());
    t.data.push_back(0);
    return t;
}

TiffTag
tag_offsets(uint16_t tag, const std::vector<uint64_t>& vals, bool big)
{
    TiffTag t { tag, uint16_t(big ? 16 : 4), vals.size(), {} };
    for (uint64_t x : vals) {
        if (big)
            put(t.data, x);
        else
            put(t.data, uint32_t(x));
    }
    return t;
}

struct File {
    FILE* f      = nullptr;
    uint64_t pos = 0;
    bool ok      = true;
    ~File()
    {
        if (f)
            fclose(f);
    }
    void write(const void* p, size_t n)
    {
        if (n && fwrite(p, 1, n, f) != n)
            ok = false;
        pos += n;
    }
    void align2()
    {
        if (pos & 1) {
            const char z = 0;
            write(&z, 1);
        }
    }
};

}  // namespace

extern "C" int
b2r_tileset_write_tiff(const b2r_tileset* ts, const char* path, const b2r_tiff_options* opt,
                       char* err, size_t errlen)
{
    if (!ts || !path) {
        set_err(err, errlen, "write_tiff: null argument");
        return B2R_ERR_INVALID;
    }
    if (ts->levels.empty() || ts->first_level != 0) {
        set_err(err, errlen, "write_tiff: the tile set must hold every level of the chain");
        return B2R_ERR_INVALID;
    }
    if (ts->tile_w % 16 || ts->tile_h % 16) {
        set_err(err, errlen, "write_tiff: TIFF tile sizes are multiples of 16 (%d x %d)",
                ts->tile_w, ts->tile_h);
        return B2R_ERR_INVALID;
    }
    if (ts->predictor >= 2 && !ts->compression) {
        set_err(err, errlen, "write_tiff: TIFF's predictor goes with compressed tiles only");
        return B2R_ERR_INVALID;
    }
    const int nch = ts->nch;
    const int bps     = 8 * tile_elem_size(ts->out_basetype);
    const bool isfloat = ts->out_basetype == B2R_HALF || ts->out_basetype == B2R_FLOAT;
    // classic TIFF holds 32-bit offsets; leave room for the directories
    uint64_t payload = 0, ntiles = 0;
    for (const auto& L : ts->levels)
        for (const auto& t : L.tiles) {
            payload += t.size() + 1;
            ++ntiles;
        }
    const bool big = (opt && opt->bigtiff)
                     || payload + ntiles * 16 + ts->levels.size() * 4096 + 65536 > 0xfff00000ull;
    const std::string software = (opt && opt->software) ? opt->software : "openimageio_b200";
    const std::string wrap     = (opt && opt->wrapmodes) ? opt->wrapmodes : "black,black";

    File out;
    out.f = fopen(path, "wb");
    if (!out.f) {
        set_err(err, errlen, "write_tiff: could not open \"%s\"", path);
        return B2R_ERR_INVALID;
    }
    std::vector<unsigned char> hdr;
    hdr.push_back('I');
    hdr.push_back('I');
    if (big) {
        put(hdr, uint16_t(43));
        put(hdr, uint16_t(8));
        put(hdr, uint16_t(0));
        put(hdr, uint64_t(0));
    } else {
        put(hdr, uint16_t(42));
        put(hdr, uint32_t(0));
    }
    out.write(hdr.data(), hdr.size());
    uint64_t link_pos = big ? 8 : 4;  // where the offset of the next directory goes
    const size_t inline_bytes = big ? 8 : 4;

    for (const auto& L : ts->levels) {
        std::vector<uint64_t> offs, cnts;
        offs.reserve(L.tiles.size());
        cnts.reserve(L.tiles.size());
        for (const auto& t : L.tiles) {
            out.align2();
            offs.push_back(out.pos);
            cnts.push_back(t.size());
            out.write(t.data(), t.size());
        }
        std::vector<TiffTag> tags;
        tags.push_back(tag_long(256, (uint32_t)L.w));
        tags.push_back(tag_long(257, (uint32_t)L.h));
        tags.push_back(tag_shorts(258, std::vector<uint16_t>(nch, (uint16_t)bps)));
        tags.push_back(tag_shorts(259, { uint16_t(ts->compression ? 8 : 1) }));
        tags.push_back(tag_shorts(262, { uint16_t(nch >= 3 ? 2 : 1) }));
        if (opt && opt->image_description)
            tags.push_back(tag_ascii(270, opt->image_description));
        tags.push_back(tag_shorts(277, { (uint16_t)nch }));
        tags.push_back(tag_shorts(284, { 1 }));
        tags.push_back(tag_ascii(305, software));
        if (ts->predictor >= 2)
            tags.push_back(tag_shorts(317, { (uint16_t)ts->predictor }));
        tags.push_back(tag_long(322, (uint32_t)ts->tile_w));
        tags.push_back(tag_long(323, (uint32_t)ts->tile_h));
        tags.push_back(tag_offsets(324, offs, big));
        tags.push_back(tag_offsets(325, cnts, big));
        // channels beyond the colour ones: alpha (associated, the reference's
        // default: tiffoutput.cpp ExtraSamples) then unspecified
        const int ncolor = nch >= 3 ? 3 : 1;
        if (nch > ncolor) {
            std::vector<uint16_t> extra(nch - ncolor, 0);
            extra[0] = 1;
            tags.push_back(tag_shorts(338, extra));
        }
        tags.push_back(tag_shorts(339, std::vector<uint16_t>(nch, uint16_t(isfloat ? 3 : 1))));
        tags.push_back(tag_ascii(33302, "Plain Texture"));
        tags.push_back(tag_ascii(33303, wrap));
        std::sort(tags.begin(), tags.end(),
                  [](const TiffTag& a, const TiffTag& b) { return a.tag < b.tag; });
        std::vector<uint64_t> placed(tags.size(), 0);
        for (size_t i = 0; i < tags.size(); ++i)
            if (tags[i].data.size() > inline_bytes) {
                out.align2();
                placed[i] = out.pos;
                out.write(tags[i].data.data(), tags[i].data.size());
            }
        out.align2();
        const uint64_t ifd = out.pos;
        std::vector<unsigned char> dir;
        if (big)
            put(dir, uint64_t(tags.size()));
        else
            put(dir, uint16_t(tags.size()));
        for (size_t i = 0; i < tags.size(); ++i) {
            put(dir, tags[i].tag);
            put(dir, tags[i].type);
            if (big)
                put(dir, uint64_t(tags[i].count));
            else
                put(dir, uint32_t(tags[i].count));
            if (tags[i].data.size() <= inline_bytes) {
                std::vector<unsigned char> v = tags[i].data;
                v.resize(inline_bytes, 0);
                dir.insert(dir.end(), v.begin(), v.end());
            } else if (big) {
                put(dir, uint64_t(placed[i]));
            } else {
                put(dir, uint32_t(placed[i]));
            }
        }
        const uint64_t next_link = ifd + dir.size();
        if (big)
            put(dir, uint64_t(0));
        else
            put(dir, uint32_t(0));
        out.write(dir.data(), dir.size());
        // link the previous directory (or the header) to this one
        if (fseeko(out.f, (off_t)link_pos, SEEK_SET) != 0)
            out.ok = false;
        if (big) {
            const uint64_t v = ifd;
            if (fwrite(&v, 1, 8, out.f) != 8)
                out.ok = false;
        } else {
            const uint32_t v = (uint32_t)ifd;
            if (fwrite(&v, 1, 4, out.f) != 4)
                out.ok = false;
        }
        if (fseeko(out.f, 0, SEEK_END) != 0)
            out.ok = false;
        link_pos = next_link;
        if (!big && out.pos > 0xffffffffull)
            out.ok = false;
    }
    if (fflush(out.f) != 0)
        out.ok = false;
    if (!out.ok) {
        set_err(err, errlen, "write_tiff: write to \"%s\" failed", path);
        return B2R_ERR_INVALID;
    }
    return B2R_OK;
}


// ---------------------------------------------------------------------------
// OpenEXR container around a raw (uncompressed) tile set
// ---------------------------------------------------------------------------
namespace {

void
exr_attr(std::vector<unsigned char>& h, const char* name, const char* type,
         const std::vector<unsigned char>& data)
{
    h.insert(h.end(), name, name + strlen(name) + 1);
    h.insert(h.end(), type, type + strlen(type) + 1);
    put(h, int32_t(data.size()));
    h.insert(h.end(), data.begin(), data.end());
}

// OpenEXR's zip codec on one tile: bytes interleaved (even ones, then odd
// ones), delta-coded, deflated; a block that does not shrink is stored raw
void
exr_zip(const std::vector<unsigned char>& raw, std::vector<unsigned char>& tmp,
        std::vector<unsigned char>& out, int zlevel)
{
    const size_t n = raw.size();
    tmp.resize(n);
    unsigned char* t1 = tmp.data();
    unsigned char* t2 = tmp.data() + (n + 1) / 2;
    for (size_t i = 0; i < n; i += 2) {
        *t1++ = raw[i];
        if (i + 1 < n)
            *t2++ = raw[i + 1];
    }
    unsigned char prev = tmp.empty() ? 0 : tmp[0];
    for (size_t i = 1; i < n; ++i) {
        const unsigned char cur = tmp[i];
        tmp[i] = (unsigned char)(cur - prev + 128);  // (cur - prev + 384) mod 256
        prev   = cur;
    }
    out.resize(compressBound((uLong)n));
    uLongf zn = (uLongf)out.size();
    if (n == 0 || compress2(out.data(), &zn, tmp.data(), (uLong)n, zlevel) != Z_OK || zn >= n)
        out = raw;
    else
        out.resize(zn);
}

}  // namespace

extern "C" int
b2r_tileset_write_exr(const b2r_tileset* ts, const char* path, const b2r_exr_options* opt,
                      char* err, size_t errlen)
{
    if (!ts || !path) {
        set_err(err, errlen, "write_exr: null argument");
        return B2R_ERR_INVALID;
    }
    if (ts->levels.empty() || ts->first_level != 0) {
        set_err(err, errlen, "write_exr: the tile set must hold every level of the chain");
        return B2R_ERR_INVALID;
    }
    if (ts->compression || ts->predictor
        || (ts->out_basetype != B2R_HALF && ts->out_basetype != B2R_FLOAT)) {
        set_err(err, errlen,
                "write_exr: needs stored (compression = 0, predictor = 0) float or half tiles "
                "-- OpenEXR's zip codec re-orders the bytes of a tile itself");
        return B2R_ERR_INVALID;
    }
    const int nch = ts->nch;
    if (nch < 1 || nch > 4) {
        set_err(err, errlen, "write_exr: 1 to 4 channels (Y, YA, RGB, RGBA), not %d", nch);
        return B2R_ERR_UNSUPPORTED;
    }
    // the chain must be the full MIP pyramid OpenEXR expects (ROUND_DOWN)
    {
        int w = ts->levels[0].w, h = ts->levels[0].h;
        size_t n = 1;
        while (w > 1 || h > 1) {
            w = std::max(1, w / 2), h = std::max(1, h / 2);
            if (n >= ts->levels.size() || ts->levels[n].w != w || ts->levels[n].h != h) {
                set_err(err, errlen, "write_exr: the chain is not a full MIP pyramid");
                return B2R_ERR_INVALID;
            }
            ++n;
        }
        if (n != ts->levels.size()) {
            set_err(err, errlen, "write_exr: the chain is not a full MIP pyramid");
            return B2R_ERR_INVALID;
        }
    }
    static const char* const names_by_nch[4][4] = { { "Y" }, { "Y", "A" }, { "R", "G", "B" },
                                                    { "R", "G", "B", "A" } };
    // channels are stored in alphabetical order
    int order[4] = { 0, 1, 2, 3 };
    std::sort(order, order + nch, [&](int a, int b) {
        return strcmp(names_by_nch[nch - 1][a], names_by_nch[nch - 1][b]) < 0;
    });
    const int es    = ts->out_basetype == B2R_HALF ? 2 : 4;
    const int ptype = ts->out_basetype == B2R_HALF ? 1 : 2;
    const int W = ts->levels[0].w, H = ts->levels[0].h;

    // attribute names: 31 characters in a plain version-2 file, 255 with the
    // "long names" version flag
    bool long_names = false;
    for (int i = 0; opt && i < opt->nattrs; ++i) {
        if (!opt->attr_names || !opt->attr_names[i])
            continue;
        const size_t n = strlen(opt->attr_names[i]);
        if (n > 255) {
            set_err(err, errlen, "write_exr: attribute name longer than 255 characters");
            return B2R_ERR_INVALID;
        }
        long_names = long_names || n > 31;
    }
    std::vector<unsigned char> hdr;
    put(hdr, uint32_t(20000630));
    put(hdr, uint32_t(2 | 0x200 | (long_names ? 0x400 : 0)));  // version 2, single-part tiled
    {
        std::vector<unsigned char> d;
        for (int k = 0; k < nch; ++k) {
            const char* nm = names_by_nch[nch - 1][order[k]];
            d.insert(d.end(), nm, nm + strlen(nm) + 1);
            put(d, int32_t(ptype));
            d.push_back(0);  // pLinear
            d.push_back(0), d.push_back(0), d.push_back(0);
            put(d, int32_t(1));
            put(d, int32_t(1));
        }
        d.push_back(0);
        exr_attr(hdr, "channels", "chlist", d);
    }
    exr_attr(hdr, "compression", "compression", { 3 });  // ZIP_COMPRESSION
    {
        std::vector<unsigned char> d;
        put(d, int32_t(0)), put(d, int32_t(0)), put(d, int32_t(W - 1)), put(d, int32_t(H - 1));
        exr_attr(hdr, "dataWindow", "box2i", d);
        exr_attr(hdr, "displayWindow", "box2i", d);
    }
    exr_attr(hdr, "lineOrder", "lineOrder", { 0 });  // INCREASING_Y
    {
        std::vector<unsigned char> d;
        put(d, 1.0f);
        exr_attr(hdr, "pixelAspectRatio", "float", d);
        exr_attr(hdr, "screenWindowWidth", "float", d);
        d.clear();
        put(d, 0.0f), put(d, 0.0f);
        exr_attr(hdr, "screenWindowCenter", "v2f", d);
    }
    {
        std::vector<unsigned char> d;
        put(d, uint32_t(ts->tile_w)), put(d, uint32_t(ts->tile_h));
        d.push_back(1);  // MIPMAP_LEVELS | ROUND_DOWN << 4
        exr_attr(hdr, "tiles", "tiledesc", d);
    }
    for (int i = 0; opt && i < opt->nattrs; ++i) {
        if (!opt->attr_names || !opt->attr_values || !opt->attr_names[i] || !opt->attr_values[i]
            || !opt->attr_names[i][0])
            continue;
        const char* v = opt->attr_values[i];
        exr_attr(hdr, opt->attr_names[i], "string", std::vector<unsigned char>(v, v + strlen(v)));
    }
    hdr.push_back(0);

    // every tile: crop to the level, planar by channel per scanline, zip -- on
    // host threads; then the chunks and the offset table
    struct Job {
        int level, tx, ty;
    };
    std::vector<Job> jobs;
    for (int lv = 0; lv < (int)ts->levels.size(); ++lv)
        for (int ty = 0; ty < ts->levels[lv].nty; ++ty)
            for (int tx = 0; tx < ts->levels[lv].ntx; ++tx)
                jobs.push_back({ lv, tx, ty });
    std::vector<std::vector<unsigned char>> blocks(jobs.size());
    const int zlevel = (opt && opt->zlevel > 0) ? std::min(9, (int)opt->zlevel) : 1;
    int nthreads     = (opt && opt->nthreads > 0) ? opt->nthreads
                                                  : (int)std::thread::hardware_concurrency();
    nthreads         = std::max(1, std::min({ nthreads, 256, (int)jobs.size() }));
    std::atomic<size_t> next(0);
    auto work = [&]() {
        std::vector<unsigned char> raw, tmp;
        for (;;) {
            const size_t j = next.fetch_add(1);
            if (j >= jobs.size())
                break;
            const auto& L = ts->levels[jobs[j].level];
            const auto& t = L.tiles[size_t(jobs[j].ty) * L.ntx + jobs[j].tx];
            const int vw  = std::min(ts->tile_w, L.w - jobs[j].tx * ts->tile_w);
            const int vh  = std::min(ts->tile_h, L.h - jobs[j].ty * ts->tile_h);
            raw.resize(size_t(vw) * vh * nch * es);
            unsigned char* o = raw.data();
            for (int y = 0; y < vh; ++y) {
                const unsigned char* row = t.data() + size_t(y) * ts->tile_w * nch * es;
                for (int k = 0; k < nch; ++k) {
                    const unsigned char* p = row + size_t(order[k]) * es;
                    for (int x = 0; x < vw; ++x, p += size_t(nch) * es, o += es)
                        memcpy(o, p, es);
                }
            }
            exr_zip(raw, tmp, blocks[j], zlevel);
        }
    };
    {
        std::vector<std::thread> th;
        try {
            for (int k = 1; k < nthreads; ++k)
                th.emplace_back(work);
        } catch (...) {
        }
        work();
        for (auto& x : th)
            x.join();
    }

    File out;
    out.f = fopen(path, "wb");
    if (!out.f) {
        set_err(err, errlen, "write_exr: could not open \"%s\"", path);
        return B2R_ERR_INVALID;
    }
    out.write(hdr.data(), hdr.size());
    std::vector<unsigned char> table;
    uint64_t pos = out.pos + 8ull * jobs.size();
    for (size_t j = 0; j < jobs.size(); ++j) {
        put(table, uint64_t(pos));
        pos += 20 + blocks[j].size();
    }
    out.write(table.data(), table.size());
    for (size_t j = 0; j < jobs.size(); ++j) {
        std::vector<unsigned char> c;
        put(c, int32_t(jobs[j].tx)), put(c, int32_t(jobs[j].ty));
        put(c, int32_t(jobs[j].level)), put(c, int32_t(jobs[j].level));
        put(c, int32_t(blocks[j].size()));
        out.write(c.data(), c.size());
        out.write(blocks[j].data(), blocks[j].size());
    }
    if (fflush(out.f) != 0)
        out.ok = false;
    if (!out.ok) {
        set_err(err, errlen, "write_exr: write to \"%s\" failed", path);
        return B2R_ERR_INVALID;
    }
    return B2R_OK;
}
