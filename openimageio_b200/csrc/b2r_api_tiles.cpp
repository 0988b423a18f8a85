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
// The result is an in-memory tile set (per level: a grid of compressed tiles);
// the container formats themselves (TIFF / OpenEXR headers, offsets tables)
// stay with the reference's file plugins.
#include "api_internal.h"

#include <zlib.h>

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
    int nch = 0, tile_w = 0, tile_h = 0, out_basetype = 0, compression = 0;
    std::vector<Level> levels;
    long long raw_bytes = 0, stored_bytes = 0;
    float total_ms = 0, device_ms = 0, deflate_busy_ms = 0, setup_ms = 0, wait_ms = 0;
};

namespace {

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
    if (obt != B2R_FLOAT && obt != B2R_HALF) {
        set_err(err, errlen, "encode_tiles: tiles are float32 or half");
        return B2R_ERR_TYPES;
    }
    int nthreads = (to && to->nthreads > 0) ? to->nthreads
                                            : (int)std::thread::hardware_concurrency();
    nthreads     = std::max(1, std::min(nthreads, 256));
    const int nch = chain->nch;
    const int oes = obt == B2R_HALF ? 2 : 4;
    const size_t tile_bytes = size_t(tile_w) * tile_h * nch * oes;

    auto t_begin = std::chrono::steady_clock::now();
    DeviceGuard guard(chain->device);
    std::unique_ptr<b2r_tileset> ts(new b2r_tileset);
    ts->nch          = nch;
    ts->tile_w       = tile_w;
    ts->tile_h       = tile_h;
    ts->out_basetype = obt;
    ts->compression  = comp;

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
        launch_tile_relayout((const float*)CL.px, CL.w, CL.h, nch, tile_w, tile_h, c.ty0, c.ty1,
                             obt == B2R_HALF, dslot[s], sk);
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
