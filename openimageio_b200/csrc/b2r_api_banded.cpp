// b2r_api_banded.cpp -- the MIP chain of write_mipmap
// (src/libOpenImageIO/maketexture.cpp:823-986) cut into row bands over the
// GPUs of one box, driven from ONE process:
//
//   * level k is split into G contiguous row bands, band g lives in the HBM of
//     device g together with the halo rows level k+1's band g will read
//     (filter radius + the kernels' 8-row margin);
//   * level 0: every device uploads ITS rows (band + halo) straight from the
//     host image -- G PCIe links in parallel, nothing is exchanged;
//   * after a device has computed its band of level k+1 it PUSHES the rows its
//     neighbours need into their buffers with peer copies over NVLink
//     (cudaMemcpyPeerAsync, stream ordered, peer access enabled) and records
//     an event; the neighbour's next level waits on that event on the device
//     (cudaStreamWaitEvent) -- no NCCL, no host wait per level;
//   * when bands get shorter than the halo the level is gathered onto the
//     first device by the same push mechanism and the chain is finished there.
//
// Every level is bit-identical to the single-device chain: the same kernels
// run on the same rows with the same tables (a band's data window is a crop
// of the level, its display window the whole level).
#include "api_internal.h"

#include <chrono>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

using namespace b2r;
using namespace b2r::api;

struct b2r_mipchain_banded {
    struct Band {
        int y0 = 0, y1 = 0;  // rows owned [y0, y1): computed (or uploaded) here
        int a0 = 0, a1 = 0;  // rows held  [a0, a1): owned + halo for the next level
        void* px = nullptr;  // address of row a0
    };
    struct Level {
        int w = 0, h = 0;
        std::vector<Band> bands;  // one per device; y0 == y1 for a device that sits the level out
    };
    int nch = 0;
    std::vector<int> devices;
    std::vector<cudaStream_t> streams;
    std::vector<Level> levels;
    float upload_ms = 0, chain_ms = 0, total_ms = 0;
    float setup_ms = 0, phase_a_ms = 0, phase_b_ms = 0;  // host wall clock of the three parts
    long long halo_bytes = 0;
    int launches = 0;
};

namespace {

// Band buffers are multi-gigabyte cudaMalloc blocks that every peer maps: 10-60
// ms per chain when allocated fresh, against ~40 ms for the whole banded chain
// at two GPUs.  Blocks of destroyed chains are kept per device (exact-size
// reuse, bounded) so a run of same-sized textures -- what maketx jobs are --
// allocates once.
struct BandBlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> blocks[64];
    size_t held[64] = {};
    static constexpr size_t kMaxHeld = size_t(12) << 30;  // per device
    void* take(int dev, size_t bytes)
    {
        std::lock_guard<std::mutex> lock(mu);
        auto& m = blocks[dev & 63];
        auto it = m.find(bytes);
        if (it == m.end())
            return nullptr;
        void* p = it->second;
        m.erase(it);
        held[dev & 63] -= bytes;
        return p;
    }
    // caller has made `dev` current
    void give(int dev, void* p, size_t bytes)
    {
        {
            std::lock_guard<std::mutex> lock(mu);
            if (held[dev & 63] + bytes <= kMaxHeld) {
                blocks[dev & 63].emplace(bytes, p);
                held[dev & 63] += bytes;
                return;
            }
        }
        cudaFree(p);
    }
};
BandBlockCache g_band_blocks;

}  // namespace

namespace b2r {
namespace api {
// b2r_trim_caches: every cached band block goes back to the driver
void
trim_band_blocks()
{
    int prev = 0;
    cudaGetDevice(&prev);
    std::lock_guard<std::mutex> lock(g_band_blocks.mu);
    for (int d = 0; d < 64; ++d) {
        if (g_band_blocks.blocks[d].empty())
            continue;
        cudaSetDevice(d);
        for (auto& kv : g_band_blocks.blocks[d])
            cudaFree(kv.second);
        g_band_blocks.blocks[d].clear();
        g_band_blocks.held[d] = 0;
    }
    cudaSetDevice(prev);
}
}  // namespace api
}  // namespace b2r

namespace {

size_t
band_bytes(const b2r_mipchain_banded& c, int level, int g)
{
    const auto& L = c.levels[level];
    const auto& B = L.bands[g];
    return size_t(B.a1 - B.a0) * L.w * c.nch * 4;
}

b2r_image
level_image(const b2r_mipchain_banded& c, int level, int g)
{
    const auto& L = c.levels[level];
    const auto& B = L.bands[g];
    b2r_image im;
    memset(&im, 0, sizeof(im));
    im.pixels      = B.px;
    im.basetype    = B2R_FLOAT;
    im.nchannels   = c.nch;
    im.x           = 0;
    im.y           = B.a0;
    im.width       = L.w;
    im.height      = B.a1 - B.a0;
    im.full_x      = 0;
    im.full_y      = 0;
    im.full_width  = L.w;
    im.full_height = L.h;
    im.xstride     = (int64_t)c.nch * 4;
    im.ystride     = (int64_t)L.w * c.nch * 4;
    im.memspace    = B2R_MEM_DEVICE;
    im.device      = c.devices[g];
    return im;
}

// Rows [r0, r1) of level `k` that computing rows [y0, y1) of level k+1 reads.
int
needed_rows(const b2r_mipchain_banded& c, int k, int y0, int y1, const std::string& fname,
            int* r0, int* r1, char* err, size_t errlen)
{
    const auto& S = c.levels[k];
    const auto& D = c.levels[k + 1];
    if (fname == "box") {
        // resize_block (:304-334): 2x2 average reads rows 2y, 2y+1; the bilinear
        // form (odd sizes) reads the two rows around the NDC centre
        const bool two_pass = (2 * D.w == S.w) && !(S.w % 2 || S.h % 2);
        if (two_pass) {
            *r0 = 2 * y0;
            *r1 = 2 * y1;
        } else {
            *r0 = std::max(0, int((long long)y0 * S.h / D.h) - 2);
            *r1 = std::min(S.h, int(((long long)y1 * S.h + D.h - 1) / D.h) + 2);
        }
        return B2R_OK;
    }
    b2r_image s, d;
    memset(&s, 0, sizeof(s));
    memset(&d, 0, sizeof(d));
    s.width = s.full_width = S.w;
    s.height = s.full_height = S.h;
    d.width = d.full_width = D.w;
    d.height = d.full_height = D.h;
    s.nchannels = d.nchannels = c.nch;
    b2r_roi roi { 0, D.w, y0, y1, 0, c.nch };
    return b2r_source_rows(&d, &s, fname.c_str(), 0.0f, &roi, r0, r1, err, errlen);
}

}  // namespace

extern "C" int
b2r_mipchain_create_banded(b2r_mipchain_banded** out, const b2r_image* level0,
                           const b2r_mip_options* mo, const int* devices, int ndevices,
                           char* err, size_t errlen)
{
    if (!out || !mo) {
        set_err(err, errlen, "mipchain_banded: null argument");
        return B2R_ERR_INVALID;
    }
    *out   = nullptr;
    int rc = check_image(level0, true, err, errlen);
    if (rc)
        return rc;
    if (level0->basetype != B2R_FLOAT) {
        set_err(err, errlen, "mipchain: levels are float32 (maketx:forcefloat); convert first");
        return B2R_ERR_TYPES;
    }
    int dummy = 0;
    if (memspace_of(level0, &dummy) != B2R_MEM_HOST
        || level0->xstride != (int64_t)level0->nchannels * 4) {
        set_err(err, errlen, "mipchain_banded: level 0 must be a host image, contiguous in x");
        return B2R_ERR_UNSUPPORTED;
    }
    if (mo->sharpen > 0.0f) {
        set_err(err, errlen, "mipchain_banded: sharpening is not supported on banded chains");
        return B2R_ERR_UNSUPPORTED;
    }
    const int ndev_all = device_count_cached();
    if (ndev_all <= 0) {
        set_err(err, errlen, "no CUDA device available: the B200 MIP path has no CPU fallback");
        return B2R_ERR_NO_DEVICE;
    }
    if (ndevices <= 0 || ndevices > 64) {
        set_err(err, errlen, "mipchain_banded: bad device count %d", ndevices);
        return B2R_ERR_INVALID;
    }
    std::unique_ptr<b2r_mipchain_banded> c(new b2r_mipchain_banded);
    for (int g = 0; g < ndevices; ++g) {
        const int d = devices ? devices[g] : g;
        if (d < 0 || d >= ndev_all) {
            set_err(err, errlen, "invalid CUDA device ordinal %d", d);
            return B2R_ERR_INVALID;
        }
        c->devices.push_back(d);
    }
    if (level0->x || level0->y || level0->full_x || level0->full_y
        || level0->full_width != level0->width || level0->full_height != level0->height) {
        // orig_was_overscan (maketexture.cpp:700-703): level 0 keeps its origin
        // through the first resize; the bands here are crops of zero-origin levels
        set_err(err, errlen, "mipchain_banded: level 0 must have a zero origin and "
                             "display window == data window");
        return B2R_ERR_UNSUPPORTED;
    }
    const int G   = ndevices;
    const int nch = level0->nchannels;
    c->nch        = nch;
    std::string fname = (mo->filtername && mo->filtername[0]) ? mo->filtername : "box";
    if (fname != "box") {
        bool known = false;
        for (int i = 0, e = Filter2D::num_filters(); i < e; ++i)
            known |= (fname == Filter2D::get_filterdesc(i).name);
        if (!known) {
            set_err(err, errlen, "Could not make filter \"%s\"", fname.c_str());
            return B2R_ERR_FILTER;
        }
    }
    int prev_dev = 0;
    cudaGetDevice(&prev_dev);
    auto t_begin = std::chrono::steady_clock::now();

    // ---- peer access, streams ---------------------------------------------
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(c->devices[g]);
        device_info(c->devices[g]);
        for (int h = 0; h < G; ++h) {
            if (h == g || c->devices[h] == c->devices[g])
                continue;
            int can = 0;
            cudaDeviceCanAccessPeer(&can, c->devices[g], c->devices[h]);
            if (can) {
                cudaError_t e = cudaDeviceEnablePeerAccess(c->devices[h], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                } else {
                    cudaGetLastError();
                }
            }
        }
        cudaStream_t st = nullptr;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        c->streams.push_back(st);
    }
    auto destroy = [&](int code) {
        for (size_t k = 0; k < c->levels.size(); ++k)
            for (int g = 0; g < G; ++g)
                if (c->levels[k].bands[g].px) {
                    cudaSetDevice(c->devices[g]);
                    // error path: work may still be in flight on the block, so
                    // it is freed (which waits), not handed to the cache
                    cudaFree(c->levels[k].bands[g].px);
                }
        for (int g = 0; g < G; ++g)
            if (c->streams[g]) {
                cudaSetDevice(c->devices[g]);
                cudaStreamDestroy(c->streams[g]);
            }
        cudaSetDevice(prev_dev);
        return code;
    };

    // ---- level sizes and ownership ------------------------------------------
    {
        int w = level0->width, h = level0->height;
        for (;;) {
            b2r_mipchain_banded::Level L;
            L.w = w;
            L.h = h;
            L.bands.resize(G);
            c->levels.push_back(L);
            if (w <= 1 && h <= 1)
                break;
            w = w > 1 ? w / 2 : w;
            h = h > 1 ? h / 2 : h;
        }
    }
    const int nlev = (int)c->levels.size();
    // a level is banded while every band keeps at least `min_band` rows (well
    // above the halo); from the first level that is not, everything is on device 0
    const int min_band = 64;
    bool banded = true;
    for (int k = 0; k < nlev; ++k) {
        auto& L = c->levels[k];
        if (L.h / G < min_band)
            banded = false;
        for (int g = 0; g < G; ++g) {
            auto& B = L.bands[g];
            if (banded) {
                B.y0 = int((long long)L.h * g / G);
                B.y1 = int((long long)L.h * (g + 1) / G);
            } else {
                B.y0 = 0;
                B.y1 = (g == 0) ? L.h : 0;
            }
            B.a0 = B.y0;
            B.a1 = B.y1;
        }
    }
    // rows held = owned + what the next level's band on the same device reads
    for (int k = 0; k + 1 < nlev; ++k) {
        for (int g = 0; g < G; ++g) {
            const auto& N = c->levels[k + 1].bands[g];
            if (N.y1 <= N.y0)
                continue;
            int r0 = 0, r1 = 0;
            rc = needed_rows(*c, k, N.y0, N.y1, fname, &r0, &r1, err, errlen);
            if (rc)
                return destroy(rc);
            auto& B = c->levels[k].bands[g];
            if (B.y1 <= B.y0) {
                B.a0 = r0;
                B.a1 = r1;
            } else {
                B.a0 = std::min(B.a0, r0);
                B.a1 = std::max(B.a1, r1);
            }
        }
    }
    for (int k = 0; k < nlev; ++k) {
        auto& L = c->levels[k];
        for (int g = 0; g < G; ++g) {
            auto& B = L.bands[g];
            if (B.a1 <= B.a0)
                continue;
            cudaSetDevice(c->devices[g]);
            B.px = g_band_blocks.take(c->devices[g], band_bytes(*c, k, g));
            if (!B.px && cudaMalloc(&B.px, band_bytes(*c, k, g)) != cudaSuccess) {
                cudaGetLastError();
                set_err(err, errlen, "mipchain_banded: out of device memory (level %d, device %d)",
                        k, c->devices[g]);
                return destroy(B2R_ERR_CUDA);
            }
        }
    }

    // ---- events: up[g] level 0 landed, done[k][g] level k's band and pushes --
    std::vector<cudaEvent_t> ev_start(G), ev_up(G), ev_end(G);
    std::vector<std::vector<cudaEvent_t>> done(nlev, std::vector<cudaEvent_t>(G, nullptr));
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(c->devices[g]);
        cudaEventCreate(&ev_start[g]);
        cudaEventCreate(&ev_up[g]);
        cudaEventCreate(&ev_end[g]);
        for (int k = 0; k < nlev; ++k)
            cudaEventCreateWithFlags(&done[k][g], cudaEventDisableTiming);
    }
    auto free_events = [&]() {
        for (int g = 0; g < G; ++g) {
            cudaSetDevice(c->devices[g]);
            cudaEventDestroy(ev_start[g]);
            cudaEventDestroy(ev_up[g]);
            cudaEventDestroy(ev_end[g]);
            for (int k = 0; k < nlev; ++k)
                cudaEventDestroy(done[k][g]);
        }
    };

    auto t_setup = std::chrono::steady_clock::now();
    c->setup_ms  = std::chrono::duration<float, std::milli>(t_setup - t_begin).count();
    // ---- phase A (one host thread per device): upload the device's rows of
    // level 0 and build + upload the tables of all its levels -----------------
    std::vector<int> trc(G, B2R_OK);
    std::vector<std::string> terr(G);
    const size_t row0 = size_t(level0->width) * nch * 4;
    auto level_step = [&](int k, int g, bool plan_only, char* e, size_t elen) -> int {
        // rows [y0, y1) of level k+1 on device g from its rows of level k
        const auto& N = c->levels[k + 1].bands[g];
        b2r_image S = level_image(*c, k, g), D = level_image(*c, k + 1, g);
        b2r_roi roi { 0, c->levels[k + 1].w, N.y0, N.y1, 0, nch };
        b2r_options opt;
        memset(&opt, 0, sizeof(opt));
        opt.device      = c->devices[g];
        opt.cuda_stream = (void*)c->streams[g];
        opt.reserved[0] = plan_only ? 1 : 0;
        if (fname == "box") {
            if (plan_only)
                return B2R_OK;
            return b2r_mip_level(&D, &S, "box", &roi, &opt, e, elen);
        }
        return b2r_resize(&D, &S, fname.c_str(), 0.0f, &roi, &opt, nullptr, e, elen);
    };
    {
        std::vector<std::thread> th;
        for (int g = 0; g < G; ++g) {
            th.emplace_back([&, g]() {
                char e[512] = "";
                cudaSetDevice(c->devices[g]);
                cudaEventRecord(ev_start[g], c->streams[g]);
                const auto& B = c->levels[0].bands[g];
                if (B.a1 > B.a0) {
                    cudaError_t ce = copy_h2d_2d(
                        B.px, row0,
                        (const unsigned char*)level0->pixels + (long long)B.a0 * level0->ystride,
                        level0->ystride, row0, B.a1 - B.a0, c->streams[g]);
                    if (ce != cudaSuccess) {
                        trc[g]  = B2R_ERR_CUDA;
                        terr[g] = std::string("CUDA error: ") + cudaGetErrorString(ce);
                        return;
                    }
                }
                cudaEventRecord(ev_up[g], c->streams[g]);
                cudaEventRecord(done[0][g], c->streams[g]);
                for (int k = 0; k + 1 < nlev; ++k) {
                    if (c->levels[k + 1].bands[g].y1 <= c->levels[k + 1].bands[g].y0)
                        continue;
                    int r = level_step(k, g, true, e, sizeof(e));
                    if (r) {
                        trc[g]  = r;
                        terr[g] = e;
                        return;
                    }
                }
            });
        }
        for (auto& t : th)
            t.join();
    }
    for (int g = 0; g < G; ++g)
        if (trc[g]) {
            set_err(err, errlen, "%s", terr[g].c_str());
            free_events();
            return destroy(trc[g]);
        }

    auto t_a      = std::chrono::steady_clock::now();
    c->phase_a_ms = std::chrono::duration<float, std::milli>(t_a - t_setup).count();
    // ---- phase B: the level loop, issued level by level ---------------------
    // level 0's rows came from the host, so nothing is pushed for it; from
    // level 1 on, the owner of a row pushes it to every device that holds it
    for (int k = 0; k + 1 < nlev; ++k) {
        // wait for the rows of level k other devices pushed into my buffer
        for (int g = 0; g < G; ++g) {
            const auto& N = c->levels[k + 1].bands[g];
            if (N.y1 <= N.y0 || k == 0)
                continue;
            const auto& B = c->levels[k].bands[g];
            cudaSetDevice(c->devices[g]);
            for (int h = 0; h < G; ++h) {
                if (h == g)
                    continue;
                const auto& O = c->levels[k].bands[h];
                if (std::max(O.y0, B.a0) < std::min(O.y1, B.a1))
                    cudaStreamWaitEvent(c->streams[g], done[k][h], 0);
            }
        }
        for (int g = 0; g < G; ++g) {
            const auto& N = c->levels[k + 1].bands[g];
            if (N.y1 <= N.y0)
                continue;
            cudaSetDevice(c->devices[g]);
            rc = level_step(k, g, false, err, errlen);
            if (rc) {
                free_events();
                return destroy(rc);
            }
            c->launches += 1;
            if (mo->clamp_half) {
                const auto& L = c->levels[k + 1];
                float* px     = (float*)N.px + size_t(N.y0 - N.a0) * L.w * nch;
                launch_clamp(px, (long long)(N.y1 - N.y0) * L.w * nch, nch, mo->alpha_channel,
                             -65504.0f, 65504.0f, c->streams[g]);
                c->launches += 1;
            }
            // push my rows of level k+1 to the devices that hold them too
            const auto& L  = c->levels[k + 1];
            const size_t rb = size_t(L.w) * nch * 4;
            for (int h = 0; h < G; ++h) {
                if (h == g)
                    continue;
                const auto& T = L.bands[h];
                if (!T.px)
                    continue;
                const int p0 = std::max(N.y0, T.a0), p1 = std::min(N.y1, T.a1);
                if (p0 >= p1)
                    continue;
                cudaMemcpyPeerAsync((unsigned char*)T.px + size_t(p0 - T.a0) * rb, c->devices[h],
                                    (const unsigned char*)N.px + size_t(p0 - N.a0) * rb,
                                    c->devices[g], size_t(p1 - p0) * rb, c->streams[g]);
                c->halo_bytes += (long long)(p1 - p0) * rb;
            }
            cudaEventRecord(done[k + 1][g], c->streams[g]);
        }
    }
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(c->devices[g]);
        cudaEventRecord(ev_end[g], c->streams[g]);
    }
    cudaError_t ce = cudaSuccess;
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(c->devices[g]);
        cudaError_t e2 = cudaStreamSynchronize(c->streams[g]);
        if (e2 == cudaSuccess)
            e2 = cudaGetLastError();
        if (e2 != cudaSuccess)
            ce = e2;
    }
    if (ce != cudaSuccess) {
        set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(ce));
        free_events();
        return destroy(B2R_ERR_CUDA);
    }
    auto t_end  = std::chrono::steady_clock::now();
    c->total_ms = std::chrono::duration<float, std::milli>(t_end - t_begin).count();
    c->phase_b_ms = std::chrono::duration<float, std::milli>(t_end - t_a).count();
    for (int g = 0; g < G; ++g) {
        cudaSetDevice(c->devices[g]);
        float up = 0, ch = 0;
        cudaEventElapsedTime(&up, ev_start[g], ev_up[g]);
        cudaEventElapsedTime(&ch, ev_up[g], ev_end[g]);
        c->upload_ms = std::max(c->upload_ms, up);
        c->chain_ms  = std::max(c->chain_ms, ch);
    }
    free_events();
    cudaSetDevice(prev_dev);
    *out = c.release();
    return B2R_OK;
}

extern "C" int
b2r_mipchain_banded_nlevels(const b2r_mipchain_banded* c)
{
    return c ? (int)c->levels.size() : 0;
}

extern "C" int
b2r_mipchain_banded_level(const b2r_mipchain_banded* c, int level, int* width, int* height)
{
    if (!c || level < 0 || level >= (int)c->levels.size())
        return B2R_ERR_INVALID;
    if (width)
        *width = c->levels[level].w;
    if (height)
        *height = c->levels[level].h;
    return B2R_OK;
}

extern "C" int
b2r_mipchain_banded_band(const b2r_mipchain_banded* c, int level, int band, int* device,
                         int* row_begin, int* row_end, void** device_pixels)
{
    if (!c || level < 0 || level >= (int)c->levels.size() || band < 0
        || band >= (int)c->devices.size())
        return B2R_ERR_INVALID;
    const auto& L = c->levels[level];
    const auto& B = L.bands[band];
    if (device)
        *device = c->devices[band];
    if (row_begin)
        *row_begin = B.y0;
    if (row_end)
        *row_end = B.y1;
    if (device_pixels)
        *device_pixels = (B.px && B.y1 > B.y0)
                             ? (void*)((unsigned char*)B.px
                                       + size_t(B.y0 - B.a0) * L.w * c->nch * 4)
                             : nullptr;
    return B2R_OK;
}

extern "C" int
b2r_mipchain_banded_download(const b2r_mipchain_banded* c, int level, const b2r_image* out,
                             char* err, size_t errlen)
{
    if (!c || level < 0 || level >= (int)c->levels.size() || !out || !out->pixels) {
        set_err(err, errlen, "mipchain_banded: bad level or output image");
        return B2R_ERR_INVALID;
    }
    const auto& L = c->levels[level];
    if (out->basetype != B2R_FLOAT || out->nchannels != c->nch || out->width != L.w
        || out->height != L.h || out->xstride != (int64_t)c->nch * 4) {
        set_err(err, errlen, "mipchain_banded: output image does not match level %d", level);
        return B2R_ERR_INVALID;
    }
    int prev = 0;
    cudaGetDevice(&prev);
    const size_t rb = size_t(L.w) * c->nch * 4;
    const int G     = (int)c->devices.size();
    std::vector<cudaError_t> res(G, cudaSuccess);
    std::vector<std::thread> th;
    for (int g = 0; g < G; ++g) {
        const auto& B = L.bands[g];
        if (B.y1 <= B.y0)
            continue;
        th.emplace_back([&, g]() {
            const auto& Bb = L.bands[g];
            cudaSetDevice(c->devices[g]);
            res[g] = copy_d2h_2d((unsigned char*)out->pixels + (long long)Bb.y0 * out->ystride,
                                 out->ystride,
                                 (const unsigned char*)Bb.px + size_t(Bb.y0 - Bb.a0) * rb, rb, rb,
                                 Bb.y1 - Bb.y0, c->streams[g]);
            if (res[g] == cudaSuccess)
                res[g] = cudaStreamSynchronize(c->streams[g]);
        });
    }
    for (auto& t : th)
        t.join();
    cudaSetDevice(prev);
    for (int g = 0; g < G; ++g)
        if (res[g] != cudaSuccess) {
            set_err(err, errlen, "CUDA error: %s", cudaGetErrorString(res[g]));
            return B2R_ERR_CUDA;
        }
    return B2R_OK;
}

extern "C" float
b2r_mipchain_banded_ms(const b2r_mipchain_banded* c, int what)
{
    if (!c)
        return 0.0f;
    switch (what) {
    case 0: return c->total_ms;   // host wall clock of the whole call
    case 1: return c->upload_ms;  // level-0 upload, device events, max over devices
    case 2: return c->chain_ms;   // level loop incl. pushes, device events, max over devices
    case 3: return float(c->halo_bytes);
    case 4: return c->setup_ms;    // host: peer access, streams, band buffers, events
    case 5: return c->phase_a_ms;  // host: per-device upload threads + tables
    case 6: return c->phase_b_ms;  // host: level loop issue + final synchronise
    }
    return 0.0f;
}

extern "C" void
b2r_mipchain_banded_destroy(b2r_mipchain_banded* c)
{
    if (!c)
        return;
    int prev = 0;
    cudaGetDevice(&prev);
    for (size_t k = 0; k < c->levels.size(); ++k) {
        auto& L = c->levels[k];
        for (size_t g = 0; g < L.bands.size(); ++g)
            if (L.bands[g].px) {
                cudaSetDevice(c->devices[g]);
                // the chain's work is complete (create synchronised its streams;
                // downloads are synchronous), the block can go back for reuse
                g_band_blocks.give(c->devices[g], L.bands[g].px, band_bytes(*c, (int)k, (int)g));
            }
    }
    for (size_t g = 0; g < c->streams.size(); ++g)
        if (c->streams[g]) {
            cudaSetDevice(c->devices[g]);
            cudaStreamDestroy(c->streams[g]);
        }
    cudaSetDevice(prev);
    delete c;
}
