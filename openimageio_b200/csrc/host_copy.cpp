// host_copy.cpp -- see host_copy.h.
#include "host_copy.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

namespace b2r {
namespace api {

bool
host_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

static int
host_copy_threads()
{
    static const int n = [] {
        if (const char* e = getenv("B2R_COPY_THREADS")) {
            int v = atoi(e);
            if (v > 0)
                return std::min(v, 64);
        }
        unsigned hc = std::thread::hardware_concurrency();
        return (int)std::max(1u, std::min(8u, hc / 2));
    }();
    return n;
}

void
parallel_copy_rows(unsigned char* d, size_t dpitch, const unsigned char* s, size_t spitch,
                   size_t rowbytes, int nrows)
{
    const int nt = (size_t)nrows * rowbytes < (size_t(1) << 20)
                       ? 1
                       : std::min(host_copy_threads(), nrows);
    auto work = [=](int t) {
        const int a = (int)((long long)nrows * t / nt), b = (int)((long long)nrows * (t + 1) / nt);
        if (dpitch == rowbytes && spitch == rowbytes) {
            memcpy(d + (size_t)a * dpitch, s + (size_t)a * spitch, (size_t)(b - a) * rowbytes);
            return;
        }
        for (int r = a; r < b; ++r)
            memcpy(d + (size_t)r * dpitch, s + (size_t)r * spitch, rowbytes);
    };
    if (nt == 1) {
        work(0);
        return;
    }
    std::vector<std::thread> th;
    int started = 1;  // part 0 runs on this thread
    try {
        th.reserve(nt - 1);
        for (; started < nt; ++started)
            th.emplace_back(work, started);
    } catch (...) {
        // no more threads to be had: this thread does the parts not handed out
    }
    work(0);
    for (int t = started; t < nt; ++t)
        work(t);
    for (auto& x : th)
        x.join();
}

bool
BounceRing::reserve(size_t bytes)
{
    if (bytes <= cap)
        return true;
    for (int i = 0; i < BOUNCE_SLOTS; ++i) {
        if (buf[i])
            cudaFreeHost(buf[i]);
        buf[i] = nullptr;
        if (!done[i])
            cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming);
    }
    cap = 0;
    for (int i = 0; i < BOUNCE_SLOTS; ++i)
        if (cudaHostAlloc(&buf[i], bytes, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
    cap = bytes;
    return true;
}

void*
BounceRing::acquire(int i)
{
    if (busy[i]) {
        cudaEventSynchronize(done[i]);
        busy[i] = false;
    }
    return buf[i];
}

void
BounceRing::release_after(int i, cudaStream_t st)
{
    cudaEventRecord(done[i], st);
    busy[i] = true;
}

void
BounceRing::idle()
{
    for (int i = 0; i < BOUNCE_SLOTS; ++i)
        busy[i] = false;
}

Bounce&
pipeline_bounce(int dev)
{
    static Bounce b[64];
    return b[dev & 63];
}

namespace {

constexpr size_t SLOT_BYTES = size_t(32) << 20;  // per bounce slot
constexpr size_t MIN_BOUNCE = size_t(4) << 20;   // below this the driver's path is fine

struct CopyRings {
    std::mutex m;
    BounceRing ring;
};
CopyRings&
copy_rings()
{
    static CopyRings r[64];
    int dev = 0;
    cudaGetDevice(&dev);
    return r[dev & 63];
}

}  // namespace

cudaError_t
copy_h2d_2d(void* ddst, size_t dpitch, const void* hsrc, size_t spitch, size_t rowbytes,
            size_t nrows, cudaStream_t st)
{
    if (!rowbytes || !nrows)
        return cudaSuccess;
    if (rowbytes * nrows >= MIN_BOUNCE && rowbytes <= SLOT_BYTES && !host_is_pinned(hsrc)) {
        CopyRings& R = copy_rings();
        std::lock_guard<std::mutex> lock(R.m);
        if (R.ring.reserve(SLOT_BYTES)) {
            const size_t chunk = std::max<size_t>(1, SLOT_BYTES / rowbytes);
            int slot           = 0;
            cudaError_t e      = cudaSuccess;
            for (size_t r = 0; r < nrows && e == cudaSuccess; r += chunk) {
                const size_t n    = std::min(chunk, nrows - r);
                unsigned char* pb = (unsigned char*)R.ring.acquire(slot);
                parallel_copy_rows(pb, rowbytes, (const unsigned char*)hsrc + r * spitch, spitch,
                                   rowbytes, (int)n);
                e = cudaMemcpy2DAsync((unsigned char*)ddst + r * dpitch, dpitch, pb, rowbytes,
                                      rowbytes, n, cudaMemcpyHostToDevice, st);
                R.ring.release_after(slot, st);
                slot = (slot + 1) % BOUNCE_SLOTS;
            }
            return e;
        }
    }
    return cudaMemcpy2DAsync(ddst, dpitch, hsrc, spitch, rowbytes, nrows, cudaMemcpyHostToDevice,
                             st);
}

cudaError_t
copy_d2h_2d(void* hdst, size_t dpitch, const void* dsrc, size_t spitch, size_t rowbytes,
            size_t nrows, cudaStream_t st)
{
    if (!rowbytes || !nrows)
        return cudaSuccess;
    if (rowbytes * nrows >= MIN_BOUNCE && rowbytes <= SLOT_BYTES && !host_is_pinned(hdst)) {
        CopyRings& R = copy_rings();
        std::lock_guard<std::mutex> lock(R.m);
        if (R.ring.reserve(SLOT_BYTES)) {
            const size_t chunk = std::max<size_t>(1, SLOT_BYTES / rowbytes);
            struct Pending {
                unsigned char* host = nullptr;
                size_t rows         = 0;
            } pending[BOUNCE_SLOTS];
            auto drain = [&](int slot) {
                if (!pending[slot].host)
                    return;
                const unsigned char* pb = (const unsigned char*)R.ring.acquire(slot);  // waits
                parallel_copy_rows(pending[slot].host, dpitch, pb, rowbytes, rowbytes,
                                   (int)pending[slot].rows);
                pending[slot].host = nullptr;
            };
            int slot      = 0;
            cudaError_t e = cudaSuccess;
            for (size_t r = 0; r < nrows && e == cudaSuccess; r += chunk) {
                const size_t n = std::min(chunk, nrows - r);
                drain(slot);
                // the ring is shared with copy_h2d_2d, which returns with its
                // slots still in flight (possibly on another stream): wait for
                // whoever used this slot last before the DMA overwrites it
                unsigned char* pb = (unsigned char*)R.ring.acquire(slot);
                e = cudaMemcpy2DAsync(pb, rowbytes, (const unsigned char*)dsrc + r * spitch,
                                      spitch, rowbytes, n, cudaMemcpyDeviceToHost, st);
                if (e != cudaSuccess)
                    break;
                R.ring.release_after(slot, st);
                pending[slot].host = (unsigned char*)hdst + r * dpitch;
                pending[slot].rows = n;
                // the previous chunk's copy-out overlaps this chunk's DMA
                drain((slot + BOUNCE_SLOTS - 1) % BOUNCE_SLOTS);
                slot = (slot + 1) % BOUNCE_SLOTS;
            }
            for (int q = 0; q < BOUNCE_SLOTS; ++q)
                drain((slot + q) % BOUNCE_SLOTS);
            return e;
        }
    }
    return cudaMemcpy2DAsync(hdst, dpitch, dsrc, spitch, rowbytes, nrows, cudaMemcpyDeviceToHost,
                             st);
}

}  // namespace api
}  // namespace b2r
