// host_copy.cpp -- see host_copy.h.
#include "host_copy.h"

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#    include <emmintrin.h>
#endif

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

namespace {

// Large copies between pageable and pinned host memory: the destination is
// not read again by this CPU (a bounce slot the DMA engine reads, or the
// caller's image), so the lines are written with non-temporal stores -- no
// read-for-ownership, a third less memory traffic than memcpy's cached
// stores for the chunk sizes a copy thread gets (glibc only switches to
// them above its own, larger, threshold).
void
copy_stream(unsigned char* d, const unsigned char* s, size_t n)
{
#if defined(__SSE2__)
    if (n >= (size_t(64) << 10)) {
        const size_t head = (16 - ((uintptr_t)d & 15)) & 15;
        if (head) {
            memcpy(d, s, head);
            d += head, s += head, n -= head;
        }
        size_t blocks = n / 64;
        for (; blocks; --blocks, d += 64, s += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(s));
            const __m128i b = _mm_loadu_si128((const __m128i*)(s + 16));
            const __m128i c = _mm_loadu_si128((const __m128i*)(s + 32));
            const __m128i e = _mm_loadu_si128((const __m128i*)(s + 48));
            _mm_stream_si128((__m128i*)(d), a);
            _mm_stream_si128((__m128i*)(d + 16), b);
            _mm_stream_si128((__m128i*)(d + 32), c);
            _mm_stream_si128((__m128i*)(d + 48), e);
        }
        n &= 63;
        if (n)
            memcpy(d, s, n);
        _mm_sfence();
        return;
    }
#endif
    memcpy(d, s, n);
}

// Persistent copy threads.  A std::thread per part per call cost ~0.2 ms of
// spawning in front of every 0.3-1.2 ms band copy of the host pipeline; the
// workers here wait on a condition variable and are never joined (they hold no
// CUDA state; process exit ends them).
class CopyPool {
    std::mutex m_;
    std::condition_variable work_cv_, done_cv_;
    std::mutex run_m_;  // one job at a time
    const std::function<void(int)>* job_ = nullptr;
    int nparts_ = 0, next_ = 0, running_ = 0;
    unsigned long long gen_ = 0;
    int nworkers_ = 0;

    void worker()
    {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(m_);
        for (;;) {
            work_cv_.wait(lk, [&] { return gen_ != seen; });
            seen = gen_;
            while (job_ && next_ < nparts_) {
                const int part = next_++;
                const std::function<void(int)>* job = job_;
                lk.unlock();
                (*job)(part);
                lk.lock();
                if (--running_ == 0)
                    done_cv_.notify_all();
            }
        }
    }

public:
    void ensure(int n)
    {
        std::lock_guard<std::mutex> lk(m_);
        while (nworkers_ < n) {
            try {
                std::thread(&CopyPool::worker, this).detach();
            } catch (...) {
                break;  // the caller's own thread does whatever is left
            }
            ++nworkers_;
        }
    }
    // runs job(0 .. nparts-1), the calling thread included; returns when all are done
    void run(int nparts, const std::function<void(int)>& job)
    {
        std::lock_guard<std::mutex> one(run_m_);
        {
            std::lock_guard<std::mutex> lk(m_);
            job_     = &job;
            nparts_  = nparts;
            next_    = 0;
            running_ = nparts;
            ++gen_;
        }
        work_cv_.notify_all();
        std::unique_lock<std::mutex> lk(m_);
        while (next_ < nparts_) {
            const int part = next_++;
            lk.unlock();
            job(part);
            lk.lock();
            --running_;
        }
        done_cv_.wait(lk, [&] { return running_ == 0; });
        job_ = nullptr;
    }
};

CopyPool&
copy_pool()
{
    static CopyPool* pool = new CopyPool;  // never destroyed, see above
    return *pool;
}

}  // namespace

void
parallel_copy_rows(unsigned char* d, size_t dpitch, const unsigned char* s, size_t spitch,
                   size_t rowbytes, int nrows)
{
    const int nt = (size_t)nrows * rowbytes < (size_t(1) << 20)
                       ? 1
                       : std::min(host_copy_threads(), nrows);
    const std::function<void(int)> work = [=](int t) {
        const int a = (int)((long long)nrows * t / nt), b = (int)((long long)nrows * (t + 1) / nt);
        if (dpitch == rowbytes && spitch == rowbytes) {
            copy_stream(d + (size_t)a * dpitch, s + (size_t)a * spitch,
                        (size_t)(b - a) * rowbytes);
            return;
        }
        for (int r = a; r < b; ++r)
            copy_stream(d + (size_t)r * dpitch, s + (size_t)r * spitch, rowbytes);
    };
    if (nt == 1) {
        work(0);
        return;
    }
    CopyPool& pool = copy_pool();
    pool.ensure(nt - 1);
    pool.run(nt, work);
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
