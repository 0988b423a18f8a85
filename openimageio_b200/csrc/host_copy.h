// host_copy.h -- host <-> device pixel copies that are fast for PAGEABLE host
// memory too.
//
// An ImageBuf normally owns pageable memory.  cudaMemcpyAsync from / to it is
// staged by the driver through one thread (~9 GB/s measured: 58 ms for an 8K
// RGBA float frame against 10.4 ms from pinned memory).  Here rows are moved
// by a few host threads into / out of a small per-device ring of pinned
// bounce buffers, so the DMA engines only ever see pinned memory (3x faster
// end to end).  Pinned (or registered) host memory goes straight to the DMA.
#pragma once

#include <cuda_runtime.h>

#include <cstddef>

namespace b2r {
namespace api {

// true for cudaHostAlloc'ed / cudaHostRegister'ed memory
bool
host_is_pinned(const void* p);

// rows x rowbytes, split over host threads (B2R_COPY_THREADS, default
// min(8, cores / 2))
void
parallel_copy_rows(unsigned char* d, size_t dpitch, const unsigned char* s, size_t spitch,
                   size_t rowbytes, int nrows);

constexpr int BOUNCE_SLOTS = 3;

// A ring of pinned slots; a slot is reusable once the DMA that last used it
// has completed (an event per slot).
struct BounceRing {
    void* buf[BOUNCE_SLOTS]        = { nullptr, nullptr, nullptr };
    size_t cap                     = 0;
    cudaEvent_t done[BOUNCE_SLOTS] = { nullptr, nullptr, nullptr };
    bool busy[BOUNCE_SLOTS]        = { false, false, false };
    bool reserve(size_t bytes);  // idle ring only: make every slot >= bytes
    void* acquire(int slot);     // waits for the slot's last DMA
    void release_after(int slot, cudaStream_t st);
    void idle();                 // caller has synchronised: nothing is in flight
};
struct Bounce {
    BounceRing in, out;
};
// the resize pipeline's rings (used under that device's pipe mutex)
Bounce&
pipeline_bounce(int device);

// Drop-in for cudaMemcpy2DAsync(..., cudaMemcpyHostToDevice, st).  Like that
// call with pageable memory, the host buffer may be reused on return.
cudaError_t
copy_h2d_2d(void* ddst, size_t dpitch, const void* hsrc, size_t spitch, size_t rowbytes,
            size_t nrows, cudaStream_t st);
// Drop-in for cudaMemcpy2DAsync(..., cudaMemcpyDeviceToHost, st).  With
// pinned memory it is asynchronous as usual; with pageable memory it returns
// when the host rows are complete (it has to wait for the data to copy it).
cudaError_t
copy_d2h_2d(void* hdst, size_t dpitch, const void* dsrc, size_t spitch, size_t rowbytes,
            size_t nrows, cudaStream_t st);

}  // namespace api
}  // namespace b2r
