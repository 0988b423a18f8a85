// stream_inst.cu -- instantiations of resize_stream_kernel for one (source
// type, channel count); compiled 16 times with -DB2R_INST_ST=<0..3>
// -DB2R_INST_NCH=<1..4> so the build parallelises (see Makefile).
#include "stream_kernel.cuh"

#ifndef B2R_INST_ST
#    error "compile with -DB2R_INST_ST=<0..3> -DB2R_INST_NCH=<1..4>"
#endif

namespace b2r {

template<int VK, int W2, bool HC>
static StreamKernel
pick_ht(int ht, int* smem)
{
    *smem = StreamSmem<B2R_INST_ST, VK, W2, B2R_INST_NCH>::TOTAL;
    switch (ht) {
    case 8: return resize_stream_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 8, W2, HC>;
    case 10: return resize_stream_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 10, W2, HC>;
    case 14: return resize_stream_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 14, W2, HC>;
    case 16: return resize_stream_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 16, W2, HC>;
    }
    return nullptr;
}

template<int W2, bool HC>
static StreamKernel
pick_vk(int vk, int ht, int* smem)
{
    switch (vk) {
    case 4: return pick_ht<4, W2, HC>(ht, smem);
    case 6: return pick_ht<6, W2, HC>(ht, smem);
    case 8: return pick_ht<8, W2, HC>(ht, smem);
    }
    return nullptr;
}

#define B2R_CAT3(a, b, c) a##b##_##c
#define B2R_NAME(st, nch) B2R_CAT3(stream_lookup_, st, nch)

StreamKernel
B2R_NAME(B2R_INST_ST, B2R_INST_NCH)(int vk, int ht, int w2, int hc, int* smem)
{
    // highlight compensation fused in: float / half sources (integer types
    // run the three passes)
#if B2R_INST_ST == 3 || B2R_INST_ST == 2
    if (hc)
        return w2 == 2 ? pick_vk<2, true>(vk, ht, smem)
                       : (w2 == 1 ? pick_vk<1, true>(vk, ht, smem) : nullptr);
#else
    if (hc)
        return nullptr;
#endif
    if (w2 == 1)
        return pick_vk<1, false>(vk, ht, smem);
    if (w2 == 2)
        return pick_vk<2, false>(vk, ht, smem);
    return nullptr;
}

}  // namespace b2r
