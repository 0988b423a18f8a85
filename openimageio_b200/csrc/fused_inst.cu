// fused_inst.cu -- instantiations of resize_fused_kernel for one (source type,
// channel count); compiled 16 times with -DB2R_INST_ST=<0..3>
// -DB2R_INST_NCH=<1..4> so the build parallelises (see Makefile).
#include "fused_kernel.cuh"

#ifndef B2R_INST_ST
#    error "compile with -DB2R_INST_ST=<0..3> -DB2R_INST_NCH=<1..4>"
#endif

namespace b2r {

// Gather mode (VK = 0, upsizing shapes the expand kernel does not cover):
// compile-time H windows.  Ring mode (downsizing): only the variant with a
// run-time H window (strong minification, 17..128 taps) -- narrower windows
// belong to resize_stream_kernel.
template<int VK>
static FusedKernel
pick_ht(int ht)
{
    if (VK == 0) {
        switch (ht) {
        case 4: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, 0, 4>;
        case 6: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, 0, 6>;
        case 8: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, 0, 8>;
        case 10: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, 0, 10>;
        case 14: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, 0, 14>;
        case 16: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, 0, 16>;
        }
        return nullptr;
    }
    if (ht > 16 && ht <= FUSED_MAX_HT && ht % 4 == 0)
        return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, (VK > 0 ? VK : 4), 0>;
    return nullptr;
}

#define B2R_CAT3(a, b, c) a##b##_##c
#define B2R_NAME(st, nch) B2R_CAT3(fused_lookup_, st, nch)

FusedKernel
B2R_NAME(B2R_INST_ST, B2R_INST_NCH)(int vk, int ht)
{
    switch (vk) {
    case 0: return pick_ht<0>(ht);
    case 4: return pick_ht<4>(ht);
    case 6: return pick_ht<6>(ht);
    case 8: return pick_ht<8>(ht);
    }
    return nullptr;
}

}  // namespace b2r
