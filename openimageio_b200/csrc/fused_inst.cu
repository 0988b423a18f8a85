// fused_inst.cu -- instantiations of resize_fused_kernel for one (source type,
// channel count); compiled 16 times with -DB2R_INST_ST=<0..3>
// -DB2R_INST_NCH=<1..4> so the build parallelises (see Makefile).
#include "fused_kernel.cuh"

#ifndef B2R_INST_ST
#    error "compile with -DB2R_INST_ST=<0..3> -DB2R_INST_NCH=<1..4>"
#endif

namespace b2r {

template<int VK>
static FusedKernel
pick_ht(int ht)
{
    switch (ht) {
    case 4: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 4>;
    case 6: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 6>;
    case 8: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 8>;
    case 10: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 10>;
    case 14: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 14>;
    case 16: return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 16>;
    }
    // wide H windows (strong minification): one run-time-tap-count variant per
    // ring size, only where it can occur (downsizing: ring mode)
    if (VK > 0 && ht > 16 && ht <= FUSED_MAX_HT && ht % 4 == 0)
        return resize_fused_kernel<B2R_INST_ST, B2R_INST_NCH, VK, 0>;
    return nullptr;
}

#define B2R_CAT3(a, b, c) a##b##_##c
#define B2R_NAME(st, nch) B2R_CAT3(fused_lookup_, st, nch)

FusedKernel
B2R_NAME(B2R_INST_ST, B2R_INST_NCH)(int vk, int ht)
{
    switch (vk) {
    case 0: return pick_ht<0>(ht);
    case 2: return pick_ht<2>(ht);
    case 4: return pick_ht<4>(ht);
    case 6: return pick_ht<6>(ht);
    case 8: return pick_ht<8>(ht);
    }
    return nullptr;
}

}  // namespace b2r
