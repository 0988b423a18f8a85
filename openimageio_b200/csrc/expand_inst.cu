// expand_inst.cu -- instantiations of resize_expand_kernel for one channel
// count; compiled 4 times with -DB2R_INST_NCH=<1..4> (see Makefile).
#include "expand_kernel.cuh"

#ifndef B2R_INST_NCH
#    error "compile with -DB2R_INST_NCH=<1..4>"
#endif

namespace b2r {

#define B2R_CAT2(a, b) a##b
#define B2R_XNAME(nch) B2R_CAT2(expand_lookup_, nch)

ExpandKernel
B2R_XNAME(B2R_INST_NCH)(int ht)
{
    switch (ht) {
    case 4: return resize_expand_kernel<B2R_INST_NCH, 4>;
    case 5: return resize_expand_kernel<B2R_INST_NCH, 5>;
    case 6: return resize_expand_kernel<B2R_INST_NCH, 6>;
    case 8: return resize_expand_kernel<B2R_INST_NCH, 8>;
    }
    return nullptr;
}

}  // namespace b2r
