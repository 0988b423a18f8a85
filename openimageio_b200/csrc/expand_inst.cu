// expand_inst.cu -- instantiations of resize_expand_kernel for one (channel
// count, V window); compiled 16 times with -DB2R_INST_NCH=<1..4>
// -DB2R_INST_VT=<2,4,6,8> (see Makefile).
#include "expand_kernel.cuh"

#if !defined(B2R_INST_NCH) || !defined(B2R_INST_VT)
#    error "compile with -DB2R_INST_NCH=<1..4> -DB2R_INST_VT=<2,4,6,8>"
#endif

namespace b2r {

#define B2R_CAT4(a, b, c) a##b##_##c
#define B2R_XNAME(nch, vt) B2R_CAT4(expand_lookup_, nch, vt)

ExpandKernel
B2R_XNAME(B2R_INST_NCH, B2R_INST_VT)(int ht)
{
    switch (ht) {
    case 4: return resize_expand_kernel<B2R_INST_NCH, B2R_INST_VT, 4>;
    case 5: return resize_expand_kernel<B2R_INST_NCH, B2R_INST_VT, 5>;
    case 6: return resize_expand_kernel<B2R_INST_NCH, B2R_INST_VT, 6>;
    case 8: return resize_expand_kernel<B2R_INST_NCH, B2R_INST_VT, 8>;
    }
    return nullptr;
}

}  // namespace b2r
