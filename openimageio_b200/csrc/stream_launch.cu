// stream_launch.cu -- host side of resize_stream_kernel: tensor map, dispatch.
#include "kernels.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <mutex>

namespace b2r {

#define B2R_DECL_LOOKUP(ST, NCH) StreamKernel stream_lookup_##ST##_##NCH(int, int, int, int, int*);
#define B2R_DECL_ST(ST) \
    B2R_DECL_LOOKUP(ST, 1) B2R_DECL_LOOKUP(ST, 2) B2R_DECL_LOOKUP(ST, 3) B2R_DECL_LOOKUP(ST, 4)
B2R_DECL_ST(0) B2R_DECL_ST(1) B2R_DECL_ST(2) B2R_DECL_ST(3)

static int
st_index(int basetype)
{
    switch (basetype) {
    case BT_UINT8: return 0;
    case BT_UINT16: return 1;
    case BT_HALF: return 2;
    case BT_FLOAT: return 3;
    }
    return -1;
}

static StreamKernel
pick_stream(int srctype, int nch, int vk, int ht, int w2, int hc, int* smem)
{
    typedef StreamKernel (*Lookup)(int, int, int, int, int*);
#define B2R_ROW(ST) \
    { stream_lookup_##ST##_1, stream_lookup_##ST##_2, stream_lookup_##ST##_3, stream_lookup_##ST##_4 }
    static const Lookup table[4][4] = { B2R_ROW(0), B2R_ROW(1), B2R_ROW(2), B2R_ROW(3) };
    int st = st_index(srctype);
    if (st < 0 || nch < 1 || nch > 4)
        return nullptr;
    return table[st][nch - 1](vk, ht, w2, hc, smem);
}

bool
stream_supported(int srctype, int nch, int vk, int ht, int w2, int hicomp)
{
    int smem = 0;
    return pick_stream(srctype, nch, vk, ht, w2, hicomp, &smem) != nullptr;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point lookup (the
// library does not link libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn
encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q)
                == cudaSuccess
            && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
        else
            cudaGetLastError();
    });
    return fn;
}

int
stream_encode_map(void* mapv, const FusedParams& p, int w2, const void* tma_base,
                  int tma_rows)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || tma_rows <= 0)
        return -1;
    const int es = basetype_size(p.srctype);
    if (((uintptr_t)tma_base & 15) || (p.src_ystride & 15) || p.src_ystride <= 0)
        return -2;
    const cuuint64_t dims[2]    = { (cuuint64_t)p.src_w * p.nch, (cuuint64_t)tma_rows };
    const cuuint64_t strides[1] = { (cuuint64_t)p.src_ystride };
    const cuuint32_t box[2]     = { 256u, (cuuint32_t)stream_rows(w2) };
    const cuuint32_t estr[2]    = { 1u, 1u };
    const CUtensorMapDataType dt = es == 4   ? CU_TENSOR_MAP_DATA_TYPE_UINT32
                                   : es == 2 ? CU_TENSOR_MAP_DATA_TYPE_UINT16
                                             : CU_TENSOR_MAP_DATA_TYPE_UINT8;
    if (enc((CUtensorMap*)mapv, dt, 2, const_cast<void*>(tma_base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
        != CUDA_SUCCESS)
        return -3;
    return 0;
}

static int
launch_with_map(const FusedParams& p, int w2, const CUtensorMap& map, int sms, void* stream)
{
    int smem = 0;
    StreamKernel k = pick_stream(p.srctype, p.nch, p.vk, p.ht, w2, p.hicomp, &smem);
    if (!k || p.nstrips <= 0 || p.nbands <= 0)
        return -1;
    if (cudaFuncSetAttribute((const void*)k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
        != cudaSuccess) {
        cudaGetLastError();
        return -4;
    }
    const long long items = (long long)p.nstrips * p.nbands * (p.nframes > 1 ? p.nframes : 1);
    dim3 grid((unsigned)std::min<long long>(items, std::max(1, sms)), 1, 1);
    // four-element groups per V thread
    const int e2      = 1;  // see StreamE2 in stream_kernel.cuh
    const int threads = STREAM_PTHREADS + stream_vthreads(p.nch, w2) / e2 + STREAM_HTHREADS;
    launch_pdl(k, grid, dim3(threads), (size_t)smem, (cudaStream_t)stream, map, p);
    return 0;
}

int
launch_resize_stream(const FusedParams& p, int w2, const void* tma_base, int tma_rows,
                     int sms, void* stream)
{
    alignas(64) CUtensorMap map;
    int rc = stream_encode_map(&map, p, w2, tma_base, tma_rows);
    if (rc)
        return rc;
    return launch_with_map(p, w2, map, sms, stream);
}

int
launch_resize_stream_batch(const FusedParams& p, int w2, int sms, void* stream)
{
    if (p.nframes < 1 || !p.batch_maps || !p.batch_dst)
        return -1;
    alignas(64) CUtensorMap unused;
    memset(&unused, 0, sizeof(unused));
    return launch_with_map(p, w2, unused, sms, stream);
}

}  // namespace b2r
