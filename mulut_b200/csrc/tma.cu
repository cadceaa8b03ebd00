// Host side of tma.cuh: encodes the tensor map of a uint8 frame batch through the
// driver entry point (no link-time dependency on libcuda).
#include <cudaTypedefs.h>

#include "common.cuh"
#include "tma.cuh"

namespace mulut {

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn()
{
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

bool tma_mappable(const void *base, int H, int row_bytes) { return tma_frame_ok(base, H, row_bytes); }

int tma_encode_frames(CUtensorMap *map, const void *base, int N, int H, int WC, int pitch, int box_w, int box_h)
{
    PFN_cuTensorMapEncodeTiled_v12000 fn = encode_fn();
    if (!fn || !tma_frame_ok(base, H, pitch) || pitch < WC || N < 1) return 1;
    const cuuint64_t dims[3] = {(cuuint64_t)WC, (cuuint64_t)H, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * (cuuint64_t)H};   // bytes, dims 1 and 2
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace mulut
