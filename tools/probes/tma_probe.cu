// Standalone probe: uint8 TMA box loads with negative / out-of-frame coordinates.
//   tma_probe <rank 2|3> <box_w> <box_h> <l2promo 0..3> <prefetch 0|1>
// nvcc -gencode arch=compute_100a,code=sm_100a -I../../mulut_b200/csrc tma_probe.cu -o tma_probe
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cudaTypedefs.h>
#include "tma.cuh"
using namespace mulut;

__global__ void probe3(const __grid_constant__ CUtensorMap tmap, int x, int y, int n, uint8_t *out, int bytes, int pf)
{
    extern __shared__ __align__(1024) uint8_t tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (pf) tma_prefetch_desc(&tmap);
        mbar_expect_tx(smem_u32(&bar), bytes);
        tma_load_3d(smem_u32(tile), &tmap, x, y, n, smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}
__global__ void probe2(const __grid_constant__ CUtensorMap tmap, int x, int y, uint8_t *out, int bytes, int pf)
{
    extern __shared__ __align__(1024) uint8_t tile[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_fence_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (pf) tma_prefetch_desc(&tmap);
        mbar_expect_tx(smem_u32(&bar), bytes);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(tile)),
            "l"(&tmap), "r"(x), "r"(y), "r"(smem_u32(&bar))
            : "memory");
    }
    mbar_wait(smem_u32(&bar), 0);
    for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = tile[i];
}

int main(int argc, char **argv)
{
    const int rank = argc > 1 ? atoi(argv[1]) : 3, BW = argc > 2 ? atoi(argv[2]) : 112, BH = argc > 3 ? atoi(argv[3]) : 36;
    const int promo = argc > 4 ? atoi(argv[4]) : 2, pf = argc > 5 ? atoi(argv[5]) : 0;
    const int N = 2, H = 64, WC = 288;
    std::vector<uint8_t> h(N * H * WC);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + i / WC);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, BW * BH);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    auto fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
    alignas(64) CUtensorMap map;
    cuuint64_t dims[3] = {(cuuint64_t)WC, (cuuint64_t)(rank == 2 ? N * H : H), (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)WC, (cuuint64_t)WC * H};
    cuuint32_t box[3] = {(cuuint32_t)BW, (cuuint32_t)BH, 1}, es[3] = {1, 1, 1};
    CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, (CUtensorMapL2promotion)promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("rank %d box %dx%d promo %d pf %d: encode=%d\n", rank, BW, BH, promo, pf, (int)r);
    if (r) return 1;
    for (int t = 0; t < 3; ++t) {
        int x = t == 0 ? 96 : t == 1 ? -16 : 176, y = t == 0 ? 10 : t == 1 ? -2 : 30, n = t == 2 ? 1 : 0;
        if (rank == 3) probe3<<<1, 128, BW * BH>>>(map, x, y, n, o, BW * BH, pf);
        else probe2<<<1, 128, BW * BH>>>(map, x, y + n * H, o, BW * BH, pf);
        cudaError_t e = cudaDeviceSynchronize();
        printf("  case %d: %s\n", t, cudaGetErrorString(e));
        if (e != cudaSuccess) return 2;
        std::vector<uint8_t> rr(BW * BH);
        cudaMemcpy(rr.data(), o, rr.size(), cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < BH; ++i)
            for (int j = 0; j < BW; ++j) {
                int gy = y + i, gx = x + j;
                bool oob = gx < 0 || gx >= WC || (rank == 3 ? (gy < 0 || gy >= H) : (gy + n * H < 0 || gy + n * H >= N * H));
                uint8_t want = oob ? 0 : h[((size_t)n * H + gy) * WC + gx];
                bad += rr[i * BW + j] != want;
            }
        printf("     mismatches %d\n", bad);
    }
    return 0;
}
