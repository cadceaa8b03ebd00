// TMA (cp.async.bulk / cp.async.bulk.tensor) and mbarrier plumbing for the tiled
// kernels: PTX wrappers on the device side, tensor-map encoding on the host side.
//
// The frames are uint8 (N, H, W, C) interleaved, i.e. H x (W*C) byte rows.  A
// halo'd tile is a 3-D box {box_w bytes, box_h rows, 1 frame} of the tensor
// {W*C, H, N}; coordinates may be negative or run past the edge (the hardware
// zero-fills; the kernels patch the replicate border in shared memory).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace mulut {

// ---------------------------------------------------------------------------
// host: tensor map for a uint8 frame batch
// ---------------------------------------------------------------------------
// TMA needs a 16-byte aligned base and 16-byte multiples for every global stride.
inline bool tma_frame_ok(const void *base, int H, int WC)
{
    return base && (reinterpret_cast<uintptr_t>(base) & 15) == 0 && WC > 0 && H > 0 && (WC & 15) == 0;
}

// Returns 0 on success.  box_w must be a multiple of 16 and <= 256, box_h <= 256.
// Frames of H rows of WC bytes, rows `pitch` bytes apart (pitch % 16 == 0, base 16-byte aligned).
int tma_encode_frames(CUtensorMap *map, const void *base, int N, int H, int WC, int pitch, int box_w, int box_h);

// ---------------------------------------------------------------------------
// device: mbarrier + bulk copies (shared::cta addresses as 32-bit)
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity)
{
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    while (!mbar_try_wait(bar, parity)) {}
}

// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16)
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// 3-D tiled tensor copy global -> shared: box at (x, y, n)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap *map, int x, int y, int n, uint32_t bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(dst),
        "l"(map), "r"(x), "r"(y), "r"(n), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

}  // namespace mulut
