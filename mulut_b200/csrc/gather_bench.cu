// Gather micro-benchmark: measures how fast an sm_100a chip can serve random
// LUT-vertex / LUT-cell fetches from shared memory, L1 and L2 under the access
// shapes the inference kernels can choose from.  Its numbers are the measured
// denominators of the gather roofline (there is no datasheet figure) and the
// evidence behind the layout choices in DESIGN.md.
#include "common.cuh"
#include "tma.cuh"

namespace mulut {

__device__ __forceinline__ uint32_t lcg(uint32_t &s)
{
    s = s * 1664525u + 1013904223u;
    return s;
}

constexpr int GB_U = 4;   // independent gathers in flight per thread and iteration

template <int VARIANT>
__global__ void gather_kernel(const uint8_t *__restrict__ table, uint32_t n_entries, int iters,
                              uint32_t *__restrict__ sink)
{
    extern __shared__ __align__(128) uint8_t gb_smem[];
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0;

    if constexpr (VARIANT == MULUT_GB_LDS_U8 || VARIANT == MULUT_GB_LDS_U32) {
        // table lives in shared memory: n_entries bytes (U8) or words (U32)
        const uint32_t bytes = VARIANT == MULUT_GB_LDS_U8 ? n_entries : n_entries * 4u;
        for (uint32_t i = threadIdx.x * 16u; i < bytes; i += blockDim.x * 16u)
            *reinterpret_cast<uint4 *>(gb_smem + i) = __ldg(reinterpret_cast<const uint4 *>(table + i));
        __syncthreads();
        uint32_t s = gtid * 2654435761u + 12345u;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < GB_U; ++u) {
                const uint32_t idx = __umulhi(lcg(s), n_entries);
                if constexpr (VARIANT == MULUT_GB_LDS_U8) acc += gb_smem[idx];
                else acc += reinterpret_cast<const uint32_t *>(gb_smem)[idx];
            }
        }
    } else if constexpr (VARIANT == MULUT_GB_CPASYNC_CELL64) {
        // per warp: ring of DEPTH stages x 32 cells x 64 B
        constexpr int DEPTH = 4;
        const int warp = threadIdx.x >> 5;
        uint8_t *ring = gb_smem + (size_t)warp * DEPTH * 2048;
        uint32_t sq = (gtid >> 2) * 2654435761u + 777u;   // same stream for the 4 lanes of a quad
        uint32_t sl = gtid * 40503u + 99u;
        auto issue = [&](int stage) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {      // 4 rounds: the quad serves its 4 lanes in turn
                const uint32_t cell = __umulhi(lcg(sq), n_entries);
                const uint8_t *src = table + (size_t)cell * 64 + (lane & 3) * 16;
                // cell slot = owner lane: quad*4 + u
                uint8_t *dst = ring + stage * 2048 + ((lane & ~3) + u) * 64 + (lane & 3) * 16;
                const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int p = 0; p < DEPTH - 1; ++p) issue(p);
        for (int it = 0; it < iters; ++it) {
            issue((it + DEPTH - 1) % DEPTH);
            asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
            __syncwarp();
            const uint32_t *cellw = reinterpret_cast<const uint32_t *>(ring + (it % DEPTH) * 2048 + lane * 64);
            const uint32_t rnd = lcg(sl);
#pragma unroll
            for (int k = 0; k < 5; ++k) acc += cellw[(rnd >> (4 * k)) & 15];
            __syncwarp();
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else if constexpr (VARIANT == MULUT_GB_BULK_CELL256 || VARIANT == MULUT_GB_BULK_ROWS64X3) {
        // the x4 cell fetch through the TMA unit instead of the L1TEX request path: every lane bulk-copies
        // (cp.async.bulk -> UBLKCP) its random 256-B cell - whole, or as the three 64-B row-blocks K1e reads -
        // into a per-warp ring in shared memory, one mbarrier per warp and stage counts the bytes; the lanes
        // then read 3 x 16 B of their cell back with LDS.128
        constexpr int DEPTH = 2;
        const int warp = threadIdx.x >> 5;
        uint8_t *ring = gb_smem + (size_t)warp * DEPTH * 8192;
        uint64_t *bars = reinterpret_cast<uint64_t *>(gb_smem + (size_t)(blockDim.x >> 5) * DEPTH * 8192) + warp * DEPTH;
        if (lane == 0) {
            for (int d = 0; d < DEPTH; ++d) mbar_init(smem_u32(bars + d), 1);
            mbar_fence_init();
        }
        __syncwarp();
        uint32_t sl = gtid * 40503u + 99u;
        constexpr uint32_t per_lane = VARIANT == MULUT_GB_BULK_CELL256 ? 256u : 192u;
        auto issue = [&](int stage) {
            const uint32_t bar = smem_u32(bars + stage);
            if (lane == 0) mbar_expect_tx(bar, 32u * per_lane);
            __syncwarp();
            const uint32_t cell = __umulhi(lcg(sl), n_entries);
            const uint8_t *src = table + (size_t)cell * 256;
            const uint32_t dst = smem_u32(ring + stage * 8192 + lane * 256);
            if constexpr (VARIANT == MULUT_GB_BULK_CELL256) {
                bulk_g2s(dst, src, 256u, bar);
            } else {
                const uint32_t mid = 64u + 64u * ((cell >> 3) & 1u);
                bulk_g2s(dst, src, 64u, bar);
                bulk_g2s(dst + 64u, src + mid, 64u, bar);
                bulk_g2s(dst + 128u, src + 192, 64u, bar);
            }
        };
        for (int p = 0; p < DEPTH - 1; ++p) issue(p);
        for (int it = 0; it < iters; ++it) {
            issue((it + DEPTH - 1) % DEPTH);
            mbar_wait(smem_u32(bars + it % DEPTH), (uint32_t)(it / DEPTH) & 1u);
            const uint4 *c = reinterpret_cast<const uint4 *>(ring + (it % DEPTH) * 8192 + lane * 256);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const uint4 v = c[(lane + 5 * k) & (per_lane / 16 - 1) & 7];
                acc += v.x ^ v.y ^ v.z ^ v.w;
            }
            __syncwarp();
        }
        mbar_wait(smem_u32(bars + (iters + DEPTH - 2) % DEPTH), (uint32_t)((iters + DEPTH - 2) / DEPTH) & 1u);
    } else {
        const bool coop4 = VARIANT == MULUT_GB_QUAD_CELL64 || VARIANT == MULUT_GB_QUAD_CELL256_3ROWS ||
                           VARIANT == MULUT_GB_QUAD_CELL256_4SECT;
        const bool coop2 = VARIANT == MULUT_GB_PAIR_CELL64;
        const bool coop8 = VARIANT == MULUT_GB_OCT_CELL128;
        const uint32_t seed_id = coop4 ? (gtid >> 2) : coop2 ? (gtid >> 1) : coop8 ? (gtid >> 3) : gtid;
        uint32_t s = seed_id * 2654435761u + 12345u;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < GB_U; ++u) {
                const uint32_t idx = __umulhi(lcg(s), n_entries);
                if constexpr (VARIANT == MULUT_GB_LDG_U8) {
                    acc += __ldg(table + idx);
                } else if constexpr (VARIANT == MULUT_GB_LDG_U32) {
                    acc += __ldg(reinterpret_cast<const uint32_t *>(table) + idx);
                } else if constexpr (VARIANT == MULUT_GB_LDG_U128) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(table) + idx);
                    acc += v.x ^ v.y ^ v.z ^ v.w;
                } else if constexpr (VARIANT == MULUT_GB_QUAD_CELL64) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(table + (size_t)idx * 64 + (lane & 3) * 16));
                    acc += v.x ^ v.y ^ v.z ^ v.w;
                } else if constexpr (VARIANT == MULUT_GB_QUAD_CELL256_3ROWS) {
                    // K1e's shape: three of the four 64-B row-blocks of a 256-B cell, 4 lanes x LDG.128 each
                    const uint8_t *c = table + (size_t)idx * 256 + (lane & 3) * 16;
                    const uint32_t mid = 64u + 64u * ((idx >> 3) & 1u);
                    const uint4 v0 = __ldg(reinterpret_cast<const uint4 *>(c));
                    const uint4 v1 = __ldg(reinterpret_cast<const uint4 *>(c + mid));
                    const uint4 v2 = __ldg(reinterpret_cast<const uint4 *>(c + 192));
                    acc += v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w ^ v2.x ^ v2.y ^ v2.z ^ v2.w;
                } else if constexpr (VARIANT == MULUT_GB_QUAD_CELL256_4SECT) {
                    // candidate shape: the four 32-B sectors a simplex path touches, one LDG.256 per lane
                    const uint32_t q = lane & 3;
                    const uint32_t sect = q == 0 ? 0u : q == 3 ? 7u : q == 1 ? 1u + (idx % 3u) : 4u + ((idx >> 2) % 3u);
                    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
                    const uint8_t *p = table + (size_t)idx * 256 + sect * 32;
                    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                                 : "l"(p));
                    acc += r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
                } else if constexpr (VARIANT == MULUT_GB_OCT_CELL128) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(table + (size_t)idx * 128 + (lane & 7) * 16));
                    acc += v.x ^ v.y ^ v.z ^ v.w;
                } else if constexpr (VARIANT == MULUT_GB_PAIR_CELL64) {
                    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
                    const uint8_t *p = table + (size_t)idx * 64 + (lane & 1) * 32;
                    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                                 : "l"(p));
                    acc += r0 ^ r1 ^ r2 ^ r3 ^ r4 ^ r5 ^ r6 ^ r7;
                }
            }
        }
    }
    if (acc == 0x9e3779b9u) sink[gtid & 1023] = acc;   // defeat dead-code elimination
}

template <int V>
static int run_variant(const uint8_t *d_table, uint32_t n_entries, int iters, int blocks, int threads,
                       size_t smem, uint32_t *d_sink, int repeats, float *ms)
{
    if (smem > 48 * 1024)
        MULUT_CUDA(cudaFuncSetAttribute(gather_kernel<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    MULUT_CUDA(cudaEventCreate(&e0));
    MULUT_CUDA(cudaEventCreate(&e1));
    for (int w = 0; w < 2; ++w) gather_kernel<V><<<blocks, threads, smem>>>(d_table, n_entries, iters, d_sink);
    MULUT_CUDA(cudaGetLastError());
    MULUT_CUDA(cudaEventRecord(e0));
    for (int r = 0; r < repeats; ++r) gather_kernel<V><<<blocks, threads, smem>>>(d_table, n_entries, iters, d_sink);
    MULUT_CUDA(cudaEventRecord(e1));
    MULUT_CUDA(cudaEventSynchronize(e1));
    MULUT_CUDA(cudaGetLastError());
    MULUT_CUDA(cudaEventElapsedTime(ms, e0, e1));
    *ms /= repeats;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return MULUT_OK;
}

}  // namespace mulut

using namespace mulut;

extern "C" int mulut_gather_bench(int device, int variant, size_t table_bytes, int iters, int blocks_per_sm,
                                  int threads, int repeats, double *out3)
{
    if (!out3 || iters < 1 || blocks_per_sm < 1 || threads < 32 || threads > 1024 || (threads & 31) || repeats < 1 ||
        table_bytes < 256) {
        set_error("mulut_gather_bench: bad argument");
        return MULUT_E_BAD_ARG;
    }
    MULUT_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    MULUT_CUDA(cudaGetDeviceProperties(&prop, device));
    const int blocks = blocks_per_sm * prop.multiProcessorCount;
    table_bytes = (table_bytes + 127) / 128 * 128;
    uint8_t *d_table = nullptr;
    uint32_t *d_sink = nullptr;
    MULUT_CUDA(cudaMalloc(&d_table, table_bytes));
    MULUT_CUDA(cudaMalloc(&d_sink, 1024 * sizeof(uint32_t)));
    MULUT_CUDA(cudaMemset(d_table, 0x5a, table_bytes));
    float ms = 0.f;
    int rc = MULUT_OK;
    double gathers_per_thread_iter = GB_U, bytes_per_gather = 0;
    switch (variant) {
    case MULUT_GB_LDG_U8:
        rc = run_variant<MULUT_GB_LDG_U8>(d_table, (uint32_t)table_bytes, iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 1; break;
    case MULUT_GB_LDG_U32:
        rc = run_variant<MULUT_GB_LDG_U32>(d_table, (uint32_t)(table_bytes / 4), iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 4; break;
    case MULUT_GB_LDG_U128:
        rc = run_variant<MULUT_GB_LDG_U128>(d_table, (uint32_t)(table_bytes / 16), iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 16; break;
    case MULUT_GB_QUAD_CELL64:
        rc = run_variant<MULUT_GB_QUAD_CELL64>(d_table, (uint32_t)(table_bytes / 64), iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 64; gathers_per_thread_iter = GB_U / 4.0; break;
    case MULUT_GB_PAIR_CELL64:
        rc = run_variant<MULUT_GB_PAIR_CELL64>(d_table, (uint32_t)(table_bytes / 64), iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 64; gathers_per_thread_iter = GB_U / 2.0; break;
    case MULUT_GB_OCT_CELL128:
        rc = run_variant<MULUT_GB_OCT_CELL128>(d_table, (uint32_t)(table_bytes / 128), iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 128; gathers_per_thread_iter = GB_U / 8.0; break;
    case MULUT_GB_QUAD_CELL256_3ROWS:
        rc = run_variant<MULUT_GB_QUAD_CELL256_3ROWS>(d_table, (uint32_t)(table_bytes / 256), iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 192; gathers_per_thread_iter = GB_U / 4.0;
        break;
    case MULUT_GB_QUAD_CELL256_4SECT:
        rc = run_variant<MULUT_GB_QUAD_CELL256_4SECT>(d_table, (uint32_t)(table_bytes / 256), iters, blocks, threads, 0, d_sink, repeats, &ms);
        bytes_per_gather = 128; gathers_per_thread_iter = GB_U / 4.0;
        break;
    case MULUT_GB_LDS_U8:
    case MULUT_GB_LDS_U32: {
        if (table_bytes > 200 * 1024) { set_error("shared-memory table must be <= 200 KB"); rc = MULUT_E_BAD_ARG; break; }
        if (variant == MULUT_GB_LDS_U8)
            rc = run_variant<MULUT_GB_LDS_U8>(d_table, (uint32_t)table_bytes, iters, blocks, threads, table_bytes, d_sink, repeats, &ms);
        else
            rc = run_variant<MULUT_GB_LDS_U32>(d_table, (uint32_t)(table_bytes / 4), iters, blocks, threads, table_bytes, d_sink, repeats, &ms);
        bytes_per_gather = variant == MULUT_GB_LDS_U8 ? 1 : 4;
        break;
    }
    case MULUT_GB_CPASYNC_CELL64: {
        const size_t smem = (size_t)(threads / 32) * 4 * 2048;
        if (smem > 200 * 1024) { set_error("too many threads for the cp.async ring"); rc = MULUT_E_BAD_ARG; break; }
        rc = run_variant<MULUT_GB_CPASYNC_CELL64>(d_table, (uint32_t)(table_bytes / 64), iters, blocks, threads, smem, d_sink, repeats, &ms);
        bytes_per_gather = 64; gathers_per_thread_iter = 1.0; break;   // one staged cell per lane per iteration
    }
    case MULUT_GB_BULK_CELL256:
    case MULUT_GB_BULK_ROWS64X3: {
        const size_t smem = (size_t)(threads / 32) * 2 * 8192 + (size_t)(threads / 32) * 2 * 8;
        if (smem > 200 * 1024) { set_error("too many threads for the bulk-copy ring"); rc = MULUT_E_BAD_ARG; break; }
        if (variant == MULUT_GB_BULK_CELL256)
            rc = run_variant<MULUT_GB_BULK_CELL256>(d_table, (uint32_t)(table_bytes / 256), iters, blocks, threads, smem, d_sink, repeats, &ms);
        else
            rc = run_variant<MULUT_GB_BULK_ROWS64X3>(d_table, (uint32_t)(table_bytes / 256), iters, blocks, threads, smem, d_sink, repeats, &ms);
        bytes_per_gather = variant == MULUT_GB_BULK_CELL256 ? 256 : 192; gathers_per_thread_iter = 1.0; break;   // one cell per lane per iteration
    }
    default:
        set_error("mulut_gather_bench: unknown variant %d", variant);
        rc = MULUT_E_BAD_ARG;
    }
    cudaFree(d_table);
    cudaFree(d_sink);
    if (rc) return rc;
    const double sec = ms * 1e-3;
    const double gathers = (double)blocks * threads * iters * gathers_per_thread_iter;
    out3[0] = gathers / sec;
    out3[1] = gathers * bytes_per_gather / sec;
    out3[2] = sec;
    return MULUT_OK;
}
