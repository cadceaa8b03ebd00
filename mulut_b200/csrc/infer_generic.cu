// Generic stage kernel: any mode list / scale 1..4 / interval 1..7 / channel count.
// One thread per input sample; the five simplex vertices are gathered
// vertex-major from the int8 LUT (reference layout) through L1/L2.
// This is the always-applicable path and the in-library cross-check for the
// tiled sm_100a kernels (infer_tiled.cu).
//
// Arithmetic = SURVEY.md 8-SPEC, restating sr/4_test_lut.py:14-237 (one pass)
// and :279-306 (sum over modes x rotations, epilogue) in integers.
#include <stdlib.h>

#include "common.cuh"
#include "infer.cuh"

namespace mulut {

template <int UP>
__device__ __forceinline__ void load_row(const int8_t *__restrict__ p, int (&v)[UP * UP])
{
    if constexpr (UP == 1) {
        v[0] = __ldg(p);
    } else if constexpr (UP == 2) {
        int w = __ldg(reinterpret_cast<const int *>(p));
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (int)(int8_t)(w >> (8 * j));
    } else if constexpr (UP == 4) {
        int4 w = __ldg(reinterpret_cast<const int4 *>(p));
        int ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (int)(int8_t)(ww[j >> 2] >> (8 * (j & 3)));
    } else {
#pragma unroll
        for (int j = 0; j < UP * UP; ++j) v[j] = __ldg(p + j);
    }
}

// K0s: rows of a table staged in SHARED memory, padded to 4-byte multiples (9 -> 12 bytes for up = 3)
__host__ __device__ constexpr int smem_row_pitch(int up) { return up == 1 ? 1 : (up * up + 3) / 4 * 4; }
template <int UP>
__device__ __forceinline__ void load_row_smem(const int8_t *p, int (&v)[UP * UP])
{
    if constexpr (UP == 1) {
        v[0] = *p;
    } else {
        constexpr int W = (UP * UP + 3) / 4;
        int ww[W];
#pragma unroll
        for (int k = 0; k < W; ++k) ww[k] = reinterpret_cast<const int *>(p)[k];
#pragma unroll
        for (int j = 0; j < UP * UP; ++j) v[j] = (int)(int8_t)(ww[j >> 2] >> (8 * (j & 3)));
    }
}

// rotated tap offset of (mode, rotation r, tap k), compile-time (same rule as common.cuh::build_tap_table)
__host__ __device__ constexpr int k0_tap_off(char mode, int r, int k, bool want_dy)
{
    int dy = mode == 's' ? (k >> 1) : mode == 'd' ? 2 * (k >> 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 1 : 2);
    int dx = mode == 's' ? (k & 1) : mode == 'd' ? 2 * (k & 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 2 : 1);
    for (int i = 0; i < r; ++i) { int t = dy; dy = dx; dx = -t; }
    return want_dy ? dy : dx;
}

// the four rotations of one mode on the sample's 5x5 neighbourhood nb[(dy+2)*5 + (dx+2)] (register-resident: every
// index is a compile-time constant), any interval, vertex rows gathered from the reference-layout table
template <char MODE, int UP, bool SMEM = false>
__device__ __forceinline__ void generic_mode(const uint32_t (&nb)[25], const int8_t *__restrict__ lut, int interval,
                                             const uint32_t (&stride)[4], int (&acc)[UP * UP])
{
    constexpr int UP2 = UP * UP;
    constexpr uint32_t SBITS = 22, SMASK = (1u << SBITS) - 1u;   // L^3 <= 129^3 < 2^22
    const int q = 1 << interval;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t key[4];
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t t = nb[(k0_tap_off(MODE, r, k, true) + 2) * 5 + k0_tap_off(MODE, r, k, false) + 2];
            v += (t >> interval) * stride[k];
            key[k] = ((t & (uint32_t)(q - 1)) << SBITS) | stride[k];
        }
        sort4_desc(key[0], key[1], key[2], key[3]);
        const int f1 = key[0] >> SBITS, f2 = key[1] >> SBITS, f3 = key[2] >> SBITS, f4 = key[3] >> SBITS;
        const int w[5] = {q - f1, f1 - f2, f2 - f3, f3 - f4, f4};
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            int vals[UP2];
            if constexpr (SMEM) load_row_smem<UP>(lut + v * smem_row_pitch(UP), vals);
            else load_row<UP>(lut + (size_t)v * UP2, vals);
#pragma unroll
            for (int j = 0; j < UP2; ++j) acc[subpixel_perm<UP>(r, j)] += w[k] * vals[j];
            if (k < 4) v += key[k] & SMASK;
        }
    }
}

// one sample i = ((n*H + y)*W + x)*C + c through all modes x rotations + epilogue
template <int UP, bool SMEM = false>
__device__ __forceinline__ void generic_sample(const StageArgs &a, size_t i, const int8_t *s_luts = nullptr,
                                               int s_table_bytes = 0)
{
    constexpr int UP2 = UP * UP;
    const int WC = a.W * a.C;
    const int q = 1 << a.interval;
    const int L = (1 << (8 - a.interval)) + 1;
    const uint32_t stride[4] = {(uint32_t)(L * L * L), (uint32_t)(L * L), (uint32_t)L, 1u};

    const int xb = (int)(i % WC);
    const size_t row = i / WC;
    const int y = (int)(row % a.H);
    const int n = (int)(row / a.H);
    const int x = xb / a.C, c = xb - x * a.C;
    const uint8_t *__restrict__ img = a.in + (size_t)n * a.H * WC;

    // the 5x5 neighbourhood with replicate padding (sr/4_test_lut.py:293-296: np.pad(..., mode='edge') of the
    // rotated image = clamped coordinates of the un-rotated one): 10 clamps and 25 loads per sample instead of
    // 96 clamps, 48 loads and 96 run-time tap-table reads
    uint32_t nb[25];
    {
        size_t yo[5];
        int xo[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) {
            yo[d] = (size_t)clampi(y + d - 2, 0, a.H - 1) * WC;
            xo[d] = clampi(x + d - 2, 0, a.W - 1) * a.C + c;
        }
#pragma unroll
        for (int dy = 0; dy < 5; ++dy)
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) nb[dy * 5 + dx] = img[yo[dy] + xo[dx]];
    }

    int acc[UP2];
#pragma unroll
    for (int j = 0; j < UP2; ++j) acc[j] = 0;

    for (int m = 0; m < a.n_modes; ++m) {
        const int8_t *__restrict__ lut = SMEM ? s_luts + (size_t)m * s_table_bytes : a.lut[m];
        switch (a.modes[m]) {                      // validated at mulut_create: only s, d, y exist
        case 's': generic_mode<'s', UP, SMEM>(nb, lut, a.interval, stride, acc); break;
        case 'd': generic_mode<'d', UP, SMEM>(nb, lut, a.interval, stride, acc); break;
        default: generic_mode<'y', UP, SMEM>(nb, lut, a.interval, stride, acc); break;
        }
    }

    // epilogue, sr/4_test_lut.py:281-286,300-306
    const uint32_t den = a.last ? (uint32_t)(q * a.n_modes) : (uint32_t)(q * a.n_modes * 4);
    const int bias = a.last ? 0 : 127 * (int)den;
    uint8_t *__restrict__ out = a.out + (size_t)n * a.H * UP * (size_t)WC * UP;
#pragma unroll
    for (int u = 0; u < UP; ++u)
#pragma unroll
        for (int vv = 0; vv < UP; ++vv) {
            const uint32_t o = rhe_div_clamp_u8(acc[u * UP + vv] + bias, den);
            out[((size_t)(y * UP + u) * (a.W * UP) + (x * UP + vv)) * a.C + c] = (uint8_t)o;
        }
}

template <int UP>
__global__ void __launch_bounds__(256) stage_generic_kernel(const __grid_constant__ StageArgs a)
{
    const size_t total = (size_t)a.N * a.H * a.W * a.C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x)
        generic_sample<UP>(a, i);
}

// K0s: the tables of intervals 5-7 are tiny (L = 9, 5, 3: 6 561, 625, 81 rows): when every mode's table of the
// stage fits one CTA's shared memory the five vertex gathers of an interpolation are LDS instead of L1 hits
template <int UP>
__global__ void __launch_bounds__(512) stage_generic_smem_kernel(const __grid_constant__ StageArgs a, int rows, int table_bytes)
{
    extern __shared__ __align__(16) int8_t k0s_luts[];
    constexpr int UP2 = UP * UP, PITCH = smem_row_pitch(UP);
    for (int m = 0; m < a.n_modes; ++m)
        for (int e = threadIdx.x; e < rows * PITCH; e += blockDim.x) {
            const int r = e / PITCH, j = e - r * PITCH;
            k0s_luts[(size_t)m * table_bytes + e] = j < UP2 ? a.lut[m][(size_t)r * UP2 + j] : (int8_t)0;
        }
    __syncthreads();
    const size_t total = (size_t)a.N * a.H * a.W * a.C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x)
        generic_sample<UP, true>(a, i, k0s_luts, table_bytes);
}

// returns MULUT_OK, an error, or +1 when the tables do not fit (K0 runs)
template <int UP>
static int launch_generic_smem_t(const StageArgs &a, cudaStream_t stream)
{
    const int L = (1 << (8 - a.interval)) + 1;
    const long long rows = (long long)L * L * L * L;
    const long long table = (rows * smem_row_pitch(UP) + 15) / 16 * 16;
    const long long smem = table * a.n_modes;
    if (smem > 100 * 1024) return 1;
    MULUT_CUDA(cudaFuncSetAttribute(stage_generic_smem_kernel<UP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    MULUT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_generic_smem_kernel<UP>, 512, (size_t)smem));
    if (per_sm < 1) return 1;
    const size_t total = (size_t)a.N * a.H * a.W * a.C;
    size_t blocks = (total + 511) / 512;
    if (blocks > (size_t)per_sm * a.num_sms) blocks = (size_t)per_sm * a.num_sms;
    stage_generic_smem_kernel<UP><<<(unsigned)blocks, 512, (size_t)smem, stream>>>(a, (int)rows, (int)table);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

// The same per-sample path over an explicit list of sample indices whose length lives
// on the device (the "orphan" samples the binned kernel K1f leaves to L2 gathers).
template <int UP>
__global__ void __launch_bounds__(256)
stage_generic_list_kernel(const __grid_constant__ StageArgs a, const uint32_t *__restrict__ list,
                          const uint32_t *__restrict__ count)
{
    const uint32_t total = *count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
        generic_sample<UP>(a, (size_t)list[i]);
}

int launch_stage_generic_list2(const StageArgs &a, const uint32_t *list, const uint32_t *count, cudaStream_t stream)
{
    stage_generic_list_kernel<2><<<(unsigned)a.num_sms * 8, 256, 0, stream>>>(a, list, count);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

int launch_stage_generic(const StageArgs &a, int up, cudaStream_t stream)
{
    const size_t total = (size_t)a.N * a.H * a.W * a.C;
    if (total == 0) return MULUT_OK;
    if (a.interval >= 5 && total >= (size_t)1 << 18 && up >= 1 && up <= 4) {      // small tables, big launch: K0s
        const char *e = getenv("MULUT_K0_SMEM");
        if (!e || e[0] != '0') {
            const int rc = up == 1 ? launch_generic_smem_t<1>(a, stream) : up == 2 ? launch_generic_smem_t<2>(a, stream)
                         : up == 3 ? launch_generic_smem_t<3>(a, stream) : launch_generic_smem_t<4>(a, stream);
            if (rc <= 0) return rc;
        }
    }
    const int threads = 256;
    size_t blocks = (total + threads - 1) / threads;
    const size_t cap = (size_t)a.num_sms * 64;      // grid-stride beyond this
    if (blocks > cap) blocks = cap;
    switch (up) {
    case 1: stage_generic_kernel<1><<<(unsigned)blocks, threads, 0, stream>>>(a); break;
    case 2: stage_generic_kernel<2><<<(unsigned)blocks, threads, 0, stream>>>(a); break;
    case 3: stage_generic_kernel<3><<<(unsigned)blocks, threads, 0, stream>>>(a); break;
    case 4: stage_generic_kernel<4><<<(unsigned)blocks, threads, 0, stream>>>(a); break;
    default: set_error("scale %d not supported (1..4)", up); return MULUT_E_BAD_ARG;
    }
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

}  // namespace mulut
