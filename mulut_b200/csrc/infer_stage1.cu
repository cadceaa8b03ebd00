// K1g: non-last stage (up = 1), shared-memory LUT, TMA tile ring.
//
// Same decomposition as K1a (infer_tiled.cu): every CTA owns ONE sampling mode for the
// whole launch, keeps that mode's 83 521-byte int8 LUT in shared memory (one TMA bulk
// copy) and emits the int16 partial sum of its four rotations per sample; K1b
// (combine_kernel) adds the modes and applies the stage epilogue.  What changed:
//
//   * the halo'd input tiles arrive through a ring of TMA tensor-tile loads
//     (cp.async.bulk.tensor.3d + mbarrier) issued two tiles ahead, so the fill costs no
//     instructions and its latency hides behind the previous tiles' arithmetic; the
//     hardware zero-fills outside the frame and the replicate border is patched in
//     shared memory on the ~9 % of tiles that touch an edge;
//   * leaner integer form: sort keys are f<<28 | stride (one shift-or per tap, the
//     fraction needs no masking), the vertex chain is v0, v0+s1, v0+s1+s2, v4-s4, v4 with
//     v4 = v0 + 5220 (all four taps incremented) - three masks instead of four.
//
// TMA needs 16-byte aligned frames with a row pitch that is a multiple of 16: other frames are
// first copied into a pitched staging buffer (capi.cu: tma_view).
// Arithmetic follows SURVEY.md 8-SPEC (sr/4_test_lut.py:14-237, :279-306).
#include "common.cuh"
#include "infer.cuh"
#include "tma.cuh"

namespace mulut {

constexpr int G1_TW = 96;                       // tile width, byte columns
constexpr int G1_TH = 32;
constexpr int G1_HX = 16;                       // box column of the tile's first sample (16-B aligned box start)
constexpr int G1_BOXW = 128;
constexpr int G1_BOXH = G1_TH + 4;
constexpr int G1_SLOT = G1_BOXW * G1_BOXH;      // 4608 B
constexpr int G1_RING = 4;
constexpr int G1_AHEAD = 2;
constexpr int G1_THREADS = 384;                 // 96 columns x 4 row phases
constexpr int G1_RW = G1_THREADS / G1_TW;
constexpr int G1_LUT = 83584;                   // 17^4 bytes padded to 16 B
constexpr size_t G1_SMEM = (size_t)G1_RING * G1_SLOT + G1_LUT;

__host__ __device__ constexpr int g1_tap_off(char mode, int r, int k, bool want_dy)
{
    int dy = mode == 's' ? (k >> 1) : mode == 'd' ? 2 * (k >> 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 1 : 2);
    int dx = mode == 's' ? (k & 1) : mode == 'd' ? 2 * (k & 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 2 : 1);
    for (int i = 0; i < r; ++i) { int t = dy; dy = dx; dx = -t; }
    return want_dy ? dy : dx;
}

// the four rotations of one mode for one sample; sp -> the sample's byte in the tile
template <char MODE, int CT>
__device__ __forceinline__ int g1_sample(const uint8_t *__restrict__ sp, const int8_t *__restrict__ slut)
{
    constexpr int P = G1_BOXW;
    constexpr uint32_t SA = 4913u, SB = 289u, SC = 17u, SD = 1u, KM = 0x0FFFFFFFu;
    const uint32_t t0 = sp[0];
    const uint32_t k0c = (t0 << 28) | SA;
    const uint32_t va = (t0 >> 4) * SA;
    int acc = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t1 = sp[g1_tap_off(MODE, r, 1, true) * P + g1_tap_off(MODE, r, 1, false) * CT];
        const uint32_t t2 = sp[g1_tap_off(MODE, r, 2, true) * P + g1_tap_off(MODE, r, 2, false) * CT];
        const uint32_t t3 = sp[g1_tap_off(MODE, r, 3, true) * P + g1_tap_off(MODE, r, 3, false) * CT];
        const uint32_t v0 = va + (t1 >> 4) * SB + (t2 >> 4) * SC + (t3 >> 4);
        uint32_t k0 = k0c, k1 = (t1 << 28) | SB, k2 = (t2 << 28) | SC, k3 = (t3 << 28) | SD;
        sort4_desc(k0, k1, k2, k3);
        const int f1 = k0 >> 28, f2 = k1 >> 28, f3 = k2 >> 28, f4 = k3 >> 28;
        const uint32_t v1 = v0 + (k0 & KM), v2 = v1 + (k1 & KM);
        const uint32_t v4 = v0 + (SA + SB + SC + SD), v3 = v4 - (k3 & KM);
        acc += (16 - f1) * (int)slut[v0];
        acc += (f1 - f2) * (int)slut[v1];
        acc += (f2 - f3) * (int)slut[v2];
        acc += (f3 - f4) * (int)slut[v3];
        acc += f4 * (int)slut[v4];
    }
    return acc;
}

struct Stage1Args {
    int16_t *partial;                        // n_modes planes of N*H*W*C int16
    int N, H, W, C;
    int n_modes;
    int ctas_per_mode;
    char modes[MULUT_MAX_MODES];
    const uint8_t *lut_pad[MULUT_MAX_MODES]; // int8 tables padded to G1_LUT bytes
};

template <int CT>
__global__ void __launch_bounds__(G1_THREADS, 2)
stage_smem_tma_kernel(const __grid_constant__ Stage1Args a, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) uint8_t g1_smem[];
    uint8_t *s_ring = g1_smem;
    const int8_t *slut = reinterpret_cast<const int8_t *>(g1_smem + G1_RING * G1_SLOT);
    __shared__ __align__(8) uint64_t s_full[G1_RING];
    __shared__ __align__(8) uint64_t s_lutbar;
    __shared__ int4 s_coord[G1_RING];            // frame, y0, X0 of the tile in each ring slot

    const int tid = threadIdx.x;
    const int WC = a.W * CT;
    const int m = blockIdx.x % a.n_modes;
    const int me = blockIdx.x / a.n_modes;
    const char mode = a.modes[m];
    const int tiles_x = (WC + G1_TW - 1) / G1_TW;
    const int tiles_y = (a.H + G1_TH - 1) / G1_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;
    if (me >= n_tiles) return;
    const int n_my = (int)((n_tiles - me + a.ctas_per_mode - 1) / a.ctas_per_mode);

    if (tid == 0) {
        for (int i = 0; i < G1_RING; ++i) mbar_init(smem_u32(&s_full[i]), 1);
        mbar_init(smem_u32(&s_lutbar), 1);
        mbar_fence_init();
    }
    __syncthreads();

    // tile -> (frame, y0, X0): computed once per tile by the issuing thread and published with the
    // tile itself (the mbarrier's release/acquire orders it), not divided out by all 384 threads
    auto issue = [&](int i) {                              // thread 0 only
        const unsigned tile = (unsigned)(me + (long long)i * a.ctas_per_mode);      // n_tiles < 2^31 (checked on the host)
        const unsigned tr = tile / (unsigned)tiles_x;
        const int X0 = (int)(tile - tr * (unsigned)tiles_x) * G1_TW;
        const int n = (int)(tr / (unsigned)tiles_y);
        const int y0 = (int)(tr - (unsigned)n * (unsigned)tiles_y) * G1_TH;
        const int slot = i % G1_RING;
        s_coord[slot] = make_int4(n, y0, X0, 0);
        const uint32_t bar = smem_u32(&s_full[slot]);
        mbar_expect_tx(bar, G1_SLOT);
        tma_load_3d(smem_u32(s_ring + slot * G1_SLOT), &tmap, X0 - G1_HX, y0 - 2, n, bar);
    };
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        const uint32_t lb = smem_u32(&s_lutbar);
        mbar_expect_tx(lb, G1_LUT);
        bulk_g2s(smem_u32(slut), a.lut_pad[m], G1_LUT, lb);
        for (int i = 0; i < G1_AHEAD && i < n_my; ++i) issue(i);
    }

    const int rp = tid / G1_TW, lx = tid - rp * G1_TW;
    int16_t *__restrict__ plane = a.partial + (size_t)m * a.N * a.H * WC;

    for (int i = 0; i < n_my; ++i) {
        const int slot = i % G1_RING;
        mbar_wait(smem_u32(&s_full[slot]), (uint32_t)(i / G1_RING) & 1u);
        const int4 tc = s_coord[slot];
        const int n = tc.x, y0 = tc.y, X0 = tc.z;
        uint8_t *tile = s_ring + slot * G1_SLOT;
        const bool border = (y0 < 2) || (y0 + G1_TH + 2 > a.H) || (X0 < 2 * CT) || (X0 + G1_TW + 2 * CT > WC);
        if (border) {
            // replicate padding: copy the clamped in-frame cell over every zero-filled out-of-frame cell
            for (int idx = tid; idx < G1_SLOT; idx += G1_THREADS) {
                const int r = idx / G1_BOXW, j = idx - r * G1_BOXW;
                const int gy = y0 - 2 + r, gx = X0 - G1_HX + j;
                const int cy = clampi(gy, 0, a.H - 1);
                int cx = gx;
                if (gx < 0) cx = (gx + G1_HX * CT) % CT;
                else if (gx >= WC) cx = WC - CT + (gx % CT);
                if (cy != gy || cx != gx) {
                    const int j2 = cx - X0 + G1_HX;
                    if (j2 >= 0 && j2 < G1_BOXW) tile[r * G1_BOXW + j] = tile[(cy - y0 + 2) * G1_BOXW + j2];
                }
            }
        }
        if (i == 0) mbar_wait(smem_u32(&s_lutbar), 0u);
        // one barrier per tile: publishes the border patch and proves every thread is done with
        // the tile two steps back, whose ring slot the next TMA load overwrites
        __syncthreads();
        if (tid == 0 && i + G1_AHEAD < n_my) issue(i + G1_AHEAD);

        const int xb = X0 + lx;
        const uint8_t *sp0 = tile + (rp + 2) * G1_BOXW + lx + G1_HX;
        if (xb < WC) {
            int16_t *__restrict__ op = plane + ((size_t)n * a.H + y0 + rp) * WC + xb;
            const int rows = min(G1_TH, a.H - y0);
#define G1_ROWS(MODE)                                                                        \
    _Pragma("unroll 2") for (int ly = rp; ly < rows; ly += G1_RW) {                          \
        const int acc = g1_sample<MODE, CT>(sp0 + (ly - rp) * G1_BOXW, slut);               \
        op[(size_t)(ly - rp) * WC] = (int16_t)acc;                                           \
    }
            switch (mode) {
            case 's': G1_ROWS('s') break;
            case 'd': G1_ROWS('d') break;
            default: G1_ROWS('y') break;
            }
#undef G1_ROWS
        }
    }
}

// ---------------------------------------------------------------------------
// K1h: the same stage with a-PAIRED table entries.
//
// Tap a of every interpolation is the sample itself, and simplex interpolation is a sum over
// the 16 thresholds th = 0..15 of LUT[base + sum_i [f_i > th] * stride_i].  Sorting only the
// OTHER three fractions g1 >= g2 >= g3 gives the 3-D vertex chain u0..u3 (u_j holds for th in
// [g_{j+1}, g_j), g0 = 16, g4 = 0); inside that interval the a-step is taken for th < fa.  With
// c_j = min(g_j, fa) vertex (u_j, a not stepped) weighs (g_j - g_{j+1}) - (c_j - c_{j+1}) and
// (u_j, a stepped) weighs c_j - c_{j+1}: the same integers as the reference's sorted-weight
// form (zero-width intervals are the ties).  The table stores both as one 16-bit entry
// pair[u] = LUT[u] | LUT[u + 17^3] << 8, so an interpolation is FOUR 16-bit gathers and four
// dp4a instead of five byte gathers, a 3-key sort instead of a 4-key one, and the packed weights
// alpha | beta << 8 fall out of Q_j = g_j + 255 c_j as Q_j - Q_{j+1}.
//
// 157 216 bytes of table: one CTA of 768 threads per SM, run as TWO independent groups of 384
// threads (own tile ring, own named barrier) so one group's per-tile barrier never idles the SM.
// ---------------------------------------------------------------------------
constexpr int G1P_GROUPS = 2;
constexpr int G1P_THREADS = G1_THREADS * G1P_GROUPS;
constexpr int G1P_ENTRIES = 78608;              // u <= 15*4913 + 4605 + 307
constexpr int G1P_LUT = G1P_ENTRIES * 2;        // 157 216 B (a multiple of 16)
constexpr size_t G1P_SMEM = (size_t)G1P_GROUPS * G1_RING * G1_SLOT + G1P_LUT;

size_t stage1_pair_bytes() { return G1P_LUT; }

__global__ void build_pair_table_kernel(const int8_t *__restrict__ lut, uint16_t *__restrict__ pair)
{
    for (int u = blockIdx.x * blockDim.x + threadIdx.x; u < G1P_ENTRIES; u += gridDim.x * blockDim.x)
        pair[u] = (uint16_t)((uint8_t)lut[u] | ((uint32_t)(uint8_t)lut[u + 4913] << 8));
}

int build_pair_table(const int8_t *d_lut, uint8_t *d_pair, cudaStream_t stream)
{
    build_pair_table_kernel<<<128, 256, 0, stream>>>(d_lut, reinterpret_cast<uint16_t *>(d_pair));
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

__device__ __forceinline__ void sort3_desc(uint32_t &a, uint32_t &b, uint32_t &c)
{
    uint32_t t;
    t = max(a, b); b = min(a, b); a = t;
    t = max(b, c); c = min(b, c); b = t;
    t = max(a, b); b = min(a, b); a = t;
}

template <char MODE, int CT>
__device__ __forceinline__ int g1p_sample(const uint8_t *__restrict__ sp, const uint8_t *__restrict__ spair)
{
    constexpr int P = G1_BOXW;
    constexpr uint32_t SA = 2u * 4913u, SB = 2u * 289u, SC = 2u * 17u, SD = 2u, KM = 0x0FFFFFFFu;   // byte strides
    const uint32_t t0 = sp[0];
    const uint32_t fa = t0 & 15u;
    const uint32_t va = (t0 >> 4) * SA;
    const uint32_t q0 = 16u + 255u * fa;
    int acc = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t1 = sp[g1_tap_off(MODE, r, 1, true) * P + g1_tap_off(MODE, r, 1, false) * CT];
        const uint32_t t2 = sp[g1_tap_off(MODE, r, 2, true) * P + g1_tap_off(MODE, r, 2, false) * CT];
        const uint32_t t3 = sp[g1_tap_off(MODE, r, 3, true) * P + g1_tap_off(MODE, r, 3, false) * CT];
        const uint32_t u0 = va + (t1 >> 4) * SB + (t2 >> 4) * SC + (t3 >> 4) * SD;
        uint32_t k1 = (t1 << 28) | SB, k2 = (t2 << 28) | SC, k3 = (t3 << 28) | SD;
        sort3_desc(k1, k2, k3);
        const uint32_t g1 = k1 >> 28, g2 = k2 >> 28, g3 = k3 >> 28;
        const uint32_t u1 = u0 + (k1 & KM);
        const uint32_t u3 = u0 + (SB + SC + SD), u2 = u3 - (k3 & KM);
        const uint32_t x0 = *reinterpret_cast<const uint16_t *>(spair + u0);
        const uint32_t x1 = *reinterpret_cast<const uint16_t *>(spair + u1);
        const uint32_t x2 = *reinterpret_cast<const uint16_t *>(spair + u2);
        const uint32_t x3 = *reinterpret_cast<const uint16_t *>(spair + u3);
        const uint32_t q1 = g1 + 255u * min(g1, fa), q2 = g2 + 255u * min(g2, fa), q3 = g3 + 255u * min(g3, fa);
        acc = __dp4a((int)x0, (int)(q0 - q1), acc);
        acc = __dp4a((int)x1, (int)(q1 - q2), acc);
        acc = __dp4a((int)x2, (int)(q2 - q3), acc);
        acc = __dp4a((int)x3, (int)q3, acc);
    }
    return acc;
}

template <int CT>
__global__ void __launch_bounds__(G1P_THREADS, 1)
stage_pair_tma_kernel(const __grid_constant__ Stage1Args a, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) uint8_t g1_smem[];
    const uint8_t *spair = g1_smem + G1P_GROUPS * G1_RING * G1_SLOT;
    __shared__ __align__(8) uint64_t s_full[G1P_GROUPS][G1_RING];
    __shared__ __align__(8) uint64_t s_lutbar;
    __shared__ int4 s_coord[G1P_GROUPS][G1_RING];

    const int grp = threadIdx.x / G1_THREADS, tid = threadIdx.x - grp * G1_THREADS;
    uint8_t *s_ring = g1_smem + grp * G1_RING * G1_SLOT;
    const int WC = a.W * CT;
    const int m = blockIdx.x % a.n_modes;
    const int me = (blockIdx.x / a.n_modes) * G1P_GROUPS + grp;         // my tile stream among the mode's streams
    const int streams = a.ctas_per_mode * G1P_GROUPS;
    const char mode = a.modes[m];
    const int tiles_x = (WC + G1_TW - 1) / G1_TW;
    const int tiles_y = (a.H + G1_TH - 1) / G1_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;
    const int n_my = me < n_tiles ? (int)((n_tiles - me + streams - 1) / streams) : 0;

    if (threadIdx.x == 0) {
        for (int g = 0; g < G1P_GROUPS; ++g)
            for (int i = 0; i < G1_RING; ++i) mbar_init(smem_u32(&s_full[g][i]), 1);
        mbar_init(smem_u32(&s_lutbar), 1);
        mbar_fence_init();
    }
    __syncthreads();

    auto issue = [&](int i) {                              // the group's thread 0 only
        const unsigned tile = (unsigned)(me + (long long)i * streams);
        const unsigned tr = tile / (unsigned)tiles_x;
        const int X0 = (int)(tile - tr * (unsigned)tiles_x) * G1_TW;
        const int n = (int)(tr / (unsigned)tiles_y);
        const int y0 = (int)(tr - (unsigned)n * (unsigned)tiles_y) * G1_TH;
        const int slot = i % G1_RING;
        s_coord[grp][slot] = make_int4(n, y0, X0, 0);
        const uint32_t bar = smem_u32(&s_full[grp][slot]);
        mbar_expect_tx(bar, G1_SLOT);
        tma_load_3d(smem_u32(s_ring + slot * G1_SLOT), &tmap, X0 - G1_HX, y0 - 2, n, bar);
    };
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        const uint32_t lb = smem_u32(&s_lutbar);
        mbar_expect_tx(lb, G1P_LUT);
        bulk_g2s(smem_u32(spair), a.lut_pad[m], G1P_LUT, lb);
    }
    if (tid == 0)
        for (int i = 0; i < G1_AHEAD && i < n_my; ++i) issue(i);

    const int rp = tid / G1_TW, lx = tid - rp * G1_TW;
    int16_t *__restrict__ plane = a.partial + (size_t)m * a.N * a.H * WC;
    mbar_wait(smem_u32(&s_lutbar), 0u);                    // every thread, also of a group without tiles

    for (int i = 0; i < n_my; ++i) {
        const int slot = i % G1_RING;
        mbar_wait(smem_u32(&s_full[grp][slot]), (uint32_t)(i / G1_RING) & 1u);
        const int4 tc = s_coord[grp][slot];
        const int n = tc.x, y0 = tc.y, X0 = tc.z;
        uint8_t *tile = s_ring + slot * G1_SLOT;
        const bool border = (y0 < 2) || (y0 + G1_TH + 2 > a.H) || (X0 < 2 * CT) || (X0 + G1_TW + 2 * CT > WC);
        if (border) {
            for (int idx = tid; idx < G1_SLOT; idx += G1_THREADS) {
                const int r = idx / G1_BOXW, j = idx - r * G1_BOXW;
                const int gy = y0 - 2 + r, gx = X0 - G1_HX + j;
                const int cy = clampi(gy, 0, a.H - 1);
                int cx = gx;
                if (gx < 0) cx = (gx + G1_HX * CT) % CT;
                else if (gx >= WC) cx = WC - CT + (gx % CT);
                if (cy != gy || cx != gx) {
                    const int j2 = cx - X0 + G1_HX;
                    if (j2 >= 0 && j2 < G1_BOXW) tile[r * G1_BOXW + j] = tile[(cy - y0 + 2) * G1_BOXW + j2];
                }
            }
        }
        // one barrier per tile and GROUP (named barrier 1 + grp, 384 threads): the other group keeps issuing
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(G1_THREADS) : "memory");
        if (tid == 0 && i + G1_AHEAD < n_my) issue(i + G1_AHEAD);

        const int xb = X0 + lx;
        const uint8_t *sp0 = tile + (rp + 2) * G1_BOXW + lx + G1_HX;
        if (xb < WC) {
            int16_t *__restrict__ op = plane + ((size_t)n * a.H + y0 + rp) * WC + xb;
            const int rows = min(G1_TH, a.H - y0);
#define G1P_ROWS(MODE)                                                                       \
    _Pragma("unroll 2") for (int ly = rp; ly < rows; ly += G1_RW) {                          \
        const int acc = g1p_sample<MODE, CT>(sp0 + (ly - rp) * G1_BOXW, spair);             \
        op[(size_t)(ly - rp) * WC] = (int16_t)acc;                                           \
    }
            switch (mode) {
            case 's': G1P_ROWS('s') break;
            case 'd': G1P_ROWS('d') break;
            default: G1P_ROWS('y') break;
            }
#undef G1P_ROWS
        }
    }
}

bool stage1_tma_supported(const StageArgs &a, int up)
{
    return up == 1 && a.interval == 4 && a.n_modes >= 1 && a.C >= 1 && a.C <= 4 && a.lut_alt[0] != nullptr &&
           a.in_tma != nullptr;
}

template <int CT>
static int launch_stage1_pair_t(Stage1Args &s, const CUtensorMap &tmap, int num_sms, long long n_tiles, cudaStream_t stream)
{
    MULUT_CUDA(cudaFuncSetAttribute(stage_pair_tma_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)G1P_SMEM));
    int ctas_per_mode = num_sms / s.n_modes;
    if (ctas_per_mode < 1) ctas_per_mode = 1;
    const long long need = (n_tiles + G1P_GROUPS - 1) / G1P_GROUPS;
    if (ctas_per_mode > need) ctas_per_mode = (int)need;
    s.ctas_per_mode = ctas_per_mode;
    stage_pair_tma_kernel<CT><<<ctas_per_mode * s.n_modes, G1P_THREADS, G1P_SMEM, stream>>>(s, tmap);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

template <int CT>
static int launch_stage1_t(Stage1Args &s, const CUtensorMap &tmap, int num_sms, long long n_tiles, cudaStream_t stream)
{
    // per device and cheap: set on every launch (one process may own several devices)
    MULUT_CUDA(cudaFuncSetAttribute(stage_smem_tma_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)G1_SMEM));
    int per_sm = 0;
    MULUT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_smem_tma_kernel<CT>, G1_THREADS, G1_SMEM));
    if (per_sm < 1) per_sm = 1;
    int ctas_per_mode = per_sm * num_sms / s.n_modes;
    if (ctas_per_mode < 1) ctas_per_mode = 1;
    if (ctas_per_mode > n_tiles) ctas_per_mode = (int)n_tiles;
    s.ctas_per_mode = ctas_per_mode;
    stage_smem_tma_kernel<CT><<<ctas_per_mode * s.n_modes, G1_THREADS, G1_SMEM, stream>>>(s, tmap);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

// Returns MULUT_OK, an error (< 0) or +1 (not applicable: caller runs K1a).
int launch_stage1_tma(const StageArgs &a, int16_t *partial, cudaStream_t stream)
{
    CUtensorMap tmap;
    if (!a.in_tma || tma_encode_frames(&tmap, a.in_tma, a.N, a.H, a.W * a.C, a.in_pitch, G1_BOXW, G1_BOXH) != 0) return 1;
    Stage1Args s;
    memset(&s, 0, sizeof s);
    s.partial = partial; s.N = a.N; s.H = a.H; s.W = a.W; s.C = a.C; s.n_modes = a.n_modes;
    for (int m = 0; m < a.n_modes; ++m) { s.modes[m] = a.modes[m]; s.lut_pad[m] = a.lut_alt[m]; }
    const int WC = a.W * a.C;
    const long long n_tiles = (long long)a.N * ((a.H + G1_TH - 1) / G1_TH) * ((WC + G1_TW - 1) / G1_TW);
    if (n_tiles >= 0x7fffffffLL) return 1;                 // 32-bit tile arithmetic in the kernel
    static const bool pair = [] { const char *e = getenv("MULUT_K1_PAIR"); return !e || e[0] != '0'; }();
    if (pair) {                                            // K1h: the pair table sits behind the padded byte table
        for (int m = 0; m < a.n_modes; ++m) s.lut_pad[m] = a.lut_alt[m] + G1_LUT;
        return a.C == 3 ? launch_stage1_pair_t<3>(s, tmap, a.num_sms, n_tiles, stream)
             : a.C == 1 ? launch_stage1_pair_t<1>(s, tmap, a.num_sms, n_tiles, stream)
             : a.C == 4 ? launch_stage1_pair_t<4>(s, tmap, a.num_sms, n_tiles, stream)
                        : launch_stage1_pair_t<2>(s, tmap, a.num_sms, n_tiles, stream);
    }
    return a.C == 3 ? launch_stage1_t<3>(s, tmap, a.num_sms, n_tiles, stream)
         : a.C == 1 ? launch_stage1_t<1>(s, tmap, a.num_sms, n_tiles, stream)
         : a.C == 4 ? launch_stage1_t<4>(s, tmap, a.num_sms, n_tiles, stream)
                    : launch_stage1_t<2>(s, tmap, a.num_sms, n_tiles, stream);
}

}  // namespace mulut
