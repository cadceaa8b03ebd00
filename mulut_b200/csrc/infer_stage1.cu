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
#include <stdlib.h>

#include "binned.cuh"
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

// ---------------------------------------------------------------------------
// K1i: K1h with the mode combine FUSED (no int16 partial planes in HBM, no K1b launch).
//
// The three mode CTAs that walk the same tile stream exchange their int16 partial sums through a small
// ring in global memory that never leaves L2 (streams x 6 slots x modes x 6 KB = a few MB, rewritten every
// few microseconds): every CTA writes its partial tile into the slot, the tile's OWNER (rotating: tile i of
// a stream belongs to mode i % M) waits until all M partials are there, adds them, applies the stage
// epilogue (sr/4_test_lut.py:281-286,300-302) and writes the uint8 image - and, when the next stage is K1f,
// counts the 8-bin histogram of the bytes it writes; the last CTA turns it into K1f's plan, as K1b did.
//
// Hand-off per (stream, slot): two monotonic counters in global memory, `ready` (+1 by every mode once its
// partial tile is written: st.global, __threadfence, group barrier, red.release.gpu) and `freed` (+1 by the
// owner once it has consumed the slot).  Writers wait for freed >= uses so far (6 slots of slack: practically
// never), the owner - three tiles later, so that nobody ever waits for
// somebody else's epilogue - waits for ready == M * (uses + 1) with ld.acquire.gpu and reads the partials with
// ld.global.cg (L2; L1 is not coherent).  CTAs of one launch wait on each other, so the kernel is launched
// COOPERATIVELY (cudaLaunchCooperativeKernel guarantees that all CTAs are co-resident; the grid is one CTA
// per SM at most); the launcher falls back to K1h + K1b when that is not possible (stream capture, no
// cooperative launch).  A thread-block cluster with the exchange in distributed shared memory was the first
// design: cudaOccupancyMaxActiveClusters allows only 45 clusters of 3 such CTAs on this part (135 of 148
// SMs, profiles/r02_cluster_probe.txt) - a 9 % loss against the 3 % K1b costs.
// ---------------------------------------------------------------------------
constexpr int G1F_R = 6;                        // exchange slots per tile stream
constexpr int G1F_DEFER = 3;                    // the owner finishes tile j while the group is at tile j + 3
constexpr int G1F_TILE = G1_TW * G1_TH;         // samples per tile
constexpr int G1F_FLAG_WORDS = 8;               // one 32-byte sector per (stream, slot): [0] ready, [1] freed

size_t stage1_fused_ws_bytes(int num_sms)
{
    // streams x modes <= 2 x num_sms (one CTA per SM at most, two tile streams per CTA)
    const size_t slots = (size_t)2 * num_sms * G1F_R;
    return slots * G1F_TILE * sizeof(int16_t) + slots * G1F_FLAG_WORDS * sizeof(uint32_t) + 256;
}

struct Stage1FusedArgs {
    uint8_t *out;                            // (N, H, W, C) uint8: the stage output
    int N, H, W, C;
    int n_modes;
    int ctas_per_mode;
    int last;                                // epilogue form (a last stage with up = 1: scale 1)
    char modes[MULUT_MAX_MODES];
    const uint8_t *lut_pad[MULUT_MAX_MODES]; // a-paired tables
    int16_t *exch;                           // [stream][slot][mode][G1F_TILE]
    uint32_t *flags;                         // [stream][slot][G1F_FLAG_WORDS]
    uint32_t *ticket;                        // CTAs that have finished (the last one resets the flags)
    BinPlanArgs pa;                          // histogram + plan for K1f on this stage's output, or ctl = null
};

__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void red_release_gpu_add(uint32_t *p, uint32_t v)
{
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// lane 0 of every warp polls, the warp barrier carries the acquire to the other lanes
__device__ __forceinline__ void warp_wait_ge(const uint32_t *p, uint32_t target)
{
    if ((threadIdx.x & 31) == 0) {
        unsigned spins = 0;
        while ((int)(ld_acquire_gpu(p) - target) < 0) {
            __nanosleep(40);
            if (++spins > (1u << 22)) asm volatile("trap;");   // seconds: a lost partner must not hang the GPU
        }
    }
    __syncwarp();
}

template <int CT>
__global__ void __launch_bounds__(G1P_THREADS, 1)
stage_pair_fused_kernel(const __grid_constant__ Stage1FusedArgs a, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) uint8_t g1_smem[];
    const uint8_t *spair = g1_smem + G1P_GROUPS * G1_RING * G1_SLOT;
    __shared__ __align__(8) uint64_t s_full[G1P_GROUPS][G1_RING];
    __shared__ __align__(8) uint64_t s_lutbar;
    __shared__ int4 s_coord[G1P_GROUPS][G1_RING];
    __shared__ uint32_t s_hist[BN_BINS];

    const int grp = threadIdx.x / G1_THREADS, tid = threadIdx.x - grp * G1_THREADS;
    uint8_t *s_ring = g1_smem + grp * G1_RING * G1_SLOT;
    const int WC = a.W * CT;
    const int M = a.n_modes;
    const int m = blockIdx.x % M;
    const int me = (blockIdx.x / M) * G1P_GROUPS + grp;                 // my tile stream (shared with the other modes)
    const int streams = a.ctas_per_mode * G1P_GROUPS;
    const char mode = a.modes[m];
    const int tiles_x = (WC + G1_TW - 1) / G1_TW;
    const int tiles_y = (a.H + G1_TH - 1) / G1_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;
    const int n_my = me < n_tiles ? (int)((n_tiles - me + streams - 1) / streams) : 0;
    const uint32_t den = a.last ? 16u * M : 64u * M;
    const uint32_t magic = rhe_magic(den);
    const int bias = a.last ? 0 : 127 * (int)den;

    if (threadIdx.x == 0) {
        for (int g = 0; g < G1P_GROUPS; ++g)
            for (int i = 0; i < G1_RING; ++i) mbar_init(smem_u32(&s_full[g][i]), 1);
        mbar_init(smem_u32(&s_lutbar), 1);
        mbar_fence_init();
    }
    if (threadIdx.x < BN_BINS) s_hist[threadIdx.x] = 0;
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
    BinCounter bc;
    mbar_wait(smem_u32(&s_lutbar), 0u);

    // The owner's part of tile j runs G1F_DEFER iterations later, after this group has finished its own partial
    // of tile j + G1F_DEFER: by then the other modes' partials of tile j are normally long there, so the wait costs
    // nothing and - more important - no CTA's next tile ever waits for another CTA's epilogue (run right away,
    // the rotation of owners chained every tile behind the previous owner's wait + epilogue: 1.7x slower).
    for (int i = 0; i < n_my + G1F_DEFER; ++i) {
        if (i < n_my) {
        const int slot = i % G1_RING;
        mbar_wait(smem_u32(&s_full[grp][slot]), (uint32_t)(i / G1_RING) & 1u);
        const int4 tc = s_coord[grp][slot];
        const int y0 = tc.y, X0 = tc.z;
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
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(G1_THREADS) : "memory");
        if (tid == 0 && i + G1_AHEAD < n_my) issue(i + G1_AHEAD);

        // ---- my mode's partial sums of this tile -> exchange slot ----
        const int xs = i % G1F_R;                          // exchange slot and its use count so far
        const uint32_t use = (uint32_t)(i / G1F_R);
        const size_t ring_idx = (size_t)me * G1F_R + xs;
        uint32_t *fl = a.flags + ring_idx * G1F_FLAG_WORDS;
        int16_t *__restrict__ xtile = a.exch + ring_idx * (size_t)M * G1F_TILE;
        if (use > 0) warp_wait_ge(fl + 1, use);            // the slot's previous tile has been consumed (long ago)
        const int xb = X0 + lx;
        const uint8_t *sp0 = tile + (rp + 2) * G1_BOXW + lx + G1_HX;
        const int rows = min(G1_TH, a.H - y0);
        if (xb < WC) {
            int16_t *__restrict__ op = xtile + (size_t)m * G1F_TILE + rp * G1_TW + lx;
#define G1F_ROWS(MODE)                                                                       \
    _Pragma("unroll 2") for (int ly = rp; ly < rows; ly += G1_RW) {                          \
        const int acc = g1p_sample<MODE, CT>(sp0 + (ly - rp) * G1_BOXW, spair);             \
        op[(ly - rp) * G1_TW] = (int16_t)acc;                                                \
    }
            switch (mode) {
            case 's': G1F_ROWS('s') break;
            case 'd': G1F_ROWS('d') break;
            default: G1F_ROWS('y') break;
            }
#undef G1F_ROWS
        }
        // release: the group's stores happen-before thread 0's red.release through the group barrier (cumulativity)
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(G1_THREADS) : "memory");
        if (tid == 0) red_release_gpu_add(fl, 1u);
        }

        // ---- owner of tile j = i - G1F_DEFER: all M partials -> stage epilogue -> uint8 image (+ K1f's histogram) ----
        const int j = i - G1F_DEFER;
        if (j < 0 || j >= n_my || (j % M) != m) continue;
        const uint32_t use = (uint32_t)(j / G1F_R);
        const size_t ring_idx = (size_t)me * G1F_R + (j % G1F_R);
        uint32_t *fl = a.flags + ring_idx * G1F_FLAG_WORDS;
        const int16_t *__restrict__ xtile = a.exch + ring_idx * (size_t)M * G1F_TILE;
        const unsigned tile_j = (unsigned)(me + (long long)j * streams);
        const unsigned tr = tile_j / (unsigned)tiles_x;
        const int X0 = (int)(tile_j - tr * (unsigned)tiles_x) * G1_TW;
        const int n = (int)(tr / (unsigned)tiles_y);
        const int y0 = (int)(tr - (unsigned)n * (unsigned)tiles_y) * G1_TH;
        const int rows = min(G1_TH, a.H - y0);
        warp_wait_ge(fl, (uint32_t)M * (use + 1u));
        // epilogue mapping: thread -> 8 consecutive samples of one row (16-byte loads of the partials, one 8-byte store)
        const int er = tid / (G1_TW / 8), ec = (tid - er * (G1_TW / 8)) * 8;
        if (er < rows && X0 + ec < WC) {
            uint8_t *__restrict__ outp = a.out + ((size_t)n * a.H + y0 + er) * WC + X0 + ec;
            const int16_t *__restrict__ xp = xtile + er * G1_TW + ec;
            const int nv = min(8, WC - (X0 + ec));
            int sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int mm = 0; mm < M; ++mm) {
                const int4 v = __ldcg(reinterpret_cast<const int4 *>(xp + (size_t)mm * G1F_TILE));
                const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    sum[2 * k] += (int)(int16_t)(w[k] & 0xffff);
                    sum[2 * k + 1] += w[k] >> 16;
                }
            }
            uint32_t lo = 0, hi = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                lo |= rhe_div_clamp_u8_magic(sum[k] + bias, den, magic) << (8 * k);
                hi |= rhe_div_clamp_u8_magic(sum[4 + k] + bias, den, magic) << (8 * k);
            }
            if (nv == 8 && (reinterpret_cast<uintptr_t>(outp) & 7) == 0) {
                *reinterpret_cast<uint2 *>(outp) = make_uint2(lo, hi);
                if (a.pa.ctl) { bc.add_word(lo); bc.add_word(hi); }
            } else {
                for (int k = 0; k < nv; ++k) {
                    const uint32_t o = ((k < 4 ? lo : hi) >> (8 * (k & 3))) & 0xFFu;
                    outp[k] = (uint8_t)o;
                    if (a.pa.ctl) bc.add_byte(o);
                }
            }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(G1_THREADS) : "memory");
        if (tid == 0) red_release_gpu_add(fl + 1, 1u);
    }

    // ---- launch epilogue: histogram -> plan (last CTA), exchange flags back to zero ----
    __syncthreads();
    if (a.pa.ctl) {
        bc.reduce_into(s_hist);
        __syncthreads();
        if (threadIdx.x < BN_BINS && s_hist[threadIdx.x])
            atomicAdd(&a.pa.ctl->hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
    }
    __threadfence();
    __syncthreads();
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {                                          // every other CTA is past its last wait
        __threadfence();
        const size_t words = (size_t)streams * G1F_R * G1F_FLAG_WORDS;
        for (size_t w = threadIdx.x; w < words; w += blockDim.x) a.flags[w] = 0u;
        if (threadIdx.x == 0) {
            *a.ticket = 0u;
            if (a.pa.ctl) bin_plan(a.pa.ctl, a.pa.n_tiles, a.pa.G, a.pa.list_cap, a.pa.allow_orphans);
        }
    }
}

template <int CT>
static int launch_stage1_fused_t(Stage1FusedArgs &s, const CUtensorMap &tmap, int num_sms, long long n_tiles, cudaStream_t stream)
{
    MULUT_CUDA(cudaFuncSetAttribute(stage_pair_fused_kernel<CT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)G1P_SMEM));
    int per_sm = 0;
    MULUT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_pair_fused_kernel<CT>, G1P_THREADS, G1P_SMEM));
    if (per_sm < 1) return 1;
    int ctas_per_mode = per_sm * num_sms / s.n_modes;      // co-resident by construction (cooperative launch checks it)
    if (ctas_per_mode > num_sms / s.n_modes) ctas_per_mode = num_sms / s.n_modes;
    if (ctas_per_mode < 1) return 1;
    const long long need = (n_tiles + G1P_GROUPS - 1) / G1P_GROUPS;
    if (ctas_per_mode > need) ctas_per_mode = (int)need;
    s.ctas_per_mode = ctas_per_mode;
    void *args[] = {(void *)&s, (void *)&tmap};
    cudaError_t e = cudaLaunchCooperativeKernel((const void *)stage_pair_fused_kernel<CT>, dim3(ctas_per_mode * s.n_modes),
                                                dim3(G1P_THREADS), args, G1P_SMEM, stream);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) { cudaGetLastError(); return 1; }
    MULUT_CUDA(e);
    return MULUT_OK;
}

bool stage1_fused_enabled()
{
    const char *ef = getenv("MULUT_K1_FUSED"), *ep = getenv("MULUT_K1_PAIR");
    return ef && ef[0] == '1' && (!ep || ep[0] != '0');
}

// K1i.  ws: stage1_fused_ws_bytes(num_sms) of zero-initialised device memory (the kernel leaves it zeroed).
// Returns MULUT_OK, an error (< 0) or +1 (not applicable here: run launch_stage1_tma + K1b).
int launch_stage1_fused(const StageArgs &a, void *ws, const BinPlanArgs *plan, cudaStream_t stream)
{
    // OPT-IN (MULUT_K1_FUSED=1, read on every call: the tests flip it).  Measured on cfg 2 (16 x 1080p per launch):
    // K1i 2.85 ms against 2.35 + 0.18 ms for K1h + K1b - the owner's epilogue (L2 round trip of the partials) and the
    // release fence sit on the critical path of a shared-memory-bound kernel, while the HBM traffic they save
    // (12 of 22 B per sample) was never the bottleneck.  K1i needs the a-paired tables.
    if (!stage1_fused_enabled() || !ws || a.n_modes > 4 || a.n_modes > a.num_sms) return 1;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return 1; }
    int dev = 0, coop = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop)
        return 1;
    CUtensorMap tmap;
    if (!a.in_tma || tma_encode_frames(&tmap, a.in_tma, a.N, a.H, a.W * a.C, a.in_pitch, G1_BOXW, G1_BOXH) != 0) return 1;
    const int WC = a.W * a.C;
    const long long n_tiles = (long long)a.N * ((a.H + G1_TH - 1) / G1_TH) * ((WC + G1_TW - 1) / G1_TW);
    if (n_tiles >= 0x7fffffffLL) return 1;
    Stage1FusedArgs s;
    memset(&s, 0, sizeof s);
    s.out = a.out; s.N = a.N; s.H = a.H; s.W = a.W; s.C = a.C; s.n_modes = a.n_modes; s.last = a.last;
    for (int m = 0; m < a.n_modes; ++m) { s.modes[m] = a.modes[m]; s.lut_pad[m] = a.lut_alt[m] + G1_LUT; }
    const size_t slots = (size_t)2 * a.num_sms * G1F_R;
    uint8_t *base = static_cast<uint8_t *>(ws);
    s.exch = reinterpret_cast<int16_t *>(base);
    s.flags = reinterpret_cast<uint32_t *>(base + slots * G1F_TILE * sizeof(int16_t));
    s.ticket = s.flags + slots * G1F_FLAG_WORDS;
    if (plan && plan->ctl) {
        s.pa = *plan;
        MULUT_CUDA(cudaMemsetAsync(s.pa.ctl, 0, sizeof(BinCtl), stream));
    }
    return a.C == 3 ? launch_stage1_fused_t<3>(s, tmap, a.num_sms, n_tiles, stream)
         : a.C == 1 ? launch_stage1_fused_t<1>(s, tmap, a.num_sms, n_tiles, stream)
         : a.C == 4 ? launch_stage1_fused_t<4>(s, tmap, a.num_sms, n_tiles, stream)
                    : launch_stage1_fused_t<2>(s, tmap, a.num_sms, n_tiles, stream);
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
