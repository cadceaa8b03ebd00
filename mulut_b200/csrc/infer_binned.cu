// K1f: last stage, up = 2, LUT slabs resident in SHARED MEMORY ("binned" kernel).
//
// Tap a of every one of a sample's 12 interpolations (3 modes x 4 rotations) is
// the sample itself (sr/4_test_lut.py:18-51: tap a has offset (0,0) in every
// mode), so all 60 vertex rows a sample can touch lie in the LUT slabs
// a in {m_a, m_a + 1}, m_a = sample >> 4.  The 17 slabs of a x2 table are 19.6 KB
// each: three of them for each of three modes fit one SM's shared memory
// (3 x 3 x 19 664 B = 177 KB).  So the kernel splits the samples by value instead
// of by position:
//
//   * bin b = sample >> 5 (8 bins); a persistent CTA serves ONE bin and keeps
//     slabs 2b .. 2b+2 of every mode in shared memory (one TMA bulk copy each);
//   * CTAs are dealt to the bins in proportion to a cost model (bin_plan, binned.cuh)
//     fed by an 8-bin histogram of the stage input.  The histogram is counted by the
//     kernel that writes that input (K1b) and its last block runs the plan, or by
//     bin_hist_kernel for single-stage models - no host round trip; bins too sparse
//     to pay for walking every tile are left to a list ("orphans") that every CTA
//     fills for its slice of the input and stage_generic_list_kernel finishes;
//   * the CTA runs as BN_GROUPS independent groups of BN_GT threads (own tile ring, queue and named
//     barrier, shared LUT slabs), so one group's scan / barrier / TMA wait hides behind the other's
//     interpolation rounds; every group draws tiles of its bin from the bin's atomic counter;
//   * the halo'd tiles arrive through
//     a ring of TMA tensor-tile loads (cp.async.bulk.tensor.3d + mbarrier, zero
//     fill outside the frame patched to replicate padding in shared memory);
//   * a scan compacts the tile's samples of this bin into a queue (one 4-sample word
//     per thread, zero-byte trick, warp prefix);
//     full rounds of BN_GT queue entries are interpolated, the remainder is
//     carried to the next tile while its ring slot is still resident;
//   * per interpolation: keys f<<28 | byte-stride are sorted by a 5-comparator
//     min/max network, the vertex chain is v0, v0+s1, v0+s1+s2, v4-s4, v4 with
//     v4 = v0 + (sum of strides), five LDS.32 fetch the biased rows (4 outputs in
//     one word) and two multiply-adds per row accumulate all four outputs:
//     even bytes in a 32-bit lane pair, the whole word in a 64-bit accumulator
//     (fields overlap; the even lanes are subtracted at the end);
//   * the 2x2 pixel-shuffle is fused into four byte stores per sample (L2 merges
//     the partial sectors written by the CTAs of different bins).
//
// Gather cost: 5 LDS.32 per interpolation at the shared-memory bank-conflict rate
// (~10 lanes/clk/SM) instead of one 64-byte L1/L2 cell per interpolation.
// Arithmetic follows SURVEY.md 8-SPEC (sr/4_test_lut.py:14-237, :279-306).
#include <stdlib.h>

#include "binned.cuh"
#include "common.cuh"
#include "infer.cuh"
#include "tma.cuh"

namespace mulut {

constexpr int BN_HX = 16;                       // box column of the tile's first sample: TMA needs the box to
                                                // start on a 16-byte boundary, so the left halo (2*C <= 16 B)
                                                // is fetched as a whole 16-byte granule
constexpr int BN_BOXW = 128;                    // HX + TW + 2*C rounded up to 16 B (C <= 4)
constexpr int BN_BOXH = BN_TH + 4;
constexpr int BN_SLOT = BN_BOXW * BN_BOXH;      // 2560 B per ring slot, 128-B aligned
#ifndef MULUT_BN_RING
#define MULUT_BN_RING 6
#endif
#ifndef MULUT_BN_AHEAD
#define MULUT_BN_AHEAD 2
#endif
#ifndef MULUT_BN_GROUPS
#define MULUT_BN_GROUPS 2
#endif
#ifndef MULUT_BN_GT
#define MULUT_BN_GT 384
#endif
#ifndef MULUT_BN_QCAP
#define MULUT_BN_QCAP 4096
#endif
constexpr int BN_RING = MULUT_BN_RING;
constexpr int BN_AHEAD = MULUT_BN_AHEAD;        // TMA prefetch distance in tiles
constexpr int BN_CARRY = BN_RING - BN_AHEAD - 1;   // tiles a queue entry may outlive its scan
constexpr int BN_GROUPS = MULUT_BN_GROUPS;      // independent tile pipelines per CTA (own ring, queue, named barrier): one
                                                // group's scan / barrier / TMA wait hides behind the other's rounds
constexpr int BN_GT = MULUT_BN_GT;              // threads per group
constexpr int BN_THREADS = BN_GROUPS * BN_GT;   // 24 warps (1024 threads fit at 62 registers but measured no faster)
constexpr int BN_SCAN_THREADS = BN_TW / 4 * BN_TH;   // TW/4 x TH words of a tile, one per scanning thread
constexpr int BN_QCAP = MULUT_BN_QCAP;          // queue capacity per group: power of two > (GT - 1 + TW*TH) + TW*TH
constexpr int BN_SLAB_WORDS = 4916;             // 17^3 = 4913 rows, padded so a slab is a 16-B multiple
constexpr int BN_SLAB_BYTES = BN_SLAB_WORDS * 4;
constexpr int BN_BIN_BYTES = 3 * BN_SLAB_BYTES; // slabs 2b, 2b+1, 2b+2 of one mode
constexpr int BN_MAX_MODES = 3;
constexpr size_t BN_SMEM = (size_t)BN_GROUPS * (BN_RING * BN_SLOT + BN_QCAP * 2) + (size_t)BN_MAX_MODES * BN_BIN_BYTES;

static_assert(BN_TW / 4 * BN_TH == BN_SCAN_THREADS && BN_SCAN_THREADS % 32 == 0 && BN_SCAN_THREADS <= BN_GT,
              "scan mapping: one 4-sample word per scanning thread, whole warps");
static_assert(BN_QCAP >= BN_GT + 2 * BN_TW * BN_TH, "queue too small");
static_assert(BN_HX + BN_TW + 2 * 4 <= BN_BOXW && 2 * 4 <= BN_HX, "box too narrow for C <= 4");

size_t slab_major_bytes() { return (size_t)17 * BN_SLAB_BYTES; }

// byte[(a*4916 + b*289 + c*17 + d)*4 + j] = LUT[a*17^3 + b*17^2 + c*17 + d][j] + 128
__global__ void build_slab_major2_kernel(const int8_t *__restrict__ lut, uint8_t *__restrict__ slabs)
{
    const int total = 17 * BN_SLAB_BYTES;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int w = i >> 2, j = i & 3;
        const int a = w / BN_SLAB_WORDS, rem = w - a * BN_SLAB_WORDS;
        slabs[i] = rem < 4913 ? (uint8_t)((int)lut[((size_t)a * 4913 + rem) * 4 + j] + 128) : (uint8_t)0;
    }
}

int build_slab_major(const int8_t *d_lut, uint8_t *d_slabs, cudaStream_t stream)
{
    build_slab_major2_kernel<<<256, 256, 0, stream>>>(d_lut, d_slabs);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

// ---------------------------------------------------------------------------
// launch preparation: stand-alone histogram + plan, orphan collection (the control block, the plan
// and the bin counter live in binned.cuh)
// ---------------------------------------------------------------------------
size_t binned_ctl_bytes() { return 256; }

// Stand-alone histogram + plan, for a stage whose input is not produced by K1b (single-stage models).
__global__ void __launch_bounds__(256)
bin_hist_kernel(const uint8_t *__restrict__ img, size_t total, BinPlanArgs pa)
{
    __shared__ uint32_t s_hist[BN_BINS];
    if (threadIdx.x < BN_BINS) s_hist[threadIdx.x] = 0;
    __syncthreads();
    BinCounter bc;
    const size_t n16 = total / 16;
    const uint4 *__restrict__ v = reinterpret_cast<const uint4 *>(img);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 q = __ldg(v + i);
        bc.add_word(q.x); bc.add_word(q.y); bc.add_word(q.z); bc.add_word(q.w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = n16 * 16; i < total; ++i) bc.add_byte(img[i]);
    bc.reduce_into(s_hist);
    __syncthreads();
    if (threadIdx.x < BN_BINS && s_hist[threadIdx.x]) atomicAdd(&pa.ctl->hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&pa.ctl->ticket, 1u) == gridDim.x - 1) {      // last block: plan
        __threadfence();
        bin_plan(pa.ctl, pa.n_tiles, pa.G, pa.list_cap, pa.allow_orphans);
    }
}

// Orphan list: linear indices of the samples whose bin has no resident CTAs.  Every CTA of K1f scans
// its slice of the stage input before it starts on its tiles (the scan overlaps the LUT bulk copies).
__device__ __forceinline__ void collect_orphans(const uint8_t *__restrict__ img, size_t total, uint32_t mask,
                                                uint32_t *__restrict__ count, uint32_t *__restrict__ list)
{
    const size_t n16 = total / 16;
    const uint4 *__restrict__ v = reinterpret_cast<const uint4 *>(img);
    const int lane = threadIdx.x & 31;
    constexpr int U = 4;                                     // independent 16-byte loads in flight per thread
    // byte b of {lut_lo, lut_hi} = 0xFF if bin b is an orphan bin: one PRMT looks up four samples at once
    uint32_t lut_lo = 0, lut_hi = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        if ((mask >> b) & 1u) lut_lo |= 0xFFu << (8 * b);
        if ((mask >> (b + 4)) & 1u) lut_hi |= 0xFFu << (8 * b);
    }
    const size_t stride = (size_t)gridDim.x * blockDim.x * U;
    for (size_t i0 = ((size_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) * U; i0 < n16; i0 += stride) {
        uint4 q[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t i = i0 + u * 32 + lane;
            q[u] = i < n16 ? __ldg(v + i) : make_uint4(0, 0, 0, 0);
        }
        uint32_t hits[U];
        uint32_t cnt = 0;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
            uint32_t h[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t x = (w[k] >> 5) & 0x07070707u;      // the four bins, one per byte
                x |= x >> 4;                                 // byte 0 = b0 | b1 << 4, byte 2 = b2 | b3 << 4
                h[k] = __byte_perm(lut_lo, lut_hi, __byte_perm(x, 0u, 0x4420));
            }
            hits[u] = 0;
            if ((h[0] | h[1] | h[2] | h[3]) && (i0 + u * 32 + lane) < n16) {
#pragma unroll
                for (int k = 0; k < 4; ++k) hits[u] |= (((h[k] & 0x08040201u) * 0x01010101u) >> 24) << (4 * k);
            }
            cnt += __popc(hits[u]);
        }
        if (!__any_sync(0xffffffffu, cnt != 0)) continue;
        uint32_t incl = cnt;                                 // one warp scan + one atomic for the U chunks
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t base = 0;
        if (lane == 31) base = atomicAdd(count, incl);
        base = __shfl_sync(0xffffffffu, base, 31);
        uint32_t pos = base + incl - cnt;
#pragma unroll
        for (int u = 0; u < U; ++u) {
            uint32_t hh = hits[u];
            const size_t i = i0 + u * 32 + lane;
            while (hh) {
                const int e = __ffs(hh) - 1;
                hh &= hh - 1;
                list[pos++] = (uint32_t)(i * 16 + e);
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = n16 * 16; i < total; ++i)
            if ((mask >> (img[i] >> 5)) & 1u) list[atomicAdd(count, 1u)] = (uint32_t)i;
}

// ---------------------------------------------------------------------------
// the interpolation body: one mode, four rotations, LUT rows from shared memory
// ---------------------------------------------------------------------------
__host__ __device__ constexpr int bn_tap_off(char mode, int r, int k, bool want_dy)
{
    int dy = mode == 's' ? (k >> 1) : mode == 'd' ? 2 * (k >> 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 1 : 2);
    int dx = mode == 's' ? (k & 1) : mode == 'd' ? 2 * (k & 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 2 : 1);
    for (int i = 0; i < r; ++i) { int t = dy; dy = dx; dx = -t; }
    return want_dy ? dy : dx;
}

template <char MODE, int CT>
__device__ __forceinline__ void binned_mode(const uint8_t *__restrict__ sp, uint32_t t0,
                                            const uint8_t *__restrict__ lut_a /* mode table + a_rel slab */,
                                            uint32_t (&AE)[4], unsigned long long (&AO)[4])
{
    constexpr int P = BN_BOXW;
    constexpr uint32_t SA = BN_SLAB_BYTES, SB = 289u * 4u, SC = 17u * 4u, SD = 4u;
    const uint32_t k0c = (t0 << 28) | SA;          // fraction in the top nibble, byte stride below
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t1 = sp[bn_tap_off(MODE, r, 1, true) * P + bn_tap_off(MODE, r, 1, false) * CT];
        const uint32_t t2 = sp[bn_tap_off(MODE, r, 2, true) * P + bn_tap_off(MODE, r, 2, false) * CT];
        const uint32_t t3 = sp[bn_tap_off(MODE, r, 3, true) * P + bn_tap_off(MODE, r, 3, false) * CT];
        // base vertex through three IMADs on the m = t >> 4 the taps share (left to itself the compiler packs
        // two of them for a dp2a with PRMTs and masks the third out of t >> 2: all on the ALU pipe; 3.57 -> 3.48 ms)
        uint32_t v0;
        {
            const uint32_t m1 = t1 >> 4, m2 = t2 >> 4, m3 = t3 >> 4;
            uint32_t x;
            asm("mad.lo.u32 %0, %1, %2, 0;" : "=r"(x) : "r"(m3), "r"(SD));
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(x) : "r"(m2), "r"(SC), "r"(x));
            asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(v0) : "r"(m1), "r"(SB), "r"(x));
        }
        // (forcing these through mad.lo to unload the ALU pipe measured 4 % slower: ptxas loses the shared shifts)
        // (the keys themselves through mad.lo: ptxas turns them into LEA on the ALU pipe again, 3.60 ms)
        uint32_t k0 = k0c, k1 = (t1 << 28) | SB, k2 = (t2 << 28) | SC, k3 = (t3 << 28) | SD;
        sort4_desc(k0, k1, k2, k3);
        const uint32_t f1 = k0 >> 28, f2 = k1 >> 28, f3 = k2 >> 28, f4 = k3 >> 28;
        // stride = key - (f << 28) as ONE IMAD (f * 0xF0000000 + key) instead of a LOP3 mask: the kernel is bound
        // by the ALU pipe (72 %), the FMA pipe has room - 3.66 -> 3.57 ms on cfg 2 although it costs three more adds
        uint32_t s1, s2, s4;
        asm("mad.lo.u32 %0, %1, 0xF0000000, %2;" : "=r"(s1) : "r"(f1), "r"(k0));
        asm("mad.lo.u32 %0, %1, 0xF0000000, %2;" : "=r"(s2) : "r"(f2), "r"(k1));
        asm("mad.lo.u32 %0, %1, 0xF0000000, %2;" : "=r"(s4) : "r"(f4), "r"(k3));
        const uint32_t v1 = v0 + s1, v2 = v1 + s2;
        const uint32_t v4 = v0 + (SA + SB + SC + SD), v3 = v4 - s4;
        const uint32_t r0 = *reinterpret_cast<const uint32_t *>(lut_a + v0);
        const uint32_t r1 = *reinterpret_cast<const uint32_t *>(lut_a + v1);
        const uint32_t r2 = *reinterpret_cast<const uint32_t *>(lut_a + v2);
        const uint32_t r3 = *reinterpret_cast<const uint32_t *>(lut_a + v3);
        const uint32_t r4 = *reinterpret_cast<const uint32_t *>(lut_a + v4);
        const uint32_t w0 = 16u - f1, w1 = f1 - f2, w2 = f2 - f3, w3 = f3 - f4, w4 = f4;
        constexpr uint32_t EM = 0x00FF00FFu;
        // (the even outputs through IDP.2A.LO/HI with the bare weight - no mask, two FMA-pipe ops - measured 3.70 ms
        // against 3.57; the weights through mad.lo 3.60)
        AE[r] += (r0 & EM) * w0 + (r1 & EM) * w1 + (r2 & EM) * w2 + (r3 & EM) * w3 + (r4 & EM) * w4;
        AO[r] += (unsigned long long)r0 * w0 + (unsigned long long)r1 * w1 + (unsigned long long)r2 * w2 +
                 (unsigned long long)r3 * w3 + (unsigned long long)r4 * w4;
    }
}

struct BinnedArgs {
    uint8_t *out;                            // (N, 2H, 2W, C)
    int N, H, W, C;
    int n_modes;
    char modes[BN_MAX_MODES + 1];
    const uint8_t *slabs[BN_MAX_MODES];      // slab-major biased tables
    BinCtl *ctl;                             // plan: CTAs per bin, orphan mask and counter
    const uint8_t *in;                       // the stage input (orphan scan) and its size in samples
    size_t total;
    uint32_t *list;                          // orphan list, or null
};

// Optional phase timing (build with -DMULUT_BN_TIMING): thread 0 of every CTA accumulates
// clock64() deltas per phase; read back with mulut_debug_bn_timing().
#ifdef MULUT_BN_TIMING
__device__ unsigned long long g_bn_timing[256 * 8];
#define BN_T0() unsigned long long t_prev = clock64(), t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define BN_T(k) do { if (threadIdx.x == 0) { const unsigned long long t_now = clock64(); t_acc[k] += t_now - t_prev; t_prev = t_now; } } while (0)
#define BN_TN(k, n) do { if (threadIdx.x == 0) t_acc[k] += (n); } while (0)
#define BN_TEND() do { if (threadIdx.x == 0) { t_acc[7] |= (unsigned long long)bin << 48; for (int k = 0; k < 8; ++k) g_bn_timing[(blockIdx.x & 255) * 8 + k] = t_acc[k]; } } while (0)
#else
#define BN_T0() do {} while (0)
#define BN_T(k) do {} while (0)
#define BN_TN(k, n) do {} while (0)
#define BN_TEND() do {} while (0)
#endif

// SDY: the model's modes are exactly "sdy" (the shipped configuration): the three mode bodies run back to back
// with no per-sample dispatch and share the tap loads they have in common
template <int CT, bool SDY>
__global__ void __launch_bounds__(BN_THREADS, 1)
stage_last2_binned_kernel(const __grid_constant__ BinnedArgs a, const __grid_constant__ CUtensorMap tmap)
{
    extern __shared__ __align__(128) uint8_t bn_smem[];
    const int grp = threadIdx.x / BN_GT, tid = threadIdx.x - grp * BN_GT, lane = tid & 31;   // tid: index within the group
    uint8_t *s_ring = bn_smem + grp * (BN_RING * BN_SLOT);
    uint16_t *s_queue = reinterpret_cast<uint16_t *>(bn_smem + BN_GROUPS * BN_RING * BN_SLOT + grp * (BN_QCAP * 2));
    uint8_t *s_lut = bn_smem + BN_GROUPS * (BN_RING * BN_SLOT + BN_QCAP * 2);
    __shared__ __align__(8) uint64_t s_full_all[BN_GROUPS][BN_RING];
    __shared__ __align__(8) uint64_t s_lutbar;
    __shared__ int4 s_info_all[BN_GROUPS][BN_RING];   // n, y0, X0, border
    __shared__ uint32_t s_cnt_all[BN_GROUPS][2];      // entries queued by even / odd tiles (monotonic)
    __shared__ int s_alloc[3];                   // bin, index within bin, CTAs of the bin
    uint64_t *s_full = s_full_all[grp];
    int4 *s_info = s_info_all[grp];
    uint32_t *s_cnt = s_cnt_all[grp];

    const int WC = a.W * CT;
    const int tiles_x = (WC + BN_TW - 1) / BN_TW;
    const int tiles_y = (a.H + BN_TH - 1) / BN_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;

    // ---- my bin: the plan kernel dealt g[b] CTAs to bin b (same answer in every CTA) ----
    if (threadIdx.x == 0) {
        int g[BN_BINS];
        for (int b = 0; b < BN_BINS; ++b) g[b] = a.ctl->g[b];
        int idx = blockIdx.x, b = 0;
        while (b < BN_BINS && idx >= g[b]) { idx -= g[b]; ++b; }
        s_alloc[0] = b; s_alloc[1] = idx; s_alloc[2] = b < BN_BINS ? g[b] : 1;
        for (int g2 = 0; g2 < BN_GROUPS; ++g2) {
            for (int i = 0; i < BN_RING; ++i) mbar_init(smem_u32(&s_full_all[g2][i]), 1);
            s_cnt_all[g2][0] = 0; s_cnt_all[g2][1] = 0;
        }
        mbar_init(smem_u32(&s_lutbar), 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int bin = s_alloc[0], me = s_alloc[1], gb = s_alloc[2];
    const bool idle = bin >= BN_BINS;
    if (idle) {                                            // no tiles for this CTA: it still scans its slice for orphans
        if (a.list && a.ctl->orphan_mask) collect_orphans(a.in, a.total, a.ctl->orphan_mask, &a.ctl->list_count, a.list);
        return;
    }
    (void)me; (void)gb;
    // Tiles are claimed dynamically: the CTAs of a bin draw tile indices from one atomic counter, so a CTA
    // that lands on tiles rich in its bin (natural frames cluster values in space) simply draws fewer of them.
    // Thread 0 keeps one claim in flight (`pending`) so the atomic's latency hides behind the rounds.
    unsigned pending = 0;                                   // thread 0: the tile index claimed for the next issue

    auto issue = [&](int i, unsigned tile) {               // thread 0 only: TMA load of my i-th tile, or the end marker
        const int slot = i % BN_RING;
        if (tile >= (unsigned)n_tiles) { s_info[slot] = make_int4(0, 0, 0, -1); return; }
        const unsigned tr = tile / (unsigned)tiles_x;
        const int tx = (int)(tile - tr * (unsigned)tiles_x);
        const int n = (int)(tr / (unsigned)tiles_y);
        const int ty = (int)(tr - (unsigned)n * (unsigned)tiles_y);
        const int y0 = ty * BN_TH, X0 = tx * BN_TW;
        const int border = (y0 < 2) || (y0 + BN_TH + 2 > a.H) || (X0 < 2 * CT) || (X0 + BN_TW + 2 * CT > WC);
        s_info[slot] = make_int4(n, y0, X0, border);
        const uint32_t bar = smem_u32(&s_full[slot]);
        mbar_expect_tx(bar, BN_BOXW * BN_BOXH);
        tma_load_3d(smem_u32(s_ring + slot * BN_SLOT), &tmap, X0 - BN_HX, y0 - 2, n, bar);
    };
    auto claim = [&]() -> unsigned { return atomicAdd(&a.ctl->next_tile[bin], 1u); };     // thread 0 only

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap);
        const uint32_t lb = smem_u32(&s_lutbar);
        mbar_expect_tx(lb, (uint32_t)a.n_modes * BN_BIN_BYTES);
        for (int m = 0; m < a.n_modes; ++m)
            bulk_g2s(smem_u32(s_lut + m * BN_BIN_BYTES), a.slabs[m] + (size_t)bin * 2 * BN_SLAB_BYTES, BN_BIN_BYTES, lb);
    }
    if (tid == 0) {                                         // each group's leader starts its own pipeline
        unsigned t = claim();
        for (int i = 0; i < BN_AHEAD; ++i) {
            issue(i, t);
            if (t < (unsigned)n_tiles) t = claim();        // a CTA draws past the end once, then stops
        }
        pending = t;
    }
    __syncthreads();                                        // s_info of the first tiles is read before any mbarrier wait
    // orphan scan of my slice of the input while the LUT slabs and the first tiles are in flight
    if (a.list && a.ctl->orphan_mask) collect_orphans(a.in, a.total, a.ctl->orphan_mask, &a.ctl->list_count, a.list);

    // scan mapping: the tile's TW x TH samples are TW/4 x TH = BN_THREADS words, one per thread
    const int srow = tid / (BN_TW / 4), swc = tid - srow * (BN_TW / 4);
    const int scan_off = (srow + 2) * BN_BOXW + BN_HX + swc * 4;
    const uint32_t bin_pat = (uint32_t)bin * 0x01010101u;
    const int oWC = 2 * WC;
    const uint32_t den = 16u * a.n_modes;
    const uint32_t magic = 0xFFFFFFFFu / den + 1u;                 // exact n / den for n < 2^16
    const int bias_total = a.n_modes * 4 * 2048;                   // 16 * 128 per interpolation
    uint32_t head = 0, tail = 0;
    uint32_t seen0 = 0, seen1 = 0;                                 // s_cnt[0/1] as of my last look
    uint32_t bound[BN_CARRY + 1];                                  // bound[k] = queue tail after the scan of tile i-k
#pragma unroll
    for (int k = 0; k <= BN_CARRY; ++k) bound[k] = 0;

    // One block barrier per tile: [wait TMA] [patch border] [scan -> queue] [barrier] [issue TMA] [rounds].
    // Threads drift by at most one scan: the queue holds a whole tile beyond the unconsumed entries,
    // the two entry counters alternate, and a ring slot is re-filled only after the barrier that
    // follows the rounds that drained it.
    BN_T0();
    for (int i = 0;; ++i) {
        const int slot = i % BN_RING;
        const int4 info = s_info[slot];                    // published by a block barrier at least one tile ago
        const bool end = info.w < 0;                       // the bin has no more tiles: drain the queue and leave
        if (!end) {
        mbar_wait(smem_u32(&s_full[slot]), (uint32_t)(i / BN_RING) & 1u);
        BN_T(0);
        uint8_t *tile = s_ring + slot * BN_SLOT;
        if (info.w) {
            // replicate padding: overwrite the zero-filled out-of-frame cells with the clamped
            // in-frame value (always inside this tile, never written by this pass, never read by the scan)
            for (int idx = tid; idx < BN_BOXW * BN_BOXH; idx += BN_GT) {
                const int r = idx / BN_BOXW, j = idx - r * BN_BOXW;
                const int gy = info.y - 2 + r, gx = info.z - BN_HX + j;
                const int cy = clampi(gy, 0, a.H - 1);
                int cx = gx;
                if (gx < 0) cx = (gx + BN_HX * CT) % CT;
                else if (gx >= WC) cx = WC - CT + (gx % CT);
                if (cy != gy || cx != gx) {
                    const int j2 = cx - info.z + BN_HX;
                    if (j2 >= 0 && j2 < BN_BOXW) tile[r * BN_BOXW + j] = tile[(cy - info.y + 2) * BN_BOXW + j2];
                }
            }
        }

        BN_T(1);
        // ---- scan: queue the samples of my bin (4 samples = one word per thread of the first 24 warps) ----
        if (tid < BN_SCAN_THREADS) {
            const uint32_t w = *reinterpret_cast<const uint32_t *>(tile + scan_off);
            const int nv = WC - (info.z + swc * 4);                                   // samples of this word inside the frame
            const bool valid = (info.y + srow < a.H) && nv > 0;
            const uint32_t x = ((w >> 5) & 0x07070707u) ^ bin_pat;                    // byte == 0  <=>  sample in my bin
            uint32_t m = valid ? (~(x + 0x7F7F7F7Fu) & 0x80808080u) : 0u;
            if (nv < 4) m &= 0x80808080u >> (8 * (4 - max(nv, 1)));                   // a row that does not end on a word
            const uint32_t cnt = __popc(m);
            uint32_t incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            uint32_t old = 0;
            if (lane == 31 && incl)        // raw atom: nvcc would wrap atomicAdd in its own warp aggregation
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(&s_cnt[i & 1])), "r"(incl) : "memory");
            old = __shfl_sync(0xffffffffu, old, 31);
            uint32_t pos = tail + (old - ((i & 1) ? seen1 : seen0)) + incl - cnt;
            const uint32_t ebase = (uint32_t)(slot << 12) | (uint32_t)(srow * BN_TW + swc * 4);
            while (m) {
                const int b = __ffs(m) - 1;                                            // bit 7, 15, 23 or 31
                m &= m - 1;
                s_queue[pos & (BN_QCAP - 1)] = (uint16_t)(ebase + (b >> 3));
                ++pos;
            }
        }
        BN_T(2);
        if (i == 0) mbar_wait(smem_u32(&s_lutbar), 0u);
        asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "r"(BN_GT) : "memory");     // the group's barrier
        BN_T(3);
        {
            const uint32_t c = *reinterpret_cast<volatile uint32_t *>(&s_cnt[i & 1]);
            if (i & 1) { tail += c - seen1; seen1 = c; } else { tail += c - seen0; seen0 = c; }
        }
        if (tid == 0) {                                    // the claim made one tile ago lands here; the next one flies during the rounds
            const unsigned t = pending;
            issue(i + BN_AHEAD, t);
            if (t < (unsigned)n_tiles) pending = claim();
        }
#pragma unroll
        for (int k = BN_CARRY; k > 0; --k) bound[k] = bound[k - 1];
        bound[0] = tail;
        } else if (i == 0) {
            mbar_wait(smem_u32(&s_lutbar), 0u);            // never leave with the slab copy in flight
        }
        const uint32_t must = end ? tail : bound[BN_CARRY];

        // ---- interpolate: full rounds; a partial round only to release an old ring slot ----
        for (;;) {
            const uint32_t avail = tail - head;
            uint32_t n;
            if (avail >= BN_GT) n = BN_GT;
            else if ((int)(must - head) > 0) n = avail;
            else break;
            if ((uint32_t)tid < n) {
                const uint32_t e = s_queue[(head + tid) & (BN_QCAP - 1)];
                const int eslot = e >> 12, s = e & 4095;
                const int row = s / BN_TW, col = s - row * BN_TW;
                const int4 ti = s_info[eslot];
                const uint8_t *sp = s_ring + eslot * BN_SLOT + (row + 2) * BN_BOXW + col + BN_HX;
                const uint32_t t0 = sp[0];
                const int a_rel = (int)(t0 >> 4) - 2 * bin;
                uint32_t AE[4] = {0u, 0u, 0u, 0u};
                unsigned long long AO[4] = {0ull, 0ull, 0ull, 0ull};
                if (SDY) {
                    const uint8_t *lut_a = s_lut + a_rel * BN_SLAB_BYTES;
                    binned_mode<'s', CT>(sp, t0, lut_a, AE, AO);
                    binned_mode<'d', CT>(sp, t0, lut_a + BN_BIN_BYTES, AE, AO);
                    binned_mode<'y', CT>(sp, t0, lut_a + 2 * BN_BIN_BYTES, AE, AO);
                } else {
                    for (int m = 0; m < a.n_modes; ++m) {
                        const uint8_t *lut_a = s_lut + m * BN_BIN_BYTES + a_rel * BN_SLAB_BYTES;
                        switch (a.modes[m]) {
                        case 's': binned_mode<'s', CT>(sp, t0, lut_a, AE, AO); break;
                        case 'd': binned_mode<'d', CT>(sp, t0, lut_a, AE, AO); break;
                        default: binned_mode<'y', CT>(sp, t0, lut_a, AE, AO); break;
                        }
                    }
                }
                // fields: AE = S0 | S2<<16;  AO - AE = S1<<8 | S3<<24 (S_j < 2^16: 12 interp x 16 x 255)
                int S[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const unsigned long long D = AO[r] - (unsigned long long)AE[r];
                    S[subpixel_perm<2>(r, 0)] += (int)(AE[r] & 0xFFFFu);
                    S[subpixel_perm<2>(r, 1)] += (int)((uint32_t)(D >> 8) & 0xFFFFu);
                    S[subpixel_perm<2>(r, 2)] += (int)(AE[r] >> 16);
                    S[subpixel_perm<2>(r, 3)] += (int)((uint32_t)(D >> 24) & 0xFFFFu);
                }
                const int y = ti.y + row, xb = ti.z + col;
                const int x = xb / CT, c = xb - x * CT;
                uint8_t *__restrict__ op = a.out + ((size_t)ti.x * (2 * a.H) + 2 * y) * (size_t)oWC + (size_t)(2 * x) * CT + c;
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int v = 0; v < 2; ++v) {
                        const int val = S[u * 2 + v] - bias_total;
                        const uint32_t nn = (uint32_t)max(val, 0);
                        uint32_t qd = __umulhi(nn, magic);
                        const uint32_t rm = nn - qd * den;
                        qd += (2u * rm > den) || ((2u * rm == den) && (qd & 1u));
                        op[(size_t)u * oWC + v * CT] = (uint8_t)min(qd, 255u);
                    }
            }
            head += n;
            BN_TN(5, 1);
            BN_TN(6, n);
        }
        BN_T(4);
        BN_TN(7, 1);
        if (end) break;
    }
    BN_TEND();
}

// ---------------------------------------------------------------------------
// launcher
// ---------------------------------------------------------------------------
bool binned_supported(const StageArgs &a, int up)
{
    return up == 2 && a.last && a.interval == 4 && a.n_modes >= 1 && a.n_modes <= BN_MAX_MODES &&
           a.C >= 1 && a.C <= 4 && a.lut_slab[0] != nullptr;
}

template <int CT, bool SDY>
static int launch_binned_ts(const BinnedArgs &b, const CUtensorMap &tmap, int num_sms, cudaStream_t stream)
{
    // per device and cheap: set on every launch (one process may own several devices)
    MULUT_CUDA(cudaFuncSetAttribute(stage_last2_binned_kernel<CT, SDY>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)BN_SMEM));
    stage_last2_binned_kernel<CT, SDY><<<num_sms, BN_THREADS, BN_SMEM, stream>>>(b, tmap);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

template <int CT>
static int launch_binned_t(const BinnedArgs &b, const CUtensorMap &tmap, int num_sms, cudaStream_t stream)
{
    const bool sdy = b.n_modes == 3 && b.modes[0] == 's' && b.modes[1] == 'd' && b.modes[2] == 'y';
    return sdy ? launch_binned_ts<CT, true>(b, tmap, num_sms, stream) : launch_binned_ts<CT, false>(b, tmap, num_sms, stream);
}

int launch_stage_generic_list2(const StageArgs &a, const uint32_t *list, const uint32_t *count, cudaStream_t stream);

static int binned_allow_orphans(const uint32_t *list, size_t total)
{
    const char *env = getenv("MULUT_BN_ORPHANS");
    return (!env || env[0] != '0') && list && total < 0xFFFFFFFFull;
}

// What K1b needs to produce the histogram + plan of the NEXT stage's input (N,H,W,C as this stage's).
BinPlanArgs binned_plan_args(const StageArgs &a, void *ctl_mem, const uint32_t *list, size_t list_cap)
{
    BinPlanArgs pa;
    const int WC = a.W * a.C;
    pa.ctl = static_cast<BinCtl *>(ctl_mem);
    pa.n_tiles = (long long)a.N * ((a.H + BN_TH - 1) / BN_TH) * ((WC + BN_TW - 1) / BN_TW);
    pa.G = a.num_sms;
    pa.list_cap = (unsigned long long)list_cap;
    pa.allow_orphans = binned_allow_orphans(list, (size_t)a.N * a.H * WC);
    return pa;
}

// ctl: binned_ctl_bytes() of device workspace; list: list_cap uint32 entries of device workspace.
// planned: the control block already holds the histogram + plan of a.in (K1b produced them).
int launch_stage_binned(const StageArgs &a, void *ctl_mem, uint32_t *list, size_t list_cap, bool planned,
                        cudaStream_t stream, int *launches, Prof *prof)
{
    const size_t total = (size_t)a.N * a.H * a.W * a.C;
    if (total == 0) return MULUT_OK;
    CUtensorMap tmap;
    if (!a.in_tma || tma_encode_frames(&tmap, a.in_tma, a.N, a.H, a.W * a.C, a.in_pitch, BN_BOXW, BN_BOXH) != 0)
        return 1;                                          // caller falls back
    BinCtl *ctl = static_cast<BinCtl *>(ctl_mem);
    const BinPlanArgs pa = binned_plan_args(a, ctl_mem, list, list_cap);
    if (pa.n_tiles >= 0x7fffffffLL) return 1;              // 32-bit tile arithmetic and tile counters in the kernel
    // the histogram and the orphan scan read the DENSE stage input with 16-byte loads (the orphan list holds
    // dense indices, so the pitched copy cannot stand in): a misaligned caller pointer - only possible for a
    // single-stage model, every later stage reads the workspace - goes to K1c
    if (reinterpret_cast<uintptr_t>(a.in) & 15u) return 1;
    BinnedArgs b;
    memset(&b, 0, sizeof b);
    b.out = a.out; b.N = a.N; b.H = a.H; b.W = a.W; b.C = a.C; b.n_modes = a.n_modes; b.ctl = ctl;
    b.in = a.in; b.total = total; b.list = pa.allow_orphans ? list : nullptr;
    for (int m = 0; m < a.n_modes; ++m) { b.modes[m] = a.modes[m]; b.slabs[m] = a.lut_slab[m]; }

    if (!planned) {
        prof->begin(MULUT_PROF_BIN_HIST, stream);
        MULUT_CUDA(cudaMemsetAsync(ctl, 0, binned_ctl_bytes(), stream));
        size_t blocks = (total / 16 + 255) / 256 + 1;
        if (blocks > (size_t)a.num_sms * 8) blocks = (size_t)a.num_sms * 8;
        bin_hist_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.in, total, pa);
        prof->end(stream);
        MULUT_CUDA(cudaGetLastError());
        *launches += 1;
    }
    prof->begin(MULUT_PROF_LAST_BINNED, stream);
    int rc = a.C == 3 ? launch_binned_t<3>(b, tmap, a.num_sms, stream)
           : a.C == 1 ? launch_binned_t<1>(b, tmap, a.num_sms, stream)
           : a.C == 4 ? launch_binned_t<4>(b, tmap, a.num_sms, stream)
                      : launch_binned_t<2>(b, tmap, a.num_sms, stream);
    prof->end(stream);
    if (rc) return rc;
    *launches += 1;
    if (pa.allow_orphans) {
        prof->begin(MULUT_PROF_BIN_ORPHANS, stream);
        rc = launch_stage_generic_list2(a, list, &ctl->list_count, stream);
        prof->end(stream);
        if (rc) return rc;
        *launches += 1;
    }
    return MULUT_OK;
}

}  // namespace mulut

// Host mirror of the plan the device computes (same code: bin_plan is __host__ __device__) - lets the
// CPU test-suite check the allocation logic without a GPU.
extern "C" int mulut_plan_bins(const unsigned long long *hist8, long long n_tiles, int n_ctas, unsigned long long list_cap,
                               int allow_orphans, int *ctas_per_bin8, unsigned *orphan_mask)
{
    if (!hist8 || !ctas_per_bin8 || !orphan_mask || n_tiles < 0 || n_ctas < 1) {
        mulut::set_error("mulut_plan_bins: bad argument");
        return MULUT_E_BAD_ARG;
    }
    mulut::BinCtl ctl;
    memset(&ctl, 0, sizeof ctl);
    for (int b = 0; b < mulut::BN_BINS; ++b) ctl.hist[b] = hist8[b];
    mulut::bin_plan(&ctl, n_tiles, n_ctas, list_cap, allow_orphans);
    for (int b = 0; b < mulut::BN_BINS; ++b) ctas_per_bin8[b] = ctl.g[b];
    *orphan_mask = ctl.orphan_mask;
    return MULUT_OK;
}

#ifdef MULUT_BN_TIMING
// out[8]: mean over CTAs of {wait, fixup, scan, barrier, rounds} cycles, rounds, entries, visits;
// out[8] = max over CTAs of the five phases' sum, out[9] = mean of it (load balance = out[9] / out[8])
extern "C" int mulut_debug_bn_timing(double *out, int n_ctas)
{
    static unsigned long long h[256 * 8];
    if (cudaMemcpyFromSymbol(h, mulut::g_bn_timing, sizeof h) != cudaSuccess) return -1;
    for (int k = 0; k < 8; ++k) {
        double acc = 0;
        for (int b = 0; b < n_ctas && b < 256; ++b) acc += (double)(k == 7 ? h[b * 8 + k] & 0xFFFFFFFFFFFFull : h[b * 8 + k]);
        out[k] = acc / n_ctas;
    }
    double mx = 0, mean = 0;
    for (int b = 0; b < n_ctas && b < 256; ++b) {
        double t = 0;
        for (int k = 0; k < 5; ++k) t += (double)h[b * 8 + k];
        mean += t / n_ctas;
        if (t > mx) mx = t;
    }
    out[8] = mx;
    out[9] = mean;
    return 0;
}
// the raw per-CTA counters (256 x 8; the CTA's bin sits in bits 48.. of counter 7)
extern "C" int mulut_debug_bn_timing_raw(unsigned long long *out)
{
    return cudaMemcpyFromSymbol(out, mulut::g_bn_timing, sizeof(unsigned long long) * 256 * 8) == cudaSuccess ? 0 : -1;
}
#endif
