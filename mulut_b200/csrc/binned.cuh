// Shared between the binned last-stage kernel K1f (infer_binned.cu) and the kernels that prepare its
// launch: the control block, the plan (which value bins get shared-memory CTAs, how many) and the
// orphan collection.  K1b (combine_kernel, infer_tiled.cu) counts the histogram of the bytes it
// writes and its last block runs the plan, so a two-stage pipeline needs no extra launch for it.
#pragma once
#include "common.cuh"

namespace mulut {

constexpr int BN_TW = 96;                       // tile width, byte columns (multiple of C for C <= 4)
#ifndef MULUT_BN_TH
#define MULUT_BN_TH 16
#endif
constexpr int BN_TH = MULUT_BN_TH;              // tile rows
constexpr int BN_BINS = 8;

struct BinCtl {                         // device, 256 B, zeroed before every launch
    unsigned long long hist[BN_BINS];   // samples per bin (sample >> 5)
    int g[BN_BINS];                     // CTAs dealt to each bin; 0 = bin not resident
    uint32_t orphan_mask;               // bins left to the L2-gather list kernel
    uint32_t list_count;                // orphan samples collected
    uint32_t ticket;                    // blocks of the histogram producer that have finished
    uint32_t next_tile[BN_BINS];        // K1f: the next unclaimed tile of each bin (CTAs of a bin claim tiles dynamically)
};
static_assert(sizeof(BinCtl) <= 256, "control block");

struct BinPlanArgs {                    // what the plan needs besides the histogram
    BinCtl *ctl;                        // null: no histogram / plan wanted
    long long n_tiles;
    int G;                              // CTAs of the binned kernel (one per SM)
    unsigned long long list_cap;        // capacity of the orphan list
    int allow_orphans;
};

// Plan (one thread): which bins get shared-memory CTAs, how many, and which are "orphans".
// A bin is worth a resident CTA group only if its samples outweigh the fixed cost of
// walking every tile once more (scan + barriers); sparse bins go to a list that the
// generic L2-gather kernel finishes (stage_generic_list_kernel).  Costs in SM cycles.
#ifndef MULUT_BN_CV
#define MULUT_BN_CV 1050
#endif
constexpr unsigned long long BN_CV = MULUT_BN_CV;   // per visit of a 96 x 16 tile (fit of tools/bn_timing.py's per-bin totals: 975-1085)
constexpr unsigned long long BN_CS = 10;        // per sample interpolated from shared memory (same fit: 10.1-10.2)
constexpr unsigned long long BN_CG = 26;        // per sample interpolated by the list kernel

__host__ __device__ inline void bin_plan(BinCtl *__restrict__ ctl, long long n_tiles, int G, unsigned long long list_cap,
                                int allow_orphans)
{
    // (this runs in ONE thread at the tail of the kernel that produced the histogram: keep it short -
    // independent loads first, float instead of emulated 64-bit division)
    unsigned long long n[BN_BINS], cost[BN_BINS], total = 0, orph = 0;
    bool res[BN_BINS];
#pragma unroll
    for (int b = 0; b < BN_BINS; ++b) n[b] = *reinterpret_cast<volatile unsigned long long *>(&ctl->hist[b]);
#pragma unroll
    for (int b = 0; b < BN_BINS; ++b) {
        res[b] = n[b] > 0 && (!allow_orphans || n[b] * (BN_CG - BN_CS) > (unsigned long long)n_tiles * BN_CV);
        if (n[b] && !res[b]) orph += n[b];
    }
    while (orph > list_cap) {                               // the list is bounded: promote the largest orphan bin
        int best = -1;
        for (int b = 0; b < BN_BINS; ++b)
            if (n[b] && !res[b] && (best < 0 || n[b] > n[best])) best = b;
        res[best] = true;
        orph -= n[best];
    }
    uint32_t mask = 0;
    for (int b = 0; b < BN_BINS; ++b) {
        cost[b] = res[b] ? (unsigned long long)n_tiles * BN_CV + n[b] * BN_CS : 0ull;
        total += cost[b];
        if (n[b] && !res[b]) mask |= 1u << b;
    }
    int g[BN_BINS], used = 0;
    for (int b = 0; b < BN_BINS; ++b) {
        // floor share in float (24-bit mantissa, G <= a few hundred: at most one CTA off, and the loops
        // below settle the sum at exactly G either way)
        const int share = (int)((float)cost[b] * ((float)G / (float)(total ? total : 1)));
        g[b] = cost[b] ? (share > 1 ? share : 1) : 0;
        used += g[b];
    }
    while (total && used < G) {                             // hand the rest to the most loaded groups
        int best = -1;
        for (int b = 0; b < BN_BINS; ++b)
            if (g[b] && (best < 0 || cost[b] * g[best] > cost[best] * g[b])) best = b;
        ++g[best]; ++used;
    }
    while (used > G) {
        int best = -1;
        for (int b = 0; b < BN_BINS; ++b)
            if (g[b] > 1 && (best < 0 || cost[b] * g[best] < cost[best] * g[b])) best = b;
        if (best < 0) break;
        --g[best]; --used;
    }
    for (int b = 0; b < BN_BINS; ++b) ctl->g[b] = g[b];
    ctl->orphan_mask = mask;
}

// Per-thread bin counting without dynamic register indexing: eight 8-bit counters in one 64-bit word.
struct BinCounter {
    unsigned long long packed = 0;
    uint32_t cnt[BN_BINS] = {0, 0, 0, 0, 0, 0, 0, 0};
    int pending = 0;
    __device__ __forceinline__ void add_word(uint32_t w)    // four samples
    {
#pragma unroll
        for (int e = 0; e < 4; ++e) packed += 1ull << (((w >> (8 * e + 5)) & 7u) * 8);
        if (++pending == 60) flush();                       // 240 samples: no 8-bit field can overflow
    }
    __device__ __forceinline__ void add_byte(uint32_t b) { packed += 1ull << ((b >> 5) * 8); if (++pending == 240) flush(); }
    __device__ __forceinline__ void flush()
    {
#pragma unroll
        for (int b = 0; b < BN_BINS; ++b) cnt[b] += (uint32_t)(packed >> (8 * b)) & 0xFFu;
        packed = 0;
        pending = 0;
    }
    // warp shuffle + one shared-memory atomic per warp and bin; s_hist: BN_BINS zeroed uint32
    __device__ __forceinline__ void reduce_into(uint32_t *s_hist)
    {
        flush();
#pragma unroll
        for (int b = 0; b < BN_BINS; ++b) {
            uint32_t c = cnt[b];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
            if ((threadIdx.x & 31) == 0 && c) atomicAdd(s_hist + b, c);
        }
    }
};

}  // namespace mulut
