// Tiled sm_100a inference kernels (interval = 4: q = 16, L = 17).
//
//  K1a  stage_smem_kernel      any stage with up = 1.  "Mode-split": every CTA
//       owns ONE sampling mode for the whole launch and keeps that mode's
//       83 521-byte int8 LUT in shared memory (staged once by a TMA bulk copy,
//       cp.async.bulk + mbarrier -> UBLKCP).  It walks halo'd tiles and emits the
//       int16 partial sum of its 4 rotations per sample.  All five vertex
//       gathers of an interpolation are LDS.S8 (bank-conflict bound, ~20x the
//       rate of divergent global gathers).
//  K1b  combine_kernel         sums the per-mode int16 planes and applies the
//       stage epilogue (integer round-half-even), writing the uint8 image.
//  K1c  stage_last2_quad_kernel  last stage, up = 2.  LUTs are re-laid out
//       cell-major (one aligned 64-byte block per 16^4 cell = all 16 corners x 4
//       outputs, biased to uint8).  Four adjacent lanes fetch one cell with one
//       LDG.128 each, so a warp-level gather touches 8 cache lines instead of
//       32.  The sample's own lane sorts the four fractions once (branch-free
//       min/max network), packs the five simplex weights into bytes with one
//       subtraction and names the order by an 8-bit code; each quad lane turns
//       (code, lane) into a PRMT selector from a 2 KB shared table, so ONE PRMT
//       yields the weights of its 4 corners, folded with dp4a; partial sums meet
//       through warp shuffles.  The 2x2
//       pixel-shuffle is fused into a shared-memory output tile that is stored
//       with coalesced 16-byte writes.
//
// Arithmetic follows SURVEY.md 8-SPEC (sr/4_test_lut.py:14-237, :279-306).
#include <cuda/barrier>

#include "binned.cuh"
#include "common.cuh"
#include "infer.cuh"

namespace mulut {

constexpr int TWB = 96;                 // tile width in byte columns (multiple of C for C in {1,2,3,4,6})
constexpr int LUT1_ROWS = 83521;        // 17^4
constexpr int LUT1_SMEM = 83584;        // padded to a multiple of 128 B
constexpr uint32_t FULL = 0xffffffffu;

// compile-time tap offsets per (mode, rotation, tap); see common.cuh::build_tap_table
__host__ __device__ constexpr int tap_off(char mode, int r, int k, bool want_dy)
{
    int dy = mode == 's' ? (k >> 1) : mode == 'd' ? 2 * (k >> 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 1 : 2);
    int dx = mode == 's' ? (k & 1) : mode == 'd' ? 2 * (k & 1) : (k == 0 ? 0 : k == 1 ? 1 : k == 2 ? 2 : 1);
    for (int i = 0; i < r; ++i) { int t = dy; dy = dx; dx = -t; }
    return want_dy ? dy : dx;
}

template <int CT> __host__ __device__ constexpr int tile_pitch() { return CT > 0 ? ((TWB + 4 * CT + 3) / 4) * 4 : TWB + 16; }

// smem(r, j) = img[clamp(y0-2+r)][clamp_pixel(X0 - 2C + j)], channel preserved.
__device__ __forceinline__ void fill_tile(uint8_t *s, int pitch, int rows, int cols,
                                          const uint8_t *__restrict__ img, int H, int C, int WC,
                                          int y0, int X0)
{
    for (int idx = threadIdx.x; idx < rows * cols; idx += blockDim.x) {
        const int r = idx / cols, j = idx - r * cols;
        const int yy = clampi(y0 - 2 + r, 0, H - 1);
        int xb = X0 - 2 * C + j;
        if (xb < 0) xb = (xb + 2 * C) % C;
        else if (xb >= WC) xb = WC - C + (xb % C);
        s[r * pitch + j] = __ldg(img + (size_t)yy * WC + xb);
    }
}

// ---------------------------------------------------------------------------
// K1a: shared-memory LUT, one mode per CTA, up = 1
// ---------------------------------------------------------------------------
template <char MODE, int CT>
__device__ __forceinline__ int smem_lut_sample(const uint8_t *__restrict__ sp, int C,
                                               const int8_t *__restrict__ slut)
{
    constexpr int P = tile_pitch<CT>();
    constexpr uint32_t SB = 20, SM = (1u << SB) - 1u;
    const int Cc = CT > 0 ? CT : C;
    const uint32_t t0 = sp[0];
    int acc = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t1 = sp[tap_off(MODE, r, 1, true) * P + tap_off(MODE, r, 1, false) * Cc];
        const uint32_t t2 = sp[tap_off(MODE, r, 2, true) * P + tap_off(MODE, r, 2, false) * Cc];
        const uint32_t t3 = sp[tap_off(MODE, r, 3, true) * P + tap_off(MODE, r, 3, false) * Cc];
        const uint32_t v0 = (((t0 >> 4) * 17u + (t1 >> 4)) * 17u + (t2 >> 4)) * 17u + (t3 >> 4);
        uint32_t k0 = ((t0 & 15u) << SB) | 4913u;
        uint32_t k1 = ((t1 & 15u) << SB) | 289u;
        uint32_t k2 = ((t2 & 15u) << SB) | 17u;
        uint32_t k3 = ((t3 & 15u) << SB) | 1u;
        sort4_desc(k0, k1, k2, k3);
        const int f1 = k0 >> SB, f2 = k1 >> SB, f3 = k2 >> SB, f4 = k3 >> SB;
        // vertex chain: the fraction bits pile up above bit 20 and are masked off
        const uint32_t v1 = v0 + k0, v2 = v1 + k1, v3 = v2 + k2, v4 = v3 + k3;
        acc += (16 - f1) * (int)slut[v0];
        acc += (f1 - f2) * (int)slut[v1 & SM];
        acc += (f2 - f3) * (int)slut[v2 & SM];
        acc += (f3 - f4) * (int)slut[v3 & SM];
        acc += f4 * (int)slut[v4 & SM];
    }
    return acc;
}

constexpr int S1_TH = 32;               // tile rows of K1a
constexpr int S1_RW = 4;                // row phases (warps = 3 * S1_RW)
constexpr int S1_THREADS = 96 * S1_RW;

template <int CT>
__global__ void __launch_bounds__(S1_THREADS, 2)
stage_smem_kernel(const __grid_constant__ StageArgs a, int16_t *__restrict__ partial, int ctas_per_mode)
{
    extern __shared__ __align__(128) uint8_t smem_raw[];
    int8_t *slut = reinterpret_cast<int8_t *>(smem_raw);
    uint8_t *s_in = smem_raw + LUT1_SMEM;
    __shared__ __align__(8) uint64_t lut_bar;

    constexpr int P = tile_pitch<CT>();
    const int C = CT > 0 ? CT : a.C;
    const int WC = a.W * C;
    const int m = blockIdx.x % a.n_modes;
    const int slot = blockIdx.x / a.n_modes;
    const char mode = a.modes[m];

    // --- stage this CTA's LUT into shared memory with one TMA bulk copy ---
    const uint32_t bar_addr = (uint32_t)__cvta_generic_to_shared(&lut_bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slut);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_addr), "r"((uint32_t)LUT1_SMEM)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
            "l"(a.lut_alt[m]), "r"((uint32_t)LUT1_SMEM), "r"(bar_addr)
            : "memory");
    }

    const int tiles_x = (WC + TWB - 1) / TWB;
    const int tiles_y = (a.H + S1_TH - 1) / S1_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp % 3) * 32 + lane;
    const int rp = warp / 3;
    const int cols = TWB + 4 * C;
    int16_t *__restrict__ plane = partial + (size_t)m * a.N * a.H * WC;
    bool lut_ready = false;

    for (long long tile = slot; tile < n_tiles; tile += ctas_per_mode) {
        const int tx = (int)(tile % tiles_x);
        const long long tr = tile / tiles_x;
        const int ty = (int)(tr % tiles_y);
        const int n = (int)(tr / tiles_y);
        const int y0 = ty * S1_TH, X0 = tx * TWB;
        const uint8_t *__restrict__ img = a.in + (size_t)n * a.H * WC;

        __syncthreads();                       // previous tile fully consumed
        fill_tile(s_in, P, S1_TH + 4, cols, img, a.H, C, WC, y0, X0);
        __syncthreads();
        if (!lut_ready) {                      // wait for the bulk copy (phase 0)
            uint32_t done = 0;
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\t"
                    "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                    "selp.b32 %0, 1, 0, p;\n\t}"
                    : "=r"(done)
                    : "r"(bar_addr), "r"(0u)
                    : "memory");
            }
            lut_ready = true;
        }
        const int xb = X0 + lx;
#pragma unroll 1
        for (int ly = rp; ly < S1_TH; ly += S1_RW) {
            const int y = y0 + ly;
            const uint8_t *sp = s_in + (ly + 2) * P + lx + 2 * C;
            int acc;
            switch (mode) {
            case 's': acc = smem_lut_sample<'s', CT>(sp, C, slut); break;
            case 'd': acc = smem_lut_sample<'d', CT>(sp, C, slut); break;
            default: acc = smem_lut_sample<'y', CT>(sp, C, slut); break;
            }
            if (y < a.H && xb < WC) plane[((size_t)n * a.H + y) * WC + xb] = (int16_t)acc;
        }
    }
    if (!lut_ready) {                          // never leave with a bulk copy in flight
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.b32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(bar_addr), "r"(0u)
                : "memory");
        }
    }
}

// ---------------------------------------------------------------------------
// K1b: sum the per-mode planes + stage epilogue (sr/4_test_lut.py:281-286,300-306)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
combine_kernel(const int16_t *__restrict__ partial, uint8_t *__restrict__ out, size_t total,
               int n_modes, int last, BinPlanArgs pa)
{
    const uint32_t den = last ? 16u * n_modes : 64u * n_modes;
    const uint32_t magic = rhe_magic(den);
    const int bias = last ? 0 : 127 * (int)den;
    // When the next stage is the binned kernel K1f, this kernel also counts the 8-bin histogram of
    // the bytes it writes (pa.ctl != null) and its last block turns it into K1f's plan.
    BinCounter bc;
    // 8 samples per thread when the planes allow 16-byte loads
    const bool vec = (total % 8 == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
    const size_t total8 = vec ? total / 8 : 0;
    // (the kernel runs at 75 % ALU pipe with HBM at 60 %: the mode sum is done on packed int16 pairs - |sum| <= 127 * 64 * M
    // fits 16 bits for M <= 4 - and the non-last epilogue skips both clamps)
    const bool packed_ok = n_modes <= 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total8;
         i += (size_t)gridDim.x * blockDim.x) {
        int s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (packed_ok) {
            int4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (j < n_modes) ? __ldg(reinterpret_cast<const int4 *>(partial + (size_t)j * total) + i)
                                     : make_int4(0, 0, 0, 0);
            uint32_t p[4] = {(uint32_t)v[0].x, (uint32_t)v[0].y, (uint32_t)v[0].z, (uint32_t)v[0].w};
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                p[0] = __vadd2(p[0], (uint32_t)v[j].x); p[1] = __vadd2(p[1], (uint32_t)v[j].y);
                p[2] = __vadd2(p[2], (uint32_t)v[j].z); p[3] = __vadd2(p[3], (uint32_t)v[j].w);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                s[2 * k] = (int)(int16_t)(p[k] & 0xffffu);
                s[2 * k + 1] = (int)p[k] >> 16;
            }
        } else {
        // the planes' loads are issued together (groups of four modes): bytes in flight, not instructions, bound this kernel
        for (int m0 = 0; m0 < n_modes; m0 += 4) {
            int4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                v[j] = (m0 + j < n_modes) ? __ldg(reinterpret_cast<const int4 *>(partial + (size_t)(m0 + j) * total) + i)
                                          : make_int4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int w[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    s[2 * k] += (int)(int16_t)(w[k] & 0xffff);
                    s[2 * k + 1] += w[k] >> 16;
                }
            }
        }
        }
        uint32_t lo = 0, hi = 0;
        if (!last) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                lo |= rhe_div_nonneg_magic((uint32_t)(s[k] + bias), den, magic) << (8 * k);
                hi |= rhe_div_nonneg_magic((uint32_t)(s[4 + k] + bias), den, magic) << (8 * k);
            }
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                lo |= rhe_div_clamp_u8_magic(s[k] + bias, den, magic) << (8 * k);
                hi |= rhe_div_clamp_u8_magic(s[4 + k] + bias, den, magic) << (8 * k);
            }
        }
        reinterpret_cast<uint2 *>(out)[i] = make_uint2(lo, hi);
        if (pa.ctl) { bc.add_word(lo); bc.add_word(hi); }
    }
    if (total8 == 0) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
             i += (size_t)gridDim.x * blockDim.x) {
            int s = 0;
            for (int m = 0; m < n_modes; ++m) s += partial[(size_t)m * total + i];
            const uint32_t o = rhe_div_clamp_u8_magic(s + bias, den, magic);
            out[i] = (uint8_t)o;
            if (pa.ctl) bc.add_byte(o);
        }
    }
    if (pa.ctl) {                                          // uniform: a kernel argument
        __shared__ uint32_t s_hist[BN_BINS];
        if (threadIdx.x < BN_BINS) s_hist[threadIdx.x] = 0;
        __syncthreads();
        bc.reduce_into(s_hist);
        __syncthreads();
        if (threadIdx.x < BN_BINS && s_hist[threadIdx.x])
            atomicAdd(&pa.ctl->hist[threadIdx.x], (unsigned long long)s_hist[threadIdx.x]);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0 && atomicAdd(&pa.ctl->ticket, 1u) == gridDim.x - 1) {   // last block: plan
            __threadfence();
            bin_plan(pa.ctl, pa.n_tiles, pa.G, pa.list_cap, pa.allow_orphans);
        }
    }
}

// ---------------------------------------------------------------------------
// K1c: last stage, up = 2, cell-major LUT, quad-cooperative gathers
// ---------------------------------------------------------------------------
constexpr int Q2_TH = 16;
constexpr int Q2_RW = 4;
constexpr int Q2_THREADS = 96 * Q2_RW;
constexpr int Q2_OP = 2 * TWB;          // output tile pitch in bytes

// coalesced copy-out of the (UP*TH) x (UP*TWB) byte output tile (pixel-shuffle already applied)
template <int UP>
__device__ __forceinline__ void store_out_tile(const uint8_t *s_out, uint8_t *__restrict__ out, int n, int H, int oWC,
                                               int y0, int X0)
{
    constexpr int OP = UP * TWB;
    uint8_t *__restrict__ outn = out + (size_t)n * (UP * H) * oWC;
    const int oy0 = UP * y0, oX0 = UP * X0;
    const int rows_valid = min(UP * Q2_TH, UP * H - oy0);
    const int cols_valid = min(OP, oWC - oX0);
    const bool vec_ok = ((oWC & 15) == 0) && ((reinterpret_cast<uintptr_t>(outn) & 15) == 0) && (cols_valid == OP);
    if (vec_ok) {
        constexpr int VPR = OP / 16;
        for (int idx = threadIdx.x; idx < rows_valid * VPR; idx += blockDim.x) {
            const int r = idx / VPR, c16 = idx - r * VPR;
            const uint4 v = *reinterpret_cast<const uint4 *>(s_out + r * OP + c16 * 16);
            *reinterpret_cast<uint4 *>(outn + (size_t)(oy0 + r) * oWC + oX0 + c16 * 16) = v;
        }
    } else {
        for (int idx = threadIdx.x; idx < rows_valid * OP; idx += blockDim.x) {
            const int r = idx / OP, c = idx - r * OP;
            if (c < cols_valid) outn[(size_t)(oy0 + r) * oWC + oX0 + c] = s_out[r * OP + c];
        }
    }
}

// Selector table: for every descending order of the four taps (code = t0 | t1<<2 |
// t2<<4 | t3<<6, t_k = tap with the k-th largest fraction) and every quad lane l
// (la = l>>1, lb = l&1), a PRMT selector that picks, for the lane's four corners
// (c,d) in {00,01,10,11}, the weight of the path vertex sitting on that corner:
// nibble k (0..4) = "this corner is the k-th vertex of the sorted-simplex walk"
// (bytes 0..3 of the packed weights w0..w3, byte 4 = w4), nibble 0xC = "not on the
// path" (sign-replicate of a byte whose top bit is clear -> 0x00).
__device__ __forceinline__ void build_selector_table(uint16_t *s_sel)
{
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) {
        const int code = i >> 2, l = i & 3;
        uint32_t sel = 0;
#pragma unroll
        for (int cd = 0; cd < 4; ++cd) {
            // bit t of S = tap t belongs to the corner (taps a,b,c,d = 0,1,2,3)
            const int S = ((l >> 1) & 1) | ((l & 1) << 1) | (((cd >> 1) & 1) << 2) | ((cd & 1) << 3);
            const int k = __popc(S);
            int prefix = 0;
            for (int j = 0; j < k; ++j) prefix |= 1 << ((code >> (2 * j)) & 3);
            sel |= (uint32_t)(prefix == S ? k : 0xC) << (4 * cd);
        }
        s_sel[i] = (uint16_t)sel;
    }
}

template <char MODE, int CT>
__device__ __forceinline__ void quad_mode(const uint8_t *__restrict__ sp, int C,
                                          const uint8_t *__restrict__ cells, int ql,
                                          const uint16_t *__restrict__ s_sel_lane, uint32_t (&acc)[4][4])
{
    constexpr int P = tile_pitch<CT>();
    const int Cc = CT > 0 ? CT : C;
    const uint8_t *__restrict__ cells_lane = cells + ql * 16;     // my 16 B of every cell
    const uint32_t t0 = sp[0];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t1 = sp[tap_off(MODE, r, 1, true) * P + tap_off(MODE, r, 1, false) * Cc];
        const uint32_t t2 = sp[tap_off(MODE, r, 2, true) * P + tap_off(MODE, r, 2, false) * Cc];
        const uint32_t t3 = sp[tap_off(MODE, r, 3, true) * P + tap_off(MODE, r, 3, false) * Cc];
        // ---- owner side: sort the fractions once per interpolation ----
        uint32_t k0 = ((t0 & 15u) << 2) | 0u, k1 = ((t1 & 15u) << 2) | 1u;
        uint32_t k2 = ((t2 & 15u) << 2) | 2u, k3 = ((t3 & 15u) << 2) | 3u;
        sort4_desc(k0, k1, k2, k3);
        const uint32_t K = k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);
        const uint32_t Fs = (K >> 2) & 0x0F0F0F0Fu;               // f(1)..f(4) in bytes 0..3
        const uint32_t w0123 = ((Fs << 8) | 16u) - Fs;            // bytes: 16-f1, f1-f2, f2-f3, f3-f4 (no borrows)
        const uint32_t code = ((K & 0x03030303u) * 0x01041040u) >> 24;   // t0 | t1<<2 | t2<<4 | t3<<6
        const uint32_t cell = ((t0 >> 4) << 12) | ((t1 >> 4) << 8) | ((t2 >> 4) << 4) | (t3 >> 4);
        const uint32_t req = (cell << 16) | (code << 8) | (Fs >> 24);    // low byte = w4 = f(4)
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            // ---- quad side: fetch my 16 B of sample s's cell, weight my 4 corners ----
            const uint32_t rq = __shfl_sync(FULL, req, s, 4);
            const uint32_t rw = __shfl_sync(FULL, w0123, s, 4);
            uint64_t addr;                                        // one IMAD.WIDE instead of a 64-bit add chain
            asm("mad.wide.u32 %0, %1, 64, %2;" : "=l"(addr) : "r"(rq >> 16), "l"(cells_lane));
            const uint4 d = __ldg(reinterpret_cast<const uint4 *>(addr));
            const uint32_t sel = s_sel_lane[((rq >> 8) & 0xFFu) * 4];
            uint32_t wp;                                          // byte 4 of {rw,rq} = w4
            // raw prmt.b32: __byte_perm() masks the selector to 3 bits per nibble and
            // would drop the sign-replicate bit the "not on the path" nibble 0xC relies on
            asm("prmt.b32 %0, %1, %2, %3;" : "=r"(wp) : "r"(rw), "r"(rq), "r"(sel));
            acc[s][subpixel_perm<2>(r, 0)] = __dp4a(d.x, wp, acc[s][subpixel_perm<2>(r, 0)]);
            acc[s][subpixel_perm<2>(r, 1)] = __dp4a(d.y, wp, acc[s][subpixel_perm<2>(r, 1)]);
            acc[s][subpixel_perm<2>(r, 2)] = __dp4a(d.z, wp, acc[s][subpixel_perm<2>(r, 2)]);
            acc[s][subpixel_perm<2>(r, 3)] = __dp4a(d.w, wp, acc[s][subpixel_perm<2>(r, 3)]);
        }
    }
}

template <int CT>
__global__ void __launch_bounds__(Q2_THREADS, 2)
stage_last2_quad_kernel(const __grid_constant__ StageArgs a)
{
    constexpr int P = tile_pitch<CT>();
    __shared__ __align__(16) uint8_t s_in[(Q2_TH + 4) * P];
    __shared__ __align__(16) uint8_t s_out[2 * Q2_TH * Q2_OP];
    __shared__ __align__(8) uint16_t s_sel[1024];

    build_selector_table(s_sel);                 // published by the first __syncthreads below
    const int C = CT > 0 ? CT : a.C;
    const int WC = a.W * C;
    const int tiles_x = (WC + TWB - 1) / TWB;
    const int tiles_y = (a.H + Q2_TH - 1) / Q2_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp % 3) * 32 + lane;
    const int rp = warp / 3;
    const int ql = lane & 3;
    const bool la = (ql & 2) != 0, lb = (ql & 1) != 0;
    const int ocol0 = lx + C * (lx / C);             // (2*(lx/C))*C + lx%C
    const int cols = TWB + 4 * C;
    const int oWC = 2 * WC;                          // output row pitch in bytes
    const uint32_t bias_total = (uint32_t)a.n_modes * 4u * 2048u;   // 16*128 per interpolation
    const uint32_t den = 16u * a.n_modes;
    const uint32_t magic = rhe_magic(den);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long tr = tile / tiles_x;
        const int ty = (int)(tr % tiles_y);
        const int n = (int)(tr / tiles_y);
        const int y0 = ty * Q2_TH, X0 = tx * TWB;
        const uint8_t *__restrict__ img = a.in + (size_t)n * a.H * WC;

        __syncthreads();                       // previous tile's s_in / s_out fully consumed
        fill_tile(s_in, P, Q2_TH + 4, cols, img, a.H, C, WC, y0, X0);
        __syncthreads();

#pragma unroll 1
        for (int ly = rp; ly < Q2_TH; ly += Q2_RW) {
            const uint8_t *sp = s_in + (ly + 2) * P + lx + 2 * C;
            uint32_t acc[4][4];
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[s][j] = 0u;
            for (int m = 0; m < a.n_modes; ++m) {
                const uint8_t *__restrict__ cells = a.lut_alt[m];
                switch (a.modes[m]) {
                case 's': quad_mode<'s', CT>(sp, C, cells, ql, s_sel + ql, acc); break;
                case 'd': quad_mode<'d', CT>(sp, C, cells, ql, s_sel + ql, acc); break;
                default: quad_mode<'y', CT>(sp, C, cells, ql, s_sel + ql, acc); break;
                }
            }
            // quad butterfly: lane ql ends with the complete sums of sample slot ql
            uint32_t tot[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t keep01 = lb ? acc[1][j] : acc[0][j];
                const uint32_t send01 = lb ? acc[0][j] : acc[1][j];
                const uint32_t keep23 = lb ? acc[3][j] : acc[2][j];
                const uint32_t send23 = lb ? acc[2][j] : acc[3][j];
                const uint32_t k01 = keep01 + __shfl_xor_sync(FULL, send01, 1);
                const uint32_t k23 = keep23 + __shfl_xor_sync(FULL, send23, 1);
                const uint32_t keep = la ? k23 : k01;
                const uint32_t send = la ? k01 : k23;
                tot[j] = keep + __shfl_xor_sync(FULL, send, 2);
            }
            uint8_t *so = s_out + (2 * ly) * Q2_OP + ocol0;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int S = (int)tot[u * 2 + v] - (int)bias_total;
                    so[u * Q2_OP + v * C] = (uint8_t)rhe_div_clamp_u8_magic(S, den, magic);
                }
        }
        __syncthreads();

        store_out_tile<2>(s_out, a.out, n, a.H, oWC, y0, X0);
    }
}

// ---------------------------------------------------------------------------
// K1d: last stage, up = 2, cell-major LUT, one lane per sample ("owner-only").
// The sample's lane fetches its whole 64-byte cell with two 256-bit loads
// (LDG.E.256: the a=0 and a=1 halves), builds the four per-row weight words with
// four PRMTs from one 8-byte selector fetch and folds 16 dp4a.  No shuffles, no
// cross-lane reduction: ~40 % fewer instructions than K1c, but MEASURED SLOWER on
// B200 (7.73 ms vs 6.83 ms per 16 x 1080p): a 256-bit load whose 32 lanes touch 32
// different lines costs ~1 L1 cycle per lane even on hits, so the L1 data pipe, not
// the ALU, bounds it.  Kept as a selectable cross-check (MULUT_KERNEL_TILED_CELL).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ldg256(const uint8_t *p, uint32_t (&w)[8])
{
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
                 : "l"(p));
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

template <char MODE, int CT>
__device__ __forceinline__ void cell_mode(const uint8_t *__restrict__ sp, int C, const uint8_t *__restrict__ cells,
                                          const uint16_t *__restrict__ s_sel, uint32_t (&acc)[4])
{
    constexpr int P = tile_pitch<CT>();
    const int Cc = CT > 0 ? CT : C;
    const uint32_t t0 = sp[0];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t1 = sp[tap_off(MODE, r, 1, true) * P + tap_off(MODE, r, 1, false) * Cc];
        const uint32_t t2 = sp[tap_off(MODE, r, 2, true) * P + tap_off(MODE, r, 2, false) * Cc];
        const uint32_t t3 = sp[tap_off(MODE, r, 3, true) * P + tap_off(MODE, r, 3, false) * Cc];
        uint32_t k0 = ((t0 & 15u) << 2) | 0u, k1 = ((t1 & 15u) << 2) | 1u;
        uint32_t k2 = ((t2 & 15u) << 2) | 2u, k3 = ((t3 & 15u) << 2) | 3u;
        sort4_desc(k0, k1, k2, k3);
        const uint32_t K = k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);
        const uint32_t Fs = (K >> 2) & 0x0F0F0F0Fu;               // f(1)..f(4) in bytes 0..3
        const uint32_t w0123 = ((Fs << 8) | 16u) - Fs;            // bytes: 16-f1, f1-f2, f2-f3, f3-f4
        const uint32_t w4 = Fs >> 24;                             // f(4)
        const uint32_t code = ((K & 0x03030303u) * 0x01041040u) >> 24;
        const uint32_t cell = ((t0 >> 4) << 12) | ((t1 >> 4) << 8) | ((t2 >> 4) << 4) | (t3 >> 4);
        uint64_t addr;
        asm("mad.wide.u32 %0, %1, 64, %2;" : "=l"(addr) : "r"(cell), "l"(cells));
        uint32_t A0[8], A1[8];
        ldg256(reinterpret_cast<const uint8_t *>(addr), A0);       // rows (a,b) = 00, 01
        ldg256(reinterpret_cast<const uint8_t *>(addr) + 32, A1);  // rows (a,b) = 10, 11
        const uint2 sel = *reinterpret_cast<const uint2 *>(s_sel + code * 4);   // selectors of rows 0..3
        const uint32_t wp0 = prmt(w0123, w4, sel.x), wp1 = prmt(w0123, w4, sel.x >> 16);
        const uint32_t wp2 = prmt(w0123, w4, sel.y), wp3 = prmt(w0123, w4, sel.y >> 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t v = acc[subpixel_perm<2>(r, j)];
            v = __dp4a(A0[j], wp0, v);
            v = __dp4a(A0[4 + j], wp1, v);
            v = __dp4a(A1[j], wp2, v);
            v = __dp4a(A1[4 + j], wp3, v);
            acc[subpixel_perm<2>(r, j)] = v;
        }
    }
}

template <int CT>
__global__ void __launch_bounds__(Q2_THREADS, 2)
stage_last2_cell_kernel(const __grid_constant__ StageArgs a)
{
    constexpr int P = tile_pitch<CT>();
    __shared__ __align__(16) uint8_t s_in[(Q2_TH + 4) * P];
    __shared__ __align__(16) uint8_t s_out[2 * Q2_TH * Q2_OP];
    __shared__ __align__(8) uint16_t s_sel[1024];

    build_selector_table(s_sel);                 // published by the first __syncthreads below
    const int C = CT > 0 ? CT : a.C;
    const int WC = a.W * C;
    const int tiles_x = (WC + TWB - 1) / TWB;
    const int tiles_y = (a.H + Q2_TH - 1) / Q2_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp % 3) * 32 + lane;
    const int rp = warp / 3;
    const int ocol0 = lx + C * (lx / C);
    const int cols = TWB + 4 * C;
    const int oWC = 2 * WC;
    const uint32_t bias_total = (uint32_t)a.n_modes * 4u * 2048u;
    const uint32_t den = 16u * a.n_modes;
    const uint32_t magic = rhe_magic(den);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long tr = tile / tiles_x;
        const int ty = (int)(tr % tiles_y);
        const int n = (int)(tr / tiles_y);
        const int y0 = ty * Q2_TH, X0 = tx * TWB;
        const uint8_t *__restrict__ img = a.in + (size_t)n * a.H * WC;

        __syncthreads();
        fill_tile(s_in, P, Q2_TH + 4, cols, img, a.H, C, WC, y0, X0);
        __syncthreads();

#pragma unroll 1
        for (int ly = rp; ly < Q2_TH; ly += Q2_RW) {
            const uint8_t *sp = s_in + (ly + 2) * P + lx + 2 * C;
            uint32_t acc[4] = {0u, 0u, 0u, 0u};
            for (int m = 0; m < a.n_modes; ++m) {
                const uint8_t *__restrict__ cells = a.lut_alt[m];
                switch (a.modes[m]) {
                case 's': cell_mode<'s', CT>(sp, C, cells, s_sel, acc); break;
                case 'd': cell_mode<'d', CT>(sp, C, cells, s_sel, acc); break;
                default: cell_mode<'y', CT>(sp, C, cells, s_sel, acc); break;
                }
            }
            uint8_t *so = s_out + (2 * ly) * Q2_OP + ocol0;
#pragma unroll
            for (int u = 0; u < 2; ++u)
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                    const int S = (int)acc[u * 2 + v] - (int)bias_total;
                    so[u * Q2_OP + v * C] = (uint8_t)rhe_div_clamp_u8_magic(S, den, magic);
                }
        }
        __syncthreads();
        store_out_tile<2>(s_out, a.out, n, a.H, oWC, y0, X0);
    }
}

// ---------------------------------------------------------------------------
// K1e: last stage, up = 4, cell-major LUT (256 B per cell), quad-cooperative.
// A cell = 4 row-blocks l = (a,b) of 64 B; a row-block = 16 words, word = the four
// (c,d) corners of ONE output column (dp4a operand).  The 16 output columns are
// stored in ROTATION-ORBIT order: lane q of the quad owns orbit q of the 4x4
// sub-pixel block, (u,v) -> (v,3-u) walks an orbit, so for rotation r the word i of
// a lane lands on the lane's own position (i+r)&3: static register indexing, no
// cross-lane reduction.  Per interpolation the quad fetches the 3 row-blocks the
// simplex walk can touch (rows 00 and 11, and 10 or 01 depending on whether tap a
// or tap b has the larger fraction): 3 x (4 lanes x LDG.128).
//   byte[cell*256 + l*64 + q*16 + i*4 + cd] = LUT[v(l,cd)][col(q,i)] + 128
// ---------------------------------------------------------------------------
// up = 3 (K1e3) uses the same 256-B cells and the same fetch: the 3x3 block has a corner orbit (lane 0), an edge
// orbit (lane 1) and the rotation-invariant centre (lane 2, the same column in all four words, so that every
// rotation adds it to every position of the lane: any one of the four accumulators holds it); lane 3 carries
// nothing.  7 of 16 table bytes are padding: the price of reusing the x4 gather, which is bound by requests
// and L2 bytes per CELL, not by the bytes used.
template <int UP>
__host__ __device__ constexpr int orbit_start_u(int q) { return UP == 4 ? (q == 3 ? 1 : 0) : (q == 2 ? 1 : 0); }
template <int UP>
__host__ __device__ constexpr int orbit_start_v(int q) { return UP == 4 ? (q == 3 ? 1 : q) : (q == 2 ? 1 : q == 3 ? 1 : q); }

template <int UP>
__global__ void build_cell_major4_kernel(const int8_t *__restrict__ lut, uint8_t *__restrict__ cells)
{
    const size_t total = (size_t)65536 * 256;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int cd = idx & 3, i = (idx >> 2) & 3, q = (idx >> 4) & 3, l = (idx >> 6) & 3;
        const int cell = (int)(idx >> 8);
        const int ma = cell >> 12, mb = (cell >> 8) & 15, mc = (cell >> 4) & 15, md = cell & 15;
        const int v = (ma + (l >> 1)) * 4913 + (mb + (l & 1)) * 289 + (mc + (cd >> 1)) * 17 + (md + (cd & 1));
        int u = orbit_start_u<UP>(q), w = orbit_start_v<UP>(q);
        for (int k = 0; k < i; ++k) { int t = u; u = w; w = UP - 1 - t; }
        cells[idx] = (UP == 3 && q == 3) ? (uint8_t)128 : (uint8_t)((int)lut[(size_t)v * (UP * UP) + u * UP + w] + 128);
    }
}

template <int UP> __host__ __device__ constexpr int q4_op() { return UP * TWB; }

template <char MODE, int CT>
__device__ __forceinline__ void quad4_mode(const uint8_t *__restrict__ sp, int C,
                                           const uint8_t *__restrict__ cells_lane,
                                           const uint16_t *__restrict__ s_sel, uint32_t (&acc)[4][4])
{
    constexpr int P = tile_pitch<CT>();
    const int Cc = CT > 0 ? CT : C;
    const uint32_t t0 = sp[0];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const uint32_t t1 = sp[tap_off(MODE, r, 1, true) * P + tap_off(MODE, r, 1, false) * Cc];
        const uint32_t t2 = sp[tap_off(MODE, r, 2, true) * P + tap_off(MODE, r, 2, false) * Cc];
        const uint32_t t3 = sp[tap_off(MODE, r, 3, true) * P + tap_off(MODE, r, 3, false) * Cc];
        uint32_t k0 = ((t0 & 15u) << 2) | 0u, k1 = ((t1 & 15u) << 2) | 1u;
        uint32_t k2 = ((t2 & 15u) << 2) | 2u, k3 = ((t3 & 15u) << 2) | 3u;
        sort4_desc(k0, k1, k2, k3);
        const uint32_t K = k0 | (k1 << 8) | (k2 << 16) | (k3 << 24);
        const uint32_t Fs = (K >> 2) & 0x0F0F0F0Fu;
        const uint32_t w0123 = ((Fs << 8) | 16u) - Fs;
        const uint32_t code = ((K & 0x03030303u) * 0x01041040u) >> 24;
        const uint32_t cell = ((t0 >> 4) << 12) | ((t1 >> 4) << 8) | ((t2 >> 4) << 4) | (t3 >> 4);
        const uint32_t req = (cell << 16) | (code << 8) | (Fs >> 24);
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint32_t rq = __shfl_sync(FULL, req, s, 4);
            const uint32_t rw = __shfl_sync(FULL, w0123, s, 4);
            const uint2 sel = *reinterpret_cast<const uint2 *>(s_sel + ((rq >> 8) & 0xFFu) * 4);
            // row 01 is off the path (all nibbles 0xC) exactly when tap a precedes tap b: then use row 10
            const bool a_first = (sel.x >> 16) == 0xCCCCu;
            const uint32_t selm = a_first ? sel.y : (sel.x >> 16);
            uint64_t addr;
            asm("mad.wide.u32 %0, %1, 256, %2;" : "=l"(addr) : "r"(rq >> 16), "l"(cells_lane));
            const uint8_t *cp = reinterpret_cast<const uint8_t *>(addr);
            const uint4 d0 = __ldg(reinterpret_cast<const uint4 *>(cp));
            const uint4 dm = __ldg(reinterpret_cast<const uint4 *>(cp + (a_first ? 128 : 64)));
            const uint4 d3 = __ldg(reinterpret_cast<const uint4 *>(cp + 192));
            const uint32_t wp0 = prmt(rw, rq, sel.x), wpm = prmt(rw, rq, selm), wp3 = prmt(rw, rq, sel.y >> 16);
            const uint32_t x0[4] = {d0.x, d0.y, d0.z, d0.w}, xm[4] = {dm.x, dm.y, dm.z, dm.w}, x3[4] = {d3.x, d3.y, d3.z, d3.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t v = acc[s][(i + r) & 3];
                v = __dp4a(x0[i], wp0, v);
                v = __dp4a(xm[i], wpm, v);
                v = __dp4a(x3[i], wp3, v);
                acc[s][(i + r) & 3] = v;
            }
        }
    }
}

template <int CT, int UP>
__global__ void __launch_bounds__(Q2_THREADS, 2)
stage_last4_quad_kernel(const __grid_constant__ StageArgs a)
{
    static_assert(UP == 3 || UP == 4, "K1e serves up = 4 and, with padded cells, up = 3");
    constexpr int P = tile_pitch<CT>();
    constexpr int Q4_OP = q4_op<UP>();
    extern __shared__ __align__(16) uint8_t q4_smem[];
    uint8_t *s_out = q4_smem;                                  // UP*TH x UP*TWB bytes
    uint8_t *s_in = q4_smem + UP * Q2_TH * Q4_OP;
    __shared__ __align__(8) uint16_t s_sel[1024];

    build_selector_table(s_sel);
    const int C = CT > 0 ? CT : a.C;
    const int WC = a.W * C;
    const int tiles_x = (WC + TWB - 1) / TWB;
    const int tiles_y = (a.H + Q2_TH - 1) / Q2_TH;
    const long long n_tiles = (long long)a.N * tiles_y * tiles_x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int lx = (warp % 3) * 32 + lane;
    const int rp = warp / 3;
    const int ql = lane & 3;
    const int cols = TWB + 4 * C;
    const int oWC = UP * WC;
    const uint32_t bias_total = (uint32_t)a.n_modes * 4u * 2048u;
    const uint32_t den = 16u * a.n_modes;
    const uint32_t magic = rhe_magic(den);
    // my orbit's four sub-pixel positions (u,v), in rotation order (up = 3: lane 2 holds the centre in every
    // word - only word 0 is stored - and lane 3 nothing)
    int pu[4], pv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int u = UP == 4 ? (ql == 3 ? 1 : 0) : (ql >= 2 ? 1 : 0), v = UP == 4 ? (ql == 3 ? 1 : ql) : (ql >= 2 ? 1 : ql);
        for (int k = 0; k < i; ++k) { int t = u; u = v; v = UP - 1 - t; }
        pu[i] = u; pv[i] = v;
    }
    const int n_store = UP == 4 ? 4 : (ql < 2 ? 4 : ql == 2 ? 1 : 0);     // accumulators of this lane that are outputs

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int tx = (int)(tile % tiles_x);
        const long long tr = tile / tiles_x;
        const int ty = (int)(tr % tiles_y);
        const int n = (int)(tr / tiles_y);
        const int y0 = ty * Q2_TH, X0 = tx * TWB;
        const uint8_t *__restrict__ img = a.in + (size_t)n * a.H * WC;

        __syncthreads();
        fill_tile(s_in, P, Q2_TH + 4, cols, img, a.H, C, WC, y0, X0);
        __syncthreads();

#pragma unroll 1
        for (int ly = rp; ly < Q2_TH; ly += Q2_RW) {
            const uint8_t *sp = s_in + (ly + 2) * P + lx + 2 * C;
            uint32_t acc[4][4];
#pragma unroll
            for (int s = 0; s < 4; ++s)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[s][j] = 0u;
            for (int m = 0; m < a.n_modes; ++m) {
                const uint8_t *__restrict__ cells_lane = a.lut_alt[m] + ql * 16;
                switch (a.modes[m]) {
                case 's': quad4_mode<'s', CT>(sp, C, cells_lane, s_sel, acc); break;
                case 'd': quad4_mode<'d', CT>(sp, C, cells_lane, s_sel, acc); break;
                default: quad4_mode<'y', CT>(sp, C, cells_lane, s_sel, acc); break;
                }
            }
            // lane ql holds orbit ql of the 4x4 block of each of the quad's four samples
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int sx = (lx & ~3) + s;                          // local byte column of sample s
                uint8_t *so = s_out + (UP * ly) * Q4_OP + sx + (UP - 1) * C * (sx / C);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int S = (int)acc[s][i] - (int)bias_total;
                    if (i < n_store) so[pu[i] * Q4_OP + pv[i] * C] = (uint8_t)rhe_div_clamp_u8_magic(S, den, magic);
                }
            }
        }
        __syncthreads();
        store_out_tile<UP>(s_out, a.out, n, a.H, oWC, y0, X0);
    }
}

// ---------------------------------------------------------------------------
// LUT re-layouts
// ---------------------------------------------------------------------------
// up = 1: the "alt" table is the int8 LUT itself padded to LUT1_SMEM bytes.
// up = 2: cell-major, 64 B per cell:  byte[cell*64 + l*16 + j*4 + cd] =
//         LUT[(ma+la)*17^3 + (mb+lb)*17^2 + (mc+c)*17 + (md+d)][j] + 128,
//         l = la*2+lb, cd = c*2+d, cell = ma<<12 | mb<<8 | mc<<4 | md.
__global__ void build_cell_major2_kernel(const int8_t *__restrict__ lut, uint8_t *__restrict__ cells)
{
    const size_t total = (size_t)65536 * 64;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int cd = i & 3, j = (i >> 2) & 3, l = (i >> 4) & 3;
        const int cell = (int)(i >> 6);
        const int ma = cell >> 12, mb = (cell >> 8) & 15, mc = (cell >> 4) & 15, md = cell & 15;
        const int v = (ma + (l >> 1)) * 4913 + (mb + (l & 1)) * 289 + (mc + (cd >> 1)) * 17 + (md + (cd & 1));
        cells[i] = (uint8_t)((int)lut[(size_t)v * 4 + j] + 128);
    }
}

size_t cell_major_bytes(int up)
{
    if (up == 1) return LUT1_SMEM + stage1_pair_bytes();     // [padded int8 table | a-paired 16-bit table of K1h]
    if (up == 2) return (size_t)65536 * 64;
    if (up == 4 || up == 3) return (size_t)65536 * 256;     // up = 3: the x4 cell with 9 of 16 columns used (K1e3)
    return 0;
}

int build_cell_major(const int8_t *d_lut, uint8_t *d_alt, int up, cudaStream_t stream)
{
    if (up == 1) {
        MULUT_CUDA(cudaMemsetAsync(d_alt, 0, LUT1_SMEM, stream));
        MULUT_CUDA(cudaMemcpyAsync(d_alt, d_lut, LUT1_ROWS, cudaMemcpyDeviceToDevice, stream));
        return build_pair_table(d_lut, d_alt + LUT1_SMEM, stream);
    }
    if (up == 2) {
        build_cell_major2_kernel<<<1024, 256, 0, stream>>>(d_lut, d_alt);
        MULUT_CUDA(cudaGetLastError());
        return MULUT_OK;
    }
    if (up == 4 || up == 3) {
        if (up == 4) build_cell_major4_kernel<4><<<2048, 256, 0, stream>>>(d_lut, d_alt);
        else build_cell_major4_kernel<3><<<2048, 256, 0, stream>>>(d_lut, d_alt);
        MULUT_CUDA(cudaGetLastError());
        return MULUT_OK;
    }
    return MULUT_OK;
}

bool tiled_supported(int up, int interval, int n_modes)
{
    return interval == 4 && n_modes >= 1 && up >= 1 && up <= 4;
}

// ---------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------
template <typename K>
static int set_smem(K kernel, size_t bytes)
{
    MULUT_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return MULUT_OK;
}

template <int CT>
static int launch_smem_stage(const StageArgs &a, int16_t *partial, cudaStream_t stream)
{
    constexpr int P = tile_pitch<CT>();
    const size_t smem = LUT1_SMEM + (size_t)(S1_TH + 4) * P;
    {
        int rc = set_smem(stage_smem_kernel<CT>, smem);
        if (rc) return rc;
    }
    int per_sm = 0;
    MULUT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_smem_kernel<CT>, S1_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    int resident = per_sm * a.num_sms;
    int ctas_per_mode = resident / a.n_modes;
    if (ctas_per_mode < 1) ctas_per_mode = 1;
    const int WC = a.W * a.C;
    const long long n_tiles = (long long)a.N * ((a.H + S1_TH - 1) / S1_TH) * ((WC + TWB - 1) / TWB);
    if (ctas_per_mode > n_tiles) ctas_per_mode = (int)n_tiles;
    stage_smem_kernel<CT><<<ctas_per_mode * a.n_modes, S1_THREADS, smem, stream>>>(a, partial, ctas_per_mode);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

template <int CT>
static int launch_quad_stage(const StageArgs &a, bool owner_only, cudaStream_t stream)
{
    int per_sm = 0;
    if (owner_only)
        MULUT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_last2_cell_kernel<CT>, Q2_THREADS, 0));
    else
        MULUT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_last2_quad_kernel<CT>, Q2_THREADS, 0));
    if (per_sm < 1) per_sm = 1;
    const int WC = a.W * a.C;
    const long long n_tiles = (long long)a.N * ((a.H + Q2_TH - 1) / Q2_TH) * ((WC + TWB - 1) / TWB);
    long long grid = (long long)per_sm * a.num_sms;
    if (grid > n_tiles) grid = n_tiles;
    if (owner_only)
        stage_last2_cell_kernel<CT><<<(unsigned)grid, Q2_THREADS, 0, stream>>>(a);
    else
        stage_last2_quad_kernel<CT><<<(unsigned)grid, Q2_THREADS, 0, stream>>>(a);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

template <int CT, int UP>
static int launch_quad4_stage(const StageArgs &a, cudaStream_t stream)
{
    constexpr int P = tile_pitch<CT>();
    const size_t smem = (size_t)UP * Q2_TH * q4_op<UP>() + (size_t)(Q2_TH + 4) * P;
    {
        int rc = set_smem(stage_last4_quad_kernel<CT, UP>, smem);
        if (rc) return rc;
    }
    int per_sm = 0;
    MULUT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, stage_last4_quad_kernel<CT, UP>, Q2_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    const int WC = a.W * a.C;
    const long long n_tiles = (long long)a.N * ((a.H + Q2_TH - 1) / Q2_TH) * ((WC + TWB - 1) / TWB);
    long long grid = (long long)per_sm * a.num_sms;
    if (grid > n_tiles) grid = n_tiles;
    stage_last4_quad_kernel<CT, UP><<<(unsigned)grid, Q2_THREADS, smem, stream>>>(a);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

// partial: workspace of n_modes * N*H*W*C int16 (only used for up == 1)
int launch_stage_tiled_ws(const StageArgs &a, int up, int16_t *partial, cudaStream_t stream, int *launches,
                          Prof *prof, bool owner_only, const BinPlanArgs *plan)
{
    const size_t total = (size_t)a.N * a.H * a.W * a.C;
    if (total == 0) return MULUT_OK;
    if (!tiled_supported(up, a.interval, a.n_modes)) return 1;
    if (a.C != 1 && a.C != 2 && a.C != 3 && a.C != 4) return 1;
    if (up == 1) {
        prof->begin(MULUT_PROF_SMEM_STAGE, stream);
        int rc = 1;
        if (!a.no_tma && stage1_tma_supported(a, up)) rc = launch_stage1_tma(a, partial, stream);   // K1g
        if (rc == 1)                                                                                // K1a
            rc = a.C == 3 ? launch_smem_stage<3>(a, partial, stream)
               : a.C == 1 ? launch_smem_stage<1>(a, partial, stream)
                          : launch_smem_stage<0>(a, partial, stream);
        prof->end(stream);
        if (rc) return rc;
        size_t blocks = (total / 8 + 255) / 256 + 1;
        const size_t cap = (size_t)a.num_sms * 16;
        if (blocks > cap) blocks = cap;
        BinPlanArgs pa;
        memset(&pa, 0, sizeof pa);
        if (plan && plan->ctl) {                       // the next stage is K1f: histogram + plan ride on K1b
            pa = *plan;
            MULUT_CUDA(cudaMemsetAsync(pa.ctl, 0, sizeof(BinCtl), stream));
        }
        prof->begin(MULUT_PROF_COMBINE, stream);
        combine_kernel<<<(unsigned)blocks, 256, 0, stream>>>(partial, a.out, total, a.n_modes, a.last, pa);
        prof->end(stream);
        MULUT_CUDA(cudaGetLastError());
        *launches += 2;
        return MULUT_OK;
    }
    if (!a.last) return 1;
    prof->begin(MULUT_PROF_LAST_TILED, stream);
    int rc;
    if (up == 4)
        rc = a.C == 3 ? launch_quad4_stage<3, 4>(a, stream)
           : a.C == 1 ? launch_quad4_stage<1, 4>(a, stream)
                      : launch_quad4_stage<0, 4>(a, stream);
    else if (up == 3)
        rc = a.C == 3 ? launch_quad4_stage<3, 3>(a, stream)
           : a.C == 1 ? launch_quad4_stage<1, 3>(a, stream)
                      : launch_quad4_stage<0, 3>(a, stream);
    else
        rc = a.C == 3 ? launch_quad_stage<3>(a, owner_only, stream)
           : a.C == 1 ? launch_quad_stage<1>(a, owner_only, stream)
                      : launch_quad_stage<0>(a, owner_only, stream);
    prof->end(stream);
    if (rc) return rc;
    *launches += 1;
    return MULUT_OK;
}

}  // namespace mulut
