// Shared device helpers and host-side error plumbing for libmulut_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/mulut.h"

namespace mulut {

// ---------------------------------------------------------------------------
// host side: thread-local last error
// ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define MULUT_CUDA(call)                                                              \
    do {                                                                              \
        cudaError_t e__ = (call);                                                     \
        if (e__ != cudaSuccess) return ::mulut::cuda_fail(e__, #call, __FILE__, __LINE__); \
    } while (0)

// Tap offsets (dy, dx) of taps a,b,c,d per mode, sr/4_test_lut.py:18-51.
// Returns false for an unknown mode (reference raises ValueError, :52-54).
inline bool mode_taps(char mode, int dy[4], int dx[4])
{
    static const int S[2][4] = {{0, 0, 1, 1}, {0, 1, 0, 1}};
    static const int D[2][4] = {{0, 0, 2, 2}, {0, 2, 0, 2}};
    static const int Y[2][4] = {{0, 1, 1, 2}, {0, 1, 2, 1}};
    const int(*t)[4] = nullptr;
    switch (mode) {
    case 's': t = S; break;
    case 'd': t = D; break;
    case 'y': t = Y; break;
    default: return false;
    }
    for (int k = 0; k < 4; ++k) { dy[k] = t[0][k]; dx[k] = t[1][k]; }
    return true;
}
inline int mode_pad(char mode) { return mode == 's' ? 1 : 2; }   // sr/4_test_lut.py:289-292

// ---------------------------------------------------------------------------
// device side
// ---------------------------------------------------------------------------
// Rotated tap tables: the reference's rot90 -> pad(edge) -> interp -> rot90-back
// (sr/4_test_lut.py:293-298,235) equals sampling the UN-rotated image at tap
// offsets rotated r times by (dy,dx) <- (dx,-dy), with clamped coordinates.
struct TapTable {
    int8_t dy[MULUT_MAX_MODES][4][4];   // [mode][rot][tap]
    int8_t dx[MULUT_MAX_MODES][4][4];
};

inline bool build_tap_table(const char *modes, int n_modes, TapTable *tt)
{
    for (int m = 0; m < n_modes; ++m) {
        int dy[4], dx[4];
        if (!mode_taps(modes[m], dy, dx)) return false;
        for (int r = 0; r < 4; ++r)
            for (int k = 0; k < 4; ++k) {
                int a = dy[k], b = dx[k];
                for (int i = 0; i < r; ++i) { int t = a; a = b; b = -t; }
                tt->dy[m][r][k] = (int8_t)a;
                tt->dx[m][r][k] = (int8_t)b;
            }
    }
    return true;
}

// LUT column j = u*UP+v of rotation r lands at sub-pixel (u,v) rotated r times
// by (u,v) <- (v, UP-1-u)   (np.rot90(out, 4-r) of the pixel-shuffled block).
template <int UP>
__host__ __device__ constexpr int subpixel_perm(int r, int j)
{
    int u = j / UP, v = j % UP;
    for (int i = 0; i < r; ++i) { int t = u; u = v; v = UP - 1 - t; }
    return u * UP + v;
}

__device__ __forceinline__ void cswap_desc(uint32_t &a, uint32_t &b)
{
    uint32_t hi = max(a, b), lo = min(a, b);
    a = hi; b = lo;
}
// 5-comparator network, descending.
__device__ __forceinline__ void sort4_desc(uint32_t &k0, uint32_t &k1, uint32_t &k2, uint32_t &k3)
{
    cswap_desc(k0, k1); cswap_desc(k2, k3);
    cswap_desc(k0, k2); cswap_desc(k1, k3);
    cswap_desc(k1, k2);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

// clamp(round_half_even(num/den), 0, 255), den > 0; np.round semantics of
// sr/4_test_lut.py:300-302 in integers (negative -> 0 after the clip).
__device__ __forceinline__ uint32_t rhe_div_clamp_u8(int num, uint32_t den)
{
    if (num <= 0) return 0u;
    uint32_t n = (uint32_t)num;
    uint32_t qd = n / den, rm = n - qd * den;
    qd += (2u * rm > den) || ((2u * rm == den) && (qd & 1u));
    return min(qd, 255u);
}

// Same with the division replaced by a multiply-high: magic = floor(2^32 / den) + 1 is exact for
// num * den < 2^32 (here num < 2^16: |S| <= 12 * 16 * 127 plus the 127 * den bias).
__device__ __forceinline__ uint32_t rhe_magic(uint32_t den) { return 0xFFFFFFFFu / den + 1u; }
__device__ __forceinline__ uint32_t rhe_div_clamp_u8_magic(int num, uint32_t den, uint32_t magic)
{
    const uint32_t n = (uint32_t)max(num, 0);
    uint32_t qd = __umulhi(n, magic);
    const uint32_t rm = n - qd * den;
    qd += (2u * rm > den) || ((2u * rm == den) && (qd & 1u));
    return min(qd, 255u);
}

// Non-last stage epilogue, 0 <= t <= 254 * den, den even (den = q * 4M): round-half-even(t / den) needs neither the
// clamp at 0 nor the one at 255.  With u = t + den/2, q' = floor(u / den) is the half-up rounding; it is one too many
// exactly on a tie (u % den == 0) whose q' is odd.  Checked exhaustively for M = 1..8 in tests/test_host_logic.py.
__device__ __forceinline__ uint32_t rhe_div_nonneg_magic(uint32_t t, uint32_t den, uint32_t magic)
{
    const uint32_t u = t + (den >> 1);
    const uint32_t qd = __umulhi(u, magic);
    const uint32_t rm = u - qd * den;
    return qd - ((rm == 0u) ? (qd & 1u) : 0u);
}

}  // namespace mulut
