// Floating-point twins of the interpolation:
//   K2  interp_fwd_kernel   MuLUT.InterpTorchBatch forward        sr/model.py:69-287
//   K3  interp_bwd_kernel   its autograd backward (d/d weight scatter-add through
//                           the quantiser's STE + clamp mask, d/d img_in through
//                           the LSB fractions)                    SURVEY.md 8-SPEC
//   interp_pass_f64_kernel  FourSimplexInterpFaster-compatible single pass
//                           (float32 in, float64 out, rotated)    sr/4_test_lut.py:14-237
//
// The reference enumerates 24 strict-inequality cases; that is a descending
// sort of the four fractions with ties broken "higher tap index first".
#include <stdlib.h>

#include "common.cuh"

namespace mulut {

struct F32Args {
    const float *weight;      // (n_rows, up^2) raw parameter
    const float *img;         // (B*C, h+bd, w+bd)
    int n_rows, up, BC, h, w, bd, interval;
    int dy[4], dx[4];
};

// order[rank] = tap; element i precedes j if f_i > f_j or (f_i == f_j and i > j)
__device__ __forceinline__ void sort_taps(const float (&f)[4], int (&order)[4])
{
    int rank[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (i != j) rank[i] += (f[j] > f[i]) || (f[j] == f[i] && j > i);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int p = 0; p < 4; ++p)
            if (rank[i] == p) order[p] = i;
}

// torch.floor_divide / torch.remainder on float32 for a positive divisor
__device__ __forceinline__ void split_msb_lsb(float t, float q, int &m, float &f)
{
    float mod = fmodf(t, q);
    if (mod != 0.f && mod < 0.f) mod += q;       // python-style remainder
    f = mod;
    m = (int)floorf(__fdiv_rn(t, q));         // q is a power of two: exact
}

// model.py:74-76: clamp(round_half_even(w * 127), -127, 127)
__device__ __forceinline__ float quant(float w) { return fminf(fmaxf(rintf(__fmul_rn(w, 127.f)), -127.f), 127.f); }
__device__ __forceinline__ bool quant_pass(float w)
{
    const float r = rintf(__fmul_rn(w, 127.f));
    return r >= -127.f && r <= 127.f;            // clamp backward is inclusive
}

struct Simplex {
    int v[5];          // LUT rows
    float w[5];        // weights (q-f1, f1-f2, f2-f3, f3-f4, f4)
    int order[4];      // taps by descending fraction
};

__device__ __forceinline__ void simplex_setup(const F32Args &a, const float *__restrict__ px, int pitch,
                                              Simplex &s)
{
    const float q = (float)(1 << a.interval);
    const int L = (1 << (8 - a.interval)) + 1;
    const int stride[4] = {L * L * L, L * L, L, 1};
    float f[4];
    int v0 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float t = px[a.dy[k] * pitch + a.dx[k]];
        int m;
        split_msb_lsb(t, q, m, f[k]);
        v0 += m * stride[k];
    }
    sort_taps(f, s.order);
    float fs[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        fs[p] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (s.order[p] == k) fs[p] = f[k];
    }
    s.v[0] = v0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        int st = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (s.order[p] == k) st = stride[k];
        s.v[p + 1] = s.v[p] + st;
    }
    // memory safety only: in-range inputs (0..255) never clamp
#pragma unroll
    for (int p = 0; p < 5; ++p) s.v[p] = min(max(s.v[p], 0), a.n_rows - 1);
    s.w[0] = q - fs[0];
    s.w[1] = fs[0] - fs[1];
    s.w[2] = fs[1] - fs[2];
    s.w[3] = fs[2] - fs[3];
    s.w[4] = fs[3];
}

template <int UP>
__global__ void __launch_bounds__(256) interp_fwd_kernel(const __grid_constant__ F32Args a, float *__restrict__ out)
{
    constexpr int UP2 = UP * UP;
    const int pitch = a.w + a.bd;
    const size_t plane_in = (size_t)(a.h + a.bd) * pitch;
    const size_t total = (size_t)a.BC * a.h * a.w;
    const float inv_q = 1.f / (float)(1 << a.interval);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % a.w);
        const size_t r = i / a.w;
        const int y = (int)(r % a.h);
        const size_t bc = r / a.h;
        Simplex s;
        simplex_setup(a, a.img + bc * plane_in + (size_t)y * pitch + x, pitch, s);
        float o[UP2];
#pragma unroll
        for (int j = 0; j < UP2; ++j) o[j] = 0.f;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float *__restrict__ row = a.weight + (size_t)s.v[k] * UP2;
#pragma unroll
            for (int j = 0; j < UP2; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(s.w[k], quant(__ldg(row + j))));
        }
        float *__restrict__ op = out + (bc * a.h * UP + (size_t)y * UP) * ((size_t)a.w * UP) + (size_t)x * UP;
#pragma unroll
        for (int u = 0; u < UP; ++u)
#pragma unroll
            for (int v = 0; v < UP; ++v) op[(size_t)u * a.w * UP + v] = o[u * UP + v] * inv_q;
    }
}

__device__ __forceinline__ void red_add_v4(float *addr, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

template <int UP>
__global__ void __launch_bounds__(256)
interp_bwd_kernel(const __grid_constant__ F32Args a, const float *__restrict__ gout,
                  float *__restrict__ gweight, float *__restrict__ gimg)
{
    constexpr int UP2 = UP * UP;
    const int pitch = a.w + a.bd;
    const size_t plane_in = (size_t)(a.h + a.bd) * pitch;
    const size_t total = (size_t)a.BC * a.h * a.w;
    const float inv_q = 1.f / (float)(1 << a.interval);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % a.w);
        const size_t r = i / a.w;
        const int y = (int)(r % a.h);
        const size_t bc = r / a.h;
        const size_t in_off = bc * plane_in + (size_t)y * pitch + x;
        Simplex s;
        simplex_setup(a, a.img + in_off, pitch, s);
        // g = dL/d(sum) = grad_out / q   (out = sum / q, model.py:286)
        float g[UP2];
        const float *__restrict__ gp = gout + (bc * a.h * UP + (size_t)y * UP) * ((size_t)a.w * UP) + (size_t)x * UP;
#pragma unroll
        for (int u = 0; u < UP; ++u)
#pragma unroll
            for (int v = 0; v < UP; ++v) g[u * UP + v] = __ldg(gp + (size_t)u * a.w * UP + v) * inv_q;

        float dot_prev = 0.f;                     // sum_j g[j] * Wq[v_{k-1}][j]
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const float *__restrict__ row = a.weight + (size_t)s.v[k] * UP2;
            float wraw[UP2];
#pragma unroll
            for (int j = 0; j < UP2; ++j) wraw[j] = __ldg(row + j);
            float dot = 0.f;
#pragma unroll
            for (int j = 0; j < UP2; ++j) dot += g[j] * quant(wraw[j]);
            if (gweight) {
                float *__restrict__ gw = gweight + (size_t)s.v[k] * UP2;
                if constexpr (UP2 % 4 == 0) {
#pragma unroll
                    for (int j = 0; j < UP2; j += 4) {
                        float c[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            c[e] = quant_pass(wraw[j + e]) ? 127.f * (s.w[k] * g[j + e]) : 0.f;
                        red_add_v4(gw + j, c[0], c[1], c[2], c[3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < UP2; ++j)
                        if (quant_pass(wraw[j])) atomicAdd(gw + j, 127.f * (s.w[k] * g[j]));
                }
            }
            if (gimg && k > 0) {
                // d(sum)/d f_(k) = Wq[v_k] - Wq[v_{k-1}];  f = t % q has unit slope in t
                int tap_dy = 0, tap_dx = 0;
#pragma unroll
                for (int t = 0; t < 4; ++t)
                    if (s.order[k - 1] == t) { tap_dy = a.dy[t]; tap_dx = a.dx[t]; }
                atomicAdd(gimg + in_off + (size_t)tap_dy * pitch + tap_dx, dot - dot_prev);
            }
            dot_prev = dot;
        }
    }
}

// FourSimplexInterpFaster-compatible pass: float32 arithmetic like numpy's, then
// float64 container, np.rot90(out, rot, [1,2]) and / q (sr/4_test_lut.py:232-236).
template <int UP>
__global__ void __launch_bounds__(256)
interp_pass_f64_kernel(const __grid_constant__ F32Args a, int rot, double *__restrict__ out)
{
    constexpr int UP2 = UP * UP;
    const int pitch = a.w + a.bd;
    const size_t plane_in = (size_t)(a.h + a.bd) * pitch;
    const size_t total = (size_t)a.BC * a.h * a.w;
    const double q = (double)(1 << a.interval);
    const int Hh = a.h * UP, Ww = a.w * UP;
    const int k = ((rot % 4) + 4) % 4;
    const int oH = (k & 1) ? Ww : Hh, oW = (k & 1) ? Hh : Ww;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % a.w);
        const size_t r = i / a.w;
        const int y = (int)(r % a.h);
        const size_t c = r / a.h;
        Simplex s;
        simplex_setup(a, a.img + c * plane_in + (size_t)y * pitch + x, pitch, s);
        float o[UP2];
#pragma unroll
        for (int j = 0; j < UP2; ++j) o[j] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) {
            const float *__restrict__ row = a.weight + (size_t)s.v[kk] * UP2;
#pragma unroll
            for (int j = 0; j < UP2; ++j) o[j] = __fadd_rn(o[j], __fmul_rn(s.w[kk], __ldg(row + j)));
        }
#pragma unroll
        for (int u = 0; u < UP; ++u)
#pragma unroll
            for (int v = 0; v < UP; ++v) {
                const int ii = y * UP + u, jj = x * UP + v;      // position before rotation
                int oi, oj;
                switch (k) {
                case 0: oi = ii; oj = jj; break;
                case 1: oi = Ww - 1 - jj; oj = ii; break;
                case 2: oi = Hh - 1 - ii; oj = Ww - 1 - jj; break;
                default: oi = jj; oj = Hh - 1 - ii; break;
                }
                out[(c * oH + oi) * (size_t)oW + oj] = (double)o[u * UP + v] / q;
            }
    }
}

// ===========================================================================
// K4: one whole stage of MuLUT.forward (sr/model.py:296-310) in ONE kernel per
// direction.  The reference runs, per stage, 12 x { rot90 -> replicate-pad ->
// InterpTorchBatch -> rot90 back -> pred += -> round_func }.  Here every input pixel
// walks its 12 (mode, rotation) interpolations with rotated tap offsets and clamped
// coordinates (no image is rotated or padded), accumulating pred in registers with
// the reference's rounding after every pass, then applies
// x' = round(clamp(pred / avg + bias, 0, 255)) and records the clamp mask.
// The backward kernel mirrors it (all round_func are identity, BPDA model.py:59-67):
// G_pred = G * mask / avg feeds all 12 passes; LUT gradients leave through vector
// red.global.add, input gradients are accumulated per tile in shared memory first.
// ===========================================================================
struct StageF32Args {
    const float *x;                          // (BC, h, w), integer-valued 0..255
    int BC, h, w, n_modes, interval, n_rows;
    float avg, bias;
    int aggregate;                           // backward: 0 direct reds, 1 warp-aggregated, 2 decide from `stats`
    uint32_t *stats;                         // workspace tail: {lanes sharing their first LUT row with another lane, lanes}
    const float *weight[MULUT_MAX_MODES];    // raw parameters (n_rows, up^2)
    const int8_t *wq[MULUT_MAX_MODES];       // quantised once per call: clamp(rint(w*127), +-127), rows of q_pitch(up) bytes
    const uint16_t *wflag[MULUT_MAX_MODES];  // bit j: the clamp passes the gradient of column j (|rint(w*127)| <= 127)
    float *gweight[MULUT_MAX_MODES];         // backward only: accumulated into
    float *replica;                          // backward only: (n_replicas - 1) x n_modes scratch tables, or null
    size_t replica_stride;                   // floats per scratch table (n_rows * up^2 rounded up to 4)
    int n_replicas;                          // 1 = every CTA adds into gweight directly
    int merge;                               // hot rows: lanes of a warp on one row merge before the red (match.any tree)
    TapTable taps;
};

// Quantised-row workspace.  MuLUT.InterpTorchBatch re-quantises the whole table on every call
// (model.py:74-76); K4 does it ONCE per stage and direction into int8 rows, so a vertex row of the
// x4 tables is one 16-byte load instead of 16 scalar fp32 loads + 16 rint/clamp (ncu: the fp32
// version was L1-bound at 92 %, 31 sectors per request).
__host__ __device__ constexpr int q_pitch(int up) { return up == 1 ? 1 : up == 2 ? 4 : up == 3 ? 12 : 16; }
static size_t align16(size_t b) { return (b + 15) / 16 * 16; }
static size_t stage_ws_mode_bytes(int n_rows, int up) { return align16((size_t)n_rows * q_pitch(up)) + align16((size_t)n_rows * 2); }
constexpr size_t STAGE_WS_TAIL = 16;       // row-sharing statistics of the forward pass (see StageF32Args::stats)
// Gradient replicas (backward, hot rows): on natural / smooth patches a few hundred LUT rows take most of the updates and
// the L2 serialises atomics on one address - the same 141 M vector reds that take 0.82 ms on noise took 2.07 ms (1.56 ms
// with the warp merge).  When the forward's row-sharing statistic says so, CTA b adds into table copy b % K4_REPLICAS
// (copy 0 is the caller's gradient buffer, the others live in the workspace, zeroed before and summed into copy 0 after
// the kernel): a quarter of the collisions per address, spread over four times as many L2 slices.
constexpr int K4_REPLICAS = 4;
static size_t replica_table_floats(int n_rows, int up) { return ((size_t)n_rows * up * up + 3) / 4 * 4; }
static size_t stage_ws_replica_bytes(int n_modes, int n_rows, int up)
{
    return (size_t)(K4_REPLICAS - 1) * n_modes * replica_table_floats(n_rows, up) * sizeof(float);
}

template <int UP>
__global__ void __launch_bounds__(256)
quantize_rows_kernel(const float *__restrict__ w, int n_rows, int8_t *__restrict__ q, uint16_t *__restrict__ flag,
                     uint32_t *__restrict__ stats_to_clear)
{
    constexpr int UP2 = UP * UP, RB = q_pitch(UP);
    if (stats_to_clear && blockIdx.x == 0 && threadIdx.x < 4) stats_to_clear[threadIdx.x] = 0u;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        uint32_t f = 0;
#pragma unroll
        for (int j = 0; j < RB; ++j) {
            int8_t v = 0;
            if (j < UP2) {
                const float x = __ldg(w + (size_t)r * UP2 + j);
                v = (int8_t)(int)quant(x);
                f |= (quant_pass(x) ? 1u : 0u) << j;
            }
            q[(size_t)r * RB + j] = v;
        }
        flag[r] = (uint16_t)f;
    }
}

template <int UP>
__device__ __forceinline__ void load_qrow(const int8_t *__restrict__ q, int row, int (&v)[UP * UP])
{
    if constexpr (UP == 1) {
        v[0] = __ldg(q + row);
    } else if constexpr (UP == 2) {
        const int w = __ldg(reinterpret_cast<const int *>(q) + row);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (int)(int8_t)(w >> (8 * j));
    } else if constexpr (UP == 3) {
        const int *p = reinterpret_cast<const int *>(q + (size_t)row * 12);
        const int w[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
#pragma unroll
        for (int j = 0; j < 9; ++j) v[j] = (int)(int8_t)(w[j >> 2] >> (8 * (j & 3)));
    } else {
        const int4 w4 = __ldg(reinterpret_cast<const int4 *>(q) + row);
        const int w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = (int)(int8_t)(w[j >> 2] >> (8 * (j & 3)));
    }
}

// simplex of four tap VALUES (model.py:122-160 without the pad/rotate bookkeeping).  K4 only: its taps have been rounded
// to the integer grid, so the split (floor_divide / python remainder), the sort and the weights are done in integers -
// the same values the float path of K2/K3 computes, at a sixth of the instructions (the float form with its fmodf /
// floorf / rank sort cost 555 instructions per interpolation: the forward kernels were bound by them).
// Order: fractions descending, ties -> higher tap index first (what the reference's 24 strict-inequality cases resolve
// to; it only matters for the input gradient): key = f << 24 | tap << 22 | stride, sorted descending.
__device__ __forceinline__ void simplex_from_taps(const float (&t)[4], int interval, int n_rows, Simplex &s)
{
    const int q = 1 << interval;
    const int L = (1 << (8 - interval)) + 1;
    const uint32_t stride[4] = {(uint32_t)(L * L * L), (uint32_t)(L * L), (uint32_t)L, 1u};
    uint32_t key[4];
    int v0 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int ti = __float2int_rn(t[k]);
        v0 += (ti >> interval) * (int)stride[k];                 // arithmetic shift = floor division (any sign)
        key[k] = ((uint32_t)(ti & (q - 1)) << 24) | ((uint32_t)k << 22) | stride[k];
    }
    sort4_desc(key[0], key[1], key[2], key[3]);
    s.v[0] = v0;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        s.order[p] = (int)((key[p] >> 22) & 3u);
        s.v[p + 1] = s.v[p] + (int)(key[p] & 0x3FFFFFu);
    }
#pragma unroll
    for (int p = 0; p < 5; ++p) s.v[p] = min(max(s.v[p], 0), n_rows - 1);   // memory safety only
    const int f1 = (int)(key[0] >> 24), f2 = (int)(key[1] >> 24), f3 = (int)(key[2] >> 24), f4 = (int)(key[3] >> 24);
    s.w[0] = (float)(q - f1);
    s.w[1] = (float)(f1 - f2);
    s.w[2] = (float)(f2 - f3);
    s.w[3] = (float)(f3 - f4);
    s.w[4] = (float)f4;
}

template <int UP>
__global__ void __launch_bounds__(256)
stage_fwd_kernel(const __grid_constant__ StageF32Args a, float *__restrict__ out, uint8_t *__restrict__ mask)
{
    constexpr int UP2 = UP * UP;
    const size_t total = (size_t)a.BC * a.h * a.w;
    const float inv_q = 1.f / (float)(1 << a.interval);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % a.w);
        const size_t rr = i / a.w;
        const int y = (int)(rr % a.h);
        const size_t bc = rr / a.h;
        const float *__restrict__ plane = a.x + bc * (size_t)a.h * a.w;
        float pred[UP2];
#pragma unroll
        for (int j = 0; j < UP2; ++j) pred[j] = 0.f;
        for (int m = 0; m < a.n_modes; ++m) {
            const int8_t *__restrict__ Q = a.wq[m];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float t[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int yy = clampi(y + a.taps.dy[m][r][k], 0, a.h - 1);
                    const int xx = clampi(x + a.taps.dx[m][r][k], 0, a.w - 1);
                    // K4 works on the integer grid (see mulut.h): an input a rounding error away from k (x/255*255
                    // computed with a reciprocal) is taken as k, as the reference's continuous interpolation would
                    t[k] = rintf(__ldg(plane + (size_t)yy * a.w + xx));
                }
                Simplex s;
                simplex_from_taps(t, a.interval, a.n_rows, s);
                if (m == 0 && r == 0) {
                    // how many lanes share their first LUT row with another lane of the warp: the backward
                    // uses it to choose between direct and warp-aggregated gradient scatter
                    const unsigned am = __activemask();
                    const unsigned peers = __match_any_sync(am, s.v[0]);
                    const unsigned shared_rows = __ballot_sync(am, (peers & (peers - 1u)) != 0u);
                    if ((threadIdx.x & 31) == (unsigned)(__ffs(am) - 1)) {
                        atomicAdd(a.stats, (uint32_t)__popc(shared_rows));
                        atomicAdd(a.stats + 1, (uint32_t)__popc(am));
                    }
                }
                // weights and quantised LUT values are small integers: the sum is exact in int32
                // (and equal to the reference's fp32 sum, |o| <= q * 127)
                int o[UP2];
#pragma unroll
                for (int j = 0; j < UP2; ++j) o[j] = 0;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    int qv[UP2];
                    load_qrow<UP>(Q, s.v[k], qv);
                    const int wk = (int)s.w[k];
#pragma unroll
                    for (int j = 0; j < UP2; ++j) o[j] += wk * qv[j];
                }
                // pred += rot_back(o / q); pred = round(pred)   (model.py:306-308, half-to-even)
#pragma unroll
                for (int j = 0; j < UP2; ++j) {
                    const int p = subpixel_perm<UP>(r, j);
                    pred[p] = rintf(__fadd_rn(pred[p], __fmul_rn((float)o[j], inv_q)));
                }
            }
        }
        // x = round(clamp(pred / avg + bias, 0, 255))   (model.py:309)
#pragma unroll
        for (int u = 0; u < UP; ++u)
#pragma unroll
            for (int v = 0; v < UP; ++v) {
                const float pre = __fadd_rn(__fdiv_rn(pred[u * UP + v], a.avg), a.bias);
                const size_t o_idx = (bc * a.h * UP + (size_t)y * UP + u) * ((size_t)a.w * UP) + (size_t)x * UP + v;
                out[o_idx] = rintf(fminf(fmaxf(pre, 0.f), 255.f));
                mask[o_idx] = (pre >= 0.f && pre <= 255.f) ? 1 : 0;
            }
    }
}

// Warp-aggregated scatter: lanes of a warp that update the SAME LUT row (natural patches sit on
// the LUT diagonal: neighbouring pixels share cell and fraction order) merge their contributions
// through shuffles - a tree over the peers found by match.any - and only the first peer issues the
// red.  When every lane has a row of its own (uniform random patches) the cost is one match + one vote.
// Returns true on the lane that must issue the update; c[] then holds the group's sum.
template <int N>
__device__ __noinline__ bool warp_merge_tree(unsigned active, unsigned peers, float *c)
{
    const int lane = threadIdx.x & 31;
    const int rel = __popc(peers & ((1u << lane) - 1u));        // my rank among the peers
    const int size = __popc(peers);
    for (int s = 1; s < 32; s <<= 1) {
        if (!__any_sync(active, size > s)) break;
        const unsigned src = __fns(peers, 0, rel + s + 1);       // lane of the peer ranked rel + s (or ~0u)
        const bool take = ((rel & (2 * s - 1)) == 0) && (src != 0xffffffffu);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const float v = __shfl_sync(active, c[j], take ? (int)src : lane);
            if (take) c[j] += v;
        }
    }
    return rel == 0;
}

template <int N>
__device__ __forceinline__ bool warp_merge_rows(unsigned active, int row, float (&c)[N])
{
    const unsigned peers = __match_any_sync(active, row);
    if (!__any_sync(active, (peers & (peers - 1u)) != 0u)) return true;     // every lane alone: nothing to merge
    return warp_merge_tree<N>(active, peers, c);                            // out of line: keeps the unrolled passes small
}

constexpr int K4_T = 16;                 // backward tile: 16 x 16 input pixels per CTA
constexpr int K4_P = K4_T + 4;           // + 2-pixel halo on every side

template <int UP, bool AGG>
__device__ __forceinline__ void stage_bwd_body(const StageF32Args &a, const float *__restrict__ gout,
                                               const uint8_t *__restrict__ mask, float *__restrict__ gx, float *s_gx,
                                               int rep)
{
    constexpr int UP2 = UP * UP;
    const int tiles_x = (a.w + K4_T - 1) / K4_T, tiles_y = (a.h + K4_T - 1) / K4_T;
    const int tx = blockIdx.x % tiles_x;
    const int ty = (blockIdx.x / tiles_x) % tiles_y;
    const size_t bc = blockIdx.x / (tiles_x * tiles_y);
    const int lx = threadIdx.x % K4_T, ly = threadIdx.x / K4_T;
    const int x = tx * K4_T + lx, y = ty * K4_T + ly;
    const float inv_q = 1.f / (float)(1 << a.interval);
    for (int i = threadIdx.x; i < K4_P * K4_P; i += blockDim.x) s_gx[i] = 0.f;
    __syncthreads();

    const unsigned active = __ballot_sync(0xffffffffu, x < a.w && y < a.h);   // lanes that own a pixel
    if (x < a.w && y < a.h) {
        const float *__restrict__ plane = a.x + bc * (size_t)a.h * a.w;
        // g = dL/d(o)  = G * mask / avg / q   (o enters pred as o/q; the rounds are identity)
        float g[UP2];
#pragma unroll
        for (int u = 0; u < UP; ++u)
#pragma unroll
            for (int v = 0; v < UP; ++v) {
                const size_t o_idx = (bc * a.h * UP + (size_t)y * UP + u) * ((size_t)a.w * UP) + (size_t)x * UP + v;
                g[u * UP + v] = mask[o_idx] ? __fdiv_rn(__ldg(gout + o_idx), a.avg) * inv_q : 0.f;
            }
        bool any_g = false;
#pragma unroll
        for (int j = 0; j < UP2; ++j) any_g |= g[j] != 0.f;
        for (int m = 0; m < a.n_modes; ++m) {
            const int8_t *__restrict__ Q = a.wq[m];
            const uint16_t *__restrict__ FL = a.wflag[m];
            float *__restrict__ GW = a.gweight[m];
            if (GW && rep) GW = a.replica + ((size_t)(rep - 1) * a.n_modes + m) * a.replica_stride;   // hot rows: this CTA's copy
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float t[4];
                int sidx[4];                                  // tile-local index of each tap (clamped coordinate)
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int yy = clampi(y + a.taps.dy[m][r][k], 0, a.h - 1);
                    const int xx = clampi(x + a.taps.dx[m][r][k], 0, a.w - 1);
                    // K4 works on the integer grid (see mulut.h): an input a rounding error away from k (x/255*255
                    // computed with a reciprocal) is taken as k, as the reference's continuous interpolation would
                    t[k] = rintf(__ldg(plane + (size_t)yy * a.w + xx));
                    sidx[k] = (yy - ty * K4_T + 2) * K4_P + (xx - tx * K4_T + 2);
                }
                Simplex s;
                simplex_from_taps(t, a.interval, a.n_rows, s);
                float gr[UP2];                                // this rotation's view of g: column j <- sub-pixel perm(r, j)
#pragma unroll
                for (int j = 0; j < UP2; ++j) gr[j] = g[subpixel_perm<UP>(r, j)];
                float dot_prev = 0.f;
#pragma unroll
                for (int k = 0; k < 5; ++k) {
                    int qv[UP2];
                    load_qrow<UP>(Q, s.v[k], qv);
                    float dot = 0.f;
#pragma unroll
                    for (int j = 0; j < UP2; ++j) dot += gr[j] * (float)qv[j];
                    if (GW) {
                        const uint32_t fl = __ldg(FL + s.v[k]);
                        float *__restrict__ gw = GW + (size_t)s.v[k] * UP2;
                        // d/dweight = 127 * [clamp passes] * (w_k / q) * g   (straight-through quantiser)
                        float c[UP2];
#pragma unroll
                        for (int j = 0; j < UP2; ++j) c[j] = ((fl >> j) & 1u) ? 127.f * (s.w[k] * gr[j]) : 0.f;
                        // a vertex with weight 0 (tied fractions, f = 0: 23 % of the last vertices on noise) or a
                        // pixel whose output was clamped (g = 0) adds nothing: no atomic for it
                        bool live = s.w[k] != 0.f && any_g;
                        if (AGG) live = warp_merge_rows<UP2>(active, live ? s.v[k] : -1 - (int)(threadIdx.x & 31), c) && live;
                        if (live) {
                            if constexpr (UP2 % 4 == 0) {
#pragma unroll
                                for (int j = 0; j < UP2; j += 4) red_add_v4(gw + j, c[j], c[j + 1], c[j + 2], c[j + 3]);
                            } else {
#pragma unroll
                                for (int j = 0; j < UP2; ++j)
                                    if (c[j] != 0.f) atomicAdd(gw + j, c[j]);
                            }
                        }
                    }
                    if (gx && k > 0) {
                        int si = 0;
#pragma unroll
                        for (int tt = 0; tt < 4; ++tt)
                            if (s.order[k - 1] == tt) si = sidx[tt];
                        atomicAdd(&s_gx[si], dot - dot_prev);
                    }
                    dot_prev = dot;
                }
            }
        }
    }
    if (gx) {
        __syncthreads();
        float *__restrict__ gplane = gx + bc * (size_t)a.h * a.w;
        for (int i = threadIdx.x; i < K4_P * K4_P; i += blockDim.x) {
            const int yy = ty * K4_T - 2 + i / K4_P, xx = tx * K4_T - 2 + i % K4_P;
            const float v = s_gx[i];
            if (v != 0.f && yy >= 0 && yy < a.h && xx >= 0 && xx < a.w) atomicAdd(gplane + (size_t)yy * a.w + xx, v);
        }
    }
}

template <int UP>
__global__ void __launch_bounds__(K4_T * K4_T)
stage_bwd_kernel(const __grid_constant__ StageF32Args a, const float *__restrict__ gout,
                 const uint8_t *__restrict__ mask, float *__restrict__ gx)
{
    __shared__ float s_gx[K4_P * K4_P];
    // aggregate when more than a quarter of the forward's lanes shared a LUT row with a neighbour
    const bool hot = a.aggregate == 2 ? (4ull * a.stats[0] > (unsigned long long)a.stats[1]) : a.aggregate != 0;
    const int rep = (hot && a.n_replicas > 1) ? (int)(blockIdx.x % a.n_replicas) : 0;
    if (hot && a.merge) stage_bwd_body<UP, true>(a, gout, mask, gx, s_gx, rep);
    else stage_bwd_body<UP, false>(a, gout, mask, gx, s_gx, rep);
}

// ---------------------------------------------------------------------------
// K4p: the x1 stage's LUT gradients accumulated in SHARED MEMORY ("private" backward).
// stage_bwd_kernel<1> issues 60 scalar red.global.add per pixel and runs at the L2's atomic rate
// (131 G ops/s, lts 75 %).  A x1 table at interval 4 is 83 521 floats; the rows one pixel touches lie in the
// slabs a in {m_a, m_a + 1} of its centre tap, so the pixels with a centre value < 128 only touch slabs 0..8 and
// the others slabs 8..16: 9 x 4913 floats = 177 KB, one SM's shared memory.  A persistent CTA serves ONE
// (mode, half): its warps scan their share of the pixels, queue the ones of their half whose output gradient is
// live (per-warp ring in shared memory, so the 32 lanes always work on 32 queued pixels), add the 4 x 5 vertex
// contributions with shared-memory atomics (a CAS loop for fp32; lanes on one row simply retry - the merge tree of
// the global kernel costs more than it saves here) and the CTA flushes its slab range once with 16-byte vector reds:
// 6.5 M floats per launch instead of 35 M scalar reds.  Used when the stage needs no input gradient (stage 1)
// and there are enough pixels to pay for the flush (see use_private_bwd).  stage_bwd_kernel<1> 269 -> 117 us at batch
// 256 on noise patches, 450 -> 278 us on smooth ones.
// ---------------------------------------------------------------------------
constexpr int K4P_SLAB = 17 * 17 * 17;           // rows per value of tap a at interval 4
constexpr int K4P_ROWS = 9 * K4P_SLAB;           // one half: slabs 0..8 or 8..16
constexpr int K4P_THREADS = 768;
constexpr int K4P_WARPS = K4P_THREADS / 32;
constexpr int K4P_RING = 64;                     // per-warp queue of pixel indices (< 32 left + 32 appended)
constexpr int K4P_FLAG_WORDS = (K4P_ROWS + 31) / 32;   // the half's clamp flags as a bitmap (5.5 KB): 20 random 2-byte
                                                        // global gathers per pixel otherwise - they were half the kernel
constexpr size_t K4P_TAB_BYTES = (size_t)(K4P_ROWS + 3) / 4 * 16;
constexpr size_t K4P_RING_BYTES = (size_t)K4P_WARPS * K4P_RING * 4;
constexpr size_t K4P_SMEM = K4P_TAB_BYTES + K4P_RING_BYTES + (size_t)K4P_FLAG_WORDS * 4;

template <bool AGG>
__device__ __forceinline__ void private_bwd_pixel(const StageF32Args &a, int m, int half_base, size_t p, float g,
                                                  unsigned active, float *__restrict__ s_tab,
                                                  const uint32_t *__restrict__ s_flag)
{
    const int x = (int)(p % a.w);
    const size_t rr = p / a.w;
    const int y = (int)(rr % a.h);
    const float *__restrict__ plane = a.x + (rr / a.h) * (size_t)a.h * a.w;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        float t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int yy = clampi(y + a.taps.dy[m][r][k], 0, a.h - 1);
            const int xx = clampi(x + a.taps.dx[m][r][k], 0, a.w - 1);
            t[k] = rintf(__ldg(plane + (size_t)yy * a.w + xx));
        }
        Simplex s;
        simplex_from_taps(t, 4, a.n_rows, s);
#pragma unroll
        for (int k = 0; k < 5; ++k) {
            const int rel = min(max(s.v[k] - half_base, 0), K4P_ROWS - 1);      // in range for inputs 0..255
            float c[1];
            c[0] = ((s_flag[rel >> 5] >> (rel & 31)) & 1u) ? 127.f * (s.w[k] * g) : 0.f;   // same product as stage_bwd_body
            bool live = s.w[k] != 0.f;
            if (AGG) live = warp_merge_rows<1>(active, live ? rel : -1 - (int)(threadIdx.x & 31), c) && live;
            if (live && c[0] != 0.f) atomicAdd(s_tab + rel, c[0]);
        }
    }
}

template <bool AGG>
__device__ __forceinline__ void private_bwd_body(const StageF32Args &a, const float *__restrict__ gout,
                                                 const uint8_t *__restrict__ mask, float *s_tab, uint32_t *s_ring,
                                                 uint32_t *s_flag)
{
    const int kinds = 2 * a.n_modes;
    const int kind = blockIdx.x % kinds, slot = blockIdx.x / kinds;
    const int nslots = ((int)gridDim.x - kind + kinds - 1) / kinds;
    const int m = kind >> 1, half = kind & 1;
    const int half_base = half ? 8 * K4P_SLAB : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t total = (size_t)a.BC * a.h * a.w;
    const float inv_q = 1.f / 16.f;
    for (int i = threadIdx.x; i < (K4P_ROWS + 3) / 4 * 4; i += blockDim.x) s_tab[i] = 0.f;
    {
        const uint16_t *__restrict__ FL = a.wflag[m] + half_base;
        constexpr int U = 8;                                                 // bitmap words per warp pass: 8 loads in flight
        for (int w0 = warp * U; w0 < K4P_FLAG_WORDS; w0 += K4P_WARPS * U) {
            uint32_t f[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int row = (w0 + u) * 32 + lane;
                f[u] = row < K4P_ROWS ? (uint32_t)__ldg(FL + row) : 0u;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned bits = __ballot_sync(0xffffffffu, (f[u] & 1u) != 0u);
                if (lane == 0 && w0 + u < K4P_FLAG_WORDS) s_flag[w0 + u] = bits;
            }
        }
    }
    __syncthreads();

    uint32_t *ring = s_ring + warp * K4P_RING;
    int head = 0, qn = 0;                                      // warp-uniform
    const size_t stride = (size_t)nslots * K4P_WARPS * 32;
    for (size_t base = ((size_t)slot * K4P_WARPS + warp) * 32; base < total; base += stride) {
        const size_t p = base + lane;
        bool mine = false;
        if (p < total && mask[p]) {
            const int ta = __float2int_rn(__ldg(a.x + p));
            mine = ((ta >= 128) == (half != 0)) && __ldg(gout + p) != 0.f;
        }
        const unsigned b = __ballot_sync(0xffffffffu, mine);
        // (pixel indices as 32 bits: the launcher only takes this path below 2^32 pixels)
        if (mine) ring[(head + qn + __popc(b & ((1u << lane) - 1u))) & (K4P_RING - 1)] = (uint32_t)p;
        qn += __popc(b);
        __syncwarp();
        if (qn >= 32) {
            const size_t pp = ring[(head + lane) & (K4P_RING - 1)];
            const float g = __fdiv_rn(__ldg(gout + pp), a.avg) * inv_q;
            private_bwd_pixel<AGG>(a, m, half_base, pp, g, 0xffffffffu, s_tab, s_flag);
            head = (head + 32) & (K4P_RING - 1);
            qn -= 32;
            __syncwarp();
        }
    }
    {
        const unsigned active = __ballot_sync(0xffffffffu, lane < qn);
        if (lane < qn) {
            const size_t pp = ring[(head + lane) & (K4P_RING - 1)];
            const float g = __fdiv_rn(__ldg(gout + pp), a.avg) * inv_q;
            private_bwd_pixel<AGG>(a, m, half_base, pp, g, active, s_tab, s_flag);
        }
    }
    __syncthreads();
    float *__restrict__ gw = a.gweight[m] + half_base;
    const float4 *s4 = reinterpret_cast<const float4 *>(s_tab);
    for (int i = threadIdx.x; i < K4P_ROWS / 4; i += blockDim.x) {
        const float4 v = s4[i];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) red_add_v4(gw + 4 * i, v.x, v.y, v.z, v.w);
    }
    if (threadIdx.x < K4P_ROWS % 4) {
        const int i = K4P_ROWS / 4 * 4 + threadIdx.x;
        if (s_tab[i] != 0.f) atomicAdd(gw + i, s_tab[i]);
    }
}

__global__ void __launch_bounds__(K4P_THREADS, 1)
stage1_bwd_private_kernel(const __grid_constant__ StageF32Args a, const float *__restrict__ gout,
                          const uint8_t *__restrict__ mask)
{
    extern __shared__ __align__(16) unsigned char k4p_smem[];
    float *s_tab = reinterpret_cast<float *>(k4p_smem);
    uint32_t *s_ring = reinterpret_cast<uint32_t *>(k4p_smem + K4P_TAB_BYTES);
    uint32_t *s_flag = reinterpret_cast<uint32_t *>(k4p_smem + K4P_TAB_BYTES + K4P_RING_BYTES);
    // no warp merge here: lanes on one row retry their shared-memory CAS a few times, which costs less than the
    // match + shuffle tree (smooth patches: 278 us without, 477 us with; noise: 147 us)
    private_bwd_body<false>(a, gout, mask, s_tab, s_ring, s_flag);
}

__device__ __forceinline__ bool stage_rows_are_hot(const StageF32Args &a)
{
    return a.aggregate == 2 ? (4ull * a.stats[0] > (unsigned long long)a.stats[1]) : a.aggregate != 0;
}

// Both run on the stage's stream around stage_bwd_kernel and return at once when the replicas are not in use.
__global__ void __launch_bounds__(256) replica_zero_kernel(const __grid_constant__ StageF32Args a)
{
    if (!stage_rows_are_hot(a)) return;
    const size_t n4 = (size_t)(a.n_replicas - 1) * a.n_modes * a.replica_stride / 4;
    float4 *__restrict__ r4 = reinterpret_cast<float4 *>(a.replica);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        r4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(256) replica_sum_kernel(const __grid_constant__ StageF32Args a, int up2)
{
    if (!stage_rows_are_hot(a)) return;
    const size_t n = (size_t)a.n_rows * up2;
    for (int m = 0; m < a.n_modes; ++m) {
        float *__restrict__ gw = a.gweight[m];
        if (!gw) continue;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            float acc = 0.f;
            for (int r = 1; r < a.n_replicas; ++r) acc += a.replica[((size_t)(r - 1) * a.n_modes + m) * a.replica_stride + i];
            if (acc != 0.f) gw[i] += acc;
        }
    }
}

// MULUT_K4_PRIVATE: 0 never, 1 whenever the stage qualifies, default: when it has >= 2^16 pixels.  Measured on 48 x 48
// patches, whole step: batch 256 1.42 -> 1.29 ms, 128 0.78 -> 0.72, 64 0.51 -> 0.48, 32 (73 728 pixels) 0.27 -> 0.26;
// the fixed cost (zeroing and flushing 148 x 177 KB) is not measured below that, so smaller launches keep the reds.
static bool use_private_bwd(size_t total)
{
    const char *e = getenv("MULUT_K4_PRIVATE");
    if (e) return atoi(e) != 0;
    return total >= ((size_t)1 << 16);
}

static int fill_stage_args(StageF32Args &a, const float *const *weights, int n_modes, const char *modes, int n_rows,
                           int up, int interval, const float *x, int B, int C, int h, int w, float avg, float bias,
                           void *workspace)
{
    if (!workspace) { set_error("stage: null workspace (mulut_stage_workspace_bytes)"); return MULUT_E_BAD_ARG; }
    if (reinterpret_cast<uintptr_t>(workspace) & 15) { set_error("stage: workspace must be 16-byte aligned"); return MULUT_E_BAD_ARG; }
    if (!weights || !modes || !x || n_modes < 1 || n_modes > MULUT_MAX_MODES || B < 0 || C < 0 || h < 0 || w < 0 ||
        interval < 1 || interval > 7 || !(avg > 0.f)) {
        set_error("stage: bad argument");
        return MULUT_E_BAD_ARG;
    }
    if (up < 1 || up > 4) { set_error("stage: upscale %d not supported (1..4)", up); return MULUT_E_BAD_ARG; }
    memset(&a, 0, sizeof a);
    if (!build_tap_table(modes, n_modes, &a.taps)) {
        for (int m = 0; m < n_modes; ++m) {
            int dy[4], dx[4];
            if (!mode_taps(modes[m], dy, dx)) { set_error("Mode %c not implemented.", modes[m]); break; }
        }
        return MULUT_E_BAD_MODE;
    }
    const int L = (1 << (8 - interval)) + 1;
    if ((long long)n_rows < (long long)L * L * L * L) {
        set_error("LUT too small: need %lld rows, have %d", (long long)L * L * L * L, n_rows);
        return MULUT_E_LUT_SMALL;
    }
    for (int m = 0; m < n_modes; ++m) {
        if (!weights[m]) { set_error("stage: weights[%d] is null", m); return MULUT_E_BAD_ARG; }
        a.weight[m] = weights[m];
        uint8_t *base = static_cast<uint8_t *>(workspace) + (size_t)m * stage_ws_mode_bytes(n_rows, up);
        a.wq[m] = reinterpret_cast<const int8_t *>(base);
        a.wflag[m] = reinterpret_cast<const uint16_t *>(base + align16((size_t)n_rows * q_pitch(up)));
    }
    a.stats = reinterpret_cast<uint32_t *>(static_cast<uint8_t *>(workspace) + (size_t)n_modes * stage_ws_mode_bytes(n_rows, up));
    a.replica = reinterpret_cast<float *>(static_cast<uint8_t *>(workspace) + (size_t)n_modes * stage_ws_mode_bytes(n_rows, up) +
                                          STAGE_WS_TAIL);
    a.replica_stride = replica_table_floats(n_rows, up);
    a.n_replicas = 1;
    a.x = x; a.BC = B * C; a.h = h; a.w = w; a.n_modes = n_modes; a.interval = interval; a.n_rows = n_rows;
    a.avg = avg; a.bias = bias;
    return MULUT_OK;
}

static int fill_args(F32Args &a, const float *weight, int n_rows, int up, char mode, const float *img, int BC,
                     int h, int w, int bd, int interval)
{
    if (!weight || !img || n_rows < 1 || BC < 0 || h < 0 || w < 0 || interval < 1 || interval > 7) {
        set_error("interp: bad argument");
        return MULUT_E_BAD_ARG;
    }
    if (up < 1 || up > 4) { set_error("interp: upscale %d not supported (1..4)", up); return MULUT_E_BAD_ARG; }
    if (!mode_taps(mode, a.dy, a.dx)) { set_error("Mode %c not implemented.", mode); return MULUT_E_BAD_MODE; }
    if (bd < mode_pad(mode)) { set_error("interp: bd=%d smaller than the mode's padding", bd); return MULUT_E_BAD_ARG; }
    const int L = (1 << (8 - interval)) + 1;
    if ((long long)n_rows < (long long)L * L * L * L) {
        set_error("LUT too small: need %lld rows, have %d", (long long)L * L * L * L, n_rows);
        return MULUT_E_LUT_SMALL;
    }
    a.weight = weight; a.img = img; a.n_rows = n_rows; a.up = up; a.BC = BC; a.h = h; a.w = w; a.bd = bd;
    a.interval = interval;
    return MULUT_OK;
}

static int sm_count()
{
    // multiProcessorCount of the current device (B200: 148), cached per device
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cached[dev]) {
        int n = 0;
        cached[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) ? n : 148;
    }
    return cached[dev];
}

static unsigned grid_for(size_t total)
{
    size_t b = (total + 255) / 256;
    const size_t cap = (size_t)sm_count() * 32;
    if (b > cap) b = cap;
    return (unsigned)(b ? b : 1);
}

}  // namespace mulut

using namespace mulut;

extern "C" int mulut_interp_fwd_f32(const float *d_weight, int n_rows, int up, char mode, const float *d_img_in,
                                    int B, int C, int h, int w, int bd, int interval, float *d_out, void *stream)
{
    F32Args a;
    int rc = fill_args(a, d_weight, n_rows, up, mode, d_img_in, B * C, h, w, bd, interval);
    if (rc) return rc;
    if (!d_out) { set_error("interp_fwd: null output"); return MULUT_E_BAD_ARG; }
    const size_t total = (size_t)B * C * h * w;
    if (total == 0) return MULUT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (up) {
    case 1: interp_fwd_kernel<1><<<grid_for(total), 256, 0, st>>>(a, d_out); break;
    case 2: interp_fwd_kernel<2><<<grid_for(total), 256, 0, st>>>(a, d_out); break;
    case 3: interp_fwd_kernel<3><<<grid_for(total), 256, 0, st>>>(a, d_out); break;
    default: interp_fwd_kernel<4><<<grid_for(total), 256, 0, st>>>(a, d_out); break;
    }
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

extern "C" int mulut_interp_bwd_f32(const float *d_weight, int n_rows, int up, char mode, const float *d_img_in,
                                    int B, int C, int h, int w, int bd, int interval, const float *d_grad_out,
                                    float *d_grad_weight, float *d_grad_img_in, void *stream)
{
    F32Args a;
    int rc = fill_args(a, d_weight, n_rows, up, mode, d_img_in, B * C, h, w, bd, interval);
    if (rc) return rc;
    if (!d_grad_out) { set_error("interp_bwd: null grad_out"); return MULUT_E_BAD_ARG; }
    const size_t total = (size_t)B * C * h * w;
    if (total == 0 || (!d_grad_weight && !d_grad_img_in)) return MULUT_OK;
    if ((up == 2 || up == 4) && (reinterpret_cast<uintptr_t>(d_grad_weight) & 15)) {
        set_error("interp_bwd: grad_weight must be 16-byte aligned (vector red.global.add)");
        return MULUT_E_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (up) {
    case 1: interp_bwd_kernel<1><<<grid_for(total), 256, 0, st>>>(a, d_grad_out, d_grad_weight, d_grad_img_in); break;
    case 2: interp_bwd_kernel<2><<<grid_for(total), 256, 0, st>>>(a, d_grad_out, d_grad_weight, d_grad_img_in); break;
    case 3: interp_bwd_kernel<3><<<grid_for(total), 256, 0, st>>>(a, d_grad_out, d_grad_weight, d_grad_img_in); break;
    default: interp_bwd_kernel<4><<<grid_for(total), 256, 0, st>>>(a, d_grad_out, d_grad_weight, d_grad_img_in); break;
    }
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

extern "C" int mulut_interp_pass_f64(const float *d_weight, int n_rows, const float *d_img_in, int C, int h, int w,
                                     int interval, int rot, int upscale, char mode, double *d_out, void *stream)
{
    F32Args a;
    // img_in is (C, h+p, w+p) with p = the mode's padding (sr/4_test_lut.py:289-296)
    int dy[4], dx[4];
    if (!mode_taps(mode, dy, dx)) { set_error("Mode %c not implemented.", mode); return MULUT_E_BAD_MODE; }
    int rc = fill_args(a, d_weight, n_rows, upscale, mode, d_img_in, C, h, w, mode_pad(mode), interval);
    if (rc) return rc;
    if (!d_out) { set_error("interp_pass: null output"); return MULUT_E_BAD_ARG; }
    const size_t total = (size_t)C * h * w;
    if (total == 0) return MULUT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    switch (upscale) {
    case 1: interp_pass_f64_kernel<1><<<grid_for(total), 256, 0, st>>>(a, rot, d_out); break;
    case 2: interp_pass_f64_kernel<2><<<grid_for(total), 256, 0, st>>>(a, rot, d_out); break;
    case 3: interp_pass_f64_kernel<3><<<grid_for(total), 256, 0, st>>>(a, rot, d_out); break;
    default: interp_pass_f64_kernel<4><<<grid_for(total), 256, 0, st>>>(a, rot, d_out); break;
    }
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

extern "C" size_t mulut_stage_workspace_bytes(int n_modes, int n_rows, int up)
{
    if (n_modes < 1 || n_rows < 1 || up < 1 || up > 4) return 0;
    return (size_t)n_modes * stage_ws_mode_bytes(n_rows, up) + STAGE_WS_TAIL + stage_ws_replica_bytes(n_modes, n_rows, up);
}

extern "C" int mulut_stage_fwd_f32(const float *const *d_weights, int n_modes, const char *modes, int n_rows, int up,
                                   int interval, const float *d_x, int B, int C, int h, int w, float avg, float bias,
                                   float *d_out, uint8_t *d_mask, void *d_workspace, void *stream)
{
    StageF32Args a;
    int rc = fill_stage_args(a, d_weights, n_modes, modes, n_rows, up, interval, d_x, B, C, h, w, avg, bias, d_workspace);
    if (rc) return rc;
    if (!d_out || !d_mask) { set_error("stage_fwd: null output"); return MULUT_E_BAD_ARG; }
    const size_t total = (size_t)B * C * h * w;
    cudaStream_t st = (cudaStream_t)stream;
    for (int m = 0; m < n_modes; ++m) {          // quantise the tables once for this stage (read again by the backward)
        int8_t *q = const_cast<int8_t *>(a.wq[m]);
        uint16_t *f = const_cast<uint16_t *>(a.wflag[m]);
        const unsigned qb = (unsigned)((n_rows + 255) / 256);
        switch (up) {
        case 1: quantize_rows_kernel<1><<<qb, 256, 0, st>>>(a.weight[m], n_rows, q, f, m == 0 ? a.stats : nullptr); break;
        case 2: quantize_rows_kernel<2><<<qb, 256, 0, st>>>(a.weight[m], n_rows, q, f, m == 0 ? a.stats : nullptr); break;
        case 3: quantize_rows_kernel<3><<<qb, 256, 0, st>>>(a.weight[m], n_rows, q, f, m == 0 ? a.stats : nullptr); break;
        default: quantize_rows_kernel<4><<<qb, 256, 0, st>>>(a.weight[m], n_rows, q, f, m == 0 ? a.stats : nullptr); break;
        }
    }
    MULUT_CUDA(cudaGetLastError());
    if (total == 0) return MULUT_OK;
    switch (up) {
    case 1: stage_fwd_kernel<1><<<grid_for(total), 256, 0, st>>>(a, d_out, d_mask); break;
    case 2: stage_fwd_kernel<2><<<grid_for(total), 256, 0, st>>>(a, d_out, d_mask); break;
    case 3: stage_fwd_kernel<3><<<grid_for(total), 256, 0, st>>>(a, d_out, d_mask); break;
    default: stage_fwd_kernel<4><<<grid_for(total), 256, 0, st>>>(a, d_out, d_mask); break;
    }
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

extern "C" int mulut_stage_bwd_f32(const float *const *d_weights, int n_modes, const char *modes, int n_rows, int up,
                                   int interval, const float *d_x, int B, int C, int h, int w, float avg, float bias,
                                   const float *d_grad_out, const uint8_t *d_mask, const void *d_workspace,
                                   float *const *d_grad_weights, float *d_grad_x, void *stream)
{
    StageF32Args a;
    int rc = fill_stage_args(a, d_weights, n_modes, modes, n_rows, up, interval, d_x, B, C, h, w, avg, bias,
                             const_cast<void *>(d_workspace));
    if (rc) return rc;
    if (!d_grad_out || !d_mask) { set_error("stage_bwd: null grad_out / mask"); return MULUT_E_BAD_ARG; }
    bool any = d_grad_x != nullptr;
    { const char *e = getenv("MULUT_K4_AGGREGATE"); a.aggregate = e ? atoi(e) : 2; }   // 0 / 1 force, default: decide from the forward's statistic
    for (int m = 0; m < n_modes; ++m) {
        a.gweight[m] = d_grad_weights ? d_grad_weights[m] : nullptr;
        any |= a.gweight[m] != nullptr;
        if ((up == 2 || up == 4) && (reinterpret_cast<uintptr_t>(a.gweight[m]) & 15)) {
            set_error("stage_bwd: grad_weights must be 16-byte aligned (vector red.global.add)");
            return MULUT_E_BAD_ARG;
        }
    }
    const size_t total = (size_t)B * C * h * w;
    if (total == 0 || !any) return MULUT_OK;
    const size_t blocks = (size_t)B * C * ((h + K4_T - 1) / K4_T) * ((w + K4_T - 1) / K4_T);
    if (blocks > 0x7fffffffull) { set_error("stage_bwd: too many tiles"); return MULUT_E_BAD_ARG; }
    cudaStream_t st = (cudaStream_t)stream;
    if (up == 1 && !d_grad_x && interval == 4 && total < 0xffffffffull && use_private_bwd(total)) {
        bool ok = sm_count() >= 2 * n_modes;
        for (int m = 0; m < n_modes; ++m) ok &= a.gweight[m] != nullptr && !(reinterpret_cast<uintptr_t>(a.gweight[m]) & 15);
        if (ok) {
            static bool attr_set[64] = {false};
            int dev = 0;
            MULUT_CUDA(cudaGetDevice(&dev));
            if (dev >= 0 && dev < 64 && !attr_set[dev]) {
                MULUT_CUDA(cudaFuncSetAttribute(stage1_bwd_private_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                (int)K4P_SMEM));
                attr_set[dev] = true;
            }
            stage1_bwd_private_kernel<<<(unsigned)sm_count(), K4P_THREADS, K4P_SMEM, st>>>(a, d_grad_out, d_mask);
            MULUT_CUDA(cudaGetLastError());
            return MULUT_OK;
        }
    }
    bool any_gw = false;
    for (int m = 0; m < n_modes; ++m) any_gw |= a.gweight[m] != nullptr;
    { const char *e = getenv("MULUT_K4_REPLICAS"); const int r = e ? atoi(e) : K4_REPLICAS; a.n_replicas = r < 1 ? 1 : r > K4_REPLICAS ? K4_REPLICAS : r; }
    if (!any_gw || a.aggregate == 0) a.n_replicas = 1;
    { const char *e = getenv("MULUT_K4_MERGE"); a.merge = e ? atoi(e) != 0 : 1; }
    if (a.n_replicas > 1) {
        replica_zero_kernel<<<(unsigned)sm_count() * 4, 256, 0, st>>>(a);
        MULUT_CUDA(cudaGetLastError());
    }
    switch (up) {
    case 1: stage_bwd_kernel<1><<<(unsigned)blocks, K4_T * K4_T, 0, st>>>(a, d_grad_out, d_mask, d_grad_x); break;
    case 2: stage_bwd_kernel<2><<<(unsigned)blocks, K4_T * K4_T, 0, st>>>(a, d_grad_out, d_mask, d_grad_x); break;
    case 3: stage_bwd_kernel<3><<<(unsigned)blocks, K4_T * K4_T, 0, st>>>(a, d_grad_out, d_mask, d_grad_x); break;
    default: stage_bwd_kernel<4><<<(unsigned)blocks, K4_T * K4_T, 0, st>>>(a, d_grad_out, d_mask, d_grad_x); break;
    }
    MULUT_CUDA(cudaGetLastError());
    if (a.n_replicas > 1) {
        replica_sum_kernel<<<(unsigned)sm_count() * 4, 256, 0, st>>>(a, up * up);
        MULUT_CUDA(cudaGetLastError());
    }
    return MULUT_OK;
}

// ---------------------------------------------------------------------------
// Fused Adam over the flat LUT parameter buffer (sr/3_finetune_lut.py:85-87,134: torch.optim.Adam
// with betas (0.9, 0.999), eps 1e-8, optional L2 weight decay, no amsgrad), one pass instead of the
// ~10 foreach kernels of the stock optimizer.  The learning rate and the step counter live on the
// device so the call can be captured in a CUDA graph and replayed with a new rate.
//   g' = g + wd * p;  m = b1 m + (1-b1) g';  v = b2 v + (1-b2) g'^2
//   p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// ---------------------------------------------------------------------------
namespace mulut {
__global__ void __launch_bounds__(256)
adam_step_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m, float *__restrict__ v,
                 size_t n, const float *__restrict__ d_lr, float b1, float b2, float eps, float wd,
                 const float *__restrict__ d_step /* already incremented */)
{
    const float t = *d_step, lr = *d_lr;
    const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
    const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = g[i] + wd * pi;
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        p[i] = pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
    }
}
// ---------------------------------------------------------------------------
// Loss head of the training step (sr/model.py:312 `x / 255.0` + sr/3_finetune_lut.py:132 F.mse_loss) in one pass per
// direction: ATen runs six elementwise / reduction kernels over the 9.4 M outputs of a cfg-4 step for it (77 us).
//   forward : loss = mean((x * scale - label)^2)                      (scale = 1/255)
//   backward: grad_x = grad_loss * 2 / n * scale * (x * scale - label)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
mse_head_fwd_kernel(const float *__restrict__ x, const float *__restrict__ label, size_t n, float scale,
                    double *__restrict__ acc /* [0] sum, [1] ticket as bits */, float *__restrict__ loss)
{
    double part = 0.0;
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(x) + i), b = __ldg(reinterpret_cast<const float4 *>(label) + i);
        const float d0 = a.x * scale - b.x, d1 = a.y * scale - b.y, d2 = a.z * scale - b.z, d3 = a.w * scale - b.w;
        part += (double)(d0 * d0 + d1 * d1) + (double)(d2 * d2 + d3 * d3);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = n4 * 4; i < n; ++i) { const float d = x[i] * scale - label[i]; part += (double)(d * d); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    __shared__ double s_part[8];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += s_part[w];
        atomicAdd(acc, t);
        __threadfence();
        unsigned long long *ticket = reinterpret_cast<unsigned long long *>(acc + 1);
        if (atomicAdd(ticket, 1ull) == gridDim.x - 1) {             // last block: the mean
            __threadfence();
            *loss = (float)(*reinterpret_cast<volatile double *>(acc) / (double)n);
        }
    }
}

__global__ void __launch_bounds__(256)
mse_head_bwd_kernel(const float *__restrict__ x, const float *__restrict__ label, size_t n, float scale,
                    const float *__restrict__ grad_loss, float *__restrict__ grad_x)
{
    const float k = __ldg(grad_loss) * (2.0f / (float)n) * scale;
    const size_t n4 = n / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(x) + i), b = __ldg(reinterpret_cast<const float4 *>(label) + i);
        reinterpret_cast<float4 *>(grad_x)[i] = make_float4(k * (a.x * scale - b.x), k * (a.y * scale - b.y),
                                                            k * (a.z * scale - b.z), k * (a.w * scale - b.w));
    }
    if (blockIdx.x == 0 && threadIdx.x == 0)
        for (size_t i = n4 * 4; i < n; ++i) grad_x[i] = k * (x[i] * scale - label[i]);
}

__global__ void adam_tick_kernel(float *d_step) { *d_step += 1.f; }
}  // namespace mulut

extern "C" int mulut_mse_head_fwd_f32(const float *d_x, const float *d_label, size_t n, float scale, void *d_work16,
                                      float *d_loss, void *stream)
{
    if (!d_x || !d_label || !d_work16 || !d_loss || n == 0) { set_error("mse_head_fwd: bad argument"); return MULUT_E_BAD_ARG; }
    if ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_label)) & 15) {
        set_error("mse_head_fwd: x and label must be 16-byte aligned");
        return MULUT_E_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    MULUT_CUDA(cudaMemsetAsync(d_work16, 0, 16, st));
    size_t blocks = (n / 4 + 255) / 256 + 1;
    if (blocks > (size_t)sm_count() * 8) blocks = (size_t)sm_count() * 8;
    mse_head_fwd_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_x, d_label, n, scale, static_cast<double *>(d_work16), d_loss);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

extern "C" int mulut_mse_head_bwd_f32(const float *d_x, const float *d_label, size_t n, float scale,
                                      const float *d_grad_loss, float *d_grad_x, void *stream)
{
    if (!d_x || !d_label || !d_grad_loss || !d_grad_x || n == 0) { set_error("mse_head_bwd: bad argument"); return MULUT_E_BAD_ARG; }
    if ((reinterpret_cast<uintptr_t>(d_x) | reinterpret_cast<uintptr_t>(d_label) | reinterpret_cast<uintptr_t>(d_grad_x)) & 15) {
        set_error("mse_head_bwd: buffers must be 16-byte aligned");
        return MULUT_E_BAD_ARG;
    }
    size_t blocks = (n / 4 + 255) / 256 + 1;
    if (blocks > (size_t)sm_count() * 8) blocks = (size_t)sm_count() * 8;
    mse_head_bwd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_x, d_label, n, scale, d_grad_loss, d_grad_x);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}

extern "C" int mulut_adam_step_f32(float *d_param, const float *d_grad, float *d_exp_avg, float *d_exp_avg_sq, size_t n,
                                   const float *d_lr, float beta1, float beta2, float eps, float weight_decay,
                                   float *d_step, void *stream)
{
    if (!d_param || !d_grad || !d_exp_avg || !d_exp_avg_sq || !d_lr || !d_step) {
        set_error("adam: null argument");
        return MULUT_E_BAD_ARG;
    }
    if (n == 0) return MULUT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 1, 0, st>>>(d_step);
    size_t blocks = (n + 255) / 256;
    if (blocks > (size_t)sm_count() * 16) blocks = (size_t)sm_count() * 16;
    adam_step_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_param, d_grad, d_exp_avg, d_exp_avg_sq, n, d_lr, beta1, beta2, eps,
                                                       weight_decay, d_step);
    MULUT_CUDA(cudaGetLastError());
    return MULUT_OK;
}
