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

static unsigned grid_for(size_t total)
{
    size_t b = (total + 255) / 256;
    if (b > 148 * 32) b = 148 * 32;
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
