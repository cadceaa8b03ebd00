// On-device evaluation of the reference's report metrics (common/utils.py:42-101):
//   Y  = BT.601 luma of an RGB uint8 frame          (_rgb2ycbcr, :42-60, float64)
//   PSNR on Y with a border shave                    (PSNR, :63-72, float32 differences)
//   SSIM on Y, 11x11 Gaussian window (sigma 1.5), 'valid' convolution, float64   (cal_ssim, :75-101)
// so the CLI can print "AVG LUT PSNR/SSIM" without copying both frames to the host and
// running scipy.  One CTA per 16x16 tile of the SSIM map: the 26x26 luma patches of both
// images go to shared memory, a separable 11-tap pass produces the five windowed moments,
// block reductions feed double-precision atomics.
#include "common.cuh"

namespace mulut {

constexpr int EV_T = 16, EV_K = 11, EV_P = EV_T + EV_K - 1;     // tile, window, patch edge

__device__ __forceinline__ double luma(const uint8_t *__restrict__ p)
{
    // row 0 of T in _rgb2ycbcr (65.481, 128.553, 24.966) / 255, offset 16
    return 0.256788235294118 * (double)p[0] + 0.504129411764706 * (double)p[1] + 0.097905882352941 * (double)p[2] + 16.0;
}

struct EvalAcc {
    double ssim_sum;         // sum of the SSIM map
    double sq_sum;           // sum of float32(diff)^2 over the shaved area
    unsigned long long ssim_n, sq_n;
};

__global__ void __launch_bounds__(EV_T * EV_T)
eval_psnr_ssim_kernel(const uint8_t *__restrict__ a, const uint8_t *__restrict__ b, int H, int W, int shave,
                      EvalAcc *__restrict__ acc)
{
    __shared__ double s_a[EV_P][EV_P + 1], s_b[EV_P][EV_P + 1];
    __shared__ double s_h[5][EV_P][EV_T + 1];                    // horizontal pass: a, b, aa, bb, ab
    __shared__ double s_g[EV_K];
    __shared__ double s_red[2][EV_T * EV_T / 32];
    const int tid = threadIdx.x, lx = tid % EV_T, ly = tid / EV_T;
    const int mapW = W - EV_K + 1, mapH = H - EV_K + 1;          // 'valid' output size (may be <= 0)
    const int x0 = blockIdx.x * EV_T, y0 = blockIdx.y * EV_T;
    if (tid < EV_K) {                                            // cv2.getGaussianKernel(11, 1.5)
        double sum = 0.0;
        for (int i = 0; i < EV_K; ++i) { const double d = i - (EV_K - 1) / 2.0; sum += exp(-(d * d) / (2.0 * 1.5 * 1.5)); }
        const double d = tid - (EV_K - 1) / 2.0;
        s_g[tid] = exp(-(d * d) / (2.0 * 1.5 * 1.5)) / sum;
    }
    for (int i = tid; i < EV_P * EV_P; i += blockDim.x) {
        const int r = i / EV_P, c = i % EV_P;
        const int y = y0 + r, x = x0 + c;
        double va = 0.0, vb = 0.0;
        if (y < H && x < W) {
            va = luma(a + ((size_t)y * W + x) * 3);
            vb = luma(b + ((size_t)y * W + x) * 3);
        }
        s_a[r][c] = va;
        s_b[r][c] = vb;
    }
    __syncthreads();
    // PSNR part: every pixel belongs to exactly one tile's top-left 16x16 block
    double sq = 0.0;
    unsigned long long sqn = 0;
    {
        const int y = y0 + ly, x = x0 + lx;
        if (y >= shave && y < H - shave && x >= shave && x < W - shave) {
            const float d = (float)s_b[ly][lx] - (float)s_a[ly][lx];
            sq = (double)(d * d);
            sqn = 1;
        }
    }
    for (int i = tid; i < EV_P * EV_T; i += blockDim.x) {       // horizontal 11-tap pass
        const int r = i / EV_T, c = i % EV_T;
        double m[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < EV_K; ++k) {
            const double g = s_g[k], va = s_a[r][c + k], vb = s_b[r][c + k];
            m[0] += g * va; m[1] += g * vb; m[2] += g * va * va; m[3] += g * vb * vb; m[4] += g * va * vb;
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) s_h[q][r][c] = m[q];
    }
    __syncthreads();
    double ss = 0.0;
    unsigned long long ssn = 0;
    if (x0 + lx < mapW && y0 + ly < mapH) {
        double m[5] = {0, 0, 0, 0, 0};
#pragma unroll
        for (int k = 0; k < EV_K; ++k) {
            const double g = s_g[k];
#pragma unroll
            for (int q = 0; q < 5; ++q) m[q] += g * s_h[q][ly + k][lx];
        }
        const double C1 = (0.01 * 255) * (0.01 * 255), C2 = (0.03 * 255) * (0.03 * 255);
        const double mu1 = m[0], mu2 = m[1];
        const double s1 = m[2] - mu1 * mu1, s2 = m[3] - mu2 * mu2, s12 = m[4] - mu1 * mu2;
        ss = ((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 * mu1 + mu2 * mu2 + C1) * (s1 + s2 + C2));
        ssn = 1;
    }
    // block reduction, then one atomic per quantity
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        ssn += __shfl_xor_sync(0xffffffffu, ssn, o);
        sqn += __shfl_xor_sync(0xffffffffu, sqn, o);
    }
    __shared__ unsigned long long s_cnt[2][EV_T * EV_T / 32];
    if ((tid & 31) == 0) { s_red[0][tid >> 5] = ss; s_red[1][tid >> 5] = sq; s_cnt[0][tid >> 5] = ssn; s_cnt[1][tid >> 5] = sqn; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0, t1 = 0;
        unsigned long long n0 = 0, n1 = 0;
        for (int w = 0; w < EV_T * EV_T / 32; ++w) { t0 += s_red[0][w]; t1 += s_red[1][w]; n0 += s_cnt[0][w]; n1 += s_cnt[1][w]; }
        if (n0) { atomicAdd(&acc->ssim_sum, t0); atomicAdd(&acc->ssim_n, n0); }
        if (n1) { atomicAdd(&acc->sq_sum, t1); atomicAdd(&acc->sq_n, n1); }
    }
}

}  // namespace mulut

using namespace mulut;

// d_work: 32 bytes of device scratch (zeroed here); out2: HOST doubles {PSNR, SSIM}.  Synchronous.
extern "C" int mulut_eval_psnr_ssim_y_u8(const uint8_t *d_gt, const uint8_t *d_img, int H, int W, int shave_border,
                                         void *d_work, double *out2, void *stream)
{
    if (!d_gt || !d_img || !d_work || !out2 || H < 1 || W < 1 || shave_border < 0) {
        set_error("eval: bad argument");
        return MULUT_E_BAD_ARG;
    }
    cudaStream_t st = (cudaStream_t)stream;
    EvalAcc *acc = static_cast<EvalAcc *>(d_work);
    MULUT_CUDA(cudaMemsetAsync(acc, 0, sizeof(EvalAcc), st));
    dim3 grid((W + EV_T - 1) / EV_T, (H + EV_T - 1) / EV_T);
    eval_psnr_ssim_kernel<<<grid, EV_T * EV_T, 0, st>>>(d_gt, d_img, H, W, shave_border, acc);
    MULUT_CUDA(cudaGetLastError());
    EvalAcc h;
    MULUT_CUDA(cudaMemcpyAsync(&h, acc, sizeof h, cudaMemcpyDeviceToHost, st));
    MULUT_CUDA(cudaStreamSynchronize(st));
    // PSNR: 20 log10(255 / sqrt(mean(diff^2)))  (inf for identical frames, as numpy prints)
    const double mse = h.sq_n ? h.sq_sum / (double)h.sq_n : 0.0;
    out2[0] = mse > 0.0 ? 20.0 * log10(255.0 / sqrt(mse)) : INFINITY;
    out2[1] = h.ssim_n ? h.ssim_sum / (double)h.ssim_n : NAN;
    return MULUT_OK;
}
