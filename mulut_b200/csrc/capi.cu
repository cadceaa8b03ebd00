// C ABI of libmulut_b200.so: handle management and the inference entry points
// (include/mulut.h).  The floating-point entry points live in interp_f32.cu, the
// micro-benchmark in gather_bench.cu.
#include <stdarg.h>
#include <stdlib.h>

#include <mutex>

#include "binned.cuh"
#include "common.cuh"
#include "infer.cuh"

namespace mulut {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char *what, const char *file, int line)
{
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return e == cudaErrorMemoryAllocation ? MULUT_E_NOMEM : MULUT_E_CUDA;
}

constexpr int HOST_LANES = 3;      // frames in flight on the host path
constexpr int MAX_DEV_STREAMS = 8; // distinct caller streams one handle serves concurrently (device API)

struct Workspace {
    uint8_t *img[2] = {nullptr, nullptr};   // stage intermediates (ping-pong)
    size_t img_bytes = 0;
    int16_t *partial = nullptr;             // per-mode int16 planes of the smem-LUT kernel
    size_t partial_bytes = 0;
    void *bn_ctl = nullptr;                 // control block of the binned kernel (histogram, plan)
    uint32_t *bn_list = nullptr;            // its orphan-sample list
    size_t bn_list_cap = 0;
    uint8_t *pitched = nullptr;             // staging copy of a stage input whose rows TMA cannot map in place
    size_t pitched_bytes = 0;
    void *fused = nullptr;                  // K1i: L2-resident exchange ring + hand-off counters (kept zeroed by the kernel)
};

}  // namespace mulut

using namespace mulut;

struct mulut_handle_s {
    int device = 0, stages = 0, n_modes = 0, scale = 0, interval = 0, lut_rows = 0, num_sms = 0;
    char modes[MULUT_MAX_MODES + 1] = {0};
    int kernel = MULUT_KERNEL_AUTO;
    uint8_t *d_luts = nullptr;              // one allocation: reference-layout tables, then re-layouts
    size_t lut_bytes = 0;
    const int8_t *lut[MULUT_MAX_STAGES][MULUT_MAX_MODES] = {};
    const uint8_t *lut_alt[MULUT_MAX_STAGES][MULUT_MAX_MODES] = {};
    const uint8_t *lut_slab[MULUT_MAX_MODES] = {};   // last stage, up = 2 only
    TapTable taps;
    Workspace ws[HOST_LANES];               // host-path lanes
    // device API: one workspace per caller stream, so calls on different streams never share intermediates
    // ("concurrent calls on different streams allowed").  Slot 0 is what mulut_reserve sizes; the
    // first stream to call adopts it.
    Workspace dev_ws[MAX_DEV_STREAMS];
    cudaStream_t dev_stream[MAX_DEV_STREAMS] = {};
    bool dev_used[MAX_DEV_STREAMS] = {};
    std::mutex mu;                          // serialises the HOST side of every call on this handle (enqueue only)
    bool l2_limit_set = false;
    cudaStream_t lane_stream[HOST_LANES] = {};
    uint8_t *lane_in[HOST_LANES] = {}, *lane_out[HOST_LANES] = {};
    size_t lane_in_bytes = 0, lane_out_bytes = 0;
    unsigned long long host_k = 0;          // chunks issued on the host path so far (lane = host_k % HOST_LANES)
    long long launches = 0;
    Prof prof;
};

static int ws_reserve(Workspace &w, int stages, int n_modes, size_t frame_samples, bool want_partial,
                      bool want_binned = false, int fused_sms = 0)
{
    if (fused_sms > 0 && !w.fused) {        // never inside a stream capture (the caller checks): it synchronises
        const size_t fb = stage1_fused_ws_bytes(fused_sms);
        MULUT_CUDA(cudaMalloc(&w.fused, fb));
        MULUT_CUDA(cudaMemset(w.fused, 0, fb));
        MULUT_CUDA(cudaDeviceSynchronize());
    }
    if (stages > 1 && w.img_bytes < frame_samples) {
        for (int i = 0; i < 2; ++i) { cudaFree(w.img[i]); w.img[i] = nullptr; }
        w.img_bytes = 0;
        const int nbuf = stages > 2 ? 2 : 1;
        for (int i = 0; i < nbuf; ++i) MULUT_CUDA(cudaMalloc(&w.img[i], frame_samples));
        if (nbuf == 1) w.img[1] = nullptr;
        w.img_bytes = frame_samples;
    }
    const size_t need = want_partial ? frame_samples * n_modes * sizeof(int16_t) : 0;
    if (w.partial_bytes < need) {
        cudaFree(w.partial); w.partial = nullptr; w.partial_bytes = 0;
        MULUT_CUDA(cudaMalloc(&w.partial, need));
        w.partial_bytes = need;
    }
    if (want_binned) {
        if (!w.bn_ctl) MULUT_CUDA(cudaMalloc(&w.bn_ctl, binned_ctl_bytes()));
        const size_t cap = frame_samples / 4 + 1024;       // the plan keeps the orphan list below this
        if (w.bn_list_cap < cap) {
            cudaFree(w.bn_list); w.bn_list = nullptr; w.bn_list_cap = 0;
            MULUT_CUDA(cudaMalloc(&w.bn_list, cap * sizeof(uint32_t)));
            w.bn_list_cap = cap;
        }
    }
    return MULUT_OK;
}

static void ws_free(Workspace &w)
{
    cudaFree(w.img[0]); cudaFree(w.img[1]); cudaFree(w.partial); cudaFree(w.bn_ctl); cudaFree(w.bn_list);
    cudaFree(w.pitched); cudaFree(w.fused);
    w = Workspace();
}

static bool uses_tiled(const mulut_handle_s *h, int up, int C)
{
    if (h->kernel == MULUT_KERNEL_GENERIC) return false;
    return tiled_supported(up, h->interval, h->n_modes) && (C >= 1 && C <= 4);
}

// dense rows -> rows `pitch` bytes apart (the padding is never read: TMA's tensor is WC wide)
__global__ void repack_pitch_kernel(const uint8_t *__restrict__ in, uint8_t *__restrict__ out, size_t rows, int WC, int pitch)
{
    const int wpr = pitch / 4;                                     // words per output row
    const size_t words = rows * (size_t)wpr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / wpr;
        const int c = (int)(i - r * wpr) * 4;
        const uint8_t *src = in + r * (size_t)WC + c;
        uint32_t v = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (c + k < WC) v |= (uint32_t)src[k] << (8 * k);
        reinterpret_cast<uint32_t *>(out)[i] = v;
    }
}

// Where TMA can read the frames at `img`: in place when they are 16-byte aligned with a row length that is
// a multiple of 16, otherwise a pitched copy in the workspace (one extra read + write of the input, < 1 % of a stage).
static int tma_view(Workspace &w, StageArgs &a, int num_sms, cudaStream_t stream, int *launches)
{
    const int WC = a.W * a.C;
    if (tma_mappable(a.in, a.H, WC)) { a.in_tma = a.in; a.in_pitch = WC; return MULUT_OK; }
    const int pitch = (WC + 15) / 16 * 16;
    const size_t rows = (size_t)a.N * a.H, need = rows * (size_t)pitch;
    if (w.pitched_bytes < need) {
        cudaFree(w.pitched); w.pitched = nullptr; w.pitched_bytes = 0;
        MULUT_CUDA(cudaMalloc(&w.pitched, need));
        w.pitched_bytes = need;
    }
    size_t blocks = (need / 4 + 255) / 256;
    if (blocks > (size_t)num_sms * 16) blocks = (size_t)num_sms * 16;
    repack_pitch_kernel<<<(unsigned)blocks, 256, 0, stream>>>(a.in, w.pitched, rows, WC, pitch);
    MULUT_CUDA(cudaGetLastError());
    a.in_tma = w.pitched; a.in_pitch = pitch;
    *launches += 1;
    return MULUT_OK;
}

static int run_stages(mulut_handle_s *h, Workspace &w, const uint8_t *d_in, uint8_t *d_out, int N, int H, int W,
                      int C, cudaStream_t stream)
{
    const size_t samples = (size_t)N * H * W * C;
    if (samples == 0) return MULUT_OK;
    bool want_partial = false;
    for (int s = 0; s < h->stages; ++s) {
        const bool last = s + 1 == h->stages;
        const int up = last ? h->scale : 1;
        if (up == 1 && uses_tiled(h, up, C)) want_partial = true;
    }
    const bool want_binned = h->scale == 2 && h->interval == 4 && h->kernel != MULUT_KERNEL_GENERIC;
    // K1i (the opt-in fused stage kernel, MULUT_K1_FUSED=1) needs its exchange ring; it is not used (nor allocated)
    // inside a stream capture.  The int16 partial planes of K1h + K1b are reserved when that path is taken.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusActive; }
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    const bool want_fused = stage1_fused_enabled() && want_partial && h->interval == 4 && C <= 4 && !capturing &&
                            h->kernel != MULUT_KERNEL_TILED_QUAD && h->kernel != MULUT_KERNEL_TILED_CELL;
    int rc = ws_reserve(w, h->stages, h->n_modes, samples, false, want_binned, want_fused ? h->num_sms : 0);
    if (rc) return rc;

    const uint8_t *cur = d_in;
    bool planned = false;                      // the previous stage's K1b left K1f's histogram + plan in w.bn_ctl
    for (int s = 0; s < h->stages; ++s) {
        const bool last = s + 1 == h->stages;
        const int up = last ? h->scale : 1;
        StageArgs a;
        memset(&a, 0, sizeof a);
        a.in = cur;
        a.out = last ? d_out : w.img[(h->stages > 2) ? (s & 1) : 0];
        a.N = N; a.H = H; a.W = W; a.C = C;
        a.n_modes = h->n_modes; a.interval = h->interval; a.last = last ? 1 : 0; a.num_sms = h->num_sms;
        for (int m = 0; m < h->n_modes; ++m) {
            a.lut[m] = h->lut[s][m];
            a.lut_alt[m] = h->lut_alt[s][m];
            a.lut_slab[m] = last ? h->lut_slab[m] : nullptr;
            a.modes[m] = h->modes[m];
        }
        a.taps = h->taps;
        // the explicit TILED selections keep the pre-TMA stage kernel (K1a) as a cross-check of K1g
        a.no_tma = (h->kernel == MULUT_KERNEL_TILED_QUAD || h->kernel == MULUT_KERNEL_TILED_CELL) ? 1 : 0;
        int done = 1;
        // K1f (binned, shared-memory slabs) when forced, or on AUTO for launches big enough to
        // amortise one 177 KB LUT load per SM; it falls through to K1c when TMA cannot map the frames.
        const bool bin_policy = h->kernel == MULUT_KERNEL_TILED_BINNED ||
                                (h->kernel == MULUT_KERNEL_AUTO && samples >= (1u << 20));
        // the TMA-fed kernels (K1h for up = 1, K1f for the x2 last stage) read the input through a tensor map
        const bool k1f = uses_tiled(h, up, C) && binned_supported(a, up) && bin_policy;
        if (uses_tiled(h, up, C) && !a.no_tma && h->interval == 4 && (k1f || up == 1)) {
            int launches = 0;
            rc = tma_view(w, a, h->num_sms, stream, &launches);
            if (rc) return rc;
            h->launches += launches;
        }
        if (k1f && a.in_tma) {
            int launches = 0;
            done = launch_stage_binned(a, w.bn_ctl, w.bn_list, w.bn_list_cap, planned, stream, &launches, &h->prof);
            if (done < 0) return done;
            h->launches += launches;
        }
        planned = false;
        if (done == 1 && uses_tiled(h, up, C)) {
            int launches = 0;
            // K1c (quad-cooperative) is the default: measured 6.83 ms vs 7.73 ms for K1d on
            // 16 x 1080p (profiles/r01_bench_quad_vs_cell.txt); K1d only on request.
            const bool owner_only = h->kernel == MULUT_KERNEL_TILED_CELL;
            // if the NEXT stage is the last one and will run K1f on this stage's output, K1b also
            // produces K1f's histogram and plan (no extra launches)
            BinPlanArgs pa;
            memset(&pa, 0, sizeof pa);
            if (up == 1 && s + 2 == h->stages && bin_policy && w.bn_ctl) {
                StageArgs nx = a;
                nx.in = a.out; nx.last = 1;
                for (int m = 0; m < h->n_modes; ++m) nx.lut_slab[m] = h->lut_slab[m];
                if (binned_supported(nx, h->scale)) pa = binned_plan_args(nx, w.bn_ctl, w.bn_list, w.bn_list_cap);
            }
            done = 1;
            if (up == 1 && a.in_tma && w.fused && !capturing) {            // K1i: stage + combine (+ plan) in one kernel
                h->prof.begin(MULUT_PROF_FUSED_STAGE, stream);
                done = launch_stage1_fused(a, w.fused, pa.ctl ? &pa : nullptr, stream);
                if (done == 0) { h->prof.end(stream); launches = 1; } else h->prof.cancel();
                if (done < 0) return done;
            }
            if (done == 1) {
                if (up == 1) {                                             // K1h / K1a + K1b: int16 partial planes
                    rc = ws_reserve(w, h->stages, h->n_modes, samples, true, false);
                    if (rc) return rc;
                }
                done = launch_stage_tiled_ws(a, up, w.partial, stream, &launches, &h->prof, owner_only, &pa);
                if (done < 0) return done;
            }
            planned = done == 0 && pa.ctl != nullptr && up == 1;
            h->launches += launches;
        }
        if (done == 1) {
            h->prof.begin(last ? MULUT_PROF_GENERIC_LAST : MULUT_PROF_GENERIC_STAGE, stream);
            rc = launch_stage_generic(a, up, stream);
            h->prof.end(stream);
            if (rc) return rc;
            h->launches += 1;
        }
        cur = a.out;
    }
    return MULUT_OK;
}

extern "C" {

int mulut_version(void) { return 100; }
const char *mulut_last_error(void) { return g_err; }

int mulut_create(mulut_handle_t *handle, int device, int stages, const char *modes, int scale, int interval,
                 const int8_t *const *host_luts, int lut_rows)
{
    if (!handle || !modes || !host_luts) { set_error("mulut_create: null argument"); return MULUT_E_BAD_ARG; }
    *handle = nullptr;
    const int n_modes = (int)strlen(modes);
    if (stages < 1 || stages > MULUT_MAX_STAGES || n_modes < 1 || n_modes > MULUT_MAX_MODES || scale < 1 ||
        scale > 4 || interval < 1 || interval > 7) {
        set_error("mulut_create: stages=%d modes='%s' scale=%d interval=%d out of range", stages, modes, scale,
                  interval);
        return MULUT_E_BAD_ARG;
    }
    mulut_handle_s *h = new mulut_handle_s();
    if (!build_tap_table(modes, n_modes, &h->taps)) {
        for (int m = 0; m < n_modes; ++m) {
            int dy[4], dx[4];
            if (!mode_taps(modes[m], dy, dx)) { set_error("Mode %c not implemented.", modes[m]); break; }
        }
        delete h;
        return MULUT_E_BAD_MODE;
    }
    const long long L = (1 << (8 - interval)) + 1;
    if ((long long)lut_rows < L * L * L * L) {
        set_error("LUT too small: need %lld rows, have %d", L * L * L * L, lut_rows);
        delete h;
        return MULUT_E_LUT_SMALL;
    }
    for (int i = 0; i < stages * n_modes; ++i)
        if (!host_luts[i]) { set_error("mulut_create: host_luts[%d] is null", i); delete h; return MULUT_E_BAD_ARG; }

    h->device = device; h->stages = stages; h->n_modes = n_modes; h->scale = scale; h->interval = interval;
    h->lut_rows = lut_rows;
    memcpy(h->modes, modes, n_modes);

    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaSetDevice", __FILE__, __LINE__); }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaGetDeviceProperties", __FILE__, __LINE__); }
    h->num_sms = prop.multiProcessorCount;

    // one contiguous allocation: [reference-layout tables | re-layouts], 256-byte aligned pieces
    auto align256 = [](size_t b) { return (b + 255) / 256 * 256; };
    size_t off = 0, lut_off[MULUT_MAX_STAGES][MULUT_MAX_MODES], alt_off[MULUT_MAX_STAGES][MULUT_MAX_MODES];
    for (int s = 0; s < stages; ++s) {
        const int up2 = (s + 1 == stages) ? scale * scale : 1;
        for (int m = 0; m < n_modes; ++m) { lut_off[s][m] = off; off += align256((size_t)lut_rows * up2); }
    }
    for (int s = 0; s < stages; ++s) {
        const int up = (s + 1 == stages) ? scale : 1;
        const size_t ab = interval == 4 ? cell_major_bytes(up) : 0;
        for (int m = 0; m < n_modes; ++m) { alt_off[s][m] = off; off += align256(ab); }
    }
    const bool want_slabs = interval == 4 && scale == 2;
    size_t slab_off[MULUT_MAX_MODES];
    for (int m = 0; m < n_modes && want_slabs; ++m) { slab_off[m] = off; off += align256(slab_major_bytes()); }
    h->lut_bytes = off;
    e = cudaMalloc(&h->d_luts, off);
    if (e != cudaSuccess) { delete h; return cuda_fail(e, "cudaMalloc(luts)", __FILE__, __LINE__); }
    int rc = MULUT_OK;
    for (int s = 0; s < stages && !rc; ++s) {
        const int up = (s + 1 == stages) ? scale : 1;
        const int up2 = up * up;
        for (int m = 0; m < n_modes && !rc; ++m) {
            int8_t *dst = reinterpret_cast<int8_t *>(h->d_luts + lut_off[s][m]);
            e = cudaMemcpy(dst, host_luts[s * n_modes + m], (size_t)lut_rows * up2, cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { rc = cuda_fail(e, "cudaMemcpy(lut)", __FILE__, __LINE__); break; }
            h->lut[s][m] = dst;
            if (interval == 4 && cell_major_bytes(up) > 0) {
                uint8_t *alt = h->d_luts + alt_off[s][m];
                rc = build_cell_major(dst, alt, up, 0);
                h->lut_alt[s][m] = alt;
            }
            if (!rc && want_slabs && s + 1 == stages) {
                uint8_t *sl = h->d_luts + slab_off[m];
                rc = build_slab_major(dst, sl, 0);
                h->lut_slab[m] = sl;
            }
        }
    }
    if (!rc) {
        e = cudaDeviceSynchronize();
        if (e != cudaSuccess) rc = cuda_fail(e, "cudaDeviceSynchronize", __FILE__, __LINE__);
    }
    if (rc) { cudaFree(h->d_luts); delete h; return rc; }

    // L2 residency: keep the LUT allocation in the persisting carve-out for the
    // library's own streams (the tables are 1.3-14 MB; L2 is 126 MB).
    // Opt-in (MULUT_L2_PERSIST=1): the carve-out is a DEVICE-GLOBAL limit shared with every other user of the
    // context, and it measured unnecessary - without it K1e's LUT reads hit L2 at 98 % (the tables are 1.3-50 MB
    // of a 126 MB L2 and the frames stream through once); see DESIGN.md section 4.
    const char *env = getenv("MULUT_L2_PERSIST");
    const bool persist = env && env[0] == '1';
    for (int i = 0; i < HOST_LANES; ++i) {
        e = cudaStreamCreateWithFlags(&h->lane_stream[i], cudaStreamNonBlocking);
        if (e != cudaSuccess) { mulut_destroy(h); return cuda_fail(e, "cudaStreamCreate", __FILE__, __LINE__); }
    }
    if (persist && prop.persistingL2CacheMaxSize > 0) {
        size_t want = h->lut_bytes < (size_t)prop.persistingL2CacheMaxSize ? h->lut_bytes
                                                                           : (size_t)prop.persistingL2CacheMaxSize;
        size_t cur = 0;
        cudaDeviceGetLimit(&cur, cudaLimitPersistingL2CacheSize);
        if (cur >= want || cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) {
            h->l2_limit_set = cur < want;
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof attr);
            attr.accessPolicyWindow.base_ptr = h->d_luts;
            size_t win = h->lut_bytes;
            if (win > (size_t)prop.accessPolicyMaxWindowSize) win = (size_t)prop.accessPolicyMaxWindowSize;
            attr.accessPolicyWindow.num_bytes = win;
            attr.accessPolicyWindow.hitRatio = 1.0f;
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            for (int i = 0; i < HOST_LANES; ++i)
                cudaStreamSetAttribute(h->lane_stream[i], cudaStreamAttributeAccessPolicyWindow, &attr);
        }
        cudaGetLastError();   // residency is best-effort
    }
    *handle = h;
    return MULUT_OK;
}

int mulut_destroy(mulut_handle_t h)
{
    if (!h) return MULUT_OK;
    cudaSetDevice(h->device);
    for (int i = 0; i < HOST_LANES; ++i) {
        if (h->lane_stream[i]) { cudaStreamSynchronize(h->lane_stream[i]); cudaStreamDestroy(h->lane_stream[i]); }
        cudaFree(h->lane_in[i]);
        cudaFree(h->lane_out[i]);
    }
    for (auto &w : h->ws) ws_free(w);
    for (auto &w : h->dev_ws) ws_free(w);
    if (h->l2_limit_set) {                  // MULUT_L2_PERSIST=1 raised the device-global carve-out: give it back
        cudaCtxResetPersistingL2Cache();
        cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);
    }
    for (int i = 0; i < h->prof.n_events; ++i) { cudaEventDestroy(h->prof.recs[i].e0); cudaEventDestroy(h->prof.recs[i].e1); }
    cudaFree(h->d_luts);
    delete h;
    return MULUT_OK;
}

int mulut_set_kernel(mulut_handle_t h, int kernel)
{
    if (!h || kernel < MULUT_KERNEL_AUTO || kernel > MULUT_KERNEL_TILED_BINNED) {
        set_error("mulut_set_kernel: bad argument");
        return MULUT_E_BAD_ARG;
    }
    h->kernel = kernel;
    return MULUT_OK;
}

long long mulut_launch_count(mulut_handle_t h) { return h ? h->launches : 0; }

int mulut_profile_enable(mulut_handle_t h, int on)
{
    if (!h) { set_error("null handle"); return MULUT_E_BAD_ARG; }
    h->prof.n = 0;
    h->prof.open = false;
    h->prof.on = on != 0;
    return MULUT_OK;
}

int mulut_profile_read(mulut_handle_t h, int kind, double *total_ms, long long *launches)
{
    if (!h || !total_ms || !launches || kind < 0 || kind >= MULUT_PROF_KINDS) {
        set_error("mulut_profile_read: bad argument");
        return MULUT_E_BAD_ARG;
    }
    *total_ms = 0.0;
    *launches = 0;
    for (int i = 0; i < h->prof.n; ++i) {
        if (h->prof.recs[i].kind != kind) continue;
        MULUT_CUDA(cudaEventSynchronize(h->prof.recs[i].e1));
        float ms = 0.f;
        MULUT_CUDA(cudaEventElapsedTime(&ms, h->prof.recs[i].e0, h->prof.recs[i].e1));
        *total_ms += ms;
        *launches += 1;
    }
    return MULUT_OK;
}

static int check_shape(mulut_handle_t h, const void *in, const void *out, int N, int H, int W, int C)
{
    if (!h) { set_error("null handle"); return MULUT_E_BAD_ARG; }
    if (N < 0 || H < 0 || W < 0 || C < 1) { set_error("bad shape N=%d H=%d W=%d C=%d", N, H, W, C); return MULUT_E_BAD_ARG; }
    if ((size_t)N * H * W * C > 0 && (!in || !out)) { set_error("null image pointer"); return MULUT_E_BAD_ARG; }
    if ((long long)W * C * h->scale > 0x3fffffffLL || (long long)H * h->scale > 0x3fffffffLL) {
        set_error("frame too large");
        return MULUT_E_BAD_ARG;
    }
    return MULUT_OK;
}

int mulut_reserve(mulut_handle_t h, int N, int H, int W, int C)
{
    int rc = check_shape(h, (void *)1, (void *)1, N, H, W, C);
    if (rc) return rc;
    MULUT_CUDA(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> lock(h->mu);
    // every workspace a stream already uses, plus the next unclaimed one (the one a new stream - e.g. the
    // stream of a CUDA-graph capture - will adopt): after this call the hot path allocates nothing
    bool spare_done = false;
    for (int i = 0; i < MAX_DEV_STREAMS; ++i) {
        if (!h->dev_used[i]) {
            if (spare_done) continue;
            spare_done = true;
        }
        Workspace &w = h->dev_ws[i];
        rc = ws_reserve(w, h->stages, h->n_modes, (size_t)N * H * W * C, h->interval == 4,
                        h->scale == 2 && h->interval == 4,
                        (stage1_fused_enabled() && h->interval == 4 && C <= 4) ? h->num_sms : 0);
        if (rc) return rc;
        // the pitched staging copy of frames TMA cannot map in place (sized for the worst case: the caller's
        // pointer may turn out to be misaligned even when W*C is a multiple of 16)
        const size_t need = (size_t)N * H * (((size_t)W * C + 15) / 16 * 16);
        if (h->interval == 4 && C <= 4 && w.pitched_bytes < need) {
            cudaFree(w.pitched); w.pitched = nullptr; w.pitched_bytes = 0;
            MULUT_CUDA(cudaMalloc(&w.pitched, need));
            w.pitched_bytes = need;
        }
    }
    return MULUT_OK;
}

int mulut_sr_infer_u8(mulut_handle_t h, const uint8_t *d_in, uint8_t *d_out, int N, int H, int W, int C,
                      void *stream)
{
    int rc = check_shape(h, d_in, d_out, N, H, W, C);
    if (rc) return rc;
    MULUT_CUDA(cudaSetDevice(h->device));
    std::lock_guard<std::mutex> lock(h->mu);
    // the workspace of this stream: the one it used before, else the first one no stream has claimed yet
    // (slot 0 is the one mulut_reserve sized)
    cudaStream_t st = (cudaStream_t)stream;
    int slot = -1;
    for (int i = 0; i < MAX_DEV_STREAMS && slot < 0; ++i)
        if (h->dev_used[i] && h->dev_stream[i] == st) slot = i;
    for (int i = 0; i < MAX_DEV_STREAMS && slot < 0; ++i)
        if (!h->dev_used[i]) { h->dev_used[i] = true; h->dev_stream[i] = st; slot = i; }
    if (slot < 0) {
        set_error("mulut_sr_infer_u8: more than %d distinct streams on one handle (create another handle)", MAX_DEV_STREAMS);
        return MULUT_E_BAD_ARG;
    }
    return run_stages(h, h->dev_ws[slot], d_in, d_out, N, H, W, C, st);
}

// kernels = false: the same copies on the same lanes with no kernels between them - the host-link ceiling of
// the streaming path (mulut_host_copy_probe_async; bench.py's e2e.copy_ceiling)
static int host_async(mulut_handle_t h, const uint8_t *h_in, uint8_t *h_out, int N, int H, int W, int C, bool kernels)
{
    int rc = check_shape(h, h_in, h_out, N, H, W, C);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(h->mu);
    const size_t fin = (size_t)H * W * C, fout = fin * h->scale * h->scale;
    if (N == 0 || fin == 0) return MULUT_OK;
    MULUT_CUDA(cudaSetDevice(h->device));
    // Frames travel in chunks: H2D(chunk k+2), kernels(chunk k+1) and D2H(chunk k) overlap on
    // HOST_LANES streams.  Measured on B200 (PCIe 5: 55 GB/s each way, 16 x 1080p -> 4K): one frame
    // per chunk gives 14.4 Gpix/s, three give 12.1 - the exposed first H2D + kernels and the last
    // D2H grow with the chunk while D2H (7.1 ms of a 9 ms step) hides the per-launch overheads.
    // MULUT_HOST_CHUNK overrides (experiments).
    int chunk = 1;
    {
        const char *e = getenv("MULUT_HOST_CHUNK");
        if (e && atoi(e) > 0) chunk = atoi(e);
        if (chunk > N) chunk = N;
    }
    while (chunk > 1 && (size_t)chunk * fout > ((size_t)1 << 30)) --chunk;        // bound the staging buffers
    const size_t cin = fin * chunk, cout = fout * chunk;
    if (h->lane_in_bytes < cin || h->lane_out_bytes < cout) {
        for (int i = 0; i < HOST_LANES; ++i) {
            MULUT_CUDA(cudaStreamSynchronize(h->lane_stream[i]));
            cudaFree(h->lane_in[i]); cudaFree(h->lane_out[i]);
            h->lane_in[i] = h->lane_out[i] = nullptr;
        }
        h->lane_in_bytes = h->lane_out_bytes = 0;
        for (int i = 0; i < HOST_LANES; ++i) {
            MULUT_CUDA(cudaMalloc(&h->lane_in[i], cin));
            MULUT_CUDA(cudaMalloc(&h->lane_out[i], cout));
        }
        h->lane_in_bytes = cin; h->lane_out_bytes = cout;
    }
    // the lane counter lives in the handle: consecutive calls keep the round robin going, so the
    // first frames of call j+1 overlap the last D2H of call j (streams order the reuse of a lane)
    for (int n = 0; n < N; n += chunk, ++h->host_k) {
        const int lane = (int)(h->host_k % HOST_LANES);
        const int cn = N - n < chunk ? N - n : chunk;
        cudaStream_t st = h->lane_stream[lane];
        MULUT_CUDA(cudaMemcpyAsync(h->lane_in[lane], h_in + (size_t)n * fin, fin * cn, cudaMemcpyHostToDevice, st));
        if (kernels) {
            rc = run_stages(h, h->ws[lane], h->lane_in[lane], h->lane_out[lane], cn, H, W, C, st);
            if (rc) return rc;
        }
        MULUT_CUDA(cudaMemcpyAsync(h_out + (size_t)n * fout, h->lane_out[lane], fout * cn, cudaMemcpyDeviceToHost, st));
    }
    return MULUT_OK;
}

int mulut_sr_infer_u8_host_async(mulut_handle_t h, const uint8_t *h_in, uint8_t *h_out, int N, int H, int W, int C)
{
    return host_async(h, h_in, h_out, N, H, W, C, true);
}

int mulut_host_copy_probe_async(mulut_handle_t h, const uint8_t *h_in, uint8_t *h_out, int N, int H, int W, int C)
{
    return host_async(h, h_in, h_out, N, H, W, C, false);
}

int mulut_sr_host_sync(mulut_handle_t h)
{
    if (!h) { set_error("mulut_sr_host_sync: null handle"); return MULUT_E_BAD_ARG; }
    MULUT_CUDA(cudaSetDevice(h->device));
    for (int i = 0; i < HOST_LANES; ++i) MULUT_CUDA(cudaStreamSynchronize(h->lane_stream[i]));
    return MULUT_OK;
}

int mulut_sr_infer_u8_host(mulut_handle_t h, const uint8_t *h_in, uint8_t *h_out, int N, int H, int W, int C)
{
    const int rc = mulut_sr_infer_u8_host_async(h, h_in, h_out, N, H, W, C);
    if (rc) { if (h) mulut_sr_host_sync(h); return rc; }
    return mulut_sr_host_sync(h);
}

void *mulut_host_alloc(size_t bytes)
{
    void *p = nullptr;
    cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
    if (e != cudaSuccess) { cuda_fail(e, "cudaMallocHost", __FILE__, __LINE__); return nullptr; }
    return p;
}

int mulut_host_free(void *p)
{
    if (p) MULUT_CUDA(cudaFreeHost(p));
    return MULUT_OK;
}

}  // extern "C"
