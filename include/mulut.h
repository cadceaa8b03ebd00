/*
 * mulut.h — C ABI of libmulut_b200.so: the B200-native (sm_100a) replacement for
 * MuLUT's LUT-retrieval hot path.  Plain pointers and sizes only; no torch or
 * numpy types cross this boundary.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository root).
 *
 * Conventions
 *   - every function returns MULUT_OK (0) or a negative MULUT_E_* code;
 *     mulut_last_error() returns a thread-local, human-readable message.
 *   - "d_" pointers are device pointers on the handle's device, "h_" pointers
 *     are host pointers.  All buffers are caller-owned; the library owns only
 *     its LUT replica and a reusable workspace inside the handle.
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *     Calls are asynchronous on that stream unless stated otherwise.
 *   - LUT tables use the reference's on-disk layout unchanged: int8, C order,
 *     (L^4, 1) for non-last stages and (L^4, scale^2) for the last stage,
 *     L = 2^(8-interval)+1, row = a*L^3 + b*L^2 + c*L + d, column = u*scale+v
 *     (sr/2_transfer_to_lut.py:12-42, sr/4_test_lut.py:61-63,323-333).
 */
#ifndef MULUT_B200_H
#define MULUT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MULUT_OK            0
#define MULUT_E_BAD_MODE   -1   /* reference: ValueError("Mode {} not implemented.") sr/4_test_lut.py:52-54 */
#define MULUT_E_BAD_ARG    -2   /* null pointer / bad shape / bad scale / bad interval */
#define MULUT_E_CUDA       -3   /* a CUDA runtime call failed (see mulut_last_error) */
#define MULUT_E_NOMEM      -4
#define MULUT_E_LUT_SMALL  -5   /* reference: numpy IndexError when a LUT has < L^4 rows */

#define MULUT_MAX_MODES     8
#define MULUT_MAX_STAGES    8

typedef struct mulut_handle_s *mulut_handle_t;

/* Kernel selection for mulut_sr_infer_u8 (parity tests run all of them). */
#define MULUT_KERNEL_AUTO     (-1)  /* fastest applicable */
#define MULUT_KERNEL_GENERIC    0   /* any modes/scale/interval/C; LUTs gathered vertex-major from L2 */
#define MULUT_KERNEL_TILED      1   /* interval 4: smem-resident up=1 LUTs, cell-major up=2 LUTs; alias of _QUAD */
#define MULUT_KERNEL_TILED_QUAD 1   /* ... up=2 last stage: four lanes fetch one 64-B cell (K1c)          */
#define MULUT_KERNEL_TILED_CELL 2   /* ... up=2 last stage: one lane fetches its cell, 2 x LDG.256 (K1d)  */
#define MULUT_KERNEL_TILED_BINNED 3 /* ... up=2 last stage: samples binned by value, LUT slabs in shared
                                       memory, TMA tile ring (K1f); needs 16-B aligned frames with
                                       W*C % 16 == 0 and C <= 4, otherwise runs K1c              */

int mulut_version(void);
const char *mulut_last_error(void);

/*
 * Replaces the LUT loading block of sr/4_test_lut.py:323-333 (the caller still
 * reads the .npy files with the reference's naming rule).  Uploads the tables
 * to `device` and builds the device-side re-layouts.  With MULUT_L2_PERSIST=1 in the environment the
 * tables are also pinned in L2 (cudaAccessPolicyWindow on the streams used by *_host; this raises the
 * device-global persisting-L2 limit until the handle is destroyed - off by default, measured unnecessary).
 *   modes      NUL-terminated string over {s,d,y}, e.g. "sdy"  (--modes)
 *   host_luts  stage-major, mode-minor: host_luts[s*strlen(modes)+m]
 *   lut_rows   rows of every table (must be >= L^4), normally 83521
 */
int mulut_create(mulut_handle_t *handle, int device, int stages, const char *modes,
                 int scale, int interval, const int8_t *const *host_luts, int lut_rows);
int mulut_destroy(mulut_handle_t handle);
int mulut_set_kernel(mulut_handle_t handle, int kernel);
/* Pre-size the workspace (stage intermediates, staging buffers) so that the
 * hot call allocates nothing. Optional: the hot calls grow it on demand. */
int mulut_reserve(mulut_handle_t handle, int N, int H, int W, int C);

/*
 * Replaces the whole per-image loop of eltr._worker, sr/4_test_lut.py:279-306
 * (stages x modes x 4 rotations of FourSimplexInterpFaster + epilogue).
 *   d_in   uint8 (N, H, W, C) interleaved (PIL / HWC order)
 *   d_out  uint8 (N, H*scale, W*scale, C)
 * Bit-exact with the reference's numpy path.  Asynchronous on `stream`.  One handle serves up to 8
 * distinct streams concurrently (each gets its own intermediates); the host side of a call is
 * serialised per handle.
 */
int mulut_sr_infer_u8(mulut_handle_t handle, const uint8_t *d_in, uint8_t *d_out,
                      int N, int H, int W, int C, void *stream);

/* Same, from/to HOST buffers (pinned memory recommended: mulut_host_alloc):
 * frames are pipelined H2D -> kernels -> D2H over internal streams.
 * Synchronous: returns when h_out is complete. */
int mulut_sr_infer_u8_host(mulut_handle_t handle, const uint8_t *h_in, uint8_t *h_out,
                           int N, int H, int W, int C);

/* Streaming form of the host path (a video pipeline; the reference keeps its worker pool busy the
 * same way, sr/4_test_lut.py:257-259): enqueue the batch and return.  Consecutive calls overlap -
 * the first frames of a call are copied in and computed while the previous call's last frames are
 * still on their way out.  h_in / h_out must stay valid (and should be pinned) until
 * mulut_sr_host_sync() returns; results of every enqueued call are complete after it. */
int mulut_sr_infer_u8_host_async(mulut_handle_t handle, const uint8_t *h_in, uint8_t *h_out,
                                 int N, int H, int W, int C);
int mulut_sr_host_sync(mulut_handle_t handle);

/* The copies of mulut_sr_infer_u8_host_async with NO kernels between them (same lanes, same per-frame
 * cudaMemcpyAsync calls, same bytes): the host-link ceiling of the streaming path on this machine.
 * Diagnostic entry for bench.py's e2e.copy_ceiling; h_out receives stale device bytes. */
int mulut_host_copy_probe_async(mulut_handle_t handle, const uint8_t *h_in, uint8_t *h_out,
                                int N, int H, int W, int C);

/* number of kernel launches issued by this handle since creation */
long long mulut_launch_count(mulut_handle_t handle);

/*
 * Per-kernel device timing for the roofline report (bench.py): while enabled,
 * every kernel launch of this handle is bracketed by CUDA events on the stream
 * it is launched on.  mulut_profile_read synchronises those events and returns
 * the summed duration and launch count of one kernel kind.
 */
#define MULUT_PROF_GENERIC_STAGE  0   /* stage_generic_kernel, non-last stage */
#define MULUT_PROF_GENERIC_LAST   1   /* stage_generic_kernel, last stage     */
#define MULUT_PROF_SMEM_STAGE     2   /* stage_smem_kernel (K1a)              */
#define MULUT_PROF_COMBINE        3   /* combine_kernel (K1b)                 */
#define MULUT_PROF_LAST_TILED     4   /* tiled last-stage kernel (K1c)        */
#define MULUT_PROF_BIN_HIST       5   /* K1f preparation: histogram, plan, orphan list */
#define MULUT_PROF_LAST_BINNED    6   /* stage_last2_binned_kernel (K1f)               */
#define MULUT_PROF_BIN_ORPHANS    7   /* stage_generic_list_kernel: K1f's sparse bins  */
#define MULUT_PROF_FUSED_STAGE    8   /* stage_pair_fused_kernel (K1i): stage + mode combine + epilogue */
#define MULUT_PROF_KINDS          9
int mulut_profile_enable(mulut_handle_t handle, int on);
int mulut_profile_read(mulut_handle_t handle, int kind, double *total_ms, long long *launches);

/*
 * Signature-compatible single pass: FourSimplexInterpFaster(weight, img_in, h, w,
 * interval, rot, upscale, mode), sr/4_test_lut.py:14-237.
 *   d_weight  float32 (n_rows, upscale^2)  (int8 values stored as float32, :333)
 *   d_img_in  float32 (C, h+p, w+p), already rotated and edge padded (p = 1 for s, 2 for d,y)
 *   d_out     float64 (C, h*up, w*up) if rot is even else (C, w*up, h*up): np.rot90(out, rot, [1,2]) / q
 */
int mulut_interp_pass_f64(const float *d_weight, int n_rows, const float *d_img_in,
                          int C, int h, int w, int interval, int rot, int upscale,
                          char mode, double *d_out, void *stream);

/*
 * MuLUT.InterpTorchBatch forward, sr/model.py:69-287.
 *   d_weight  float32 (n_rows, up^2), the raw nn.Parameter (quantised on the fly:
 *             clamp(round_half_even(w*127), -127, 127), model.py:74-76)
 *   d_img_in  float32 (B, C, h+bd, w+bd);   d_out float32 (B, C, h*up, w*up)
 */
int mulut_interp_fwd_f32(const float *d_weight, int n_rows, int up, char mode,
                         const float *d_img_in, int B, int C, int h, int w, int bd,
                         int interval, float *d_out, void *stream);
/*
 * Backward of the above (what torch autograd derives for model.py:69-287).
 *   d_grad_out     float32 (B, C, h*up, w*up)
 *   d_grad_weight  float32 (n_rows, up^2), ACCUMULATED INTO (+=); may be NULL
 *   d_grad_img_in  float32 (B, C, h+bd, w+bd), ACCUMULATED INTO (+=); may be NULL
 * Ties between LSB fractions are broken "higher tap index first", which is what
 * the reference's 24 strict-inequality cases resolve to.
 */
int mulut_interp_bwd_f32(const float *d_weight, int n_rows, int up, char mode,
                         const float *d_img_in, int B, int C, int h, int w, int bd,
                         int interval, const float *d_grad_out,
                         float *d_grad_weight, float *d_grad_img_in, void *stream);

/*
 * K4: one whole stage of MuLUT.forward, sr/model.py:296-310, fused: for every mode (in the
 * order of `modes`) and rotation 0..3
 *     pred = round(pred + rot90_back(InterpTorchBatch(weight_mode, up, mode, pad(rot90(x, r)))))
 * then  x' = round(clamp(pred / avg + bias, 0, 255))  - no rotated or padded copies are made.
 *   d_weights  n_modes raw parameter tables, float32 (n_rows, up^2) each
 *   d_x        float32 (B, C, h, w), integer-valued 0..255 (the reference's x * 255).  PRECONDITION: the
 *              kernel works on the integer grid and rounds every sample to the nearest integer first; the
 *              reference flow only ever feeds integers (an input that is off by a float rounding error
 *              gives the reference's result to ~1e-5).  Truly fractional inputs need the per-pass
 *              entry points mulut_interp_fwd_f32 / _bwd_f32, which keep float fractions.
 *   d_out      float32 (B, C, h*up, w*up) = x'
 *   d_mask     uint8, same shape: 1 where 0 <= pred/avg + bias <= 255 (clamp passes the gradient)
 *   d_workspace  mulut_stage_workspace_bytes() bytes, 16-byte aligned, caller-owned: the forward
 *              writes the quantised tables clamp(round(w*127), -127, 127) (model.py:74-76) there as
 *              int8 rows + clamp flags; hand the SAME buffer to the matching backward call (which also uses
 *              its tail as scratch: copies of the gradient tables for inputs that put most updates on a few
 *              hot LUT rows - one backward call per workspace at a time).
 */
size_t mulut_stage_workspace_bytes(int n_modes, int n_rows, int up);
int mulut_stage_fwd_f32(const float *const *d_weights, int n_modes, const char *modes, int n_rows,
                        int up, int interval, const float *d_x, int B, int C, int h, int w,
                        float avg, float bias, float *d_out, uint8_t *d_mask, void *d_workspace,
                        void *stream);
/*
 * Backward of the fused stage (what autograd derives for model.py:296-310 with BPDA rounding,
 * model.py:59-67): G_pred = grad_out * mask / avg feeds all 4*n_modes passes.
 *   d_grad_weights  n_modes tables float32 (n_rows, up^2), ACCUMULATED INTO; entries may be NULL
 *   d_grad_x        float32 (B, C, h, w), ACCUMULATED INTO; may be NULL
 *   d_workspace     the forward call's buffer (const as far as the quantised tables go; see above)
 */
int mulut_stage_bwd_f32(const float *const *d_weights, int n_modes, const char *modes, int n_rows,
                        int up, int interval, const float *d_x, int B, int C, int h, int w,
                        float avg, float bias, const float *d_grad_out, const uint8_t *d_mask,
                        const void *d_workspace, float *const *d_grad_weights, float *d_grad_x,
                        void *stream);

/*
 * The optimiser step of sr/3_finetune_lut.py:85-87,134 (torch.optim.Adam, betas/eps as given there, L2
 * weight decay, no amsgrad) fused into one pass over flat buffers of n floats.
 *   d_lr    device scalar: the learning rate of this step (LambdaLR value, :88-95)
 *   d_step  device scalar float: step counter, incremented by the call (bias corrections use the new value)
 * All pointers are device pointers; asynchronous on `stream`, capturable in a CUDA graph.
 */
int mulut_adam_step_f32(float *d_param, const float *d_grad, float *d_exp_avg, float *d_exp_avg_sq, size_t n,
                        const float *d_lr, float beta1, float beta2, float eps, float weight_decay,
                        float *d_step, void *stream);

/*
 * Loss head of the training step: `x / 255.0` (sr/model.py:312) followed by F.mse_loss(pred, label)
 * (sr/3_finetune_lut.py:132) and their backward, one kernel per direction.
 *   d_x      float32, n elements: the last stage's output (0..255), 16-byte aligned
 *   d_label  float32, n elements in [0, 1]
 *   scale    1/255
 *   forward : *d_loss = mean((x * scale - label)^2);  d_work16 = 16 bytes of device scratch
 *   backward: d_grad_x[i] = *d_grad_loss * 2 / n * scale * (x[i] * scale - label[i])   (overwritten)
 */
int mulut_mse_head_fwd_f32(const float *d_x, const float *d_label, size_t n, float scale, void *d_work16,
                           float *d_loss, void *stream);
int mulut_mse_head_bwd_f32(const float *d_x, const float *d_label, size_t n, float scale,
                           const float *d_grad_loss, float *d_grad_x, void *stream);

/*
 * On-device report metrics of eltr._worker, sr/4_test_lut.py:309-314: PSNR and SSIM on the BT.601
 * luma of two RGB uint8 frames, definitions of common/utils.py:42-101 (_rgb2ycbcr, PSNR with
 * shave_border, cal_ssim: 11x11 Gaussian window sigma 1.5, 'valid', float64).
 *   d_gt, d_img  uint8 (H, W, 3) device frames;  d_work  32 bytes of device scratch
 *   out2         HOST: {PSNR in dB, mean SSIM}.  Synchronous on `stream`.
 */
int mulut_eval_psnr_ssim_y_u8(const uint8_t *d_gt, const uint8_t *d_img, int H, int W, int shave_border,
                              void *d_work, double *out2, void *stream);

/*
 * The CTA allocation of the binned last-stage kernel (K1f), computed on the HOST with the same code
 * the device runs: given the 8-bin histogram of the stage input (bin = sample >> 5), the number of
 * tiles, the CTAs available (one per SM) and the capacity of the orphan list, returns the CTAs dealt
 * to each bin (0 = bin not resident) and the mask of bins left to the orphan list kernel.
 * Diagnostic / test entry: no device work.
 */
int mulut_plan_bins(const unsigned long long *hist8, long long n_tiles, int n_ctas,
                    unsigned long long list_cap, int allow_orphans, int *ctas_per_bin8,
                    unsigned *orphan_mask);

/* Pinned host memory for the *_host entry points. */
void *mulut_host_alloc(size_t bytes);
int mulut_host_free(void *p);

/*
 * L2/L1/shared-memory gather micro-benchmark: the measured denominator of the
 * gather roofline (no datasheet figure exists for it).
 *   variant      see MULUT_GB_* below
 *   table_bytes  size of the random-access table
 *   out[0] = gathers per second (one gather = one LUT vertex / cell fetch)
 *   out[1] = useful bytes per second, out[2] = seconds per launch
 */
#define MULUT_GB_LDG_U8        0   /* 1-byte loads, one lane per random address              */
#define MULUT_GB_LDG_U32       1   /* 4-byte loads, one lane per random address              */
#define MULUT_GB_LDG_U128      2   /* 16-byte loads, one lane per random address             */
#define MULUT_GB_QUAD_CELL64   3   /* 4 lanes x 16 B cover one random 64-byte cell           */
#define MULUT_GB_LDS_U8        4   /* 1-byte loads from a shared-memory table                */
#define MULUT_GB_PAIR_CELL64   5   /* 2 lanes x 32 B (256-bit loads) cover one 64-byte cell  */
#define MULUT_GB_OCT_CELL128   6   /* 8 lanes x 16 B cover one random 128-byte line          */
#define MULUT_GB_CPASYNC_CELL64 7  /* 4 lanes x cp.async 16 B stage a cell in smem + 5 LDS   */
#define MULUT_GB_LDS_U32       8   /* 4-byte loads from a shared-memory table                */
#define MULUT_GB_QUAD_CELL256_3ROWS 9  /* 3 of the 4 64-B row-blocks of a 256-B cell, 4 lanes x 16 B (K1e) */
#define MULUT_GB_QUAD_CELL256_4SECT 10 /* 4 of the 8 32-B sectors of a 256-B cell, one LDG.256 per lane   */
#define MULUT_GB_BULK_CELL256   11 /* one cp.async.bulk (TMA unit) of a random 256-B cell per lane into a smem ring, then LDS */
#define MULUT_GB_BULK_ROWS64X3  12 /* the same as three 64-B bulk copies (the row-blocks K1e reads)                     */
int mulut_gather_bench(int device, int variant, size_t table_bytes, int iters_per_thread,
                       int blocks_per_sm, int threads_per_block, int repeats, double *out3);

#ifdef __cplusplus
}
#endif
#endif /* MULUT_B200_H */
