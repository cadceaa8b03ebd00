// Internal interfaces between capi.cu and the inference kernels.
#pragma once
#include "common.cuh"

namespace mulut {

// Arguments of one stage launch (passed by value as a __grid_constant__).
struct StageArgs {
    const uint8_t *in;                       // (N, H, W, C) uint8
    uint8_t *out;                            // (N, H*up, W*up, C) uint8
    int N, H, W, C;
    int n_modes;
    int interval;
    int last;                                // last stage: up = scale, avg = M, bias 0
    int no_tma;                              // force the pre-TMA kernels (K1a) for cross-checks
    const uint8_t *in_tma;                   // the same frames where TMA can map them (16-B aligned, rows in_pitch apart), or null
    int in_pitch;
    int num_sms;
    const int8_t *lut[MULUT_MAX_MODES];      // reference layout: int8 (L^4, up^2)
    const uint8_t *lut_alt[MULUT_MAX_MODES]; // device re-layout used by the tiled kernels
    const uint8_t *lut_slab[MULUT_MAX_MODES];// slab-major biased re-layout of an up = 2 table (binned kernel), or null
    char modes[MULUT_MAX_MODES];
    TapTable taps;
};

// Optional per-launch event bracketing (mulut_profile_*).
struct Prof {
    bool on = false;
    struct Rec { int kind; cudaEvent_t e0, e1; };
    Rec recs[4096];
    int n = 0, n_events = 0;           // records in use / records whose events exist
    void begin(int kind, cudaStream_t st)
    {
        if (!on || n >= 4096) return;
        if (n >= n_events) {
            if (cudaEventCreate(&recs[n].e0) != cudaSuccess || cudaEventCreate(&recs[n].e1) != cudaSuccess) { on = false; return; }
            n_events = n + 1;
        }
        recs[n].kind = kind;
        cudaEventRecord(recs[n].e0, st);
        open = true;
    }
    void end(cudaStream_t st)
    {
        if (!open) return;
        cudaEventRecord(recs[n].e1, st);
        ++n;
        open = false;
    }
    void cancel() { open = false; }       // begin() without a launch: drop the record
    bool open = false;
};

int launch_stage_generic(const StageArgs &a, int up, cudaStream_t stream);

// Tiled sm_100a kernels (interval 4 only).  Return MULUT_OK, an error, or
// +1 when the configuration is not covered (caller falls back to generic).
// partial: workspace of n_modes * N*H*W*C int16 (used when up == 1).
// owner_only selects K1d (one lane per sample) over K1c (quad-cooperative) for the up = 2 last stage.
struct BinPlanArgs;   // binned.cuh
int launch_stage_tiled_ws(const StageArgs &a, int up, int16_t *partial, cudaStream_t stream, int *launches,
                          Prof *prof, bool owner_only, const BinPlanArgs *plan = nullptr);
bool tiled_supported(int up, int interval, int n_modes);

// K1g, the TMA-fed shared-memory kernel for up = 1 stages (infer_stage1.cu); same int16 partial
// planes as K1a.  Returns MULUT_OK, an error (< 0) or +1 (frames not TMA-mappable: run K1a).
bool tma_mappable(const void *base, int H, int row_bytes);   // tma.cu: can a tensor map address these frames in place?
bool stage1_tma_supported(const StageArgs &a, int up);
size_t stage1_pair_bytes();
int build_pair_table(const int8_t *d_lut_vertex_major, uint8_t *d_pair, cudaStream_t stream);
int launch_stage1_tma(const StageArgs &a, int16_t *partial, cudaStream_t stream);
// K1i: K1h with the mode combine, the stage epilogue and K1f's histogram + plan fused (cooperative launch; the
// modes exchange partial tiles through an L2-resident ring).  ws: stage1_fused_ws_bytes() of zeroed device memory.
// Returns MULUT_OK, an error (< 0) or +1 (not applicable: run launch_stage1_tma + K1b).
bool stage1_fused_enabled();            // MULUT_K1_FUSED=1 (opt-in; read on every call)
size_t stage1_fused_ws_bytes(int num_sms);
int launch_stage1_fused(const StageArgs &a, void *ws, const BinPlanArgs *plan, cudaStream_t stream);

// K1f, the binned shared-memory kernel for the up = 2 last stage (infer_binned.cu).
// binned_supported: configuration + TMA preconditions (16-byte aligned frames, W*C % 16 == 0).
// launch_stage_binned returns MULUT_OK, an error (< 0) or +1 (not applicable: caller falls back).
// ctl: binned_ctl_bytes() of device workspace; list: list_cap uint32 entries (orphan samples), may be null.
bool binned_supported(const StageArgs &a, int up);
// planned: the control block already holds histogram + plan of a.in (K1b made them: binned_plan_args).
int launch_stage_binned(const StageArgs &a, void *ctl, uint32_t *list, size_t list_cap, bool planned,
                        cudaStream_t stream, int *launches, Prof *prof);
size_t binned_ctl_bytes();
struct BinPlanArgs binned_plan_args(const StageArgs &next_stage, void *ctl, const uint32_t *list, size_t list_cap);
size_t slab_major_bytes();
int build_slab_major(const int8_t *d_lut_vertex_major, uint8_t *d_slabs, cudaStream_t stream);

// Device-side LUT re-layouts (run once at mulut_create).
//  cell-major: for each of the 16^4 cells, the 16 corner rows laid out so one
//  aligned 16*up^2-byte block holds everything an interpolation can touch.
size_t cell_major_bytes(int up);
int build_cell_major(const int8_t *d_lut_vertex_major, uint8_t *d_cells, int up, cudaStream_t stream);

}  // namespace mulut
