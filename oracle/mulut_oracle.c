/*
 * CPU oracle for the MuLUT LUT-retrieval inference path, plain C.
 * TEST INFRASTRUCTURE ONLY: linked/called only from tests/, from
 * __graft_entry__.smoke() and from bench.py's cpu_baseline / --impl reference
 * legs.  The product library (libmulut_b200.so) never links or calls this.
 *
 * Restates (all-integer) the reference's numpy path:
 *   FourSimplexInterpFaster            /root/reference/sr/4_test_lut.py:14-237
 *   stage/mode/rotation loop+epilogue  /root/reference/sr/4_test_lut.py:279-306
 * The reference rotates the image (np.rot90), pads bottom/right with 'edge',
 * interpolates and rotates back; here the rotation is folded into the tap
 * offsets and the sub-pixel placement, and the padding into coordinate clamps
 * (SURVEY.md 8-SPEC).  Parity is pinned by tests/test_oracle_pinned.py against
 * the reference's golden PNGs and against oracle/mulut_oracle.py.
 *
 * Threading: pthreads over image rows (the reference parallelises over images
 * with multiprocessing.Pool, 4_test_lut.py:257-259; rows are the same thing
 * for a batch of frames and let one frame use every host core).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#define ORACLE_MAX_MODES 8

static int mode_taps(char mode, int dy[4], int dx[4]) {
    /* 4_test_lut.py:18-51 */
    switch (mode) {
    case 's': { int y[4] = {0, 0, 1, 1}, x[4] = {0, 1, 0, 1}; memcpy(dy, y, sizeof y); memcpy(dx, x, sizeof x); return 0; }
    case 'd': { int y[4] = {0, 0, 2, 2}, x[4] = {0, 2, 0, 2}; memcpy(dy, y, sizeof y); memcpy(dx, x, sizeof x); return 0; }
    case 'y': { int y[4] = {0, 1, 1, 2}, x[4] = {0, 1, 2, 1}; memcpy(dy, y, sizeof y); memcpy(dx, x, sizeof x); return 0; }
    default: return -1;   /* "Mode {} not implemented." 4_test_lut.py:52-54 */
    }
}

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* integer round-half-to-even of num/den, den > 0 (np.round, 4_test_lut.py:302) */
static inline long rhe_div(long num, long den) {
    long qd = num / den, rm = num % den;
    if (rm < 0) { rm += den; qd -= 1; }
    if (2 * rm > den || (2 * rm == den && (qd & 1))) qd += 1;
    return qd;
}

typedef struct {
    const uint8_t *in; uint8_t *out; int H, W, C, n_modes; const int8_t *const *luts;
    int up, interval, last;
    int tdy[ORACLE_MAX_MODES][4][4], tdx[ORACLE_MAX_MODES][4][4];
    int sub[4][16][2];
    int next_row;            /* shared work counter (rows handed out in chunks) */
} stage_ctx;

static void stage_rows(stage_ctx *cx, int y0, int y1);

static void *stage_worker(void *arg)
{
    stage_ctx *cx = (stage_ctx *)arg;
    for (;;) {
        int y0 = __atomic_fetch_add(&cx->next_row, 4, __ATOMIC_RELAXED);
        if (y0 >= cx->H) break;
        stage_rows(cx, y0, y0 + 4 < cx->H ? y0 + 4 : cx->H);
    }
    return NULL;
}

/* One stage over one frame.  in: H x W x C uint8; out: (H*up) x (W*up) x C uint8. */
static void oracle_stage(const uint8_t *in, uint8_t *out, int H, int W, int C,
                         int n_modes, const char *modes, const int8_t *const *luts,
                         int up, int interval, int last, int nthreads)
{
    stage_ctx ctx, *cx = &ctx;
    cx->in = in; cx->out = out; cx->H = H; cx->W = W; cx->C = C; cx->n_modes = n_modes;
    cx->luts = luts; cx->up = up; cx->interval = interval; cx->last = last; cx->next_row = 0;
    int (*tdy)[4][4] = cx->tdy, (*tdx)[4][4] = cx->tdx;
    int (*sub)[16][2] = cx->sub;
    for (int m = 0; m < n_modes; ++m) {
        int dy[4], dx[4];
        mode_taps(modes[m], dy, dx);
        for (int r = 0; r < 4; ++r)
            for (int k = 0; k < 4; ++k) {
                int a = dy[k], b = dx[k];
                for (int i = 0; i < r; ++i) { int t = a; a = b; b = -t; }   /* (dy,dx) <- (dx,-dy) */
                tdy[m][r][k] = a; tdx[m][r][k] = b;
            }
    }
    for (int r = 0; r < 4; ++r)
        for (int u = 0; u < up; ++u)
            for (int v = 0; v < up; ++v) {
                int a = u, b = v;
                for (int i = 0; i < r; ++i) { int t = a; a = b; b = up - 1 - t; }  /* (u,v) <- (v,up-1-u) */
                sub[r][u * up + v][0] = a; sub[r][u * up + v][1] = b;
            }
    if (nthreads <= 1) { stage_rows(cx, 0, H); return; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    int started = 0;
    for (int i = 0; i < nthreads - 1; ++i)
        if (pthread_create(&th[started], NULL, stage_worker, cx) == 0) ++started;
    stage_worker(cx);
    for (int i = 0; i < started; ++i) pthread_join(th[i], NULL);
}

static void stage_rows(stage_ctx *cx, int y0, int y1)
{
    const uint8_t *in = cx->in; uint8_t *out = cx->out;
    const int H = cx->H, W = cx->W, C = cx->C, n_modes = cx->n_modes, up = cx->up;
    const int interval = cx->interval, last = cx->last;
    const int8_t *const *luts = cx->luts;
    int (*tdy)[4][4] = cx->tdy, (*tdx)[4][4] = cx->tdx;
    int (*sub)[16][2] = cx->sub;
    const int q = 1 << interval, L = (1 << (8 - interval)) + 1;
    const int strides[4] = {L * L * L, L * L, L, 1};
    const int up2 = up * up;
    const long den = last ? (long)q * n_modes : (long)q * n_modes * 4;
    for (int y = y0; y < y1; ++y) {
        long acc[16];
        for (int x = 0; x < W; ++x)
            for (int c = 0; c < C; ++c) {
                for (int j = 0; j < up2; ++j) acc[j] = 0;
                for (int m = 0; m < n_modes; ++m) {
                    const int8_t *lut = luts[m];
                    for (int r = 0; r < 4; ++r) {
                        int f[4], s[4], v0 = 0;
                        for (int k = 0; k < 4; ++k) {
                            int yy = clampi(y + tdy[m][r][k], 0, H - 1);
                            int xx = clampi(x + tdx[m][r][k], 0, W - 1);
                            int t = in[((size_t)yy * W + xx) * C + c];
                            v0 += (t >> interval) * strides[k];
                            f[k] = t & (q - 1);
                            s[k] = strides[k];
                        }
                        /* sort descending by fraction; ties -> higher tap first */
                        for (int i = 1; i < 4; ++i)
                            for (int j = i; j > 0 && f[j] >= f[j - 1]; --j) {
                                int t = f[j]; f[j] = f[j - 1]; f[j - 1] = t;
                                t = s[j]; s[j] = s[j - 1]; s[j - 1] = t;
                            }
                        int wk[5] = {q - f[0], f[0] - f[1], f[1] - f[2], f[2] - f[3], f[3]};
                        int vk = v0;
                        for (int k = 0; k < 5; ++k) {
                            const int8_t *row = lut + (size_t)vk * up2;
                            for (int j = 0; j < up2; ++j) {
                                int *p = sub[r][j];
                                acc[p[0] * up + p[1]] += (long)wk[k] * row[j];
                            }
                            if (k < 4) vk += s[k];
                        }
                    }
                }
                for (int u = 0; u < up; ++u)
                    for (int v = 0; v < up; ++v) {
                        long S = acc[u * up + v];
                        long xo = last ? rhe_div(S, den) : rhe_div(S + 127 * den, den);
                        xo = xo < 0 ? 0 : (xo > 255 ? 255 : xo);
                        out[(((size_t)(y * up + u)) * (W * up) + (x * up + v)) * C + c] = (uint8_t)xo;
                    }
            }
    }
}

int mulut_oracle_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

/*
 * Whole path for N frames (N,H,W,C uint8, interleaved) -> (N,H*scale,W*scale,C).
 * luts: stage-major, mode-minor array of int8 tables, (L^4 x 1) for non-last
 * stages and (L^4 x scale^2) for the last one, C order.
 * returns 0, or -1 bad mode, -2 bad argument.
 */
int mulut_oracle_sr_u8(const uint8_t *in, uint8_t *out, int N, int H, int W, int C,
                       int stages, const char *modes, int scale, int interval,
                       const int8_t *const *luts, int nthreads)
{
    if (!in || !out || !modes || !luts || N < 0 || H <= 0 || W <= 0 || C <= 0 ||
        stages < 1 || scale < 1 || scale > 4 || interval < 1 || interval > 7)
        return -2;
    int n_modes = (int)strlen(modes);
    if (n_modes < 1 || n_modes > ORACLE_MAX_MODES) return -2;
    int dy[4], dx[4];
    for (int m = 0; m < n_modes; ++m)
        if (mode_taps(modes[m], dy, dx)) return -1;
    if (nthreads <= 0) nthreads = mulut_oracle_max_threads();
    size_t fin = (size_t)H * W * C, fout = fin * scale * scale;
    uint8_t *tmp_a = NULL, *tmp_b = NULL;
    if (stages > 1) {
        tmp_a = (uint8_t *)malloc(fin);
        tmp_b = (uint8_t *)malloc(fin);
        if (!tmp_a || !tmp_b) { free(tmp_a); free(tmp_b); return -2; }
    }
    for (int n = 0; n < N; ++n) {
        const uint8_t *cur = in + (size_t)n * fin;
        for (int s = 0; s < stages; ++s) {
            int last = (s + 1 == stages);
            uint8_t *dst = last ? out + (size_t)n * fout : ((s & 1) ? tmp_b : tmp_a);
            oracle_stage(cur, dst, H, W, C, n_modes, modes, luts + (size_t)s * n_modes,
                         last ? scale : 1, interval, last, nthreads);
            cur = dst;
        }
    }
    free(tmp_a); free(tmp_b);
    return 0;
}

