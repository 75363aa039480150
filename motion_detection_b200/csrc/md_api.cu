// md_api.cu -- the C ABI of libmotion_b200.so (include/motion_b200.h): context, staging, kernel orchestration.
// No CPU fallback anywhere: every compute entry point runs CUDA kernels on the context's stream.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "md_internal.h"
#include "tma.h"

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char buf_[256];                                                                           \
            snprintf(buf_, sizeof buf_, "%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            ctx->err = buf_;                                                                          \
            return MD_ERR_CUDA;                                                                       \
        }                                                                                             \
    } while (0)

#define FAIL(code, msg)     \
    do {                    \
        ctx->err = (msg);   \
        return (code);      \
    } while (0)

std::atomic<long long> g_md_launches(0);

static inline int align_up(int v, int a) { return (v + a - 1) / a * a; }

extern "C" const char *md_version(void) { return "motion_b200 0.1 (sm_100a)"; }

extern "C" int md_config_default(md_config *c)
{
    if (!c) return MD_ERR_INVALID;
    memset(c, 0, sizeof *c);
    c->width = 640; c->height = 480; c->max_batch = 1;
    c->pixel_step = 10;            // ros/launch/bag.launch (and camera/video/towercam/tu)
    c->min_vector_size = 1.0;      // ros/src/motion_detection_node.cpp:44
    c->lk_win = 40; c->lk_max_level = 5; c->lk_max_iters = 10; c->lk_eps = 0.03; c->lk_min_eig = 0.001f;   // cpp:40-44,71
    c->diff_threshold = 190;       // cpp:127
    c->morph = 1;                  // background_subtractor.cpp:31-32
    c->ego_mode = MD_EGO_RANSAC_HOMOGRAPHY;
    c->ransac_iters = 50;          // outlier_detector.cpp:250
    c->ransac_thresh = 0.5;
    c->seed = 1;
    c->vf_max_level = 4; c->vf_start_level = 0; c->vf_n1 = 2; c->vf_n2 = 2;      // cpp:422-425
    c->vf_rho = 2.8f; c->vf_alpha = 1400.f; c->vf_sigma = 1.5f;                  // cpp:427-429
    c->vf_literal = 1;
    c->cuda_graphs = 1;
    return MD_OK;
}

static int pyr_levels(int w, int h, int win, int max_level)
{
    for (int level = 0; level <= max_level; level++) {
        w = (w + 1) / 2; h = (h + 1) / 2;
        if (w <= win || h <= win) return level;
    }
    return max_level;
}

// ---- live-path workspace ------------------------------------------------------------------------------------------------
struct LiveWs {
    int *ints = nullptr;          // one arena, see live_ensure
    float *floats = nullptr;
    uint8_t *bytes = nullptr;
    int cap = 0, capF = 0;        // capacity in points / frames
    int *blockcnt, *idx, *total, *oidx, *ototal, *cand, *ncand, *label, *nclusters, *sizes, *box, *nout, *out_box, *out_size,
        *out_id, *best, *ninl, *n_dev;
    float *traj_c, *opts, *m2, *res;
    uint8_t *outl;
};

static void live_free(void *w)
{
    LiveWs *ws = (LiveWs *)w;
    if (!ws) return;
    if (ws->ints) cudaFree(ws->ints);
    if (ws->floats) cudaFree(ws->floats);
    if (ws->bytes) cudaFree(ws->bytes);
    delete ws;
}

// ---- sub-pixel phase planes of the tracked grid (k_lk_phase.cu) -----------------------------------------------------
static void phase_geometry(md_ctx *ctx)
{
    PhaseGeom &pg = ctx->pg;
    const int ps = ctx->cfg.pixel_step;
    int tz = 0;
    while (tz < 30 && !((ps >> tz) & 1)) tz++;
    size_t off = 0, px = 0;
    for (int l = 0; l < ctx->g.nlev; l++) {
        PhaseLevel &PL = pg.lv[l];
        PL.w = ctx->g.lv[l].w + 2 * MD_PH_MARGIN;
        PL.h = ctx->g.lv[l].h + 2 * MD_PH_MARGIN;
        PL.pitch = align_up(PL.w, 8);
        PL.shift = tz < l ? tz : l;
        PL.ncx = 1 << (l - PL.shift);
        PL.off = off;
        const size_t n = (size_t)PL.ncx * PL.ncx * PL.pitch * PL.h;
        off += 3 * n; px += n;
    }
    pg.pair_elems = off;
    pg.px_per_pair = px;
    ctx->phase_state = 0;
}

// Allocates the phase-plane arenas and encodes their tensor maps on the first grid-mode LK call.  The planes pay off
// when the per-point window builds they replace (P x levels x 1600 taps) outweigh the per-pixel evaluations.
static bool ensure_phase(md_ctx *ctx)
{
    if (ctx->phase_state) return ctx->phase_state > 0;
    ctx->phase_state = -1;
    if (!ctx->lk_maps.valid || ctx->cfg.lk_win != 40) return false;
    const PhaseGeom &pg = ctx->pg;
    if ((double)ctx->P * ctx->g.nlev * 1600.0 < 4.0 * (double)pg.px_per_pair) return false;
    const size_t bytes = pg.pair_elems * sizeof(int16_t) * ctx->cfg.max_batch;
    if (bytes > ((size_t)24 << 30)) return false;
    for (int l = 0; l < ctx->g.nlev; l++)
        if ((ctx->cfg.pixel_step >> pg.lv[l].shift) > 470) return false;              // k_window_sums: one lattice step must fit its 512-column block
    if (cudaMalloc((void **)&ctx->d_phase, bytes) != cudaSuccess) { cudaGetLastError(); ctx->d_phase = nullptr; return false; }
    if (cudaMalloc((void **)&ctx->d_wsum, sizeof(LkLevelRec) * (size_t)ctx->P * ctx->g.nlev * ctx->cfg.max_batch) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(ctx->d_phase); ctx->d_phase = nullptr; ctx->d_wsum = nullptr;
        return false;
    }
    for (int l = 0; l < ctx->g.nlev; l++) {
        const PhaseLevel &PL = pg.lv[l];
        const uint64_t plane = (uint64_t)PL.pitch * PL.h * 2;
        const uint64_t dims[5] = {(uint64_t)PL.pitch, (uint64_t)PL.h, 3, (uint64_t)PL.ncx * PL.ncx, (uint64_t)ctx->cfg.max_batch};
        const uint64_t str[4] = {(uint64_t)PL.pitch * 2, plane, 3 * plane, (uint64_t)pg.pair_elems * 2};
        if (!tma_encode_5d_u16(&ctx->ph_maps.ph[l], ctx->d_phase + PL.off, dims, str, MD_PH_BOX_W, 40, 2)) {
            cudaFree(ctx->d_phase); ctx->d_phase = nullptr; cudaFree(ctx->d_wsum); ctx->d_wsum = nullptr;
            return false;
        }
        ctx->ph_maps.imgJ[l] = ctx->lk_maps.imgJ[l];
    }
    ctx->ph_maps.valid = 1;
    ctx->phase_state = 1;
    return true;
}

static int ensure_scratch(md_ctx *ctx, size_t bytes)
{
    if (ctx->scratch && ctx->scratch_bytes >= bytes) return MD_OK;
    if (ctx->scratch) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->scratch); ctx->scratch = nullptr; ctx->scratch_bytes = 0; }
    bytes = (bytes + 4095) / 4096 * 4096;
    if (cudaMalloc(&ctx->scratch, bytes) != cudaSuccess) { cudaGetLastError(); ctx->scratch = nullptr; FAIL(MD_ERR_NOMEM, "out of device memory (scratch)"); }
    ctx->scratch_bytes = bytes;
    return MD_OK;
}

static void graphs_free(void *p);

static void free_ctx(md_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->scratch) cudaFree(ctx->scratch);
    void *ptrs[] = {ctx->d_img, ctx->d_der, ctx->d_frames, ctx->d_mask, ctx->d_pts_in, ctx->d_next, ctx->d_status, ctx->d_keep,
                    ctx->d_inlier_mask, ctx->d_blockcnt, ctx->d_kept_idx, ctx->d_M, ctx->d_hyp_valid, ctx->d_counts,
                    ctx->d_inliers, ctx->d_valid, ctx->d_hyp, ctx->d_partial, ctx->d_H, ctx->d_Hinv, ctx->d_stats,
                    ctx->d_traj, ctx->d_traj_len, ctx->d_phase, ctx->d_wsum, ctx->d_pair_ctr};
    graphs_free(ctx->graphs);
    for (void *p : ptrs) if (p) cudaFree(p);
    for (int i = 0; i < MD_VF_LANES; i++) {
        vf_free_workspace(ctx->vf_ws[i]);
        if (ctx->vf_stream[i]) cudaStreamDestroy(ctx->vf_stream[i]);
        if (ctx->vf_ev[i]) cudaEventDestroy(ctx->vf_ev[i]);
    }
    if (ctx->vf_fork) cudaEventDestroy(ctx->vf_fork);
    sub_free_workspace(ctx->sub_ws);
    mad_free_workspace(ctx->mad_ws);
    live_free(ctx->live_ws);
    for (int i = 0; i < 5; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i < 8; i++) { if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]); if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]); }
    for (int i = 0; i < 8; i++) { if (ctx->ev_k1[i]) cudaEventDestroy(ctx->ev_k1[i]); if (ctx->ev_lk[i]) cudaEventDestroy(ctx->ev_lk[i]); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_out) cudaEventDestroy(ctx->ev_out);
    for (int i = 0; i < 4; i++) if (ctx->ev_lv[i]) cudaEventDestroy(ctx->ev_lv[i]);
    for (int i = 0; i < 2; i++) if (ctx->aux_lv[i]) cudaStreamDestroy(ctx->aux_lv[i]);
    if (ctx->aux_pyr) cudaStreamDestroy(ctx->aux_pyr);
    if (ctx->aux_post) cudaStreamDestroy(ctx->aux_post);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

extern "C" int md_create(const md_config *cfg, int device, md_ctx **out)
{
    if (!cfg || !out) return MD_ERR_INVALID;
    *out = nullptr;
    if (cfg->width < 8 || cfg->height < 8 || cfg->width > 16384 || cfg->height > 16384) return MD_ERR_INVALID;
    if (cfg->max_batch < 1 || cfg->max_batch > 4096 || cfg->pixel_step < 1) return MD_ERR_INVALID;
    if (cfg->lk_win < 4 || cfg->lk_win > 44 || cfg->lk_max_level < 0 || cfg->lk_max_level >= MD_MAX_LEVELS) return MD_ERR_INVALID;
    if (cfg->ransac_iters < 1 || cfg->ransac_iters > MD_MAX_HYP) return MD_ERR_INVALID;
    if (cfg->ego_mode < 0 || cfg->ego_mode > 2) return MD_ERR_INVALID;
    if (cfg->vf_start_level != 0) return MD_ERR_UNSUPPORTED;
    if (cfg->flow_engine != MD_FLOW_LK && cfg->flow_engine != MD_FLOW_VARFLOW) return MD_ERR_INVALID;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return MD_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MD_ERR_CUDA;
    if (prop.major != 10) return MD_ERR_CUDA;     // sm_100a code only
    if (cudaSetDevice(device) != cudaSuccess) return MD_ERR_CUDA;

    md_ctx *ctx = new (std::nothrow) md_ctx();
    if (!ctx) return MD_ERR_NOMEM;
    ctx->cfg = *cfg;
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MD_ERR_CUDA; }
    ctx->stream = ctx->own_stream;
    bool sok = cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking) == cudaSuccess &&
               cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking) == cudaSuccess;
    {
        // the side streams of the kernel pipeline outrank the LK stream: their small kernels are dispatched as soon as an SM
        // has room, instead of queueing behind LK's tens of thousands of CTAs
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        sok = sok && cudaStreamCreateWithPriority(&ctx->aux_lv[0], cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&ctx->aux_lv[1], cudaStreamNonBlocking, hi) == cudaSuccess;
        for (int i = 0; i < 4 && sok; i++) sok = cudaEventCreateWithFlags(&ctx->ev_lv[i], cudaEventDisableTiming) == cudaSuccess;
        sok = sok && cudaStreamCreateWithPriority(&ctx->aux_pyr, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaStreamCreateWithPriority(&ctx->aux_post, cudaStreamNonBlocking, hi) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_out, cudaEventDisableTiming) == cudaSuccess;
    }
    for (int i = 0; i < 8 && sok; i++)
        sok = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_k1[i], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_lk[i], cudaEventDisableTiming) == cudaSuccess;
    if (!sok) { free_ctx(ctx); return MD_ERR_CUDA; }

    const int w = cfg->width, h = cfg->height, B = cfg->max_batch;
    PyrGeom &g = ctx->g;
    g.nlev = pyr_levels(w, h, cfg->lk_win, cfg->lk_max_level) + 1;
    // window reach (win) + the drift margin of the shared-memory J tiles staged by k_lk_tiled (k_lk.cu)
    g.padx = align_up(cfg->lk_win + 8, 16) + 16;
    g.pady = cfg->lk_win + 8;
    g.nslots = B + 1;
    size_t ioff = 0, doff = 0;
    int lw = w, lh = h;
    for (int l = 0; l < g.nlev; l++) {
        LevelGeom &L = g.lv[l];
        L.w = lw; L.h = lh;
        L.pitch = align_up(lw + 2 * g.padx, 128);
        L.rows = lh + 2 * g.pady;
        L.img_off = ioff; L.der_off = doff;
        ioff += (size_t)L.pitch * L.rows;
        doff += (size_t)L.pitch * L.rows;
        lw = (lw + 1) / 2; lh = (lh + 1) / 2;
    }
    g.slot_img_bytes = ioff;
    g.slot_der_elems = doff;
    ctx->gx = (w + cfg->pixel_step - 1) / cfg->pixel_step;
    ctx->gy = (h + cfg->pixel_step - 1) / cfg->pixel_step;
    ctx->P = ctx->gx * ctx->gy;
    ctx->pts_cap = (size_t)ctx->P;
    ctx->fpitch = align_up(w, 128);
    ctx->nblk_scan = (ctx->P + 2047) / 2048;
    ctx->nblk_acc = (ctx->P + 255) / 256 < 64 ? (ctx->P + 255) / 256 : 64;

    const size_t P = ctx->P;
    const int it = cfg->ransac_iters;
    bool ok = true;
    auto A = [&](void **p, size_t bytes) { if (ok && cudaMalloc(p, bytes ? bytes : 16) != cudaSuccess) ok = false; };
    A((void **)&ctx->d_img, g.slot_img_bytes * g.nslots);
    A((void **)&ctx->d_der, g.slot_der_elems * sizeof(short2) * g.nslots);
    A((void **)&ctx->d_mask, (size_t)B * h * ctx->fpitch);
    A((void **)&ctx->d_pts_in, sizeof(float2) * P * B);
    A((void **)&ctx->d_next, sizeof(float2) * P * B);
    A((void **)&ctx->d_status, P * B);
    A((void **)&ctx->d_keep, P * B);
    A((void **)&ctx->d_inlier_mask, P * B);
    A((void **)&ctx->d_blockcnt, sizeof(int) * ctx->nblk_scan * B);
    A((void **)&ctx->d_kept_idx, sizeof(int) * P * B);
    A((void **)&ctx->d_M, sizeof(int) * B);
    A((void **)&ctx->d_hyp_valid, sizeof(int) * it * B);
    A((void **)&ctx->d_counts, sizeof(int) * it * B);
    A((void **)&ctx->d_inliers, sizeof(int) * B);
    A((void **)&ctx->d_valid, sizeof(int) * B);
    A((void **)&ctx->d_hyp, sizeof(double) * 9 * it * B);
    A((void **)&ctx->d_partial, sizeof(double) * 24 * ctx->nblk_acc * B);
    A((void **)&ctx->d_H, sizeof(double) * 9 * B);
    A((void **)&ctx->d_Hinv, sizeof(double) * 9 * B);
    A((void **)&ctx->d_stats, sizeof(unsigned long long) * 136);
    A((void **)&ctx->d_pair_ctr, sizeof(unsigned long long));
    if (ok) {
        // derivative planes keep a ZERO frame forever (derivBorder = BORDER_CONSTANT); image frames are rewritten per build
        ok = cudaMemsetAsync(ctx->d_der, 0, g.slot_der_elems * sizeof(short2) * g.nslots, ctx->stream) == cudaSuccess &&
             cudaMemsetAsync(ctx->d_stats, 0, sizeof(unsigned long long) * 136, ctx->stream) == cudaSuccess &&
             cudaMemsetAsync(ctx->d_pair_ctr, 0, sizeof(unsigned long long), ctx->stream) == cudaSuccess &&
             cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    if (!ok) { free_ctx(ctx); return MD_ERR_NOMEM; }
    // TMA tensor maps over the padded planes (x, y, slot) for the production LK kernel (window 40)
    ctx->lk_maps.valid = 0;
    if (cfg->lk_win == 40) {
        bool mok = true;
        for (int l = 0; l < g.nlev && mok; l++) {
            const LevelGeom &L = g.lv[l];
            mok = tma_encode_3d(&ctx->lk_maps.imgI[l], 1, ctx->d_img + L.img_off, L.pitch, L.rows, g.nslots, L.pitch,
                                g.slot_img_bytes, MD_LK_I_BOX_W, 41) &&
                  tma_encode_3d(&ctx->lk_maps.imgJ[l], 1, ctx->d_img + L.img_off, L.pitch, L.rows, g.nslots, L.pitch,
                                g.slot_img_bytes, MD_LK_J_BOX_W, 41 + 2 * MD_LK_J_MARGIN_Y) &&
                  tma_encode_3d(&ctx->lk_maps.der[l], 4, ctx->d_der + L.der_off, L.pitch, L.rows, g.nslots, (uint64_t)L.pitch * 4,
                                g.slot_der_elems * 4, MD_LK_D_BOX_W, 41);
        }
        if (!mok) { free_ctx(ctx); return MD_ERR_CUDA; }
        ctx->lk_maps.valid = 1;
    }
    // K4: tensor maps over the INTERIOR of the ring's level-0 planes (the frames themselves); failure -> gather path
    {
        const uint8_t *f0 = ctx->d_img + g.lv[0].img_off + (size_t)g.pady * g.lv[0].pitch + g.padx;
        mask_encode_maps(&ctx->mask_maps, f0, f0, w, h, g.lv[0].pitch, (long long)g.slot_img_bytes, g.nslots);
    }
    phase_geometry(ctx);
    {
        // tuning / tracing aids, read once per context (not per call, and not into process-wide statics)
        const char *e = getenv("MD_PIPE_CHUNKS");
        ctx->pipe_chunks = e ? atoi(e) : 0;
        if (ctx->pipe_chunks < 2 || ctx->pipe_chunks > 8) ctx->pipe_chunks = 0;
        e = getenv("MD_PIPE_FIRST");
        ctx->pipe_first = e ? atoi(e) : 0;
        e = getenv("MD_TRACE");
        ctx->trace_calls = e ? atoi(e) : 0;
        e = getenv("MD_PIPE_BOUNDS");
        ctx->pipe_nbounds = 0;
        while (e && *e && ctx->pipe_nbounds < 8) {
            ctx->pipe_bounds[ctx->pipe_nbounds++] = atoi(e);
            e = strchr(e, ',');
            if (e) e++;
        }
        e = getenv("MD_GRAPHS");
        if (e) ctx->cfg.cuda_graphs = atoi(e) != 0;
    }
    ctx->stats.device = device;
    *out = ctx;
    return MD_OK;
}

extern "C" int md_destroy(md_ctx *ctx)
{
    if (!ctx) return MD_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_ctx(ctx);
    return MD_OK;
}

extern "C" int md_set_stream(md_ctx *ctx, void *s)
{
    if (!ctx) return MD_ERR_INVALID;
    ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
    return MD_OK;
}

extern "C" int md_sync(md_ctx *ctx)
{
    if (!ctx) return MD_ERR_INVALID;
    CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

extern "C" const char *md_last_error(const md_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" int md_grid_size(const md_ctx *ctx) { return ctx ? ctx->P : MD_ERR_INVALID; }

extern "C" int md_grid_points(const md_ctx *ctx, float *pts)
{
    if (!ctx || !pts) return MD_ERR_INVALID;
    int n = 0;
    for (int i = 0; i < ctx->cfg.width; i += ctx->cfg.pixel_step)
        for (int j = 0; j < ctx->cfg.height; j += ctx->cfg.pixel_step) { pts[2 * n] = (float)i; pts[2 * n + 1] = (float)j; n++; }
    return n;
}

extern "C" int md_pyramid_levels(const md_ctx *ctx) { return ctx ? ctx->g.nlev : MD_ERR_INVALID; }

extern "C" int md_pyramid_level_size(const md_ctx *ctx, int level, int32_t *w, int32_t *h)
{
    if (!ctx || level < 0 || level >= ctx->g.nlev) return MD_ERR_INVALID;
    if (w) *w = ctx->g.lv[level].w;
    if (h) *h = ctx->g.lv[level].h;
    return MD_OK;
}

// ---- staging helpers -----------------------------------------------------------------------------------------------
static int ensure_frames(md_ctx *ctx, int channels)
{
    if (ctx->d_frames && ctx->frames_channels >= channels) return MD_OK;
    if (ctx->d_frames) { CK(cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->d_frames); ctx->d_frames = nullptr; }
    size_t bytes = (size_t)(ctx->cfg.max_batch + 1) * ctx->cfg.height * ctx->fpitch * channels;
    CK(cudaMalloc((void **)&ctx->d_frames, bytes));
    ctx->frames_channels = channels;
    return MD_OK;
}

static inline uint8_t *slot_plane(md_ctx *ctx, int slot, int level)
{
    const LevelGeom &L = ctx->g.lv[level];
    return ctx->d_img + (size_t)slot * ctx->g.slot_img_bytes + L.img_off + (size_t)ctx->g.pady * L.pitch + ctx->g.padx;
}

// ---- K0 ------------------------------------------------------------------------------------------------------------
extern "C" int md_gray_u8(md_ctx *ctx, const uint8_t *src3, int32_t src_pitch, int32_t w, int32_t h, uint8_t *dst,
                          int32_t dst_pitch, int mem)
{
    if (!ctx) return MD_ERR_INVALID;
    if (!src3 || !dst || w < 1 || h < 1 || src_pitch < 3 * w || dst_pitch < w) FAIL(MD_ERR_INVALID, "md_gray_u8: bad arguments");
    CK(cudaSetDevice(ctx->device));
    if (mem == MD_MEM_DEVICE) {
        CK(launch_gray(src3, src_pitch, w, h, dst, dst_pitch, ctx->stream));
        return MD_OK;
    }
    const size_t in_bytes = ((size_t)3 * w * h + 255) / 256 * 256;
    { int r = ensure_scratch(ctx, in_bytes + (size_t)w * h); if (r != MD_OK) return r; }
    uint8_t *d_in = (uint8_t *)ctx->scratch, *d_out = d_in + in_bytes;
    CK(cudaMemcpy2DAsync(d_in, 3 * w, src3, src_pitch, 3 * w, h, cudaMemcpyHostToDevice, ctx->stream));
    CK(launch_gray(d_in, 3 * w, w, h, d_out, w, ctx->stream));
    CK(cudaMemcpy2DAsync(dst, dst_pitch, d_out, w, w, h, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

// ---- K1 ------------------------------------------------------------------------------------------------------------
static int stage_frames(md_ctx *ctx, const uint8_t *data, int channels, int pitch, long long fstride, int count, int mem,
                        const uint8_t **d_frames, int *d_pitch, long long *d_stride)
{
    if (mem == MD_MEM_DEVICE) { *d_frames = data; *d_pitch = pitch; *d_stride = fstride; return MD_OK; }
    int r = ensure_frames(ctx, channels);
    if (r != MD_OK) return r;
    const int dp = ctx->fpitch * channels;
    const long long ds = (long long)dp * ctx->cfg.height;
    for (int f = 0; f < count; f++)
        CK(cudaMemcpy2DAsync(ctx->d_frames + f * ds, dp, data + f * fstride, pitch, (size_t)ctx->cfg.width * channels,
                             ctx->cfg.height, cudaMemcpyHostToDevice, ctx->stream));
    *d_frames = ctx->d_frames; *d_pitch = dp; *d_stride = ds;
    return MD_OK;
}

extern "C" int md_pyramid_u8(md_ctx *ctx, const uint8_t *gray, int32_t pitch, int slot, int mem)
{
    if (!ctx) return MD_ERR_INVALID;
    if (!gray || pitch < ctx->cfg.width || slot < 0 || slot >= ctx->g.nslots) FAIL(MD_ERR_INVALID, "md_pyramid_u8: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const uint8_t *df; int dp; long long ds;
    int r = stage_frames(ctx, gray, 1, pitch, 0, 1, mem, &df, &dp, &ds);
    if (r != MD_OK) return r;
    CK(launch_pyramid(ctx->g, ctx->d_img, ctx->d_der, slot, 1, df, 1, dp, ds, ctx->stream));
    if (mem == MD_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_cached = 0;
    return MD_OK;
}

extern "C" int md_pyramid_read(md_ctx *ctx, int slot, int level, uint8_t *dst, int32_t dst_pitch, int mem)
{
    if (!ctx) return MD_ERR_INVALID;
    if (!dst || slot < 0 || slot >= ctx->g.nslots || level < 0 || level >= ctx->g.nlev || dst_pitch < ctx->g.lv[level].w)
        FAIL(MD_ERR_INVALID, "md_pyramid_read: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const LevelGeom &L = ctx->g.lv[level];
    CK(cudaMemcpy2DAsync(dst, dst_pitch, slot_plane(ctx, slot, level), L.pitch, L.w, L.h,
                         mem == MD_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    if (mem == MD_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

extern "C" int md_pyramid_read_deriv(md_ctx *ctx, int slot, int level, int16_t *dst, int mem)
{
    if (!ctx) return MD_ERR_INVALID;
    if (!dst || slot < 0 || slot >= ctx->g.nslots || level < 0 || level >= ctx->g.nlev) FAIL(MD_ERR_INVALID, "md_pyramid_read_deriv: bad arguments");
    CK(cudaSetDevice(ctx->device));
    const LevelGeom &L = ctx->g.lv[level];
    const short2 *src = ctx->d_der + (size_t)slot * ctx->g.slot_der_elems + L.der_off + (size_t)ctx->g.pady * L.pitch + ctx->g.padx;
    CK(cudaMemcpy2DAsync(dst, (size_t)L.w * 4, src, (size_t)L.pitch * 4, (size_t)L.w * 4, L.h,
                         mem == MD_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    if (mem == MD_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

// ---- K2 ------------------------------------------------------------------------------------------------------------
static void fill_lk(md_ctx *ctx, LkParams &p, int prev_slot0, int next_slot0, const float2 *pts_in, int P, float2 *next,
                    uint8_t *status, int ph_pair0 = 0)
{
    p.g = ctx->g;
    p.pg = ctx->pg;
    p.ph = (!pts_in && ensure_phase(ctx)) ? ctx->d_phase : nullptr;
    p.wsum = ctx->d_wsum;
    p.ph_pair0 = ph_pair0;
    p.ph_ready = 0;
    p.stat_iters = ctx->d_stats + 8;
    p.img = ctx->d_img; p.der = ctx->d_der;
    p.prev_slot0 = prev_slot0; p.next_slot0 = next_slot0;
    p.pts_in = pts_in;
    p.ps = ctx->cfg.pixel_step; p.gy = ctx->gy;
    p.P = P;
    p.next = next; p.status = status;
    p.win = ctx->cfg.lk_win;
    int it = ctx->cfg.lk_max_iters;
    p.max_iters = it < 0 ? 0 : (it > 100 ? 100 : it);
    double eps = ctx->cfg.lk_eps;
    eps = eps < 0 ? 0 : (eps > 10 ? 10 : eps);
    p.eps2 = eps * eps;
    p.min_eig = ctx->cfg.lk_min_eig;
}

extern "C" int md_lk_flow(md_ctx *ctx, int slot_prev, int slot_next, const float *pts_in, int32_t npts, float *pts_out,
                          uint8_t *status, int mem)
{
    MD_NVTX("md_lk_flow (K2)");
    if (!ctx) return MD_ERR_INVALID;
    if (!pts_out || !status || npts < 1 || (size_t)npts > ctx->pts_cap || slot_prev < 0 || slot_next < 0 ||
        slot_prev >= ctx->g.nslots || slot_next >= ctx->g.nslots || (!pts_in && npts != ctx->P))
        FAIL(MD_ERR_INVALID, "md_lk_flow: bad arguments (npts must be <= md_grid_size)");
    CK(cudaSetDevice(ctx->device));
    const float2 *d_in = nullptr;
    float2 *d_out = (float2 *)pts_out;
    uint8_t *d_st = status;
    if (mem == MD_MEM_HOST) {
        if (pts_in) { CK(cudaMemcpyAsync(ctx->d_pts_in, pts_in, sizeof(float2) * npts, cudaMemcpyHostToDevice, ctx->stream)); d_in = ctx->d_pts_in; }
        d_out = ctx->d_next; d_st = ctx->d_status;
    } else d_in = (const float2 *)pts_in;
    LkParams p;
    fill_lk(ctx, p, slot_prev, slot_next, d_in, npts, d_out, d_st);
    CK(launch_lk(p, &ctx->lk_maps, &ctx->ph_maps, 1, ctx->stream));
    if (mem == MD_MEM_HOST) {
        CK(cudaMemcpyAsync(pts_out, d_out, sizeof(float2) * npts, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(status, d_st, npts, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return MD_OK;
}

// ---- K3 ------------------------------------------------------------------------------------------------------------
static void fill_ego(md_ctx *ctx, EgoParams &p, const float2 *pts_in, int P, const float2 *next, const uint8_t *status,
                     uint8_t *keep, int keep_given, int mode, uint32_t seed0, double *H, int *inliers, uint8_t *inlier_mask)
{
    p.P = P; p.ps = ctx->cfg.pixel_step; p.gy = ctx->gy;
    p.pts_in = pts_in; p.next = next; p.status = status; p.keep = keep; p.keep_given = keep_given;
    p.min_vec = ctx->cfg.min_vector_size;
    p.mode = mode; p.iters = ctx->cfg.ransac_iters; p.minimal = mode == MD_EGO_RANSAC_AFFINE ? 3 : 4;
    p.thr2 = ctx->cfg.ransac_thresh * ctx->cfg.ransac_thresh;
    p.seed0 = seed0;
    p.pair_ctr = nullptr;
    p.w = ctx->cfg.width; p.h = ctx->cfg.height;
    p.nblk_scan = (P + 2047) / 2048;
    p.nblk_acc = (P + 255) / 256 < 64 ? (P + 255) / 256 : 64;
    p.blockcnt = ctx->d_blockcnt; p.kept_idx = ctx->d_kept_idx; p.M = ctx->d_M;
    p.hyp = ctx->d_hyp; p.hyp_valid = ctx->d_hyp_valid; p.counts = ctx->d_counts; p.partial = ctx->d_partial;
    p.H = H ? H : ctx->d_H; p.Hinv = ctx->d_Hinv;
    p.inliers = inliers ? inliers : ctx->d_inliers; p.valid = ctx->d_valid;
    p.inlier_mask = inlier_mask;
    p.stat_tracked = ctx->d_stats + 1; p.stat_inliers = ctx->d_stats + 2;
}

extern "C" int md_fit_egomotion(md_ctx *ctx, const float *src, const float *dst, const uint8_t *status, const uint8_t *keep,
                                int32_t npts, int mode, uint32_t seed, double *H9, int32_t *num_vectors, int32_t *inliers,
                                uint8_t *inlier_mask, int mem)
{
    MD_NVTX("md_fit_egomotion (K3)");
    if (!ctx) return MD_ERR_INVALID;
    if (!src || !dst || (!status && !keep) || npts < 1 || (size_t)npts > ctx->pts_cap || mode < 0 || mode > 2)
        FAIL(MD_ERR_INVALID, "md_fit_egomotion: bad arguments (npts must be <= md_grid_size)");
    CK(cudaSetDevice(ctx->device));
    const cudaMemcpyKind in = mem == MD_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const cudaMemcpyKind outk = mem == MD_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(ctx->d_pts_in, src, sizeof(float2) * npts, in, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_next, dst, sizeof(float2) * npts, in, ctx->stream));
    if (status) CK(cudaMemcpyAsync(ctx->d_status, status, npts, in, ctx->stream));
    else CK(cudaMemsetAsync(ctx->d_status, 1, npts, ctx->stream));
    if (keep) CK(cudaMemcpyAsync(ctx->d_keep, keep, npts, in, ctx->stream));
    CK(cudaMemsetAsync(ctx->d_inlier_mask, 0, npts, ctx->stream));
    EgoParams p;
    fill_ego(ctx, p, ctx->d_pts_in, npts, ctx->d_next, ctx->d_status, ctx->d_keep, keep ? 1 : 0, mode, seed, nullptr, nullptr,
             ctx->d_inlier_mask);
    p.stat_tracked = nullptr; p.stat_inliers = nullptr;
    CK(launch_ego(p, 1, ctx->stream));
    if (H9) CK(cudaMemcpyAsync(H9, ctx->d_H, sizeof(double) * 9, outk, ctx->stream));
    if (num_vectors) CK(cudaMemcpyAsync(num_vectors, ctx->d_M, sizeof(int), outk, ctx->stream));
    if (inliers) CK(cudaMemcpyAsync(inliers, ctx->d_inliers, sizeof(int), outk, ctx->stream));
    if (inlier_mask) CK(cudaMemcpyAsync(inlier_mask, ctx->d_inlier_mask, npts, outk, ctx->stream));
    if (mem == MD_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

// ---- K4 ------------------------------------------------------------------------------------------------------------
static bool invert3_host(const double *S, double *D)
{
    // cv::invert 3x3 (cofactors), same operation order as k_ego.cu:invert3 and the oracle
    double c0 = S[4] * S[8] - S[5] * S[7], c1 = S[3] * S[8] - S[5] * S[6], c2 = S[3] * S[7] - S[4] * S[6];
    double d = S[0] * c0 - S[1] * c1 + S[2] * c2;
    if (d == 0) { for (int i = 0; i < 9; i++) D[i] = 0; return false; }
    d = 1. / d;
    double t[9];
    t[0] = c0 * d;
    t[1] = (S[2] * S[7] - S[1] * S[8]) * d;
    t[2] = (S[1] * S[5] - S[2] * S[4]) * d;
    t[3] = (S[5] * S[6] - S[3] * S[8]) * d;
    t[4] = (S[0] * S[8] - S[2] * S[6]) * d;
    t[5] = (S[2] * S[3] - S[0] * S[5]) * d;
    t[6] = c2 * d;
    t[7] = (S[1] * S[6] - S[0] * S[7]) * d;
    t[8] = (S[0] * S[4] - S[1] * S[3]) * d;
    memcpy(D, t, sizeof t);
    return true;
}

extern "C" int md_motion_mask(md_ctx *ctx, const uint8_t *prev, const uint8_t *cur, int32_t pitch, const double *H9,
                              int32_t thresh, int32_t morph, uint8_t *mask, int32_t mask_pitch, int mem)
{
    MD_NVTX("md_motion_mask (K4)");
    if (!ctx) return MD_ERR_INVALID;
    const int w = ctx->cfg.width, h = ctx->cfg.height;
    if (!prev || !cur || !H9 || !mask || pitch < w || mask_pitch < w) FAIL(MD_ERR_INVALID, "md_motion_mask: bad arguments");
    CK(cudaSetDevice(ctx->device));
    double Hinv[9];
    invert3_host(H9, Hinv);
    // d_Hinv row 0 is scratch for this call; the copy is stream ordered (a pageable source is staged before the call returns)
    double *d_hi = ctx->d_Hinv;
    CK(cudaMemcpyAsync(d_hi, Hinv, sizeof Hinv, cudaMemcpyHostToDevice, ctx->stream));
    MaskParams p;
    memset(&p, 0, sizeof p);
    p.w = w; p.h = h; p.Hinv = d_hi; p.valid = nullptr; p.thresh = thresh; p.morph = morph; p.nslots = 0;
    p.stat_mask = nullptr;
    auto user_maps = [&](const uint8_t *a, const uint8_t *c, int pt) -> const MaskTmaMaps * {
        if (ctx->mask_user_key[0] != a || ctx->mask_user_key[1] != c || ctx->mask_user_pitch != pt) {
            mask_encode_maps(&ctx->mask_maps_user, a, c, w, h, pt, 0, 1);
            ctx->mask_user_key[0] = a; ctx->mask_user_key[1] = c; ctx->mask_user_pitch = pt;
        }
        return &ctx->mask_maps_user;
    };
    if (mem == MD_MEM_HOST) {
        if (ctx->cfg.max_batch < 1) FAIL(MD_ERR_STATE, "md_motion_mask: no staging");
        int r = ensure_frames(ctx, 1);
        if (r != MD_OK) return r;
        const long long fs = (long long)ctx->fpitch * h;
        CK(cudaMemcpy2DAsync(ctx->d_frames, ctx->fpitch, prev, pitch, w, h, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemcpy2DAsync(ctx->d_frames + fs, ctx->fpitch, cur, pitch, w, h, cudaMemcpyHostToDevice, ctx->stream));
        p.prev = ctx->d_frames; p.cur = ctx->d_frames + fs; p.pitch = ctx->fpitch; p.stride = 0;
        p.mask = ctx->d_mask; p.mask_pitch = ctx->fpitch; p.mask_stride = 0;
        CK(launch_mask(p, 1, user_maps(p.prev, p.cur, p.pitch), ctx->stream));
        CK(cudaMemcpy2DAsync(mask, mask_pitch, ctx->d_mask, ctx->fpitch, w, h, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    } else {
        p.prev = prev; p.cur = cur; p.pitch = pitch; p.stride = 0;
        p.mask = mask; p.mask_pitch = mask_pitch; p.mask_stride = 0;
        CK(launch_mask(p, 1, user_maps(prev, cur, pitch), ctx->stream));
    }
    return MD_OK;
}

// ---- the chain -------------------------------------------------------------------------------------------------------
// Everything LK of pairs [p0, p1) needs: the pyramids of the chunk's `nframes` new frames (slots slot0 ...) and the phase
// planes + window sums of the pairs' first frames, spread over three streams by pyramid level -- level 0 on `s`, levels 1-2
// and levels 3+ on the two level streams (forked from / joined into `s`).  The levels only meet in the pyrDown chain, so
// when a batch starts and nothing else runs yet, the exposed head is the longest level group instead of the sum of all.
static int run_prep(md_ctx *ctx, int prev0, int p0, int p1, int slot0, int nframes, const uint8_t *frames, int channels, int fpitch,
                    long long fstride, cudaStream_t s)
{
    MD_NVTX("K1 pyramid + Scharr + phase planes + window sums");
    const PyrGeom &g = ctx->g;
    const int ns = g.nslots, nl = g.nlev;
    LkParams lp;
    bool planes = false;
    if (ctx->cfg.flow_engine != MD_FLOW_VARFLOW) {
        fill_lk(ctx, lp, (prev0 + p0) % ns, (prev0 + p0 + 1) % ns, nullptr, ctx->P, ctx->d_next, ctx->d_status, p0);
        planes = lp.ph && ctx->ph_maps.valid && lp.win == 40;
    }
    if (nl < 3) {
        if (nframes > 0) CK(launch_pyramid(g, ctx->d_img, ctx->d_der, slot0, nframes, frames, channels, fpitch, fstride, s));
        if (planes) CK(launch_lk_planes(lp, p1 - p0, s));
        return MD_OK;
    }
    cudaStream_t s1 = ctx->aux_lv[0], s2 = ctx->aux_lv[1];
    const int lmid = 2;                                          // levels 1..2 on s1, 3.. on s2
    if (nframes > 0) CK(launch_pyramid_level0(g, ctx->d_img, slot0, nframes, frames, channels, fpitch, fstride, s));
    CK(cudaEventRecord(ctx->ev_lv[0], s));
    CK(cudaStreamWaitEvent(s1, ctx->ev_lv[0], 0));
    if (nframes > 0) CK(launch_pyramid_down(g, ctx->d_img, slot0, nframes, 1, lmid, s1));
    CK(cudaEventRecord(ctx->ev_lv[1], s1));
    CK(cudaStreamWaitEvent(s2, ctx->ev_lv[1], 0));
    if (nframes > 0) {
        CK(launch_pyramid_down(g, ctx->d_img, slot0, nframes, lmid + 1, nl - 1, s2));
        CK(launch_pyramid_scharr(g, ctx->d_img, ctx->d_der, slot0, nframes, 0, 0, s));
        CK(launch_pyramid_scharr(g, ctx->d_img, ctx->d_der, slot0, nframes, 1, lmid, s1));
        CK(launch_pyramid_scharr(g, ctx->d_img, ctx->d_der, slot0, nframes, lmid + 1, nl - 1, s2));
    }
    if (planes) {
        CK(launch_lk_planes_levels(lp, p1 - p0, 0, 0, s));
        CK(launch_lk_planes_levels(lp, p1 - p0, 1, lmid, s1));
        CK(launch_lk_planes_levels(lp, p1 - p0, lmid + 1, nl - 1, s2));
    }
    CK(cudaEventRecord(ctx->ev_lv[2], s1));
    CK(cudaEventRecord(ctx->ev_lv[3], s2));
    CK(cudaStreamWaitEvent(s, ctx->ev_lv[2], 0));
    CK(cudaStreamWaitEvent(s, ctx->ev_lv[3], 0));
    return MD_OK;
}

// K2 for pairs [p0, p1) of the current batch (pyramids of the frames involved are already built) on stream s
static int run_flow(md_ctx *ctx, int prev0, int p0, int p1, float2 *d_next, uint8_t *d_status, cudaStream_t s, bool planes_done = false)
{
    MD_NVTX("K2 optical flow (LK / VarFlow)");
    const int P = ctx->P, ns = ctx->g.nslots, n = p1 - p0;
    if (ctx->cfg.flow_engine == MD_FLOW_VARFLOW) {
        // dense variational flow per pair, sampled at the grid points (the gray frames are the level-0 planes).  The pairs are
        // independent and one pair's wavefront occupies a single 16-CTA cluster, so up to MD_VF_LANES pairs run side by side,
        // each on its own workspace and stream, forked from / joined into s.
        const int lanes = n < MD_VF_LANES ? n : MD_VF_LANES;
        if (lanes > 1) {
            for (int k = 0; k < lanes; k++) {
                if (!ctx->vf_stream[k]) CK(cudaStreamCreateWithFlags(&ctx->vf_stream[k], cudaStreamNonBlocking));
                if (!ctx->vf_ev[k]) CK(cudaEventCreateWithFlags(&ctx->vf_ev[k], cudaEventDisableTiming));
            }
            if (!ctx->vf_fork) CK(cudaEventCreateWithFlags(&ctx->vf_fork, cudaEventDisableTiming));
            CK(cudaEventRecord(ctx->vf_fork, s));
            for (int k = 0; k < lanes; k++) CK(cudaStreamWaitEvent(ctx->vf_stream[k], ctx->vf_fork, 0));
        }
        for (int q = p0; q < p1; q++) {
            const int k = lanes > 1 ? (q - p0) % lanes : 0;
            cudaStream_t sk = lanes > 1 ? ctx->vf_stream[k] : s;
            const int rc = vf_compute_device(ctx, k, sk, slot_plane(ctx, (prev0 + q) % ns, 0), slot_plane(ctx, (prev0 + q + 1) % ns, 0), ctx->g.lv[0].pitch);
            if (rc != MD_OK) return rc;
            CK(vf_sample_grid(ctx, k, d_next + (size_t)q * P, d_status + (size_t)q * P, sk));
        }
        if (lanes > 1)
            for (int k = 0; k < lanes; k++) {
                CK(cudaEventRecord(ctx->vf_ev[k], ctx->vf_stream[k]));
                CK(cudaStreamWaitEvent(s, ctx->vf_ev[k], 0));
            }
    } else {
        LkParams lp;
        fill_lk(ctx, lp, (prev0 + p0) % ns, (prev0 + p0 + 1) % ns, nullptr, P, d_next + (size_t)p0 * P, d_status + (size_t)p0 * P, p0);
        lp.ph_ready = planes_done ? 1 : 0;
        CK(launch_lk(lp, &ctx->lk_maps, &ctx->ph_maps, n, s));
    }
    return MD_OK;
}

// K3 -> K4 for pairs [p0, p1) on stream s
static int run_post(md_ctx *ctx, int prev0, int p0, int p1, float2 *d_next, uint8_t *d_status, uint8_t *d_keep, uint8_t *d_mask,
                    int mask_pitch, long long mask_stride, bool want_mask, cudaStream_t s)
{
    const int P = ctx->P, ns = ctx->g.nslots, n = p1 - p0, it = ctx->cfg.ransac_iters;
    EgoParams ep;
    fill_ego(ctx, ep, nullptr, P, d_next + (size_t)p0 * P, d_status + (size_t)p0 * P, d_keep + (size_t)p0 * P, 0, ctx->cfg.ego_mode,
             ctx->cfg.seed + (uint32_t)p0, nullptr, nullptr, nullptr);
    ep.pair_ctr = ctx->d_pair_ctr;
    ep.blockcnt += (size_t)p0 * ep.nblk_scan; ep.kept_idx += (size_t)p0 * P; ep.M += p0;
    ep.hyp += (size_t)p0 * it * 9; ep.hyp_valid += (size_t)p0 * it; ep.counts += (size_t)p0 * it;
    ep.partial += (size_t)p0 * ep.nblk_acc * 24; ep.H += 9 * p0; ep.Hinv += 9 * p0; ep.inliers += p0; ep.valid += p0;
    { MD_NVTX("K3 egomotion fit"); CK(launch_ego(ep, n, s)); }
    if (ctx->profile) CK(cudaEventRecord(ctx->ev[3], s));
    if (want_mask) {
        MaskParams mp;
        memset(&mp, 0, sizeof mp);
        mp.prev = slot_plane(ctx, 0, 0); mp.cur = mp.prev;
        mp.pitch = ctx->g.lv[0].pitch; mp.stride = (long long)ctx->g.slot_img_bytes;
        mp.nslots = ns; mp.prev_slot0 = (prev0 + p0) % ns; mp.cur_slot0 = (prev0 + p0 + 1) % ns;
        mp.w = ctx->cfg.width; mp.h = ctx->cfg.height; mp.Hinv = ctx->d_Hinv + 9 * p0; mp.valid = ctx->d_valid + p0;
        mp.thresh = ctx->cfg.diff_threshold; mp.morph = ctx->cfg.morph; mp.packed = ctx->cfg.mask_packed ? 1 : 0;
        mp.mask = d_mask + (size_t)p0 * mask_stride; mp.mask_pitch = mask_pitch; mp.mask_stride = mask_stride;
        mp.stat_mask = ctx->d_stats;
        MD_NVTX("K4 warp + diff + threshold + morphology");
        CK(launch_mask(mp, n, &ctx->mask_maps, s));
    }
    if (ctx->profile) CK(cudaEventRecord(ctx->ev[4], s));
    return MD_OK;
}

// The batch is cut into chunks and software-pipelined.  K2 (LK) is instruction-issue bound and fills the GPU; the pyramid
// build (K1, bandwidth bound), the egomotion fit (K3, a chain of small latency-bound kernels) and the mask (K4) of the
// neighbouring chunks run beside it on two higher-priority streams, so their latencies hide behind LK instead of
// adding up per chunk.  With host buffers the H2D of chunk i+1 and the D2H of chunk i-1 overlap as well (PCIe is full
// duplex; the results are complete on return).
//   copy_in : H2D(c)                                (host buffers only)
//   aux_pyr (+ two level streams): K1(c) + phase planes(c) + window sums(c)  after H2D(c) / after the work already queued on the context's stream
//   stream  : LK(c)                                 after K1(c)
//   aux_post: K3(c), K4(c)     after LK(c)
//   copy_out: D2H(c)           after K4(c)          (host buffers only)
// The context's stream joins aux_post at the end, so the call stays stream-ordered for the caller.
#define MD_PIPE_CHUNKS_DEVICE 2   // device-resident frames: measured 4 193 pairs/s with 2 even chunks, 4 146 with 3, 4 060 with 4, 3 675 with 6
#define MD_PIPE_CHUNKS 7      // host buffers, 32 pairs: measured e2e 4 649 / 4 821 / 4 858 / 4 824 frames/s with 5 / 6 / 7 / 8 chunks (round 2; round 1: 3 297 ... 4 003 with 3 ... 8); a short first chunk (its K1 / H2D is exposed), even middle chunks, a short last chunk (its K3 / K4 / D2H is)

// Enqueues one batch (frames in ring slots prev0 ...) on the context's stream and its side streams; no allocation, no host
// synchronisation, every side stream joined back into the context's stream: the body can run eagerly or under stream capture.
static int batch_enqueue(md_ctx *ctx, const md_frames *fr, const md_outputs *out, int mem, int prev0, bool serial)
{
    const int pairs = fr->chain ? fr->count : fr->count - 1;
    cudaStream_t s = ctx->stream;
    const int w = ctx->cfg.width, h = ctx->cfg.height, P = ctx->P, ns = ctx->g.nslots;
    const int new0 = fr->chain ? (prev0 + 1) % ns : prev0;
    const bool host = mem != MD_MEM_DEVICE;
    const bool want_mask = out->mask != nullptr;
    int r;

    // where the frames and the results live on the device
    const uint8_t *df = fr->data;
    int dp = fr->pitch;
    long long ds = fr->frame_stride;
    float2 *d_next = ctx->d_next;
    uint8_t *d_status = ctx->d_status, *d_keep = ctx->d_keep, *d_mask = ctx->d_mask;
    // staging layout of the masks in host mode: byte rows of the frame pitch, or packed rows ((w + 7) / 8 bytes, 16-byte aligned)
    const int packed = ctx->cfg.mask_packed ? 1 : 0;
    const int mrow = packed ? (w + 7) / 8 : w;                 // payload bytes per mask row
    int mpitch = packed ? align_up(mrow, 16) : ctx->fpitch;
    long long mstride = (long long)mpitch * h;
    if (host) {
        df = ctx->d_frames; dp = ctx->fpitch * fr->channels; ds = (long long)dp * h;
    } else {
        if (out->next_pts) d_next = (float2 *)out->next_pts;
        if (out->status) d_status = out->status;
        if (out->keep) d_keep = out->keep;
        if (out->mask) { d_mask = out->mask; mpitch = out->mask_pitch; mstride = out->mask_stride; }
    }

    // chunk boundaries
    int bounds[12];                   // at most 8 chunks (the per-chunk events)
    int nch = 0;
    bounds[0] = 0;
    if (serial) { nch = 1; bounds[1] = pairs; }
    else if (ctx->pipe_nbounds > 0 && !host) {
        // MD_PIPE_BOUNDS="2,8,32": explicit chunk ends (tuning aid); ends beyond the batch are clipped, the batch end is appended
        for (int i = 0; i < ctx->pipe_nbounds && nch < 7; i++)
            if (ctx->pipe_bounds[i] > bounds[nch] && ctx->pipe_bounds[i] < pairs) { bounds[nch + 1] = ctx->pipe_bounds[i]; nch++; }
        bounds[++nch] = pairs;
    }
    else if (pairs < 8) { nch = 2; bounds[1] = pairs / 2; bounds[2] = pairs; }
    else {
        const int tuned = ctx->pipe_chunks;        // MD_PIPE_CHUNKS=<2..8> in the environment overrides the default (tuning aid)
        const int total = tuned ? tuned : (host ? MD_PIPE_CHUNKS : MD_PIPE_CHUNKS_DEVICE);
        if (host) {
            // a short first chunk (its H2D is exposed) and a short last chunk (its D2H is exposed) around even middle chunks
            const int mid = total - 2 > 0 ? total - 2 : 1;
            bounds[++nch] = 1;
            for (int i = 1; i <= mid; i++) bounds[++nch] = 1 + (int)((long long)(pairs - 2) * i / mid);
            bounds[++nch] = pairs;
        } else {
            // device-resident frames: a short first chunk (its pyramids, planes and sums are exposed), then even chunks
            const int first = ctx->pipe_first > 0 && ctx->pipe_first < pairs ? ctx->pipe_first : 0;
            if (first) bounds[++nch] = first;
            const int rest = pairs - first, nrest = first ? total - 1 : total;
            for (int i = 1; i <= nrest; i++) bounds[++nch] = first + (int)((long long)rest * i / nrest);
        }
    }
    {
        // every chunk holds at least one pair (a tuning knob or a small batch can produce empty ones: drop them)
        int m = 0;
        for (int i = 1; i <= nch; i++)
            if (bounds[i] > bounds[m]) bounds[++m] = bounds[i];
        nch = m;
    }
    cudaStream_t s_pyr = serial ? s : ctx->aux_pyr, s_post = serial ? s : ctx->aux_post;

    auto d2h = [&](int p0, int p1) -> int {
        const int n = p1 - p0;
        cudaStream_t so = ctx->copy_out;
        if (out->next_pts) CK(cudaMemcpyAsync(out->next_pts + (size_t)2 * p0 * P, ctx->d_next + (size_t)p0 * P, sizeof(float2) * P * n, cudaMemcpyDeviceToHost, so));
        if (out->status) CK(cudaMemcpyAsync(out->status + (size_t)p0 * P, ctx->d_status + (size_t)p0 * P, (size_t)P * n, cudaMemcpyDeviceToHost, so));
        if (out->keep) CK(cudaMemcpyAsync(out->keep + (size_t)p0 * P, ctx->d_keep + (size_t)p0 * P, (size_t)P * n, cudaMemcpyDeviceToHost, so));
        if (out->H) CK(cudaMemcpyAsync(out->H + 9 * p0, ctx->d_H + 9 * p0, sizeof(double) * 9 * n, cudaMemcpyDeviceToHost, so));
        if (out->num_vectors) CK(cudaMemcpyAsync(out->num_vectors + p0, ctx->d_M + p0, sizeof(int) * n, cudaMemcpyDeviceToHost, so));
        if (out->inliers) CK(cudaMemcpyAsync(out->inliers + p0, ctx->d_inliers + p0, sizeof(int) * n, cudaMemcpyDeviceToHost, so));
        if (out->mask) {
            if (out->mask_stride == (long long)out->mask_pitch * h)
                CK(cudaMemcpy2DAsync(out->mask + p0 * out->mask_stride, out->mask_pitch, ctx->d_mask + p0 * mstride, mpitch, mrow,
                                     (size_t)h * n, cudaMemcpyDeviceToHost, so));
            else
                for (int b = p0; b < p1; b++)
                    CK(cudaMemcpy2DAsync(out->mask + b * out->mask_stride, out->mask_pitch, ctx->d_mask + b * mstride, mpitch, mrow, h,
                                         cudaMemcpyDeviceToHost, so));
        }
        return MD_OK;
    };

    const bool trace = ctx->trace_calls > 0 && !serial && !ctx->capturing;
    cudaEvent_t tev[8][6];            // per chunk: K1 begin, K1+planes end, LK begin, LK end, post begin, post end
    cudaEvent_t tev0 = nullptr;
    if (trace) {
        cudaEventCreate(&tev0);
        for (int i = 0; i < nch; i++) for (int k = 0; k < 6; k++) cudaEventCreate(&tev[i][k]);
    }
    if (!serial) {
        if (trace) cudaEventRecord(tev0, s);
        // fork: the side streams start after whatever the caller already queued on the context's stream
        CK(cudaEventRecord(ctx->ev_fork, s));
        CK(cudaStreamWaitEvent(s_pyr, ctx->ev_fork, 0));
        CK(cudaStreamWaitEvent(s_post, ctx->ev_fork, 0));
    }
    if (host) {
        // the H2D stream forks from the context's stream as well: the staging buffer may still be read by work queued there, and
        // under capture every stream of the DAG has to descend from the capturing one
        if (serial) CK(cudaEventRecord(ctx->ev_fork, s));
        CK(cudaStreamWaitEvent(ctx->copy_in, ctx->ev_fork, 0));
    }
    int fdone = 0, pprev0 = 0, pprev1 = 0;
    for (int i = 0; i < nch; i++) {
        const int p0 = bounds[i], p1 = bounds[i + 1];
        const int fa = fdone, fb = fr->chain ? p1 : p1 + 1;          // new frames this chunk needs
        if (host) {
            for (int f = fa; f < fb; f++)
                CK(cudaMemcpy2DAsync(ctx->d_frames + f * ds, dp, fr->data + f * fr->frame_stride, fr->pitch, (size_t)w * fr->channels, h,
                                     cudaMemcpyHostToDevice, ctx->copy_in));
            CK(cudaEventRecord(ctx->ev_in[i], ctx->copy_in));
            CK(cudaStreamWaitEvent(s_pyr, ctx->ev_in[i], 0));
        }
        if (ctx->profile) CK(cudaEventRecord(ctx->ev[0], s));
        if (trace) cudaEventRecord(tev[i][0], s_pyr);
        if (serial) {
            if (fb > fa) CK(launch_pyramid(ctx->g, ctx->d_img, ctx->d_der, (new0 + fa) % ns, fb - fa, df + fa * ds, fr->channels, dp, ds, s_pyr));
        }
        if (ctx->profile) CK(cudaEventRecord(ctx->ev[1], s));
        if (!serial) {
            // pyramids, phase planes and window sums ride on the side streams: the context's stream runs LK kernels back to back
            r = run_prep(ctx, prev0, p0, p1, (new0 + fa) % ns, fb - fa, df + fa * ds, fr->channels, dp, ds, s_pyr);
            if (r != MD_OK) return r;
            if (trace) cudaEventRecord(tev[i][1], s_pyr);
            CK(cudaEventRecord(ctx->ev_k1[i], s_pyr));
            CK(cudaStreamWaitEvent(s, ctx->ev_k1[i], 0));
        }
        if (trace) cudaEventRecord(tev[i][2], s);
        r = run_flow(ctx, prev0, p0, p1, d_next, d_status, s, !serial);
        if (r != MD_OK) return r;
        if (trace) cudaEventRecord(tev[i][3], s);
        if (ctx->profile) CK(cudaEventRecord(ctx->ev[2], s));
        if (!serial) {
            CK(cudaEventRecord(ctx->ev_lk[i], s));
            CK(cudaStreamWaitEvent(s_post, ctx->ev_lk[i], 0));
        }
        if (trace) cudaEventRecord(tev[i][4], s_post);
        r = run_post(ctx, prev0, p0, p1, d_next, d_status, d_keep, d_mask, mpitch, mstride, want_mask, s_post);
        if (r != MD_OK) return r;
        if (trace) cudaEventRecord(tev[i][5], s_post);
        if (host) {
            CK(cudaEventRecord(ctx->ev_comp[i], s_post));
            // results of the PREVIOUS chunk are copied out only now, so that (with pageable host memory, where the copy call
            // blocks) this chunk's kernels are already queued behind it
            if (i > 0) {
                CK(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[i - 1], 0));
                r = d2h(pprev0, pprev1);
                if (r != MD_OK) return r;
            }
        }
        fdone = fb; pprev0 = p0; pprev1 = p1;
    }
    if (!serial) {
        // join: everything of this call is ordered before later work on the context's stream
        CK(cudaEventRecord(ctx->ev_join, s_post));
        CK(cudaStreamWaitEvent(s, ctx->ev_join, 0));
    }
    if (trace) {
        cudaStreamSynchronize(s);
        for (int i = 0; i < nch; i++) {
            float t[6];
            for (int k = 0; k < 6; k++) cudaEventElapsedTime(&t[k], tev0, tev[i][k]);
            fprintf(stderr, "[md trace] chunk %d pairs %2d..%2d  K1+planes %.3f-%.3f  LK %.3f-%.3f  K3+K4 %.3f-%.3f ms\n", i, bounds[i],
                    bounds[i + 1], t[0], t[1], t[2], t[3], t[4], t[5]);
            for (int k = 0; k < 6; k++) cudaEventDestroy(tev[i][k]);
        }
        cudaEventDestroy(tev0);
        ctx->trace_calls--;           // MD_TRACE=n traces the first n pipelined calls of a context
    }
    if (host) {
        CK(cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[nch - 1], 0));
        r = d2h(pprev0, pprev1);
        if (r != MD_OK) return r;
        // the copies join the context's stream: the caller (md_process_batch) synchronises it once
        CK(cudaEventRecord(ctx->ev_out, ctx->copy_out));
        CK(cudaStreamWaitEvent(s, ctx->ev_out, 0));
    } else {
        if (out->num_vectors) CK(cudaMemcpyAsync(out->num_vectors, ctx->d_M, sizeof(int) * pairs, cudaMemcpyDeviceToDevice, s));
        if (out->H) CK(cudaMemcpyAsync(out->H, ctx->d_H, sizeof(double) * 9 * pairs, cudaMemcpyDeviceToDevice, s));
        if (out->inliers) CK(cudaMemcpyAsync(out->inliers, ctx->d_inliers, sizeof(int) * pairs, cudaMemcpyDeviceToDevice, s));
    }
    CK(launch_advance_pairs(ctx->d_pair_ctr, pairs, s));
    return MD_OK;
}

// ---- captured batch graphs ---------------------------------------------------------------------------------------------------
// A batch is ~60 kernel launches and ~40 event operations over seven streams.  The second time a call with the same buffers,
// shapes and ring position arrives, its whole DAG is captured (cudaStreamBeginCapture on the context's stream; the side streams
// join the capture through their event waits) and from then on replayed with ONE cudaGraphLaunch: no per-kernel launch latency
// between the dependent small kernels at the head and the tail of a batch, no host launch jitter.  What varies from call to call
// is kept out of the kernel parameters: the ring position is constant (the last frame's pyramid is moved to slot 0 before a
// chained call), the RANSAC seeds come from the pair counter on the device.
struct BatchKey {
    const void *data; int channels, pitch; long long fstride; int count, chain, mem, prev0;
    const void *o[7]; int mask_pitch; long long mask_stride;
    const void *stream;
    bool operator==(const BatchKey &k) const
    {
        if (data != k.data || channels != k.channels || pitch != k.pitch || fstride != k.fstride || count != k.count || chain != k.chain ||
            mem != k.mem || prev0 != k.prev0 || mask_pitch != k.mask_pitch || mask_stride != k.mask_stride || stream != k.stream)
            return false;
        for (int i = 0; i < 7; i++) if (o[i] != k.o[i]) return false;
        return true;
    }
};
struct BatchGraph { BatchKey key; cudaGraphExec_t exec; long long launches; unsigned long long stamp; };
struct GraphCache {
    std::vector<BatchKey> seen, failed;      // seen once (eager) / could not be captured (e.g. pageable host buffers): stay eager
    std::vector<BatchGraph> graphs;
    unsigned long long clock = 0;
};
#define MD_GRAPH_CACHE 16

static void graphs_free(void *p)
{
    GraphCache *gc = static_cast<GraphCache *>(p);
    if (!gc) return;
    for (auto &g : gc->graphs) cudaGraphExecDestroy(g.exec);
    delete gc;
}

extern "C" int md_process_batch(md_ctx *ctx, const md_frames *fr, const md_outputs *out, int mem)
{
    MD_NVTX("md_process_batch");
    if (!ctx) return MD_ERR_INVALID;
    if (!fr || !out || !fr->data || (fr->channels != 1 && fr->channels != 3) || fr->count < 1 ||
        fr->pitch < ctx->cfg.width * fr->channels)
        FAIL(MD_ERR_INVALID, "md_process_batch: bad frame descriptor");
    const int pairs = fr->chain ? fr->count : fr->count - 1;
    if (pairs < 1 || pairs > ctx->cfg.max_batch) FAIL(MD_ERR_INVALID, "md_process_batch: pairs must be in [1, max_batch]");
    if (mem != MD_MEM_HOST && mem != MD_MEM_DEVICE && mem != MD_MEM_HOST_ASYNC) FAIL(MD_ERR_INVALID, "md_process_batch: bad mem");
    if (fr->chain && !ctx->have_cached) FAIL(MD_ERR_STATE, "md_process_batch: chain=1 without a cached previous frame");
    if (out->mask && out->mask_pitch < (ctx->cfg.mask_packed ? (ctx->cfg.width + 7) / 8 : ctx->cfg.width))
        FAIL(MD_ERR_INVALID, "md_process_batch: mask_pitch too small");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int ns = ctx->g.nslots;
    const bool host = mem != MD_MEM_DEVICE;
    if (host) { int r0 = ensure_frames(ctx, fr->channels); if (r0 != MD_OK) return r0; }
    // The ring position is the same for every call: an unchained batch rewrites every slot it uses, a chained one finds the
    // previous call's last pyramid in slot 0 (two device copies of one slot, ~16 MB at 1080p, when it is not there already).
    if (fr->chain && ctx->slot_base != 0) {
        const PyrGeom &g = ctx->g;
        CK(cudaMemcpyAsync(ctx->d_img, ctx->d_img + (size_t)ctx->slot_base * g.slot_img_bytes, g.slot_img_bytes, cudaMemcpyDeviceToDevice, s));
        CK(cudaMemcpyAsync(ctx->d_der, ctx->d_der + (size_t)ctx->slot_base * g.slot_der_elems, g.slot_der_elems * sizeof(short2),
                           cudaMemcpyDeviceToDevice, s));
    }
    const int prev0 = 0;
    ctx->slot_base = 0;
    const bool serial = ctx->profile || pairs < 2 || ctx->cfg.flow_engine == MD_FLOW_VARFLOW;

    int r = MD_OK;
    bool done = false;
    const bool graph_ok = ctx->cfg.cuda_graphs && !ctx->profile && ctx->trace_calls <= 0 && ctx->cfg.flow_engine != MD_FLOW_VARFLOW;
    if (graph_ok) {
        if (!ctx->graphs) ctx->graphs = new (std::nothrow) GraphCache();
        GraphCache *gc = static_cast<GraphCache *>(ctx->graphs);
        BatchKey key;
        memset(&key, 0, sizeof key);
        key.data = fr->data; key.channels = fr->channels; key.pitch = fr->pitch; key.fstride = fr->frame_stride; key.count = fr->count;
        key.chain = fr->chain; key.mem = mem; key.prev0 = prev0;
        key.o[0] = out->next_pts; key.o[1] = out->status; key.o[2] = out->keep; key.o[3] = out->H; key.o[4] = out->num_vectors;
        key.o[5] = out->inliers; key.o[6] = out->mask; key.mask_pitch = out->mask_pitch; key.mask_stride = out->mask_stride;
        key.stream = (const void *)s;
        BatchGraph *hit = nullptr;
        if (gc) for (auto &g : gc->graphs) if (g.key == key) { hit = &g; break; }
        if (gc && !hit) {
            bool seen = false;
            for (auto &k : gc->seen) if (k == key) { seen = true; break; }
            bool failed = false;
            for (auto &k : gc->failed) if (k == key) { failed = true; break; }
            if (failed) {
            } else if (!seen) {
                // first sight of these buffers: run eagerly (this also allocates the phase planes etc. outside any capture)
                if (gc->seen.size() >= 64) gc->seen.erase(gc->seen.begin());
                gc->seen.push_back(key);
            } else {
                const long long l0 = g_md_launches.load(std::memory_order_relaxed);
                cudaGraph_t graph = nullptr;
                cudaGraphExec_t exec = nullptr;
                if (cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed) == cudaSuccess) {
                    ctx->capturing = 1;
                    const int rc = batch_enqueue(ctx, fr, out, mem, prev0, serial);
                    ctx->capturing = 0;
                    const cudaError_t ee = cudaStreamEndCapture(s, &graph);
                    if (rc == MD_OK && ee == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                        if (gc->graphs.size() >= MD_GRAPH_CACHE) {
                            size_t old = 0;
                            for (size_t i = 1; i < gc->graphs.size(); i++) if (gc->graphs[i].stamp < gc->graphs[old].stamp) old = i;
                            cudaGraphExecDestroy(gc->graphs[old].exec);
                            gc->graphs.erase(gc->graphs.begin() + old);
                        }
                        BatchGraph bg;
                        bg.key = key; bg.exec = exec; bg.launches = g_md_launches.load(std::memory_order_relaxed) - l0; bg.stamp = 0;
                        gc->graphs.push_back(bg);
                        hit = &gc->graphs.back();
                    } else {
                        (void)cudaGetLastError();     // capture failed (pageable host buffers, a foreign capture ...): stay eager
                        if (exec) cudaGraphExecDestroy(exec);
                        if (gc->failed.size() >= 64) gc->failed.erase(gc->failed.begin());
                        gc->failed.push_back(key);
                    }
                    if (graph) cudaGraphDestroy(graph);
                    // the launches counted while capturing did not run; the replay below counts them
                    g_md_launches.fetch_sub(g_md_launches.load(std::memory_order_relaxed) - l0, std::memory_order_relaxed);
                } else (void)cudaGetLastError();
            }
        }
        if (hit) {
            hit->stamp = ++gc->clock;
            CK(cudaGraphLaunch(hit->exec, s));
            MD_COUNT_LAUNCH(hit->launches);
            ctx->stats.graph_replays++;
            done = true;
        }
    }
    if (!done) {
        r = batch_enqueue(ctx, fr, out, mem, prev0, serial);
        if (r != MD_OK) return r;
    }
    if (host && mem != MD_MEM_HOST_ASYNC) CK(cudaStreamSynchronize(s));       // MD_MEM_HOST_ASYNC: the caller waits with md_sync()
    ctx->slot_base = (prev0 + pairs) % ns;
    ctx->have_cached = 1;
    ctx->win_fill = 0;               // the ring now belongs to the batch API
    ctx->pair_counter += pairs;
    ctx->stats.pairs += pairs;
    return MD_OK;
}

extern "C" int md_process_pair(md_ctx *ctx, const uint8_t *prev, const uint8_t *cur, int32_t channels, int32_t pitch,
                               const md_outputs *out, int mem)
{
    if (!ctx) return MD_ERR_INVALID;
    if (!prev || !cur) FAIL(MD_ERR_INVALID, "md_process_pair: null frame");
    md_frames fr;
    memset(&fr, 0, sizeof fr);
    fr.data = prev; fr.channels = channels; fr.pitch = pitch;
    fr.frame_stride = (int64_t)(cur - prev);
    fr.count = 2; fr.chain = 0;
    return md_process_batch(ctx, &fr, out, mem);
}

// ---- trajectories ----------------------------------------------------------------------------------------------------
// The tracking loop of calculateOpticalFlowTrajectory (cpp:161-241) over F frames whose pyramids sit in the ring slots
// slot0, slot0+1, ... (mod nslots).  d_traj [P][F], d_len [P] are device memory.
static int track_window(md_ctx *ctx, int slot0, int F, float2 *d_traj, int32_t *d_len, float *last_prev, float *last_next,
                        uint8_t *last_status, bool dev, cudaStream_t s)
{
    const int P = ctx->P, ns = ctx->g.nslots;
    float2 *cur = ctx->d_pts_in;
    CK(launch_traj_init(cur, d_traj, d_len, P, F, ctx->cfg.pixel_step, ctx->gy, s));
    const cudaMemcpyKind outk = dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    for (int j = 0; j < F - 1; j++) {
        LkParams lp;
        // the first pair starts at the grid itself (cpp:148-159): phase-plane kernel; later pairs track arbitrary points
        fill_lk(ctx, lp, (slot0 + j) % ns, (slot0 + j + 1) % ns, j == 0 ? nullptr : cur, P, ctx->d_next, ctx->d_status);
        CK(launch_lk(lp, &ctx->lk_maps, &ctx->ph_maps, 1, s));
        if (j == F - 2) {
            if (last_prev) CK(cudaMemcpyAsync(last_prev, cur, sizeof(float2) * P, outk, s));
            if (last_next) CK(cudaMemcpyAsync(last_next, ctx->d_next, sizeof(float2) * P, outk, s));
            if (last_status) CK(cudaMemcpyAsync(last_status, ctx->d_status, P, outk, s));
        }
        CK(launch_traj_step(cur, ctx->d_next, ctx->d_status, d_traj, d_len, P, F, ctx->cfg.width, ctx->cfg.height, s));
    }
    return MD_OK;
}

static int ensure_traj(md_ctx *ctx, int F)
{
    if (ctx->d_traj && ctx->traj_F >= F) return MD_OK;
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->d_traj) cudaFree(ctx->d_traj);
    if (ctx->d_traj_len) cudaFree(ctx->d_traj_len);
    ctx->d_traj = nullptr; ctx->d_traj_len = nullptr; ctx->traj_F = 0;
    CK(cudaMalloc((void **)&ctx->d_traj, sizeof(float2) * ctx->P * F));
    CK(cudaMalloc((void **)&ctx->d_traj_len, sizeof(int32_t) * ctx->P));
    ctx->traj_F = F;
    return MD_OK;
}

extern "C" int md_track_trajectories(md_ctx *ctx, const md_frames *fr, float *traj, int32_t *traj_len, float *last_prev,
                                     float *last_next, uint8_t *last_status, int mem)
{
    MD_NVTX("md_track_trajectories");
    if (!ctx) return MD_ERR_INVALID;
    if (!fr || !fr->data || !traj || !traj_len || (fr->channels != 1 && fr->channels != 3) || fr->count < 2 ||
        fr->count > ctx->g.nslots || fr->pitch < ctx->cfg.width * fr->channels || fr->chain)
        FAIL(MD_ERR_INVALID, "md_track_trajectories: need 2..max_batch+1 frames, chain=0");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const int P = ctx->P, F = fr->count;
    { int ra = ensure_traj(ctx, F); if (ra != MD_OK) return ra; }
    const uint8_t *df; int dp; long long ds;
    int r = stage_frames(ctx, fr->data, fr->channels, fr->pitch, fr->frame_stride, F, mem, &df, &dp, &ds);
    if (r != MD_OK) return r;
    CK(launch_pyramid(ctx->g, ctx->d_img, ctx->d_der, 0, F, df, fr->channels, dp, ds, s));
    ctx->slot_base = 0; ctx->have_cached = 0; ctx->win_fill = 0;
    const bool dev = mem == MD_MEM_DEVICE;
    float2 *d_traj = dev ? (float2 *)traj : ctx->d_traj;
    int32_t *d_len = dev ? traj_len : ctx->d_traj_len;
    r = track_window(ctx, 0, F, d_traj, d_len, last_prev, last_next, last_status, dev, s);
    if (r != MD_OK) return r;
    if (!dev) {
        CK(cudaMemcpyAsync(traj, d_traj, sizeof(float2) * P * F, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(traj_len, d_len, sizeof(int32_t) * P, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    }
    return MD_OK;
}

// ---- the live path (imageCallback) -------------------------------------------------------------------------------------
static int live_ensure(md_ctx *ctx, int npts, int F)
{
    if (!ctx->live_ws) ctx->live_ws = new (std::nothrow) LiveWs();
    LiveWs *ws = (LiveWs *)ctx->live_ws;
    if (!ws) FAIL(MD_ERR_NOMEM, "live path: out of host memory");
    if (ws->cap < npts || ws->capF < F) {
        CK(cudaStreamSynchronize(ctx->stream));
        if (ws->ints) cudaFree(ws->ints);
        if (ws->floats) cudaFree(ws->floats);
        if (ws->bytes) cudaFree(ws->bytes);
        ws->ints = nullptr; ws->floats = nullptr; ws->bytes = nullptr; ws->cap = 0; ws->capF = 0;
        const size_t n = (size_t)npts, nblk = (n + 2047) / 2048 + 1;
        const size_t ni = nblk + n * 2 + 8 * n + n * 3 + 4 * n + 4 * n + 2 * n + 64 + 16;
        CK(cudaMalloc((void **)&ws->ints, ni * sizeof(int)));
        CK(cudaMalloc((void **)&ws->floats, (n * F * 2 + 2 * n + n + n) * sizeof(float)));
        CK(cudaMalloc((void **)&ws->bytes, n));
        int *q = ws->ints;
        ws->cand = q; q += 8 * n;                      // 32-byte aligned rows first
        ws->blockcnt = q; q += nblk;
        ws->idx = q; q += n; ws->oidx = q; q += n; ws->ncand = q; q += n; ws->label = q; q += n; ws->sizes = q; q += n;
        ws->box = q; q += 4 * n; ws->out_box = q; q += 4 * n; ws->out_size = q; q += n; ws->out_id = q; q += n;
        ws->best = q; q += 64;
        ws->total = q++; ws->ototal = q++; ws->nclusters = q++; ws->nout = q++; ws->ninl = q++; ws->n_dev = q++;
        float *f = ws->floats;
        ws->traj_c = f; f += n * F * 2; ws->opts = f; f += 2 * n; ws->m2 = f; f += n; ws->res = f; f += n;
        ws->outl = ws->bytes;
        ws->cap = npts; ws->capF = F;
    }
    return MD_OK;
}

extern "C" int md_live_params_default(md_live_params *p)
{
    if (!p) return MD_ERR_INVALID;
    memset(p, 0, sizeof *p);
    p->num_motions = 2;              // ros/src/motion_detection_node.cpp:239
    p->sigma = 0.5;                  // node.cpp:346
    p->distance_threshold = 50.0;    // ros/launch/bag.launch:28
    p->seed = 1;
    p->subspace_iters = 50;          // outlier_detector.cpp:250
    p->min_cluster_size = 5;         // flow_clusterer.cpp:264
    return MD_OK;
}

extern "C" int md_window_reset(md_ctx *ctx)
{
    if (!ctx) return MD_ERR_INVALID;
    ctx->win_fill = 0;
    ctx->win_head = 0;
    return MD_OK;
}

extern "C" int md_window_push(md_ctx *ctx, const uint8_t *frame, int32_t channels, int32_t pitch, int32_t *fill, int mem)
{
    MD_NVTX("md_window_push");
    if (!ctx) return MD_ERR_INVALID;
    if (!frame || (channels != 1 && channels != 3) || pitch < ctx->cfg.width * channels) FAIL(MD_ERR_INVALID, "md_window_push: bad frame");
    CK(cudaSetDevice(ctx->device));
    const int ns = ctx->g.nslots;
    const int slot = ctx->win_fill ? (ctx->win_head + 1) % ns : 0;
    const uint8_t *df; int dp; long long ds;
    int r = stage_frames(ctx, frame, channels, pitch, 0, 1, mem, &df, &dp, &ds);
    if (r != MD_OK) return r;
    CK(launch_pyramid(ctx->g, ctx->d_img, ctx->d_der, slot, 1, df, channels, dp, ds, ctx->stream));
    // the pageable staging buffer is reused by the next push: host-memory pushes complete before returning
    if (mem == MD_MEM_HOST) CK(cudaStreamSynchronize(ctx->stream));
    ctx->win_head = slot;
    if (ctx->win_fill < ns) ctx->win_fill++;
    ctx->have_cached = 0;            // the ring now belongs to the window API
    if (fill) *fill = ctx->win_fill;
    return MD_OK;
}

// clusterEuclidean + boxes on ws->opts[0 .. *n_dev), results to the caller
static int live_cluster(md_ctx *ctx, LiveWs *ws, const int *n_dev, int n_max, double thr, int min_size, cudaStream_t s)
{
    CK(launch_cluster((const float2 *)ws->opts, n_dev, n_max, thr, min_size, ws->m2, ws->cand, ws->ncand, ws->label, ws->nclusters,
                      ws->sizes, ws->box, ws->nout, ws->out_box, ws->out_size, ws->out_id, s));
    return MD_OK;
}

extern "C" int md_window_detect(md_ctx *ctx, const md_live_params *lp, md_live_result *res, int mem)
{
    MD_NVTX("md_window_detect");
    if (!ctx) return MD_ERR_INVALID;
    if (!lp || !res || lp->num_motions < 1) FAIL(MD_ERR_INVALID, "md_window_detect: bad arguments");
    const int F = 2 * lp->num_motions + 1, P = ctx->P, ns = ctx->g.nslots;
    if (F > ns) FAIL(MD_ERR_INVALID, "md_window_detect: 2 * num_motions + 1 frames need max_batch >= 2 * num_motions");
    if (ctx->win_fill < F) FAIL(MD_ERR_STATE, "md_window_detect: fewer than 2 * num_motions + 1 frames pushed");
    if (2 * F > 32 || 4 * lp->num_motions > 2 * F) FAIL(MD_ERR_INVALID, "md_window_detect: num_motions too large");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int r = ensure_traj(ctx, F);
    if (r != MD_OK) return r;
    r = live_ensure(ctx, P, F);
    if (r != MD_OK) return r;
    LiveWs *ws = (LiveWs *)ctx->live_ws;
    const cudaMemcpyKind outk = mem == MD_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    res->num_trajectories = res->subspace_inliers = res->num_outliers = res->num_clusters_all = res->num_clusters = 0;

    const int slot0 = ((ctx->win_head - (F - 1)) % ns + ns) % ns;
    r = track_window(ctx, slot0, F, ctx->d_traj, ctx->d_traj_len, nullptr, nullptr, nullptr, true, s);
    if (r != MD_OK) return r;
    CK(launch_compact_trajectories(ctx->d_traj, ctx->d_traj_len, P, F, ws->blockcnt, ws->idx, ws->total, (float2 *)ws->traj_c, s));
    int T = 0;
    CK(cudaMemcpyAsync(&T, ws->total, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    res->num_trajectories = T;
    if (T < 1) return MD_OK;                                    // "no trajectories found" (node.cpp:296-321)

    r = sub_enqueue(ctx, ws->traj_c, T, F, lp->num_motions, lp->sigma, lp->seed, nullptr, lp->subspace_iters > 0 ? lp->subspace_iters : 50,
                    ws->res, ws->outl, ws->best, ws->ninl);
    if (r != MD_OK) return r;
    CK(launch_compact_outliers((const float2 *)ws->traj_c, ws->outl, T, F, ws->blockcnt, ws->oidx, ws->ototal, (float2 *)ws->opts, s));
    r = live_cluster(ctx, ws, ws->ototal, T, lp->distance_threshold, lp->min_cluster_size, s);
    if (r != MD_OK) return r;
    int counts[4] = {0, 0, 0, 0};
    CK(cudaMemcpyAsync(&counts[0], ws->ninl, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&counts[1], ws->ototal, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&counts[2], ws->nclusters, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&counts[3], ws->nout, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    res->subspace_inliers = counts[0]; res->num_outliers = counts[1]; res->num_clusters_all = counts[2]; res->num_clusters = counts[3];
    const int no = counts[1], K = counts[3];
    if (res->traj) CK(cudaMemcpyAsync(res->traj, ws->traj_c, sizeof(float2) * (size_t)T * F, outk, s));
    if (res->traj_index) CK(cudaMemcpyAsync(res->traj_index, ws->idx, sizeof(int) * T, outk, s));
    if (res->residual) CK(cudaMemcpyAsync(res->residual, ws->res, sizeof(float) * T, outk, s));
    if (res->outlier) CK(cudaMemcpyAsync(res->outlier, ws->outl, T, outk, s));
    if (res->best_cols) CK(cudaMemcpyAsync(res->best_cols, ws->best, sizeof(int) * 4 * lp->num_motions, outk, s));
    if (res->outlier_points && no) CK(cudaMemcpyAsync(res->outlier_points, ws->opts, sizeof(float2) * no, outk, s));
    if (res->labels && no) CK(cudaMemcpyAsync(res->labels, ws->label, sizeof(int) * no, outk, s));
    if (res->boxes && K) CK(cudaMemcpyAsync(res->boxes, ws->out_box, sizeof(int) * 4 * K, outk, s));
    if (res->cluster_sizes && K) CK(cudaMemcpyAsync(res->cluster_sizes, ws->out_size, sizeof(int) * K, outk, s));
    if (res->cluster_ids && K) CK(cudaMemcpyAsync(res->cluster_ids, ws->out_id, sizeof(int) * K, outk, s));
    CK(cudaStreamSynchronize(s));
    return MD_OK;
}

extern "C" int md_cluster_points(md_ctx *ctx, const float *pts, int32_t n, double distance_threshold, int32_t min_cluster_size,
                                 int32_t *labels, int32_t *num_clusters_all, int32_t *num_clusters, int32_t *boxes, int32_t *sizes,
                                 int32_t *ids, int mem)
{
    if (!ctx) return MD_ERR_INVALID;
    if (!pts || n < 0 || n > (1 << 24)) FAIL(MD_ERR_INVALID, "md_cluster_points: bad arguments");
    if (num_clusters_all) *num_clusters_all = 0;
    if (num_clusters) *num_clusters = 0;
    if (n == 0) return MD_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    int r = live_ensure(ctx, n > ctx->P ? n : ctx->P, 1);
    if (r != MD_OK) return r;
    LiveWs *ws = (LiveWs *)ctx->live_ws;
    const cudaMemcpyKind ink = mem == MD_MEM_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    const cudaMemcpyKind outk = mem == MD_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(ws->opts, pts, sizeof(float2) * n, ink, s));
    CK(cudaMemcpyAsync(ws->n_dev, &n, sizeof(int), cudaMemcpyHostToDevice, s));
    r = live_cluster(ctx, ws, ws->n_dev, n, distance_threshold, min_cluster_size, s);
    if (r != MD_OK) return r;
    int counts[2] = {0, 0};
    CK(cudaMemcpyAsync(&counts[0], ws->nclusters, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&counts[1], ws->nout, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (num_clusters_all) *num_clusters_all = counts[0];
    if (num_clusters) *num_clusters = counts[1];
    const int K = counts[1];
    if (labels) CK(cudaMemcpyAsync(labels, ws->label, sizeof(int) * n, outk, s));
    if (boxes && K) CK(cudaMemcpyAsync(boxes, ws->out_box, sizeof(int) * 4 * K, outk, s));
    if (sizes && K) CK(cudaMemcpyAsync(sizes, ws->out_size, sizeof(int) * K, outk, s));
    if (ids && K) CK(cudaMemcpyAsync(ids, ws->out_id, sizeof(int) * K, outk, s));
    CK(cudaStreamSynchronize(s));
    return MD_OK;
}

// FlowClusterer::getClusters (flow_clusterer.cpp:178-227) on the participating vectors
extern "C" int md_cluster_vectors(md_ctx *ctx, const double *vec4, int32_t n, double distance_threshold, double angular_threshold,
                                  int32_t *labels, int32_t *num_clusters_all, int mem)
{
    if (!ctx) return MD_ERR_INVALID;
    if (!vec4 || !labels || n < 0) FAIL(MD_ERR_INVALID, "md_cluster_vectors: bad arguments");
    if (n > 200000) FAIL(MD_ERR_UNSUPPORTED, "md_cluster_vectors: more than 200 000 vectors (the grouping is O(n^2) and sequential)");
    if (num_clusters_all) *num_clusters_all = 0;
    if (n == 0) return MD_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const size_t vb = sizeof(double) * 4 * (size_t)n, ab = sizeof(double) * (size_t)n, lb = (sizeof(int) * (size_t)n + 15) / 16 * 16;
    { int r = ensure_scratch(ctx, vb + ab + lb + 16); if (r != MD_OK) return r; }
    double *d_vec = (double *)ctx->scratch, *d_ang = d_vec + 4 * (size_t)n;
    int *d_label = (int *)((uint8_t *)ctx->scratch + vb + ab), *d_ncl = (int *)((uint8_t *)d_label + lb);
    const double *dv = vec4;
    if (mem == MD_MEM_HOST) { CK(cudaMemcpyAsync(d_vec, vec4, vb, cudaMemcpyHostToDevice, s)); dv = d_vec; }
    else if (((uintptr_t)vec4 & 31) != 0) { CK(cudaMemcpyAsync(d_vec, vec4, vb, cudaMemcpyDeviceToDevice, s)); dv = d_vec; }   // double4 loads
    CK(launch_cluster_vectors(dv, n, distance_threshold, angular_threshold, d_ang, d_label, d_ncl, s));
    const cudaMemcpyKind outk = mem == MD_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(labels, d_label, sizeof(int) * (size_t)n, outk, s));
    int ncl = 0;
    CK(cudaMemcpyAsync(&ncl, d_ncl, sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (num_clusters_all) *num_clusters_all = ncl;
    return MD_OK;
}

// ---- statistics --------------------------------------------------------------------------------------------------------
// ---- OpticalFlowVisualizer::showOpticalFlowVectors (k_draw.cu) ------------------------------------------------------------------
extern "C" int md_draw_flow(md_ctx *ctx, const uint8_t *image, int32_t channels, int32_t pitch, const double *vec4, int32_t n,
                            const float *next_pts, const uint8_t *status, const uint8_t *keep, const uint8_t *colour, uint8_t *out,
                            int32_t out_pitch, int32_t *drawn, int mem)
{
    MD_NVTX("md_draw_flow");
    if (!ctx) return MD_ERR_INVALID;
    const int w = ctx->cfg.width, h = ctx->cfg.height;
    const bool grid = vec4 == nullptr;
    if (!image || !out || !colour || (channels != 1 && channels != 3) || pitch < w * channels || out_pitch < w * channels ||
        (grid ? (!next_pts || !status || !keep) : n < 0))
        FAIL(MD_ERR_INVALID, "md_draw_flow: bad arguments (a Vec4d list, or next_pts + status + keep of one pair)");
    if (grid) n = ctx->P;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    const bool host = mem != MD_MEM_DEVICE;
    const size_t segb = draw_segs_bytes(n), boxb = ((size_t)n * sizeof(int4) + 255) / 256 * 256;
    const size_t row = (size_t)w * channels, imgb = (row * h + 255) / 256 * 256;
    const size_t vecb = host ? (grid ? ((size_t)n * 10 + 255) / 256 * 256 : ((size_t)n * 32 + 255) / 256 * 256) : 0;
    { int r = ensure_scratch(ctx, segb + boxb + 256 + vecb + (host ? 2 * imgb : 0)); if (r != MD_OK) return r; }
    uint8_t *base = (uint8_t *)ctx->scratch;
    void *segs = base, *abox = base + segb;
    int *d_drawn = (int *)(base + segb + boxb);
    uint8_t *d_vec = base + segb + boxb + 256, *d_in = d_vec + vecb, *d_out = d_in + imgb;
    const uint8_t *src = image;
    uint8_t *dst = out;
    int sp = pitch, dp = out_pitch;
    const double *dv = vec4;
    const float2 *dn = (const float2 *)next_pts;
    const uint8_t *dst_status = status, *dst_keep = keep;
    if (host) {
        CK(cudaMemcpy2DAsync(d_in, row, image, pitch, row, h, cudaMemcpyHostToDevice, s));
        src = d_in; dst = d_out; sp = dp = (int)row;
        if (grid) {
            CK(cudaMemcpyAsync(d_vec, next_pts, (size_t)n * 8, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_vec + (size_t)n * 8, status, n, cudaMemcpyHostToDevice, s));
            CK(cudaMemcpyAsync(d_vec + (size_t)n * 9, keep, n, cudaMemcpyHostToDevice, s));
            dn = (const float2 *)d_vec; dst_status = d_vec + (size_t)n * 8; dst_keep = d_vec + (size_t)n * 9;
        } else if (n > 0) {
            CK(cudaMemcpyAsync(d_vec, vec4, (size_t)n * 32, cudaMemcpyHostToDevice, s));
            dv = (const double *)d_vec;
        }
    }
    CK(launch_draw_flow(src, channels, sp, dst, dp, w, h, n, ctx->cfg.pixel_step, ctx->cfg.min_vector_size, dv, dn, dst_status, dst_keep,
                        ctx->gx, ctx->gy, colour, segs, abox, d_drawn, s));
    if (host) {
        CK(cudaMemcpy2DAsync(out, out_pitch, d_out, row, row, h, cudaMemcpyDeviceToHost, s));
        if (drawn) CK(cudaMemcpyAsync(drawn, d_drawn, sizeof(int), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
    } else if (drawn) CK(cudaMemcpyAsync(drawn, d_drawn, sizeof(int), cudaMemcpyDeviceToDevice, s));
    return MD_OK;
}

extern "C" int md_stats_get(md_ctx *ctx, md_stats *out)
{
    if (!ctx || !out) return MD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    unsigned long long v[136];
    CK(cudaMemcpyAsync(v, ctx->d_stats, sizeof v, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(ctx->stats.last_H, ctx->d_H, sizeof(double) * 9, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->stats.mask_pixels = (int64_t)v[0];
    ctx->stats.tracked = (int64_t)v[1];
    ctx->stats.inliers = (int64_t)v[2];
    ctx->stats.lk_iterations = 0; ctx->stats.lk_levels = 0;
    for (int i = 0; i < 64; i++) { ctx->stats.lk_iterations += (int64_t)v[8 + 2 * i]; ctx->stats.lk_levels += (int64_t)v[9 + 2 * i]; }
    ctx->stats.kernel_launches = g_md_launches.load(std::memory_order_relaxed);
    *out = ctx->stats;
    return MD_OK;
}

extern "C" int md_stats_reset(md_ctx *ctx)
{
    if (!ctx) return MD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_stats, 0, sizeof(unsigned long long) * 136, ctx->stream));
    int dev = ctx->stats.device;
    memset(&ctx->stats, 0, sizeof ctx->stats);
    ctx->stats.device = dev;
    return MD_OK;
}

extern "C" int md_set_pair_index(md_ctx *ctx, uint64_t index)
{
    if (!ctx) return MD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(launch_set_pairs(ctx->d_pair_ctr, (unsigned long long)index, ctx->stream));
    ctx->pair_counter = index;
    return MD_OK;
}

// Host utility (no context, no GPU): one table look-up turns a byte of 8 mask bits into 8 bytes of 0 / 255.
extern "C" int md_unpack_mask_host(const uint8_t *bits, int32_t bits_pitch, int32_t width, int32_t rows, uint8_t *mask, int32_t mask_pitch)
{
    if (!bits || !mask || width < 1 || rows < 0 || bits_pitch < (width + 7) / 8 || mask_pitch < width) return MD_ERR_INVALID;
    static const struct Lut {
        uint64_t v[256];
        Lut()
        {
            for (int b = 0; b < 256; b++) {
                uint64_t x = 0;
                for (int k = 0; k < 8; k++) if (b >> k & 1) x |= 0xffull << (8 * k);      // bit k -> byte k (little endian: pixel 8 i + k)
                v[b] = x;
            }
        }
    } lut;
    const int full = width >> 3, rest = width & 7;
    for (int y = 0; y < rows; y++) {
        const uint8_t *src = bits + (size_t)y * bits_pitch;
        uint8_t *dst = mask + (size_t)y * mask_pitch;
        for (int i = 0; i < full; i++) memcpy(dst + 8 * i, &lut.v[src[i]], 8);
        if (rest) memcpy(dst + 8 * full, &lut.v[src[full]], rest);
    }
    return MD_OK;
}

// ---- measurement hook ----------------------------------------------------------------------------------------------------
extern "C" int md_profile(md_ctx *ctx, int enable)
{
    if (!ctx) return MD_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (enable)
        for (int i = 0; i < 5; i++)
            if (!ctx->ev[i]) CK(cudaEventCreate(&ctx->ev[i]));
    ctx->profile = enable ? 1 : 0;
    return MD_OK;
}

extern "C" int md_profile_read(md_ctx *ctx, float *ms4)
{
    if (!ctx || !ms4) return MD_ERR_INVALID;
    if (!ctx->ev[4]) FAIL(MD_ERR_STATE, "md_profile_read: profiling was never enabled");
    CK(cudaSetDevice(ctx->device));
    CK(cudaEventSynchronize(ctx->ev[4]));
    for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(&ms4[i], ctx->ev[i], ctx->ev[i + 1]));
    return MD_OK;
}
