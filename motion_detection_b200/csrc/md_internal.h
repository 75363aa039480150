// md_internal.h -- context layout and kernel launch prototypes shared by the .cu files of libmotion_b200.so.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/motion_b200.h"

#define MD_MAX_LEVELS 8
#define MD_MAX_HYP 256

// Pyramid planes are stored PADDED, exactly like cv::buildOpticalFlowPyramid does (winSize border):
// image planes carry a BORDER_REFLECT_101 frame, derivative planes a zero frame, so the LK window reads
// (which reach from -win to w+win-1) never need border logic.
struct LevelGeom {
    int w, h;          // interior size
    int pitch;         // elements per padded row (u8 for images, short2 for derivatives); multiple of 128
    int rows;          // h + 2*pady
    size_t img_off;    // byte offset of the padded plane inside a slot's image arena
    size_t der_off;    // short2 offset of the padded plane inside a slot's derivative arena
};

struct PyrGeom {
    int nlev;
    int padx, pady;
    LevelGeom lv[MD_MAX_LEVELS];
    size_t slot_img_bytes;   // bytes per slot (all levels)
    size_t slot_der_elems;   // short2 per slot
    int nslots;
};

// Shared-memory tile geometry of k_lk_tma (boxes of the per-level TMA tensor maps), window = 40.
#define MD_LK_I_BOX_W 64        // bytes: <= 15 alignment offset + 41 window columns + the over-read word, multiple of 16
#define MD_LK_J_BOX_W 80        // bytes: window + drift margin
#define MD_LK_D_BOX_W 44        // short2 elements (176 bytes)
#define MD_LK_J_MARGIN_X 8
#define MD_LK_J_MARGIN_Y 3

// Per-level tensor maps over the padded pyramid planes: dims (x, y, slot).  128 bytes each, passed as a
// __grid_constant__ kernel parameter.
struct alignas(64) LkTmaMaps {
    CUtensorMap imgI[MD_MAX_LEVELS];   // u8, box 64 x 41
    CUtensorMap imgJ[MD_MAX_LEVELS];   // u8, box 80 x 47
    CUtensorMap der[MD_MAX_LEVELS];    // u32 (short2), box 44 x 41
    int valid;
};

// ---- sub-pixel phase planes (k_lk_phase.cu) -------------------------------------------------------------------------
// Grid points sit at integer multiples of pixel_step, so at pyramid level l the fractional part of a point's window
// origin takes only (2^l / gcd(ps, 2^l))^2 distinct values ("phase classes").  For every class the bilinearly
// interpolated window samples (I with 5 extra bits, Ix, Iy: the values LKTrackerInvoker builds per point) are
// computed ONCE per pixel into s16 planes; a point's 40x40 window is then a plain TMA box of its class's planes.
// Plane pixel (X, Y) holds the samples of the window tap whose integer origin is (X - margin, Y - margin).
#define MD_PH_MARGIN 24         // window origins of in-image grid points reach from -20 to w - 20
#define MD_PH_BOX_W 48          // s16 elements: <= 7 alignment offset + 40 window columns, multiple of 8 (16 bytes)
struct PhaseLevel {
    int w, h;            // plane size: level size + 2 * margin
    int pitch;           // s16 elements per row, multiple of 8
    int ncx;             // classes per axis
    int shift;           // log2 gcd(ps, 2^level): class = (coordinate & (2^level - 1)) >> shift
    size_t off;          // s16 offset of (class 0, plane 0) inside a pair's arena; layout [class][I, Ix, Iy][row][pitch]
};
struct PhaseGeom {
    PhaseLevel lv[MD_MAX_LEVELS];
    size_t pair_elems;   // s16 per pair (all levels, classes, planes)
    size_t px_per_pair;  // plane pixels per pair (work estimate)
};
struct alignas(64) LkPhaseMaps {
    CUtensorMap imgJ[MD_MAX_LEVELS];   // u8, box 80 x 47 (same as LkTmaMaps::imgJ)
    CUtensorMap ph[MD_MAX_LEVELS];     // s16, dims (x, y, plane, class, pair), box 48 x 40 x 2 (the Ix, Iy planes)
    int valid;
};

// Per (pair, level, grid point) scalars of the LK normal equations, written by k_window_sums and read by k_lk_phase: the f32 matrix
// entries (window sums * 2^-20), the reciprocal determinant (0 = the level fails calcOpticalFlowPyrLK's minimum-eigenvalue / determinant
// test and is skipped) and the exact integer sums of I * Ix, I * Iy.  32 bytes, two 16-byte loads.
struct alignas(16) LkLevelRec { float A11, A12, A22, Dinv; long long C1, C2; };

struct LkParams {
    PyrGeom g;
    const uint8_t *img;      // slot 0 image arena
    const short2 *der;       // slot 0 derivative arena
    int prev_slot0, next_slot0;   // pair b uses slots (prev_slot0 + b) % nslots, (next_slot0 + b) % nslots
    const float2 *pts_in;    // [pairs][P] or nullptr (grid)
    int ps, gy;              // grid: x = ps * (k / gy), y = ps * (k % gy)
    int P;
    float2 *next;            // [pairs][P]
    uint8_t *status;         // [pairs][P]
    int win, max_iters;
    double eps2;
    float min_eig;
    // grid mode with phase planes (pts_in == nullptr): pair b uses arena ph + b * pg.pair_elems
    PhaseGeom pg;
    int16_t *ph;
    LkLevelRec *wsum;        // [max_batch][nlev][P] per-point level scalars (k_window_sums)
    int ph_pair0;            // pair b of this launch uses arena ph_pair0 + b
    int ph_ready;            // 1 = the planes were already computed (launch_lk_planes on another stream)
    unsigned long long *stat_iters;   // [64][2] nullable, striped: iterations executed, levels iterated (summed over points)
};

struct EgoParams {
    int P, ps, gy;
    const float2 *pts_in;    // nullptr = grid
    const float2 *next;      // [pairs][P]
    const uint8_t *status;   // [pairs][P]
    uint8_t *keep;           // [pairs][P]
    int keep_given;          // 1: keep[] is an input
    double min_vec;
    int mode, iters, minimal;
    double thr2;
    uint32_t seed0;          // pair b uses seed0 + b (+ *pair_ctr when given: the context's count of pairs processed so far, on the device
                             // so that a captured graph of the batch replays with fresh seeds)
    const unsigned long long *pair_ctr;
    int w, h;
    int nblk_scan;           // blocks per pair for keep/compact (2048 items each)
    int nblk_acc;            // blocks per pair for the normal-equation accumulation
    int *blockcnt;           // [pairs][nblk_scan]
    int *kept_idx;           // [pairs][P]
    int *M;                  // [pairs]
    double *hyp;             // [pairs][iters][9]
    int *hyp_valid;          // [pairs][iters]
    int *counts;             // [pairs][iters]
    double *partial;         // [pairs][nblk_acc][48]
    double *H, *Hinv;        // [pairs][9]
    int *inliers;            // [pairs]
    int *valid;              // [pairs] 1 = mask may be computed
    uint8_t *inlier_mask;    // [pairs][P] or nullptr
    unsigned long long *stat_tracked, *stat_inliers;
};

// K4 TMA boxes: the source bounding box of a 128 x 40 computed region in `prev` (16-byte aligned start: up to 15 columns of slack)
#define MD_MASK_PREV_BOX_W 256
#define MD_MASK_PREV_BOX_H 48
struct alignas(64) MaskTmaMaps {
    CUtensorMap prev, cur;   // u8, dims (x, y, frame) over the image INTERIOR: out-of-image elements are zero-filled
    int valid;
};

struct MaskParams {
    // frame b of prev lives at prev + ((prev_slot0 + b) % nslots) * stride when nslots > 0 (pyramid ring),
    // at prev + b * stride when nslots == 0 (plain frame arrays); same for cur.
    const uint8_t *prev, *cur;
    int pitch;
    long long stride;
    int nslots, prev_slot0, cur_slot0;
    int w, h;
    const double *Hinv;      // [pairs][9] device
    const int *valid;        // [pairs] device (nullptr = always valid)
    int thresh, morph;
    uint8_t *mask;
    int mask_pitch;
    long long mask_stride;
    unsigned long long *stat_mask;   // nullable
    int packed;              // 1 = the mask leaves as 1 bit per pixel (LSB first), mask_pitch in bytes of packed rows
    int use_tma;             // set by launch_mask: 1 = the tensor maps are usable (TMA-staged fast path)
};

struct md_ctx {
    md_config cfg;
    int device;
    cudaStream_t own_stream, stream;
    cudaStream_t copy_in, copy_out;     // H2D / D2H streams of the pipelined host-memory path
    cudaEvent_t ev_in[8], ev_comp[8];
    cudaStream_t aux_pyr, aux_post;     // side streams of the kernel pipeline (K1 / K3+K4 beside LK), higher priority
    cudaEvent_t ev_k1[8], ev_lk[8], ev_fork, ev_join, ev_out;
    int capturing;                      // batch_enqueue runs under stream capture
    cudaStream_t aux_lv[2];             // level groups of the phase planes / window sums (launch_lk_planes)
    cudaEvent_t ev_lv[4];
    std::string err;
    int sm_count;

    PyrGeom g;
    uint8_t *d_img;
    short2 *d_der;
    int P, gx, gy;
    int slot_base;        // slot holding frame 0 of the current batch
    int have_cached;      // slot_base holds a valid pyramid of the last frame of the previous batch
    uint64_t pair_counter;
    unsigned long long *d_pair_ctr;      // device copy of pair_counter (k_hypotheses reads it, k_advance_pairs bumps it)
    void *graphs;                        // cache of captured batch graphs (md_api.cu)

    // staging for host-memory calls / plain frames
    uint8_t *d_frames;    // [(max_batch+1)][h][fpitch*channels], allocated on first host-memory call
    int frames_channels;
    int fpitch;
    uint8_t *d_mask;      // [max_batch][h][fpitch]
    float2 *d_pts_in;     // [max_batch][P]
    float2 *d_next;
    uint8_t *d_status, *d_keep, *d_inlier_mask;
    size_t pts_cap;       // capacity (points per pair) of the per-point buffers

    // egomotion workspace
    int nblk_scan, nblk_acc;
    int *d_blockcnt, *d_kept_idx, *d_M, *d_hyp_valid, *d_counts, *d_inliers, *d_valid;
    double *d_hyp, *d_partial, *d_H, *d_Hinv;
    unsigned long long *d_stats;   // [0] mask px, [1] tracked, [2] inliers, [3] spare, [8 .. 136) 64 striped (LK iterations, LK levels) pairs
    float2 *d_traj;       // [P][F] trajectory staging (host-memory calls)
    int32_t *d_traj_len;
    int traj_F;
    int win_head, win_fill;   // live-path ring: slot of the newest frame, frames held
    md_stats stats;
    int profile;
    cudaEvent_t ev[5];
    LkTmaMaps lk_maps;
    LkPhaseMaps ph_maps;
    MaskTmaMaps mask_maps;        // level-0 planes of the pyramid ring (frames of md_process_batch)
    MaskTmaMaps mask_maps_user;   // plain frame arrays of md_motion_mask, re-encoded when the buffers change
    const void *mask_user_key[2];
    int mask_user_pitch;
    PhaseGeom pg;
    int16_t *d_phase;     // [max_batch] pair arenas of phase planes, allocated on the first grid-mode LK call
    LkLevelRec *d_wsum;   // [max_batch][nlev][P] per-point level scalars
    int phase_state;      // 0 = not tried, 1 = ready, -1 = not used (not worth it / allocation failed)
    void *sub_ws;         // fitSubspace workspace (k_subspace.cu)
    void *mad_ws;         // findOutliers workspace (k_mad.cu)
    void *live_ws;        // live-path workspace (md_api.cu: md_window_*)
    // VarFlow workspaces (k_varflow.cu), allocated on first use.  Slot 0 serves md_varflow; a batch on the variational engine runs
    // up to MD_VF_LANES pairs at once, each on its own workspace and stream (a pair's Gauss-Seidel wavefront keeps one 16-CTA
    // cluster busy, a ninth of the GPU)
    void *vf_ws[8];
    cudaStream_t vf_stream[8];
    cudaEvent_t vf_ev[8], vf_fork;
    void *scratch;        // grow-only device scratch of the small host-memory entry points (md_gray_u8, md_cluster_vectors)
    size_t scratch_bytes;
    long long launches;   // kernels launched through this context's calls (md_stats.kernel_launches)
    int pipe_bounds[8], pipe_nbounds;           // MD_PIPE_BOUNDS
    int pipe_chunks, pipe_first, trace_calls;   // MD_PIPE_CHUNKS / MD_PIPE_FIRST / MD_TRACE, read once in md_create
};

void vf_free_workspace(void *ws);
void sub_free_workspace(void *ws);
void mad_free_workspace(void *ws);
int sub_enqueue(md_ctx *ctx, const float *d_traj, int T, int F, int num_motions, double sigma, uint32_t seed,
                const int32_t *forced_host, int iters, float *d_res, uint8_t *d_out, int *d_best, int *d_ninl);
#define MD_VF_LANES 8
int vf_compute_device(md_ctx *ctx, int lane, cudaStream_t s, const uint8_t *dA, const uint8_t *dB, int dpitch);
cudaError_t vf_sample_grid(md_ctx *ctx, int lane, float2 *next, uint8_t *status, cudaStream_t s);

// kernels launched by this library, process wide (md_stats.kernel_launches): contexts may live on different threads
extern std::atomic<long long> g_md_launches;
// NVTX ranges (tracing): one host-side range per stage of the chain and per entry point; free when no tool is attached
#include <nvtx3/nvToolsExt.h>
struct MdNvtxRange {
    explicit MdNvtxRange(const char *name) { nvtxRangePushA(name); }
    ~MdNvtxRange() { nvtxRangePop(); }
};
#define MD_NVTX_CAT2(a, b) a##b
#define MD_NVTX_CAT(a, b) MD_NVTX_CAT2(a, b)
#define MD_NVTX(name) MdNvtxRange MD_NVTX_CAT(md_nvtx_range_, __LINE__)(name)

#define MD_COUNT_LAUNCH(n) (g_md_launches.fetch_add((n), std::memory_order_relaxed))

// kernel launchers (each in its own .cu)
cudaError_t launch_gray(const uint8_t *src3, int src_pitch, int w, int h, uint8_t *dst, int dst_pitch, cudaStream_t s);
cudaError_t launch_pyramid(const PyrGeom &g, uint8_t *img, short2 *der, int slot0, int nframes, const uint8_t *frames,
                           int channels, int fpitch, long long fstride, cudaStream_t s);
cudaError_t launch_lk(const LkParams &p, const LkTmaMaps *maps, const LkPhaseMaps *pmaps, int pairs, cudaStream_t s);
cudaError_t launch_ego(const EgoParams &p, int pairs, cudaStream_t s);
cudaError_t launch_draw_flow(const uint8_t *src, int channels, int spitch, uint8_t *dst, int dpitch, int w, int h, int n, int ps, double min_vec,
                             const double *vec4, const float2 *next, const uint8_t *status, const uint8_t *keep, int gx, int gy,
                             const uint8_t *colour, void *segs, void *abox, int *drawn, cudaStream_t s);
size_t draw_segs_bytes(int n);
cudaError_t launch_advance_pairs(unsigned long long *ctr, int pairs, cudaStream_t s);
cudaError_t launch_set_pairs(unsigned long long *ctr, unsigned long long v, cudaStream_t s);
cudaError_t launch_mask(const MaskParams &p, int pairs, const MaskTmaMaps *maps, cudaStream_t s);
bool mask_encode_maps(MaskTmaMaps *m, const uint8_t *prev, const uint8_t *cur, int w, int h, int pitch, long long stride, int nframes);
cudaError_t launch_compact_trajectories(const float2 *traj, const int32_t *len, int P, int F, int *blockcnt, int *idx, int *total,
                                        float2 *traj_c, cudaStream_t s);
cudaError_t launch_compact_outliers(const float2 *traj_c, const uint8_t *outlier, int T, int F, int *blockcnt, int *oidx, int *total,
                                    float2 *pts, cudaStream_t s);
cudaError_t launch_cluster(const float2 *pts, const int *n_ptr, int n_max, double thr, int min_size, float *m2, int *cand, int *ncand,
                           int *label, int *nclusters, int *sizes, int *box, int *nout, int32_t *out_box, int32_t *out_size,
                           int32_t *out_id, cudaStream_t s);
cudaError_t launch_cluster_vectors(const double *vec4, int n, double dthr, double athr, double *ang, int *label, int *nclusters,
                                   cudaStream_t s);
cudaError_t launch_lk_planes(const LkParams &p, int pairs, cudaStream_t s);
cudaError_t launch_lk_planes_levels(const LkParams &p, int pairs, int l0, int l1, cudaStream_t s);
cudaError_t launch_pyramid_level0(const PyrGeom &g, uint8_t *img, int slot0, int nframes, const uint8_t *frames, int channels,
                                  int fpitch, long long fstride, cudaStream_t s);
cudaError_t launch_pyramid_down(const PyrGeom &g, uint8_t *img, int slot0, int nframes, int l0, int l1, cudaStream_t s);
cudaError_t launch_pyramid_scharr(const PyrGeom &g, uint8_t *img, short2 *der, int slot0, int nframes, int l0, int l1, cudaStream_t s);
cudaError_t launch_lk_phase(const LkParams &p, const LkPhaseMaps *maps, int pairs, cudaStream_t s);
cudaError_t launch_traj_step(float2 *pts_cur, const float2 *next, const uint8_t *status, float2 *traj, int32_t *len,
                             int P, int F, int w, int h, cudaStream_t s);
cudaError_t launch_traj_init(float2 *pts_cur, float2 *traj, int32_t *len, int P, int F, int ps, int gy, cudaStream_t s);
