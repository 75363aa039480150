/*
 * motion_b200.h -- C ABI of libmotion_b200.so: the B200 (sm_100a) implementation of the motion-mask hot path of
 * shadimsaleh/motion_detection's common/ library.
 *
 * The reference is ONE executable whose hot path is a handful of public C++ methods of classes the ROS node
 * holds by value (ros/include/motion_detection/motion_detection_node.h:87-96).  Each entry point below replaces
 * the body of one of those methods (or the OpenCV routine that method calls); the C++ adapter classes in
 * adapter/ keep the reference's signatures and forward to this ABI.  Citations are path:line in the reference.
 *
 * Conventions
 *  - Every function returns an md_status (0 = OK, < 0 = error); nothing throws or aborts; nothing is written to
 *    the outputs on error.  md_last_error() gives a text for the last failure on a context.
 *  - Plain pointers and sizes only.  `mem` says where ALL data pointers of that call live: MD_MEM_HOST (pageable
 *    or pinned host memory; the call copies in/out and returns when the results are there) or MD_MEM_DEVICE
 *    (device memory of the context's GPU; the call only enqueues work on the context's stream and returns --
 *    use md_sync() or your own stream/event to wait).  Small parameter structs (md_config, H[9] inputs) are
 *    always host memory.
 *  - Images are row-major u8 with an explicit pitch in bytes.  Points are (x, y) float pairs.
 *  - A context is not thread-safe; contexts are independent (one per camera stream / GPU).
 *  - There is no CPU fallback: md_create fails with MD_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef MOTION_B200_H_
#define MOTION_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum md_status {
    MD_OK = 0,
    MD_ERR_INVALID = -1,     /* bad argument / size / NULL */
    MD_ERR_CUDA = -2,        /* CUDA runtime error (see md_last_error) */
    MD_ERR_NOMEM = -3,
    MD_ERR_UNSUPPORTED = -4,
    MD_ERR_STATE = -5        /* call order (e.g. chained batch without a cached frame) */
} md_status;

/* MD_MEM_HOST_ASYNC (md_process_batch only): host pointers like MD_MEM_HOST, but the call returns as soon as the batch -- copies
 * in, kernels, copies out -- is queued on the context's streams; md_sync() waits for it.  The frames and the outputs must stay valid
 * (and should be pinned: a pageable copy blocks the call) until then.  With two contexts fed alternately this keeps a batch queued
 * on the GPU while the host consumes the previous one: the head (first upload, first pyramids) and the tail (last egomotion fit,
 * last mask, last download) of a batch run beside the other context's flow kernels instead of on an idle GPU
 * (motion_detection_b200/streams.py: BatchPipeline). */
typedef enum md_mem { MD_MEM_HOST = 0, MD_MEM_DEVICE = 1, MD_MEM_HOST_ASYNC = 2 } md_mem;

/* Egomotion model fitted to the kept flow vectors.
 *  FIRST4: literal cv::getPerspectiveTransform(&src[0], &dst[0]) on the first four kept vectors
 *          (common/src/optical_flow_calculator.cpp:120).  Degenerate (collinear) input -> identity, inliers = 0.
 *  RANSAC_*: minimal-sample hypotheses drawn with the reference's rand() % M pattern
 *          (common/src/outlier_detector.cpp:223-234), inlier counting, first-best-wins (:300), LS refit. */
typedef enum md_ego_mode { MD_EGO_FIRST4 = 0, MD_EGO_RANSAC_HOMOGRAPHY = 1, MD_EGO_RANSAC_AFFINE = 2 } md_ego_mode;

/* Flow engine of the chain: grid pyramidal LK (calculateOpticalFlow, cpp:71) or the dense variational flow
 * (VarFlow::CalcFlow, VarFlow.cpp:600) sampled at the grid points. */
typedef enum md_flow_engine { MD_FLOW_LK = 0, MD_FLOW_VARFLOW = 1 } md_flow_engine;

typedef struct md_config {
    int32_t width, height;      /* frame size in pixels */
    int32_t max_batch;          /* max frame PAIRS per md_process_batch call (>= 1) */
    int32_t pixel_step;         /* grid step of tracked points; ROS param `pixel_step` (node.cpp:29), launch default 10 */
    double  min_vector_size;    /* ROS param `min_vector_size` (node.cpp:44); video.launch:23 uses 0.2 */
    /* cv::calcOpticalFlowPyrLK constants, optical_flow_calculator.cpp:40-44,71 */
    int32_t lk_win;             /* 40 */
    int32_t lk_max_level;       /* 5 */
    int32_t lk_max_iters;       /* 10 */
    double  lk_eps;             /* 0.03 */
    float   lk_min_eig;         /* 0.001 */
    int32_t diff_threshold;     /* 190, optical_flow_calculator.cpp:127 */
    int32_t morph;              /* 1 = erode+dilate 3x3 after the threshold (background_subtractor.cpp:31-32) */
    int32_t ego_mode;           /* md_ego_mode */
    int32_t ransac_iters;       /* 50, outlier_detector.cpp:250 */
    double  ransac_thresh;      /* reprojection threshold in px (0.5) */
    uint32_t seed;              /* srand() seed; pair i of a context uses seed + i */
    /* VarFlow parameters, optical_flow_calculator.cpp:422-429 */
    int32_t vf_max_level;       /* 4 */
    int32_t vf_start_level;     /* 0 (only 0 is supported) */
    int32_t vf_n1, vf_n2;       /* 2, 2 */
    float   vf_rho, vf_alpha, vf_sigma; /* 2.8, 1400, 1.5 */
    int32_t vf_literal;         /* 1 = literal multigrid schedule incl. the (numerically inert) coarse corrections */
    int32_t flow_engine;        /* md_flow_engine: which flow feeds the egomotion fit in md_process_batch */
    int32_t vf_grid_barrier;    /* test hook: 1 = run every Gauss-Seidel launch of md_varflow as the cooperative grid with the
                                   counting barrier (the path large levels take by themselves) instead of one cluster */
    int32_t cuda_graphs;        /* 1 (default) = md_process_batch captures the DAG of a repeated call into a CUDA graph and replays it */
    int32_t mask_packed;        /* 1 = md_process_batch writes the mask as 1 bit per pixel (bit x & 7 of byte x >> 3, LSB first, row pitch
                                   md_outputs.mask_pitch >= (width + 7) / 8 bytes): an eighth of the device-to-host bytes.  0 (default) =
                                   one byte per pixel, 0 / 255, like the reference's cv::Mat */
    int32_t reserved[4];
} md_config;

typedef struct md_ctx md_ctx;

/* Frames handed to md_process_batch. */
typedef struct md_frames {
    const uint8_t *data;        /* frame f starts at data + f * frame_stride */
    int32_t channels;           /* 1 = gray, 3 = interleaved 8UC3 (converted like cv::cvtColor(CV_BGR2GRAY), cpp:50-51) */
    int32_t pitch;              /* bytes per row */
    int64_t frame_stride;       /* bytes between frames */
    int32_t count;              /* number of frames */
    int32_t chain;              /* 1 = pair 0 is (last frame of the previous call, frame 0): count pairs; 0 = count-1 pairs */
} md_frames;

/* Per-pair outputs of md_process_batch; any pointer may be NULL (that output is skipped). P = md_grid_size(). */
typedef struct md_outputs {
    float   *next_pts;          /* [pairs][P][2] tracked position in the second frame */
    uint8_t *status;            /* [pairs][P]    LK status (cpp:80) */
    uint8_t *keep;              /* [pairs][P]    1 = vector entered src/dst (cpp:86-97) */
    double  *H;                 /* [pairs][9]    egomotion, maps frame k -> frame k+1, H[8] = 1 */
    int32_t *num_vectors;       /* [pairs]       return value of calculateOpticalFlow (cpp:129) */
    int32_t *inliers;           /* [pairs]       inliers of the winning hypothesis (0 = no egomotion, mask is all 0) */
    uint8_t *mask;              /* [pairs] images, 0 / 255: the motion mask `comp` (cpp:125-127) after morphology */
    int32_t mask_pitch;
    int64_t mask_stride;
} md_outputs;

typedef struct md_stats {
    int64_t pairs;              /* frame pairs processed */
    int64_t mask_pixels;        /* sum of set mask pixels over those pairs */
    int64_t tracked;            /* sum of LK status==1 points */
    int64_t inliers;            /* sum of egomotion inliers */
    double  last_H[9];
    int64_t kernel_launches;    /* CUDA kernels launched by this library in this process (all contexts) */
    int32_t device;
    int32_t reserved0;
    int64_t lk_iterations;      /* sum over tracked points and pyramid levels of the LK iterations executed (each = 1600 taps) */
    int64_t lk_levels;          /* sum over tracked points of the pyramid levels whose window was evaluated */
    int64_t graph_replays;      /* md_process_batch calls served by a captured CUDA graph */
} md_stats;

/* ---- lifetime -------------------------------------------------------------------------------------------- */
int md_config_default(md_config *cfg);                           /* fills the reference's constants */
int md_create(const md_config *cfg, int device, md_ctx **out);   /* replaces construction of ofc_/od_/bs_ (node.h:87-96) */
int md_destroy(md_ctx *ctx);
int md_set_stream(md_ctx *ctx, void *cuda_stream);               /* cudaStream_t; NULL = the context's own stream */
int md_sync(md_ctx *ctx);
const char *md_last_error(const md_ctx *ctx);
const char *md_version(void);

/* ---- geometry of the tracked grid / pyramid ----------------------------------------------------------------- */
int md_grid_size(const md_ctx *ctx);                             /* P = ceil(w/ps) * ceil(h/ps), x outer / y inner (cpp:56-64) */
int md_grid_points(const md_ctx *ctx, float *pts /* [P][2] host */);
int md_pyramid_levels(const md_ctx *ctx);                        /* levels buildOpticalFlowPyramid keeps (max level + 1) */
int md_pyramid_level_size(const md_ctx *ctx, int level, int32_t *w, int32_t *h);

/* ---- K0: cv::cvtColor(CV_BGR2GRAY), optical_flow_calculator.cpp:50-51,166-167 ------------------------------ */
int md_gray_u8(md_ctx *ctx, const uint8_t *src3, int32_t src_pitch, int32_t w, int32_t h, uint8_t *dst, int32_t dst_pitch, int mem);

/* ---- K1: cv::buildOpticalFlowPyramid(gray, pyr, Size(40,40), 5, true), cpp:67,170 --------------------------- */
/* Builds the u8 pyramid and the Scharr derivative planes of one gray frame into pyramid slot `slot`
 * (0 <= slot <= max_batch). */
int md_pyramid_u8(md_ctx *ctx, const uint8_t *gray, int32_t pitch, int slot, int mem);
int md_pyramid_read(md_ctx *ctx, int slot, int level, uint8_t *dst, int32_t dst_pitch, int mem);        /* parity probes */
int md_pyramid_read_deriv(md_ctx *ctx, int slot, int level, int16_t *dst /* [h][w][2] */, int mem);

/* ---- K2: cv::calcOpticalFlowPyrLK(pyr, gray2, pts1, pts2, status, err, win, 5, termcrit, 0, 0.001), cpp:71,172 */
/* pts_in == NULL tracks the context's grid (npts must then be md_grid_size). */
int md_lk_flow(md_ctx *ctx, int slot_prev, int slot_next, const float *pts_in, int32_t npts, float *pts_out,
               uint8_t *status, int mem);

/* ---- K3: vector filter (cpp:78-117) + egomotion fit (cpp:120 / SURVEY 8c "intended") ------------------------- */
/* keep == NULL: computed from status and min_vector_size.  H9 / inliers / inlier_mask may be NULL. */
int md_fit_egomotion(md_ctx *ctx, const float *src, const float *dst, const uint8_t *status, const uint8_t *keep,
                     int32_t npts, int mode, uint32_t seed, double *H9, int32_t *num_vectors, int32_t *inliers,
                     uint8_t *inlier_mask, int mem);

/* ---- K4: cv::warpPerspective + cv::absdiff + cv::threshold (cpp:124-127) + erode/dilate (bgsub.cpp:31-32) ----- */
/* H9 is host memory (maps prev -> cur, inverted internally like cv::warpPerspective). */
int md_motion_mask(md_ctx *ctx, const uint8_t *prev, const uint8_t *cur, int32_t pitch, const double *H9,
                   int32_t thresh, int32_t morph, uint8_t *mask, int32_t mask_pitch, int mem);

/* ---- the chain: OpticalFlowCalculator::calculateOpticalFlow (cpp:30-130) + morphology, batched ---------------- */
int md_process_batch(md_ctx *ctx, const md_frames *frames, const md_outputs *out, int mem);
/* Convenience for one pair (prev, cur), host or device. */
int md_process_pair(md_ctx *ctx, const uint8_t *prev, const uint8_t *cur, int32_t channels, int32_t pitch,
                    const md_outputs *out, int mem);

/* ---- OpticalFlowCalculator::calculateOpticalFlowTrajectory (cpp:133-257) ------------------------------------- */
/* Tracks the grid through `frames->count` frames; traj gets [P][F][2] positions, traj_len[P] the number of valid
 * entries (a trajectory is complete when traj_len == F, cpp:246); last_next/last_status are the LK outputs of the
 * last pair (cpp:183-206). */
int md_track_trajectories(md_ctx *ctx, const md_frames *frames, float *traj, int32_t *traj_len, float *last_prev,
                          float *last_next, uint8_t *last_status, int mem);

/* ---- the node's LIVE path: MotionDetectionNode::imageCallback (ros/src/motion_detection_node.cpp:235-430) --------- */
/* The node keeps the last F = 2 * num_motions + 1 frames in a deque (node.cpp:248-261), re-tracks the grid through them on
 * every callback (runOpticalFlowTrajectory -> calculateOpticalFlowTrajectory, cpp:133-257), fits the motion subspace
 * (od_.fitSubspace, node.cpp:348), groups the outlier points (fc_.clusterEuclidean, node.cpp:355 /
 * common/src/flow_clusterer.cpp:227-269) and boxes every group of more than 5 points (ofv_.showBoundingBoxes, node.cpp:395 /
 * common/src/optical_flow_visualizer.cpp:223-240); ml_.writeBoundingBox logs "frame, id, x0, y0, x1, y1"
 * (common/src/motion_logger.cpp:43-47).  Here the deque is a ring of pyramid slots: every frame is converted and its
 * pyramid built ONCE (the reference rebuilds F-1 pyramids per callback), the rest runs on the device. */
typedef struct md_live_params {
    int32_t num_motions;        /* ROS param num_motions (node.cpp:239), F = 2 * num_motions + 1 frames */
    double  sigma;              /* ROS param sigma (node.cpp:346), default 0.5 */
    double  distance_threshold; /* ROS param distance_threshold (node.cpp:317), launch files: 50 */
    uint32_t seed;              /* srand() seed of fitSubspace's rand() % T draws */
    int32_t subspace_iters;     /* 50 (outlier_detector.cpp:250) */
    int32_t min_cluster_size;   /* clusters with MORE than this many points are reported: 5 (flow_clusterer.cpp:264) */
    int32_t reserved[4];
} md_live_params;

typedef struct md_live_result {
    /* counts: always written (host memory) */
    int32_t num_trajectories;   /* complete trajectories T (cpp:244-254); 0 = "no trajectories found" (node.cpp:298) */
    int32_t subspace_inliers;   /* inliers of the winning subspace hypothesis */
    int32_t num_outliers;       /* outlier_points.size() (outlier_detector.cpp:318-324) */
    int32_t num_clusters_all;   /* clusters founded by clusterEuclidean, any size */
    int32_t num_clusters;       /* clusters with more than min_cluster_size points = rectangles */
    int32_t reserved[3];
    /* arrays: each may be NULL; capacity md_grid_size() entries; memory space = `mem` of the call */
    float   *traj;              /* [T][F][2] complete trajectories in grid order */
    int32_t *traj_index;        /* [T] grid index of each */
    float   *residual;          /* [T] */
    uint8_t *outlier;           /* [T] */
    int32_t *best_cols;         /* [4 * num_motions] trajectories spanning the winning subspace (return value of fitSubspace) */
    float   *outlier_points;    /* [num_outliers][2] */
    int32_t *labels;            /* [num_outliers] cluster id (creation order, any size) of every outlier point */
    int32_t *boxes;             /* [num_clusters][4] tl.x, tl.y, br.x, br.y of cv::boundingRect (the CSV columns) */
    int32_t *cluster_sizes;     /* [num_clusters] */
    int32_t *cluster_ids;       /* [num_clusters] id (as in labels) of every reported cluster */
} md_live_result;

int md_live_params_default(md_live_params *p);
/* Forgets the frames pushed so far. */
int md_window_reset(md_ctx *ctx);
/* raw_images_.push_back(image) (+ pop_front when full): converts (channels 3: cv::cvtColor(CV_BGR2GRAY) on the rgb8 data,
 * cpp:166-167) and builds the pyramid of ONE frame into the next ring slot.  *fill = frames held (<= max_batch + 1). */
int md_window_push(md_ctx *ctx, const uint8_t *frame, int32_t channels, int32_t pitch, int32_t *fill, int mem);
/* Runs the callback body over the newest F frames (needs F <= frames held).  Returns after the results are complete. */
int md_window_detect(md_ctx *ctx, const md_live_params *params, md_live_result *res, int mem);

/* ---- FlowClusterer::clusterEuclidean + showBoundingBoxes on a caller-supplied point list ----------------------------- */
/* pts [n][2]; labels [n]; boxes [n][4] / sizes [n] / ids [n] capacity n; counts are host memory. */
int md_cluster_points(md_ctx *ctx, const float *pts, int32_t n, double distance_threshold, int32_t min_cluster_size,
                      int32_t *labels, int32_t *num_clusters_all, int32_t *num_clusters, int32_t *boxes, int32_t *sizes,
                      int32_t *ids, int mem);

/* ---- FlowClusterer::getClusters (common/src/flow_clusterer.cpp:178-227; node.cpp:127,169,375) --------------------------------- */
/* vec4 [n][4] f64 = (x, y, dx, dy) of the participating flow vectors in the reference's traversal order (rows outer, columns
 * inner, steps of pixel_step, only |dx| > 0 or |dy| > 0, :180-185).  labels[n] = id (creation order) of the cluster every
 * vector joins: the FIRST cluster with a member nearer than distance_threshold and a member whose orientation differs by
 * less than angular_threshold (VectorCluster::getClosestDistance / getClosestOrientation, vector_cluster.cpp:25-50).
 * The reference returns the clusters with more than 5 members (:210-217); sizes follow from the labels. */
int md_cluster_vectors(md_ctx *ctx, const double *vec4, int32_t n, double distance_threshold, double angular_threshold,
                       int32_t *labels, int32_t *num_clusters_all, int mem);

/* ---- OpticalFlowVisualizer::showOpticalFlowVectors (common/src/optical_flow_visualizer.cpp:23-71; node.cpp:83,101) --------------- */
/* out = image (1 or 3 interleaved channels, width x height of the context) with one anti-aliased arrow per flow vector that passes
 * the reference's test (|dx| or |dy| > min_vector_size, both < 5 * pixel_step; :37), drawn like cv::line(..., colour, 1, CV_AA):
 * shaft start -> end, two 3-pixel head strokes at +-45 degrees, in the reference's visiting order (overlapping arrows blend in
 * that order).  The vectors are EITHER vec4 [n][4] f64 = (x, y, dx, dy), the non-empty elements of the flow field row by row
 * (rows outer, columns inner), OR -- vec4 == NULL -- one pair's next_pts [P][2] / status [P] / keep [P] as md_process_batch wrote
 * them (the field of optical_flow_calculator.cpp:78-117 is formed on the device; nothing has to visit the host to be drawn).
 * pixel_step and min_vector_size are the context's.  colour[channels].  *drawn (nullable, same memory space as the buffers) =
 * arrows that passed the test.  MD_MEM_DEVICE: stream ordered, image / out may not alias. */
int md_draw_flow(md_ctx *ctx, const uint8_t *image, int32_t channels, int32_t pitch, const double *vec4, int32_t n,
                 const float *next_pts, const uint8_t *status, const uint8_t *keep, const uint8_t *colour, uint8_t *out,
                 int32_t out_pitch, int32_t *drawn, int mem);

/* ---- OutlierDetector::fitSubspace (common/src/outlier_detector.cpp:236-331) ------------------------------------ */
/* traj [T][F][2]; forced_cols NULL (rand() % T after srand(seed)) or [iters][4*num_motions] host indices.
 * residual[T] f32, best_cols[4*num_motions] i32, outlier[T] u8 (residual > threshold, :318-324). */
int md_fit_subspace(md_ctx *ctx, const float *traj, int32_t T, int32_t F, int32_t num_motions, double sigma,
                    uint32_t seed, const int32_t *forced_cols, int32_t iters, float *residual, int32_t *best_cols,
                    uint8_t *outlier, int32_t *num_inliers, int mem);

/* ---- OutlierDetector::findOutliers / createMask (common/src/outlier_detector.cpp:37-186) ---------------------------- */
/* The median / MAD test on the angles atan2(dy, dx) and magnitudes sqrt(dy^2 + dx^2) of the grid's flow vectors.
 * flow_dxdy [n][2] f64 = components 2, 3 of the Vec4d field at the grid points (zero for filtered / failed vectors,
 * cpp:99-115); include_zeros = 0 leaves zero vectors out of the statistics (:132-139).  outlier[n] = 1 where
 * 0.6745 |v - median| / MAD > 3.5 for the angle or the magnitude (= outlier_probabilities > 0.5, :54-70).
 * stats4 (nullable): median angle, MAD angle, median magnitude, MAD magnitude. */
int md_find_outliers(md_ctx *ctx, const double *flow_dxdy, int32_t n, int32_t include_zeros, uint8_t *outlier,
                     double *stats4, int mem);

/* ---- VarFlow::CalcFlow (common/src/VarFlow.cpp:600-697) with the varFlow() parameters (cpp:422-429) ----------- */
/* A, B gray u8; U (+x) and V (y-UP, VarFlow.cpp:103-107) f32 [h][w] dense. */
int md_varflow(md_ctx *ctx, const uint8_t *A, const uint8_t *B, int32_t pitch, float *U, float *V, int mem);

/* ---- per-stream statistics (what NCCL gathers at report time) --------------------------------------------------- */
int md_stats_get(md_ctx *ctx, md_stats *out);
int md_stats_reset(md_ctx *ctx);
/* The context's count of pairs processed so far, which numbers the RANSAC seeds (pair i draws after srand(seed + i),
 * outlier_detector.cpp:223-234 pattern).  Setting it (stream ordered) lets several contexts share one camera sequence -- batch k on
 * context k % n -- and draw exactly the hypotheses a single context would. */
int md_set_pair_index(md_ctx *ctx, uint64_t index);

/* ---- host utility: a mask_packed = 1 mask (1 bit per pixel, LSB first) expanded to the reference's cv::Mat form, 0 / 255 bytes ----- */
/* Pure host code (no context, no GPU): `rows` rows of `width` pixels; bits_pitch >= (width + 7) / 8, mask_pitch >= width.  For users who
 * take the packed mask over a saturated host link and still need `comp` as optical_flow_calculator.cpp:127 leaves it. */
int md_unpack_mask_host(const uint8_t *bits, int32_t bits_pitch, int32_t width, int32_t rows, uint8_t *mask, int32_t mask_pitch);

/* ---- measurement hook: CUDA events around the four stages of md_process_batch ---------------------------------- */
/* enable != 0: the next md_process_batch calls record events on the context's stream around
 * K1 (pyramid), K2 (LK), K3 (egomotion), K4 (mask).  md_profile_read waits for the last batch and returns the
 * four stage durations in milliseconds (of that last batch). */
int md_profile(md_ctx *ctx, int enable);
int md_profile_read(md_ctx *ctx, float *ms4);

#ifdef __cplusplus
}
#endif
#endif /* MOTION_B200_H_ */
