/* md_oracle.h -- CPU ORACLE (test infrastructure only; see md_oracle.c header). */
#ifndef MD_ORACLE_H_
#define MD_ORACLE_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int32_t r[34]; int f, b; } orc_rand;
void orc_glibc_srand(orc_rand *st, uint32_t seed);
int orc_glibc_rand(orc_rand *st);

void orc_gray_bgr2gray(const uint8_t *src, int w, int h, int pitch, uint8_t *dst, int dpitch);
void orc_pyr_down(const uint8_t *src, int w, int h, int pitch, uint8_t *dst, int dpitch);
int orc_pyr_levels(int w, int h, int win, int max_level);
void orc_scharr(const uint8_t *src, int w, int h, int pitch, int16_t *dst);
int orc_lk_pyr(const uint8_t *prev, const uint8_t *next, int w, int h, int pitch, const float *pts_in, int npts,
               float *pts_out, uint8_t *status, int win, int max_level, int max_iters, double eps, float min_eig_thr);
int orc_grid_points(int w, int h, int ps, float *pts);
int orc_flow_filter(const float *p1, const float *p2, const uint8_t *status, int npts, double min_vec,
                    uint8_t *keep, double *flow4);
int orc_perspective_4pt(const double *src, const double *dst, double *H);
int orc_fit_egomotion(const float *p1, const float *p2, const uint8_t *keep, int npts, int w, int h, int mode,
                      int iters, double thr, uint32_t seed, double *H, uint8_t *inlier);
int orc_invert3(const double *S, double *D);
void orc_warp_perspective_inv(const uint8_t *src, int w, int h, int pitch, const double *M, uint8_t *dst, int dpitch);
void orc_warp_perspective(const uint8_t *src, int w, int h, int pitch, const double *H, uint8_t *dst, int dpitch);
void orc_absdiff_threshold(const uint8_t *a, const uint8_t *b, int w, int h, int pitch, int thresh, uint8_t *dst, int dpitch);
void orc_erode3(const uint8_t *src, int w, int h, int pitch, uint8_t *dst, int dpitch);
void orc_dilate3(const uint8_t *src, int w, int h, int pitch, uint8_t *dst, int dpitch);
void orc_motion_mask(const uint8_t *prev, const uint8_t *cur, int w, int h, int pitch, const double *H, int thresh,
                     int morph, uint8_t *mask, int mpitch);

/* md_oracle_varflow.c */
int orc_varflow(const uint8_t *A, const uint8_t *B, int w, int h, int pitch, int max_level, int start_level,
                int n1, int n2, float rho, float alpha, float sigma, float *U, float *V, int literal_corrections);
void orc_gaussian_blur_f32(const float *src, int w, int h, float *dst, double sigma);
void orc_resize_linear_f32(const float *src, int sw, int sh, float *dst, int dw, int dh);

/* md_oracle_subspace.c */
int orc_fit_subspace(const float *traj, int T, int F, int num_motions, double sigma, uint32_t seed,
                     const int *forced_cols, int iters, float *residual, int *best_cols, uint8_t *outlier,
                     double *threshold_out);
void orc_subspace_projector(const float *data, int n, int T, const int *cols, int d, float *Pnd);

/* md_oracle_live.c */
void orc_traj_step(float *cur, const float *next, const uint8_t *status, float *traj, int32_t *len, int P, int F, int w, int h);
int orc_cluster_euclidean(const float *pts, int n, double distance_threshold, int32_t *labels);
int orc_cluster_vectors(const double *vec4, int n, double distance_threshold, double angular_threshold, int32_t *labels);
int orc_find_outliers(const double *dxdy, int n, int include_zeros, uint8_t *outlier, double *stats4);
int orc_bounding_boxes(const float *pts, int n, const int32_t *labels, int nclusters, int min_size, int32_t *boxes,
                       int32_t *sizes, int32_t *ids);

/* md_oracle_draw.c */
void orc_line_aa(uint8_t *img, int w, int h, int pitch, int nch, int px1, int py1, int px2, int py2, const uint8_t *colour);
int orc_arrow_segments(const double *elem, int pixel_step, double min_vector_size, int *seg);
int orc_draw_flow(const uint8_t *src, int w, int h, int nch, int pitch, const double *vec4, int n, int pixel_step,
                  double min_vector_size, const uint8_t *colour, uint8_t *dst, int dpitch);

#ifdef __cplusplus
}
#endif
#endif
