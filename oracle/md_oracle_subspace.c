/*
 * md_oracle_subspace.c -- CPU ORACLE (TEST INFRASTRUCTURE ONLY) for the reference's RANSAC:
 *   OutlierDetector::fitSubspace   common/src/outlier_detector.cpp:236-331
 *   fillMatrix :188-200, meanSubtract :202-221, fillSubset :223-234, chi-square table :19-30
 *
 * Deviations from the literal Eigen/float arithmetic, all stated in DESIGN.md:
 *  - the SVD (Eigen::JacobiSVD<MatrixXf>, un-vendored, :266) is restated as a one-sided (Hestenes) Jacobi
 *    in f64 on the f32 data; only span(U[:, :d]) enters the result, so for a full-rank sample any SVD
 *    gives the same projector.  Columns whose singular value is < 1e-12 * sigma_max (duplicate draws,
 *    sampling is WITH replacement, :228) contribute nothing, where Eigen would pick an arbitrary
 *    orthonormal completion -- "parity asserted on full-rank samples" (SURVEY.md section 7).
 *  - row means and the quadratic form d^T Pnd d are accumulated in f64 (Eigen: f32, whose rounding noise
 *    on |d|^2 ~ 1e6 is itself ~0.1, the same order as the sigma=0.5 threshold).
 * Pinned against numpy.linalg.svd in tests/test_oracle_vs_cv2.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "md_oracle.h"

#define MAXN 32 /* 2F <= 32 */

/* chi_square_table p99, outlier_detector.cpp:19-30 */
static const double CHI2_P99[10] = {0.0, 0.020, 0.115, 0.297, 0.554, 0.872, 1.239, 1.646, 2.088, 2.558};

/* Orthonormal basis of span(A[:, 0..d)) by one-sided (Hestenes) Jacobi; A is n x d column-major in f64 (destroyed).
 * On return column k of A is sigma_k * u_k; norms[k] = sigma_k.
 * Sweep order: ROUND-ROBIN.  A sweep is d' - 1 rounds (d' = d rounded up to even) of d'/2 disjoint column pairs -- the
 * chess-tournament schedule: round r pairs (r, d'-1) and ((r + i) mod (d'-1), (r - i) mod (d'-1)), i = 1 .. d'/2 - 1.
 * The pairs of a round touch different columns, so their rotations commute EXACTLY: any order (or all at once, as the
 * device does) gives the same bits.  (The reference uses Eigen's JacobiSVD in f32; any convergent Jacobi order spans the
 * same subspace -- tests/test_oracle_vs_cv2.py pins the projector against numpy's SVD.) */
static int jacobi_pair(double *A, int n, int p, int q)
{
    double a = 0, b = 0, g = 0;
    for (int i = 0; i < n; i++) {
        a += A[p * n + i] * A[p * n + i];
        b += A[q * n + i] * A[q * n + i];
        g += A[p * n + i] * A[q * n + i];
    }
    if (g == 0 || fabs(g) <= 1e-15 * sqrt(a * b)) return 0;
    double zeta = (b - a) / (2 * g);
    double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1 + zeta * zeta));
    double c = 1 / sqrt(1 + t * t), s = c * t;
    for (int i = 0; i < n; i++) {
        double vp = A[p * n + i], vq = A[q * n + i];
        A[p * n + i] = c * vp - s * vq;
        A[q * n + i] = s * vp + c * vq;
    }
    return 1;
}

static void hestenes(double *A, int n, int d, double *norms)
{
    const int de = d + (d & 1), m = de - 1;          /* an odd d gets a bye (the pair with the phantom column is skipped) */
    for (int sweep = 0; sweep < 60; sweep++) {
        int rotated = 0;
        for (int r = 0; r < m; r++)
            for (int i = 0; i < de / 2; i++) {
                int p = i == 0 ? r : (r + i) % m, q = i == 0 ? m : (r - i + m) % m;
                if (p > q) { int t = p; p = q; q = t; }
                if (q >= d || p == q) continue;
                rotated |= jacobi_pair(A, n, p, q);
            }
        if (!rotated) break;
    }
    for (int k = 0; k < d; k++) {
        double a = 0;
        for (int i = 0; i < n; i++) a += A[k * n + i] * A[k * n + i];
        norms[k] = sqrt(a);
    }
}

/* data: n x T row-major f32 (already mean-subtracted); cols: d sampled column indices.
 * Pnd (n x n row-major, f32 out for inspection; computed in f64) = I - sum_k u_k u_k^T  (:269-282). */
static void projector_f64(const float *data, int n, int T, const int *cols, int d, double *P)
{
    double A[MAXN * MAXN], norms[MAXN];
    for (int k = 0; k < d; k++)
        for (int i = 0; i < n; i++) A[k * n + i] = data[(size_t)i * T + cols[k]];
    hestenes(A, n, d, norms);
    double smax = 0;
    for (int k = 0; k < d; k++) if (norms[k] > smax) smax = norms[k];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) P[i * n + j] = i == j ? 1.0 : 0.0;
    for (int k = 0; k < d; k++) {
        if (!(norms[k] > 1e-12 * smax) || norms[k] == 0) continue;
        double inv = 1.0 / norms[k];
        for (int i = 0; i < n; i++)
            for (int j = 0; j < n; j++) P[i * n + j] -= (A[k * n + i] * inv) * (A[k * n + j] * inv);
    }
}

void orc_subspace_projector(const float *data, int n, int T, const int *cols, int d, float *Pnd)
{
    double P[MAXN * MAXN];
    projector_f64(data, n, T, cols, d, P);
    for (int i = 0; i < n * n; i++) Pnd[i] = (float)P[i];
}

/* traj: [T][F][2] f32 (x, y).  forced_cols: NULL (draw rand() % T after srand(seed)) or [iters][d] indices.
 * residual[T]: |d_i^T Pnd d_i| of the winning hypothesis; best_cols[d]; outlier[T] = residual > threshold.
 * Returns the winner's inlier count (0: nothing won, outputs zeroed / -1). */
int orc_fit_subspace(const float *traj, int T, int F, int num_motions, double sigma, uint32_t seed,
                     const int *forced_cols, int iters, float *residual, int *best_cols, uint8_t *outlier,
                     double *threshold_out)
{
    int n = 2 * F, d = 4 * num_motions;
    if (T < 1 || n > MAXN || d > MAXN || d < 1) return -1;
    float *data = (float *)malloc(sizeof(float) * (size_t)n * T);
    /* fillMatrix :188-200 */
    for (int i = 0; i < T; i++)
        for (int j = 0; j < F; j++) {
            data[(size_t)(2 * j) * T + i] = traj[((size_t)i * F + j) * 2];
            data[(size_t)(2 * j + 1) * T + i] = traj[((size_t)i * F + j) * 2 + 1];
        }
    /* meanSubtract :202-221: frame-0 means; x rows -= mean_x, y rows = mean_y - y */
    double xs = 0, ys = 0;
    for (int i = 0; i < T; i++) { xs += data[i]; ys += data[(size_t)T + i]; }
    xs /= T; ys /= T;
    float xm = (float)xs, ym = (float)ys;
    for (int r = 0; r < n; r++)
        for (int i = 0; i < T; i++) {
            float *p = &data[(size_t)r * T + i];
            *p = (r % 2 == 0) ? (*p - xm) : (ym - *p);
        }
    orc_rand rs;
    orc_glibc_srand(&rs, seed);
    double *res = (double *)malloc(sizeof(double) * (size_t)T);
    int max_points = 0;
    int cols[MAXN];
    for (int k = 0; k < d; k++) best_cols[k] = -1;
    memset(residual, 0, sizeof(float) * (size_t)T);
    double inl_thr = (n - d) * sigma * sigma;    /* :294 */
    for (int it = 0; it < iters; it++) {
        for (int k = 0; k < d; k++) cols[k] = forced_cols ? forced_cols[it * d + k] : orc_glibc_rand(&rs) % T;   /* :228 */
        double P[MAXN * MAXN];
        projector_f64(data, n, T, cols, d, P);
        int cnt = 0;
        for (int i = 0; i < T; i++) {
            double acc = 0;
            for (int r = 0; r < n; r++) {
                double pr = 0;
                for (int c = 0; c < n; c++) pr += P[r * n + c] * data[(size_t)c * T + i];
                acc += data[(size_t)r * T + i] * pr;
            }
            res[i] = fabs(acc);
            if (res[i] < inl_thr) cnt++;
        }
        if (cnt > max_points) {                  /* :300-308 */
            max_points = cnt;
            for (int i = 0; i < T; i++) residual[i] = (float)res[i];
            memcpy(best_cols, cols, sizeof(int) * (size_t)d);
        }
    }
    double thr = 0.2;                            /* :312-317 */
    if (n - d < 11 && n - d > 0) thr = sigma * sigma * CHI2_P99[n - d >= 10 ? 9 : n - d];
    if (threshold_out) *threshold_out = thr;
    for (int i = 0; i < T; i++) outlier[i] = (max_points > 0 && residual[i] > thr) ? 1 : 0;   /* :318-324 */
    free(res); free(data);
    return max_points;
}
