/* md_oracle_live.c -- CPU restatement (TEST INFRASTRUCTURE ONLY) of the steps of the node's live path that follow the
 * trajectories (MotionDetectionNode::imageCallback, ros/src/motion_detection_node.cpp:294-414):
 *   orc_traj_step          calculateOpticalFlowTrajectory bookkeeping, common/src/optical_flow_calculator.cpp:178-241
 *   orc_cluster_euclidean  FlowClusterer::clusterEuclidean, common/src/flow_clusterer.cpp:227-269, with
 *                          PointCluster::getClosestDistance / getDistance, common/src/point_cluster.cpp:26-38,62-65
 *   orc_cluster_vectors    FlowClusterer::getClusters, common/src/flow_clusterer.cpp:178-227 (the live code after the
 *                          #ifdef NEW_THING block), with VectorCluster::getClosestDistance / getClosestOrientation /
 *                          getDistance / getAngularDistance / getAngle, common/src/vector_cluster.cpp:25-50,116-136
 *   orc_bounding_boxes     OpticalFlowVisualizer::showBoundingBoxes, common/src/optical_flow_visualizer.cpp:223-240
 *                          (cv::Mat(Point2f).copyTo(vector<Point>) = saturate_cast<int> = round half to even;
 *                          cv::boundingRect of integer points: tl = min, size = max - min + 1; br = tl + size) --
 *                          the row MotionLogger::writeBoundingBox logs (common/src/motion_logger.cpp:43-47).
 * Written as the reference writes it (explicit cluster member lists, clusters scanned in creation order); the GPU path
 * uses a different but equivalent formulation, which is what the parity test checks.
 * The reference holds no tests / fixtures for these functions.  orc_cluster_euclidean and orc_cluster_vectors are PINNED TO THE
 * REFERENCE'S OWN CODE: oracle/_ref/libcluster_ref.so is the unmodified flow_clusterer.cpp + vector_cluster.cpp +
 * point_cluster.cpp compiled against the cv:: shim of oracle/ref_shim/opencv2, and the clusters it returns equal the ones
 * derived from these labels exactly (tests/test_oracle_ref.py, frozen in tests/golden/golden_ref.npz).  orc_traj_step,
 * orc_bounding_boxes, orc_find_outliers: hand-worked cases in tests/test_oracle_golden.py (parity unpinned for those).
 */
#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "md_oracle.h"

/* One pair of calculateOpticalFlowTrajectory (cpp:178-241): a tracked point inside the 10 px margin extends its
 * trajectory and becomes the new start point; any other point keeps its old position. */
void orc_traj_step(float *cur, const float *next, const uint8_t *status, float *traj, int32_t *len, int P, int F, int w, int h)
{
    for (int i = 0; i < P; i++) {
        if (!status[i]) continue;                                      /* cpp:219-236: temp.push_back(points_image1) */
        const float x = next[2 * i], y = next[2 * i + 1];
        if (x > 10.0 && y > 10.0 && x < w - 10 && y < h - 10) {      /* cpp:207-208 */
            if (len[i] < F) { traj[((size_t)i * F + len[i]) * 2] = x; traj[((size_t)i * F + len[i]) * 2 + 1] = y; len[i]++; }
            cur[2 * i] = x; cur[2 * i + 1] = y;                        /* cpp:210-211 */
        }
    }
}

typedef struct { int *member; int n, cap; } orc_cluster;

static double cl_distance(const float *a, const float *b)
{
    /* point_cluster.cpp:62-65: float arithmetic inside; the unqualified sqrt() gets a float argument, so with a libstdc++
     * whose <math.h> exports the std:: overloads (gcc >= 6) the FLOAT overload is chosen and the distance is rounded to
     * f32 before it becomes the double return value.  (gcc 4.8-era headers would pick sqrt(double); the two differ only
     * when a distance sits within one f32 ulp of the threshold or of another cluster's distance.) */
    const float dx = a[0] - b[0], dy = a[1] - b[1];
    const float s = dx * dx + dy * dy;
    return (double)sqrtf(s);
}

/* labels[i] = id (creation order) of the cluster point i joined; returns the number of clusters founded */
int orc_cluster_euclidean(const float *pts, int n, double distance_threshold, int32_t *labels)
{
    orc_cluster *cl = NULL;
    int ncl = 0, capcl = 0;
    for (int i = 0; i < n; i++) {
        const float *p = pts + 2 * i;
        double closest = DBL_MAX;                                      /* flow_clusterer.cpp:234 */
        int index = -1;
        for (int k = 0; k < ncl; k++) {                                /* :236-244 */
            double md = DBL_MAX;                                       /* point_cluster.cpp:28-37 */
            for (int m = 0; m < cl[k].n; m++) {
                const double d = cl_distance(p, pts + 2 * cl[k].member[m]);
                if (d < md) md = d;
            }
            if (md < closest) { closest = md; index = k; }
        }
        if (index != -1 && closest < distance_threshold) {             /* :245-248 */
            orc_cluster *c = &cl[index];
            if (c->n == c->cap) { c->cap *= 2; c->member = (int *)realloc(c->member, sizeof(int) * c->cap); }
            c->member[c->n++] = i;
            labels[i] = index;
        } else {                                                       /* :249-254 */
            if (ncl == capcl) { capcl = capcl ? 2 * capcl : 16; cl = (orc_cluster *)realloc(cl, sizeof(orc_cluster) * capcl); }
            cl[ncl].cap = 8; cl[ncl].n = 1;
            cl[ncl].member = (int *)malloc(sizeof(int) * 8);
            cl[ncl].member[0] = i;
            labels[i] = ncl++;
        }
    }
    for (int k = 0; k < ncl; k++) free(cl[k].member);
    free(cl);
    return ncl;
}

/* ---- FlowClusterer::getClusters (flow_clusterer.cpp:178-227), literal ---------------------------------------------------
 * vec4 [n][4] = (x, y, dx, dy) of the participating flow vectors in the reference's traversal order (rows outer, columns
 * inner, both in steps of pixel_step, only vectors with |dx| > 0 or |dy| > 0, :184-185).  A vector joins the FIRST cluster
 * (creation order) whose closest member is nearer than distance_threshold AND whose closest orientation differs by less
 * than angular_threshold -- the two minima may come from different members (:192-199); otherwise it founds a cluster.
 * VectorCluster::getClosestOrientation calls an unqualified abs() on a double (vector_cluster.cpp:43): as for sqrt above,
 * the std::abs(double) overload is assumed (an int abs() would truncate every angle below 1 rad to 0). */
static double vc_angle(const double *v)
{
    double ang = atan2(v[3], v[2]);                                    /* vector_cluster.cpp:129-136 */
    if (ang < 0.0) ang += 2 * M_PI;
    return ang;
}
int orc_cluster_vectors(const double *vec4, int n, double distance_threshold, double angular_threshold, int32_t *labels)
{
    orc_cluster *cl = NULL;
    int ncl = 0, capcl = 0;
    for (int i = 0; i < n; i++) {
        const double *v = vec4 + 4 * i;
        int added = 0;
        for (int k = 0; k < ncl && !added; k++) {                      /* flow_clusterer.cpp:190-201 */
            double md = DBL_MAX, ma = DBL_MAX;
            for (int m = 0; m < cl[k].n; m++) {
                const double *u = vec4 + 4 * cl[k].member[m];
                const double d = sqrt((v[0] - u[0]) * (v[0] - u[0]) + (v[1] - u[1]) * (v[1] - u[1]));   /* vector_cluster.cpp:116-119 */
                if (d < md) md = d;
                const double a1 = vc_angle(v), a2 = vc_angle(u);       /* :121-127 */
                const double ad = fabs(atan2(sin(a1 - a2), cos(a1 - a2)));
                if (ad < ma) ma = ad;
            }
            if (md < distance_threshold && ma < angular_threshold) {
                orc_cluster *c = &cl[k];
                if (c->n == c->cap) { c->cap *= 2; c->member = (int *)realloc(c->member, sizeof(int) * c->cap); }
                c->member[c->n++] = i;
                labels[i] = k;
                added = 1;
            }
        }
        if (!added) {                                                  /* :202-207 */
            if (ncl == capcl) { capcl = capcl ? 2 * capcl : 16; cl = (orc_cluster *)realloc(cl, sizeof(orc_cluster) * capcl); }
            cl[ncl].cap = 8; cl[ncl].n = 1;
            cl[ncl].member = (int *)malloc(sizeof(int) * 8);
            cl[ncl].member[0] = i;
            labels[i] = ncl++;
        }
    }
    for (int k = 0; k < ncl; k++) free(cl[k].member);
    free(cl);
    return ncl;
}

/* clusters with MORE than min_size members (flow_clusterer.cpp:262-267), in creation order:
 * boxes[k] = tl.x, tl.y, br.x, br.y; sizes[k]; ids[k]; returns their number */
int orc_bounding_boxes(const float *pts, int n, const int32_t *labels, int nclusters, int min_size, int32_t *boxes,
                       int32_t *sizes, int32_t *ids)
{
    int *cnt = (int *)calloc((size_t)(nclusters > 0 ? nclusters : 1), sizeof(int));
    int *bb = (int *)malloc(sizeof(int) * 4 * (size_t)(nclusters > 0 ? nclusters : 1));
    for (int k = 0; k < nclusters; k++) { bb[4 * k] = bb[4 * k + 1] = 0x7fffffff; bb[4 * k + 2] = bb[4 * k + 3] = (int)0x80000000; }
    for (int i = 0; i < n; i++) {
        const int k = labels[i];
        const int x = (int)lrintf(pts[2 * i]), y = (int)lrintf(pts[2 * i + 1]);   /* saturate_cast<int>(float) */
        cnt[k]++;
        if (x < bb[4 * k]) bb[4 * k] = x;
        if (y < bb[4 * k + 1]) bb[4 * k + 1] = y;
        if (x > bb[4 * k + 2]) bb[4 * k + 2] = x;
        if (y > bb[4 * k + 3]) bb[4 * k + 3] = y;
    }
    int out = 0;
    for (int k = 0; k < nclusters; k++)
        if (cnt[k] > min_size) {
            boxes[4 * out] = bb[4 * k]; boxes[4 * out + 1] = bb[4 * k + 1];
            boxes[4 * out + 2] = bb[4 * k + 2] + 1; boxes[4 * out + 3] = bb[4 * k + 3] + 1;
            sizes[out] = cnt[k]; ids[out] = k;
            out++;
        }
    free(cnt); free(bb);
    return out;
}

/* ---- OutlierDetector::findOutliers / createMask (common/src/outlier_detector.cpp:37-186), literal ------------------------ */
static int cmp_double(const void *a, const void *b)
{
    const double x = *(const double *)a, y = *(const double *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}
static double orc_median(const double *v, int n)            /* getMedian :99-121 (takes its vector by value and sorts it) */
{
    double *t = (double *)malloc(sizeof(double) * (size_t)n);
    memcpy(t, v, sizeof(double) * (size_t)n);
    qsort(t, (size_t)n, sizeof(double), cmp_double);
    const double m = (n % 2 == 0) ? (t[n / 2 - 1] + t[n / 2]) / 2.0 : t[(n - 1) / 2];
    free(t);
    return m;
}
/* one createMask pass (:122-186) over `values`; mask is OR-ed; returns the number of participating vectors */
static int orc_create_mask(const double *dxdy, const double *values, int n, int include_zeros, uint8_t *mask, double *median_out,
                           double *mad_out)
{
    double *vals = (double *)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; i++)
        if (include_zeros || fabs(dxdy[2 * i]) > 0.0 || fabs(dxdy[2 * i + 1]) > 0.0) vals[m++] = values[i];
    if (m == 0) { free(vals); return 0; }
    const double median = orc_median(vals, m);
    for (int i = 0; i < m; i++) vals[i] = fabs(vals[i] - median);
    const double mad = orc_median(vals, m);
    int index = 0;
    for (int i = 0; i < n; i++)
        if (include_zeros || fabs(dxdy[2 * i]) > 0.0 || fabs(dxdy[2 * i + 1]) > 0.0) {
            const double z = 0.6745 * vals[index] / mad;
            if (fabs(z) > 3.5) mask[i] = 1;
            index++;
        }
    *median_out = median; *mad_out = mad;
    free(vals);
    return m;
}
/* findOutliers (:37-52): angle pass, then magnitude pass into the same mask.  stats4: median/MAD angle, median/MAD magnitude */
int orc_find_outliers(const double *dxdy, int n, int include_zeros, uint8_t *outlier, double *stats4)
{
    double *ang = (double *)malloc(sizeof(double) * (size_t)n), *mag = (double *)malloc(sizeof(double) * (size_t)n);
    for (int i = 0; i < n; i++) {
        const double dx = dxdy[2 * i], dy = dxdy[2 * i + 1];
        ang[i] = atan2(dy, dx);                                  /* :82 */
        mag[i] = sqrt(dy * dy + dx * dx);                        /* :94 */
    }
    memset(outlier, 0, (size_t)n);
    stats4[0] = stats4[1] = stats4[2] = stats4[3] = 0;
    int m = orc_create_mask(dxdy, ang, n, include_zeros, outlier, &stats4[0], &stats4[1]);
    orc_create_mask(dxdy, mag, n, include_zeros, outlier, &stats4[2], &stats4[3]);
    free(ang); free(mag);
    return m;
}
