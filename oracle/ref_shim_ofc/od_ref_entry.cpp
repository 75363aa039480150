// od_ref_entry.cpp -- C entry points into the reference's own OutlierDetector (common/src/outlier_detector.cpp compiled unmodified
// against oracle/ref_shim_ofc: cv::Mat stand-in + the Eigen stand-in of Eigen/Dense), test infrastructure only.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <motion_detection/outlier_detector.h>      // the REFERENCE's header

// findOutliers + getOutlierVectors (outlier_detector.cpp:37-186, node.cpp:112-121) on a Vec4d field [h][w][4]:
// prob [h][w] f64 = the outlier_probabilities matrix (0 / 1 after the angle stage, overwritten by the magnitude stage for the nodes it flags)
extern "C" void ref_find_outliers(const double *flow, int w, int h, int pixel_step, int include_zeros, double *prob)
{
    cv::Mat f(h, w, CV_64FC4), p;
    memcpy(f.data, flow, sizeof(double) * 4 * (size_t)w * h);
    OutlierDetector od;
    od.findOutliers(f, p, include_zeros != 0, pixel_step, false);
    for (int y = 0; y < h; y++) memcpy(prob + (size_t)y * w, p.data + (size_t)y * p.step, sizeof(double) * (size_t)w);
}
// fitSubspace (outlier_detector.cpp:236-331) on trajectories [T][F][2] after srand(seed) (the ctor's srand(time) is overridden):
// outlier_idx [T] = indices of the trajectories it reports (recovered from the reported point = the second-to-last position), cols [4 * num_motions]
// = the sampled columns of the winning hypothesis (recovered by matching the returned trajectories); returns the number of outliers
extern "C" int ref_fit_subspace(const float *traj, int T, int F, int num_motions, double sigma, unsigned seed, int32_t *outlier_idx, int32_t *cols,
                                int *ncols)
{
    std::vector<std::vector<cv::Point2f> > tr(T);
    for (int t = 0; t < T; t++)
        for (int j = 0; j < F; j++) tr[t].push_back(cv::Point2f(traj[((size_t)t * F + j) * 2], traj[((size_t)t * F + j) * 2 + 1]));
    OutlierDetector od;
    srand(seed);
    std::vector<cv::Point2f> out;
    std::vector<std::vector<cv::Point2f> > basis = od.fitSubspace(tr, out, num_motions, sigma);
    int n = 0, scan = 0;
    for (size_t i = 0; i < out.size(); i++)                         // outliers are reported in trajectory order
        for (; scan < T; scan++)
            if (tr[scan][F - 2].x == out[i].x && tr[scan][F - 2].y == out[i].y) { outlier_idx[n++] = scan++; break; }
    *ncols = (int)basis.size();
    for (size_t i = 0; i < basis.size(); i++) {
        cols[i] = -1;
        for (int t = 0; t < T && cols[i] < 0; t++) {
            bool same = true;
            for (int j = 0; j < F && same; j++) same = tr[t][j].x == basis[i][j].x && tr[t][j].y == basis[i][j].y;
            if (same) cols[i] = t;
        }
    }
    return n == (int)out.size() ? n : -1;
}
