/* outlier_detector.h -- drop-in replacement of common/include/motion_detection/outlier_detector.h:12-32 for the
 * RANSAC entry point (fitSubspace).  findOutliers / getOutlierVectors (the MAD path, uncalled by the node) stay with
 * the reference's own sources. */
#ifndef OUTLIER_DETECTOR_H_
#define OUTLIER_DETECTOR_H_

#include <opencv2/core/core.hpp>
#include <vector>

struct md_ctx;

class OutlierDetector
{
    public:
        OutlierDetector();
        virtual ~OutlierDetector();

        /* common/src/outlier_detector.cpp:236-331 */
        std::vector<std::vector<cv::Point2f> > fitSubspace(const std::vector<std::vector<cv::Point2f> > &trajectories, std::vector<cv::Point2f> &outlier_points, int num_motions, double sigma);

        void setDevice(int device) { device_ = device; }
        void setSeed(unsigned seed) { seed_ = seed; }        /* the reference seeds with time(NULL), cpp:17 */
        int lastInliers() const { return last_inliers_; }

    private:
        OutlierDetector(const OutlierDetector &);
        OutlierDetector &operator=(const OutlierDetector &);
        md_ctx *ctx_;
        int device_;
        unsigned seed_;
        int last_inliers_;
};
#endif
