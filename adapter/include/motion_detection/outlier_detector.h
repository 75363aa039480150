/* outlier_detector.h -- drop-in replacement of common/include/motion_detection/outlier_detector.h:12-32: the RANSAC entry
 * point (fitSubspace) and the MAD path (findOutliers / getOutlierVectors, node.cpp:112-121). */
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

        /* common/src/outlier_detector.cpp:37-52: outlier_probabilities (CV_64F) = 1.0 at grid vectors whose angle or magnitude
         * fails the median / MAD test */
        void findOutliers(const cv::Mat &optical_flow_vectors, cv::Mat &outlier_probabilities, bool include_zeros, int pixel_step, bool print);
        /* :54-73 (the reference allocates CV_32FC4 and writes Vec4d; here the field is CV_64FC4) */
        void getOutlierVectors(const cv::Mat &optical_flow_vectors, const cv::Mat &outlier_probabilities, cv::Mat &outlier_vectors, int pixel_step);

        void setDevice(int device) { device_ = device; }
        void setSeed(unsigned seed) { seed_ = seed; }        /* the reference seeds with time(NULL), cpp:17 */
        int lastInliers() const { return last_inliers_; }

    private:
        bool ensureContext();
        OutlierDetector(const OutlierDetector &);
        OutlierDetector &operator=(const OutlierDetector &);
        md_ctx *ctx_;
        int device_;
        unsigned seed_;
        int last_inliers_;
};
#endif
