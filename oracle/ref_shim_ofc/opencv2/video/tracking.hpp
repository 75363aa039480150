// opencv2/video/tracking.hpp -- SHIM, see opencv2/core/core.hpp of this directory
#ifndef MD_REF_SHIM_OFC_TRACKING_HPP
#define MD_REF_SHIM_OFC_TRACKING_HPP
#include <opencv2/core/core.hpp>
namespace cv {
// keeps the gray image as level 0 of `pyramid`; the LK shim builds the levels itself (orc_lk_pyr), like cv::calcOpticalFlowPyrLK does
// with an image input -- buildOpticalFlowPyramid + calcOpticalFlowPyrLK(pyramid) and calcOpticalFlowPyrLK(image) are the same arithmetic
int buildOpticalFlowPyramid(const Mat &img, std::vector<Mat> &pyramid, Size winSize, int maxLevel, bool withDerivatives = true);
void calcOpticalFlowPyrLK(const std::vector<Mat> &prevPyr, const Mat &nextImg, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts,
                          std::vector<uchar> &status, std::vector<float> &err, Size winSize, int maxLevel, TermCriteria criteria, int flags,
                          double minEigThreshold);
void calcOpticalFlowPyrLK(const Mat &prevImg, const Mat &nextImg, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts,
                          std::vector<uchar> &status, std::vector<float> &err, Size winSize, int maxLevel, TermCriteria criteria, int flags,
                          double minEigThreshold);
}
#endif
