// opencv2/imgproc/imgproc.hpp -- SHIM, see opencv2/core/core.hpp of this directory
#ifndef MD_REF_SHIM_OFC_IMGPROC_HPP
#define MD_REF_SHIM_OFC_IMGPROC_HPP
#include <opencv2/core/core.hpp>
namespace cv {
void cvtColor(const Mat &src, Mat &dst, int code);                              // CV_BGR2GRAY of 1- or 3-channel u8
Mat getPerspectiveTransform(const Point2f src[], const Point2f dst[]);          // 3 x 3 CV_64F
void warpPerspective(const Mat &src, Mat &dst, const Mat &M, Size dsize);       // INTER_LINEAR, BORDER_CONSTANT 0, M = src -> dst
double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type);
void line(Mat &img, Point2f p1, Point2f p2, const Scalar &colour, int thickness, int lineType, int shift);   // superPixelFlow only: not on the tested path
}
#endif
