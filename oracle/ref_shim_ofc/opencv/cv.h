/* opencv/cv.h -- SHIM, see opencv2/core/core.hpp of this directory: the legacy C calls of superPixelFlow / varFlow / drawMotionField
 * (optical_flow_calculator.cpp:343-355,432-501).  cvCloneImage / cvCvtColor(BGR2Lab) / cvLine belong to superPixelFlow and the arrow
 * drawing, which are not on the tested path: they abort when reached. */
#ifndef MD_REF_SHIM_OFC_OPENCV_CV_H
#define MD_REF_SHIM_OFC_OPENCV_CV_H
#include <opencv2/core/core.hpp>
#ifdef __cplusplus
extern "C" {
#endif
IplImage *cvCloneImage(const IplImage *img);
void cvCvtColor(const CvArr *src, CvArr *dst, int code);
void cvLine(CvArr *img, CvPoint p0, CvPoint p1, CvScalar colour, int thickness, int line_type, int shift);
#ifdef __cplusplus
}
#endif
#endif
