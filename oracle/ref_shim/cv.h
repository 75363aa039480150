/* cv.h / cxcore.h -- SHIM of the OpenCV legacy C API, test infrastructure only.
 *
 * Purpose: compile the reference's UNMODIFIED common/src/VarFlow.cpp (which includes <cv.h> and <cxcore.h>,
 * common/include/motion_detection/VarFlow.h:11-12) in an image that has no OpenCV C/C++ headers, so that the oracle's
 * restatement of the in-tree arithmetic (gauss_seidel_step / _iteration / _recursive, residual_part_step,
 * calculate_residual, the CalcFlow schedule) can be checked bit for bit against the reference's own loops
 * (oracle/_ref/libvarflow_ref.so, recipe in oracle/Makefile, test tests/test_oracle_ref.py).
 *
 * Only the 16 entry points VarFlow.cpp calls exist.  The image-processing ones (cvSmooth, cvResize, cvFilter2D) are
 * implemented on the oracle's cv2-pinned primitives (orc_gaussian_blur_f32, orc_resize_linear_f32: compared with
 * cv2.GaussianBlur / cv2.resize / cv2.filter2D in tests/test_oracle_vs_cv2.py); the element-wise ones are plain f32 loops
 * with OpenCV's operation order.  Nothing here is reference code. */
#ifndef MD_REF_SHIM_CV_H
#define MD_REF_SHIM_CV_H
#include <math.h>
#include <stddef.h>

#define IPL_DEPTH_8U 8
#define IPL_DEPTH_32F 32
#define CV_32F 5
#define CV_INTER_LINEAR 1
#define CV_GAUSSIAN 2
#define CV_SWAP(a, b, t) ((t) = (a), (a) = (b), (b) = (t))

typedef struct _IplImage {
    int nChannels, depth, width, height, widthStep, imageSize;
    char *imageData;
} IplImage;
typedef struct CvSize { int width, height; } CvSize;
typedef struct CvPoint { int x, y; } CvPoint;
typedef struct CvMat {
    int type, step, rows, cols;
    union { unsigned char *ptr; float *fl; } data;
} CvMat;
typedef void CvArr;

static inline CvSize cvSize(int w, int h) { CvSize s; s.width = w; s.height = h; return s; }
static inline CvPoint cvPoint(int x, int y) { CvPoint p; p.x = x; p.y = y; return p; }
static inline CvMat cvMat(int rows, int cols, int type, void *data)
{
    CvMat m; m.type = type; m.rows = rows; m.cols = cols; m.step = cols * 4; m.data.ptr = (unsigned char *)data; return m;
}

#ifdef __cplusplus
extern "C" {
#endif
IplImage *cvCreateImage(CvSize size, int depth, int channels);     /* rows aligned to 4 bytes like OpenCV's default */
void cvReleaseImage(IplImage **img);
void cvZero(CvArr *arr);
void cvResize(const CvArr *src, CvArr *dst, int interpolation);
void cvConvertScale(const CvArr *src, CvArr *dst, double scale, double shift);
void cvSmooth(const CvArr *src, CvArr *dst, int smoothtype, int size1, int size2, double sigma1, double sigma2);
void cvFilter2D(const CvArr *src, CvArr *dst, const CvMat *kernel, CvPoint anchor);
void cvSub(const CvArr *a, const CvArr *b, CvArr *dst, const CvArr *mask);
void cvAdd(const CvArr *a, const CvArr *b, CvArr *dst, const CvArr *mask);
void cvMul(const CvArr *a, const CvArr *b, CvArr *dst, double scale);
void cvAddWeighted(const CvArr *a, double alpha, const CvArr *b, double beta, double gamma, CvArr *dst);
#ifdef __cplusplus
}
#endif
#define cvConvert(src, dst) cvConvertScale((src), (dst), 1, 0)
#define cvScale(src, dst, scale) cvConvertScale((src), (dst), (scale), 0)
#endif
