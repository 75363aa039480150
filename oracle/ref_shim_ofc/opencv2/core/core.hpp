// opencv2/core/core.hpp -- SHIM (test infrastructure only) for compiling the reference's UNMODIFIED
// common/src/optical_flow_calculator.cpp in an image without OpenCV headers (oracle/_ref/libofc_ref.so, recipe oracle/Makefile,
// test tests/test_oracle_ref.py).  The cv:: types and functions that file uses; the image-processing functions forward to the
// oracle's cv2-pinned primitives (orc_gray_bgr2gray, orc_lk_pyr, orc_perspective_4pt, orc_warp_perspective: compared with cv2 in
// tests/test_oracle_vs_cv2.py), so what the resulting library pins is the reference's own COMPOSITION: grid order, vector filter,
// Vec4d bookkeeping, first-four getPerspectiveTransform, warp / absdiff / threshold, trajectory bookkeeping.
// Nothing here is reference code.
#ifndef MD_REF_SHIM_OFC_CORE_HPP
#define MD_REF_SHIM_OFC_CORE_HPP
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <vector>

#include <cv.h>          // the legacy C types of oracle/ref_shim (IplImage, CvSize, CvPoint, ...)

typedef unsigned char uchar;
typedef struct CvScalar { double val[4]; } CvScalar;
static inline CvScalar cvScalar(double a, double b, double c, double d) { CvScalar s; s.val[0] = a; s.val[1] = b; s.val[2] = c; s.val[3] = d; return s; }
#define CV_RGB(r, g, b) cvScalar((b), (g), (r), 0)
static inline int cvRound(double v) { return (int)lrint(v); }

#define CV_8U 0
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)
#define CV_64FC4 CV_MAKETYPE(CV_64F, 4)
#define CV_TERMCRIT_ITER 1
#define CV_TERMCRIT_EPS 2
#define CV_THRESH_BINARY 0
#define CV_BGR2GRAY 6
#define CV_BGR2Lab 44
#define CV_AA 16

namespace cv {
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <typename U, typename W> Point_(U a, W b) : x((T)a), y((T)b) {}
};
typedef Point_<float> Point2f;
template <typename T, int N> struct Vec {
    T val[N];
    Vec() { for (int i = 0; i < N; i++) val[i] = T(); }
    T &operator[](int i) { return val[i]; }
    const T &operator[](int i) const { return val[i]; }
};
typedef Vec<double, 4> Vec4d;
template <typename T, int N> std::ostream &operator<<(std::ostream &o, const Vec<T, N> &v) { o << "["; for (int i = 0; i < N; i++) o << (i ? ", " : "") << v[i]; return o << "]"; }
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Scalar { double val[4]; Scalar(const CvScalar &s) { for (int i = 0; i < 4; i++) val[i] = s.val[i]; } };
struct TermCriteria { int type, maxCount; double epsilon; TermCriteria(int t, int n, double e) : type(t), maxCount(n), epsilon(e) {} };

class Mat {
public:
    int rows, cols;
    uchar *data;
    size_t step;
    Mat() : rows(0), cols(0), data(0), step(0), type_(0) {}
    Mat(int r, int c, int t) : rows(0), cols(0), data(0), step(0), type_(0) { create(r, c, t); }
    Mat(int r, int c, int t, void *ext, size_t st) : rows(r), cols(c), data((uchar *)ext), step(st), type_(t) {}    // external data, not owned
    void create(int r, int c, int t)
    {
        if (data && r == rows && c == cols && t == type_) return;
        rows = r; cols = c; type_ = t;
        step = (size_t)c * esz(t);
        buf_.reset(new std::vector<uchar>(step * (size_t)r + 64));
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }          // std::vector value-initialises to 0
    static size_t esz(int t) { static const int d[7] = {1, 1, 2, 2, 4, 4, 8}; return (size_t)d[t & 7] * (size_t)((t >> 3) + 1); }
    int type() const { return type_; }
    int channels() const { return (type_ >> 3) + 1; }
    bool empty() const { return data == 0 || rows * cols == 0; }
    Size size() const { return Size(cols, rows); }
    void copyTo(Mat &m) const
    {
        m.create(rows, cols, type_);
        for (int y = 0; y < rows; y++) memcpy(m.data + (size_t)y * m.step, data + (size_t)y * step, (size_t)cols * esz(type_));
    }
    Mat clone() const { Mat m; copyTo(m); return m; }
    template <typename T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};

void absdiff(const Mat &a, const Mat &b, Mat &dst);
Mat cvarrToMat(const IplImage *img);
}  // namespace cv
#endif
