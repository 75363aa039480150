// Minimal stand-in for the parts of <opencv2/core/core.hpp> (and the legacy IplImage) that the adapter touches.
// ONLY for building / testing the adapter in images without OpenCV (this one has no OpenCV C++ headers).
// With real OpenCV on the include path this directory is simply not added to -I.
#ifndef MD_CV_STUB_CORE_HPP
#define MD_CV_STUB_CORE_HPP
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_32F 5
#define CV_64F 6
#define CV_CN_SHIFT 3
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << CV_CN_SHIFT))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_32FC4 CV_MAKETYPE(CV_32F, 4)
#define CV_64FC4 CV_MAKETYPE(CV_64F, 4)
#define IPL_DEPTH_8U 8
#define IPL_DEPTH_32F 32

namespace cv {
typedef unsigned char uchar;
struct Point2f { float x, y; Point2f() : x(0), y(0) {} Point2f(float a, float b) : x(a), y(b) {} };
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
template <typename T, int N> struct Vec { T val[N]; T &operator[](int i) { return val[i]; } const T &operator[](int i) const { return val[i]; } };
typedef Vec<double, 4> Vec4d;
typedef Vec<int, 4> Vec4i;
struct Scalar { double val[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; } };
struct Rect { int x, y, width, height; };

class Mat {
public:
    int rows, cols;
    uchar *data;
    size_t step;
    Mat() : rows(0), cols(0), data(nullptr), step(0), type_(0) {}
    Mat(int r, int c, int t) : rows(0), cols(0), data(nullptr), step(0), type_(0) { create(r, c, t); }
    static Mat zeros(int r, int c, int t) { Mat m(r, c, t); std::memset(m.data, 0, m.step * r); return m; }
    static size_t elemSizeOf(int t) { static const int d[7] = {1, 1, 2, 2, 4, 4, 8}; return (size_t)d[t & 7] * (size_t)((t >> CV_CN_SHIFT) + 1); }
    void create(int r, int c, int t)
    {
        if (r == rows && c == cols && t == type_ && data) return;
        rows = r; cols = c; type_ = t; step = (size_t)c * elemSizeOf(t);
        buf_.reset(new std::vector<uchar>(step * (size_t)r));
        data = buf_->data();
    }
    int type() const { return type_; }
    int channels() const { return (type_ >> CV_CN_SHIFT) + 1; }
    size_t elemSize() const { return elemSizeOf(type_); }
    bool empty() const { return !data || rows == 0 || cols == 0; }
    template <typename T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    void copyTo(Mat &o) const { o.create(rows, cols, type_); if (data) std::memcpy(o.data, data, step * (size_t)rows); }
    void setTo0() { if (data) std::memset(data, 0, step * (size_t)rows); }
private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};
}  // namespace cv

// legacy C types used by VarFlow's interface (VarFlow.h:33-36) and drawMotionField (optical_flow_calculator.h:29)
struct IplImage { int nChannels, depth, width, height, widthStep; char *imageData; };
struct CvScalar { double val[4]; };
struct CvPoint { int x, y; };
inline CvPoint cvPoint(int x, int y) { CvPoint p = {x, y}; return p; }
inline CvScalar cvScalar(double a, double b = 0, double c = 0, double d = 0) { CvScalar s = {{a, b, c, d}}; return s; }
#define CV_RGB(r, g, b) cv::Scalar((b), (g), (r), 0)
#define CV_AA 16
// plain 8-connected line into an 8-bit IplImage (stand-in for cvLine; real OpenCV draws the anti-aliased one)
inline void cvLine(IplImage *img, CvPoint a, CvPoint b, CvScalar color, int = 1, int = 8, int = 0)
{
    int dx = b.x > a.x ? b.x - a.x : a.x - b.x, dy = b.y > a.y ? b.y - a.y : a.y - b.y;
    const int sx = a.x < b.x ? 1 : -1, sy = a.y < b.y ? 1 : -1;
    int err = dx - dy, x = a.x, y = a.y;
    for (;;) {
        if (x >= 0 && y >= 0 && x < img->width && y < img->height)
            for (int c = 0; c < img->nChannels; c++) img->imageData[(size_t)y * img->widthStep + (size_t)x * img->nChannels + c] = (char)(unsigned char)color.val[c];
        if (x == b.x && y == b.y) break;
        const int e2 = 2 * err;
        if (e2 > -dy) { err -= dy; x += sx; }
        if (e2 < dx) { err += dx; y += sy; }
    }
}
#endif
