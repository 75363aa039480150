// opencv2/core/core.hpp -- SHIM (test infrastructure only): the few cv:: types the reference's clustering sources
// (common/src/flow_clusterer.cpp, vector_cluster.cpp, point_cluster.cpp) use, so that those files compile UNMODIFIED in an
// image without OpenCV and the oracle's restatement of clusterEuclidean / getClusters can be checked against the reference's
// own loops (oracle/_ref/libcluster_ref.so, tests/test_oracle_ref.py).  Like the real header (which pulls in
// opencv2/core/types_c.h -> <math.h>, and <cmath>, <limits>, <algorithm>, <vector>, ...) this one includes the C and C++ math
// headers: which overload the reference's unqualified sqrt(float) (point_cluster.cpp:64) and abs(double)
// (vector_cluster.cpp:43) calls resolve to is then decided by the toolchain exactly as in a real build.
#ifndef MD_REF_SHIM_CORE_HPP
#define MD_REF_SHIM_CORE_HPP
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

#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_64FC4 CV_MAKETYPE(CV_64F, 4)

namespace cv {
typedef unsigned char uchar;
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
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

class Mat {
public:
    int rows, cols;
    uchar *data;
    size_t step;
    Mat() : rows(0), cols(0), data(0), step(0), type_(0) {}
    Mat(int r, int c, int t) : rows(r), cols(c), data(0), step(0), type_(t)
    {
        step = (size_t)c * esz(t);
        buf_.reset(new std::vector<uchar>(step * (size_t)r + 1));
        data = buf_->data();
    }
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }          // std::vector value-initialises to 0
    static size_t esz(int t) { static const int d[7] = {1, 1, 2, 2, 4, 4, 8}; return (size_t)d[t & 7] * (size_t)((t >> 3) + 1); }
    int type() const { return type_; }
    template <typename T> T &at(int y, int x) { return *reinterpret_cast<T *>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    template <typename T> const T &at(int y, int x) const { return *reinterpret_cast<const T *>(data + (size_t)y * step + (size_t)x * sizeof(T)); }
    Mat rowRange(const Range &r) const { Mat m(*this); m.rows = r.end - r.start; m.data = data + (size_t)r.start * step; return m; }
private:
    int type_;
    std::shared_ptr<std::vector<uchar> > buf_;
};
}  // namespace cv
#endif
