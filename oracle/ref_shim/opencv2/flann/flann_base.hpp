// SHIM (test infrastructure only, see opencv2/core/core.hpp here): just enough of cvflann for FlowClusterer::clusterFlowVectors
// (common/src/flow_clusterer.cpp:23-78, no call site) to COMPILE.  hierarchicalClustering is not implemented: it reports 0 centres.
#ifndef MD_REF_SHIM_FLANN_HPP
#define MD_REF_SHIM_FLANN_HPP
#include <stddef.h>
namespace cvflann {
enum flann_centers_init_t { FLANN_CENTERS_RANDOM = 0, FLANN_CENTERS_GONZALES = 1, FLANN_CENTERS_KMEANSPP = 2 };
struct KMeansIndexParams { KMeansIndexParams(int = 32, int = 11, flann_centers_init_t = FLANN_CENTERS_RANDOM, float = 0.2f) {} };
template <typename T> struct Matrix {
    size_t rows, cols;
    T *data;
    Matrix(T *d, size_t r, size_t c) : rows(r), cols(c), data(d) {}
    T *operator[](size_t i) const { return data + i * cols; }
};
template <typename T> struct L2 { typedef T ElementType; typedef float ResultType; };
template <typename Distance>
int hierarchicalClustering(const Matrix<typename Distance::ElementType> &, Matrix<typename Distance::ResultType> &, const KMeansIndexParams &) { return 0; }
}  // namespace cvflann
#endif
