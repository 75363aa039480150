// cv_shim.cpp -- implementation of oracle/ref_shim/cv.h (test infrastructure only, see that header) and the C entry point
// ref_varflow() that runs the reference's own VarFlow class.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <motion_detection/VarFlow.h>      // the REFERENCE's header (common/include), which includes this shim's <cv.h>

#include "../md_oracle.h"

static IplImage *I(const CvArr *a) { return (IplImage *)a; }
static float *row(const IplImage *m, int y) { return (float *)(m->imageData + (size_t)y * m->widthStep); }

extern "C" IplImage *cvCreateImage(CvSize size, int depth, int channels)
{
    IplImage *m = (IplImage *)calloc(1, sizeof(IplImage));
    m->nChannels = channels; m->depth = depth; m->width = size.width; m->height = size.height;
    m->widthStep = (size.width * channels * (depth / 8) + 3) & ~3;
    m->imageSize = m->widthStep * size.height;
    m->imageData = (char *)malloc(m->imageSize > 0 ? m->imageSize : 1);
    return m;
}
extern "C" void cvReleaseImage(IplImage **img)
{
    if (img && *img) { free((*img)->imageData); free(*img); *img = NULL; }
}
extern "C" void cvZero(CvArr *arr) { memset(I(arr)->imageData, 0, (size_t)I(arr)->imageSize); }

// dense copies in / out of the (possibly row-padded) IplImage
static std::vector<float> pack(const IplImage *m)
{
    std::vector<float> v((size_t)m->width * m->height);
    for (int y = 0; y < m->height; y++) memcpy(&v[(size_t)y * m->width], row(m, y), sizeof(float) * m->width);
    return v;
}
static void unpack(const std::vector<float> &v, IplImage *m)
{
    for (int y = 0; y < m->height; y++) memcpy(row(m, y), &v[(size_t)y * m->width], sizeof(float) * m->width);
}

extern "C" void cvResize(const CvArr *src, CvArr *dst, int)
{
    const IplImage *s = I(src);
    IplImage *d = I(dst);
    if (s->depth == IPL_DEPTH_8U) {
        // VarFlow.cpp:612,621-622: frame -> working size; start_level 0 (optical_flow_calculator.cpp:423) makes it a copy
        if (s->width != d->width || s->height != d->height) abort();
        for (int y = 0; y < s->height; y++) memcpy(d->imageData + (size_t)y * d->widthStep, s->imageData + (size_t)y * s->widthStep, s->width);
        return;
    }
    std::vector<float> a = pack(s), b((size_t)d->width * d->height);
    orc_resize_linear_f32(a.data(), s->width, s->height, b.data(), d->width, d->height);
    unpack(b, d);
}
extern "C" void cvConvertScale(const CvArr *src, CvArr *dst, double scale, double shift)
{
    const IplImage *s = I(src);
    IplImage *d = I(dst);
    for (int y = 0; y < s->height; y++) {
        float *o = row(d, y);
        if (s->depth == IPL_DEPTH_8U) {
            const unsigned char *p = (const unsigned char *)(s->imageData + (size_t)y * s->widthStep);
            for (int x = 0; x < s->width; x++) o[x] = (scale == 1 && shift == 0) ? (float)p[x] : (float)(p[x] * scale + shift);
        } else {
            const float *p = row(s, y);
            for (int x = 0; x < s->width; x++) o[x] = (scale == 1 && shift == 0) ? p[x] : (float)(p[x] * scale + shift);
        }
    }
}
extern "C" void cvSmooth(const CvArr *src, CvArr *dst, int, int, int, double sigma1, double)
{
    std::vector<float> a = pack(I(src)), b(a.size());
    orc_gaussian_blur_f32(a.data(), I(src)->width, I(src)->height, b.data(), sigma1);
    unpack(b, I(dst));
}
// correlation, anchor = kernel centre, BORDER_REPLICATE, taps in row-major kernel order, zero coefficients skipped (cv::Filter2D
// keeps only the non-zero taps)
extern "C" void cvFilter2D(const CvArr *src, CvArr *dst, const CvMat *k, CvPoint)
{
    const IplImage *s = I(src);
    IplImage *d = I(dst);
    const int w = s->width, h = s->height, ax = k->cols / 2, ay = k->rows / 2;
    std::vector<float> a = pack(s), b(a.size());
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++) {
            float acc = 0;
            for (int j = 0; j < k->rows; j++)
                for (int i = 0; i < k->cols; i++) {
                    const float c = k->data.fl[j * k->cols + i];
                    if (c == 0.f) continue;
                    int xx = x + i - ax, yy = y + j - ay;
                    xx = xx < 0 ? 0 : (xx > w - 1 ? w - 1 : xx);
                    yy = yy < 0 ? 0 : (yy > h - 1 ? h - 1 : yy);
                    acc += c * a[(size_t)yy * w + xx];
                }
            b[(size_t)y * w + x] = acc;
        }
    unpack(b, d);
}
#define ELEMENTWISE(expr)                                                        \
    for (int y = 0; y < I(dst)->height; y++) {                                   \
        const float *pa = row(I(a), y), *pb = row(I(b), y);                      \
        float *po = row(I(dst), y);                                              \
        for (int x = 0; x < I(dst)->width; x++) po[x] = (expr);                  \
    }
extern "C" void cvSub(const CvArr *a, const CvArr *b, CvArr *dst, const CvArr *) { ELEMENTWISE(pa[x] - pb[x]) }
extern "C" void cvAdd(const CvArr *a, const CvArr *b, CvArr *dst, const CvArr *) { ELEMENTWISE(pa[x] + pb[x]) }
extern "C" void cvMul(const CvArr *a, const CvArr *b, CvArr *dst, double scale)
{
    if (scale == 1) { ELEMENTWISE(pa[x] * pb[x]) }
    else { const float sc = (float)scale; ELEMENTWISE(sc * pa[x] * pb[x]) }
}
// cv::addWeighted on 32F: dst = a * alpha + b * beta + gamma with the scalars converted to float
extern "C" void cvAddWeighted(const CvArr *a, double alpha, const CvArr *b, double beta, double gamma, CvArr *dst)
{
    const float fa = (float)alpha, fb = (float)beta, fg = (float)gamma;
    ELEMENTWISE(pa[x] * fa + pb[x] * fb + fg)
}

// ---- the entry point tests call: the reference's VarFlow class, constructed and run as OpticalFlowCalculator::varFlow does
// (common/src/optical_flow_calculator.cpp:422-452) ------------------------------------------------------------------------------
extern "C" int ref_varflow(const uint8_t *A, const uint8_t *B, int w, int h, int pitch, int max_level, int start_level, int n1, int n2,
                           float rho, float alpha, float sigma, float *U, float *V)
{
    IplImage *a = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 1), *b = cvCreateImage(cvSize(w, h), IPL_DEPTH_8U, 1);
    IplImage *u = cvCreateImage(cvSize(w, h), IPL_DEPTH_32F, 1), *v = cvCreateImage(cvSize(w, h), IPL_DEPTH_32F, 1);
    for (int y = 0; y < h; y++) {
        memcpy(a->imageData + (size_t)y * a->widthStep, A + (size_t)y * pitch, w);
        memcpy(b->imageData + (size_t)y * b->widthStep, B + (size_t)y * pitch, w);
    }
    cvZero(u); cvZero(v);
    int rc;
    {
        VarFlow vf(w, h, max_level, start_level, n1, n2, rho, alpha, sigma);
        rc = vf.CalcFlow(a, b, u, v, 0);
    }
    for (int y = 0; y < h; y++) {
        memcpy(U + (size_t)y * w, row(u, y), sizeof(float) * w);
        memcpy(V + (size_t)y * w, row(v, y), sizeof(float) * w);
    }
    cvReleaseImage(&a); cvReleaseImage(&b); cvReleaseImage(&u); cvReleaseImage(&v);
    return rc;
}
