// ofc_shim.cpp -- implementation of the cv:: shim of this directory on the oracle's cv2-pinned primitives (test infrastructure
// only; nothing here is reference code) + the C entry points into the reference's own OpticalFlowCalculator.
#include <stdint.h>
#include <stdio.h>

#include <opencv2/video/tracking.hpp>
#include <opencv2/imgproc/imgproc.hpp>
#include <opencv/cv.h>

#include <motion_detection/optical_flow_calculator.h>      // the REFERENCE's header
#include <motion_detection/slic.h>                         // the REFERENCE's header (superPixelFlow holds a Slic by value)

extern "C" {
#include "../md_oracle.h"
}

static void off_path(const char *what)
{
    fprintf(stderr, "oracle/ref_shim_ofc: %s is not on the tested path (superPixelFlow / arrow drawing)\n", what);
    abort();
}

namespace cv {
void absdiff(const Mat &a, const Mat &b, Mat &dst)
{
    Mat out(a.rows, a.cols, a.type());
    const size_t n = (size_t)a.cols * Mat::esz(a.type());
    for (int y = 0; y < a.rows; y++) {
        const uchar *pa = a.data + (size_t)y * a.step, *pb = b.data + (size_t)y * b.step;
        uchar *po = out.data + (size_t)y * out.step;
        for (size_t x = 0; x < n; x++) po[x] = (uchar)(pa[x] > pb[x] ? pa[x] - pb[x] : pb[x] - pa[x]);
    }
    dst = out;
}
double threshold(const Mat &src, Mat &dst, double thresh, double maxval, int type)
{
    if (type != CV_THRESH_BINARY) off_path("threshold type");
    Mat out(src.rows, src.cols, src.type());
    const int t = (int)floor(thresh);                           // 8-bit THRESH_BINARY: src > floor(thresh)
    for (int y = 0; y < src.rows; y++)
        for (int x = 0; x < src.cols; x++) out.data[(size_t)y * out.step + x] = src.data[(size_t)y * src.step + x] > t ? (uchar)maxval : 0;
    dst = out;
    return thresh;
}
void cvtColor(const Mat &src, Mat &dst, int code)
{
    if (code != CV_BGR2GRAY) off_path("cvtColor code");
    Mat out(src.rows, src.cols, CV_8UC1);
    if (src.channels() == 3) orc_gray_bgr2gray(src.data, src.cols, src.rows, (int)src.step, out.data, (int)out.step);
    else off_path("cvtColor of a non-BGR image (OpenCV asserts scn == 3 || scn == 4)");
    dst = out;
}
int buildOpticalFlowPyramid(const Mat &img, std::vector<Mat> &pyramid, Size winSize, int maxLevel, bool)
{
    pyramid.clear();
    pyramid.push_back(img.clone());
    return orc_pyr_levels(img.cols, img.rows, winSize.width, maxLevel) - 1;
}
void calcOpticalFlowPyrLK(const Mat &prevImg, const Mat &nextImg, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts,
                          std::vector<uchar> &status, std::vector<float> &err, Size winSize, int maxLevel, TermCriteria criteria, int flags,
                          double minEigThreshold)
{
    if (flags != 0 || winSize.width != winSize.height || prevImg.step != nextImg.step) off_path("calcOpticalFlowPyrLK arguments");
    const int n = (int)prevPts.size();
    nextPts.resize(n); status.resize(n); err.assign(n, 0.f);
    if (n == 0) return;
    std::vector<float> in(2 * (size_t)n), out(2 * (size_t)n);
    for (int i = 0; i < n; i++) { in[2 * i] = prevPts[i].x; in[2 * i + 1] = prevPts[i].y; }
    orc_lk_pyr(prevImg.data, nextImg.data, prevImg.cols, prevImg.rows, (int)prevImg.step, in.data(), n, out.data(), status.data(),
               winSize.width, maxLevel, criteria.maxCount, criteria.epsilon, (float)minEigThreshold);
    for (int i = 0; i < n; i++) nextPts[i] = Point2f(out[2 * i], out[2 * i + 1]);
}
void calcOpticalFlowPyrLK(const std::vector<Mat> &prevPyr, const Mat &nextImg, const std::vector<Point2f> &prevPts, std::vector<Point2f> &nextPts,
                          std::vector<uchar> &status, std::vector<float> &err, Size winSize, int maxLevel, TermCriteria criteria, int flags,
                          double minEigThreshold)
{
    calcOpticalFlowPyrLK(prevPyr.at(0), nextImg, prevPts, nextPts, status, err, winSize, maxLevel, criteria, flags, minEigThreshold);
}
Mat getPerspectiveTransform(const Point2f src[], const Point2f dst[])
{
    double s[8], d[8];
    for (int i = 0; i < 4; i++) { s[2 * i] = src[i].x; s[2 * i + 1] = src[i].y; d[2 * i] = dst[i].x; d[2 * i + 1] = dst[i].y; }
    Mat H(3, 3, CV_64FC1);
    double h9[9];
    if (!orc_perspective_4pt(s, d, h9)) for (int i = 0; i < 9; i++) h9[i] = 0.0;      // singular system: the tests keep away from it
    for (int i = 0; i < 9; i++) H.at<double>(i / 3, i % 3) = h9[i];
    return H;
}
void warpPerspective(const Mat &src, Mat &dst, const Mat &M, Size dsize)
{
    if (dsize.width != src.cols || dsize.height != src.rows) off_path("warpPerspective to another size");
    double h9[9];
    for (int i = 0; i < 9; i++) h9[i] = M.at<double>(i / 3, i % 3);
    Mat out(src.rows, src.cols, src.type());
    orc_warp_perspective(src.data, src.cols, src.rows, (int)src.step, h9, out.data, (int)out.step);
    dst = out;
}
void line(Mat &, Point2f, Point2f, const Scalar &, int, int, int) { off_path("cv::line"); }
Mat cvarrToMat(const IplImage *img)
{
    return Mat(img->height, img->width, CV_MAKETYPE(img->depth == IPL_DEPTH_32F ? CV_32F : CV_8U, img->nChannels), img->imageData, (size_t)img->widthStep);
}
}  // namespace cv

extern "C" IplImage *cvCloneImage(const IplImage *) { off_path("cvCloneImage"); return 0; }
extern "C" void cvCvtColor(const CvArr *, CvArr *, int) { off_path("cvCvtColor"); }
extern "C" void cvLine(CvArr *, CvPoint, CvPoint, CvScalar, int, int, int) { off_path("cvLine"); }
// Slic (common/src/slic.cpp needs the IplImage pixel accessors of OpenCV): superPixelFlow is uncalled in the reference
Slic::Slic() {}
Slic::~Slic() {}
void Slic::generate_superpixels(IplImage *, int, int) { off_path("Slic::generate_superpixels"); }
vec2dd Slic::get_centers() { off_path("Slic::get_centers"); return vec2dd(); }

// ---- C entry points ---------------------------------------------------------------------------------------------------------------
static cv::Mat wrap_u8(const uint8_t *p, int w, int h, int channels)
{
    cv::Mat m(h, w, channels == 3 ? CV_8UC3 : CV_8UC1);
    for (int y = 0; y < h; y++) memcpy(m.data + (size_t)y * m.step, p + (size_t)y * w * channels, (size_t)w * channels);
    return m;
}
// OpticalFlowCalculator::calculateOpticalFlow (optical_flow_calculator.cpp:30-130) on two BGR images [h][w][3]; flow [h][w][4] f64 (the
// node's Mat, allocated CV_64FC4 so that the at<Vec4d> writes are in bounds); comp [h][w] u8, *comp_rows = rows of comp after the call
// (0: untouched, the num_vectors == 0 path)
extern "C" int ref_calculate_optical_flow(const uint8_t *im1, const uint8_t *im2, int w, int h, int pixel_step, double min_vector_size,
                                          double *flow, uint8_t *comp, int *comp_rows)
{
    cv::Mat a = wrap_u8(im1, w, h, 3), b = wrap_u8(im2, w, h, 3);
    cv::Mat f = cv::Mat::zeros(h, w, CV_64FC4), c;
    OpticalFlowCalculator ofc;
    const int nv = ofc.calculateOpticalFlow(a, b, f, pixel_step, c, min_vector_size);
    memcpy(flow, f.data, sizeof(double) * 4 * (size_t)w * h);
    *comp_rows = c.rows;
    for (int y = 0; y < c.rows; y++) memcpy(comp + (size_t)y * w, c.data + (size_t)y * c.step, (size_t)w);
    return nv;
}
// OpticalFlowCalculator::calculateOpticalFlowTrajectory (cpp:133-257) on n BGR images; trajectories [ntraj][n][2] f32 (capacity cap);
// returns num_vectors, *ntraj = complete trajectories
extern "C" int ref_calculate_trajectories(const uint8_t *imgs, int n, int w, int h, int pixel_step, double min_vector_size, double *flow,
                                          float *traj, int cap, int *ntraj)
{
    std::vector<cv::Mat> images;
    for (int i = 0; i < n; i++) images.push_back(wrap_u8(imgs + (size_t)i * w * h * 3, w, h, 3));
    cv::Mat f = cv::Mat::zeros(h, w, CV_64FC4), c;
    std::vector<std::vector<cv::Point2f> > tr;
    OpticalFlowCalculator ofc;
    const int nv = ofc.calculateOpticalFlowTrajectory(images, f, tr, pixel_step, c, min_vector_size);
    memcpy(flow, f.data, sizeof(double) * 4 * (size_t)w * h);
    *ntraj = (int)tr.size();
    for (int t = 0; t < (int)tr.size() && t < cap; t++)
        for (int j = 0; j < n; j++) { traj[((size_t)t * n + j) * 2] = tr[t][j].x; traj[((size_t)t * n + j) * 2 + 1] = tr[t][j].y; }
    return nv;
}
// OpticalFlowCalculator::calculateCompensatedFlow (cpp:264-335).  The reference's function has no return statement: a current gcc
// plants a trap (ud2) where control flows off its end, after the whole body -- loops and the destructors of its locals -- has run.
// That trap is caught here and the call abandoned at exactly that point; the flow field is complete by then.
#include <setjmp.h>
#include <signal.h>
static sigjmp_buf g_ofc_jmp;
static void ofc_sigill(int) { siglongjmp(g_ofc_jmp, 1); }
extern "C" void ref_calculate_compensated_flow(const uint8_t *im1, const uint8_t *im2, int w, int h, int pixel_step, double *flow)
{
    cv::Mat a = wrap_u8(im1, w, h, 3), b = wrap_u8(im2, w, h, 3);
    cv::Mat f = cv::Mat::zeros(h, w, CV_64FC4);
    static OpticalFlowCalculator ofc;
    struct sigaction sa, old;
    memset(&sa, 0, sizeof sa);
    sa.sa_handler = ofc_sigill;
    sigaction(SIGILL, &sa, &old);
    if (sigsetjmp(g_ofc_jmp, 1) == 0) ofc.calculateCompensatedFlow(a, b, f, pixel_step);
    sigaction(SIGILL, &old, 0);
    memcpy(flow, f.data, sizeof(double) * 4 * (size_t)w * h);
}
// writeFlow / writeTrajectories (cpp:509-562): the files the reference writes for a field [h][w][4] / trajectories [nt][F][2]
extern "C" void ref_write_flow(const double *flow, int w, int h, int pixel_step, const char *filename)
{
    cv::Mat f(h, w, CV_64FC4);
    memcpy(f.data, flow, sizeof(double) * 4 * (size_t)w * h);
    OpticalFlowCalculator ofc;
    ofc.writeFlow(f, filename, pixel_step);
}
extern "C" void ref_write_trajectories(const float *traj, int nt, int F, const char *filename)
{
    std::vector<std::vector<cv::Point2f> > tr(nt);
    for (int t = 0; t < nt; t++)
        for (int j = 0; j < F; j++) tr[t].push_back(cv::Point2f(traj[((size_t)t * F + j) * 2], traj[((size_t)t * F + j) * 2 + 1]));
    OpticalFlowCalculator ofc;
    ofc.writeTrajectories(tr, filename);
}
