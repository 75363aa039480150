// adapter.cpp -- the reference's common/ classes re-expressed over the C ABI of libmotion_b200.so.
// cv::Mat in / out exactly like the reference; errors of the ABI are mapped to "no vectors / empty outputs" so the
// node keeps running (the reference has no error convention of its own).
#include <motion_detection/optical_flow_calculator.h>
#include <motion_detection/outlier_detector.h>
#include <motion_detection/VarFlow.h>
#include <motion_detection/flow_clusterer.h>

#include <cmath>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <ctime>
#include <fstream>
#include <limits>

#include "motion_b200.h"

// ---------------------------------------------------------------------------------------------------------------
OpticalFlowCalculator::OpticalFlowCalculator()
    : ctx_(0), ctx2_(0), w2_(0), h2_(0), ps2_(0), w_(0), h_(0), ps_(0), batch_(0), device_(0), ego_mode_(MD_EGO_RANSAC_HOMOGRAPHY), morph_(false), seed_(1), minvec_(0)
{
    for (int i = 0; i < 9; i++) last_H_[i] = (i % 4 == 0) ? 1.0 : 0.0;
}

OpticalFlowCalculator::~OpticalFlowCalculator() { release(); }

void OpticalFlowCalculator::release()
{
    if (ctx_) md_destroy(ctx_);
    if (ctx2_) md_destroy(ctx2_);
    ctx_ = 0; ctx2_ = 0;
}

const char *OpticalFlowCalculator::lastError() const { return md_last_error(ctx_); }

bool OpticalFlowCalculator::ensure(int w, int h, int pixel_step, double min_vector_size, int max_batch)
{
    if (ctx_ && w == w_ && h == h_ && pixel_step == ps_ && min_vector_size == minvec_ && max_batch <= batch_) return true;
    release();
    md_config cfg;
    md_config_default(&cfg);
    cfg.width = w; cfg.height = h; cfg.pixel_step = pixel_step; cfg.min_vector_size = min_vector_size;
    cfg.max_batch = max_batch; cfg.ego_mode = ego_mode_; cfg.morph = morph_ ? 1 : 0; cfg.seed = seed_;
    if (md_create(&cfg, device_, &ctx_) != MD_OK) { ctx_ = 0; return false; }
    w_ = w; h_ = h; ps_ = pixel_step; minvec_ = min_vector_size; batch_ = max_batch;
    return true;
}

// The node allocates the flow field as CV_32FC4 but every accessor uses at<cv::Vec4d> (node.cpp:81,98 vs
// optical_flow_calculator.cpp:88): re-create it as CV_64FC4 so those reads are in bounds.
void OpticalFlowCalculator::ensureFlowMat(cv::Mat &m, int rows, int cols)
{
    if (m.rows != rows || m.cols != cols || m.type() != CV_64FC4) {
        m.create(rows, cols, CV_64FC4);
        std::memset(m.data, 0, m.step * (size_t)rows);
    }
}

static void write_flow(cv::Mat &flow, const float *p1, const float *p2, const uint8_t *status, const uint8_t *keep, int n, int &num_vectors)
{
    for (int i = 0; i < n; i++) {
        if (status[i]) {
            const int x = (int)p1[2 * i], y = (int)p1[2 * i + 1];
            if (x < 0 || y < 0 || x >= flow.cols || y >= flow.rows) continue;
            cv::Vec4d &e = flow.at<cv::Vec4d>(y, x);
            e[0] = p1[2 * i]; e[1] = p1[2 * i + 1];
            if (keep[i]) { e[2] = p2[2 * i] - p1[2 * i]; e[3] = p2[2 * i + 1] - p1[2 * i + 1]; num_vectors++; }
            else { e[2] = 0.0; e[3] = 0.0; }
        } else {
            const int x = (int)p1[2 * i], y = (int)p1[2 * i + 1];
            if (x < 0 || y < 0 || x >= flow.cols || y >= flow.rows) continue;
            cv::Vec4d &e = flow.at<cv::Vec4d>(y, x);
            e[0] = -1.0; e[1] = -1.0; e[2] = 0.0; e[3] = 0.0;
        }
    }
}

int OpticalFlowCalculator::calculateOpticalFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow_vectors,
                                                int pixel_step, cv::Mat &comp, double min_vector_size)
{
    const int w = image1.cols, h = image1.rows, ch = image1.channels();
    if (image1.empty() || image2.empty() || image2.cols != w || image2.rows != h || image2.channels() != ch || (ch != 1 && ch != 3)) return 0;
    if (!ensure(w, h, pixel_step, min_vector_size, 1)) return 0;
    const int P = md_grid_size(ctx_);
    std::vector<float> pts(2 * (size_t)P), nxt(2 * (size_t)P);
    std::vector<uint8_t> st(P), keep(P);
    cv::Mat mask(h, w, CV_8UC1);
    int nv = 0, inl = 0;
    md_outputs out;
    std::memset(&out, 0, sizeof out);
    out.next_pts = nxt.data(); out.status = st.data(); out.keep = keep.data(); out.H = last_H_;
    out.num_vectors = &nv; out.inliers = &inl; out.mask = mask.data; out.mask_pitch = (int)mask.step; out.mask_stride = 0;
    // md_process_pair takes one pitch for both frames: re-pack the second frame when the two Mats disagree
    std::vector<uint8_t> repack;
    const uint8_t *cur = image2.data;
    if (image2.step != image1.step) {
        repack.resize(image1.step * (size_t)h);
        for (int y = 0; y < h; y++) std::memcpy(&repack[(size_t)y * image1.step], image2.data + (size_t)y * image2.step, (size_t)w * ch);
        cur = repack.data();
    }
    if (md_process_pair(ctx_, image1.data, cur, ch, (int)image1.step, &out, MD_MEM_HOST) != MD_OK) return 0;
    md_grid_points(ctx_, pts.data());
    ensureFlowMat(optical_flow_vectors, h, w);
    int counted = 0;
    write_flow(optical_flow_vectors, pts.data(), nxt.data(), st.data(), keep.data(), P, counted);
    if (nv > 0) mask.copyTo(comp);          // the reference leaves comp untouched when there are no vectors (cpp:118)
    return nv;
}

int OpticalFlowCalculator::calculateOpticalFlowTrajectory(const std::vector<cv::Mat> &images, cv::Mat &optical_flow_vectors,
                                                          std::vector<std::vector<cv::Point2f> > &trajectories, int pixel_step,
                                                          cv::Mat &comp, double min_vector_size)
{
    (void)comp;      // untouched by the reference as well
    const int F = (int)images.size();
    if (F < 2) return 0;
    const int w = images[0].cols, h = images[0].rows, ch = images[0].channels();
    for (int i = 0; i < F; i++)
        if (images[i].empty() || images[i].cols != w || images[i].rows != h || images[i].channels() != ch) return 0;
    if (!ensure(w, h, pixel_step, min_vector_size, F - 1)) return 0;
    const int P = md_grid_size(ctx_);
    // md_frames wants one base + stride: copy the window into one contiguous block (F small frames, host side)
    const size_t fbytes = (size_t)w * ch * h;
    std::vector<uint8_t> block(fbytes * F);
    for (int i = 0; i < F; i++)
        for (int y = 0; y < h; y++) std::memcpy(&block[i * fbytes + (size_t)y * w * ch], images[i].data + (size_t)y * images[i].step, (size_t)w * ch);
    md_frames fr;
    std::memset(&fr, 0, sizeof fr);
    fr.data = block.data(); fr.channels = ch; fr.pitch = w * ch; fr.frame_stride = (int64_t)fbytes; fr.count = F; fr.chain = 0;
    std::vector<float> traj(2 * (size_t)P * F), lp(2 * (size_t)P), ln(2 * (size_t)P);
    std::vector<int32_t> len(P);
    std::vector<uint8_t> ls(P), keep(P);
    if (md_track_trajectories(ctx_, &fr, traj.data(), len.data(), lp.data(), ln.data(), ls.data(), MD_MEM_HOST) != MD_OK) return 0;
    ensureFlowMat(optical_flow_vectors, h, w);
    for (int i = 0; i < P; i++) {
        const float xd = ln[2 * i] - lp[2 * i], yd = ln[2 * i + 1] - lp[2 * i + 1];
        keep[i] = ls[i] && (std::abs(xd) > min_vector_size || std::abs(yd) > min_vector_size);      // cpp:189
    }
    int num_vectors = 0;
    write_flow(optical_flow_vectors, lp.data(), ln.data(), ls.data(), keep.data(), P, num_vectors);
    for (int i = 0; i < P; i++) {
        if (len[i] != F) continue;                                                                   // cpp:246
        std::vector<cv::Point2f> t(F);
        for (int f = 0; f < F; f++) t[f] = cv::Point2f(traj[((size_t)i * F + f) * 2], traj[((size_t)i * F + f) * 2 + 1]);
        trajectories.push_back(t);
    }
    return num_vectors;
}

// The grid LK alone (cpp:259-330): MAX_LEVEL = 2, vectors kept when |dx| or |dy| exceeds 1 px.  (The reference falls off the
// end of this non-void function; the count it computes is returned here.)
int OpticalFlowCalculator::calculateCompensatedFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow_vectors, int pixel_step)
{
    const int w = image1.cols, h = image1.rows, ch = image1.channels();
    if (image1.empty() || image2.empty() || image2.cols != w || image2.rows != h || image2.channels() != ch || (ch != 1 && ch != 3) || pixel_step < 1)
        return 0;
    if (!ctx2_ || w != w2_ || h != h2_ || pixel_step != ps2_) {
        if (ctx2_) md_destroy(ctx2_);
        ctx2_ = 0;
        md_config cfg;
        md_config_default(&cfg);
        cfg.width = w; cfg.height = h; cfg.pixel_step = pixel_step; cfg.max_batch = 1;
        cfg.lk_max_level = 2;                                   // cpp:261
        cfg.min_vector_size = 1.0;                              // cpp:301
        if (md_create(&cfg, device_, &ctx2_) != MD_OK) { ctx2_ = 0; return 0; }
        w2_ = w; h2_ = h; ps2_ = pixel_step;
    }
    std::vector<uint8_t> g1((size_t)w * h), g2((size_t)w * h);
    const cv::Mat *im[2] = {&image1, &image2};
    uint8_t *g[2] = {g1.data(), g2.data()};
    for (int k = 0; k < 2; k++) {
        if (ch == 3) {
            if (md_gray_u8(ctx2_, im[k]->data, (int32_t)im[k]->step, w, h, g[k], w, MD_MEM_HOST) != MD_OK) return 0;   // cpp:270-271
        } else {
            for (int y = 0; y < h; y++) std::memcpy(g[k] + (size_t)y * w, im[k]->data + (size_t)y * im[k]->step, (size_t)w);
        }
        if (md_pyramid_u8(ctx2_, g[k], w, k, MD_MEM_HOST) != MD_OK) return 0;
    }
    const int P = md_grid_size(ctx2_);
    std::vector<float> p1(2 * (size_t)P), p2(2 * (size_t)P);
    std::vector<uint8_t> st(P), keep(P);
    md_grid_points(ctx2_, p1.data());
    if (md_lk_flow(ctx2_, 0, 1, 0, P, p2.data(), st.data(), MD_MEM_HOST) != MD_OK) return 0;
    for (int i = 0; i < P; i++) {
        const float xd = p2[2 * i] - p1[2 * i], yd = p2[2 * i + 1] - p1[2 * i + 1];
        keep[i] = st[i] && (std::abs(xd) > 1.0 || std::abs(yd) > 1.0);
    }
    ensureFlowMat(optical_flow_vectors, h, w);
    int num_vectors = 0;
    write_flow(optical_flow_vectors, p1.data(), p2.data(), st.data(), keep.data(), P, num_vectors);
    return num_vectors;
}

void OpticalFlowCalculator::varFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow, cv::Mat &optical_flow_vectors)
{
    // The reference computes U,V and only draws arrows into `optical_flow` (cpp:455-463); here the dense field itself is
    // returned: optical_flow = 32FC1 U (+x), optical_flow_vectors = 32FC1 V (y-up, VarFlow.cpp:103-107).
    const int w = image1.cols, h = image1.rows;
    if (image1.empty() || image1.channels() != 1 || image2.cols != w || image2.rows != h || image2.channels() != 1) return;
    if (!ensure(w, h, ps_ > 0 ? ps_ : 10, minvec_, batch_ > 0 ? batch_ : 1)) return;
    optical_flow.create(h, w, CV_32FC1);
    optical_flow_vectors.create(h, w, CV_32FC1);
    std::vector<uint8_t> a((size_t)w * h), b((size_t)w * h);
    for (int y = 0; y < h; y++) {
        std::memcpy(&a[(size_t)y * w], image1.data + (size_t)y * image1.step, w);
        std::memcpy(&b[(size_t)y * w], image2.data + (size_t)y * image2.step, w);
    }
    md_varflow(ctx_, a.data(), b.data(), w, reinterpret_cast<float *>(optical_flow.data), reinterpret_cast<float *>(optical_flow_vectors.data), MD_MEM_HOST);
}

#ifndef MD_ADAPTER_EXTERNAL_SUPERPIXELFLOW
// cpp:337-416 -- SLIC superpixels + LK per superpixel centre.  Nothing in the reference calls it and it depends on the
// vendored slic.cpp; not on the accelerated path: no vectors, outputs untouched (see the header note).
int OpticalFlowCalculator::superPixelFlow(const cv::Mat &, const cv::Mat &, cv::Mat &, cv::Mat &) { return 0; }
#endif

// cpp:467-507: for every (xSpace, ySpace) grid node an arrow along (U, -V) scaled by `multiplier`, when longer than `cutoff`;
// shaft plus two 3-pixel barbs at +-45 degrees from the reversed direction.
void OpticalFlowCalculator::drawMotionField(IplImage *imgU, IplImage *imgV, IplImage *imgMotion, int xSpace, int ySpace, float cutoff,
                                            int multiplier, CvScalar color)
{
    if (!imgU || !imgV || !imgMotion || xSpace < 1 || ySpace < 1) return;
    const double pi = 3.14159265358979323846;
    for (int y = ySpace; y < imgU->height; y += ySpace) {
        const float *urow = reinterpret_cast<const float *>(imgU->imageData + (size_t)y * imgU->widthStep);
        const float *vrow = reinterpret_cast<const float *>(imgV->imageData + (size_t)y * imgV->widthStep);
        for (int x = xSpace; x < imgU->width; x += xSpace) {
            const float du = urow[x], dv = -vrow[x];                      // V is y-up
            const float len = std::sqrt(du * du + dv * dv);
            if (!(len > cutoff)) continue;
            const float dir = std::atan2(dv, du);
            const CvPoint tail = cvPoint(x, y);
            const CvPoint tip = cvPoint(x + (int)lrint(multiplier * len * std::cos(dir)), y + (int)lrint(multiplier * len * std::sin(dir)));
            cvLine(imgMotion, tail, tip, color, 1, CV_AA, 0);
            for (int side = -1; side <= 1; side += 2) {
                const double a = dir - pi - side * (pi / 4);
                const CvPoint barb = cvPoint(tip.x + (int)lrint(3 * std::cos(a)), tip.y + (int)lrint(3 * std::sin(a)));
                cvLine(imgMotion, barb, tip, color, 1, CV_AA, 0);
            }
        }
    }
}

// cpp:509-541: two text files, "<filename>_h" (dx) and "<filename>_f" (dy); one line per grid row (rows and columns in steps
// of pixel_step), values separated by ", ", failed vectors (x == -1, cpp:112-115) written as 0, default ostream formatting.
void OpticalFlowCalculator::writeFlow(const cv::Mat &flow_vectors, const std::string &filename, int pixel_step)
{
    if (pixel_step < 1 || flow_vectors.type() != CV_64FC4) return;
    std::ofstream fh((filename + "_h").c_str()), fv((filename + "_f").c_str());
    for (int r = 0; r < flow_vectors.rows; r += pixel_step) {
        for (int c = 0; c < flow_vectors.cols; c += pixel_step) {
            const cv::Vec4d &e = flow_vectors.at<cv::Vec4d>(r, c);
            const bool failed = e[0] == -1.0;
            if (c) { fh << ", "; fv << ", "; }
            fh << (failed ? 0.0 : e[2]);
            fv << (failed ? 0.0 : e[3]);
        }
        fh << std::endl;
        fv << std::endl;
    }
}

// cpp:543-562: one line per trajectory, "x0, y0, x1, y1, ..."
void OpticalFlowCalculator::writeTrajectories(const std::vector<std::vector<cv::Point2f> > &trajectories, const std::string &filename)
{
    std::ofstream ft(filename.c_str());
    for (size_t i = 0; i < trajectories.size(); i++) {
        const std::vector<cv::Point2f> &t = trajectories[i];
        for (size_t j = 0; j < t.size(); j++) ft << (j ? ", " : "") << t[j].x << ", " << t[j].y;
        ft << std::endl;
    }
}

// ---------------------------------------------------------------------------------------------------------------
OutlierDetector::OutlierDetector() : ctx_(0), device_(0), seed_((unsigned)time(NULL)), last_inliers_(0) {}

OutlierDetector::~OutlierDetector() { if (ctx_) md_destroy(ctx_); }

std::vector<std::vector<cv::Point2f> > OutlierDetector::fitSubspace(const std::vector<std::vector<cv::Point2f> > &trajectories,
                                                                     std::vector<cv::Point2f> &outlier_points, int num_motions, double sigma)
{
    std::vector<std::vector<cv::Point2f> > basis;
    const int T = (int)trajectories.size();
    if (T < 1) return basis;                       // the reference dereferences trajectories[0] unguarded (:239)
    const int F = (int)trajectories[0].size();
    if (!ensureContext()) return basis;
    std::vector<float> traj(2 * (size_t)T * F);
    for (int i = 0; i < T; i++)
        for (int f = 0; f < F; f++) { traj[((size_t)i * F + f) * 2] = trajectories[i][f].x; traj[((size_t)i * F + f) * 2 + 1] = trajectories[i][f].y; }
    const int d = 4 * num_motions;
    std::vector<float> res(T);
    std::vector<int32_t> cols(d > 0 ? d : 1);
    std::vector<uint8_t> outl(T);
    int32_t ninl = 0;
    // the reference's rand() stream continues across calls; here every call advances the seed by one
    if (md_fit_subspace(ctx_, traj.data(), T, F, num_motions, sigma, seed_++, 0, 50, res.data(), cols.data(), outl.data(), &ninl, MD_MEM_HOST) != MD_OK)
        return basis;
    last_inliers_ = ninl;
    for (int i = 0; i < T; i++)
        if (outl[i]) outlier_points.push_back(trajectories[i][F - 2]);       // second-to-last point, :322
    for (int k = 0; k < d; k++)
        if (cols[k] >= 0) basis.push_back(trajectories[cols[k]]);           // :325-330
    return basis;
}

// ---------------------------------------------------------------------------------------------------------------
VarFlow::VarFlow(int width_in, int height_in, int max_level_in, int start_level_in, int n1_in, int n2_in, float rho_in,
                 float alpha_in, float sigma_in)
    : ctx_(0), width(width_in), height(height_in), initialized(0)
{
    md_config cfg;
    md_config_default(&cfg);
    cfg.width = width_in; cfg.height = height_in;
    cfg.vf_max_level = max_level_in < start_level_in ? start_level_in : max_level_in;      // VarFlow.cpp:33-37
    cfg.vf_start_level = start_level_in; cfg.vf_n1 = n1_in; cfg.vf_n2 = n2_in;
    cfg.vf_rho = rho_in; cfg.vf_alpha = alpha_in; cfg.vf_sigma = sigma_in;
    if (md_create(&cfg, default_device_, &ctx_) == MD_OK) initialized = 1;
    else ctx_ = 0;
}

int VarFlow::default_device_ = 0;

VarFlow::~VarFlow() { if (ctx_) md_destroy(ctx_); }

int VarFlow::CalcFlow(IplImage *imgA, IplImage *imgB, IplImage *imgU, IplImage *imgV, bool saved_data)
{
    (void)saved_data;        // only a caching hint in the reference (VarFlow.cpp:608-616); results are the same
    if (!initialized || !imgA || !imgB || !imgU || !imgV) return 0;
    if (imgA->width != width || imgA->height != height || imgA->nChannels != 1 || imgB->width != width || imgB->height != height ||
        imgA->widthStep != imgB->widthStep || imgU->width != width || imgU->height != height)
        return 0;
    std::vector<float> U((size_t)width * height), V((size_t)width * height);
    if (md_varflow(ctx_, reinterpret_cast<const uint8_t *>(imgA->imageData), reinterpret_cast<const uint8_t *>(imgB->imageData),
                   imgA->widthStep, U.data(), V.data(), MD_MEM_HOST) != MD_OK)
        return 0;
    for (int y = 0; y < height; y++) {
        std::memcpy(imgU->imageData + (size_t)y * imgU->widthStep, &U[(size_t)y * width], sizeof(float) * width);
        std::memcpy(imgV->imageData + (size_t)y * imgV->widthStep, &V[(size_t)y * width], sizeof(float) * width);
    }
    return 1;
}

bool OutlierDetector::ensureContext()
{
    if (ctx_) return true;
    md_config cfg;
    md_config_default(&cfg);
    cfg.width = 64; cfg.height = 64;               // geometry is irrelevant for the subspace fit and the MAD test
    if (md_create(&cfg, device_, &ctx_) != MD_OK) { ctx_ = 0; return false; }
    return true;
}

void OutlierDetector::findOutliers(const cv::Mat &optical_flow_vectors, cv::Mat &outlier_probabilities, bool include_zeros,
                                   int pixel_step, bool print)
{
    (void)print;
    const int rows = optical_flow_vectors.rows, cols = optical_flow_vectors.cols;
    outlier_probabilities = cv::Mat::zeros(rows, cols, CV_64F);                              // :41
    if (rows < 1 || cols < 1 || pixel_step < 1 || optical_flow_vectors.type() != CV_64FC4 || !ensureContext()) return;
    std::vector<double> dxdy;
    for (int i = 0; i < rows; i += pixel_step)                                               // the reference's traversal, :77-84
        for (int j = 0; j < cols; j += pixel_step) {
            const cv::Vec4d &e = optical_flow_vectors.at<cv::Vec4d>(i, j);
            dxdy.push_back(e[2]); dxdy.push_back(e[3]);
        }
    const int n = (int)(dxdy.size() / 2);
    std::vector<uint8_t> flag(n);
    if (md_find_outliers(ctx_, dxdy.data(), n, include_zeros ? 1 : 0, flag.data(), 0, MD_MEM_HOST) != MD_OK) return;
    int k = 0;
    for (int i = 0; i < rows; i += pixel_step)
        for (int j = 0; j < cols; j += pixel_step, k++)
            if (flag[k]) outlier_probabilities.at<double>(i, j) = 1.0;
}

void OutlierDetector::getOutlierVectors(const cv::Mat &optical_flow_vectors, const cv::Mat &outlier_probabilities, cv::Mat &outlier_vectors,
                                        int pixel_step)
{
    outlier_vectors = cv::Mat::zeros(optical_flow_vectors.rows, optical_flow_vectors.cols, CV_64FC4);
    if (optical_flow_vectors.type() != CV_64FC4 || pixel_step < 1) return;
    for (int i = 0; i < optical_flow_vectors.rows; i += pixel_step)
        for (int j = 0; j < optical_flow_vectors.cols; j += pixel_step)
            if (outlier_probabilities.at<double>(i, j) > 0.5) outlier_vectors.at<cv::Vec4d>(i, j) = optical_flow_vectors.at<cv::Vec4d>(i, j);
}

// ---------------------------------------------------------------------------------------------------------------
FlowClusterer::FlowClusterer() : ctx_(0), device_(0) {}

FlowClusterer::~FlowClusterer() { if (ctx_) md_destroy(ctx_); }

bool FlowClusterer::ensureContext()
{
    if (ctx_) return true;
    md_config cfg;
    md_config_default(&cfg);
    cfg.width = 64; cfg.height = 64;               // geometry is irrelevant for the grouping
    if (md_create(&cfg, device_, &ctx_) != MD_OK) { ctx_ = 0; return false; }
    return true;
}

cv::Mat FlowClusterer::clusterFlowVectors(const cv::Mat &flow_vectors)
{
    (void)flow_vectors;                            // FLANN k-means with a rand()-seeded start: no call site, not reproducible
    return cv::Mat(0, 2, CV_32F);
}

// the vectors the reference's loops visit (flow_clusterer.cpp:180-185): rows outer, columns inner, non-zero flow only
static void gather_vectors(const cv::Mat &flow_vectors, int pixel_step, std::vector<cv::Vec4d> &out)
{
    if (pixel_step < 1 || flow_vectors.type() != CV_64FC4) return;
    for (int r = 0; r < flow_vectors.rows; r += pixel_step)
        for (int c = 0; c < flow_vectors.cols; c += pixel_step) {
            const cv::Vec4d &e = flow_vectors.at<cv::Vec4d>(r, c);
            if (std::fabs(e[2]) > 0.0 || std::fabs(e[3]) > 0.0) out.push_back(e);
        }
}

std::vector<std::vector<cv::Vec4d> > FlowClusterer::getClusters(const cv::Mat &flow_vectors, int pixel_step, double distance_threshold,
                                                                  double angular_threshold)
{
    std::vector<std::vector<cv::Vec4d> > mat_clusters;
    std::vector<cv::Vec4d> vec;
    gather_vectors(flow_vectors, pixel_step, vec);
    const int n = (int)vec.size();
    if (n < 1 || !ensureContext()) return mat_clusters;
    std::vector<int32_t> labels(n);
    int32_t nall = 0;
    if (md_cluster_vectors(ctx_, &vec[0][0], n, distance_threshold, angular_threshold, labels.data(), &nall, MD_MEM_HOST) != MD_OK)
        return mat_clusters;
    // clusters with more than 5 members in creation order, members in arrival order (:210-217)
    std::vector<int> size(nall > 0 ? nall : 1, 0), slot(nall > 0 ? nall : 1, -1);
    for (int i = 0; i < n; i++) size[labels[i]]++;
    int k = 0;
    for (int c = 0; c < nall; c++)
        if (size[c] > 5) slot[c] = k++;
    mat_clusters.resize(k);
    for (int c = 0; c < nall; c++)
        if (slot[c] >= 0) mat_clusters[slot[c]].reserve(size[c]);
    for (int i = 0; i < n; i++)
        if (slot[labels[i]] >= 0) mat_clusters[slot[labels[i]]].push_back(vec[i]);
    return mat_clusters;
}

// flow_clusterer.cpp:80-115 (no call site): every qualifying cluster takes the vector; centroids of all clusters.
std::vector<cv::Point2f> FlowClusterer::getClustersCenters(const cv::Mat &flow_vectors, int pixel_step, double distance_threshold,
                                                           double angular_threshold)
{
    std::vector<cv::Vec4d> vec;
    gather_vectors(flow_vectors, pixel_step, vec);
    const double two_pi = 2 * 3.14159265358979323846;
    std::vector<double> ang(vec.size());
    for (size_t i = 0; i < vec.size(); i++) { ang[i] = std::atan2(vec[i][3], vec[i][2]); if (ang[i] < 0.0) ang[i] += two_pi; }
    std::vector<std::vector<int> > members;
    for (size_t i = 0; i < vec.size(); i++) {
        bool added = false;
        for (size_t c = 0; c < members.size(); c++) {
            double dmin = std::numeric_limits<double>::max(), amin = dmin;
            for (size_t m = 0; m < members[c].size(); m++) {
                const int j = members[c][m];
                const double dx = vec[i][0] - vec[j][0], dy = vec[i][1] - vec[j][1], da = ang[i] - ang[j];
                dmin = std::min(dmin, std::sqrt(dx * dx + dy * dy));
                amin = std::min(amin, std::fabs(std::atan2(std::sin(da), std::cos(da))));
            }
            if (dmin < distance_threshold && amin < angular_threshold) { members[c].push_back((int)i); added = true; }
        }
        if (!added) members.push_back(std::vector<int>(1, (int)i));
    }
    std::vector<cv::Point2f> centroids;
    for (size_t c = 0; c < members.size(); c++) {
        double sx = 0.0, sy = 0.0;
        for (size_t m = 0; m < members[c].size(); m++) { sx += vec[members[c][m]][0]; sy += vec[members[c][m]][1]; }
        centroids.push_back(cv::Point2f((float)(sx / members[c].size()), (float)(sy / members[c].size())));
    }
    return centroids;
}

std::vector<std::vector<cv::Point2f> > FlowClusterer::clusterEuclidean(const std::vector<cv::Point2f> &points, double distance_threshold)
{
    std::vector<std::vector<cv::Point2f> > mat_clusters;
    boxes_.clear();
    const int n = (int)points.size();
    if (n < 1 || !ensureContext()) return mat_clusters;
    std::vector<int32_t> labels(n), boxes(4 * (size_t)n), sizes(n), ids(n);
    int32_t nall = 0, k = 0;
    if (md_cluster_points(ctx_, &points[0].x, n, distance_threshold, 5, labels.data(), &nall, &k, boxes.data(), sizes.data(), ids.data(),
                          MD_MEM_HOST) != MD_OK)
        return mat_clusters;
    // clusters with more than 5 points in creation order, members in arrival order (flow_clusterer.cpp:262-267)
    std::vector<int> slot(nall > 0 ? nall : 1, -1);
    mat_clusters.resize(k);
    for (int c = 0; c < k; c++) {
        slot[ids[c]] = c;
        mat_clusters[c].reserve(sizes[c]);
        cv::Vec4i b;
        for (int q = 0; q < 4; q++) b[q] = boxes[4 * c + q];
        boxes_.push_back(b);
    }
    for (int i = 0; i < n; i++)
        if (slot[labels[i]] >= 0) mat_clusters[slot[labels[i]]].push_back(points[i]);
    return mat_clusters;
}
