/* optical_flow_calculator.h -- drop-in replacement of common/include/motion_detection/optical_flow_calculator.h:13-33.
 * Same class name and EVERY public member of the reference class with the same signature, so that
 * ros/src/motion_detection_node.cpp compiles against this header unchanged (call sites node.cpp:82,99,209,214).
 * calculateOpticalFlow / calculateOpticalFlowTrajectory / calculateCompensatedFlow / varFlow forward to libmotion_b200.so
 * (include/motion_b200.h); writeFlow / writeTrajectories / drawMotionField are host code (file IO / drawing, no GPU work);
 * superPixelFlow (SLIC, no call site anywhere in the reference) is declared for source compatibility, see its note.
 *
 * Deviations from the reference, stated: (i) the egomotion defaults to the RANSAC homography (setEgomotionMode(0) = the
 * literal first-4 cv::getPerspectiveTransform, which is degenerate on a regular grid); (ii) gray conversion uses OpenCV 4.x's
 * 15-bit coefficients (3735/19235/9798, >> 15), OpenCV 2.4's 14-bit ones differ by at most 1 grey level; (iii) `comp` is the
 * thresholded difference exactly as cpp:124-127 -- the erode/dilate of BackgroundSubtractor (background_subtractor.cpp:31-32)
 * is applied only after setMorphology(true). */
#ifndef OPTICAL_FLOW_CALCULATOR_H_
#define OPTICAL_FLOW_CALCULATOR_H_

#include <opencv2/core/core.hpp>
#include <string>
#include <vector>

struct md_ctx;

class OpticalFlowCalculator
{
    public:
        OpticalFlowCalculator();
        virtual ~OpticalFlowCalculator();

        /* common/src/optical_flow_calculator.cpp:30-130 */
        int calculateOpticalFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow_vectors, int pixel_step, cv::Mat &comp, double min_vector_size);

        /* common/src/optical_flow_calculator.cpp:133-257 */
        int calculateOpticalFlowTrajectory(const std::vector<cv::Mat> &images, cv::Mat &optical_flow_vectors, std::vector<std::vector<cv::Point2f> > &trajectories, int pixel_step, cv::Mat &comp, double min_vector_size);

        /* common/src/optical_flow_calculator.cpp:259-330: the grid LK alone, 2 pyramid levels above the image, vectors over 1 px */
        int calculateCompensatedFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow_vectors, int pixel_step);

        /* common/src/optical_flow_calculator.cpp:337-416: SLIC superpixels + per-superpixel LK.  No call site exists in the
         * reference (node or common/), and it needs the vendored slic.cpp; declared so the class is source compatible.  This
         * build returns 0 vectors and leaves the outputs untouched; define MD_ADAPTER_EXTERNAL_SUPERPIXELFLOW when compiling
         * adapter.cpp to link the reference's own body (moved to its own translation unit) instead. */
        int superPixelFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow_image, cv::Mat &optical_flow_vectors);

        /* common/src/optical_flow_calculator.cpp:417-464 (the dense flow itself; the reference only draws it) */
        void varFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow, cv::Mat &optical_flow_vectors);

        /* common/src/optical_flow_calculator.cpp:467-507: arrows of the (U, V) field every xSpace / ySpace pixels (host drawing) */
        void drawMotionField(IplImage* imgU, IplImage* imgV, IplImage* imgMotion, int xSpace, int ySpace, float cutoff, int multiplier, CvScalar color);

        /* common/src/optical_flow_calculator.cpp:509-541: "<filename>_h" / "<filename>_f" CSV grids of dx / dy (node.cpp:209) */
        void writeFlow(const cv::Mat &flow_vectors, const std::string &filename, int pixel_step);
        /* common/src/optical_flow_calculator.cpp:543-562: one CSV row "x0, y0, x1, y1, ..." per trajectory (node.cpp:214) */
        void writeTrajectories(const std::vector<std::vector<cv::Point2f> > &trajectories, const std::string &filename);

        /* knobs that are compile-time constants in the reference */
        void setDevice(int device) { device_ = device; }
        void setEgomotionMode(int md_ego_mode) { ego_mode_ = md_ego_mode; release(); }
        void setMorphology(bool on) { morph_ = on; release(); }
        void setSeed(unsigned seed) { seed_ = seed; release(); }
        const double *lastHomography() const { return last_H_; }
        const char *lastError() const;

    private:
        OpticalFlowCalculator(const OpticalFlowCalculator &);
        OpticalFlowCalculator &operator=(const OpticalFlowCalculator &);
        bool ensure(int w, int h, int pixel_step, double min_vector_size, int max_batch);
        void release();
        static void ensureFlowMat(cv::Mat &m, int rows, int cols);
        md_ctx *ctx_;
        md_ctx *ctx2_;                 /* calculateCompensatedFlow: MAX_LEVEL = 2 */
        int w2_, h2_, ps2_;
        int w_, h_, ps_, batch_, device_, ego_mode_;
        bool morph_;
        unsigned seed_;
        double minvec_;
        double last_H_[9];
};

#endif
