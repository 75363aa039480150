/* optical_flow_calculator.h -- drop-in replacement of common/include/motion_detection/optical_flow_calculator.h:13-33.
 * Same class name and public signatures; the bodies forward to libmotion_b200.so (include/motion_b200.h).
 * Viz / file-IO members of the reference (drawMotionField, writeFlow, writeTrajectories) and superPixelFlow (SLIC, uncalled)
 * are outside the accelerated path and stay with the reference's own sources. */
#ifndef OPTICAL_FLOW_CALCULATOR_H_
#define OPTICAL_FLOW_CALCULATOR_H_

#include <opencv2/core/core.hpp>
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

        /* common/src/optical_flow_calculator.cpp:417-464 (the dense flow itself; the reference only draws it) */
        void varFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow, cv::Mat &optical_flow_vectors);

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
