/* flow_clusterer.h -- drop-in replacement of common/include/motion_detection/flow_clusterer.h:13-27 for the entry point
 * the node's live (egomotion) path calls: clusterEuclidean (ros/src/motion_detection_node.cpp:355,
 * common/src/flow_clusterer.cpp:227-269).  getClusters / getClustersCenters / clusterFlowVectors (the !egomotion_ branch
 * and dead code) stay with the reference's own sources.  boundingBoxes() returns what
 * OpticalFlowVisualizer::showBoundingBoxes (optical_flow_visualizer.cpp:223-240) computes before it draws. */
#ifndef FLOW_CLUSTERER_H_
#define FLOW_CLUSTERER_H_

#include <opencv2/core/core.hpp>
#include <vector>

struct md_ctx;

class FlowClusterer
{
    public:
        FlowClusterer();
        virtual ~FlowClusterer();

        /* common/src/flow_clusterer.cpp:227-269 */
        std::vector<std::vector<cv::Point2f> > clusterEuclidean(const std::vector<cv::Point2f> &points, double distance_threshold);

        /* tl.x, tl.y, br.x, br.y of cv::boundingRect of every cluster returned by the LAST clusterEuclidean call -- the
         * columns MotionLogger::writeBoundingBox logs (motion_logger.cpp:43-47) */
        const std::vector<cv::Vec4i> &boundingBoxes() const { return boxes_; }
        void setDevice(int device) { device_ = device; }

    private:
        FlowClusterer(const FlowClusterer &);
        FlowClusterer &operator=(const FlowClusterer &);
        md_ctx *ctx_;
        int device_;
        std::vector<cv::Vec4i> boxes_;
};
#endif
