/* flow_clusterer.h -- drop-in replacement of common/include/motion_detection/flow_clusterer.h:13-27: every public member of
 * the reference class with the same signature (call sites ros/src/motion_detection_node.cpp:127,169,355,375,497).
 * clusterEuclidean (common/src/flow_clusterer.cpp:227-269) and getClusters (:117-229, live code :178-227) run on the GPU
 * (md_cluster_points / md_cluster_vectors); getClustersCenters (:80-115) and clusterFlowVectors (:23-78) have no call site in
 * the reference and are host code, see their notes.  boundingBoxes() returns what
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

        /* common/src/flow_clusterer.cpp:23-78: FLANN hierarchical k-means (KMEANSPP seeded by rand(): not reproducible) over the
         * positions of every 20th vector with dy > 0.  No call site in the reference; FLANN is not part of this library: returns
         * an empty (0 x 2, CV_32F) centre matrix. */
        cv::Mat clusterFlowVectors(const cv::Mat &flow_vectors);

        /* common/src/flow_clusterer.cpp:80-115: like getClusters but WITHOUT the break -- a vector joins every cluster that
         * qualifies -- returning the centroids (vector_cluster.cpp:52-65).  No call site in the reference; host code. */
        std::vector<cv::Point2f> getClustersCenters(const cv::Mat &flow_vectors, int pixel_step, double distance_threshold, double angular_threshold);

        /* common/src/flow_clusterer.cpp:117-229 (live code :178-227): greedy first-fit grouping by distance and orientation of
         * the non-zero vectors of the Vec4d field; clusters with more than 5 members, creation order, members in arrival order */
        std::vector<std::vector<cv::Vec4d> > getClusters(const cv::Mat &flow_vectors, int pixel_step, double distance_threshold, double angular_threshold);

        /* common/src/flow_clusterer.cpp:227-269 */
        std::vector<std::vector<cv::Point2f> > clusterEuclidean(const std::vector<cv::Point2f> &points, double distance_threshold);

        /* tl.x, tl.y, br.x, br.y of cv::boundingRect of every cluster returned by the LAST clusterEuclidean call -- the
         * columns MotionLogger::writeBoundingBox logs (motion_logger.cpp:43-47) */
        const std::vector<cv::Vec4i> &boundingBoxes() const { return boxes_; }
        void setDevice(int device) { device_ = device; }

    private:
        bool ensureContext();
        FlowClusterer(const FlowClusterer &);
        FlowClusterer &operator=(const FlowClusterer &);
        md_ctx *ctx_;
        int device_;
        std::vector<cv::Vec4i> boxes_;
};
#endif
