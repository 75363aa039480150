// cluster_ref_entry.cpp -- C entry points into the reference's own FlowClusterer (common/src/flow_clusterer.cpp compiled
// unmodified against oracle/ref_shim/opencv2/*), test infrastructure only.
#include <stdint.h>

#include <motion_detection/flow_clusterer.h>      // the REFERENCE's header

// FlowClusterer::clusterEuclidean (flow_clusterer.cpp:231-269): returns the number of clusters (> 5 members); sizes[k] and the
// member points (concatenated, cluster after cluster, arrival order) in members [n][2]
extern "C" int ref_cluster_euclidean(const float *pts, int n, double distance_threshold, int32_t *sizes, float *members)
{
    std::vector<cv::Point2f> p(n);
    for (int i = 0; i < n; i++) p[i] = cv::Point2f(pts[2 * i], pts[2 * i + 1]);
    FlowClusterer fc;
    std::vector<std::vector<cv::Point2f> > cl = fc.clusterEuclidean(p, distance_threshold);
    int m = 0;
    for (size_t k = 0; k < cl.size(); k++) {
        sizes[k] = (int32_t)cl[k].size();
        for (size_t j = 0; j < cl[k].size(); j++, m++) { members[2 * m] = cl[k][j].x; members[2 * m + 1] = cl[k][j].y; }
    }
    return (int)cl.size();
}

// FlowClusterer::getClusters (flow_clusterer.cpp:117-229) on a Vec4d field [h][w][4] (the node's flow Mat read as Vec4d)
extern "C" int ref_get_clusters(const double *flow, int w, int h, int pixel_step, double distance_threshold, double angular_threshold,
                                int32_t *sizes, double *members)
{
    cv::Mat f(h, w, CV_64FC4);
    memcpy(f.data, flow, sizeof(double) * 4 * (size_t)w * h);
    FlowClusterer fc;
    std::vector<std::vector<cv::Vec4d> > cl = fc.getClusters(f, pixel_step, distance_threshold, angular_threshold);
    int m = 0;
    for (size_t k = 0; k < cl.size(); k++) {
        sizes[k] = (int32_t)cl[k].size();
        for (size_t j = 0; j < cl[k].size(); j++, m++)
            for (int q = 0; q < 4; q++) members[4 * m + q] = cl[k][j][q];
    }
    return (int)cl.size();
}
