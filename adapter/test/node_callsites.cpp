// node_callsites.cpp -- call-site conformance of the drop-in headers.
//
// ros/src/motion_detection_node.cpp holds OpticalFlowCalculator / OutlierDetector / FlowClusterer BY VALUE
// (ros/include/motion_detection/motion_detection_node.h:87-96) and calls their members at node.cpp:82, 99, 114, 121, 127, 169,
// 209, 214, 348, 355, 375, 494, 497.  This translation unit declares the same members the same way and repeats EVERY one of
// those call expressions with the argument types the node uses (cv::Mat, std::vector<cv::Mat>, std::vector<std::vector<
// cv::Point2f> >, int / double / bool ROS parameters, std::string file names): if it compiles and links against
// adapter/include + libmotion_adapter.so, the unchanged node source does too (ROS and the visualisers aside).
// Run without arguments it executes only the host-side members (writers, getClustersCenters, clusterFlowVectors,
// superPixelFlow, drawMotionField) and prints what they produce, which tests/test_adapter_cpu.py checks; with "gpu" it also
// drives the device members once.
#include <motion_detection/optical_flow_calculator.h>
#include <motion_detection/flow_clusterer.h>
#include <motion_detection/outlier_detector.h>
#include <motion_detection/VarFlow.h>

#include <cstdio>
#include <fstream>
#include <iostream>
#include <list>
#include <string>
#include <vector>

class NodeCallSites
{
  public:
    NodeCallSites() : pixel_step_(10), min_vector_size_(0.2), include_zeros_(false) {}

    // node.cpp:76-92
    void runOpticalFlow(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &optical_flow_vectors)
    {
        cv::Mat debug_image;
        optical_flow_vectors = cv::Mat::zeros(image1.rows, image1.cols, CV_32FC4);
        int num_vectors = ofc_.calculateOpticalFlow(image1, image2, optical_flow_vectors, pixel_step_, debug_image, min_vector_size_);
        (void)num_vectors;
    }
    // node.cpp:94-110
    void runOpticalFlowTrajectory(const std::vector<cv::Mat> &images, cv::Mat &optical_flow_vectors,
                                  std::vector<std::vector<cv::Point2f> > &trajectories, cv::Mat &optical_flow_image)
    {
        cv::Mat debug_image;
        (void)optical_flow_image;
        optical_flow_vectors = cv::Mat::zeros(images[0].rows, images[0].cols, CV_32FC4);
        int num_vectors = ofc_.calculateOpticalFlowTrajectory(images, optical_flow_vectors, trajectories, pixel_step_, debug_image, min_vector_size_);
        (void)num_vectors;
    }
    // node.cpp:112-131
    void detectOutliers(const cv::Mat &original_image, const cv::Mat &optical_flow_vectors, cv::Mat &outlier_mask, bool include_zeros)
    {
        (void)original_image;
        od_.findOutliers(optical_flow_vectors, outlier_mask, include_zeros, pixel_step_, false);
        cv::Mat outlier_vectors;
        od_.getOutlierVectors(optical_flow_vectors, outlier_mask, outlier_vectors, pixel_step_);
        std::vector<std::vector<cv::Vec4d> > clusters;
        double distance_threshold = 50.0, angular_threshold = 0.15;
        clusters = fc_.getClusters(outlier_vectors, pixel_step_, distance_threshold, angular_threshold);
    }
    // node.cpp:162-205
    void clusterFlow(const cv::Mat &image, const cv::Mat &flow_vectors, std::vector<std::vector<cv::Vec4d> > &clusters)
    {
        (void)image;
        double distance_threshold = 50.0, angular_threshold = 0.15;
        clusters = fc_.getClusters(flow_vectors, pixel_step_, distance_threshold, angular_threshold);
    }
    // node.cpp:207-215
    void writeVectors(const cv::Mat &flow_vectors, const std::string &filename) { ofc_.writeFlow(flow_vectors, filename, pixel_step_); }
    void writeTrajectories(const std::vector<std::vector<cv::Point2f> > &trajectories, const std::string &filename)
    {
        ofc_.writeTrajectories(trajectories, filename);
    }
    // node.cpp:342-391 (imageCallback, both branches of `if (egomotion_)`)
    void callbackBody(const std::vector<std::vector<cv::Point2f> > &trajectories, const cv::Mat &optical_flow_vectors, bool egomotion,
                      int num_motions, double distance_threshold, std::vector<std::vector<cv::Point2f> > &clusters)
    {
        if (egomotion) {
            std::vector<cv::Point2f> outlier_points;
            double sigma = 0.5;
            std::vector<std::vector<cv::Point2f> > trajectory_subspace_vectors;
            trajectory_subspace_vectors = od_.fitSubspace(trajectories, outlier_points, num_motions, sigma);
            clusters = fc_.clusterEuclidean(outlier_points, distance_threshold);
        } else {
            std::vector<std::vector<cv::Vec4d> > cluster_vec;
            double angular_threshold = 0.15;
            cluster_vec = fc_.getClusters(optical_flow_vectors, pixel_step_, distance_threshold, angular_threshold);
            for (int i = 0; i < (int)cluster_vec.size(); i++) {
                std::vector<cv::Point2f> cc;
                std::vector<cv::Vec4d> cc_v = cluster_vec.at(i);
                for (int j = 0; j < (int)cc_v.size(); j++) cc.push_back(cv::Point2f((float)cc_v.at(j)[0], (float)cc_v.at(j)[1]));
                clusters.push_back(cc);
            }
        }
    }
    // node.cpp:486-497 (run())
    void runBody(const std::vector<cv::Mat> &cv_images, std::vector<std::vector<cv::Point2f> > &clusters)
    {
        cv::Mat optical_flow_vectors, optical_flow_image;
        std::vector<std::vector<cv::Point2f> > trajectories;
        runOpticalFlowTrajectory(cv_images, optical_flow_vectors, trajectories, optical_flow_image);
        std::vector<cv::Point2f> outlier_points;
        double residual_threshold = 0.2;
        od_.fitSubspace(trajectories, outlier_points, 2, residual_threshold);
        double distance_threshold = 50.0;
        clusters = fc_.clusterEuclidean(outlier_points, distance_threshold);
    }
    // the public members of the reference headers that have no call site in the node (optical_flow_calculator.h:24-29,
    // flow_clusterer.h:19-21): same signatures
    void uncalledMembers(const cv::Mat &image1, const cv::Mat &image2, cv::Mat &flow, IplImage *u, IplImage *v, IplImage *motion)
    {
        cv::Mat a, b;
        int n1 = ofc_.calculateCompensatedFlow(image1, image2, flow, pixel_step_);
        int n2 = ofc_.superPixelFlow(image1, image2, a, b);
        ofc_.varFlow(image1, image2, a, b);
        ofc_.drawMotionField(u, v, motion, 10, 10, 0.5f, 5, cvScalar(255, 0, 0));
        cv::Mat centers = fc_.clusterFlowVectors(flow);
        std::vector<cv::Point2f> cc = fc_.getClustersCenters(flow, pixel_step_, 50.0, 0.15);
        (void)n1; (void)n2;
    }

    OpticalFlowCalculator ofc_;
    FlowClusterer fc_;
    OutlierDetector od_;
    int pixel_step_;
    double min_vector_size_;
    bool include_zeros_;
};

static void dump(const std::string &path)
{
    std::ifstream f(path.c_str());
    std::string line;
    std::cout << "== " << path.substr(path.find_last_of('/') + 1) << "\n";
    while (std::getline(f, line)) std::cout << line << "\n";
}

int main(int argc, char **argv)
{
    const std::string dir = argc > 2 ? argv[2] : "/tmp";
    const bool gpu = argc > 1 && std::string(argv[1]) == "gpu";
    NodeCallSites node;

    // a 30 x 20 Vec4d field on a 10-pixel grid: two tracked vectors, one failed vector, the rest zero
    cv::Mat flow = cv::Mat::zeros(20, 30, CV_64FC4);
    for (int y = 0; y < 20; y += 10)
        for (int x = 0; x < 30; x += 10) { cv::Vec4d &e = flow.at<cv::Vec4d>(y, x); e[0] = x; e[1] = y; }
    flow.at<cv::Vec4d>(0, 10)[2] = 1.5; flow.at<cv::Vec4d>(0, 10)[3] = -0.25;
    flow.at<cv::Vec4d>(10, 20)[2] = -2.0; flow.at<cv::Vec4d>(10, 20)[3] = 0.333333333;
    { cv::Vec4d &e = flow.at<cv::Vec4d>(10, 0); e[0] = -1.0; e[1] = -1.0; e[2] = 7.0; e[3] = 7.0; }     // failed: written as 0
    node.writeVectors(flow, dir + "/flow");
    dump(dir + "/flow_h");
    dump(dir + "/flow_f");

    std::vector<std::vector<cv::Point2f> > traj(2);
    traj[0].push_back(cv::Point2f(10.f, 20.f)); traj[0].push_back(cv::Point2f(11.25f, 19.5f)); traj[0].push_back(cv::Point2f(12.5f, 19.f));
    traj[1].push_back(cv::Point2f(100.f, 200.f)); traj[1].push_back(cv::Point2f(100.125f, 200.f));
    node.writeTrajectories(traj, dir + "/traj");
    dump(dir + "/traj");

    // getClustersCenters (host): two groups of parallel vectors far apart, one vector of opposite direction inside group 0
    cv::Mat field = cv::Mat::zeros(40, 200, CV_64FC4);
    for (int x = 0; x < 30; x += 10) { cv::Vec4d &e = field.at<cv::Vec4d>(0, x); e[0] = x; e[1] = 0; e[2] = 1.0; e[3] = 0.0; }
    for (int x = 150; x < 180; x += 10) { cv::Vec4d &e = field.at<cv::Vec4d>(10, x); e[0] = x; e[1] = 10; e[2] = 0.0; e[3] = 2.0; }
    { cv::Vec4d &e = field.at<cv::Vec4d>(10, 10); e[0] = 10; e[1] = 10; e[2] = -1.0; e[3] = 0.0; }
    std::vector<cv::Point2f> cc = node.fc_.getClustersCenters(field, 10, 50.0, 0.15);
    std::cout << "== centers\n";
    for (size_t i = 0; i < cc.size(); i++) std::cout << cc[i].x << ", " << cc[i].y << "\n";
    cv::Mat km = node.fc_.clusterFlowVectors(field);
    std::cout << "== kmeans " << km.rows << " x " << km.cols << "\n";
    cv::Mat a, b;
    std::cout << "== superpixel " << node.ofc_.superPixelFlow(field, field, a, b) << "\n";

    // drawMotionField: one arrow of (U, V) = (2, -1) at (10, 10), multiplier 5 -> tip at (20, 15)
    std::vector<float> U(30 * 20, 0.f), V(30 * 20, 0.f);
    std::vector<char> canvas(30 * 20, 0);
    U[10 * 30 + 10] = 2.f; V[10 * 30 + 10] = -1.f;
    IplImage iu = {1, IPL_DEPTH_32F, 30, 20, 30 * 4, reinterpret_cast<char *>(U.data())};
    IplImage iv = {1, IPL_DEPTH_32F, 30, 20, 30 * 4, reinterpret_cast<char *>(V.data())};
    IplImage im = {1, IPL_DEPTH_8U, 30, 20, 30, canvas.data()};
    node.ofc_.drawMotionField(&iu, &iv, &im, 10, 10, 0.5f, 5, cvScalar(255));
    int lit = 0;
    for (size_t i = 0; i < canvas.size(); i++) lit += canvas[i] != 0;
    std::cout << "== arrow tail " << (int)(unsigned char)canvas[10 * 30 + 10] << " tip " << (int)(unsigned char)canvas[15 * 30 + 20] << " lit " << (lit > 10) << "\n";

    if (gpu) {
        // every device member once, through the node-shaped methods
        const int w = 320, h = 240, F = 5;
        std::vector<cv::Mat> imgs(F);
        for (int f = 0; f < F; f++) {
            imgs[f].create(h, w, CV_8UC3);
            for (int y = 0; y < h; y++)
                for (int x = 0; x < w; x++) {
                    const unsigned char g = (unsigned char)(((x + 2 * f) * 7 + (y + f) * 13 + ((x + 2 * f) / 8) * 31 + ((y + f) / 8) * 17) & 63);
                    for (int c = 0; c < 3; c++) imgs[f].data[(size_t)y * imgs[f].step + 3 * x + c] = g;
                }
        }
        cv::Mat fl, mask;
        node.runOpticalFlow(imgs[0], imgs[1], fl);
        std::vector<std::vector<cv::Point2f> > tr, clusters;
        cv::Mat ofi;
        node.runOpticalFlowTrajectory(imgs, fl, tr, ofi);
        node.detectOutliers(imgs[0], fl, mask, false);
        std::vector<std::vector<cv::Vec4d> > vc;
        node.clusterFlow(imgs[0], fl, vc);
        node.callbackBody(tr, fl, true, 2, 50.0, clusters);
        node.callbackBody(tr, fl, false, 2, 50.0, clusters);
        node.runBody(imgs, clusters);
        std::cout << "== gpu trajectories " << tr.size() << " flow type " << fl.type() << "\n";
    }
    return 0;
}
