// adapter_test.cpp -- drives the reference-named C++ classes (adapter/) end to end and dumps what they return, so
// tests/test_adapter_gpu.py can compare it with the oracle.  usage: adapter_test frames.bin w h F out.bin
#include <motion_detection/optical_flow_calculator.h>
#include <motion_detection/outlier_detector.h>
#include <motion_detection/VarFlow.h>
#include <motion_detection/flow_clusterer.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static void put(FILE *f, const void *p, size_t n) { fwrite(p, 1, n, f); }

int main(int argc, char **argv)
{
    if (argc < 6) return 2;
    const int w = atoi(argv[2]), h = atoi(argv[3]), F = atoi(argv[4]);
    std::vector<cv::Mat> gray(F), rgb(F);
    FILE *fi = fopen(argv[1], "rb");
    if (!fi) return 3;
    for (int i = 0; i < F; i++) {
        gray[i].create(h, w, CV_8UC1);
        if (fread(gray[i].data, 1, (size_t)w * h, fi) != (size_t)w * h) return 4;
        rgb[i].create(h, w, CV_8UC3);                      // the node hands 8UC3 frames (node.cpp:271)
        for (int k = 0; k < w * h; k++) rgb[i].data[3 * k] = rgb[i].data[3 * k + 1] = rgb[i].data[3 * k + 2] = gray[i].data[k];
    }
    fclose(fi);
    FILE *fo = fopen(argv[5], "wb");

    // --- calculateOpticalFlow, as runOpticalFlow calls it (node.cpp:76-92): flow Mat allocated CV_32FC4 by the caller
    OpticalFlowCalculator ofc;
    ofc.setSeed(11);
    cv::Mat flow = cv::Mat::zeros(h, w, CV_32FC4), comp;
    int nv = ofc.calculateOpticalFlow(rgb[0], rgb[1], flow, 10, comp, 0.2);
    int32_t hdr[4] = {nv, comp.rows, comp.cols, flow.type()};
    put(fo, hdr, sizeof hdr);
    put(fo, ofc.lastHomography(), 9 * sizeof(double));
    if (comp.rows == h) put(fo, comp.data, (size_t)w * h);
    for (int x = 0; x < w; x += 10)
        for (int y = 0; y < h; y += 10) put(fo, &flow.at<cv::Vec4d>(y, x), sizeof(cv::Vec4d));

    // --- calculateCompensatedFlow: the grid LK alone with 2 levels above the image (cpp:259-330)
    {
        cv::Mat flowc = cv::Mat::zeros(h, w, CV_32FC4);
        int32_t nvc = ofc.calculateCompensatedFlow(rgb[0], rgb[1], flowc, 10);
        put(fo, &nvc, sizeof nvc);
        for (int x = 0; x < w; x += 10)
            for (int y = 0; y < h; y += 10) put(fo, &flowc.at<cv::Vec4d>(y, x), sizeof(cv::Vec4d));
    }

    // --- findOutliers on that flow field (detectOutliers, node.cpp:112-121)
    {
        OutlierDetector odm;
        cv::Mat prob, outv;
        odm.findOutliers(flow, prob, false, 10, false);
        odm.getOutlierVectors(flow, prob, outv, 10);
        int32_t nflag = 0, nvec = 0;
        for (int y = 0; y < h; y += 10)
            for (int x = 0; x < w; x += 10) {
                if (prob.at<double>(y, x) > 0.5) nflag++;
                const cv::Vec4d &e = outv.at<cv::Vec4d>(y, x);
                if (e[2] != 0.0 || e[3] != 0.0) nvec++;
            }
        int32_t m[2] = {nflag, nvec};
        put(fo, m, sizeof m);
    }

    // --- calculateOpticalFlowTrajectory + fitSubspace, the live path (node.cpp:94-110, 348)
    cv::Mat flow2 = cv::Mat::zeros(h, w, CV_32FC4), comp2;
    std::vector<std::vector<cv::Point2f> > traj;
    int nv2 = ofc.calculateOpticalFlowTrajectory(rgb, flow2, traj, 10, comp2, 0.2);
    OutlierDetector od;
    od.setSeed(3);
    std::vector<cv::Point2f> outliers;
    std::vector<std::vector<cv::Point2f> > basis = od.fitSubspace(traj, outliers, 2, 0.5);
    int32_t t2[5] = {nv2, (int32_t)traj.size(), (int32_t)outliers.size(), (int32_t)basis.size(), od.lastInliers()};
    put(fo, t2, sizeof t2);
    for (size_t i = 0; i < traj.size(); i++) put(fo, traj[i].data(), sizeof(cv::Point2f) * F);

    // --- clusterEuclidean on the outlier points (node.cpp:355) + the rectangles showBoundingBoxes would draw
    FlowClusterer fc;
    std::vector<std::vector<cv::Point2f> > clusters = fc.clusterEuclidean(outliers, 50.0);
    int32_t nc = (int32_t)clusters.size();
    put(fo, &nc, sizeof nc);
    for (int c = 0; c < nc; c++) {
        int32_t sz = (int32_t)clusters[c].size();
        put(fo, &sz, sizeof sz);
        put(fo, &fc.boundingBoxes()[c], 4 * sizeof(int32_t));
        put(fo, clusters[c].data(), sizeof(cv::Point2f) * sz);
    }
    if (!outliers.empty()) put(fo, outliers.data(), sizeof(cv::Point2f) * outliers.size());

    // --- VarFlow with the parameters of varFlow() (cpp:422-429)
    VarFlow vf(w, h, 4, 0, 2, 2, 2.8f, 1400.f, 1.5f);
    std::vector<float> U((size_t)w * h), V((size_t)w * h);
    IplImage a = {1, IPL_DEPTH_8U, w, h, w, (char *)gray[0].data}, b = {1, IPL_DEPTH_8U, w, h, w, (char *)gray[1].data};
    IplImage u = {1, IPL_DEPTH_32F, w, h, w * 4, (char *)U.data()}, v = {1, IPL_DEPTH_32F, w, h, w * 4, (char *)V.data()};
    int32_t ok = vf.CalcFlow(&a, &b, &u, &v, false);
    put(fo, &ok, sizeof ok);
    put(fo, U.data(), sizeof(float) * U.size());
    put(fo, V.data(), sizeof(float) * V.size());
    fclose(fo);
    printf("adapter_test: num_vectors=%d trajectories=%zu outliers=%zu varflow=%d\n", nv, traj.size(), outliers.size(), ok);
    return 0;
}
