"""CPU checks of the drop-in C++ boundary (adapter/): the reference-named classes declare EVERY public member of the reference
headers, a translation unit repeating every ofc_. / od_. / fc_. call expression of ros/src/motion_detection_node.cpp compiles
and links against them (adapter/test/node_callsites.cpp), and the host-side members (CSV writers of
common/src/optical_flow_calculator.cpp:509-562, getClustersCenters, drawMotionField) produce the reference's formats.
No GPU work is done here."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "adapter")

# public members of the reference classes: common/include/motion_detection/optical_flow_calculator.h:16-32,
# flow_clusterer.h:15-23, outlier_detector.h:15-21, VarFlow.h:33-36
REFERENCE_PUBLIC_MEMBERS = {
    "optical_flow_calculator.h": ["calculateOpticalFlow", "calculateOpticalFlowTrajectory", "calculateCompensatedFlow", "superPixelFlow",
                                  "varFlow", "drawMotionField", "writeFlow", "writeTrajectories"],
    "flow_clusterer.h": ["clusterFlowVectors", "getClustersCenters", "getClusters", "clusterEuclidean"],
    "outlier_detector.h": ["findOutliers", "getOutlierVectors", "fitSubspace"],
    "VarFlow.h": ["CalcFlow"],
}


def _build():
    subprocess.check_call(["make", "-C", ADAPTER, "-s"])


def test_headers_declare_every_reference_member():
    for hdr, members in REFERENCE_PUBLIC_MEMBERS.items():
        txt = open(os.path.join(ADAPTER, "include", "motion_detection", hdr)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        public = txt.split("private:")[0]
        for m in members:
            assert re.search(r"\b%s\s*\(" % m, public), "%s: public member %s is not declared" % (hdr, m)


def test_node_call_sites_compile_link_and_host_members_run(tmp_path):
    _build()
    exe = os.path.join(ADAPTER, "node_callsites")
    assert os.path.exists(exe)
    out = subprocess.run([exe, "cpu", str(tmp_path)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    sec = {}
    cur = None
    for line in out.stdout.splitlines():
        if line.startswith("== "):
            cur = line[3:]
            sec[cur] = []
        else:
            sec[cur].append(line)
    # writeFlow: "<name>_h" = dx, "<name>_f" = dy; rows / columns in steps of pixel_step; ", " separators; failed vectors
    # (x == -1) and untouched grid nodes print 0; default ostream precision (6 significant digits)
    assert sec["flow_h"] == ["0, 1.5, 0", "0, 0, -2"]
    assert sec["flow_f"] == ["0, -0.25, 0", "0, 0, 0.333333"]
    # writeTrajectories: one row per trajectory, "x0, y0, x1, y1, ..."
    assert sec["traj"] == ["10, 20, 11.25, 19.5, 12.5, 19", "100, 200, 100.125, 200"]
    # getClustersCenters: parallel neighbours merge, the opposite vector in their middle founds its own cluster
    assert sec["centers"] == ["10, 0", "10, 10", "160, 10"]
    assert "kmeans 0 x 2" in sec and "superpixel 0" in sec
    assert "arrow tail 255 tip 255 lit 1" in sec
    # the files exist where the node would look for them
    for f in ("flow_h", "flow_f", "traj"):
        assert os.path.getsize(os.path.join(str(tmp_path), f)) > 0


def test_node_call_sites_cover_every_call_expression_of_the_node():
    """The call expressions of node.cpp (line numbers of the reference) and the member each one needs."""
    src = open(os.path.join(ADAPTER, "test", "node_callsites.cpp")).read()
    calls = {82: "ofc_.calculateOpticalFlow(", 99: "ofc_.calculateOpticalFlowTrajectory(", 114: "od_.findOutliers(",
             121: "od_.getOutlierVectors(", 127: "fc_.getClusters(outlier_vectors", 169: "fc_.getClusters(flow_vectors",
             209: "ofc_.writeFlow(", 214: "ofc_.writeTrajectories(", 348: "od_.fitSubspace(trajectories, outlier_points, num_motions, sigma)",
             355: "fc_.clusterEuclidean(outlier_points, distance_threshold)", 375: "fc_.getClusters(optical_flow_vectors",
             494: "od_.fitSubspace(trajectories, outlier_points, 2, residual_threshold)", 497: "fc_.clusterEuclidean("}
    for line, expr in calls.items():
        assert expr in src, "node.cpp:%d call %s is not reproduced" % (line, expr)
