"""Writes tests/golden/golden_ref.npz: outputs of the REFERENCE's own code (oracle/_ref: the unmodified
/root/reference/common/src/VarFlow.cpp and flow_clusterer.cpp / vector_cluster.cpp / point_cluster.cpp compiled against
oracle/ref_shim, recipe oracle/Makefile) on the seeded inputs tests/test_oracle_ref.py rebuilds.  Run in the container that
holds /root/reference:  python tests/golden/make_golden_ref.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import oracle as O  # noqa: E402
import test_oracle_ref as T  # noqa: E402

out = {}
fr = T._vf_pair(96, 80, 5)
out["vf_U_96x80"], out["vf_V_96x80"] = O.ref_varflow(fr[0], fr[1])
fr = T._vf_pair(160, 120, 9)
out["vf_U_160x120_l2"], out["vf_V_160x120_l2"] = O.ref_varflow(fr[0], fr[1], max_level=2, n1=1, n2=3)
out["ce_sizes"], out["ce_members"] = O.ref_cluster_euclidean(T._lattice_points(3), 33.3)
out["gc_sizes"], out["gc_members"] = O.ref_get_clusters(T._flow_field(2), 10, 50.0, 0.5)
# the chain composition by the reference's own optical_flow_calculator.cpp (oracle/_ref/libofc_ref.so)
f0, f1, ps = T._mover_pair()
nv, flow, comp = O.ref_calculate_optical_flow(f0, f1, ps, 1.0)
out["ofc_nv"] = np.array([nv])
out["ofc_flow_grid"] = flow[::ps, ::ps].copy()
out["ofc_comp_bits"] = np.packbits(comp > 0)
frames = T.synth.sequence(320, 240, 5, seed=21)[0]
nv, flow, traj = O.ref_calculate_trajectories(frames, 20, 1.0)
out["traj_nv"] = np.array([nv])
out["traj_complete"] = traj
# OutlierDetector by the reference's own outlier_detector.cpp (oracle/_ref/libod_ref.so)
fld = T._mad_field(3, ps=10, zero_frac=0.3)
out["mad_flags_nozero"] = (O.ref_find_outliers(fld, 10, False)[::10, ::10].reshape(-1) == 1.0).astype(np.uint8)
out["mad_flags_zero"] = (O.ref_find_outliers(fld, 10, True)[::10, ::10].reshape(-1) == 1.0).astype(np.uint8)
tr, _ = T._two_motion_trajectories(7)
out["sub_outliers"], out["sub_cols"] = O.ref_fit_subspace(tr, 2, 0.5, 7)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_ref.npz"), **out)
print({k: v.shape for k, v in out.items()})
