#!/usr/bin/env python3
"""Regenerates tests/golden/pcl_golden.npz from the reference checkout's OWN test fixtures.

Run in the build container only (needs /root/reference); the GPU box uses the committed .npz.
Nothing here executes reference code — PCL cannot be built in this image (no Eigen/Boost/FLANN) —
it only re-encodes the data files and known-answer constants that PCL's tests pin:

  bun0 / bun4 / sac_plane   test/bun0.pcd, test/bun4.pcd, test/sac_plane_test.pcd (ASCII PCD -> float32)
  corr_original (397x2)     test/registration/test_registration_api_data.h:3-402
  corr_reciprocal (53x2)    test/registration/test_registration_api_data.h:404-459
  corr_rej_dist (97x2)      ...:461-561   (CorrespondenceRejectorDistance, max 0.01)      [next row]
  corr_rej_median (139x2)   ...:563-706   (factor 0.5)                                      [next row]
  radius_offsets/indices    test/kdtree/kdtree_unit_test_results.xml (3283 FLANN radiusSearch(r=0.02)
                            neighbour lists over sac_plane_test.pcd, test/kdtree/test_kdtree.cpp:292-328)
  icp_bun0_bun4 (4x4)       test/registration/test_registration.cpp:250-269 (tolerances 1e-3 / 1e-2)
  knn10_*                   test/kdtree/test_kdtree.cpp:229-262 (10 points, k=10 around (50,50,50))
  voxel_*                   test/filters/test_filters.cpp:576-603
  normal_bun0               test/features/test_normal_estimation.cpp:106-127
  svd_Tref                  test/registration/test_registration_api_data.h:1122-1124
"""
import re
import sys
import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = sys.argv[2] if len(sys.argv) > 2 else __file__.rsplit("/", 1)[0] + "/pcl_golden.npz"


def read_ascii_pcd(path):
    fields, rows, data = None, [], False
    for line in open(path):
        if data:
            if line.strip():
                rows.append([float(v) for v in line.split()])
        elif line.startswith("FIELDS"):
            fields = line.split()[1:]
        elif line.startswith("DATA"):
            assert line.split()[1] == "ascii"
            data = True
    a = np.asarray(rows, dtype=np.float64).astype(np.float32)
    return fields, a


def c_int_pairs(text, name):
    m = re.search(r"const int %s\[(\d+)\]\[2\] = \{(.*?)\};" % name, text, re.S)
    n = int(m.group(1))
    vals = [int(v) for v in re.findall(r"-?\d+", m.group(2))]
    a = np.asarray(vals, dtype=np.int32).reshape(-1, 2)
    assert a.shape[0] == n, (name, a.shape, n)
    return a


g = {}
for nm in ("bun0", "bun4", "sac_plane_test"):
    f, a = read_ascii_pcd(f"{REF}/test/{nm}.pcd")
    g[nm.replace("_test", "")] = np.ascontiguousarray(a[:, :3])
    if "normal_x" in f:
        g[nm.replace("_test", "") + "_normals"] = np.ascontiguousarray(a[:, 3:7])

hdr = open(f"{REF}/test/registration/test_registration_api_data.h").read()
g["corr_original"] = c_int_pairs(hdr, "correspondences_original")
g["corr_reciprocal"] = c_int_pairs(hdr, "correspondences_reciprocal")
g["corr_rej_dist"] = c_int_pairs(hdr, "correspondences_dist")
g["corr_rej_median"] = c_int_pairs(hdr, "correspondences_median_dist")
g["corr_rej_one_to_one"] = c_int_pairs(hdr, "correspondences_one_to_one")
g["corr_rej_trimmed"] = c_int_pairs(hdr, "correspondences_trimmed")
# CorrespondenceRejectorSampleConsensus (test_registration_api.cpp:225-263): inlier threshold 0.01, 1000 iterations
g["corr_rej_sac"] = c_int_pairs(hdr, "correspondences_sac")
m = re.search(r"const float transform_from_SAC\[4\]\[4\] = \{(.*?)\};", hdr, re.S)
g["sac_transform"] = np.asarray([float(v.rstrip("f")) for v in re.findall(r"-?\d+\.?\d*(?:e-?\d+)?f?", m.group(1))], dtype=np.float64).reshape(4, 4)

xml = open(f"{REF}/test/kdtree/kdtree_unit_test_results.xml").read()
offs, idx = [0], []
for m in re.finditer(r"<point_(\d+)>(.*?)</point_\1>", xml, re.S):
    body = m.group(2)
    size = int(re.search(r"<size>(\d+)</size>", body).group(1))
    nn = [int(v) for v in re.findall(r"<nn_\d+>(\d+)</nn_\d+>", body)]
    assert len(nn) == size
    idx.extend(nn)
    offs.append(len(idx))
g["radius_offsets"] = np.asarray(offs, dtype=np.int64)
g["radius_indices"] = np.asarray(idx, dtype=np.int32)
g["radius_r"] = np.float64(0.02)
assert len(offs) - 1 == g["sac_plane"].shape[0] == 3283

g["icp_bun0_bun4"] = np.array(
    [[0.8806, 0.036481287330389023, -0.4724, 0.03453],
     [-0.02354, 0.9992, 0.03326, -0.001519],
     [0.4732, -0.01817, 0.8808, 0.04116],
     [0, 0, 0, 1]], dtype=np.float64)
g["knn10_points"] = np.array(
    [[86.6, 42.1, 92.4], [63.1, 18.4, 22.3], [35.5, 72.5, 37.3], [99.7, 37.0, 8.7],
     [22.4, 84.1, 64.0], [65.2, 73.4, 18.0], [60.4, 57.1, 4.5], [38.7, 17.6, 72.3],
     [14.2, 95.7, 34.7], [2.5, 26.5, 66.0]], dtype=np.float32)
g["knn10_query"] = np.array([50.0, 50.0, 50.0], dtype=np.float32)
g["knn10_indices"] = np.array([2, 7, 5, 1, 4, 6, 9, 0, 8, 3], dtype=np.int32)
g["knn10_distances"] = np.array(
    [877.8, 1674.7, 1802.6, 1937.5, 2120.6, 2228.8, 3064.5, 3199.7, 3604.2, 4344.8], dtype=np.float32)
# rescaled representation alpha = (1,2,3): test_kdtree.cpp:273-286
g["knn10_rescaled_indices"] = np.array([2, 9, 4, 7, 1, 5, 8, 0, 3, 6], dtype=np.int32)
g["knn10_rescaled_distances"] = np.array(
    [3686.9, 6769.2, 7177.0, 8802.3, 11071.5, 11637.3, 11742.4, 17769.0, 18497.3, 18942.0], dtype=np.float32)
g["voxel_leaf"] = np.float32(0.02)
g["voxel_count_all"] = np.int64(103)
g["voxel_count_z_005_01"] = np.int64(14)
g["voxel_count_z_negative"] = np.int64(100)
g["voxel_z_first"] = np.array([-0.026125, 0.039788, 0.052827], dtype=np.float64)
g["voxel_z_last"] = np.array([-0.073202, 0.1296, 0.051333], dtype=np.float64)
g["normal_bun0"] = np.array([0.035592, 0.369596, 0.928511, -0.0622552, 0.0693136], dtype=np.float64)
q = np.array([0.9, 0.1, -0.25, 0.15], dtype=np.float64)  # w,x,y,z
q /= np.linalg.norm(q)
w, x, y, z = q
R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
              [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
              [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
T = np.eye(4)
T[:3, :3] = R
T[:3, 3] = [0.5, -2.0, 1.0]
g["svd_Tref"] = T
np.savez_compressed(OUT, **g)
print("wrote", OUT, {k: getattr(v, "shape", None) for k, v in g.items()})
