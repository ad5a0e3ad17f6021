"""Pins the CPU oracle against every golden vector PCL's own tests hold for the ICP hot path
(SURVEY.md §4 / §8c).  CPU-only."""
import numpy as np
import pytest


def test_correspondences_bun0_bun4_397(golden, orc):
    # test/registration/test_registration_api.cpp:83-104
    tgt = orc.Index(orc.to_xyz1(golden["bun4"]))
    c = tgt.correspondences(orc.to_xyz1(golden["bun0"]))
    assert c.size == 397
    assert np.array_equal(c["index_query"], np.arange(397))
    assert np.array_equal(c["index_query"], golden["corr_original"][:, 0])
    assert np.array_equal(c["index_match"], golden["corr_original"][:, 1])


def test_reciprocal_correspondences_53(golden, orc):
    # test/registration/test_registration_api.cpp:107-128
    src = orc.to_xyz1(golden["bun0"])
    tgt = orc.Index(orc.to_xyz1(golden["bun4"]))
    c = tgt.correspondences_reciprocal(src, orc.Index(src))
    assert c.size == 53
    assert np.array_equal(c["index_query"], golden["corr_reciprocal"][:, 0])
    assert np.array_equal(c["index_match"], golden["corr_reciprocal"][:, 1])


def test_radius_search_3283_lists(golden, orc):
    # test/kdtree/test_kdtree.cpp:292-328 + kdtree_unit_test_results.xml: exact count AND order
    cloud = orc.to_xyz1(golden["sac_plane"])
    offs, idx, d2 = orc.Index(cloud).radius(cloud, float(golden["radius_r"]))
    assert np.array_equal(offs, golden["radius_offsets"])
    assert np.array_equal(idx, golden["radius_indices"])
    for i in (0, 1, 1000, 3282):
        seg = d2[offs[i]:offs[i + 1]]
        assert np.all(np.diff(seg) >= 0) and np.all(seg < np.float32(0.02 * 0.02))


def test_knn10_known_answer(golden, orc):
    # test/kdtree/test_kdtree.cpp:229-262
    pts = orc.to_xyz1(golden["knn10_points"])
    q = orc.to_xyz1(golden["knn10_query"][None])
    idx, d2, keff = orc.Index(pts).knn(q, 10)
    assert keff == 10
    assert np.array_equal(idx[0], golden["knn10_indices"])
    assert np.allclose(d2[0], golden["knn10_distances"], atol=0.1)
    # rescaled representation alpha=(1,2,3) (:273-286) == scaling coordinates before indexing
    a = np.array([1, 2, 3], np.float32)
    idx, d2, _ = orc.Index(orc.to_xyz1(golden["knn10_points"] * a)).knn(orc.to_xyz1(golden["knn10_query"][None] * a), 10)
    assert np.array_equal(idx[0], golden["knn10_rescaled_indices"])
    assert np.allclose(d2[0], golden["knn10_rescaled_distances"], atol=0.1)
    # k larger than the cloud is clamped (kdtree_flann.hpp:241-242)
    idx, d2, keff = orc.Index(pts).knn(q, 12)
    assert keff == 10 and np.all(idx[0, 10:] == -1)


@pytest.mark.parametrize("scalar_is_double", [False, True])
def test_icp_bun0_bun4_matrix(golden, orc, scalar_is_double):
    # test/registration/test_registration.cpp:236-270
    r = orc.icp_align(orc.to_xyz1(golden["bun0"]), orc.to_xyz1(golden["bun4"]), max_iterations=50,
                      transformation_epsilon=1e-8, max_correspondence_distance=0.05,
                      scalar_is_double=scalar_is_double)
    T, G = r["final"], golden["icp_bun0_bun4"]
    assert r["converged"]
    tol = np.full((4, 4), 1e-3)
    tol[0, 1] = 1e-2
    tol[3, :] = 0.0
    assert np.all(np.abs(T - G) <= tol), (T, r)


def test_icp_translated(golden, orc):
    # test/registration/test_registration.cpp:161-195
    src = orc.to_xyz1(golden["bun0"])
    tgt = src.copy()
    tgt[:, 2] += np.float32(0.2)
    r = orc.icp_align(src, tgt, max_iterations=50)
    assert r["converged"]
    T = r["final"]
    assert orc.Index(tgt).fitness_score(src, T) < 1e-6
    assert np.allclose(np.diag(T)[:3], 1.0, atol=2e-3)
    assert np.allclose(T[:3, 3], [0, 0, 0.2], atol=2e-3)


def test_fitness_score_indices(orc):
    # test/registration/test_registration.cpp:198-233
    src = orc.to_xyz1(np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [10, 0, 0]], np.float32))
    tgt = orc.to_xyz1(np.array([[0, 0, 0], [0, 1, 0], [0, 0, 1], [10, 0, 0.5]], np.float32))
    t = orc.Index(tgt)
    I = np.eye(4)
    assert abs(t.fitness_score(src, I, max_range=1.0) - 0.0625) < 1e-4
    assert abs(t.fitness_score(src, I, max_range=1.0, indices=[0, 1, 2]) - 0.0) < 1e-4


def test_transformation_estimation_svd(golden, orc):
    # test/registration/test_registration_api.cpp:383-423 (tolerance 1e-6 on quaternion / translation)
    src = orc.to_xyz1(golden["bun4"])
    Tref = golden["svd_Tref"]
    tgt = orc.transform(src, Tref, mode=1)
    for dbl in (False, True):
        T = orc.estimate_svd(src, tgt, scalar_is_double=dbl)
        assert np.allclose(T[:3, 3], Tref[:3, 3], atol=2e-6)
        assert np.allclose(T[:3, :3], Tref[:3, :3], atol=2e-6)
        corr = np.zeros(src.shape[0], dtype=orc.CORR_DTYPE)
        corr["index_query"] = corr["index_match"] = np.arange(src.shape[0])
        T2 = orc.estimate_svd(src, tgt, corr=corr, scalar_is_double=dbl)
        assert np.array_equal(T, T2)


def test_point_to_plane_lls_paraboloid(orc):
    # test/registration/test_registration_api.cpp:469-518
    xs = np.arange(-5.0, 5.0 + 1e-6, 0.5, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    x, y = X.ravel(), Y.ravel()
    z = np.float32(0.1) * x ** 2 + np.float32(0.2) * x * y - np.float32(0.3) * y + np.float32(1.0)
    n = np.stack([-0.2 * x - 0.2, 0.6 * y - 0.2, np.ones_like(x)], 1).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    src = np.zeros((x.size, 12), np.float32)
    src[:, 0], src[:, 1], src[:, 2], src[:, 3] = x, y, z, 1
    src[:, 4:7] = n
    G = np.array([[0.9938, 0.0988, 0.0517, 0.1], [-0.0997, 0.9949, 0.0149, -0.2],
                  [-0.05, -0.02, 0.9986, 0.3], [0, 0, 0, 1]], np.float64)
    tgt = orc.transform(src, G, mode=1, normal_off=4)
    T, rc = orc.estimate_point_to_plane_lls(src, tgt)
    assert rc == 0
    assert np.all(np.abs(T - G) < 1e-2)


def test_voxelgrid_bun0(golden, orc):
    # test/filters/test_filters.cpp:566-603
    cloud = orc.to_xyz1(golden["bun0"])
    leaf = [0.02, 0.02, 0.02]
    out = orc.voxelgrid(cloud, leaf)
    assert out.shape[0] == int(golden["voxel_count_all"])
    # the filter-field pass ("z" in [0.05, 0.1]) selects points before the grid: emulate with indices
    z = cloud[:, 2]
    # getMinMax3D with a filter field computes min/max over the SELECTED points; equivalent to
    # running the grid on the selected subset (voxel_grid.hpp:614-617, 663-690)
    sel = np.nonzero(~((z > np.float32(0.1)) | (z < np.float32(0.05))))[0].astype(np.int32)
    out = orc.voxelgrid(cloud, leaf, indices=sel)
    assert out.shape[0] == int(golden["voxel_count_z_005_01"])
    assert np.allclose(out[0, :3], golden["voxel_z_first"], atol=1e-4)
    assert np.allclose(out[13, :3], golden["voxel_z_last"], atol=1e-4)
    neg = np.nonzero(~((z < np.float32(0.1)) & (z > np.float32(0.05))))[0].astype(np.int32)
    out = orc.voxelgrid(cloud, leaf, indices=neg)
    assert out.shape[0] == int(golden["voxel_count_z_negative"])
    # overflow guard (voxel_grid.hpp:620-629)
    assert orc.voxelgrid(cloud, [1e-5, 1e-5, 1e-5]) is None


def test_normal_bun0(golden, orc):
    # test/features/test_normal_estimation.cpp:98-163: indices = all points, k = all points
    cloud = orc.to_xyz1(golden["bun0"])
    g = golden["normal_bun0"]
    n, ok = orc.point_normal(cloud, np.arange(cloud.shape[0]))
    assert ok
    assert np.allclose(np.abs(n[:3]), g[:3], atol=1e-4)
    assert abs(n[3] - g[4]) < 1e-4
    normals, dense = orc.Index(cloud).normals_knn(cloud, cloud.shape[0])
    assert dense
    assert np.allclose(normals[:, :3], -g[:3], atol=1e-4)
    assert np.allclose(normals[:, 3], g[4], atol=1e-4)


def test_rejectors_golden(golden, orc):
    # test/registration/test_registration_api.cpp:131-380 + test_registration_api_data.h:461-1119
    c = orc.Index(orc.to_xyz1(golden["bun4"])).correspondences(orc.to_xyz1(golden["bun0"]))
    r, _ = orc.reject(c, orc.REJ_DISTANCE, p=0.01)                       # rej_dist_max_dist = 0.01f
    assert np.array_equal(np.stack([r["index_query"], r["index_match"]], 1), golden["corr_rej_dist"])
    r, med = orc.reject(c, orc.REJ_MEDIAN, p=0.5)                        # rej_median_factor = 0.5
    assert np.array_equal(np.stack([r["index_query"], r["index_match"]], 1), golden["corr_rej_median"])
    assert abs(med - 0.000465391) < 1e-4
    r, _ = orc.reject(c, orc.REJ_ONE_TO_ONE)
    assert np.array_equal(np.stack([r["index_query"], r["index_match"]], 1), golden["corr_rej_one_to_one"])
    r, _ = orc.reject(c, orc.REJ_TRIMMED, p=0.5)                         # rej_trimmed_overlap = 0.5
    assert np.array_equal(np.stack([r["index_query"], r["index_match"]], 1), golden["corr_rej_trimmed"])


def _paraboloid(orc):
    xs = np.arange(-5.0, 5.0 + 1e-6, 0.5, dtype=np.float32)
    X, Y = np.meshgrid(xs, xs, indexing="ij")
    x, y = X.ravel(), Y.ravel()
    z = np.float32(0.1) * x ** 2 + np.float32(0.2) * x * y - np.float32(0.3) * y + np.float32(1.0)
    n = np.stack([-0.2 * x - 0.2, 0.6 * y - 0.2, np.ones_like(x)], 1).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    src = np.zeros((x.size, 12), np.float32)
    src[:, 0], src[:, 1], src[:, 2], src[:, 3] = x, y, z, 1
    src[:, 4:7] = n
    G = np.array([[0.9938, 0.0988, 0.0517, 0.1], [-0.0997, 0.9949, 0.0149, -0.2],
                  [-0.05, -0.02, 0.9986, 0.3], [0, 0, 0, 1]], np.float64)
    return src, orc.transform(src, G, mode=1, normal_off=4), G


def test_symmetric_point_to_plane_lls_paraboloid(orc):
    # test/registration/test_registration_api.cpp:663-713
    src, tgt, G = _paraboloid(orc)
    for dbl in (False, True):
        T, rc = orc.estimate_symmetric_lls(src, tgt, scalar_is_double=dbl)
        assert rc == 0 and np.all(np.abs(T - G) < 1e-2), T


def test_outlier_filters_bun0(golden, orc):
    # test/filters/test_filters.cpp:1494-1515 (RadiusOutlierRemoval) and :1587-1613 (StatisticalOutlierRemoval)
    cloud = orc.to_xyz1(golden["bun0"])
    idx = orc.Index(cloud)
    assert idx.radius_outlier_removal(cloud, 0.02, 14).size == 307
    assert idx.radius_outlier_removal(cloud, 0.02, 14, negative=True).size == 90
    assert idx.radius_outlier_removal(cloud, 0.02, 14, is_dense=False).size == 307
    k = idx.statistical_outlier_removal(cloud, 50, 1.0)
    assert k.size == 352
    assert np.allclose(cloud[k[-1], :3], [-0.034667, 0.15131, -0.00071029], atol=1e-4)
    kn = idx.statistical_outlier_removal(cloud, 50, 1.0, negative=True)
    assert kn.size == 397 - 352
    assert np.allclose(cloud[kn[-1], :3], [-0.07793, 0.17516, -0.0444], atol=1e-4)


# ---------------------------------------------------------------------------------------------------------------------
# normal-based correspondence estimators, surface-normal rejector, radius-search normals
# ---------------------------------------------------------------------------------------------------------------------
def _point_normal_rows(xyz, normals=None):
    out = np.zeros((xyz.shape[0], 12), np.float32)
    out[:, :3] = xyz[:, :3]
    out[:, 3] = 1
    if normals is not None:
        out[:, 4:4 + normals.shape[1]] = normals
    return out


def test_normal_shooting_reference_planes(orc):
    """test/registration/test_correspondence_estimation.cpp:95-137: two parallel planes, normals from k = 5,
    CorrespondenceEstimationNormalShooting with k = 10 must pair every point with its counterpart."""
    ii, jj = np.meshgrid(np.arange(50), np.arange(25), indexing="ij")
    x = ii.ravel().astype(np.float32) * np.float32(0.2)
    z = jj.ravel().astype(np.float32) * np.float32(0.2)
    c1 = _point_normal_rows(np.stack([x, np.zeros_like(x), z], 1))
    c2 = _point_normal_rows(np.stack([x, np.full_like(x, 2), z], 1))
    nrm, dense = orc.Index(c1).normals_knn(c1, 5)
    assert dense and np.allclose(np.abs(nrm[:, :3]), [0, 1, 0], atol=1e-6)  # "All normals are perpendicular to the plane"
    c1[:, 4:8] = nrm
    c2[:, 4:8] = nrm
    it = orc.Index(c2)
    for kind in (orc.CORR_NORMAL_SHOOTING, orc.CORR_BACK_PROJECTION):
        c = it.correspondences_normals(kind, c1, c2, k=10)
        assert len(c) == 1250 and np.array_equal(c["index_query"], c["index_match"])
        assert np.all(c["distance"] == np.float32(4.0))  # the stored distance is the squared POINT distance (:126)
    # the gate compares the squared line distance with max_distance itself (:121): 0 passes any gate, a tilted normal fails
    c1[:, 4:7] = np.float32([0.6, 0.8, 0.0])
    assert len(it.correspondences_normals(orc.CORR_NORMAL_SHOOTING, c1, c2, k=10, max_distance=1e-3)) == 0


def test_surface_normal_rejector_and_icp_with_normal_shooting(golden, orc):
    """test_registration_api.cpp:266-317 (rejector) and test_registration.cpp:511-560 (ICP whose correspondences come from
    normal shooting, filtered by the surface-normal rejector with threshold 0; accepted when the fitness score < 0.005)."""
    b0, b4 = _point_normal_rows(golden["bun0"]), _point_normal_rows(golden["bun4"])
    i0, i4 = orc.Index(b0), orc.Index(b4)
    b0[:, 4:8] = i0.normals_knn(b0, 10)[0]
    b4[:, 4:8] = i4.normals_knn(b4, 10)[0]
    corr = i4.correspondences(b0)
    kept = orc.reject_surface_normal(corr, b0[:, 4:], b4[:, 4:], 0.5)
    dots = (b0[corr["index_query"], 4:7] * b4[corr["index_match"], 4:7]).sum(1)
    assert 0 < len(kept) < len(corr) and abs(len(kept) - int((dots > 0.5).sum())) <= 1  # fp32 vs numpy summation order
    assert np.array_equal(orc.reject_surface_normal(corr, b0[:, 4:], b4[:, 4:], -2.0), corr)
    assert len(orc.reject_surface_normal(corr, b0[:, 4:], b4[:, 4:], 1.5)) == 0
    for kind in (orc.CORR_NORMAL_SHOOTING, orc.CORR_BACK_PROJECTION):
        for dbl in (False, True):
            r = orc.icp_align_rejectors(b0, b4, [(orc.REJ_SURFACE_NORMAL, 0.0, 0)], estimator=1, source_has_normals=True,
                                        scalar_is_double=dbl, correspondence_kind=kind, correspondence_k=10,
                                        max_iterations=50, transformation_epsilon=1e-8)
            assert r["converged"] and r["iterations"] < 50
            assert i4.fitness_score(b0, r["final"], scalar_is_double=dbl) < 0.005


def test_radius_normals(golden, orc):
    """A radius that holds the whole cloud is NormalEstimation with k = all points, whose result the reference pins
    (test/features/test_normal_estimation.cpp:128-163); a radius with < 3 neighbours gives NaN and is_dense = false."""
    b0 = orc.to_xyz1(golden["bun0"])
    idx = orc.Index(b0)
    n, dense = idx.normals_radius(b0, 10.0)
    g = golden["normal_bun0"]
    assert dense
    assert np.allclose(n[:, :3], -g[:3], atol=1e-4) and np.allclose(n[:, 3], g[4], atol=1e-4)
    nk, _ = idx.normals_knn(b0, 397)
    assert np.array_equal(n, nk)  # same neighbour lists in the same (d2, index) order -> identical arithmetic
    n, dense = idx.normals_radius(b0, 1e-4)
    assert not dense and np.isnan(n).all()
    # radius lists and radius normals agree: recompute one normal from its radius list
    offs, ids, _ = idx.radius(b0[:5], 0.02)
    nr, _ = idx.normals_radius(b0[:5], 0.02)
    for i in range(5):
        lst = ids[offs[i]:offs[i + 1]]
        pn, ok = orc.point_normal(b0, lst)
        assert ok and np.allclose(np.abs(pn), np.abs(nr[i]), atol=0)


def test_cluster_labels_are_connected_components(orc):
    """extractEuclideanClusters (extract_clusters.hpp:124-223) restated as a flood fill; an independent
    scipy connected-components labelling of the same d2 < r2 graph must give the same partition."""
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    from scipy.spatial import cKDTree
    import pcl_b200 as P  # only the host-side grouping helper (no device call)
    rng = np.random.default_rng(3)
    c = orc.to_xyz1(rng.random((20000, 3), dtype=np.float32))
    c[::97, 0] = np.nan
    ok = np.isfinite(c[:, 0])
    idx = np.nonzero(ok)[0]
    tree = cKDTree(c[ok, :3].astype(np.float64))
    for tol, n_expected in ((0.0, idx.size), (0.035, None), (0.05, None)):
        lab = orc.Index(c).cluster_labels(tol)
        assert np.array_equal(lab >= 0, ok)
        assert np.all(lab[idx] <= idx)  # the label is the smallest index of the cluster
        pairs = tree.query_pairs(float(np.float32(tol)) * 1.000001 + 1e-12, output_type="ndarray")
        d = c[idx[pairs[:, 0]], :3] - c[idx[pairs[:, 1]], :3]
        d2 = ((d[:, 0] * d[:, 0]) + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]
        keep = d2 < np.float32(np.float64(np.float32(tol)) ** 2)
        g = coo_matrix((np.ones(int(keep.sum())), (pairs[keep, 0], pairs[keep, 1])), shape=(idx.size, idx.size))
        nc, l = connected_components(g, directed=False)
        first = {}
        assert all(first.setdefault(a, b) == b for a, b in zip(l, lab[idx])) and len(set(first.values())) == nc
        if n_expected is not None:
            assert nc == n_expected
        cl = P.clusters_from_labels(lab, 3, 100)
        assert all(3 <= a.size <= 100 for a in cl) and [a.size for a in cl] == sorted((a.size for a in cl), reverse=True)
