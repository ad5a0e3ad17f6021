"""The oracle's kd-tree against its own brute force (the reference's strategy: BruteForce is ground
truth, test/kdtree/test_kdtree.cpp:92-207, test/search/test_search.cpp) — including exact ties,
duplicates, NaNs and index subsets.  CPU-only."""
import numpy as np
import pytest


def _clouds(rng):
    yield "uniform", rng.random((3000, 3), dtype=np.float32)
    g = np.stack(np.meshgrid(*[np.arange(11, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    yield "grid_ties", g  # 11^3 lattice: massive exact ties
    d = rng.random((500, 3), dtype=np.float32)
    yield "duplicates", np.concatenate([d, d, d[:100]])
    yield "collinear", np.stack([np.linspace(0, 1, 777, dtype=np.float32)] * 3, 1)


@pytest.mark.parametrize("k", [1, 5, 16])
def test_tree_equals_bruteforce(orc, k):
    rng = np.random.default_rng(5)
    for name, pts in _clouds(rng):
        cloud = orc.to_xyz1(pts)
        q = orc.to_xyz1(np.concatenate([pts[:200], rng.random((200, 3), dtype=np.float32) * pts.max()]))
        i1, d1, k1 = orc.Index(cloud).knn(q, k, nthreads=2)
        i2, d2, k2 = orc.knn_bruteforce(cloud, q, k)
        assert k1 == k2, name
        assert np.array_equal(i1, i2), name
        assert np.array_equal(d1, d2), name


def test_nan_points_and_subset(orc):
    rng = np.random.default_rng(6)
    pts = rng.random((1000, 3), dtype=np.float32)
    pts[::7, 1] = np.nan
    pts[5, 0] = np.inf
    cloud = orc.to_xyz1(pts)
    idx = orc.Index(cloud)
    valid = np.isfinite(pts).all(1)
    assert idx.size == int(valid.sum())
    q = orc.to_xyz1(rng.random((100, 3), dtype=np.float32))
    i1, d1, _ = idx.knn(q, 3)
    assert valid[i1].all()  # original indices, never a dropped point (kdtree_flann.hpp:445-458)
    i2, d2, _ = orc.knn_bruteforce(cloud, q, 3)
    assert np.array_equal(i1, i2)
    sub = np.arange(0, 1000, 3, dtype=np.int32)
    i3, _, _ = orc.Index(cloud, subset=sub).knn(q, 2)
    assert np.isin(i3, sub).all() and valid[i3].all()


def test_radius_matches_bruteforce(orc):
    rng = np.random.default_rng(7)
    pts = rng.random((2000, 3), dtype=np.float32)
    cloud = orc.to_xyz1(pts)
    q = cloud[:150]
    r = 0.1
    offs, idx, d2 = orc.Index(cloud).radius(q, r)
    r2 = np.float32(r * r)
    for i in range(q.shape[0]):
        d = ((pts - pts[i]) ** 2)
        dd = (d[:, 0] + d[:, 1]) + d[:, 2]
        want = np.nonzero(dd < r2)[0]
        want = want[np.lexsort((want, dd[want]))]
        assert np.array_equal(idx[offs[i]:offs[i + 1]], want)
    # max_nn keeps the nearest max_nn (kdtree_flann.hpp:382-391)
    o2, i2, _ = orc.Index(cloud).radius(q, r, max_nn=3)
    for i in range(q.shape[0]):
        assert np.array_equal(i2[o2[i]:o2[i + 1]], idx[offs[i]:offs[i + 1]][:3])


def test_correspondence_gate_and_threads(orc):
    rng = np.random.default_rng(8)
    tgt = orc.to_xyz1(rng.random((5000, 3), dtype=np.float32))
    src = orc.to_xyz1(rng.random((4000, 3), dtype=np.float32))
    t = orc.Index(tgt)
    c1 = t.correspondences(src, max_distance=0.03, nthreads=1)
    c4 = t.correspondences(src, max_distance=0.03, nthreads=4)
    assert np.array_equal(c1, c4)  # ordered by index_query regardless of threads (:193-216)
    assert 0 < c1.size < 4000
    assert np.all(c1["distance"] <= np.float32(0.03) ** 2 * (1 + 1e-6))
    assert np.all(np.diff(c1["index_query"]) > 0)
