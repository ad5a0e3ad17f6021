"""Euclidean clustering through the C-ABI against the CPU restatement of pcl::extractEuclideanClusters:
component labels bit-exact.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    import pcl_b200
    pcl_b200.lib()
    ctx = pcl_b200.Context(0)
    yield pcl_b200, ctx
    ctx.close()


def _cases(rng):
    u = rng.random((20000, 3), dtype=np.float32)
    for tol in (0.0, 0.02, 0.035, 0.05):  # isolated points ... around the percolation threshold ... one giant component
        yield f"uniform_{tol}", u, tol
    g = np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3)
    yield "grid_at_the_tolerance", g, 1.0          # d2 == r2 is NOT an edge (strict test): 1728 singletons
    yield "grid_above_the_tolerance", g, 1.0001    # one component
    yield "grid_diagonals", g, 1.5                 # sqrt(2) neighbours join, sqrt(3) ones are redundant
    d = rng.random((700, 3), dtype=np.float32)
    yield "duplicates", np.concatenate([d, d, d[:300]]), 0.01
    blobs = np.concatenate([rng.normal(c, 0.02, (400, 3)) for c in rng.random((25, 3)) * 4]).astype(np.float32)
    yield "blobs", blobs, 0.03
    yield "tiny", rng.random((5, 3), dtype=np.float32), 0.5
    yield "single", rng.random((1, 3), dtype=np.float32), 0.5
    line = np.stack([np.arange(3000, dtype=np.float32) * np.float32(0.01)] * 3, 1)  # a chain: deep union-find paths
    yield "chain", line, 0.018


def test_cluster_labels_bit_exact(gpu, orc):
    P, ctx = gpu
    rng = np.random.default_rng(31)
    for name, pts, tol in _cases(rng):
        cloud = orc.to_xyz1(pts)
        g = P.Index(ctx, cloud).cluster_labels(tol)
        o = orc.Index(cloud).cluster_labels(tol)
        assert np.array_equal(g, o), (name, np.argwhere(g != o)[:5].ravel())
    assert len(set(P.Index(ctx, orc.to_xyz1(np.stack(np.meshgrid(*[np.arange(12, dtype=np.float32)] * 3, indexing="ij"), -1).reshape(-1, 3))).cluster_labels(1.0))) == 1728


def test_cluster_nan_points_subset_and_size_window(gpu, orc):
    P, ctx = gpu
    rng = np.random.default_rng(32)
    pts = rng.random((30000, 3), dtype=np.float32)
    cloud = orc.to_xyz1(pts)
    cloud[::53, 1] = np.nan
    sub = np.sort(rng.permutation(30000)[:18000]).astype(np.int32)
    for subset in (None, sub):
        gi, oi = P.Index(ctx, cloud, subset), orc.Index(cloud, subset)
        g, o = gi.cluster_labels(0.03), oi.cluster_labels(0.03)
        assert np.array_equal(g, o)
        held = np.isfinite(cloud[:, 1]) if subset is None else np.isin(np.arange(30000), sub) & np.isfinite(cloud[:, 1])
        assert np.array_equal(g >= 0, held)
        cl = gi.euclidean_clusters(0.03, min_size=5, max_size=200)
        ref = P.clusters_from_labels(o, 5, 200)
        assert len(cl) == len(ref) > 0 and all(np.array_equal(a, b) for a, b in zip(cl, ref))
        sizes = [c.size for c in cl]
        assert sizes == sorted(sizes, reverse=True) and 5 <= min(sizes) and max(sizes) <= 200
        assert all(np.all(np.diff(c) > 0) for c in cl)


def test_cluster_large_surface(gpu, orc):
    """1 M points of a noisy sheet with holes punched into it: the component structure of a real scan."""
    P, ctx = gpu
    rng = np.random.default_rng(33)
    n = 1_000_000
    xy = rng.random((n, 2), dtype=np.float32) * 10
    keep = (np.sin(xy[:, 0] * 3) * np.cos(xy[:, 1] * 2.5)) < 0.55
    pts = np.zeros((keep.sum(), 3), np.float32)
    pts[:, :2] = xy[keep]
    pts[:, 2] = 0.3 * np.sin(pts[:, 0]) + rng.normal(0, 0.002, keep.sum()).astype(np.float32)
    cloud = orc.to_xyz1(pts)
    g = P.Index(ctx, cloud).cluster_labels(0.02)
    o = orc.Index(cloud).cluster_labels(0.02)
    assert np.array_equal(g, o)
    dev = np.empty(cloud.shape[0], np.int32)
    import torch
    t = torch.empty(cloud.shape[0], dtype=torch.int32, device="cuda")
    P.Index(ctx, cloud).cluster_labels(0.02, out=t)  # device-resident output
    assert np.array_equal(t.cpu().numpy(), o)


def test_cluster_argument_errors(gpu, orc):
    P, ctx = gpu
    cloud = orc.to_xyz1(np.random.default_rng(1).random((100, 3), dtype=np.float32))
    idx = P.Index(ctx, cloud)
    with pytest.raises(P.Pclb200Error) as e:
        idx.cluster_labels(-1.0)
    assert e.value.code == P.ERR_INVALID
    with pytest.raises(P.Pclb200Error) as e:
        P._check(P.lib().pclb200_cluster_labels(ctx.h, idx.h, 0.1, np.empty(50, np.int32).ctypes.data, 50))
    assert e.value.code == P.ERR_INVALID
