"""Other consumers of the searcher: TransformationValidationEuclidean and the inlier count /
fitness of SampleConsensusPrerejective — one batch 1-NN on the device index each.  Needs a B200: run with -m gpu."""
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


@pytest.fixture(scope="module")
def orc():
    import oracle
    oracle.build()
    return oracle


def _scene(seed=4, n=40000):
    rng = np.random.default_rng(seed)
    tgt = rng.random((n, 3), dtype=np.float32)
    tgt[:, 2] = np.float32(0.2) * np.sin(np.float32(5) * tgt[:, 0])
    a = np.deg2rad(1.5)
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    src = (tgt[::3].astype(np.float64) @ R.T + [0.004, -0.003, 0.002]).astype(np.float32)
    T = np.eye(4)
    T[:3, :3] = R.T * 0.9999 + 1e-5   # a pose hypothesis close to the inverse motion, deliberately not orthonormal
    T[:3, 3] = [-0.0035, 0.0031, -0.0018]
    return src, tgt, T


def test_transformation_validation_euclidean(gpu, orc):
    """registration/impl/transformation_validation_euclidean.hpp:50-109, float and double Scalar."""
    P, ctx = gpu
    src, tgt, T = _scene()
    S, Tg = P.xyz1(src), P.xyz1(tgt)
    idx = P.Index(ctx, Tg)
    oidx = orc.Index(Tg)
    for dbl in (False, True):
        M = T if dbl else T.astype(np.float32)
        x, y, z = (S[:, k].astype(M.dtype) for k in range(3))
        moved = np.ones_like(S)
        for r in range(3):  # T(r,0)*x + T(r,1)*y + T(r,2)*z + T(r,3), left to right in Scalar, cast to float (:62-75)
            moved[:, r] = (((M[r, 0] * x + M[r, 1] * y) + M[r, 2] * z) + M[r, 3]).astype(np.float32)
        _, d2, _ = oidx.knn(moved, 1, nthreads=4)
        d2 = d2[:, 0].astype(np.float64)
        for max_range in (np.finfo(np.float64).max, float(np.percentile(d2, 60)), 0.0):
            keep = d2 <= max_range
            want = d2[keep].sum() / keep.sum() if keep.any() else np.finfo(np.float64).max
            got = idx.validate_transformation(S, T, max_range=max_range, scalar_is_double=dbl)
            assert got == pytest.approx(want, rel=1e-12), (dbl, max_range)


def test_sample_consensus_prerejective_inliers(gpu, orc):
    """impl/sample_consensus_prerejective.hpp:308-347: inlier list (strict <) and the float fitness, bit for bit."""
    P, ctx = gpu
    src, tgt, T = _scene(seed=9)
    S, Tg = P.xyz1(src), P.xyz1(tgt)
    S[7, 0] = np.nan
    idx = P.Index(ctx, Tg)
    oidx = orc.Index(Tg)
    moved = orc.transform(S, T.astype(np.float32).astype(np.float64), mode=1)   # pcl::transformPointCloud, float
    fin = np.isfinite(moved[:, :3]).all(1)
    _, d2, _ = oidx.knn(moved[fin], 1, nthreads=4)
    d2_all = np.full(S.shape[0], np.inf, dtype=np.float32)
    d2_all[fin] = d2[:, 0]
    for thr in (0.002, float(np.sqrt(np.median(d2_all[fin]))), 1e-9):
        max_range = np.float32(thr) * np.float32(thr)
        want = np.nonzero(d2_all < max_range)[0]
        fit = np.float32(0)
        for d in d2_all[want]:
            fit = np.float32(fit + d)
        want_fit = float(np.float32(fit / np.float32(want.size))) if want.size else float(np.finfo(np.float32).max)
        inl, got_fit = idx.inliers(S, T, thr)
        assert np.array_equal(inl, want), thr
        assert got_fit == want_fit, thr


def test_gicp_covariances(gpu, orc):
    """registration/impl/gicp.hpp:69-147: k = 20 neighbourhood covariance, regularised to singular values (1, 1, eps)."""
    P, ctx = gpu
    rng = np.random.default_rng(2)
    n = 30000
    pts = rng.random((n, 3), dtype=np.float32) * np.float32(3)
    pts[:, 2] = np.float32(0.2) * np.sin(np.float32(4) * pts[:, 0]) + np.float32(0.003) * rng.standard_normal(n).astype(np.float32)
    cloud = P.xyz1(pts)
    cloud[11, 1] = np.nan
    g = P.Index(ctx, cloud).gicp_covariances(cloud, k=20, gicp_epsilon=0.001)
    o = orc.Index(cloud).gicp_covariances(k=20, gicp_epsilon=0.001, nthreads=8)
    assert np.all(g[11] == 0) and np.all(o[11] == 0)
    ok = np.isfinite(cloud[:, 1])
    # the regularised matrix is I - (1 - eps) n n^T: well conditioned wherever the smallest singular value is isolated
    assert np.abs(g[ok] - o[ok]).max() < 1e-9, float(np.abs(g[ok] - o[ok]).max())
    w = np.linalg.eigvalsh(g[ok])
    assert np.allclose(w[:, 0], 0.001, atol=1e-9) and np.allclose(w[:, 1:], 1.0, atol=1e-9)
