import sys, numpy as np, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench, pcl_b200 as P
n = 10_000_000
ctx = P.Context(0)
tgt = bench.make_target(n); src = bench.make_source(n, 0)
tidx = P.Index(ctx, tgt)
nrm = np.empty((n, 4), np.float32); tidx.normals_knn(tgt, 16, viewpoint=(5, 5, 10), out=nrm)
s = P.Icp(ctx, max_iterations=30, max_correspondence_distance=0.05, estimator=P.EST_POINT_TO_PLANE_LLS, with_normals_transform=1, mse_threshold_absolute=0.0)
s.set_target(tidx, normals=nrm)
s.set_source(src)
prev = 0
ctx.profile(True)
for it in range(30):
    ctx.profile_reset()
    st = s.iterate(1)
    ms, _ = ctx.profile_get("icp_search")
    T = st["last"]
    ang = np.degrees(np.arccos(min(1.0, (np.trace(T[:3, :3]) - 1) / 2)))
    print(it + 1, "skipped", st["total_skipped_walks"] - prev, "ncorr", st["n_correspondences"], "search_ms %.2f" % ms, "rot_deg %.2e" % ang, "trans %.2e" % np.linalg.norm(T[:3, 3]), "mse %.3e" % st["mse"])
    prev = st["total_skipped_walks"]
    if st["state"] != 0: break
