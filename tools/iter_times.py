#!/usr/bin/env python3
"""Per-iteration device times of the ICP loop on bench.py's workload (one align, stepped one iteration at a time).
With a -DPCLB_STATS build (PCLB200_LIB=pcl_b200/libpclb200_stats.so) also prints the walk event counts per query."""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import pcl_b200 as P  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    track = int(sys.argv[3]) if len(sys.argv) > 3 else P.TRACK_AUTO
    L = P.lib()
    stats_fn = getattr(L, "pclb200_debug_walk_stats", None) if hasattr(L, "pclb200_debug_walk_stats") else None
    hist_fn = getattr(L, "pclb200_debug_walk_hist", None) if hasattr(L, "pclb200_debug_walk_hist") else None
    ctx = P.Context(0)
    dev = torch.device("cuda", 0)
    tgt = bench.make_target(n)
    ctx.profile(True)
    tidx = P.Index(ctx, tgt)
    tgt_dev = torch.from_numpy(tgt).to(dev)
    nrm = torch.empty((n, 4), dtype=torch.float32, device=dev)
    tidx.normals_knn(tgt_dev, 16, viewpoint=(5.0, 5.0, 10.0), out=nrm)
    print(json.dumps({"index": tidx.stats, "build_ms": ctx.profile_get("index_build")[0], "normals_ms": ctx.profile_get("normals")[0]}))
    src = torch.from_numpy(bench.make_source(n, 0)).to(dev)
    params = P.default_params(max_iterations=iters, max_correspondence_distance=bench.MAX_CORR_DIST,
                              estimator=P.EST_POINT_TO_PLANE_LLS, with_normals_transform=1, mse_threshold_absolute=0.0,
                              track_mode=track)
    icp = P.Icp(ctx, params=params)
    icp.set_target(tidx, normals=nrm)
    names = ["lookups", "nodes", "leaves", "pushes", "home_seeds", "rooted", "cells_pushed", "walks"]
    for rep in range(2):
        icp.set_source(src, normals=P.Field(src, 4))
        if stats_fn:
            buf = (C.c_ulonglong * 8)()
            stats_fn(buf, 1)
        if hist_fn:
            hb = (C.c_ulonglong * 66)()
            hist_fn(hb, 1)
        for it in range(iters):
            ctx.profile_reset()
            st = icp.iterate(1)
            row = {"rep": rep, "iter": it, "search_ms": round(ctx.profile_get("icp_search")[0], 4),
                   "accum_ms": round(ctx.profile_get("icp_accum")[0], 4), "n_corr": st["n_correspondences"],
                   "mse": st["mse"], "skipped": st["total_skipped_walks"]}
            if stats_fn:
                stats_fn(buf, 1)
                w = max(buf[7], 1)
                row.update({k: round(buf[i] / w, 3) for i, k in enumerate(names[:7])})
                row["walks"] = buf[7]
            if hist_fn:
                hist_fn(hb, 1)
                row["node_lane_utilisation"] = round(hb[0] / max(hb[1], 1), 3)
                h = np.array(hb[2:], dtype=np.float64)
                c = np.cumsum(h) / max(h.sum(), 1)
                row["node_visits_p50_p90_p99"] = [int(np.searchsorted(c, q)) for q in (0.5, 0.9, 0.99)]
            if rep == 1:
                print(json.dumps(row))
            if st["state"] != 0:
                break


if __name__ == "__main__":
    main()
