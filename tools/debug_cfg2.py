#!/usr/bin/env python3
"""Config 2 at its named size, GPU loop vs the oracle's loop iteration by iteration: where do the two trajectories part?
Per iteration: (a) GPU correspondences vs an oracle search on the GPU's own cloud (search exactness), (b) GPU T_k vs the
oracle's estimate from the same pairs (solve), (c) |final_gpu - final_oracle| of the two free-running loops."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle as orc  # noqa: E402
import pcl_b200 as P  # noqa: E402


def rot(axis, deg):
    axis = np.asarray(axis, dtype=np.float64)
    axis /= np.linalg.norm(axis)
    a = np.deg2rad(deg)
    K = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(a) * K + (1 - np.cos(a)) * K @ K


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    track = int(sys.argv[3]) if len(sys.argv) > 3 else P.TRACK_AUTO
    ctx = P.Context(0)
    tgt = np.random.default_rng(42).random((n, 3), dtype=np.float32)
    src = (tgt.astype(np.float64) @ rot([1, 1, 1], 5.0).T + [0.01, -0.02, 0.015] +
           np.random.default_rng(43).normal(0, 0.001, (n, 3))).astype(np.float32)
    leaf = 0.01
    vt, vs = ctx.voxelgrid(P.xyz1(tgt), leaf), ctx.voxelgrid(P.xyz1(src), leaf)
    nt = os.cpu_count() or 8
    oidx = orc.Index(vt)
    s = P.Icp(ctx, max_iterations=iters, transformation_epsilon=1e-8, max_correspondence_distance=0.05, track_mode=track)
    s.set_target(P.Index(ctx, vt))
    s.set_source(vs)
    cloud_g = vs.copy()
    cloud_o = vs.copy()
    final_o = np.eye(4)
    for it in range(iters):
        st = s.iterate(1)
        g = s.get_correspondences()
        o_on_g = oidx.correspondences(cloud_g, max_distance=0.05, nthreads=nt)
        same = g.size == o_on_g.size and np.array_equal(g, o_on_g)
        T_same_pairs = orc.estimate_svd(cloud_g, vt, g)
        o = oidx.correspondences(cloud_o, max_distance=0.05, nthreads=nt)
        T_o = orc.estimate_svd(cloud_o, vt, o)
        final_o = T_o.astype(np.float32).astype(np.float64) @ final_o
        final_o = final_o.astype(np.float32).astype(np.float64)
        cloud_o = orc.transform(cloud_o, T_o, mode=0)
        cloud_g = orc.transform(cloud_g, st["last"], mode=0)
        chk = np.empty_like(vs)
        s.get_cloud(chk)
        idx_diff = int((g["index_match"] != o["index_match"]).sum()) if g.size == o.size else -1
        print(json.dumps({"it": it, "search_exact": bool(same), "n": int(g.size),
                          "solve_dT": float(np.linalg.norm(st["last"] - T_same_pairs)),
                          "traj_dT_last": float(np.linalg.norm(st["last"] - T_o)),
                          "traj_dfinal": float(np.linalg.norm(st["final"] - final_o)),
                          "pairs_differ": idx_diff, "cloud_maxdiff_vs_oracle_transform": float(np.abs(chk[:, :3] - cloud_g[:, :3]).max()),
                          "state": st["state"], "skipped": st["total_skipped_walks"]}), flush=True)
        if st["state"] != 0:
            break


if __name__ == "__main__":
    main()
