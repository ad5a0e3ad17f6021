#!/usr/bin/env python
"""Per-stage throughput of the operators around the ICP loop (index build, k-NN, radiusSearch with and
without max_nn, normals, VoxelGrid, normal-shooting correspondences) on one B200.  Not the headline bench (bench.py);
writes one JSON object per stage to stdout.  Inputs are torch CUDA tensors (resident in HBM) unless a stage's C-ABI
returns host arrays (radius), in which case the D2H of the result is part of the time and said so."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import pcl_b200 as P


def timed(ctx, fn, reps=3):
    fn()
    ctx.synchronize()
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        ctx.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=10_000_000)
    a = ap.parse_args()
    n = a.n
    ctx = P.Context(0)
    g = torch.Generator(device="cuda").manual_seed(5)
    out = []

    def emit(stage, seconds, units, unit, **kw):
        rec = dict(stage=stage, ms=round(seconds * 1e3, 3), rate=units / seconds, unit=unit + "/s", n=n, **kw)
        out.append(rec)
        print(json.dumps(rec), flush=True)

    # surface cloud of bench.py's workload (10 M points on a 10 x 10 sheet) and a volume cloud for radius search
    xy = torch.rand((n, 2), generator=g, device="cuda") * 10
    surf = torch.ones((n, 4), device="cuda")
    surf[:, :2] = xy
    surf[:, 2] = 0.5 * torch.sin(xy[:, 0]) * torch.cos(0.7 * xy[:, 1]) + 0.002 * torch.randn(n, generator=g, device="cuda")
    vol = torch.ones((n, 4), device="cuda")
    vol[:, :3] = torch.rand((n, 3), generator=g, device="cuda") * 20
    holder = {}

    emit("index_build_surface", timed(ctx, lambda: holder.__setitem__("s", P.Index(ctx, surf))), n, "points")
    emit("index_build_volume", timed(ctx, lambda: holder.__setitem__("v", P.Index(ctx, vol))), n, "points")
    si, vi = holder["s"], holder["v"]
    q = surf[torch.randperm(n, generator=g, device="cuda")].contiguous()
    nrm = torch.empty((n, 4), device="cuda")
    for walk in ("default",):
        for k in (1, 10, 16, 32):
            oi = torch.empty((n, k), dtype=torch.int32, device="cuda")
            od = torch.empty((n, k), dtype=torch.float32, device="cuda")
            emit(f"knn_k{k}_self_surface_{walk}", timed(ctx, lambda: si.knn(q, k, oi, od)), n, "queries", k=k, walk=walk)
            del oi, od
        emit(f"normals_knn16_surface_{walk}", timed(ctx, lambda: si.normals_knn(surf, 16, (5, 5, 10), out=nrm)), n, "points", walk=walk)
        emit(f"knn_k10_volume_{walk}", timed(ctx, lambda: vi.knn(vol[:2_000_000], 10), reps=2), 2_000_000, "queries", k=10, walk=walk,
             includes="D2H of the rows")
    # radius search: density n/8000 per unit volume; r so that a ball holds ~30 points
    r = float((30.0 / (n / 8000.0) * 3 / (4 * np.pi)) ** (1 / 3))
    nq = min(n, 2_000_000)
    qv = vol[:nq].contiguous()
    res = {}
    ctx.profile(True)
    for name, scope, kw in (("radius_unlimited_volume", "radius", {}), ("radius_maxnn32_volume", "radius_knn", dict(max_nn=32))):
        t = timed(ctx, lambda: res.__setitem__("r", vi.radius(qv, r, **kw)), reps=2)
        ctx.profile_reset()
        vi.radius(qv, r, **kw)
        dev_ms, _ = ctx.profile_get(scope)  # the search itself (count + fill + sort, or the bounded k-NN) on the device
        emit(name, t, nq, "queries", radius=r, neighbours=int(res["r"][0][-1]), includes="D2H of the lists into malloc'd host arrays",
             device_search_ms=round(dev_ms, 3), device_rate=nq / (dev_ms * 1e-3) if dev_ms > 0 else None)
    ctx.profile(False)
    emit("normals_radius_volume", timed(ctx, lambda: vi.normals_radius(qv, r, out=nrm[:nq]), reps=2), nq, "points", radius=r)
    lab = torch.empty(n, dtype=torch.int32, device="cuda")
    emit("cluster_labels_volume", timed(ctx, lambda: vi.cluster_labels(r * 0.6, out=lab), reps=2), n, "points", tolerance=r * 0.6)
    emit("cluster_labels_surface", timed(ctx, lambda: si.cluster_labels(0.004, out=lab), reps=2), n, "points", tolerance=0.004)
    vg = torch.empty((n, 4), device="cuda")
    emit("voxelgrid_leaf0.05_volume", timed(ctx, lambda: ctx.voxelgrid(vol, 0.05, out=vg)), n, "points")
    emit("voxelgrid_leaf0.01_surface", timed(ctx, lambda: ctx.voxelgrid(surf, 0.01, out=vg)), n, "points")
    # normal shooting: surface against itself re-sampled, k = 10
    pn = torch.zeros((n, 12), device="cuda")
    pn[:, :4] = surf
    pn[:, 4:8] = nrm
    m = min(n, 2_000_000)
    src = pn[:m].contiguous()
    src[:, 0] += 0.003
    pi = P.Index(ctx, pn)
    emit("corr_normal_shooting_k10", timed(ctx, lambda: pi.correspondences_normals(1, src, P.Field(src, 4), None, k=10), reps=2),
         m, "queries", includes="D2H of the correspondences")
    emit("corr_nearest", timed(ctx, lambda: pi.correspondences(src), reps=2), m, "queries", includes="D2H of the correspondences")
    return out


if __name__ == "__main__":
    main()
