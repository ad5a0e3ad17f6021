#!/usr/bin/env python
"""k-NN / normals stage times on bench_stages.py's surface cloud (10 M points), for comparing builds:
PCLB200_LIB=pcl_b200/libpclb200_<tag>.so python tools/knn_times.py"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
    return best * 1e3


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
    ctx = P.Context(0)
    g = torch.Generator(device="cuda").manual_seed(5)
    xy = torch.rand((n, 2), generator=g, device="cuda") * 10
    surf = torch.ones((n, 4), device="cuda")
    surf[:, :2] = xy
    surf[:, 2] = 0.5 * torch.sin(xy[:, 0]) * torch.cos(0.7 * xy[:, 1]) + 0.002 * torch.randn(n, generator=g, device="cuda")
    si = P.Index(ctx, surf)
    q = surf[torch.randperm(n, generator=g, device="cuda")].contiguous()
    nrm = torch.empty((n, 4), device="cuda")
    row = {"lib": os.environ.get("PCLB200_LIB", "default")}
    for k in (16, 32):
        oi = torch.empty((n, k), dtype=torch.int32, device="cuda")
        od = torch.empty((n, k), dtype=torch.float32, device="cuda")
        row[f"knn{k}_ms"] = round(timed(ctx, lambda: si.knn(q, k, oi, od)), 3)
        row[f"knn{k}_checksum"] = int(oi.sum(dtype=torch.int64))
        del oi, od
    row["normals16_ms"] = round(timed(ctx, lambda: si.normals_knn(surf, 16, (5, 5, 10), out=nrm)), 3)
    row["normals16_checksum"] = float(nrm[:, :3].double().abs().sum())
    print(json.dumps(row))


if __name__ == "__main__":
    main()
