#!/usr/bin/env python
"""Exactness of the k-NN lists on a dense surface cloud (the density of tools/knn_times.py) against the oracle, for a
given build: PCLB200_LIB=... python tools/knn_check.py [n] [n_queries]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle as orc
import pcl_b200 as P


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    nq = int(sys.argv[2]) if len(sys.argv) > 2 else 300_000
    side = float(np.sqrt(n / 100_000.0))
    ctx = P.Context(0)
    g = torch.Generator(device="cuda").manual_seed(5)
    xy = torch.rand((n, 2), generator=g, device="cuda") * side
    surf = torch.ones((n, 4), device="cuda")
    surf[:, :2] = xy
    surf[:, 2] = 0.5 * torch.sin(xy[:, 0]) * torch.cos(0.7 * xy[:, 1]) + 0.002 * torch.randn(n, generator=g, device="cuda")
    si = P.Index(ctx, surf)
    q = surf[torch.randperm(n, generator=g, device="cuda")[:nq]].contiguous()
    h_surf, h_q = surf.cpu().numpy(), q.cpu().numpy()
    oidx = orc.Index(h_surf)
    row = {"lib": os.environ.get("PCLB200_LIB", "default"), "n": n, "nq": nq}
    ks = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [16, 32]
    for k in ks:
        oi = torch.empty((nq, k), dtype=torch.int32, device="cuda")
        od = torch.empty((nq, k), dtype=torch.float32, device="cuda")
        si.knn(q, k, oi, od)
        ctx.synchronize()
        ri, rd = oidx.knn(h_q, k, nthreads=os.cpu_count() or 8)[:2]
        gi, gd = oi.cpu().numpy(), od.cpu().numpy()
        bad = np.nonzero((gi != ri).any(axis=1))[0]
        row[f"k{k}_rows_differ"] = int(bad.size)
        row[f"k{k}_d2_differ"] = int((gd != rd).any(axis=1).sum())
        if bad.size:
            j = int(bad[0])
            row[f"k{k}_first"] = {"row": j, "q": h_q[j].tolist(), "gpu": gi[j].tolist(), "ref": ri[j].tolist(), "gd": gd[j].tolist(), "rd": rd[j].tolist()}
            row[f"k{k}_bad_rows"] = bad[:20].tolist()
    print(json.dumps(row))


if __name__ == "__main__":
    main()
