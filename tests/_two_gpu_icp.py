"""Helper for tests/test_multirank_gloo.py::test_two_gpu_icp_matches_single_gpu (launched under torchrun)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pcl_b200 as P  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = P.Context(local)
uid = [P.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(rank, world, uid[0])
rng = np.random.default_rng(5)
n = 200000
tgt = P.xyz1(rng.random((n, 3), dtype=np.float32))
a = np.deg2rad(3.0)
R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
src = P.xyz1((tgt[:, :3].astype(np.float64) @ R.T + [0.01, -0.01, 0.02]).astype(np.float32))
kw = dict(max_iterations=20, max_correspondence_distance=0.1)
idx = P.Index(ctx, tgt)
shard = src[rank * n // world:(rank + 1) * n // world]
r = P.icp_align(ctx, shard, idx, **kw)
T = torch.from_numpy(r["final"]).cuda()
Ts = [torch.empty_like(T) for _ in range(world)]
dist.all_gather(Ts, T)
assert all(torch.equal(Ts[0], t) for t in Ts), "ranks disagree"
if rank == 0:
    ctx1 = P.Context(local)  # no communicator: whole cloud on one GPU
    r1 = P.icp_align(ctx1, src, P.Index(ctx1, tgt), **kw)
    err = float(np.linalg.norm(r1["final"] - r["final"]))
    assert r["n_correspondences"] == r1["n_correspondences"], (r["n_correspondences"], r1["n_correspondences"])
    assert r["iterations"] == r1["iterations"] and err < 1e-6, (err, r["iterations"], r1["iterations"])
    print("TWO_GPU_OK", err)
dist.barrier()
dist.destroy_process_group()
