#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on BASELINE.json's 10 M-point configuration.

Workload (config.workload = "cfg3_10M_point_to_plane", BASELINE.json configs[2]):
  target : 10 000 000 points, (x,y) ~ U[0,10)^2, z = 0.5 sin(x) cos(0.7 y) + N(0, 0.002^2), seed 7
  source : the same surface re-sampled (seed 8 + rank), rotated 2 deg about z, t = (0.02, 0.01, -0.01)
  target normals: NormalEstimation k = 16, viewpoint (5,5,10)  (set-up, timed separately, not part of a step)
  ICP    : IterativeClosestPointWithNormals (non-symmetric => TransformationEstimationPointToPlaneLLS),
           max correspondence distance 0.05, k = 1
A STEP is one Registration::align() of ICP_ITERS = 10 iterations (PCL's default max_iterations_,
registration.h:566) over the whole source cloud: source upload -> Morton ordering of the queries -> 10 x
(1-NN search + gate + 6x6 accumulation + solve + transform) -> output cloud.  The target index is built once
before the timed region, exactly as pcl::Registration keeps its tree across align() calls
(registration.hpp:84-87).
  value : correspondences/s with source, target index and output resident in HBM (device pointers through
          the same C-ABI calls pclb200_icp_set_source / _iterate / _get_cloud the C++ facade's align() makes)
  e2e   : the same call with HOST buffers (pinned pcl::PointNormal records, 48 B/pt): H2D of the source and
          D2H of the aligned cloud inside the timed region
  N > 1 : weak scaling — every rank holds a replica of the target index and its own 10 M-point source shard;
          the 40 fp64 accumulators are all-reduced (NCCL) once per iteration.
--impl reference times the CPU restatement of PCL's own path (oracle/, the reference cannot be compiled in this
image: no Eigen/Boost/FLANN) on the box's host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ICP_ITERS = 10
MAX_CORR_DIST = 0.05
N_DEFAULT = 10_000_000
_JSON_OUT = sys.stdout
METRIC = "icp_correspondences_per_sec"
UNIT = "correspondences/s"


def surface(n, seed):
    r = np.random.default_rng(seed)
    xy = r.random((n, 2)) * 10.0
    z = 0.5 * np.sin(xy[:, 0]) * np.cos(0.7 * xy[:, 1]) + r.normal(0.0, 0.002, n)
    return xy[:, 0], xy[:, 1], z


def make_target(n):
    x, y, z = surface(n, 7)
    t = np.zeros((n, 12), dtype=np.float32)  # pcl::PointNormal records
    t[:, 0], t[:, 1], t[:, 2], t[:, 3] = x, y, z, 1.0
    return t


def make_source(n, rank):
    x, y, z = surface(n, 8 + rank)
    a = np.deg2rad(2.0)
    c, s = np.cos(a), np.sin(a)
    out = np.zeros((n, 12), dtype=np.float32)
    out[:, 0] = c * x - s * y + 0.02
    out[:, 1] = s * x + c * y + 0.01
    out[:, 2] = z - 0.01
    out[:, 3] = 1.0
    return out


def analytic_normals(t, vp=(5.0, 5.0, 10.0)):
    """Normals of z = 0.5 sin x cos 0.7y (used by the REFERENCE arm only: it may not call our normals kernel)."""
    x, y = t[:, 0].astype(np.float64), t[:, 1].astype(np.float64)
    n = np.stack([-0.5 * np.cos(x) * np.cos(0.7 * y), 0.35 * np.sin(x) * np.sin(0.7 * y), np.ones_like(x)], 1)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    flip = ((np.asarray(vp) - t[:, :3]) * n).sum(1) < 0
    n[flip] *= -1
    return n.astype(np.float32)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML (in-process thread; the same
    counters `nvidia-smi --query-gpu=clocks.sm,clocks_event_reasons.*` prints, B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index, period_s=0.05):
        self.gpu, self.period = gpu_index, period_s
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self.stop_flag = threading.Event()
        self.th = None
        self.err = None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.gpu]) if vis and vis.split(",")[0].isdigit() else self.gpu
            self.h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.nv = nv
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._run, daemon=True)
            self.th.start()
        except Exception as e:  # noqa
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "hw_power_brake_slowdown": 0x80}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
            except Exception as e:  # noqa
                self.err = repr(e)
                break
            self.stop_flag.wait(self.period)

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=2)
        out = {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
               "samples": len(self.sm), "reasons": sorted(self.reasons), "how": "NVML, 50 ms period, timed region only"}
        if self.err:
            out["error"] = self.err
        return out


def measured_peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# -----------------------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement of PCL's path (oracle/), all host threads, bounded sample
# -----------------------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    import oracle
    n = args.points
    cores = oracle.max_threads()
    tgt = make_target(n)
    tgt[:, 4:7] = analytic_normals(tgt)
    src_full = make_source(n, 0)
    t0 = time.time()
    tree = oracle.Index(tgt)  # pcl::KdTreeFLANN::setInputCloud, kept across align() calls
    build_s = time.time() - t0
    kw = dict(max_iterations=ICP_ITERS, max_correspondence_distance=MAX_CORR_DIST, estimator=1,
              with_normals_transform=True, source_has_normals=True, nthreads=cores, index=tree)
    # calibrate the sample so that (steps + warmup) aligns end within ~150 s
    probe = min(n, 100_000)
    t0 = time.time()
    r = oracle.icp_align(src_full[:probe], tgt, want_cloud=True, **kw)
    rate = max(r["total_correspondences"], 1) / max(time.time() - t0, 1e-6)
    budget_s = 150.0 / max(args.steps + args.warmup, 1)
    sample = int(min(n, max(probe, rate * budget_s / ICP_ITERS)))
    src = np.ascontiguousarray(src_full[:sample])
    out = np.empty_like(src)
    for _ in range(args.warmup):
        oracle.icp_align(src, tgt, out=out, **kw)
    total, t0 = 0, time.time()
    for _ in range(args.steps):
        r = oracle.icp_align(src, tgt, out=out, **kw)
        total += r["total_correspondences"]
    dt = time.time() - t0
    val = total / dt
    sample_desc = (f"{sample} of {n} source points (first rows) x {ICP_ITERS} iterations per step against the full "
                   f"{n}-point target; kd-tree build ({build_s:.1f} s, 1 thread) outside the timed region; "
                   "analytic target normals")
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3_10M_point_to_plane", "points_target": n, "points_source_per_step": sample,
                       "icp_iterations_per_step": ICP_ITERS, "max_correspondence_distance": MAX_CORR_DIST,
                       "estimator": "point_to_plane_lls", "k": 1},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample_desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


# -----------------------------------------------------------------------------------------------------------------
# our arm
# -----------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import pcl_b200 as P
    rank, world, local = dist_env()
    n = args.points
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    P.lib()
    ctx = P.Context(local)
    if world > 1:
        uid = [P.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    # ---- set-up (outside the timed region): target index + k=16 normals, both on the GPU ----------------------
    tgt = make_target(n)
    src_host = torch.from_numpy(make_source(n, rank)).pin_memory()           # pcl::PointNormal records, pinned
    out_host = torch.empty_like(src_host).pin_memory()
    ctx.profile(True)
    tidx = P.Index(ctx, tgt)
    tgt_dev = torch.from_numpy(tgt).to(dev)
    nrm_dev = torch.empty((n, 4), dtype=torch.float32, device=dev)
    tidx.normals_knn(tgt_dev, 16, viewpoint=(5.0, 5.0, 10.0), out=nrm_dev)
    build_ms, _ = ctx.profile_get("index_build")
    normals_ms, _ = ctx.profile_get("normals")
    del tgt_dev
    src_dev = src_host.to(dev)                    # value leg: 48-byte records already resident in HBM
    out_dev = torch.empty_like(src_dev)
    params = P.default_params(max_iterations=ICP_ITERS, max_correspondence_distance=MAX_CORR_DIST,
                              estimator=P.EST_POINT_TO_PLANE_LLS, with_normals_transform=1, mse_threshold_absolute=0.0)

    # One registration object for the whole run, like a pcl::IterativeClosestPointWithNormals instance whose target
    # was set once (Registration::setInputTarget): the target index and its normals stay on the device across
    # align() calls (registration.hpp:84-87); every step is setInputSource + align(output).
    icp = P.Icp(ctx, params=params)
    icp.set_target(tidx, normals=nrm_dev)

    def align(src, out):
        icp.set_source(src, normals=P.Field(src, 4))
        st = icp.iterate()
        icp.get_cloud(out, normals=P.Field(out, 4))
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(src, out, steps, warmup, sample_clocks):
        for _ in range(warmup):
            align(src, out)
        ctx.profile_reset()
        l0 = ctx.launches
        barrier()
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0
        e0.record(stream)
        for _ in range(steps):
            total += align(src, out)["total_correspondences"]
        e1.record(stream)
        barrier()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms, float(total)], dtype=torch.float64, device=dev)
        if world > 1:
            tm = t.clone()
            dist.all_reduce(tm, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            ms, total = float(tm[0]), float(t[1])
        # NOTE: with the all-reduce inside every iteration each rank already reports the GLOBAL pair count;
        # `total` summed over ranks would double count, so take rank 0's figure in that case.
        return ms, total, ctx.launches - l0, clocks

    # device-resident leg (value) --------------------------------------------------------------------------------
    ms_v, tot_v, launches, clocks = timed(src_dev, out_dev, args.steps, args.warmup, not args.no_clocks)
    iter_ms, iter_n = ctx.profile_get("icp_search")
    accum_ms, _ = ctx.profile_get("icp_accum")
    allreduce_ms, _ = ctx.profile_get("allreduce")
    sort_ms, _ = ctx.profile_get("query_sort")
    solve_ms, _ = ctx.profile_get("solve")
    out_ms, _ = ctx.profile_get("transform_out")
    # host-buffer leg (e2e) --------------------------------------------------------------------------------------
    ms_e, tot_e, _, _ = timed(src_host, out_host, args.steps, max(1, args.warmup // 2), False)
    if world > 1:  # stats already carry the all-reduced (global) pair counts on every rank
        tot_v /= world
        tot_e /= world
    value = tot_v / (ms_v * 1e-3)
    e2e = tot_e / (ms_e * 1e-3)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # roofline of the dominant kernel (k_search; the first search of a step is k_search_packet), algorithmic bytes per
    # correspondence:
    #   16 source read + 16 source write-back (T_k applied in place) + 16 previous match read + 16 match write
    #   + (target leaf slots + nodes, each read about once under Hilbert-ordered queries) / N_s
    st_idx = tidx.stats
    # what a search reads of the index: the padded leaf lines (16 B per slot) and the 64-byte nodes — not the auxiliary
    # arrays the index also owns (parent pointers, lazily built position maps), which the default walk never touches
    tree_bytes = st_idx["leaves"] * st_idx["leaf_size"] * 16 + st_idx["nodes"] * 64
    bytes_per_corr = 16 + 16 + 16 + 16 + tree_bytes / float(n)
    peak, peak_src = measured_peak_gbs()
    avg_iter_s = (iter_ms / max(iter_n, 1)) * 1e-3
    achieved = bytes_per_corr * n / avg_iter_s / 1e9 if avg_iter_s > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "k_search (in-place transform + exact seeded 1-NN + gate; first search of a step: k_search_packet)", "achieved": achieved, "peak": peak,
                "peak_source": peak_src, "unit": "GB/s", "frac": achieved / peak,
                "frac_of_nominal_8000_GBs": achieved / 8000.0,  # report both denominators
                # dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel on this
                # workload (profiles/r1k_k_search_ncu.txt: k_search<0,0>, 698.3 MB + 313.6 MB); only valid for the 10 M default
                "traffic": 1011.9e6 if n == N_DEFAULT else None,
                "algorithmic_bytes_per_launch": bytes_per_corr * n, "avg_launch_ms": avg_iter_s * 1e3,
                "launches_timed": iter_n, "algorithmic_bytes_per_correspondence": bytes_per_corr,
                "index": st_idx,
                "note": "BVH traversal is latency/issue bound, not HBM bound (SURVEY.md §7 hard part ii, DESIGN.md)"}

    # CPU baseline (oracle port) on a bounded sample: one align() of ICP_ITERS iterations, sample sized for ~15 s
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import oracle
        cores = oracle.max_threads()
        tgt_n = tgt.copy()
        tgt_n[:, 4:8] = nrm_dev.cpu().numpy()
        t0 = time.time()
        tree = oracle.Index(tgt_n)
        build_s = time.time() - t0
        kw = dict(max_iterations=ICP_ITERS, max_correspondence_distance=MAX_CORR_DIST, estimator=1,
                  with_normals_transform=True, source_has_normals=True, nthreads=cores, index=tree)
        src_np = src_host.numpy()
        probe = min(n, 100_000)
        t0 = time.time()
        r = oracle.icp_align(src_np[:probe], tgt_n, **kw)
        rate = max(r["total_correspondences"], 1) / max(time.time() - t0, 1e-6)
        sample = int(min(n, max(probe, rate * 15.0 / ICP_ITERS)))
        t0 = time.time()
        r = oracle.icp_align(np.ascontiguousarray(src_np[:sample]), tgt_n, want_cloud=True, **kw)
        dt = time.time() - t0
        cpu = {"value": r["total_correspondences"] / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{sample} of {n} source points x {ICP_ITERS} iterations (one align) against the full "
                         f"{n}-point target, GPU-computed normals; kd-tree build {build_s:.1f} s excluded"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_v / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg3_10M_point_to_plane", "points_target": n, "points_source_per_gpu": n,
                       "icp_iterations_per_step": ICP_ITERS, "max_correspondence_distance": MAX_CORR_DIST,
                       "estimator": "point_to_plane_lls", "k": 1, "normals_k": 16,
                       "l2_policy": "inputs larger than L2 (target 160 MB + nodes 80 MB + normals 160 MB + source 160 MB)",
                       "parallelism": f"source sharded x{world}, target replicated, 40-double all-reduce/iteration"},
            "ms_per_iter": ms_v / args.steps / ICP_ITERS,
            "breakdown_ms_per_step": {"icp_search_kernel": iter_ms / args.steps, "icp_accum_kernel": accum_ms / args.steps, "nccl_allreduce": allreduce_ms / args.steps, "query_sort": sort_ms / args.steps,
                                      "solve": solve_ms / args.steps, "transform_out": out_ms / args.steps},
            "setup_ms": {"index_build": build_ms, "normals_k16": normals_ms},
            "e2e": {"value": e2e, "unit": UNIT, "ms_per_step": ms_e / args.steps,
                    "h2d_bytes_per_step": int(src_host.numel() * 4), "d2h_bytes_per_step": int(out_host.numel() * 4 + 512 * ICP_ITERS)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if cpu:
        line["cpu_baseline"] = cpu
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()
    if world > 1:
        dist.destroy_process_group()


def _protect_stdout():
    """Keep stdout for the ONE JSON line: libraries (NCCL prints its version banner) get stderr instead."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


def main():
    global _JSON_OUT
    _JSON_OUT = _protect_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--points", type=int, default=N_DEFAULT)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="debug: do not sample clocks during the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
