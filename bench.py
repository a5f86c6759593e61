#!/usr/bin/env python
"""bench.py — points grouped per second through binarize + neighbour search + HP clustering + fragment
filter + LP assignment + centres (BASELINE.json metric) on the synthetic ScanNet-val-shaped set
(312 scenes, 50k-250k points each), sharded by scene over N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nnodes=1 --nproc-per-node N bench.py --gpus N ...

One "step" = one pass of the whole grouping path over every per-class call of every scene of the set
(one pb_binary_cluster_batched launch sequence per rank).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

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

METRIC = "points grouped/sec (binarize+search+cluster+vote)"
UNIT = "points/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ----------------------------------------------------------------------------------------------------
def cpu_port_rate(w, max_points, threads):
    """CPU oracle (faithful restatement, oracle/pb_oracle.c) on a bounded sample of the workload's calls,
    `threads` worker threads over calls (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pb_oracle as po
    from pbnet_b200 import workload
    po.build()
    r18 = np.full(18, np.float32(0.04), np.float32)
    m18 = np.full(18, 31, np.int32)
    jobs, pts = [], 0
    for c, ps, ss in workload.iter_calls(w):
        jobs.append((ps, ss))
        pts += ps.stop - ps.start
        if pts >= max_points:
            break

    def run(j):
        ps, ss = j
        xs = np.stack([w["x"][ps], w["y"][ps], w["z"][ps]], 1)
        xo = np.stack([w["xo"][ps], w["yo"][ps], w["zo"][ps]], 1)
        po.oracle_binary_cluster(xs, xo, w["sem"][ps], w["seg_counts"][ss], r18, m18)

    run(jobs[0])
    t0 = time.perf_counter()
    if threads > 1:
        jobs_sorted = sorted(jobs, key=lambda j: -(j[0].stop - j[0].start))
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, jobs_sorted))
    else:
        for j in jobs:
            run(j)
    dt = time.perf_counter() - t0
    return pts / dt, pts, len(jobs), dt


def ref_marshalled_calls(w, max_points):
    """The reference wrapper's argument marshalling (lib/PB_lib/torch_io/pbnet_ops.py:14-75) for the
    first calls of the workload, CPU tensors."""
    import torch

    from pbnet_b200 import workload
    calls, pts = [], 0
    for c, ps, ss in workload.iter_calls(w):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        calls.append(dict(xs=torch.stack([t(w["x"][ps]), t(w["y"][ps]), t(w["z"][ps])], 1),
                          xo=torch.stack([t(w["xo"][ps]), t(w["yo"][ps]), t(w["zo"][ps])], 1),
                          sem=t(w["sem"][ps]).long(), bp=t(w["seg_counts"][ss])))
        pts += ps.stop - ps.start
        if pts >= max_points:
            break
    return calls, pts


def run_reference_arm(args):
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    from pbnet_b200 import scenes, workload
    sizes = scenes.scene_sizes(args.scenes)
    sample_scenes = list(range(min(args.ref_scenes, args.scenes)))
    w = workload.build(sample_scenes, sizes, args.copies)
    line = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"C1 sample: first {len(sample_scenes)} of {args.scenes} synthetic ScanNet-val-shaped "
                                   f"scenes, per-class calls as network/PBNet.py:151-179, copies={args.copies}, "
                                   "r=0.04 min_pts=31"}}
    ref_so_dir = os.path.join(ROOT, "oracle", "_ref")
    kind, value, cores, sample = None, None, 1, ""
    try:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("no GPU")
        sys.path.insert(0, ref_so_dir)
        import PB_lib  # the UNMODIFIED compiled reference (oracle/build_ref.py)
        calls, pts = ref_marshalled_calls(w, int(w["n_points"]))

        def one_call(c):
            xs, xo = c["xs"], c["xo"]
            x, y, z = (xs[:, i].contiguous() for i in range(3))
            l1 = torch.abs(x) + torch.abs(y) + torch.abs(z)
            imap = torch.cat([torch.arange(0, int(b)) for b in c["bp"]]).type(torch.int32).contiguous()
            ox, oy, oz = (xo[:, i].contiguous() for i in range(3))
            n = xs.shape[0]
            cid = (torch.ones(n) * -1).type(torch.int32)
            cnum = torch.zeros([len(c["bp"])]).type(torch.int32)
            den = torch.zeros(n, dtype=torch.int32)
            cen = torch.zeros(n, dtype=torch.float32)
            cs = torch.zeros(n, dtype=torch.int32)
            PB_lib.binary_cluster(x, y, z, l1, imap, ox, oy, oz, c["sem"].type(torch.int32), c["bp"],
                                  (torch.ones(18) * 0.04).float(), (torch.ones(18) * 31).int(), cid, cnum, den, cen, cs,
                                  len(c["bp"]), 0.05, True)

        def step():
            for c in calls:
                one_call(c)
            torch.cuda.synchronize()
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
        kind, value, cores = "reference", pts / dt, 1
        sample = (f"UNMODIFIED reference PB_lib.binary_cluster (oracle/_ref, its only implementation: CUDA kernels "
                  f"driven by one host thread, CPU tensors in/out) on {len(calls)} per-class calls / {pts} points per step")
        ms = dt * 1e3
    except Exception as e:  # reference module unavailable -> CPU oracle port on all host cores
        threads = os.cpu_count() or 1
        t_all = []
        for _ in range(max(1, min(args.steps, 3))):
            rate, pts, ncalls, dt = cpu_port_rate(w, int(w["n_points"]), threads)
            t_all.append(dt)
        dt = float(np.mean(t_all))
        kind, value, cores = "port", pts / dt, threads
        sample = f"CPU oracle port ({type(e).__name__}: compiled reference not runnable) on {ncalls} calls / {pts} points"
        ms = dt * 1e3
    line.update({"value": value, "ms_per_step": ms, "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores,
                                                                      "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                 "gpu_launches": 0})
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=312)
    ap.add_argument("--copies", type=int, default=1)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: --scenes scenes PER GPU (a set of scenes x N_gpus scenes sharded by scene); strong: the same --scenes scenes sharded over the GPUs")
    ap.add_argument("--ref-scenes", type=int, default=8, help="scenes per step of the reference arm (bounded sample)")
    ap.add_argument("--cpu-sample-points", type=int, default=10_000_000, help="bounded CPU-baseline sample (~10-20 s on 16 cores)")
    ap.add_argument("--dropin-calls", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    rank, local_rank, world = dist_env()
    from pbnet_b200 import scenes, workload
    # scenes are independent units: they are partitioned over the ranks (LPT by point count), no data-path collective.
    # weak scaling (default): the set grows with the GPU count (args.scenes per GPU; the first args.scenes scenes are the
    # N=1 set); strong scaling: BASELINE.json configs[2] taken literally (the same args.scenes scenes over all GPUs)
    n_scenes_total = args.scenes * (world if args.scaling == "weak" else 1)
    sizes = scenes.scene_sizes(n_scenes_total)
    shards = workload.shard_scenes(sizes, world)
    # build the workload BEFORE touching CUDA (uses forked worker processes)
    w = workload.build(shards[rank], sizes, args.copies, workers=max(1, (os.cpu_count() or 1) // max(1, world)))

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (pbnet_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from pbnet_b200.cluster import Context
    ctx = Context(local_rank, profiling=True)
    n = int(w["n_points"])
    keys = ("x", "y", "z", "xo", "yo", "zo", "sem")
    d_in = [torch.from_numpy(w[k]).to(dev) for k in keys]
    seg, csc = w["seg_counts"], w["call_seg_counts"]
    S = len(seg)
    r18 = np.full(18, np.float32(scenes.RADIUS), np.float32)
    m18 = np.full(18, scenes.MIN_PTS, np.int32)
    d_out = dict(cluster_id=torch.empty(n, dtype=torch.int32, device=dev), cluster_num=torch.empty(S, dtype=torch.int32, device=dev),
                 degree=torch.empty(n, dtype=torch.int32, device=dev), center=torch.empty(3 * max(n // 32, 1024), dtype=torch.float32, device=dev),
                 clt_sem=torch.empty(max(n // 32, 1024), dtype=torch.int32, device=dev))
    stream = torch.cuda.current_stream()

    # gather of proposals to rank 0 (the only collective; NCCL over NVLink) — pbnet_b200/sharding.py
    from pbnet_b200 import sharding
    gather_ids = sharding.Rank0Gather(n, torch.int32, dev)  # size exchange + padded buffers once, not per step

    def step_device():
        out = ctx.binary_cluster(*d_in, seg, r18, m18, 0.05, True, call_seg_counts=csc, stream=stream, **d_out)
        if world > 1:
            gather_ids(d_out["cluster_id"])
        return out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    total_points = int(sum_over_ranks(float(n)))
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        out = step_device()
    launches_per_step = ctx.last_launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_acc = {}
    e0.record(stream)
    for _ in range(args.steps):
        out = step_device()
        for k, v in ctx.stage_ms().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    e1.record(stream)
    barrier()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    counters = ctx.counters()
    n_clusters = out["n_clusters"]
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}

    # ---- e2e: same call with HOST (pinned) buffers; H2D of inputs and D2H of results inside the timed region
    h_in = [torch.from_numpy(w[k]).pin_memory() for k in keys]
    h_out = dict(cluster_id=torch.empty(n, dtype=torch.int32).pin_memory(), cluster_num=torch.empty(S, dtype=torch.int32).pin_memory(),
                 degree=torch.empty(n, dtype=torch.int32).pin_memory(), center=torch.empty(3 * max(n // 32, 1024), dtype=torch.float32).pin_memory(),
                 clt_sem=torch.empty(max(n // 32, 1024), dtype=torch.int32).pin_memory())

    def step_host():
        return ctx.binary_cluster(*h_in, seg, r18, m18, 0.05, True, call_seg_counts=csc, stream=stream, **h_out)
    for _ in range(max(1, args.warmup - 1)):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oh = step_host()
    torch.cuda.synchronize()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / args.steps)
    clocks = sampler.stop() if rank == 0 else None  # sampled from the first warm-up step to the end of the e2e loop
    h2d = 28 * n
    d2h = 8 * n + 4 * S + 16 * int(oh["n_clusters"])
    e2e_stage = ctx.stage_ms()

    # ---- e2e through the reference-facing per-class operator (pbnet_ops.cluster, CPU tensors), bounded sample
    dropin = None
    if rank == 0:
        from pbnet_b200 import pbnet_ops
        calls, pts = [], 0
        for c, ps, ss in workload.iter_calls(w):
            if c >= args.dropin_calls:
                break
            t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
            calls.append((torch.stack([t(w["x"][ps]), t(w["y"][ps]), t(w["z"][ps])], 1),
                          torch.stack([t(w["xo"][ps]), t(w["yo"][ps]), t(w["zo"][ps])], 1), t(w["sem"][ps]).long(),
                          t(w["seg_counts"][ss])))
            pts += ps.stop - ps.start
        for c in calls[:8]:
            pbnet_ops.cluster(c[0], c[1], c[2], c[3], scenes.RADIUS, scenes.MIN_PTS, len(c[3]))
        dts = []
        for _ in range(3):  # host-side latency measurement: median of three passes (single passes are bimodal on shared hosts)
            t0 = time.perf_counter()
            for c in calls:
                pbnet_ops.cluster(c[0], c[1], c[2], c[3], scenes.RADIUS, scenes.MIN_PTS, len(c[3]))
            dts.append(time.perf_counter() - t0)
        dt = sorted(dts)[1]
        dropin = {"value": pts / dt, "unit": UNIT, "calls": len(calls), "points": pts,
                  "api": "pbnet_b200.pbnet_ops.cluster per (scene, class), CPU tensors in/out (reference call pattern)",
                  "us_per_call": 1e6 * dt / max(1, len(calls)), "passes_us_per_call": [round(1e6 * d / max(1, len(calls))) for d in dts]}

    # ---- voxelize / devoxelize (rows a12-a14): the HBM-bound scatter-gather of the path, rank 0 only
    vox = None
    if rank == 0:
        from pbnet_b200 import voxel
        nv = min(n, 8_000_000)
        coords = torch.stack([d_in[3][:nv], d_in[4][:nv], d_in[5][:nv]], 1).contiguous()  # original xyz
        bcol = torch.zeros(nv, dtype=torch.int32, device=dev)
        peak_v, _ = measured_peaks()

        def timed(f, reps):
            f()
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                r = f()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / reps * 1e-3, r
        t_vox, vm = timed(lambda: voxel.voxel_map(coords, scenes.VOXEL_SIZE, batch=bcol), 3)
        C = 76  # 32 + 20 + 20 + 3 + 1 channels gathered at network/PBNet.py:130-134
        vfeat = torch.randn((vm.n_voxels, C), device=dev)
        t_dev, o = timed(lambda: voxel.devoxelize_raw(vfeat, vm.inverse), 5)
        t_bwd, _ = timed(lambda: voxel.voxel_rows(o, vm, "sum"), 5)
        gb_f = (nv * C * 4 + vm.n_voxels * C * 4 + nv * 8) / 1e9
        gb_b = (nv * C * 4 + vm.n_voxels * C * 4 + nv * 4 + vm.n_voxels * 4) / 1e9
        vox = {"points": nv, "voxels": vm.n_voxels, "voxelize_points_per_s": nv / t_vox,
               "devoxelize": {"channels": C, "ms": t_dev * 1e3, "algorithmic_gb": gb_f, "achieved_gbs": gb_f / t_dev,
                              "frac_of_hbm_copy_peak": gb_f / t_dev / peak_v,
                              "note": "write-dominated (n*C*4 B written): the copy peak counts read+write, a pure write "
                                      "stream tops out near half of it"},
               "devoxelize_backward": {"ms": t_bwd * 1e3, "algorithmic_gb": gb_b, "achieved_gbs": gb_b / t_bwd,
                                       "frac_of_hbm_copy_peak": gb_b / t_bwd / peak_v}}
        del vfeat, o, coords

    # ---- the callers either side of the path on ONE scene with three rotated copies, the unit
    #      eval_map.py processes per iteration: grouping + local scenes + feature rows + proposal voxelization
    #      (grouping.propose), get_proposal, evaluation post-processing; rank 0 only
    nxt = None
    if rank == 0:
        from pbnet_b200 import evalpost, grouping
        sc0 = scenes.make_scene(scenes.BASE_SEED, int(sizes[0]))
        cp3 = 3
        t_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        xyz3 = t_(np.concatenate(scenes.rotate_copies(sc0["xyz_orig"], cp3)))
        off3 = t_(np.concatenate(scenes.rotate_copies(sc0["offset"], cp3)))
        sem3 = t_(np.tile(sc0["sem"], cp3))
        bh3 = t_(np.repeat(np.arange(cp3), sc0["sem"].shape[0]).astype(np.int32))
        gen = torch.Generator(device=dev).manual_seed(22)
        feat3 = torch.rand((xyz3.shape[0], 32), device=dev, generator=gen)
        sfp3 = torch.softmax(torch.rand((xyz3.shape[0], 20), device=dev, generator=gen), dim=1)
        t_prop, pr = timed(lambda: grouping.propose(xyz3, off3, sem3, bh3, feat3, sfp3, scenes.RADIUS, scenes.MIN_PTS, cp3), 5)
        scn = pr["scenes"]
        E3, P3 = int(scn["index"].shape[0]), int(scn["offsets"].shape[0]) - 1
        ms3 = torch.rand(E3, device=dev, generator=gen)
        t_gp, gp = timed(lambda: grouping.get_proposal(scn["offsets"], scn["index"], ms3), 5)
        score3 = torch.rand(int(gp[1].shape[0]) - 1, device=dev, generator=gen)
        n3 = xyz3.shape[0] // cp3
        # superpoints: 10 cm voxels of the original coordinates (spatially coherent, compressed ids)
        sp3 = torch.unique(torch.floor(xyz3[:n3] / 0.1).to(torch.int64), dim=0, return_inverse=True)[1].contiguous()
        t_ev, ev = timed(lambda: evalpost.postprocess(gp[0], gp[1], score3, sem3, sp3, int(xyz3.shape[0])), 5)
        nxt = {"unit": "one scene x 3 rotated copies (the per-iteration unit of eval_map.py)", "points": int(xyz3.shape[0]),
               "proposals": P3, "list_entries": E3, "voxels": int(pr["voxel_coords"].shape[0]),
               "propose_ms": t_prop * 1e3, "propose_api": "grouping.propose = network/PBNet.py:144-247 (class loop, local scenes, "
                                                          "feature rows, proposal voxelization)",
               "get_proposal_ms": t_gp * 1e3, "kept_entries": int(gp[0].shape[0]),
               "eval_postprocess_ms": t_ev * 1e3, "final_clusters": int(ev["scores"].shape[0])}
        del xyz3, off3, sem3, bh3, feat3, sfp3, pr, scn

    # ---- CPU baseline (oracle port) on a bounded sample, rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        rate, pts, ncalls, dt = cpu_port_rate(w, args.cpu_sample_points, threads)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {ncalls} per-class calls ({pts} points) of the same workload, oracle/pb_oracle.c "
                         f"(grid-accelerated faithful restatement), {threads} threads over calls, {dt:.1f} s"}

    if rank == 0:
        peak, peak_src = measured_peaks()
        cells = max(1, counters["cells"])
        chunks = max(1, counters.get("chunks", 1))
        # algorithmic bytes of k_degree: per point 16 B pts4 + 4 B row_of read, 4 B degree write;
        # per fine cell 12 B (key, coarse ordinal); per coarse cell 76 B (9 stencil rows + point offset)
        deg_bytes = (24.0 * n + 12.0 * cells + 76.0 * counters.get("coarse_cells", 0)) / chunks   # per launch
        traffic = None
        try:  # dram__bytes_read+write per launch from the committed ncu --set full capture of the same workload
            prof = json.load(open(os.path.join(ROOT, "profiles", "ncu_r01_k_degree.json")))
            if world == 1 and abs(prof["points_per_launch"] - n / chunks) < 0.02 * n:
                traffic = prof["dram_bytes_per_launch"]
        except Exception:
            pass
        deg_ms_total = stage_ms.get("degree", 0.0)               # summed over the chunks of a step
        deg_ms = deg_ms_total / chunks                           # average launch duration
        achieved = deg_bytes / (deg_ms * 1e-3) / 1e9 if deg_ms > 0 else None
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        # ceiling of the packed fp32x2 pair test measured with tools/microbench/pipes.cu on this pool's B200
        # (profiles/microbench_pipes_r01_v2.txt, slowest-warp timing, no loop-invariant operands): 0.485 warp-tests/clk/SM
        # — the fp32 pipe retires 6 operations per test (3 FADD, FMUL, 2 FFMA); the scalar form tops out at 0.332
        alu_peak = 148 * 32 * 0.485 * sm_mhz * 1e6
        tests_per_s = counters["pair_tests"] / (deg_ms_total * 1e-3) if deg_ms_total > 0 else None
        value = total_points / (ms_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"C1: {n_scenes_total} synthetic ScanNet-val-shaped scenes "
                                   f"({args.scenes} per GPU, {args.scaling} scaling; 50k-250k points, seed 22+s), "
                                   f"per-class calls as network/PBNet.py:151-179, copies={args.copies}, r=0.04, min_pts=31, "
                                   "sharded by scene (LPT) over ranks",
                       "points_total": total_points, "points_rank0": n, "calls_rank0": int(len(csc)), "segments_rank0": S,
                       "clusters_rank0": int(n_clusters),
                       "l2": "inputs + workspace of one step are ~%.1f GB per rank, far beyond the 126 MB L2; no flush needed" % (
                           (28 + 430) * n / 1e9),
                       "collective": "NCCL gather of cluster ids to rank 0 once per step (inside the timed region)" if world > 1 else "none"},
            "e2e": {"value": total_points / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "pb_binary_cluster_batched via pbnet_b200.cluster.Context.binary_cluster, pinned host buffers",
                    "ms_per_step": e2e_s * 1e3, "stage_ms": {k: round(v, 3) for k, v in e2e_stage.items()}},
            "e2e_dropin": dropin,
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()},
            "stage_ms_note": "CUDA-event intervals summed over the chunks of a step; the chunks run on two concurrent streams, so an "
                             "interval also contains the time its kernels share the GPU with the other chunk (k_degree runs alone on "
                             "the SMs, its interval is the kernel time; profiles/launches_*_summary.txt has the serialised per-kernel times)",
            "roofline": {"bound": "hbm", "kernel": "k_degree", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "launches_per_step": chunks, "avg_launch_ms": deg_ms, "share_of_step": deg_ms_total / ms_step,
                         "algorithmic_bytes_per_launch": deg_bytes,
                         "note": "k_degree is fp32-issue bound, not HBM bound: see roofline_alu"},
            "roofline_alu": {"kernel": "k_degree", "pair_tests_per_step": counters["pair_tests"],
                             "achieved": tests_per_s, "peak": alu_peak, "unit": "pair tests/s",
                             "frac": (tests_per_s / alu_peak) if tests_per_s else None,
                             "peak_def": "148 SM x 32 lanes x 0.485 warp-tests/clk/SM (measured ceiling of the packed fp32x2 pair test, "
                                         "tools/microbench/pipes.cu, profiles/microbench_pipes_r01_v2.txt) x measured SM clock"},
            "io_roofline": {"bytes_per_point": 36, "achieved_gbs": value * 36 / 1e9, "frac_of_hbm": value * 36 / 1e9 / peak},
            "counters": counters,
            "voxel": vox,
            "next_rows": nxt,
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
