#!/usr/bin/env python
"""bench.py — points grouped per second through binarize + neighbour search + HP clustering + fragment
filter + LP assignment + centres (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c3|c4]
    torchrun --nnodes=1 --nproc-per-node N bench.py --gpus N ...

Default workload C1 (BASELINE.json configs[1] / [2]): the synthetic ScanNet-val-shaped set (312 scenes, 50k-250k
points each), sharded by scene over the N ranks — the SAME 312 scenes at every N ("strong" scaling,
configs[2] taken literally; ``--scaling weak`` grows the set with N instead).  One "step" = one pass of the whole
grouping path over every per-class call of every scene of the rank's shard (one pb_binary_cluster_batched launch
sequence per rank) + the NCCL gather of the cluster ids to rank 0.  Prints ONE JSON line on rank 0.

After the timed region every output of the device-resident step is compared bit for bit with the CPU oracle
(``verify``), and — where the compiled reference travelled to the box — the first scenes with the live reference.
``--workload c3`` / ``c4`` run BASELINE.json configs[3] / [4] (large-scene radius sweep, dense HP-fraction sweep)
as one JSON line with a ``sweep`` list.
"""
from __future__ import annotations

import argparse
import glob
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
# ceilings of the packed fp32x2 pair test measured with tools/microbench/pipes.cu on this pool's B200 (slowest-warp timing, no
# loop-invariant operands; profiles/microbench_pipes_r01_v2.txt, profiles/microbench_pipes_r02.txt), in warp-tests/clk/SM:
#   0.485  the one-sided candidate loop (3 FADD2 + FMUL2 + 2 FFMA2 + 2 FSETP + 2 IADD per two tests; the scalar form: 0.332)
#   0.387  the SYMMETRIC candidate loop of round 2 at two query pairs per lane (+ IADD3 slot sums, one REDUX + RED per four
#          candidates; 0.414 at three pairs per lane) - every executed test settles BOTH directions of a pair
PAIR_TEST_CEILING = 0.485
PAIR_TEST_CEILING_SYM = 0.387


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
# checkers (test infrastructure: the only places this file touches oracle/)
# ----------------------------------------------------------------------------------------------------
def _call_tables(w):
    seg_off = np.concatenate([[0], np.cumsum(w["call_seg_counts"])]).astype(np.int64)
    pt_off = np.concatenate([[0], np.cumsum(w["seg_counts"].astype(np.int64))])
    return seg_off, pt_off


def _same_call(got, want):
    bad = [k for k in ("cluster_id", "cluster_num", "den_queue", "clt_sem") if not np.array_equal(got[k], want[k])]
    if got["center"].shape != want["center"].shape or not np.array_equal(got["center"].view(np.uint32),
                                                                           want["center"].view(np.uint32)):
        bad.append("center")
    return bad


def _slice_call(w, host, c, seg_off, pt_off, k_off):
    s0, s1 = int(seg_off[c]), int(seg_off[c + 1])
    p0, p1 = int(pt_off[s0]), int(pt_off[s1])
    k0, k1 = int(k_off[c]), int(k_off[c + 1])
    return dict(cluster_id=host["cluster_id"][p0:p1], den_queue=host["degree"][p0:p1], cluster_num=host["cluster_num"][s0:s1],
                center=host["center"][3 * k0:3 * k1], clt_sem=host["clt_sem"][k0:k1]), slice(p0, p1), slice(s0, s1)


def verify_against_oracle(w, host, r18, m18, threads, time_budget_s=None):
    """Every per-class call of the workload through oracle/pb_oracle.c (faithful CPU restatement, pinned to the compiled
    reference) on `threads` host threads; every output compared bit for bit (centres included).  Returns the verify
    block and the oracle's throughput (the `cpu_baseline` port figure)."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import pb_oracle as po
    po.build()
    seg_off, pt_off = _call_tables(w)
    n_calls = len(w["call_seg_counts"])
    k_off = np.concatenate([[0], np.cumsum(host["call_clusters"])]).astype(np.int64)
    order = sorted(range(n_calls), key=lambda c: -(int(pt_off[seg_off[c + 1]]) - int(pt_off[seg_off[c]])))
    t0 = time.perf_counter()
    state = {"pts": 0, "calls": 0, "skipped": 0}
    bad = []
    lock = threading.Lock()

    def run(c):
        if time_budget_s is not None and time.perf_counter() - t0 > time_budget_s:
            with lock:
                state["skipped"] += 1
            return
        got, ps, ss = _slice_call(w, host, c, seg_off, pt_off, k_off)
        xs = np.stack([w["x"][ps], w["y"][ps], w["z"][ps]], 1)
        xo = np.stack([w["xo"][ps], w["yo"][ps], w["zo"][ps]], 1)
        want = po.oracle_binary_cluster(xs, xo, w["sem"][ps], w["seg_counts"][ss], r18, m18)
        b = _same_call(got, want)
        with lock:
            state["pts"] += ps.stop - ps.start
            state["calls"] += 1
            if b:
                bad.append((c, b))

    if threads > 1:
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(run, order))
    else:
        for c in order:
            run(c)
    dt = time.perf_counter() - t0
    return {"against": "oracle/pb_oracle.c (CPU restatement pinned 147/147 to the compiled reference)", "calls": state["calls"],
            "calls_total": n_calls, "points": state["pts"], "mismatches": len(bad), "first_mismatches": [[int(c), b] for c, b in bad[:5]],
            "skipped_for_time": state["skipped"], "seconds": round(dt, 2), "threads": threads,
            "fields": "cluster_id, cluster_num, den_queue, clt_sem, center (bit-exact)"}, (state["pts"] / dt if dt > 0 else None)


def load_reference_module():
    """The UNMODIFIED compiled reference (oracle/_ref/PB_lib*.so, oracle/build_ref.py) or None."""
    so = glob.glob(os.path.join(ROOT, "oracle", "_ref", "PB_lib*.so"))
    if not so:
        return None
    try:
        import importlib.util

        import torch  # noqa: F401
        spec = importlib.util.spec_from_file_location("PB_lib", so[0])
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    except Exception:
        return None


def reference_call(PB_lib, xs, xo, sem, bp, radius=0.04, min_pts=31):
    """One call of the reference extension marshalled as lib/PB_lib/torch_io/pbnet_ops.py:14-75 does (CPU tensors)."""
    import torch
    x, y, z = (xs[:, i].contiguous() for i in range(3))
    l1 = torch.abs(x) + torch.abs(y) + torch.abs(z)
    imap = torch.cat([torch.arange(0, int(b)) for b in bp]).type(torch.int32).contiguous()
    ox, oy, oz = (xo[:, i].contiguous() for i in range(3))
    n = xs.shape[0]
    cid = (torch.ones(n) * -1).type(torch.int32)
    cnum = torch.zeros([len(bp)]).type(torch.int32)
    den = torch.zeros(n, dtype=torch.int32)
    cen = torch.zeros(n, dtype=torch.float32)
    cs = torch.zeros(n, dtype=torch.int32)
    PB_lib.binary_cluster(x, y, z, l1, imap, ox, oy, oz, sem.type(torch.int32), bp, (torch.ones(18) * radius).float(),
                          (torch.ones(18) * min_pts).int(), cid, cnum, den, cen, cs, len(bp), 0.05, True)
    return dict(cluster_id=cid.numpy(), cluster_num=cnum.numpy(), den_queue=den.numpy(), center=cen.numpy(), clt_sem=cs.numpy())


def verify_against_reference(w, host, max_calls):
    """The first `max_calls` calls through the live compiled reference on this GPU, compared bit for bit."""
    import torch
    PB_lib = load_reference_module()
    if PB_lib is None:
        return {"against": "oracle/_ref (compiled reference)", "unavailable": "oracle/_ref/PB_lib*.so not present"}
    seg_off, pt_off = _call_tables(w)
    k_off = np.concatenate([[0], np.cumsum(host["call_clusters"])]).astype(np.int64)
    n_calls = min(max_calls, len(w["call_seg_counts"]))
    bad, pts = [], 0
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    for c in range(n_calls):
        got, ps, ss = _slice_call(w, host, c, seg_off, pt_off, k_off)
        want = reference_call(PB_lib, torch.stack([t(w["x"][ps]), t(w["y"][ps]), t(w["z"][ps])], 1),
                              torch.stack([t(w["xo"][ps]), t(w["yo"][ps]), t(w["zo"][ps])], 1), t(w["sem"][ps]).long(),
                              t(w["seg_counts"][ss]))
        b = _same_call(got, want)
        pts += ps.stop - ps.start
        if b:
            bad.append([c, b])
    return {"against": "oracle/_ref: the UNMODIFIED compiled reference PB_lib.binary_cluster, live on this GPU", "calls": n_calls,
            "points": pts, "mismatches": len(bad), "first_mismatches": bad[:5]}


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


def c1_config(args, world, n_scenes_total):
    return (f"C1: {n_scenes_total} synthetic ScanNet-val-shaped scenes (50k-250k points, seed 22+s), per-class calls as "
            f"network/PBNet.py:151-179, copies={args.copies}, r=0.04, min_pts=31, sharded by scene (LPT) over {world} rank(s), "
            f"{args.scaling} scaling")


def c1_config_dict(args, world, n_scenes_total):
    """`config` of BOTH arms (identical by construction; per-run details live in `workload_detail` / `cpu_baseline.sample`)."""
    return {"workload": c1_config(args, world, n_scenes_total),
            "l2": "no flush needed: the inputs and the workspace of one step are gigabytes per rank, far beyond the 126 MB L2"}


def run_reference_arm(args):
    """The reference's own implementation of the path on the box: the UNMODIFIED compiled PB_lib (its only implementation
    is host-driven CUDA with CPU tensors in/out) on a bounded sample of the arm's workload; rank 0 only."""
    rank, local_rank, world = dist_env()
    if rank != 0:
        return
    from pbnet_b200 import scenes, workload
    n_scenes_total = args.scenes * (world if args.scaling == "weak" else 1)
    sizes = scenes.scene_sizes(n_scenes_total)
    sample_scenes = list(range(min(args.ref_scenes, args.scenes)))
    w = workload.build(sample_scenes, sizes, args.copies)
    line = {"metric": METRIC, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": c1_config_dict(args, world, n_scenes_total),
            "gpus_used": 1,
            "note": "the reference has no multi-GPU inference path (eval is single-GPU, config/config_test.py:32): at --gpus N "
                    "this arm still runs on ONE GPU driven by one host thread, so an N>1 ratio against it is N GPUs vs 1"}
    kind, value, cores, sample = None, None, 1, ""
    try:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("no GPU")
        PB_lib = load_reference_module()
        if PB_lib is None:
            raise RuntimeError("oracle/_ref missing")
        calls, pts = ref_marshalled_calls(w, int(w["n_points"]))

        def step():
            for c in calls:
                reference_call(PB_lib, c["xs"], c["xo"], c["sem"], c["bp"])
            torch.cuda.synchronize()
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = (time.perf_counter() - t0) / args.steps
        kind, value, cores = "reference", pts / dt, 1
        sample = (f"bounded sample: the first {len(sample_scenes)} of the {n_scenes_total} scenes = {len(calls)} per-class calls / "
                  f"{pts} points per step through the UNMODIFIED reference PB_lib.binary_cluster (oracle/_ref; CUDA kernels driven "
                  f"by one host thread, CPU tensors in/out)")
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


def ref_marshalled_calls(w, max_points):
    """CPU tensors of the first calls of the workload, as the reference wrapper receives them."""
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


# ----------------------------------------------------------------------------------------------------
class DeviceRun:
    """One workload resident in HBM + its output buffers; step() = one pb_binary_cluster_batched call."""

    def __init__(self, ctx, w, dev, r18, m18):
        import torch
        self.ctx, self.w, self.dev, self.r18, self.m18 = ctx, w, dev, r18, m18
        self.n = int(w["n_points"])
        self.keys = ("x", "y", "z", "xo", "yo", "zo", "sem")
        self.d_in = [torch.from_numpy(w[k]).to(dev) for k in self.keys]
        self.seg, self.csc = w["seg_counts"], w["call_seg_counts"]
        self.S = len(self.seg)
        n, S = self.n, self.S
        kcap = max(n // 32, 1024)
        self.d_out = dict(cluster_id=torch.empty(n, dtype=torch.int32, device=dev), cluster_num=torch.empty(S, dtype=torch.int32, device=dev),
                          degree=torch.empty(n, dtype=torch.int32, device=dev), center=torch.empty(3 * kcap, dtype=torch.float32, device=dev),
                          clt_sem=torch.empty(kcap, dtype=torch.int32, device=dev))
        self.stream = torch.cuda.current_stream()

    def step(self):
        return self.ctx.binary_cluster(*self.d_in, self.seg, self.r18, self.m18, 0.05, True, call_seg_counts=self.csc,
                                       stream=self.stream, **self.d_out)

    def host_results(self, out):
        return dict(cluster_id=self.d_out["cluster_id"].cpu().numpy(), degree=self.d_out["degree"].cpu().numpy(),
                    cluster_num=self.d_out["cluster_num"].cpu().numpy(), center=out["center"].cpu().numpy(),
                    clt_sem=out["clt_sem"].cpu().numpy(), call_clusters=np.asarray(out["call_clusters"], np.int64))


def degree_roofline(n, counters, deg_ms_total, ms_step, clocks, world):
    """`roofline` (HBM, as the contract asks) and `roofline_alu` (the bound that explains the time) of k_degree."""
    peak, peak_src = measured_peaks()
    cells = max(1, counters["cells"])
    chunks = max(1, counters.get("chunks", 1))
    # algorithmic bytes of k_degree: per point 16 B pts4 + 4 B row_of read, 4 B zero fill + 4 B degree
    # accumulation (RED); per fine cell 12 B (key, coarse ordinal); per coarse cell 76 B (9 stencil rows + point offset)
    deg_bytes = (28.0 * n + 12.0 * cells + 76.0 * counters.get("coarse_cells", 0)) / chunks   # per launch
    traffic = None
    try:  # dram__bytes_read+write per launch from the newest committed ncu --set full capture of the same workload
        cands = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_r*_k_degree.json")))
        prof = json.load(open(cands[-1]))
        if world == 1 and abs(prof["points_per_launch"] - n / chunks) < 0.02 * n:
            traffic = prof["dram_bytes_per_launch"]
    except Exception:
        pass
    deg_ms = deg_ms_total / chunks
    achieved = deg_bytes / (deg_ms * 1e-3) / 1e9 if deg_ms > 0 else None
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    intra = counters.get("intra_tests", 0)
    sym = intra > 0                                   # the symmetric kernel reports its one-sided share
    ceiling = PAIR_TEST_CEILING_SYM if sym else PAIR_TEST_CEILING
    alu_peak = 148 * 32 * ceiling * sm_mhz * 1e6
    tests_per_s = counters["pair_tests"] / (deg_ms_total * 1e-3) if deg_ms_total > 0 else None
    ordered = 2 * (counters["pair_tests"] - intra) + intra if sym else counters["pair_tests"]
    roof = {"bound": "hbm", "kernel": "k_degree", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
            "traffic_source": "profiles/ncu_r*_k_degree.json (ncu --set full capture of this workload, per launch)" if traffic else None,
            "launches_per_step": chunks, "avg_launch_ms": deg_ms, "share_of_step": deg_ms_total / ms_step if ms_step else None,
            "algorithmic_bytes_per_launch": deg_bytes,
            "note": "k_degree is fp32-pipe bound, not HBM bound: see roofline_alu (SURVEY.md §8d asks for both)"}
    alu = {"kernel": "k_degree", "counting": "symmetric" if sym else "one-sided",
           "pair_tests_per_step": counters["pair_tests"], "achieved": tests_per_s, "peak": alu_peak,
           "unit": "pair tests/s", "frac": (tests_per_s / alu_peak) if tests_per_s else None,
           "pair_tests_per_point": counters["pair_tests"] / max(1, n),
           "ordered_pairs_settled_per_step": ordered,
           "ordered_pairs_per_s": ordered / (deg_ms_total * 1e-3) if deg_ms_total > 0 else None,
           "vs_one_sided_ceiling": (ordered / (deg_ms_total * 1e-3)) / (148 * 32 * PAIR_TEST_CEILING * sm_mhz * 1e6) if deg_ms_total > 0 else None,
           "peak_def": f"148 SM x 32 lanes x {ceiling} warp-tests/clk/SM (builder-measured ceiling of the "
                       f"{'symmetric' if sym else 'one-sided'} packed fp32x2 candidate loop, tools/microbench/pipes.cu, "
                       "profiles/microbench_pipes_r02.txt) x measured SM clock; `vs_one_sided_ceiling` = ordered pairs settled "
                       f"per second / the one-sided ceiling ({PAIR_TEST_CEILING}): above 1 means the symmetric kernel does what no "
                       "one-sided kernel could"}
    return roof, alu


def timed_steps(run, steps, warmup, barrier, max_over_ranks, post_step=None, post_finish=None):
    import torch
    for _ in range(warmup):
        out = run.step()
        if post_step:
            post_step()
    if post_finish:
        post_finish()
    launches = run.ctx.last_launch_count
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    deg_ms = 0.0
    e0.record(run.stream)
    for _ in range(steps):
        out = run.step()
        if post_step:
            post_step()
        deg_ms += run.ctx.stage_ms().get("degree", 0.0)   # the one always-on event pair (per chunk, around k_degree)
    if post_finish:
        post_finish()                                      # the last step's gather completes inside the timed region
    e1.record(run.stream)
    barrier()
    return max_over_ranks(e0.elapsed_time(e1) / steps), out, launches, deg_ms / steps


# ----------------------------------------------------------------------------------------------------
def voxel_bench(d_xo, n, dev, scene_of_point=None):
    """Rows a12-a14, the HBM-bound scatter-gather of the path: preallocated outputs, >= 20 repetitions over a rotation
    of buffer sets larger than L2 (126 MB), min / median per op.  The batch column is the point's scene, as in the
    reference (ME.utils.batched_coordinates, network/PBNet.py:236-247: scenes never share a voxel); without it
    (`scene_of_point=None`) all scenes fall into ONE grid and every voxel row is fetched once per scene that touches it."""
    import torch

    from pbnet_b200 import scenes, voxel
    nv = min(n, 8_000_000)
    coords = torch.stack([d_xo[0][:nv], d_xo[1][:nv], d_xo[2][:nv]], 1).contiguous()
    if scene_of_point is None:
        bcol = torch.zeros(nv, dtype=torch.int32, device=dev)
    else:
        sc = np.ascontiguousarray(scene_of_point[:nv]).astype(np.int32)
        bcol = torch.from_numpy(sc - sc.min()).to(dev)
    peak_v, _ = measured_peaks()
    reps = 24

    def timed_each(f, nrep):
        f(0)
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nrep)]
        for i, (a, b) in enumerate(ev):
            a.record()
            f(i)
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        return ms[0] * 1e-3, ms[len(ms) // 2] * 1e-3

    vm = voxel.voxel_map(coords, scenes.VOXEL_SIZE, batch=bcol)
    t_vox = timed_each(lambda i: voxel.voxel_map(coords, scenes.VOXEL_SIZE, batch=bcol), 8)
    C = 76  # 32 + 20 + 20 + 3 + 1 channels gathered at network/PBNet.py:130-134
    V = vm.n_voxels
    sets = 3  # rotation: 3 x (V*C*4 + nv*C*4) bytes >> L2
    vfeat = [torch.randn((V, C), device=dev) for _ in range(sets)]
    outs = [torch.empty((nv, C), device=dev) for _ in range(sets)]
    gouts = [torch.empty((V, C), device=dev) for _ in range(sets)]
    t_dev = timed_each(lambda i: voxel.devoxelize_raw(vfeat[i % sets], vm.inverse, out=outs[i % sets]), reps)
    t_bwd = timed_each(lambda i: voxel.voxel_rows(outs[i % sets], vm, "sum", out=gouts[i % sets]), reps)
    gb_f = (nv * C * 4 + V * C * 4 + nv * 8) / 1e9
    gb_b = (nv * C * 4 + V * C * 4 + nv * 4 + V * 4) / 1e9
    res = {"points": nv, "voxels": V, "batches": int(bcol.max().item()) + 1, "repetitions": reps,
           "buffers": f"{sets} rotating input/output sets, outputs preallocated",
           "voxelize_points_per_s": {"best": nv / t_vox[0], "median": nv / t_vox[1]},
           "devoxelize": {"channels": C, "ms_min": t_dev[0] * 1e3, "ms_median": t_dev[1] * 1e3, "algorithmic_gb": gb_f,
                          "achieved_gbs_median": gb_f / t_dev[1], "frac_of_hbm_copy_peak_median": gb_f / t_dev[1] / peak_v,
                          "frac_of_hbm_copy_peak_best": gb_f / t_dev[0] / peak_v},
           "devoxelize_backward": {"ms_min": t_bwd[0] * 1e3, "ms_median": t_bwd[1] * 1e3, "algorithmic_gb": gb_b,
                                   "achieved_gbs_median": gb_b / t_bwd[1], "frac_of_hbm_copy_peak_median": gb_b / t_bwd[1] / peak_v,
                                   "frac_of_hbm_copy_peak_best": gb_b / t_bwd[0] / peak_v}}
    return res


def next_rows_bench(sizes, dev):
    """Rows f1/f2/f4 on ONE scene with three rotated copies (the unit eval_map.py processes per iteration)."""
    import torch

    from pbnet_b200 import evalpost, grouping, scenes
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

    def timed(f, reps):
        f()
        f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            r = f()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        return ts[len(ts) // 2], r
    t_grp, _ = timed(lambda: grouping.group_instances(xyz3, off3, sem3, bh3, scenes.RADIUS, scenes.MIN_PTS, cp3), 9)
    t_prop, pr = timed(lambda: grouping.propose(xyz3, off3, sem3, bh3, feat3, sfp3, scenes.RADIUS, scenes.MIN_PTS, cp3), 9)
    scn = pr["scenes"]
    E3, P3 = int(scn["index"].shape[0]), int(scn["offsets"].shape[0]) - 1
    ms3 = torch.rand(E3, device=dev, generator=gen)
    t_gp, gp = timed(lambda: grouping.get_proposal(scn["offsets"], scn["index"], ms3), 9)
    score3 = torch.rand(int(gp[1].shape[0]) - 1, device=dev, generator=gen)
    n3 = xyz3.shape[0] // cp3
    sp3 = torch.unique(torch.floor(xyz3[:n3] / 0.1).to(torch.int64), dim=0, return_inverse=True)[1].contiguous()
    t_ev, ev = timed(lambda: evalpost.postprocess(gp[0], gp[1], score3, sem3, sp3, int(xyz3.shape[0])), 9)
    return {"unit": "one scene x 3 rotated copies (the per-iteration unit of eval_map.py), host wall clock incl. the final sync, "
                    "median of 9", "points": int(xyz3.shape[0]), "proposals": P3, "list_entries": E3,
            "voxels": int(pr["voxel_coords"].shape[0]), "group_instances_ms": t_grp * 1e3, "propose_ms": t_prop * 1e3,
            "propose_api": "grouping.propose = network/PBNet.py:144-247 (class loop, local scenes, feature rows, proposal voxelization)",
            "get_proposal_ms": t_gp * 1e3, "kept_entries": int(gp[0].shape[0]), "eval_postprocess_ms": t_ev * 1e3,
            "final_clusters": int(ev["scores"].shape[0])}


def dropin_bench(w, n_calls):
    """e2e through the reference-facing per-class operator (pbnet_ops.cluster, CPU tensors): the reference's own call
    pattern, one call per (scene, class)."""
    import torch

    from pbnet_b200 import pbnet_ops, scenes, workload
    from pbnet_b200.cluster import default_context
    calls, pts = [], 0
    for c, ps, ss in workload.iter_calls(w):
        if c >= n_calls:
            break
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        calls.append((torch.stack([t(w["x"][ps]), t(w["y"][ps]), t(w["z"][ps])], 1),
                      torch.stack([t(w["xo"][ps]), t(w["yo"][ps]), t(w["zo"][ps])], 1), t(w["sem"][ps]).long(),
                      t(w["seg_counts"][ss])))
        pts += ps.stop - ps.start
    for c in calls[:8]:
        pbnet_ops.cluster(c[0], c[1], c[2], c[3], scenes.RADIUS, scenes.MIN_PTS, len(c[3]))
    dts, launches = [], 0
    for _ in range(3):  # host-side latency measurement: median of three passes
        t0 = time.perf_counter()
        for c in calls:
            pbnet_ops.cluster(c[0], c[1], c[2], c[3], scenes.RADIUS, scenes.MIN_PTS, len(c[3]))
        dts.append(time.perf_counter() - t0)
    launches = default_context(torch.cuda.current_device()).last_launch_count
    dt = sorted(dts)[1]
    return {"value": pts / dt, "unit": UNIT, "calls": len(calls), "points": pts,
            "api": "pbnet_b200.pbnet_ops.cluster per (scene, class), CPU tensors in/out (reference call pattern)",
            "us_per_call": 1e6 * dt / max(1, len(calls)), "passes_us_per_call": [round(1e6 * d / max(1, len(calls))) for d in dts],
            "launches_last_call": int(launches)}


# ----------------------------------------------------------------------------------------------------
def stress_workloads(args):
    """BASELINE.json configs[3] (large scene, radius sweep) and [4] (dense HP-fraction sweep) as (label, workload, r18, m18)."""
    from pbnet_b200 import scenes
    out = []
    if args.workload == "c3":
        sc = scenes.make_scene(3003, args.c3_points, hp_frac=0.25)
        calls = scenes.class_calls(sc, 1)
        xs = np.concatenate([c["xyz_shift"] for c in calls])
        xo = np.concatenate([c["xyz_orig"] for c in calls])
        w = dict(x=np.ascontiguousarray(xs[:, 0]), y=np.ascontiguousarray(xs[:, 1]), z=np.ascontiguousarray(xs[:, 2]),
                 xo=np.ascontiguousarray(xo[:, 0]), yo=np.ascontiguousarray(xo[:, 1]), zo=np.ascontiguousarray(xo[:, 2]),
                 sem=np.concatenate([c["sem"] for c in calls]).astype(np.int32),
                 seg_counts=np.concatenate([c["seg_counts"] for c in calls]).astype(np.int32),
                 call_seg_counts=np.ones(len(calls), np.int32), n_points=np.int64(len(xs)))
        for r in (0.02, 0.03, 0.04, 0.05, 0.06):
            out.append((f"r={r}", w, np.full(18, np.float32(r), np.float32), np.full(18, 31, np.int32)))
    else:
        for f in (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9):
            xs, xo, sem = scenes.make_dense_case(4004, args.c4_points, f)
            w = dict(x=np.ascontiguousarray(xs[:, 0]), y=np.ascontiguousarray(xs[:, 1]), z=np.ascontiguousarray(xs[:, 2]),
                     xo=np.ascontiguousarray(xo[:, 0]), yo=np.ascontiguousarray(xo[:, 1]), zo=np.ascontiguousarray(xo[:, 2]),
                     sem=sem.astype(np.int32), seg_counts=np.array([len(sem)], np.int32), call_seg_counts=np.ones(1, np.int32),
                     n_points=np.int64(len(sem)))
            out.append((f"hp_fraction={f}", w, np.full(18, np.float32(0.04), np.float32), np.full(18, 31, np.int32)))
    return out


def run_stress(args):
    """One JSON line for --workload c3 / c4: a sweep list, each point with throughput, roofline, roofline_alu, verify."""
    import torch

    from pbnet_b200.cluster import Context
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    ctx = Context(0)
    sampler = ClockSampler(0)
    sampler.start()
    threads = os.cpu_count() or 1
    sweep = []
    ident = lambda v: v
    for label, w, r18, m18 in stress_workloads(args):
        run = DeviceRun(ctx, w, dev, r18, m18)
        ms_step, out, launches, deg_ms = timed_steps(run, args.steps, args.warmup, torch.cuda.synchronize, ident)
        ctx.set_profiling(True)
        out = run.step()
        counters, stage = ctx.counters(), ctx.stage_ms()
        ctx.set_profiling(False)
        host = run.host_results(out)
        ver, cpu_rate = verify_against_oracle(w, host, r18, m18, threads, time_budget_s=args.verify_budget)
        roof, alu = degree_roofline(run.n, counters, deg_ms, ms_step, None, 1)
        sweep.append({"point": label, "points": run.n, "value": run.n / (ms_step * 1e-3), "ms_per_step": ms_step,
                      "clusters": int(out["n_clusters"]), "launches_per_step": int(launches), "roofline": roof, "roofline_alu": alu,
                      "sum_degree_per_point": counters["sum_deg"] / max(1, run.n), "hp_points": counters["n_hp"],
                      "centre_replay": {"halves": counters.get("centre_halves", 0),
                                        "replayed_with_div_rn": counters.get("centre_halves_replayed", 0),
                                        "replay_cycles": counters.get("centre_replay_cycles", 0),
                                        "gather_cycles": counters.get("centre_gather_cycles", 0)},
                      "stage_ms": {k: round(v, 3) for k, v in stage.items()}, "verify": ver, "cpu_oracle_points_per_s": cpu_rate})
        del run
        torch.cuda.empty_cache()
    clocks = sampler.stop()
    name = ("C3: one synthetic room scene of %d points (seed 3003, hp_frac 0.25), per-class calls, radius sweep 0.02-0.06, min_pts=31"
            % args.c3_points) if args.workload == "c3" else (
        "C4: dense single-class segment of %d points (seed 4004), HP-fraction sweep 0.1-0.9, blobs of 2001 points with sigma=r/4, "
        "r=0.04, min_pts=31" % args.c4_points)
    best = max(sweep, key=lambda s: s["value"])
    line = {"metric": METRIC, "value": float(np.mean([s["value"] for s in sweep])), "unit": UNIT, "n_gpus": 1, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": float(np.mean([s["ms_per_step"] for s in sweep])), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": name, "value_def": "mean over the sweep points; per-point figures in `sweep`",
                       "l2": "inputs + workspace of one step exceed the 126 MB L2 (>= 400 B/point); no flush needed"},
            "sweep": sweep, "roofline": best["roofline"], "roofline_alu": best["roofline_alu"], "clocks": clocks,
            "gpu_launches": int(sum(s["launches_per_step"] for s in sweep) * args.steps)}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c1", choices=["c1", "c3", "c4"])
    ap.add_argument("--scenes", type=int, default=312)
    ap.add_argument("--copies", type=int, default=1)
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="strong (default): BASELINE.json configs[2] taken literally, the SAME --scenes scenes sharded over the GPUs; "
                         "weak: --scenes scenes PER GPU (a set of scenes x N_gpus scenes sharded by scene)")
    ap.add_argument("--ref-scenes", type=int, default=8, help="scenes per step of the reference arm (bounded sample)")
    ap.add_argument("--dropin-calls", type=int, default=256)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle legs (verify + cpu_baseline)")
    ap.add_argument("--no-extras", action="store_true", help="skip the voxel / next-rows / drop-in legs")
    ap.add_argument("--verify-budget", type=float, default=None, help="seconds of oracle time per verification (default: all calls)")
    ap.add_argument("--c3-points", type=int, default=2_000_000)
    ap.add_argument("--c4-points", type=int, default=1_000_000)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    if args.workload != "c1":
        return run_stress(args)

    rank, local_rank, world = dist_env()
    from pbnet_b200 import scenes, workload
    # scenes are independent units: they are partitioned over the ranks (LPT by point count), no data-path collective.
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
    from pbnet_b200 import sharding
    # NUMA: this rank's pinned buffers and copy submissions stay on the socket its GPU hangs off (undone before the CPU legs)
    numa = None if os.environ.get("PB_NUMA_BIND", "1") == "0" else sharding.bind_to_gpu_numa_node(local_rank)
    from pbnet_b200.cluster import Context
    ctx = Context(local_rank)          # production configuration: no stage events, no counters in the timed region
    r18 = np.full(18, np.float32(scenes.RADIUS), np.float32)
    m18 = np.full(18, scenes.MIN_PTS, np.int32)
    run = DeviceRun(ctx, w, dev, r18, m18)
    n, S, seg, csc, stream = run.n, run.S, run.seg, run.csc, run.stream

    # gather of proposals to rank 0 (the only collective; NCCL over NVLink) — pbnet_b200/sharding.py
    # ids restart at 0 in every call and a call has a handful of clusters: int16 on the wire; the transfer of step i
    # overlaps the kernels of step i+1, the last one is awaited (finish) before the timed region ends
    gather_ids = sharding.Rank0Gather(n, torch.int32, dev, narrow_to=torch.int16, overlap=True)

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
    post = (lambda: gather_ids(run.d_out["cluster_id"])) if world > 1 else None
    ms_step, out, launches_per_step, deg_ms_total = timed_steps(run, args.steps, args.warmup, barrier, max_over_ranks, post,
                                                                (lambda: gather_ids.finish()) if world > 1 else None)
    if world > 1 and int(np.max(out["call_clusters"], initial=0)) >= 32768:
        raise SystemExit("a call produced >= 32768 clusters: the int16 gather of the ids would be lossy")
    n_clusters = out["n_clusters"]
    host = run.host_results(out)       # the results of the LAST timed step: what `verify` checks below

    # one extra, untimed step with stage events and counters (pair tests, cells, ...)
    ctx.set_profiling(True)
    run.step()
    counters, stage_ms = ctx.counters(), ctx.stage_ms()
    ctx.set_profiling(False)

    # ---- e2e: same call with HOST (pinned) buffers; H2D of inputs and D2H of results inside the timed region
    kcap = max(n // 32, 1024)
    h_in = [torch.from_numpy(w[k]).pin_memory() for k in run.keys]
    h_out = dict(cluster_id=torch.empty(n, dtype=torch.int32).pin_memory(), cluster_num=torch.empty(S, dtype=torch.int32).pin_memory(),
                 degree=torch.empty(n, dtype=torch.int32).pin_memory(), center=torch.empty(3 * kcap, dtype=torch.float32).pin_memory(),
                 clt_sem=torch.empty(kcap, dtype=torch.int32).pin_memory())

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
    if numa is not None:
        os.sched_setaffinity(0, numa[2])  # the verification / CPU-baseline legs use every host core
    clocks = sampler.stop() if rank == 0 else None  # sampled from the first warm-up step to the end of the e2e loop
    h2d = 28 * n
    d2h = 8 * n + 4 * S + 16 * int(oh["n_clusters"])
    e2e_identical = all(np.array_equal(h_out[k].numpy()[:len(host[k])], host[k]) for k in ("cluster_id", "degree", "cluster_num"))

    # ---- verify (outside every timed region): ALL calls of this rank's shard vs the CPU oracle, bit for bit; the oracle's
    #      throughput over the full set on all host cores doubles as the cpu_baseline figure (N=1)
    verify, cpu = None, None
    if not args.no_cpu_baseline:
        threads = max(1, (os.cpu_count() or 1) // world)
        ver, cpu_rate = verify_against_oracle(w, host, r18, m18, threads, args.verify_budget)
        ver["e2e_host_call_identical_to_device_call"] = bool(e2e_identical)
        agg = {k: int(sum_over_ranks(float(ver[k]))) for k in ("calls", "calls_total", "points", "mismatches", "skipped_for_time")}
        if rank == 0:
            ver.update(agg)
            ver["ranks"] = world
            scene_ids = sorted(set(w["call_scene"].tolist()))     # calls are stored scene by scene, ascending
            last_scene = scene_ids[min(args.ref_scenes, len(scene_ids)) - 1]
            ref_calls = int(np.sum(w["call_scene"] <= last_scene))
            ver["reference_live"] = verify_against_reference(w, host, ref_calls)
            verify = ver
            if world == 1:
                cpu = {"value": cpu_rate, "unit": UNIT, "cores": threads, "kind": "port",
                       "sample": f"ALL {ver['calls']} per-class calls ({ver['points']} points) of the workload through oracle/pb_oracle.c "
                                 f"(grid-accelerated faithful restatement), {threads} threads over calls, {ver['seconds']} s — the same "
                                 f"pass that produced `verify`"}

    dropin = vox = nxt = None
    if rank == 0 and not args.no_extras:
        dropin = dropin_bench(w, args.dropin_calls)
        vox = voxel_bench(run.d_in[3:6], n, dev, np.repeat(w["call_scene"], w["call_points"]))
        nxt = next_rows_bench(sizes, dev)

    if rank == 0:
        peak, _ = measured_peaks()
        roof, alu = degree_roofline(n, counters, deg_ms_total, ms_step, clocks, world)
        value = total_points / (ms_step * 1e-3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": c1_config_dict(args, world, n_scenes_total),
            "workload_detail": {"points_total": total_points, "points_rank0": n, "calls_rank0": int(len(csc)), "segments_rank0": S,
                                "clusters_rank0": int(n_clusters),
                                "working_set": "inputs + workspace of one step are ~%.1f GB per rank" % ((28 + 430) * n / 1e9),
                                "collective": "NCCL gather of the cluster ids (int16 on the wire) to rank 0 once per step, issued on a "
                                              "side stream so that it overlaps the next step; the last one is awaited inside the timed "
                                              "region" if world > 1 else "none"},
            "e2e": {"value": total_points / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "pb_binary_cluster_batched via pbnet_b200.cluster.Context.binary_cluster, pinned host buffers "
                           "(one batched call for the whole shard; the per-class reference call pattern is `e2e_dropin`)",
                    "ms_per_step": e2e_s * 1e3},
            "e2e_dropin": dropin,
            "verify": verify,
            "gpu_launches": int(launches_per_step * args.steps),
            "launches_per_step": int(launches_per_step),
            "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()},
            "stage_ms_note": "from ONE extra untimed step with stage events on (the timed region runs the production configuration: "
                             "only the event pair around k_degree); intervals are summed over the chunks of a step, the chunks run on "
                             "two concurrent streams, so an interval also contains time shared with the other chunk",
            "roofline": roof, "roofline_alu": alu,
            "io_roofline": {"bytes_per_point": 36, "achieved_gbs": value * 36 / 1e9, "frac_of_hbm": value * 36 / 1e9 / peak},
            "counters": counters,
            "voxel": vox,
            "next_rows": nxt,
            "cpu_baseline": cpu,
            "clocks": clocks,
            "numa": ({"node": numa[0], "cpus": numa[1]} if numa else None),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
