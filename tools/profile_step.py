"""Runs a few batched grouping steps on device-resident inputs — the command wrapped by ncu
(B200_PROFILING.md): e.g.
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
      python tools/profile_step.py --scenes 64 --steps 2
  ncu --set full --clock-control none --import-source on -k regex:k_degree -s 1 -c 1 -o gpurun_out/prof_degree \
      python tools/profile_step.py --scenes 64 --steps 2
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import scenes, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=64)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--copies", type=int, default=1)
args = ap.parse_args()
sizes = scenes.scene_sizes(312)
w = workload.build(range(args.scenes), sizes, args.copies)
import torch  # noqa: E402

from pbnet_b200.cluster import Context  # noqa: E402

dev = torch.device("cuda", 0)
ctx = Context(0, profiling=True)
if os.environ.get("PB_CHUNK_POINTS"):
    ctx.set_chunk_points(int(os.environ["PB_CHUNK_POINTS"]))
d_in = [torch.from_numpy(w[k]).to(dev) for k in ("x", "y", "z", "xo", "yo", "zo", "sem")]
r18 = np.full(18, np.float32(0.04), np.float32)
m18 = np.full(18, 31, np.int32)
import time  # noqa: E402
for i in range(args.steps):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = ctx.binary_cluster(*d_in, w["seg_counts"], r18, m18, 0.05, True, call_seg_counts=w["call_seg_counts"])
    torch.cuda.synchronize()
    t1 = time.perf_counter()
print("last step wall ms", round((t1 - t0) * 1e3, 3))
print("points", int(w["n_points"]), "clusters", out["n_clusters"], "launches", ctx.last_launch_count)
print({k: round(v, 3) for k, v in ctx.stage_ms().items()})
print(ctx.counters())
