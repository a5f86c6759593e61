"""A/B run of the host-buffer (end-to-end) call: pinned inputs and outputs, H2D + D2H inside the timed region.
  python tools/ab_e2e.py --scenes 312 --configs "PB_HOST_SPLIT=150,425,425;PB_HOST_SPLIT=100,300,600"
Environment variables are read by pb_create, so every configuration gets its own context."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import scenes, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=312)
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--configs", default="PB_HOST_SPLIT=150,425,425")
args = ap.parse_args()
sizes = scenes.scene_sizes(312)
w = workload.build(range(args.scenes), sizes, 1)
import torch  # noqa: E402

from pbnet_b200.cluster import Context  # noqa: E402

n, S = int(w["n_points"]), len(w["seg_counts"])
kcap = max(n // 32, 1024)
h_in = [torch.from_numpy(w[k]).pin_memory() for k in ("x", "y", "z", "xo", "yo", "zo", "sem")]
h_out = dict(cluster_id=torch.empty(n, dtype=torch.int32).pin_memory(), cluster_num=torch.empty(S, dtype=torch.int32).pin_memory(),
             degree=torch.empty(n, dtype=torch.int32).pin_memory(), center=torch.empty(3 * kcap, dtype=torch.float32).pin_memory(),
             clt_sem=torch.empty(kcap, dtype=torch.int32).pin_memory())
r18 = np.full(18, np.float32(0.04), np.float32)
m18 = np.full(18, 31, np.int32)
stream = torch.cuda.Stream()
ref = None
for cfg in args.configs.split(";"):
    kv = dict(x.split("=") for x in cfg.split("+") if x)
    for k, v in kv.items():
        os.environ[k] = v
    ctx = Context(0)
    ts = []
    for i in range(2 + args.steps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ctx.binary_cluster(*h_in, w["seg_counts"], r18, m18, 0.05, True, call_seg_counts=w["call_seg_counts"], stream=stream, **h_out)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ids = out["cluster_id"].numpy().copy()
    if ref is None:
        ref = ids
    print(f"{cfg}: e2e ms {[round(t, 2) for t in ts[2:]]} median {np.median(ts[2:]):.2f} chunks {ctx.counters()['chunks']} identical={bool((ids == ref).all())}", flush=True)
    del ctx
    for k in kv:
        os.environ.pop(k, None)
