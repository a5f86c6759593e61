"""A/B run of k_degree variants on device-resident inputs (one process, one workload):
  python tools/ab_degree.py --scenes 312 [--configs "PB_DEG_SYM=0;PB_DEG_SYM=1,PB_DEG_MINB_SYM=8;..."]
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
ap.add_argument("--configs", default="PB_DEG_SYM=0;PB_DEG_SYM=1,PB_DEG_MINB_SYM=8;PB_DEG_SYM=1,PB_DEG_MINB_SYM=9")
args = ap.parse_args()
sizes = scenes.scene_sizes(312)
w = workload.build(range(args.scenes), sizes, 1)
import torch  # noqa: E402

from pbnet_b200.cluster import Context  # noqa: E402

dev = torch.device("cuda", 0)
d_in = [torch.from_numpy(w[k]).to(dev) for k in ("x", "y", "z", "xo", "yo", "zo", "sem")]
r18 = np.full(18, np.float32(0.04), np.float32)
m18 = np.full(18, 31, np.int32)
ref = None
for cfg in args.configs.split(";"):
    kv = dict(x.split("=") for x in cfg.split(",") if x)
    for k, v in kv.items():
        os.environ[k] = v
    for prof in (False, True):
        ctx = Context(0, profiling=prof)
        if os.environ.get("PB_CHUNK_POINTS"):
            ctx.set_chunk_points(int(os.environ["PB_CHUNK_POINTS"]))
        ts = []
        for i in range(2 + (args.steps if not prof else 1)):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = ctx.binary_cluster(*d_in, w["seg_counts"], r18, m18, 0.05, True, call_seg_counts=w["call_seg_counts"])
            torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
        if not prof:
            ids = out["cluster_id"].cpu().numpy()
            deg = out["degree"].cpu().numpy() if "degree" in out else None
            if ref is None:
                ref = (ids, deg)
            same = bool((ids == ref[0]).all()) and (deg is None or bool((deg == ref[1]).all()))
            print(f"{cfg}: step ms {[round(t, 3) for t in ts[2:]]} median {np.median(ts[2:]):.3f}  identical_to_first={same}", flush=True)
        else:
            st = ctx.stage_ms()
            print("   stages", {k: round(v, 2) for k, v in st.items() if k in ("degree", "hp_cells", "union", "sort", "lp_nn")}, ctx.counters().get("pair_tests"), flush=True)
        del ctx
    for k in kv:
        os.environ.pop(k, None)
