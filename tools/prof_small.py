"""Phase times of the small-call kernel (globaltimer stamps) for the per-class calls of one scene."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import scenes  # noqa: E402
from pbnet_b200.cluster import Context  # noqa: E402

sc = scenes.make_scene(22, 150000)
ctx = Context(0)
r18 = np.full(18, np.float32(0.04), np.float32)
m18 = np.full(18, 31, np.int32)
for c in scenes.class_calls(sc, 1):
    xs, xo = c["xyz_shift"], c["xyz_orig"]
    cols = [np.ascontiguousarray(xs[:, i]) for i in range(3)] + [np.ascontiguousarray(xo[:, i]) for i in range(3)]
    sem = c["sem"].astype(np.int32)
    for _ in range(3):
        out = ctx.binary_cluster(*cols, sem, c["seg_counts"], r18, m18)
    t0 = time.perf_counter()
    for _ in range(10):
        out = ctx.binary_cluster(*cols, sem, c["seg_counts"], r18, m18)
    dt = (time.perf_counter() - t0) / 10 * 1e6
    st = {k: round(v * 1e3, 1) for k, v in ctx.stage_ms().items() if v > 0}
    print(f"n={len(sem):6d} K={out['n_clusters']:3d} total {dt:6.0f} us  kernel phases(us): {st}  sum {sum(st.values()):.0f}")
