"""Stage timeline of one batched step (PB_TIMELINE=1 makes the library print, per chunk, when every stage started):
  python tools/timeline.py [--host] [--scenes 312]        # --host: pinned host buffers (the e2e path)"""
import argparse
import os
import sys

import numpy as np

os.environ["PB_TIMELINE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pbnet_b200 import scenes, workload  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--scenes", type=int, default=312)
ap.add_argument("--host", action="store_true")
args = ap.parse_args()
w = workload.build(range(args.scenes), scenes.scene_sizes(312), 1)
import torch  # noqa: E402

from pbnet_b200.cluster import Context  # noqa: E402

keys = ("x", "y", "z", "xo", "yo", "zo", "sem")
if args.host:
    ins = [torch.from_numpy(w[k]).pin_memory() for k in keys]
else:
    ins = [torch.from_numpy(w[k]).cuda() for k in keys]
r18 = np.full(18, np.float32(0.04), np.float32)
m18 = np.full(18, 31, np.int32)
ctx = Context(0, profiling=False)
for i in range(3):
    if i == 2:
        ctx.set_profiling(True)
        print("--- timeline of step 3 (ms since the first event of chunk 0)", file=sys.stderr)
    ctx.binary_cluster(*ins, w["seg_counts"], r18, m18, 0.05, True, call_seg_counts=w["call_seg_counts"])
    torch.cuda.synchronize()
